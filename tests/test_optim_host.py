"""Host side of geniconet_b200.optim.Adam without a GPU: the table row it uploads is the header's GinAdamTensor, the chunk prefix
is what gin_adam_step's binary search expects, and the constructor mirrors torch.optim.Adam's argument checks and group keys."""
import os
import re

import numpy as np
import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_table_row_matches_the_header_struct():
    from geniconet_b200 import optim
    hdr = open(os.path.join(ROOT, 'include', 'geniconet_b200.h')).read()
    body = re.search(r'typedef struct \{(.*?)\} GinAdamTensor;', hdr, re.S).group(1)
    fields = re.findall(r'(?:const\s+)?(?:float\*|int64_t)\s+(\w+);', body)
    assert fields == list(optim._ROW.names) == ['p', 'g', 'm', 'v', 'step', 'n']
    assert optim._ROW.itemsize == 48 and all(optim._ROW.fields[f][1] == 8 * i for i, f in enumerate(fields))


def test_chunk_prefix_and_rows():
    from geniconet_b200 import optim
    from geniconet_b200 import _lib
    chunk = int(_lib.lib.gin_adam_chunk())                      # a host-side constant, no device needed
    assert chunk == 4096
    o = optim.Adam.__new__(optim.Adam)
    o._chunk = chunk
    rows = [(0x1000, 0x2000, 0x3000, 0x4000, 0x5000, 7), (0x1100, 0x2100, 0x3100, 0x4100, 0x5100, chunk),
            (0x1200, 0x2200, 0x3200, 0x4200, 0x5200, chunk + 1), (0x1300, 0x2300, 0x3300, 0x4300, 0x5300, 10 * chunk)]
    n_params = 6                                                 # the group holds more parameters than have gradients this step
    tb = {'rows_bytes': n_params * optim._ROW.itemsize}
    host = torch.zeros(n_params * optim._ROW.itemsize + (n_params + 1) * 4, dtype=torch.uint8)
    count, chunks = o._fill(tb, host, rows)
    assert (count, chunks) == (4, 1 + 1 + 2 + 10)
    raw = host.numpy()
    got = raw[:count * 48].view(optim._ROW)
    assert [tuple(int(v) for v in r) for r in got] == rows
    first = raw[tb['rows_bytes']:tb['rows_bytes'] + (count + 1) * 4].view(np.int32)
    assert first.tolist() == [0, 1, 2, 4, 14]
    # the kernel's search: the last tensor whose first chunk is <= the block index
    for blk, want in [(0, 0), (1, 1), (2, 2), (3, 2), (4, 3), (13, 3)]:
        assert int(np.searchsorted(first[:count], blk, side='right')) - 1 == want


def test_constructor_mirrors_torch_adam():
    from geniconet_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(3))
    ref = torch.optim.Adam([p]).param_groups[0]
    ours = Adam([p]).param_groups[0]
    assert set(ref) == set(ours)
    for k in ('lr', 'betas', 'eps', 'weight_decay', 'amsgrad', 'maximize', 'differentiable', 'decoupled_weight_decay'):
        assert ref[k] == ours[k], k
    for bad in (dict(lr=-1.0), dict(eps=-1.0), dict(betas=(1.0, 0.9)), dict(betas=(0.9, 1.0)), dict(weight_decay=-0.1)):
        with pytest.raises(ValueError):
            Adam([p], **bad)
        with pytest.raises(ValueError):
            torch.optim.Adam([p], **bad)


def test_cpu_parameters_raise():
    from geniconet_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(4))
    p.grad = torch.ones(4)
    with pytest.raises(RuntimeError, match='no CPU path'):
        Adam([p]).step()
    q = torch.nn.Parameter(torch.zeros(4))
    Adam([q]).step()                                             # nothing has a gradient: nothing to do, nothing raised


def test_reference_scheduler_accepts_the_optimizer():
    """run.py:447-449 wraps the optimizer in CyclicLR(optimizer, lr_base, lr_max, cycle_momentum=False) and steps it after every
    batch (run.py:252-254): the scheduler must construct over our optimizer and drive its group learning rate."""
    from geniconet_b200.optim import Adam
    p = torch.nn.Parameter(torch.zeros(3))
    opt = Adam([p], lr=1e-4)
    sched = torch.optim.lr_scheduler.CyclicLR(opt, 1e-5, 1e-3, cycle_momentum=False)
    ref_opt = torch.optim.Adam([torch.nn.Parameter(torch.zeros(3))], lr=1e-4)
    ref_sched = torch.optim.lr_scheduler.CyclicLR(ref_opt, 1e-5, 1e-3, cycle_momentum=False)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter('ignore')                         # "scheduler.step() before optimizer.step()": no step is taken here
        for _ in range(5):
            sched.step()
            ref_sched.step()
            assert opt.param_groups[0]['lr'] == ref_opt.param_groups[0]['lr']
