"""Run the forward hex-conv problems of one fused ico2ico step once each (ncu counts their DRAM traffic):

    ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum -k regex:patch_conv --csv --log-file t.csv python tools/measure_traffic.py
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import bench                                              # noqa: E402
from geniconet_b200 import _lib, models as gm            # noqa: E402
from geniconet_b200.ico_conv import get_plan             # noqa: E402

L = _lib.lib
B = 36
model = gm.ico2ico(gm.default_params('ico2ico', 5))
for name, ci, co, stride, lvl, cm, count in bench.conv_calls(model, True):
    Pin = 10 * 4 ** lvl
    Pout = Pin // stride ** 2
    plan = get_plan(_lib.PLAN_HEXCONV, lvl, stride, cm, 'cuda')
    w = torch.randn(co, ci, 7, device='cuda') * 0.05
    bias = torch.zeros(co, device='cuda')
    packed = torch.empty(L.gin_hexconv_packed_bytes(ci, co), dtype=torch.uint8, device='cuda')
    st = torch.cuda.current_stream().cuda_stream
    _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), ci, co, st))
    xb = (torch.randn(B * Pin + 2 * B, ci, device='cuda') * 0.5).to(_lib.forward_operand_dtype())
    y = torch.empty(B * Pout, co, device='cuda')
    flush = torch.empty(64 << 20, dtype=torch.float32, device='cuda')
    for _ in range(count):
        flush.zero_()                                    # 256 MB write: nothing of the operand is left in the 126 MB L2
        _lib.check(L.gin_hexconv_fwd_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), bias.data_ptr(), y.data_ptr(), B, ci, co, st))
    torch.cuda.synchronize()
    print('%s %d->%d s%d L%d x%d' % (name, ci, co, stride, lvl, count))
