"""The reference's data / checkpoint / evaluation-helper surface (data.py, run.py:317-409, ico_utils.py), host side: file
listing order, the .npz contract, the Dataset return conventions, checkpoint round trips and partial loads, and the
point-to-mesh oracle against hand-computed cases."""
import os

import numpy as np
import pytest
import torch

from geniconet_b200 import checkpoint as ck
from geniconet_b200 import data as gd
from geniconet_b200 import ico_utils as iu
from geniconet_b200 import models as gm
from oracle.kaolin_ref import point_to_mesh_distance as p2m_ref

LEVEL = 2
N = 2 ** LEVEL
P = 10 * 4 ** LEVEL


def _params(root, process='train', lvl=2):
    return {'process_name': process, 'model_name': 'ico2ico', 'ico2ico': {'data_instance': 'val'},
            'ico': {'dataPth': os.path.join(root, 'ico'), 'ext': '.npz', 'width': 2 * N, 'dataPthLvl': lvl, 'subdivisions': LEVEL},
            'enc': {'dataPth': os.path.join(root, 'enc'), 'ext': '.npz'}, 'out': {'dataPth': os.path.join(root, 'out_E0')},
            'ftr': {}, 'logDir': os.path.join(root, 'log')}


def _make_modelnet(root, classes=('chair', 'bed'), counts=(12, 3)):
    k = 0
    for cls in classes:
        for split, cnt in zip(('train', 'test'), counts):
            d = os.path.join(root, 'ico', cls, split)
            os.makedirs(d)
            for i in range(1, cnt + 1):
                _, tgt = gd.synthetic_mesh(LEVEL, k)
                gd.write_ico_npz(os.path.join(d, '%s_%d.npz' % (cls, i)), tgt.numpy())
                k += 1
            open(os.path.join(d, 'notes.txt'), 'w').close()


def test_natural_order_and_modelnet_listing(tmp_path):
    assert sorted(['a_10.npz', 'a_2.npz', 'a_1.npz', 'b_1.npz'], key=gd.natural_key) == ['a_1.npz', 'a_2.npz', 'a_10.npz', 'b_1.npz']
    _make_modelnet(str(tmp_path))
    params = _params(str(tmp_path))
    trn = gd.listFiles(params, 'ico', 'trn')              # 'trn' -> 'train', 'val' -> 'test' (data.py:26-29)
    val = gd.listFiles(params, 'ico', 'val')
    assert len(trn) == 24 and len(val) == 6 and all(f.endswith('.npz') for f in trn + val)
    chair = [os.path.basename(f) for f in trn if os.sep + 'chair' + os.sep in f]
    assert chair == ['chair_%d.npz' % i for i in range(1, 13)]            # chair_2 before chair_10


def test_npz_contract_and_train_dataset(tmp_path):
    _make_modelnet(str(tmp_path), classes=('chair',), counts=(3, 1))
    params = _params(str(tmp_path))
    ds = gd.createico2icoDataset(params, 'trn')
    assert len(ds) == 3
    ico, tgt = ds[1]
    assert ico.shape == (3, 5 * N, 2 * N) and tgt.shape == (9, P + 2) and ico.dtype == np.float32
    x_ref, t_ref = gd.synthetic_mesh(LEVEL, 1)
    assert np.array_equal(tgt, t_ref.numpy()) and np.array_equal(ico, x_ref.numpy())     # data.py:66-69 == our synthetic contract
    xb, tb = next(iter(torch.utils.data.DataLoader(ds, batch_size=2)))
    assert xb.shape == (2, 3, 5 * N, 2 * N) and tb.shape == (2, 9, P + 2)
    # DevicePrefetcher degrades to a pass-through on the CPU (the CUDA path is tested on the GPU box)
    got = list(gd.DevicePrefetcher(torch.utils.data.DataLoader(ds, batch_size=2), device='cpu'))
    assert len(got) == 2 and got[0][0].shape == (2, 3, 5 * N, 2 * N) and got[1][0].shape[0] == 1


def test_test_mode_and_encoding_datasets(tmp_path):
    _make_modelnet(str(tmp_path), classes=('chair',), counts=(2, 2))
    params = _params(str(tmp_path), process='test')
    ds = gd.createico2ico_vaeDataset(params, 'val')
    ico, out_stem, ico2 = ds[0]
    assert out_stem == os.path.join(params['out']['dataPth'], 'val', 'chair_1') and ico is ico2 and os.path.isdir(os.path.dirname(out_stem))
    enc_ds = gd.createico2encDataset(params, 'val')
    ico, enc_path = enc_ds[1]
    assert enc_path == os.path.join(params['enc']['dataPth'], 'val', 'chair_2.npz')
    np.savez_compressed(enc_path, np.arange(6, dtype=np.float32).reshape(2, 3))      # what the reference's save_to_file writes ('arr_0')
    assert torch.equal(gd.loadEncFile(params, enc_path), torch.arange(6, dtype=torch.float32).reshape(2, 3))
    with pytest.raises(ValueError):
        gd.loadEncFile(params, enc_path[:-4] + '.bin')
    flat = dict(params, ico=dict(params['ico'], dataPthLvl=1, dataPth=os.path.join(str(tmp_path), 'ico', 'chair', 'test')))
    dec_ds = gd.createenc2icoDataset(flat, 'val')          # encodings matched to ico files by base name (data.py:130-138)
    assert len(dec_ds) == 1
    enc, ico_path, ico = dec_ds[0]
    assert enc.shape == (2, 3) and ico_path.endswith(os.path.join('val', 'chair_2')) and ico.shape == (3, 5 * N, 2 * N)
    with pytest.raises(ValueError):
        gd.loadIcoFile(dict(params, ico=dict(params['ico'], ext='.obj')), 'x.obj')


def _tiny_model():
    p = gm.default_params('ico2ico')
    return gm.ico2ico(p), p


def test_checkpoint_round_trip_and_best_rotation(tmp_path):
    params = _params(str(tmp_path))
    torch.manual_seed(0)
    model, mp = _tiny_model()
    opt = torch.optim.Adam(model.parameters(), lr=1e-3)
    assert ck.loadModel(params, model, [0], 'ico2ico') is False                         # nothing saved yet
    assert ck.saveModel(params, model, opt, 3, 'ico2ico', 0.5, {'note': 'x'}) is True
    assert ck.saveModel(params, model, opt, 3, 'ico2ico', 0.1, None) is False            # never overwrites (run.py:335,340)
    raw = torch.load(os.path.join(params['logDir'], 'savedModel', 'ico2ico_E3.pt'), weights_only=False)
    assert set(raw) == {'model_state_dict', 'optimizer_state_dict', 'epoch', 'loss', 'misc'} and raw['epoch'] == 3 and raw['loss'] == 0.5
    best, last = [np.inf], [1.0]
    for epoch in range(1, 9):                                                           # eight improving epochs -> at most six best files kept
        last[0] = 1.0 / epoch
        ck.saveBestModel(params, model, opt, epoch, 'ico2ico', best, last)
    kept = sorted(os.listdir(os.path.join(params['logDir'], 'savedModel')), key=gd.natural_key)
    assert kept == ['ico2ico_E3.pt'] + ['ico2ico_EB%d.pt' % e for e in range(3, 9)]
    last[0] = 5.0
    ck.saveBestModel(params, model, opt, 9, 'ico2ico', best, last)                      # worse: nothing happens
    assert best[0] == 1.0 / 8 and not os.path.exists(os.path.join(params['logDir'], 'savedModel', 'ico2ico_EB9.pt'))

    torch.manual_seed(1)
    other, _ = _tiny_model()
    opt2 = torch.optim.Adam(other.parameters(), lr=1e-3)
    epoch, loss, misc = [0], [np.inf], []
    assert ck.loadModel(params, other, epoch, 'ico2ico', opt2, loss, misc) is True      # [0] -> newest best (natural order: EB8, not EB3)
    assert epoch == [8] and loss == [1.0 / 8] and misc == [None]
    assert params['out']['dataPth'].endswith('out_EB8')                                 # run.py:377-380
    for (k, a), (_, b) in zip(model.state_dict().items(), other.state_dict().items()):
        assert torch.equal(a, b), k


def test_partial_and_multi_model_load(tmp_path):
    params = _params(str(tmp_path))
    torch.manual_seed(0)
    full, mp = _tiny_model()
    opt = torch.optim.Adam(full.parameters(), lr=1e-3)
    ck.saveModel(params, full, opt, 1, 'ico2ico', 0.0, None)
    torch.manual_seed(5)
    enc = gm.ico2enc(mp)                                   # encoder-only model takes its keys from the full checkpoint
    assert ck.loadModel(params, enc, [1], 'ico2ico') is True
    fs = full.state_dict()
    assert all(torch.equal(v, fs[k]) for k, v in enc.state_dict().items())
    torch.manual_seed(6)
    full2, _ = _tiny_model()
    ck.saveModel(params, full2, opt, 2, 'other', 0.0, None)
    torch.manual_seed(7)
    mixed, _ = _tiny_model()
    assert ck.loadMultiModel(params, mixed, [1, 2], ['ico2ico', 'other'])               # first checkpoint wins every shared key
    assert all(torch.equal(v, fs[k]) for k, v in mixed.state_dict().items())
    with pytest.raises(ValueError):
        ck.loadMultiModel(params, mixed, [9], ['ico2ico'])
    bad = {k: (v[..., :3] if k.endswith('conv01.weight') else v) for k, v in fs.items()}
    torch.save({'model_state_dict': bad, 'optimizer_state_dict': {}, 'epoch': 4, 'loss': 0, 'misc': None},
               os.path.join(params['logDir'], 'savedModel', 'ico2ico_E4.pt'))
    with pytest.raises(ValueError, match='conv01.weight'):
        ck.loadModel(params, mixed, [4], 'ico2ico')
    # foreign layout: translate names and tensors on the way in
    foreign = {'module.' + k: (v * 2 if k.endswith('.bias') else v) for k, v in fs.items()}
    torch.save({'model_state_dict': foreign, 'optimizer_state_dict': {}, 'epoch': 5, 'loss': 0, 'misc': None},
               os.path.join(params['logDir'], 'savedModel', 'ico2ico_E5.pt'))
    assert ck.loadModel(params, mixed, [5], 'ico2ico', key_map=lambda k: k[len('module.'):],
                        tensor_map=lambda k, t: t / 2 if k.endswith('.bias') else t)
    assert all(torch.allclose(v, fs[k]) for k, v in mixed.state_dict().items())


def test_point_to_mesh_oracle_known_answers():
    V = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [1, 1, 1]], dtype=np.float64)
    F = np.array([[0, 1, 2], [1, 3, 2]])
    pts = np.array([[0.25, 0.25, 0.5],      # above the interior of face 0: 0.5^2
                    [-1.0, -1.0, 0.0],      # nearest to vertex 0: 2
                    [0.5, -2.0, 0.0],       # nearest to edge 0-1: 4
                    [0.25, 0.25, 0.0],      # on the face: 0
                    [1.0, 1.0, 1.0]])       # a vertex of face 1: 0
    d, f = p2m_ref(pts, V, F)
    assert np.allclose(d, [0.25, 2.0, 4.0, 0.0, 0.0], atol=1e-14)
    assert list(f[[0, 1, 2, 3, 4]]) == [0, 0, 0, 0, 1]
    # a degenerate triangle behaves like its longest edge
    d, _ = p2m_ref(np.array([[0.5, 1.0, 0.0]]), np.array([[0, 0, 0], [1, 0, 0], [2, 0, 0.0]]), np.array([[0, 1, 2]]))
    assert np.allclose(d, [1.0])


def test_eval_helpers_refuse_cpu_tensors():
    with pytest.raises(RuntimeError):
        iu.point_to_mesh_distance(torch.zeros(1, 4, 3), torch.zeros(1, 3, 3), torch.tensor([[0, 1, 2]]))
    with pytest.raises(ValueError):
        iu.point_to_mesh_distance(torch.zeros(1, 4, 2), torch.zeros(1, 3, 3), torch.tensor([[0, 1, 2]]))


def test_oracle_synthetic_meshes_match_the_product_generator():
    """oracle/synth_ref.py (what the CPU arm of bench.py trains on, no product import) and geniconet_b200.data.synthetic_mesh
    state the same SURVEY 8d recipe over two independent geometry builders (numpy oracle vs the C ABI's host tables)."""
    import torch
    from oracle import synth_ref
    from geniconet_b200 import data
    for level, idx in ((2, 0), (4, 7)):
        xo, to = synth_ref.synthetic_mesh(level, idx)
        xp, tp = data.synthetic_mesh(level, idx)
        assert xo.shape == xp.shape and to.shape == tp.shape
        assert torch.allclose(xo, xp, rtol=0, atol=2e-6) and torch.allclose(to, tp, rtol=0, atol=2e-5)
