"""The plan tables built by the C++ host code (gin_host.cpp) interpreted in numpy exactly as the kernels interpret
them (tests/plan_emulator.py) must reproduce the oracle's forward, dgrad, wgrad and upsample -- no GPU involved."""
import numpy as np
import pytest
import torch

import plan_emulator as pe
from geniconet_b200 import _lib
from oracle import icocnn_ref

CONV_CASES = [  # level, stride, corner_mode, B, Cin, Cout
    (2, 1, 'average', 5, 5, 6), (2, 1, 'zeros', 3, 4, 3), (3, 2, 'average', 3, 4, 5), (1, 1, 'average', 18, 3, 4),
    (2, 2, 'average', 5, 3, 3), (3, 1, 'average', 2, 3, 2), (0, 1, 'average', 3, 2, 2), (1, 2, 'zeros', 2, 2, 3),
]


@pytest.mark.parametrize('case', CONV_CASES)
def test_hexconv_plan_reproduces_oracle(case):
    lvl, stride, cm, B, Cin, Cout = case
    torch.manual_seed(0)
    blob = _lib.plan_blob(_lib.PLAN_HEXCONV, lvl, stride, cm)
    h = pe.parse_conv(blob)
    assert h['magic'] == 0x47494E31 and h['total_words'] == len(blob)
    m = icocnn_ref.IcoConvS2S(Cin, Cout, stride, True, lvl, cm).double()
    x = torch.randn(B, Cin, 5 * 2 ** lvl, 2 * 2 ** lvl, dtype=torch.double, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xn = x.detach().permute(0, 2, 3, 1).reshape(B, -1, Cin).numpy()
    W = m.weight.detach().permute(2, 1, 0).numpy()
    ye = pe.run_side(blob, h['fwd'], h['group'], xn, W, m.bias.detach().numpy())
    assert np.abs(ye - y.detach().permute(0, 2, 3, 1).reshape(B, -1, Cout).numpy()).max() < 1e-12
    gyn = gy.permute(0, 2, 3, 1).reshape(B, -1, Cout).numpy()
    dxe = pe.run_side(blob, h['dg'], h['group'], gyn, np.ascontiguousarray(W.transpose(0, 2, 1)))
    assert not np.isnan(dxe).any()      # every input pixel is written exactly once
    assert np.abs(dxe - x.grad.permute(0, 2, 3, 1).reshape(B, -1, Cin).numpy()).max() < 1e-12
    dWe = pe.run_wgrad(blob, h['fwd'], h['group'], xn, gyn)
    assert np.abs(dWe - m.weight.grad.permute(2, 1, 0).numpy()).max() < 1e-10


PATCH_CASES = [  # level, stride, corner_mode, B, Cin, Cout
    (2, 1, 'average', 5, 3, 4), (3, 1, 'average', 2, 3, 2), (3, 1, 'zeros', 1, 2, 3), (4, 1, 'average', 1, 2, 2),
    (3, 2, 'average', 5, 3, 2), (4, 2, 'average', 2, 2, 3), (4, 2, 'zeros', 1, 2, 2), (5, 2, 'average', 1, 1, 2),
    (5, 1, 'average', 3, 1, 2), (6, 2, 'average', 1, 1, 1),     # the bench levels: I5 stride 1 with a ragged sample group, I6 -> I5
]


@pytest.mark.parametrize('case', PATCH_CASES)
def test_patch_plan_reproduces_oracle(case):
    """The patch-mode tables (single-copy image, taps = start rows; stride 2 = four parity planes on the coarse lattice)
    interpreted as gin_conv2.cuh / gin_wgrad2.cuh interpret them."""
    lvl, stride, cm, B, Cin, Cout = case
    torch.manual_seed(1)
    blob = _lib.plan_blob(_lib.PLAN_HEXCONV, lvl, stride, cm)
    h = pe.parse_conv_full(blob)
    m = icocnn_ref.IcoConvS2S(Cin, Cout, stride, True, lvl, cm).double()
    x = torch.randn(B, Cin, 5 * 2 ** lvl, 2 * 2 ** lvl, dtype=torch.double, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    xn = x.detach().permute(0, 2, 3, 1).reshape(B, -1, Cin).numpy()
    W = m.weight.detach().permute(2, 1, 0).numpy()                      # [7][Cin][Cout]
    Wd = np.ascontiguousarray(W.transpose(0, 2, 1))                     # [7][Cout][Cin]
    yn = y.detach().permute(0, 2, 3, 1).reshape(B, -1, Cout).numpy()
    gyn = gy.permute(0, 2, 3, 1).reshape(B, -1, Cout).numpy()
    dxn = x.grad.permute(0, 2, 3, 1).reshape(B, -1, Cin).numpy()
    dWn = m.weight.grad.permute(2, 1, 0).numpy()
    if stride == 1:
        ye = pe.run_patch2(blob, h['pfwd'], h['group'], xn, W, 0, m.bias.detach().numpy())
        dxe = pe.run_patch2(blob, h['pdg'], h['group'], gyn, Wd, 1)
        dWe = pe.run_wgrad2(blob, h['pfwd'], h['group'], xn, gyn)
    else:
        p2 = pe.parse_p2(blob)
        assert p2['ntiles'] > 0
        ye = pe.run_p2_fwd(blob, h, p2, xn, W, m.bias.detach().numpy())
        dxe = pe.run_p2_dgrad(blob, h, p2, gyn, Wd)
        dWe = pe.run_p2_wgrad(blob, h, p2, xn, gyn)
    assert np.abs(ye - yn).max() < 1e-12
    assert not np.isnan(dxe).any()                                      # the in-chart pass writes every input pixel once
    px = pe.parse_px(blob)
    assert px['ntiles'] > 0 and px['nslots'] <= 16
    dxe2 = pe.run_px_accumulate(blob, h, px, gyn, Wd, dxe.copy())            # + remainder, regular form (patch kernel)
    assert np.abs(dxe2 - dxn).max() < 1e-12
    pf = pe.parse_pf(blob)                                                    # one-launch dgrad: masked in-chart tiles + boundary tiles
    assert pf['ntiles'] > 0 and pf['nslots'] <= 16 and pf['nfl'] == (1 if stride == 1 else 4)
    dxe3 = pe.run_pf(blob, h, pf, gyn, Wd, dxe)
    assert not np.isnan(dxe3).any() and np.abs(dxe3 - dxn).max() < 1e-12
    print('level %d stride %d: boundary tiles %d x %d slots per group of %d (in-chart tiles %d)' % (
        lvl, stride, pf['ntiles'], pf['nslots'], h['group'], h['pdg']['ntiles'] if stride == 1 else pe.parse_p2(blob)['ntiles']))
    dxe = pe.run_side_accumulate(blob, h['dgx'], h['group'], gyn, Wd, dxe)   # + cross-seam / pole remainder (gather kernel)
    assert np.abs(dxe - dxn).max() < 1e-12
    assert np.abs(dWe - dWn).max() < 1e-10
    # one boundary pixel per seam row: the seam pass may add without atomics
    rows = blob[h['dgx']['rows_off']:h['dgx']['rows_off'] + h['dgx']['ntiles'] * 128]
    valid = rows[rows >= 0]
    assert len(set(valid.tolist())) == len(valid)


@pytest.mark.parametrize('lvl', [2, 3, 4, 5])
def test_dgrad_plan_shape(lvl):
    """Interior rows form pure 7-slot tiles (sorted first); seam rows carry the extra slots; stride 2 needs <= 4."""
    blob = _lib.plan_blob(_lib.PLAN_HEXCONV, lvl, 1, 'average')
    h = pe.parse_conv(blob)
    tl = pe.tiles(blob, h['dg'])
    assert min(t[0] for t in tl) == 7 and max(t[0] for t in tl) <= 24
    assert all(t[0] == 7 for t in pe.tiles(blob, h['fwd']))
    rows = blob[h['dg']['rows_off']:h['dg']['rows_off'] + h['dg']['ntiles'] * 128]
    valid = rows[rows >= 0]
    assert sorted(valid.tolist()) == list(range(h['group'] * h['dg']['P_dst']))
    if lvl >= 3:
        b2 = _lib.plan_blob(_lib.PLAN_HEXCONV, lvl, 2, 'average')
        assert max(t[0] for t in pe.tiles(b2, pe.parse_conv(b2)['dg'])) <= 4


@pytest.mark.parametrize('lvl,cm', [(0, 'average'), (1, 'average'), (2, 'zeros'), (3, 'average')])
def test_upsample_plan_reproduces_oracle(lvl, cm):
    blob = _lib.plan_blob(_lib.PLAN_UPSAMPLE, lvl, 1, cm)
    m = icocnn_ref.IcoUpsampleS2S(4, lvl, cm).double()
    x = torch.randn(3, 4, 5 * 2 ** lvl, 2 * 2 ** lvl, dtype=torch.double, requires_grad=True)
    y = m(x)
    gy = torch.randn_like(y)
    y.backward(gy)
    ye = pe.run_up_fwd(blob, x.detach().permute(0, 2, 3, 1).reshape(3, -1, 4).numpy())
    assert np.abs(ye - y.detach().permute(0, 2, 3, 1).reshape(3, -1, 4).numpy()).max() < 1e-12
    dxe = pe.run_up_bwd(blob, gy.permute(0, 2, 3, 1).reshape(3, -1, 4).numpy())
    assert np.abs(dxe - x.grad.permute(0, 2, 3, 1).reshape(3, -1, 4).numpy()).max() < 1e-6   # 0.1f weights are float32


def test_loss_plan_rings():
    from oracle import ico_geometry_ref as geo
    for s in (1, 2, 3):
        blob = _lib.plan_blob(_lib.PLAN_LOSS, s)
        P, V = int(blob[3]), int(blob[4])
        ring = blob[int(blob[6]):int(blob[6]) + V * 6].reshape(V, 6)
        v, f = geo.get_icosahedral_grid(s)
        nbrs = [set() for _ in range(V)]
        for a, b, c in f.tolist():
            nbrs[a] |= {b, c}; nbrs[b] |= {a, c}; nbrs[c] |= {a, b}
        for u in range(V):
            r = [int(t) for t in ring[u] if t >= 0]
            assert set(r) == nbrs[u] and len(r) == len(nbrs[u])
            acc = sum(np.cross(v[r[i]], v[r[(i + 1) % len(r)]]) for i in range(len(r)))
            assert acc @ v[u] > 0                       # counter-clockwise seen from outside
            for i in range(len(r)):                     # consecutive ring vertices are themselves adjacent
                assert r[(i + 1) % len(r)] in nbrs[r[i]]


def test_plan_errors():
    with pytest.raises(_lib.GinError):
        _lib.plan_blob(_lib.PLAN_HEXCONV, 0, 2, 'average')
    with pytest.raises(_lib.GinError):
        _lib.plan_blob(_lib.PLAN_HEXCONV, 3, 3, 'average')
    with pytest.raises(_lib.GinError):
        _lib.plan_blob(99, 3, 1, 'average')
