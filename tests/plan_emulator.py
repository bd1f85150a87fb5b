"""Test tool: interpret plan blobs (geniconet_b200/csrc/gin_plan.h) in numpy, exactly as the device
kernels do, so the index tables can be checked against the oracle without a GPU."""
import numpy as np

TILE = 128
MAX_SLOTS = 24


def parse_side(blob, off):
    names = ['ntiles', 'tiles_off', 'src_off', 'rows_off', 'P_src', 'P_dst', 'ring_off', 'max_slots']
    return dict(zip(names, [int(x) for x in blob[off:off + 8]]))


def parse_conv(blob):
    h = dict(zip(['magic', 'kind', 'level_in', 'level_out', 'stride', 'corner_mode', 'group', 'total_words'],
                 [int(x) for x in blob[:8]]))
    h['fwd'] = parse_side(blob, 8)
    h['dg'] = parse_side(blob, 16)
    return h


def tiles(blob, side):
    out = []
    for t in range(side['ntiles']):
        w = blob[side['tiles_off'] + 8 * t: side['tiles_off'] + 8 * t + 8]
        taps = np.frombuffer(w[2:].tobytes(), dtype=np.int8)
        out.append((int(w[0]), int(w[1]), taps[:int(w[0])].astype(int)))
    return out


def gather_rows(x, codes, base_sample, ring, P_src, scale_zero=True):
    """x [B,P_src,C]; codes [128] -> [128,C] exactly like resolve_src/load_row8."""
    B, _, C = x.shape
    out = np.zeros((len(codes), C), dtype=x.dtype)
    flat = x.reshape(B * P_src, C)
    for r, c in enumerate(codes):
        c = int(c)
        if c >= 0:
            gp = base_sample * P_src + c
            if gp < B * P_src:
                out[r] = flat[gp]
        elif c <= -2:
            q = -2 - c
            smp = base_sample + (q >> 1)
            if smp < B:
                out[r] = x[smp, ring[(q & 1) * 5:(q & 1) * 5 + 5]].mean(0)
    return out


def run_side(blob, side, group, x, W, bias=None):
    """dst[B,P_dst,N] = gather-GEMM of x [B,P_src,K] with W [7,K,N]."""
    B, _, K = x.shape
    N = W.shape[2]
    y = np.full((B, side['P_dst'], N), np.nan, dtype=x.dtype)
    yf = y.reshape(B * side['P_dst'], N)
    ring = blob[side['ring_off']:side['ring_off'] + 10]
    tl = tiles(blob, side)
    groups = (B + group - 1) // group
    for G in range(groups):
        for t, (nslots, soff, taps) in enumerate(tl):
            acc = np.zeros((TILE, N), dtype=x.dtype)
            for s in range(nslots):
                codes = blob[side['src_off'] + soff + s * TILE: side['src_off'] + soff + (s + 1) * TILE]
                acc += gather_rows(x, codes, G * group, ring, side['P_src']) @ W[taps[s]]
            rows = blob[side['rows_off'] + t * TILE: side['rows_off'] + (t + 1) * TILE]
            for r, d in enumerate(rows):
                if d >= 0:
                    gd = G * group * side['P_dst'] + int(d)
                    if gd < B * side['P_dst']:
                        yf[gd] = acc[r] + (bias if bias is not None else 0)
    return y


def run_wgrad(blob, side, group, x, dy):
    """dW [7,K,N] from the forward side."""
    B, _, K = x.shape
    N = dy.shape[2]
    dW = np.zeros((7, K, N), dtype=x.dtype)
    dyf = dy.reshape(B * side['P_dst'], N)
    ring = blob[side['ring_off']:side['ring_off'] + 10]
    tl = tiles(blob, side)
    groups = (B + group - 1) // group
    for G in range(groups):
        for t, (nslots, soff, taps) in enumerate(tl):
            rows = blob[side['rows_off'] + t * TILE: side['rows_off'] + (t + 1) * TILE]
            g = np.zeros((TILE, N), dtype=x.dtype)
            for r, d in enumerate(rows):
                if d >= 0 and G * group * side['P_dst'] + int(d) < B * side['P_dst']:
                    g[r] = dyf[G * group * side['P_dst'] + int(d)]
            for s in range(nslots):
                codes = blob[side['src_off'] + soff + s * TILE: side['src_off'] + soff + (s + 1) * TILE]
                dW[taps[s]] += gather_rows(x, codes, G * group, ring, side['P_src']).T @ g
    return dW


def parse_up(blob):
    names = ['magic', 'kind', 'level', 'corner_mode', 'Pc', 'Pf', 'total_words', 'fwd_off', 'ring_off', 'bwd_deg',
             'bwd_idx_off', 'bwd_w_off']
    return dict(zip(names, [int(x) for x in blob[:12]]))


def run_up_fwd(blob, x):
    h = parse_up(blob)
    B, Pc, C = x.shape
    src = blob[h['fwd_off']:h['fwd_off'] + 2 * h['Pf']].reshape(h['Pf'], 2)
    ring = blob[h['ring_off']:h['ring_off'] + 10]
    ext = np.concatenate([x, np.zeros((B, 1, C), x.dtype), x[:, ring[5:10]].mean(1, keepdims=True),
                          x[:, ring[0:5]].mean(1, keepdims=True)], axis=1)   # index -1 zero, -3 south, -2 north (from the end)
    # python negative indexing: -1 -> last (north?) -- build explicit map instead
    def fetch(code):
        out = np.zeros((B, len(code), C), x.dtype)
        pos = code >= 0
        out[:, pos] = x[:, code[pos]]
        out[:, code == -2] = x[:, ring[0:5]].mean(1, keepdims=True)
        out[:, code == -3] = x[:, ring[5:10]].mean(1, keepdims=True)
        return out
    return 0.5 * (fetch(src[:, 0]) + fetch(src[:, 1]))


def run_up_bwd(blob, dy):
    h = parse_up(blob)
    B, Pf, C = dy.shape
    deg = h['bwd_deg']
    idx = blob[h['bwd_idx_off']:h['bwd_idx_off'] + h['Pc'] * deg].reshape(h['Pc'], deg)
    w = blob[h['bwd_w_off']:h['bwd_w_off'] + h['Pc'] * deg].view(np.float32).reshape(h['Pc'], deg)
    dx = np.zeros((B, h['Pc'], C), dy.dtype)
    for e in range(deg):
        ok = idx[:, e] >= 0
        dx[:, ok] += w[ok, e][None, :, None] * dy[:, idx[ok, e]]
    return dx


# ---------------------------------------------------------------- patch-mode tables
TAPS = ((0, 0), (-1, 0), (1, 0), (0, -1), (0, 1), (-1, 1), (1, -1))


def parse_pside(blob, off):
    names = ['R', 'Q', 'U', 'ntiles', 'src_off', 'rows_off', 'ring_off', 'pad']
    return dict(zip(names, [int(x) for x in blob[off:off + 8]]))


def parse_conv_full(blob):
    h = parse_conv(blob)
    h['pfwd'] = parse_pside(blob, 24)
    h['pdg'] = parse_pside(blob, 32)
    h['dgx'] = parse_side(blob, 40)
    return h


def run_pside(blob, ps, group, x, W, mirror, bias=None):
    """Emulates gin_gemm_tcp.cuh: three column-shifted copies of the padded patch, taps = descriptor offsets."""
    B, P, K = x.shape
    N = W.shape[2]
    R, Q, U = ps['R'], ps['Q'], ps['U']
    y = np.full((B, P, N), np.nan, dtype=x.dtype)
    yf = y.reshape(B * P, N)
    ring = blob[ps['ring_off']:ps['ring_off'] + 10]
    groups = (B + group - 1) // group
    for G in range(groups):
        for t in range(ps['ntiles']):
            codes = blob[ps['src_off'] + t * U: ps['src_off'] + (t + 1) * U]
            rows = gather_rows(x, codes, G * group, ring, P)                # [U, K]
            copies = np.zeros((3, (R + 2) * Q * 8, K), dtype=x.dtype)       # smem image: [copy][cell*8+px]
            for u in range(U):
                cell, c = divmod(u, 10)
                if c <= 7:
                    copies[0, cell * 8 + c] = rows[u]
                if 1 <= c <= 8:
                    copies[1, cell * 8 + c - 1] = rows[u]
                if c >= 2:
                    copies[2, cell * 8 + c - 2] = rows[u]
            acc = np.zeros((TILE, N), dtype=x.dtype)
            for tap, (di, dj) in enumerate(TAPS):
                if mirror:
                    di, dj = -di, -dj
                start = (1 + di) * Q * 8
                acc += copies[dj + 1, start:start + TILE] @ W[tap]
            drow = blob[ps['rows_off'] + t * TILE: ps['rows_off'] + (t + 1) * TILE]
            for r, d in enumerate(drow):
                gd = G * group * P + int(d)
                if gd < B * P:
                    yf[gd] = acc[r] + (bias if bias is not None else 0)
    return y


def run_side_accumulate(blob, side, group, x, W, y):
    """Gather-mode pass that ADDS into y (the dgx seam pass)."""
    B = x.shape[0]
    yf = y.reshape(B * side['P_dst'], -1)
    ring = blob[side['ring_off']:side['ring_off'] + 10]
    groups = (B + group - 1) // group
    for G in range(groups):
        for t, (nslots, soff, taps) in enumerate(tiles(blob, side)):
            acc = np.zeros((TILE, W.shape[2]), dtype=x.dtype)
            for s in range(nslots):
                codes = blob[side['src_off'] + soff + s * TILE: side['src_off'] + soff + (s + 1) * TILE]
                acc += gather_rows(x, codes, G * group, ring, side['P_src']) @ W[taps[s]]
            rows = blob[side['rows_off'] + t * TILE: side['rows_off'] + (t + 1) * TILE]
            for r, d in enumerate(rows):
                if d >= 0:
                    gd = G * group * side['P_dst'] + int(d)
                    if gd < B * side['P_dst']:
                        yf[gd] += acc[r]
    return y


# ---------------------------------------------------------------- second-generation patch kernels (gin_conv2.cuh, gin_wgrad2.cuh)
def image_rows(start, group_rows=10):
    """Rows of the single-copy patch image read by the 128 tile rows for a tap that starts at `start`."""
    m = np.arange(TILE)
    return start + (m // 8) * group_rows + (m % 8)


def tap_start(Q, a, b):
    return (1 + a) * Q * 10 + (1 + b)


def run_patch2(blob, ps, group, x, W, mirror, bias=None):
    """Stride-1 forward (mirror=0) / in-chart dgrad (mirror=1) as gin_conv2.cuh computes it: ONE image per tile in plan order,
    tap (di,dj) = the rows start + (m/8)*10 + m%8."""
    B, P, K = x.shape
    N = W.shape[2]
    Q, U = ps['Q'], ps['U']
    y = np.full((B, P, N), np.nan, dtype=x.dtype)
    yf = y.reshape(B * P, N)
    ring = blob[ps['ring_off']:ps['ring_off'] + 10]
    for G in range((B + group - 1) // group):
        for t in range(ps['ntiles']):
            img = gather_rows(x, blob[ps['src_off'] + t * U: ps['src_off'] + (t + 1) * U], G * group, ring, P)
            acc = np.zeros((TILE, N), dtype=x.dtype)
            for tap, (di, dj) in enumerate(TAPS):
                if mirror:
                    di, dj = -di, -dj
                acc += img[image_rows(tap_start(Q, di, dj))] @ W[tap]
            # destination rows the way the epilogue computes them: per-octet base pixel + r*W + px
            base = blob[ps['rows_off'] + t * TILE: ps['rows_off'] + (t + 1) * TILE]
            n2 = int(round((P / 10) ** 0.5)) * 2
            for m in range(TILE):
                g, px = divmod(m, 8)
                r, q = divmod(g, Q)
                gd = G * group * P + int(base[q * 8]) + r * n2 + px
                assert int(base[m]) == int(base[q * 8]) + r * n2 + px
                if gd < B * P:
                    yf[gd] = acc[m] + (bias if bias is not None else 0)
    return y


def run_wgrad2(blob, ps, group, x, dy):
    """gin_wgrad2.cuh: dW[tap] += image[tap rows]^T @ dY tile."""
    B, P, K = x.shape
    N = dy.shape[2]
    Q, U = ps['Q'], ps['U']
    dW = np.zeros((7, K, N), dtype=x.dtype)
    dyf = dy.reshape(B * P, N)
    ring = blob[ps['ring_off']:ps['ring_off'] + 10]
    for G in range((B + group - 1) // group):
        for t in range(ps['ntiles']):
            img = gather_rows(x, blob[ps['src_off'] + t * U: ps['src_off'] + (t + 1) * U], G * group, ring, P)
            rows = blob[ps['rows_off'] + t * TILE: ps['rows_off'] + (t + 1) * TILE]
            g = np.zeros((TILE, N), dtype=x.dtype)
            for m, d in enumerate(rows):
                if G * group * P + int(d) < B * P:
                    g[m] = dyf[G * group * P + int(d)]
            for tap, (di, dj) in enumerate(TAPS):
                dW[tap] += img[image_rows(tap_start(Q, di, dj))].T @ g
    return dW


# stride 2 (GinP2Side): forward tap -> (parity plane, coarse offset)
S2_TAP = {0: (2, 0, 0), 1: (0, 0, 0), 2: (0, 1, 0), 3: (3, 0, -1), 4: (3, 0, 0), 5: (1, 0, 0), 6: (1, 1, -1)}


def parse_p2(blob):
    names = ['R', 'Q', 'U', 'ntiles', 'src_off', 'dsrc_off', 'rows_off', 'frows_off']
    return dict(zip(names, [int(v) for v in blob[48:56]]))


def run_p2_fwd(blob, h, p2, x, W, bias=None):
    """Stride-2 forward: per plane an image gathered from the FINE map, 1-2 taps each, one accumulator."""
    B, Pf, K = x.shape
    N = W.shape[2]
    Pc, group, Q, U = h['fwd']['P_dst'], h['group'], p2['Q'], p2['U']
    ring = blob[h['fwd']['ring_off']:h['fwd']['ring_off'] + 10]
    y = np.full((B, Pc, N), np.nan, dtype=x.dtype)
    yf = y.reshape(B * Pc, N)
    for G in range((B + group - 1) // group):
        for t in range(p2['ntiles']):
            acc = np.zeros((TILE, N), dtype=x.dtype)
            for pl in range(4):
                off = p2['src_off'] + (t * 4 + pl) * U
                img = gather_rows(x, blob[off:off + U], G * group, ring, Pf)
                for tap, (tpl, a, b) in S2_TAP.items():
                    if tpl == pl:
                        acc += img[image_rows(tap_start(Q, a, b))] @ W[tap]
            rows = blob[p2['rows_off'] + t * TILE: p2['rows_off'] + (t + 1) * TILE]
            for m, d in enumerate(rows):
                gd = G * group * Pc + int(d)
                if gd < B * Pc:
                    yf[gd] = acc[m] + (bias if bias is not None else 0)
    return y


def run_p2_wgrad(blob, h, p2, x, dy):
    B, Pf, K = x.shape
    N = dy.shape[2]
    Pc, group, Q, U = h['fwd']['P_dst'], h['group'], p2['Q'], p2['U']
    ring = blob[h['fwd']['ring_off']:h['fwd']['ring_off'] + 10]
    dW = np.zeros((7, K, N), dtype=x.dtype)
    dyf = dy.reshape(B * Pc, N)
    for G in range((B + group - 1) // group):
        for t in range(p2['ntiles']):
            rows = blob[p2['rows_off'] + t * TILE: p2['rows_off'] + (t + 1) * TILE]
            g = np.zeros((TILE, N), dtype=x.dtype)
            for m, d in enumerate(rows):
                if G * group * Pc + int(d) < B * Pc:
                    g[m] = dyf[G * group * Pc + int(d)]
            for pl in range(4):
                off = p2['src_off'] + (t * 4 + pl) * U
                img = gather_rows(x, blob[off:off + U], G * group, ring, Pf)
                for tap, (tpl, a, b) in S2_TAP.items():
                    if tpl == pl:
                        dW[tap] += img[image_rows(tap_start(Q, a, b))].T @ g
    return dW


def run_p2_dgrad(blob, h, p2, dy, Wd):
    """In-chart part of the stride-2 dgrad: the fine pixels of plane pl gather the coarse dy image with offsets (-a, -b);
    destination = frows[t][q] + r*2*Wf + 2*px + pr*Wf + pc.  Wd is [7][Cout][Cin]."""
    B, Pc, K = dy.shape
    N = Wd.shape[2]
    Pf, group, Q, U = h['fwd']['P_src'], h['group'], p2['Q'], p2['U']
    Wf = int(round((Pf / 10) ** 0.5)) * 2
    dx = np.full((B, Pf, N), np.nan, dtype=dy.dtype)
    dxf = dx.reshape(B * Pf, N)
    ring = np.zeros(10, dtype=np.int64)
    for G in range((B + group - 1) // group):
        for t in range(p2['ntiles']):
            img = gather_rows(dy, blob[p2['dsrc_off'] + t * U: p2['dsrc_off'] + (t + 1) * U], G * group, ring, Pc)
            fr = blob[p2['frows_off'] + t * Q: p2['frows_off'] + (t + 1) * Q]
            for pl in range(4):
                pr, pc = pl >> 1, pl & 1
                acc = np.zeros((TILE, N), dtype=dy.dtype)
                for tap, (tpl, a, b) in S2_TAP.items():
                    if tpl == pl:
                        acc += img[image_rows(tap_start(Q, -a, -b))] @ Wd[tap]
                for m in range(TILE):
                    g, px = divmod(m, 8)
                    r, q = divmod(g, Q)
                    gd = G * group * Pf + int(fr[q]) + r * 2 * Wf + 2 * px + pr * Wf + pc
                    if gd < B * Pf:
                        dxf[gd] = acc[m]
    return dx


def parse_px(blob):
    d = dict(zip(['ntiles', 'nslots', 'src_off', 'dst_off'], [int(v) for v in blob[56:60]]))
    d['tap'] = np.frombuffer(blob[60:64].tobytes(), dtype=np.int8)[:d['nslots']].astype(int)
    return d


def run_px_accumulate(blob, h, px, dy, Wd, dx):
    """The regular-form seam pass (GinPxSide) as the patch kernel runs it: every tile applies ALL slots, then adds into dx."""
    B = dy.shape[0]
    P_src, P_dst, group = h['dgx']['P_src'], h['dgx']['P_dst'], h['group']
    ring = blob[h['dgx']['ring_off']:h['dgx']['ring_off'] + 10]
    dxf = dx.reshape(B * P_dst, -1)
    seen = set()
    for G in range((B + group - 1) // group):
        for t in range(px['ntiles']):
            acc = np.zeros((TILE, Wd.shape[2]), dtype=dy.dtype)
            for s in range(px['nslots']):
                off = px['src_off'] + (t * px['nslots'] + s) * TILE
                acc += gather_rows(dy, blob[off:off + TILE], G * group, ring, P_src) @ Wd[px['tap'][s]]
            dst = blob[px['dst_off'] + t * TILE: px['dst_off'] + (t + 1) * TILE]
            for r, d in enumerate(dst):
                d = int(d)
                if d < 0:
                    assert d in (-1, -3)
                    continue
                tot = acc[r].copy()
                for k in (1, 2):                            # extra rows (dst -3) directly below, inside the same 32-row group
                    if (r + k) // 32 == r // 32 and r + k < TILE and all(int(dst[r + j]) == -3 for j in range(1, k + 1)):
                        tot = tot + acc[r + k]
                gd = G * group * P_dst + d
                if gd < B * P_dst:
                    assert gd not in seen                   # one read-modify-write per pixel: race free without atomics
                    seen.add(gd)
                    dxf[gd] += tot
            for r, d in enumerate(dst):                     # every extra row has its pixel row at most two above, same group
                if int(d) == -3:
                    k = 1 if int(dst[r - 1]) >= 0 else 2
                    assert r - k >= 0 and (r - k) // 32 == r // 32 and int(dst[r - k]) >= 0 and all(int(dst[r - j]) == -3 for j in range(0, k))
    return dx


def parse_pf(blob):
    """GinPfSide (words 64..75 of the conv plan header): one-launch dgrad."""
    d = dict(zip(['ntiles', 'nslots', 'src_off', 'dst_off', 'mask_off', 'nfl', 'all'], [int(v) for v in blob[64:71]]))
    d['tap'] = np.frombuffer(blob[72:76].tobytes(), dtype=np.int8)[:d['nslots']].astype(int)
    return d


def run_pf(blob, h, pf, dy, Wd, dx_inchart):
    """One-launch dgrad as the patch kernel runs it: the in-chart tiles do NOT store the rows their mask flags; the boundary tiles
    (GinPfSide) then store the complete gradient of exactly those pixels -- plain stores, each pixel written once overall."""
    B = dy.shape[0]
    group, stride = h['group'], h['stride']
    P_src, P_dst = h['dg']['P_src'], h['dg']['P_dst']
    ring = blob[h['dgx']['ring_off']:h['dgx']['ring_off'] + 10]
    dx = dx_inchart.copy()
    if pf['all']:                                          # small levels: the in-chart tiles are not run at all
        dx[:] = np.nan
    dxf = dx.reshape(B * P_dst, -1)
    masked = set()
    groups = (B + group - 1) // group
    if stride == 1:
        ps = h['pdg']
        assert pf['nfl'] == 1
        for G in range(groups):
            for t in range(ps['ntiles']):
                rows = blob[ps['rows_off'] + t * TILE: ps['rows_off'] + (t + 1) * TILE]
                words = blob[pf['mask_off'] + t * 4: pf['mask_off'] + t * 4 + 4].astype(np.int64) & 0xffffffff
                for r in range(TILE):
                    if (int(words[r // 32]) >> (r % 32)) & 1:
                        gd = G * group * P_dst + int(rows[r])
                        if gd < B * P_dst:
                            masked.add(gd)
    else:
        p2 = parse_p2(blob)
        assert pf['nfl'] == 4
        Q, Wf = p2['Q'], 2 << h['level_in']
        for G in range(groups):
            for t in range(p2['ntiles']):
                fr = blob[p2['frows_off'] + t * Q: p2['frows_off'] + (t + 1) * Q]
                for pl in range(4):
                    words = blob[pf['mask_off'] + (t * 4 + pl) * 4: pf['mask_off'] + (t * 4 + pl) * 4 + 4].astype(np.int64) & 0xffffffff
                    for r in range(TILE):
                        if (int(words[r // 32]) >> (r % 32)) & 1:
                            g, px = divmod(r, 8)
                            r_in, q = divmod(g, Q)
                            gd = G * group * P_dst + int(fr[q]) + r_in * 2 * Wf + 2 * px + (pl >> 1) * Wf + (pl & 1)
                            if gd < B * P_dst:
                                masked.add(gd)
    for gd in masked:
        dxf[gd] = np.nan                                   # not stored by the in-chart pass
    written = set()
    for G in range(groups):
        for t in range(pf['ntiles']):
            acc = np.zeros((TILE, Wd.shape[2]), dtype=dy.dtype)
            for s in range(pf['nslots']):
                off = pf['src_off'] + (t * pf['nslots'] + s) * TILE
                acc += gather_rows(dy, blob[off:off + TILE], G * group, ring, P_src) @ Wd[pf['tap'][s]]
            dst = blob[pf['dst_off'] + t * TILE: pf['dst_off'] + (t + 1) * TILE]
            for r, d in enumerate(dst):
                d = int(d)
                if d < 0:
                    assert d in (-1, -3)
                    continue
                tot = acc[r].copy()
                for k in (1, 2):
                    if (r + k) // 32 == r // 32 and r + k < TILE and all(int(dst[r + j]) == -3 for j in range(1, k + 1)):
                        tot = tot + acc[r + k]
                gd = G * group * P_dst + d
                if gd < B * P_dst:
                    assert gd in masked and gd not in written      # exactly the pixels the in-chart pass left out, once each
                    written.add(gd)
                    dxf[gd] = tot
            for r, d in enumerate(dst):
                if int(d) == -3:
                    k = 1 if int(dst[r - 1]) >= 0 else 2
                    assert r - k >= 0 and (r - k) // 32 == r // 32 and int(dst[r - k]) >= 0
    assert written == masked
    if pf['all']:
        assert len(written) == B * P_dst
    return dx
