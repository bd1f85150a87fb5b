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
