"""Isolated hex-conv layer sweep (BASELINE.json configs[4]): levels x channel widths x stride x {fwd, dgrad, wgrad} through the
C ABI on tensors sized past the L2, device time from a replayed CUDA graph.  Writes a markdown table.

    python tools/layer_sweep.py [out.md]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geniconet_b200 import _lib                          # noqa: E402
from geniconet_b200.ico_conv import get_plan            # noqa: E402

L = _lib.lib
CH = [(64, 64), (64, 128), (128, 64), (128, 128), (128, 256), (256, 128), (256, 256), (256, 512), (512, 256)]
LEVELS = [3, 4, 5, 6, 7]
ITERS = 6


def timed(fn, st_holder):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    side = torch.cuda.Stream()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.stream(side):
        st_holder[0] = side.cuda_stream
        with torch.cuda.graph(g, stream=side):
            for _ in range(ITERS):
                fn()
    st_holder[0] = torch.cuda.current_stream().cuda_stream
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
    return e0.elapsed_time(e1) * 1e3 / ITERS


def main():
    out = sys.argv[1] if len(sys.argv) > 1 else os.path.join(ROOT, 'gpurun_out', 'layer_sweep.md')
    rows = ['| level | Cin→Cout | stride | B | GFLOP | in+out MB (bf16 in, fp32 out) | fwd µs | fwd TF/s | fwd GB/s | dgrad µs | dgrad TF/s | wgrad µs | wgrad TF/s |',
            '|---|---|---|---|---|---|---|---|---|---|---|---|---|']
    for lvl in LEVELS:
        for ci, co in CH:
            for stride in (1, 2):
                Pin = 10 * 4 ** lvl
                Pout = Pin // stride ** 2
                per_sample = 2.0 * ci * Pin + 4.0 * co * Pout
                B = int(min(36, max(1, round(300e6 / per_sample))))      # ~300 MB of operand + result: well past the 126 MB L2 (capped at 36)
                lvl_out = lvl - (1 if stride == 2 else 0)
                try:
                    plan = get_plan(_lib.PLAN_HEXCONV, lvl, stride, 'average', 'cuda')
                    w = torch.randn(co, ci, 7, device='cuda') * 0.05
                    bias = torch.zeros(co, device='cuda')
                    packed = torch.empty(L.gin_hexconv_packed_bytes(ci, co), dtype=torch.uint8, device='cuda')
                    st = [torch.cuda.current_stream().cuda_stream]
                    _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), ci, co, st[0]))
                    xb = (torch.randn(B * Pin + 2 * B, ci, device='cuda') * 0.5).to(_lib.forward_operand_dtype())
                    dyb = (torch.randn(B * Pout + 2 * B, co, device='cuda') * 0.5).to(torch.bfloat16)
                    y = torch.empty(B * Pout, co, device='cuda')
                    dx = torch.empty(B * Pin, ci, device='cuda')
                    dW = torch.empty(co, ci, 7, device='cuda')
                    ws = torch.empty(L.gin_hexconv_wgrad_ws_bytes(ci, co), dtype=torch.uint8, device='cuda')
                    t = {}
                    t['fwd'] = timed(lambda: _lib.check(L.gin_hexconv_fwd_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), bias.data_ptr(),
                                                                               y.data_ptr(), B, ci, co, st[0])), st)
                    t['dgrad'] = timed(lambda: _lib.check(L.gin_hexconv_dgrad_bf16(plan.host_ptr, plan.dev_ptr, dyb.data_ptr(), packed.data_ptr(), dx.data_ptr(),
                                                                                   B, ci, co, st[0])), st)
                    t['wgrad'] = timed(lambda: _lib.check(L.gin_hexconv_wgrad_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), dyb.data_ptr(), None, dW.data_ptr(),
                                                                                   None, ws.data_ptr(), B, ci, co, st[0])), st)
                    gf = 2.0 * 7 * ci * co * Pout * B / 1e9
                    mb = per_sample * B / 1e6
                    rows.append('| %d | %d→%d | %d | %d | %.1f | %.0f | %.1f | %.0f | %.0f | %.1f | %.0f | %.1f | %.0f |' % (
                        lvl, ci, co, stride, B, gf, mb, t['fwd'], gf / t['fwd'] * 1e3, mb / t['fwd'] * 1e3,
                        t['dgrad'], gf / t['dgrad'] * 1e3, t['wgrad'], gf / t['wgrad'] * 1e3))
                    del xb, dyb, y, dx, dW, ws
                except Exception as e:                       # keep sweeping; the table says what failed
                    rows.append('| %d | %d→%d | %d | %d | failed: %s |' % (lvl, ci, co, stride, B, str(e)[:80]))
                torch.cuda.empty_cache()
    with open(out, 'w') as f:
        f.write('\n'.join(rows) + '\n')
    print('\n'.join(rows))


if __name__ == '__main__':
    main()
