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
ITERS = 4
TARGET_BYTES = 512e6


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
                B = int(max(1, -(-TARGET_BYTES // per_sample)))          # input + output >= 512 MB: far past the 126 MB L2
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
    rows += ['', '## Memory-bound end (judged in GB/s against the measured 6546 GB/s; algorithmic bytes = one read of the input + one write of the output)', '',
             '| op | level | channels | B | MB | µs | GB/s | of HBM peak |', '|---|---|---|---|---|---|---|---|']
    peak = 6546.2
    for lvl in LEVELS:
        n = 2 ** lvl
        P = 10 * 4 ** lvl
        st = [torch.cuda.current_stream().cuda_stream]
        # ---- xyz input layer 3 -> 64 (narrow kernels, fp32): forward and wgrad (+db)
        B = int(max(1, -(-TARGET_BYTES // (4.0 * (3 + 64) * P))))
        try:
            plan = get_plan(_lib.PLAN_HEXCONV, lvl, 1, 'average', 'cuda')
            w = torch.randn(64, 3, 7, device='cuda') * 0.1
            bias = torch.zeros(64, device='cuda')
            packed = torch.empty(L.gin_hexconv_packed_bytes(3, 64), dtype=torch.uint8, device='cuda')
            _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), 3, 64, st[0]))
            x = torch.randn(B, 3, 5 * n, 2 * n, device='cuda')
            y = torch.empty(B * P, 64, device='cuda')
            dW, db = torch.empty(64, 3, 7, device='cuda'), torch.empty(64, device='cuda')
            ws = torch.empty(L.gin_hexconv_wgrad_ws_bytes(3, 64), dtype=torch.uint8, device='cuda')
            sb, sc, sp = x.stride(0), x.stride(1), x.stride(3)
            us = timed(lambda: _lib.check(L.gin_hexconv_fwd(plan.host_ptr, plan.dev_ptr, x.data_ptr(), sb, sp, sc, packed.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                                            B, 3, 64, _lib.IMPL_AUTO, st[0])), st)
            mb = 4.0 * (3 + 64) * P * B / 1e6
            rows.append('| hex-conv 3→64 fwd (narrow) | %d | 3→64 | %d | %.0f | %.1f | %.0f | %.2f |' % (lvl, B, mb, us, mb / us * 1e3, mb / us * 1e3 / peak))
            us = timed(lambda: _lib.check(L.gin_hexconv_wgrad(plan.host_ptr, plan.dev_ptr, x.data_ptr(), sb, sp, sc, y.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(),
                                                              B, 3, 64, _lib.IMPL_AUTO, st[0])), st)
            rows.append('| hex-conv 3→64 wgrad+db (narrow) | %d | 3→64 | %d | %.0f | %.1f | %.0f | %.2f |' % (lvl, B, mb, us, mb / us * 1e3, mb / us * 1e3 / peak))
            del x, y
        except Exception as e:
            rows.append('| hex-conv 3→64 | %d | failed: %s |' % (lvl, str(e)[:80]))
        torch.cuda.empty_cache()
        # ---- 64 -> 3 hex-conv (fp32 CUDA-core gather-GEMM: the carrier for channel counts that are not multiples of 64)
        B = int(max(1, -(-TARGET_BYTES // (4.0 * (3 + 64) * P))))
        try:
            plan = get_plan(_lib.PLAN_HEXCONV, lvl, 1, 'average', 'cuda')
            w = torch.randn(3, 64, 7, device='cuda') * 0.1
            bias = torch.zeros(3, device='cuda')
            packed = torch.empty(L.gin_hexconv_packed_bytes(64, 3), dtype=torch.uint8, device='cuda')
            _lib.check(L.gin_hexconv_pack_weights(w.data_ptr(), packed.data_ptr(), 64, 3, st[0]))
            x = torch.randn(B * P, 64, device='cuda')
            y = torch.empty(B * P, 3, device='cuda')
            us = timed(lambda: _lib.check(L.gin_hexconv_fwd(plan.host_ptr, plan.dev_ptr, x.data_ptr(), P * 64, 64, 1, packed.data_ptr(), bias.data_ptr(), y.data_ptr(),
                                                            B, 64, 3, _lib.IMPL_AUTO, st[0])), st)
            mb = 4.0 * (3 + 64) * P * B / 1e6
            rows.append('| hex-conv 64→3 fwd (fp32 gather-GEMM) | %d | 64→3 | %d | %.0f | %.1f | %.0f | %.2f |' % (lvl, B, mb, us, mb / us * 1e3, mb / us * 1e3 / peak))
            del x, y
        except Exception as e:
            rows.append('| hex-conv 64→3 | %d | failed: %s |' % (lvl, str(e)[:80]))
        torch.cuda.empty_cache()
        # ---- upsample level lvl -> lvl+1 (fp32 module form, its backward, and the fused operand-copy form)
        if lvl >= 7:
            continue
        for C in (64, 256):
            Pc, Pf = P, 4 * P
            B = int(max(1, -(-TARGET_BYTES // (4.0 * C * (Pc + Pf)))))
            try:
                up = get_plan(_lib.PLAN_UPSAMPLE, lvl, 1, 'average', 'cuda')
                xc = torch.randn(B * Pc, C, device='cuda')
                yf = torch.empty(B * Pf, C, device='cuda')
                mb = 4.0 * C * (Pc + Pf) * B / 1e6
                us = timed(lambda: _lib.check(L.gin_upsample_fwd(up.host_ptr, up.dev_ptr, xc.data_ptr(), yf.data_ptr(), B, C, st[0])), st)
                rows.append('| upsample fwd fp32 | %d→%d | %d | %d | %.0f | %.1f | %.0f | %.2f |' % (lvl, lvl + 1, C, B, mb, us, mb / us * 1e3, mb / us * 1e3 / peak))
                us = timed(lambda: _lib.check(L.gin_upsample_bwd(up.host_ptr, up.dev_ptr, yf.data_ptr(), xc.data_ptr(), B, C, st[0])), st)
                rows.append('| upsample bwd fp32 | %d→%d | %d | %d | %.0f | %.1f | %.0f | %.2f |' % (lvl, lvl + 1, C, B, mb, us, mb / us * 1e3, mb / us * 1e3 / peak))
                ob = torch.empty(B * Pf + 2 * B, C, dtype=torch.bfloat16, device='cuda')
                ow = torch.empty(B * Pf + 2 * B, C, dtype=torch.bfloat16, device='cuda')
                dual = bool(L.gin_forward_operand_is_fp16())
                mb2 = C * B * (4.0 * Pc + (4.0 if dual else 2.0) * Pf) / 1e6
                us = timed(lambda: _lib.check(L.gin_upsample_bf16(up.host_ptr, up.dev_ptr, xc.data_ptr(), 1, ob.data_ptr(), ow.data_ptr() if dual else None, B, C, st[0])), st)
                rows.append('| upsample into operand copies (fp32 in, %s out) | %d→%d | %d | %d | %.0f | %.1f | %.0f | %.2f |' % (
                    'fp16 + bf16' if dual else 'bf16', lvl, lvl + 1, C, B, mb2, us, mb2 / us * 1e3, mb2 / us * 1e3 / peak))
                del xc, yf, ob, ow
            except Exception as e:
                rows.append('| upsample | %d | failed: %s |' % (lvl, str(e)[:80]))
            torch.cuda.empty_cache()
    with open(out, 'w') as f:
        f.write('\n'.join(rows) + '\n')
    print('\n'.join(rows))


if __name__ == '__main__':
    main()
