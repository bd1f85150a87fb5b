"""Run one hex-conv layer pass a few times (for ncu / timing): python tools/run_layer.py cin cout stride level B pass [iters]"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geniconet_b200 import _lib
from geniconet_b200.ico_conv import IcoConvS2S, get_plan
cin, cout, stride, level, B = [int(a) for a in sys.argv[1:6]]
which = sys.argv[6]
iters = int(sys.argv[7]) if len(sys.argv) > 7 else 5
m = IcoConvS2S(cin, cout, stride, True, level, 'average').cuda()
n = 2 ** level
Pin, Pout = 10 * 4 ** level, 10 * 4 ** level // stride ** 2
x = torch.randn(B, Pin, cin, device='cuda')
dy = torch.randn(B, Pout, cout, device='cuda')
y = torch.empty_like(dy); dx = torch.empty_like(x)
dW = torch.empty(cout, cin, 7, device='cuda'); db = torch.empty(cout, device='cuda')
ws = torch.empty(_lib.lib.gin_hexconv_wgrad_ws_bytes(cin, cout), dtype=torch.uint8, device='cuda')
plan = get_plan(_lib.PLAN_HEXCONV, level, stride, 'average', 'cuda')
packed = m._packed_weights(m.weight)
st = torch.cuda.current_stream().cuda_stream
lvl_out = level - (1 if stride == 2 else 0)
xb = torch.empty(_lib.lib.gin_cast_bf16_bytes(B, level, cin) // 2, dtype=torch.bfloat16, device='cuda')
dyb = torch.empty(_lib.lib.gin_cast_bf16_bytes(B, lvl_out, cout) // 2, dtype=torch.bfloat16, device='cuda')
_lib.check(_lib.lib.gin_cast_bf16(plan.host_ptr, plan.dev_ptr, 0, x.data_ptr(), xb.data_ptr(), B, cin, st))
_lib.check(_lib.lib.gin_cast_bf16(plan.host_ptr, plan.dev_ptr, 1, dy.data_ptr(), dyb.data_ptr(), B, cout, st))
def run():
    if which == 'fwd':
        _lib.check(_lib.lib.gin_hexconv_fwd_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), packed.data_ptr(), m.bias.data_ptr(), y.data_ptr(), B, cin, cout, st))
    elif which == 'dgrad':
        _lib.check(_lib.lib.gin_hexconv_dgrad_bf16(plan.host_ptr, plan.dev_ptr, dyb.data_ptr(), packed.data_ptr(), dx.data_ptr(), B, cin, cout, st))
    else:
        _lib.check(_lib.lib.gin_hexconv_wgrad_bf16(plan.host_ptr, plan.dev_ptr, xb.data_ptr(), dyb.data_ptr(), None, dW.data_ptr(), None, ws.data_ptr(), B, cin, cout, st))
for _ in range(2): run()
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
if iters >= 4:          # GPU time only: replay a captured graph of `iters` calls (the Python/ctypes launch path costs ~20-30 us per call)
    side = torch.cuda.Stream()
    with torch.cuda.stream(side):
        st = side.cuda_stream
        g = torch.cuda.CUDAGraph()
        with torch.cuda.graph(g, stream=side):
            for _ in range(iters): run()
    g.replay(); torch.cuda.synchronize()
    e0.record(); g.replay(); e1.record(); torch.cuda.synchronize()
else:
    e0.record()
    for _ in range(iters): run()
    e1.record(); torch.cuda.synchronize()
print('%s %d->%d s%d L%d B%d: %.1f us/iter' % (which, cin, cout, stride, level, B, e0.elapsed_time(e1) * 1e3 / iters))
