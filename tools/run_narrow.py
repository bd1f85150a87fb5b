"""Run the xyz-layer kernels (Cin = 3) a few times: python tools/run_narrow.py [B] [iters]"""
import sys, os, torch
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from geniconet_b200 import _lib
from geniconet_b200.ico_conv import IcoConvS2S, get_plan
B = int(sys.argv[1]) if len(sys.argv) > 1 else 36
iters = int(sys.argv[2]) if len(sys.argv) > 2 else 10
L = _lib.lib
m = IcoConvS2S(3, 64, 1, True, 5, 'average').cuda()
x = torch.randn(B, 3, 160, 64, device='cuda')
P = 10240
y = torch.empty(B * P, 64, device='cuda'); dy = torch.randn(B * P, 64, device='cuda')
dW = torch.empty(64, 3, 7, device='cuda'); db = torch.empty(64, device='cuda')
ws = torch.empty(L.gin_hexconv_wgrad_ws_bytes(3, 64), dtype=torch.uint8, device='cuda')
plan = get_plan(_lib.PLAN_HEXCONV, 5, 1, 'average', 'cuda')
packed = m._packed_weights(m.weight)
st = torch.cuda.current_stream().cuda_stream
sb, sp, sc = 3 * P, 1, P
def fwd():
    _lib.check(L.gin_hexconv_fwd(plan.host_ptr, plan.dev_ptr, x.data_ptr(), sb, sp, sc, packed.data_ptr(), m.bias.data_ptr(), y.data_ptr(), B, 3, 64, 0, st))
def wgrad():
    _lib.check(L.gin_hexconv_wgrad(plan.host_ptr, plan.dev_ptr, x.data_ptr(), sb, sp, sc, dy.data_ptr(), dW.data_ptr(), db.data_ptr(), ws.data_ptr(), B, 3, 64, 0, st))
for name, fn in (('fwd', fwd), ('wgrad', wgrad)):
    for _ in range(2): fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters): fn()
    e1.record(); torch.cuda.synchronize()
    print('narrow %s 3->64 L5 B%d: %.1f us/iter' % (name, B, e0.elapsed_time(e1) * 1e3 / iters))
