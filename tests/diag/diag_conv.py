import sys, os, torch, numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from oracle import icocnn_ref
from geniconet_b200.ico_conv import IcoConvS2S
cin, cout, stride, level, B = [int(a) for a in sys.argv[1:6]]
torch.manual_seed(2)
from geniconet_b200 import _lib
FWD = _lib.forward_operand_dtype()                       # forward-side operands: fp16 (default) or bf16; gradient side: bf16
ref = icocnn_ref.IcoConvS2S(cin, cout, stride, True, level, 'average')
with torch.no_grad():
    ref.weight.copy_(ref.weight.to(torch.float16).to(torch.bfloat16).float())      # exactly representable in both formats
mod = IcoConvS2S(cin, cout, stride, True, level, 'average', impl='tc').cuda()
mod.load_state_dict(ref.state_dict())
n = 2 ** level
g = torch.Generator().manual_seed(9)
x = torch.randn(B, cin, 5 * n, 2 * n, generator=g).to(torch.float16).to(torch.bfloat16).float()      # exact in both formats (wgrad reads bf16)
xr = x.clone().requires_grad_(True)
yr = ref(xr); gy = torch.randn(yr.shape, generator=g).to(torch.bfloat16).float(); yr.backward(gy)
xc = x.cuda().contiguous(memory_format=torch.channels_last).requires_grad_(True)
yc = mod(xc); yc.backward(gy.cuda()); torch.cuda.synchronize()
for name, a, b in (('fwd', yc.detach().cpu(), yr.detach()), ('dgrad', xc.grad.cpu(), xr.grad), ('wgrad', mod.weight.grad.cpu(), ref.weight.grad)):
    err = (a - b).abs()
    print(name, 'max err %.3e rel %.3e' % (err.max().item(), (err.max() / b.abs().max()).item()))
    if name != 'wgrad' and err.max() / b.abs().max() > 2e-3:
        e = err.amax(1)            # [B,H,W]
        bad = (e > 2e-3 * b.abs().max()).nonzero()
        print('  bad pixels:', len(bad), 'of', e.numel())
        no = a.shape[2] // 5
        import collections
        cnt = collections.Counter((int(bb), int(h) // no, (int(h) % no) // 8, int(w) // 8) for bb, h, w in bad.tolist())
        print('  (sample, chart, rowblock, octet): count ->', sorted(cnt.items())[:40])
    if name == 'wgrad' and err.max() / b.abs().max() > 2e-3:
        e = err.amax((0, 1)); print('  per-tap err', e.tolist())
