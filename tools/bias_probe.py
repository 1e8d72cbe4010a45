"""GPU probe: signed mean relative error of the split-TF32 GEMM plans on all-positive operands (bias detector)."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from xfr_b200.kernels import CudaBackend
from xfr_b200.packing import gemm_planes
g = torch.Generator().manual_seed(0)
J, H, Cin, Cout = 64, 14, 256, 512          # GEMM M = J*H*H, N = Cin, K = Cout
y = torch.rand(J, H, H, Cout, generator=g) + 0.5
Bd = torch.rand(Cin, Cout, generator=g) + 0.5
want = (y.double().reshape(-1, Cout) @ Bd.double().t())
for impl in ('fp32', 'tf32x3full', 'tf32x3', 'tf32'):
    be = CudaBackend('cuda:0', impl=impl)
    class L: pass
    L.Bd, L.cin, L.R = gemm_planes(Bd, impl).cuda(), Cin, 1
    out = torch.zeros(J, H, H, Cin, device='cuda')
    be.dgrad_plain(y.cuda(), L, out)
    torch.cuda.synchronize()
    rel = (out.cpu().double().reshape(-1, Cin) - want) / want
    print('%-11s mean signed rel err %+.3e   max |rel| %.3e' % (impl, float(rel.mean()), float(rel.abs().max())))
