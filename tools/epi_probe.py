"""GPU probe: time the fused dgrad + JOIN epilogue kernel at the ResNet-101 layer3 shape with parts of the epilogue
switched off (debug bits in `hooks` >> 8: 1 no global loads, 2 no stores, 4 no hook math)."""
import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from xfr_b200.kernels import CudaBackend
from xfr_b200.packing import gemm_planes
be = CudaBackend('cuda:0', impl='tf32x3')
import os
J, N, H, Cin, Cout = 256, 128, 14, 1024, int(os.environ.get("PROBE_K", "256"))
g = torch.Generator().manual_seed(0)
dev = 'cuda'
y1 = torch.rand(J, H, H, Cout, generator=g).to(dev)
class L: pass
L.Bd, L.cin, L.R = gemm_planes(torch.rand(Cin, Cout, generator=g), 'tf32x3').to(dev), Cin, 1
g_res = torch.rand(J, H, H, Cin, generator=g).to(dev)
out, o3, xr3 = (torch.rand(N, H, H, Cin, generator=g).to(dev) for _ in range(3))
bn3 = torch.rand(4, Cin, generator=g).to(dev)
g_out = torch.empty(J, H, H, Cin, device=dev); y3 = torch.empty_like(g_out)
flush = torch.empty(64 * 1024 * 1024, device=dev)
for dbg in (0, 1, 2, 4, 3, 5, 6, 7):
    ts = []
    for it in range(6):
        flush.zero_()
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        be.dgrad_join(y1, L, g_res, out, o3, xr3, bn3, None, 2 | (dbg << 8), 0, g_out, y3)
        b.record(); torch.cuda.synchronize()
        ts.append(a.elapsed_time(b) * 1000)
    print('dbg %d (%s%s%s): %.1f us' % (dbg, 'noload ' if dbg & 1 else '', 'nostore ' if dbg & 2 else '', 'nomath' if dbg & 4 else '', min(ts[1:])))
