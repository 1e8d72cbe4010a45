"""A/B timing of the unfused block-boundary kernel (csrc/stages.cu: join_kernel, one thread per SAMPLE element walking the
gradient-row groups, against join_rows_kernel, one thread per gradient-row element) on the four ResNet-101 boundary shapes.
Run once per variant (the switch is read once per process):  XFRB_JOIN=0|1 python tools/join_probe.py [N] [G]
Prints per-shape microseconds, algorithmic GB/s and a checksum of both outputs (the variants must agree bit for bit)."""
import os
import sys
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfr_b200.kernels import CudaBackend  # noqa: E402

N = int(sys.argv[1]) if len(sys.argv) > 1 else 128
G = int(sys.argv[2]) if len(sys.argv) > 2 else 2
J = G * N
dev = torch.device('cuda:0')
be = CudaBackend(dev, impl='tf32x3')
gen = torch.Generator(device=dev).manual_seed(0)
rnd = lambda *s: torch.randn(*s, device=dev, generator=gen)
# (H, C, up/k, residual channels of the low-resolution shortcut gradient)
shapes = [(7, 2048, 1, 0), (14, 1024, 2, 1024), (28, 512, 2, 512), (56, 256, 2, 256)]
tot = 0.0
for H, C, up, cr in shapes:
    out, o3, xr3 = rnd(N, H, H, C), rnd(N, H, H, C), rnd(N, H, H, C).abs()
    bn = torch.stack([rnd(C).abs() + 0.1, rnd(C) * 0.1, rnd(C).abs() + 0.1, rnd(C) * 0.1]).contiguous()
    zmain = rnd(J, H // up, H // up, C).abs()
    gres = rnd(J, H // up, H // up, cr).abs() if cr else None
    g_out, y3 = torch.empty(J, H, H, C, device=dev), torch.empty(J, H, H, C, device=dev)
    hooks = 1 if up == 1 else 3
    run = lambda: be.join(zmain, up, gres, up, out, o3, xr3, bn, None, hooks, 0, g_out, y3)
    for _ in range(3):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(True), torch.cuda.Event(True)
    reps = 20
    e0.record()
    for _ in range(reps):
        run()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) / reps * 1e3
    tot += us
    alg = 4 * (3 * out.numel() + zmain.numel() + (gres.numel() if cr else 0) + 2 * g_out.numel())
    print('XFRB_JOIN=%s N=%d G=%d H=%d C=%d up=%d: %.1f us, %.0f GB/s (saved tensors counted once), checksum %.9e %.9e'
          % (os.environ.get('XFRB_JOIN', '0'), N, G, H, C, up, us, alg / us / 1e3, g_out.double().sum().item(),
             y3.double().sum().item()))
print('XFRB_JOIN=%s total %.1f us' % (os.environ.get('XFRB_JOIN', '0'), tot))
