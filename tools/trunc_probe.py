import sys, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from xfr_b200.kernels import CudaBackend
be = CudaBackend('cuda:0', impl='tf32')
vals = [1 + 2**-11, 1 + 2**-11 + 2**-20, 1 + 2**-12, 1 + 2**-10 - 2**-22, -(1 + 2**-11), -(1 + 2**-11 + 2**-20), 1 + 3 * 2**-11, 1 + 3*2**-11 - 2**-21]
J, C = 128, 64
y = torch.zeros(J, 1, 1, C)
for i, v in enumerate(vals):
    y[i, 0, 0, 0] = v
Bd = torch.zeros(64, C)
Bd[0, 0] = 1.0          # out channel 0 = y[., 0] * 1
class L: pass
L.Bd, L.cin, L.R = Bd.cuda(), 64, 1
out = torch.zeros(J, 1, 1, 64, device='cuda')
be.dgrad_plain(y.cuda(), L, out)
torch.cuda.synchronize()
o = out[:, 0, 0, 0].cpu().double()
for i, v in enumerate(vals):
    print('a=1%+.3e (x 2^-10 ulp: %+.4f) -> %+.4f ulp' % (abs(v)-1, (abs(v)-1)*2**10, (abs(float(o[i]))-1)*2**10), 'sign', float(o[i])>0)
# now B side: y = 1, B = vals
y2 = torch.zeros(J, 1, 1, C); y2[:, 0, 0, 0] = 1.0
Bd2 = torch.zeros(64, C)
for i, v in enumerate(vals): Bd2[i, 0] = v
L.Bd = Bd2.cuda()
be.dgrad_plain(y2.cuda(), L, out)
torch.cuda.synchronize()
o = out[0, 0, 0, :8].cpu().double()
for i, v in enumerate(vals):
    print('b=1%+.3e -> %+.4f ulp' % (abs(v)-1, (abs(float(o[i]))-1)*2**10))
