"""Scratch timing helper for gpurun sessions (not the contract bench)."""
import sys, time, torch
sys.path.insert(0, __import__('os').path.dirname(__import__('os').path.dirname(__import__('os').path.abspath(__file__))))
from xfr_b200 import synth
from xfr_b200.engine import StResnetEngine
from xfr_b200.kernels import CudaBackend
impl = sys.argv[1] if len(sys.argv) > 1 else 'fp32'
N = int(sys.argv[2]) if len(sys.argv) > 2 else 32
dev = torch.device('cuda:0')
sd = synth.stresnet_state_dict(0)
eng = StResnetEngine(sd, CudaBackend(dev, impl=impl), device=dev)
x = synth.synthetic_probes(N, seed=1).permute(0, 2, 3, 1).contiguous().to(dev)
g = torch.Generator().manual_seed(5)
W2 = (torch.randn(N, 2, 512, generator=g) * 0.02).to(dev)
for _ in range(2):
    eng.contrastive(x, W2)
torch.cuda.synchronize()
e0, e1, e2 = torch.cuda.Event(True), torch.cuda.Event(True), torch.cuda.Event(True)
reps = 3
tf = tb = 0.0
for _ in range(reps):
    e0.record(); eng.forward(x); e1.record()
    P = eng.priors_contrastive(N, 2, 0, 1)
    P2, _, sums = eng.ebp_backward(P, W2); e2.record()
    torch.cuda.synchronize()
    tf += e0.elapsed_time(e1); tb += e1.elapsed_time(e2)
print('impl %s N %d: fwd %.2f ms  bwd(2N rows) %.2f ms  -> %.1f maps/s ; workspace %.1f GB ; launches/sweep %d'
      % (impl, N, tf / reps, tb / reps, N / ((tf + tb) / reps / 1e3), eng.workspace_bytes() / 1e9, eng.be.launches // (reps + 2)))
