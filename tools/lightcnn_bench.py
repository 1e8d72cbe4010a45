"""GPU: throughput of the Light-CNN-29v2 path at the BASELINE config-5 shape (128-probe sweeps of the 512 per GPU):
ebp() in 'affineonly' (demo) and 'affineonly_with_prior' (create_wbnet), contrastive_ebp, device-resident inputs."""
import os
import sys
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from xfr_b200 import synth
from xfr_b200.kernels import CudaBackend
from xfr_b200.lightcnn import LightCNNEngine
dev = torch.device('cuda:0')
eng = LightCNNEngine(synth.lightcnn_state_dict(0, 2), CudaBackend(dev, impl='tf32x3'), device=dev)
N, total = 128, 512
x = synth.lightcnn_probes(N, seed=3, smooth=False).permute(0, 2, 3, 1).contiguous().to(dev)
g = torch.Generator().manual_seed(4)
W2 = torch.randn(N, 2, 256, generator=g).to(dev)
P1 = torch.zeros(N, 2, device=dev)
P1[:, 0] = 1
for name, fn in (("ebp 'affineonly'", lambda: eng.ebp(x, P1, W2, 'affineonly')),
                 ("ebp 'affineonly_with_prior'", lambda: eng.ebp(x, P1, W2, 'affineonly_with_prior')),
                 ("ebp 'all'", lambda: eng.ebp(x, P1, W2, 'all')),
                 ("contrastive_ebp 'affineonly_with_prior'", lambda: eng.contrastive(x, W2))):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    l0 = eng.be.launches
    a.record()
    for _ in range(total // N):
        fn()
    b.record()
    torch.cuda.synchronize()
    ms = a.elapsed_time(b)
    print('%-42s %7.1f maps/s  (%d probes in %.1f ms, %d launches, workspace %.1f GB)'
          % (name, total / (ms / 1e3), total, ms, eng.be.launches - l0, eng.workspace_bytes() / 1e9))
