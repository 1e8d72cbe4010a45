"""Batch-1 contrastive_ebp calls (graph replays) for an ncu launch list:
    ncu --metrics gpu__time_duration.sum --clock-control none --graph-profiling node --csv --log-file out.csv python tools/batch1_profile.py
Without ncu it prints the wall-clock latency per call."""
import sys
import time

import torch

sys.path.insert(0, '.')
from xfr_b200 import synth, whitebox  # noqa: E402

dev = torch.device('cuda:0')
sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
net = whitebox.WhiteboxSTResnet(sd)
wb = whitebox.Whitebox(net)
x = synth.synthetic_probes(1, seed=100).pin_memory()
with torch.no_grad():
    enc = net.encode(synth.synthetic_probes(2, seed=1000).to(dev))
net.set_triplet_classifier(enc[0:1] / 2500.0, enc[1:2] / 2500.0)
n = int(sys.argv[1]) if len(sys.argv) > 1 else 5
ts = []
for i in range(n):
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    wb.contrastive_ebp(x, 0, 1)
    torch.cuda.synchronize()
    ts.append(1e3 * (time.perf_counter() - t0))
print('ms per call:', ' '.join('%.2f' % t for t in ts))
