"""Timing of the inpainting-game scoring path (xfr_b200/inpaintgame.py:classified_as_inpainted_twin) on the STR ResNet-101:
one job = 101 percent-density blends of a 224x224 probe -> 101 forwards -> twin classification.  Not the contract bench."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'tests'))
from inpaintgame_fixture import PCT_DENSITY, scoring_fixture  # noqa: E402
from xfr_b200 import inpaintgame as IG  # noqa: E402
from xfr_b200 import synth, whitebox  # noqa: E402

dev = torch.device('cuda:0')
sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
snet = whitebox.Whitebox(whitebox.WhiteboxSTResnet(sd))
F = scoring_fixture()
gal_o, gal_p = snet.embeddings([F['orig']]), snet.embeddings([F['inp']])
run = lambda: IG.classified_as_inpainted_twin(snet, F['orig'], F['inp'], gal_o, gal_p, F['smap'], 'percent-density',
                                              percentiles=PCT_DENSITY, seed=0)
for _ in range(3):
    cls, pg, pr = run()
torch.cuda.synchronize()
reps = 10
t0 = time.time()
for _ in range(reps):
    run()
torch.cuda.synchronize()
job_ms = (time.time() - t0) / reps * 1e3
t0 = time.time()
for _ in range(reps):
    value, thr = IG.mask_value_map(F['smap'], 'percent-density', PCT_DENSITY, seed=0)
host_ms = (time.time() - t0) / reps * 1e3
eng, be = snet.net.engine(), snet.net.engine().be
up = lambda a: torch.from_numpy(np.ascontiguousarray(a, dtype=np.float64)).to(dev)
o, p, v, t = up(F['orig']), up(F['inp']), up(value), up(thr)
blends = torch.empty(101, 224, 224, 3, device=dev)
e = [torch.cuda.Event(True) for _ in range(3)]
tb = tf = 0.0
for _ in range(reps):
    e[0].record(); be.twin_blends(o, p, v, t, None, blends); e[1].record(); snet.net.encode_nhwc(blends); e[2].record()
    torch.cuda.synchronize()
    tb += e[0].elapsed_time(e[1]); tf += e[1].elapsed_time(e[2])
print('scoring, ResNet-101, 101 blends per job: %.1f ms per job end to end (%.1f jobs/s): host value map %.1f ms, '
      'xfrb_twin_blends %.3f ms (%.0f GB/s written), forward sweep of 101 blends %.1f ms (%.0f blends/s); first twin blend at '
      'percentile %d' % (job_ms, 1e3 / job_ms, host_ms, tb / reps, blends.numel() * 4 / (tb / reps) / 1e6, tf / reps,
                         101 / (tf / reps) * 1e3, int(PCT_DENSITY[np.argmax(cls)])))
