"""CPU estimate of what the opt-in hybrid plans of kernels.HYBRID_IMPLS cost in parity, before GPU time is spent on them: the
torch emulation of the kernel set (tests/emul_backend.py) against the reference's outputs (tests/golden).
  'tf32x2f'   two-pass forward dual convs: the signed weights rounded to TF32 (hi plane), activations exact
  'tf32x3b1'  ONE TF32 pass in the W+ dgrads: activation operand truncated to TF32 (as the tensor core does), hi weight plane
Prints max-abs and max-abs / max(ref) per map for the default and both plans."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, 'tests')]
from emul_backend import EmulBackend  # noqa: E402
from helpers import L101, L1111, golden, golden_inputs, rel_err  # noqa: E402
from xfr_b200 import synth  # noqa: E402
from xfr_b200.engine import StResnetEngine  # noqa: E402

for layers in (L1111, L101):
    G = golden(layers)
    x, W2, _ = golden_inputs(G)
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    for plan, kw in (('tf32x3 (default)', {}), ('tf32x2f', {'fwd_two_pass': True}), ('tf32x3b1', {'bwd_single_pass': True})):
        eng = StResnetEngine(synth.stresnet_state_dict(0, layers, 2), EmulBackend(impl_name='tf32x3', **kw), layers)
        s = eng.ebp(x, P1, W2).clone().numpy()
        c = eng.contrastive(x, W2).clone().numpy()
        t = eng.contrastive(x, W2, percentile=20).clone().numpy()
        rows = []
        for name, got, key in (('ebp', s, 'ebp_awp_%s'), ('contrastive', c, 'cebp_awp_%s'), ('truncated', t, 'tcebp20_awp_%s')):
            for i, p in enumerate(('smooth', 'noise')):
                if key % p in G.files:
                    ref = G[key % p]
                    rows.append('%s/%s max-abs %.2e rel %.2e' % (name, p, np.abs(got[i] - ref).max(), rel_err(got[i], ref)))
        print('layers %s  plan %s:' % (layers, plan))
        for r in rows:
            print('    ' + r)
