"""Generate tests/golden/layersweep1111_seed0.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists:

    python oracle/gen_golden_layersweep.py

BASELINE.json configs[2]: `Whitebox.layerwise_contrastive_ebp(mode='percentile',
percentile=20, k_layer=k)` (reference whitebox.py:584-644) for EVERY hook firing k of one triplet, on the reference's own
ResNet class with layers [1,1,1,1] and this repo's seeded synthetic weights; inputs as in tests/helpers.py:golden_inputs
(smooth probe, classifier rows (1/2500) * encodings of two more smooth images).  Stored per firing: sum, maximum and arg-max
of the map; the full 112x112 maps of every fourth firing; the layer names.
"""
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, 'shim'), '/root/reference/python', ROOT]
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from xfr.models.whitebox import Whitebox, WhiteboxSTResnet  # noqa: E402  (the reference)
from xfr.models.resnet import ResNet, Bottleneck  # noqa: E402  (the reference)
from xfr_b200 import synth  # noqa: E402


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    layers = (1, 1, 1, 1)
    net = ResNet(Bottleneck, list(layers), mode='encode', num_classes=2)
    net.load_state_dict(synth.stresnet_state_dict(0, layers, 2))
    net.eval()
    wb = Whitebox(WhiteboxSTResnet(net))
    imgs = synth.smooth_probes(3, seed=1)
    with torch.no_grad():
        x_mate, x_nonmate = wb.encode(imgs[1:2]).detach(), wb.encode(imgs[2:3]).detach()
    wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_nonmate)
    probe = imgs[0:1]
    P0 = torch.zeros(1, 2)
    P0[0, 0] = 1
    wb.ebp(probe, P0)
    n = len(wb.P)
    G = {'n_firings': np.array(n), 'names': np.array([str(s) for s in wb.P_layername])}
    sums, maxs, args, full_k, full = [], [], [], [], []
    for k in range(n):
        m = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode='percentile', percentile=20)
        m = np.asarray(m, dtype=np.float32)
        sums.append(float(m.astype(np.float64).sum()))
        maxs.append(float(m.max()))
        args.append(int(m.argmax()))
        if k % 4 == 1 or k == n - 2:
            full_k.append(k)
            full.append(m)
    # the other modes of the operator (whitebox.py:606-642) at two live firings, layerwise_ebp's default 'argmax' mode
    # (whitebox.py:561-581) and other truncation percentiles of truncated_contrastive_ebp (whitebox.py:529-558)
    for mode in ('copy', 'mean', 'product', 'argmax', 'argmax_product', 'percentile_argmax'):
        for k in (7, 29):
            G['lc_%s_%d' % (mode, k)] = np.asarray(wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode=mode, percentile=20),
                                                   dtype=np.float32)
    for k in (3, 15, 29, 46, -2):
        G['lw_argmax_%d' % (k % n)] = np.asarray(wb.layerwise_ebp(probe, k_layer=k, mode='argmax', k_poschannel=0, mwp=True),
                                                 dtype=np.float32)
    for pct in (0, 50, 80, 100):
        G['trunc_pct%d' % pct] = np.asarray(wb.truncated_contrastive_ebp(probe, 0, 1, percentile=pct), dtype=np.float32)
    G.update(map_sum=np.array(sums), map_max=np.array(maxs), map_argmax=np.array(args), full_k=np.array(full_k),
             full=np.stack(full))
    out = os.path.join(ROOT, 'tests', 'golden', 'layersweep1111_seed0.npz')
    np.savez_compressed(out, **G)
    print('wrote %s (%d firings, %d full maps, %.0f KB)' % (out, n, len(full_k), os.path.getsize(out) / 1024))
    print('nonzero maps at firings', [k for k in range(n) if maxs[k] > 0])


if __name__ == '__main__':
    main()
