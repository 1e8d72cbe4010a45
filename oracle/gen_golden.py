"""Generate tests/golden/*.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists:

    python oracle/gen_golden.py [--only NAME]

It imports `xfr.models.whitebox` / `xfr.models.resnet` from /root/reference/python
(missing third-party imports satisfied by oracle/shim), loads this repo's seeded
synthetic state_dict (xfr_b200/synth.py) into the reference's own ResNet class and
records what the reference's hook-based `Whitebox` returns.  The weights do not
travel (160 MB); the seed does, so the GPU box rebuilds the identical state_dict.

Golden sets
  stresnet101_seed0   STR ResNet-101 [3,4,23,3]  (reference resnet.py:277), triplet head
  stresnet1111_seed0  same class with layers [1,1,1,1]: cheap enough to pin the
                      layerwise / weighted-subtree paths (756 ebp() calls on the 101)
"""
import argparse
import os
import sys
import time
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


def ref_net(layers, seed, num_classes=2):
    net = ResNet(Bottleneck, list(layers), mode='encode', num_classes=num_classes)
    net.load_state_dict(synth.stresnet_state_dict(seed, layers, num_classes))
    net.eval()
    return net


def onehot(c, k):
    p = torch.zeros(1, c)
    p[0, k] = 1.0
    return p


def fingerprints(P):
    sums = np.array([float(p.double().sum()) for p in P])
    maxs = np.array([float(p.max()) for p in P])
    return sums, maxs


def run_set(layers, seed, out, do_subtree):
    t0 = time.time()
    torch.manual_seed(0)
    imgs = synth.smooth_probes(3, seed=1)
    noise = synth.synthetic_probes(1, seed=2)
    probe, im_mate, im_non = imgs[0:1], imgs[1:2], imgs[2:3]
    G = {}
    net = ref_net(layers, seed)
    wb = Whitebox(WhiteboxSTResnet(net))
    with torch.no_grad():
        x_mate = wb.net.encode(im_mate).detach()
        x_non = wb.net.encode(im_non).detach()
    G['enc_mate'] = x_mate.numpy()
    G['enc_nonmate'] = x_non.numpy()
    for mode in ('affineonly_with_prior', 'all', 'affineonly', 'norelu'):
        tag = {'affineonly_with_prior': 'awp'}.get(mode, mode)
        wb = Whitebox(WhiteboxSTResnet(ref_net(layers, seed)), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_non)
        for pname, x in (('smooth', probe), ('noise', noise)):
            G['ebp_mwp_%s_%s' % (tag, pname)] = wb.ebp(x, onehot(2, 0), mwp=True)
            sums, maxs = fingerprints(wb.P)
            G['Psum_%s_%s' % (tag, pname)] = sums
            G['Pmax_%s_%s' % (tag, pname)] = maxs
            if pname == 'smooth' and mode == 'affineonly_with_prior':
                names = list(wb.P_layername)
                G['P_kinds'] = np.array([n.split('(')[0] for n in names])
                G['P_numel'] = np.array([p.numel() for p in wb.P])
            G['ebp_%s_%s' % (tag, pname)] = wb.ebp(x, onehot(2, 0))
            G['cebp_%s_%s' % (tag, pname)] = wb.contrastive_ebp(x, 0, 1)
            G['tcebp20_%s_%s' % (tag, pname)] = wb.truncated_contrastive_ebp(x, 0, 1, percentile=20)
        G['meanebp_%s_smooth' % tag] = wb.ebp(probe, torch.ones(1, 2))
        print('  mode %s done (%.0fs)' % (mode, time.time() - t0))
    # hooked (non-triplet) classifier head: the network's own fc2 takes part with W+
    wb = Whitebox(WhiteboxSTResnet(ref_net(layers, seed)))
    G['ebp_mwp_awp_fc2head'] = wb.ebp(probe, onehot(2, 1), mwp=True)
    G['cebp_awp_fc2head'] = wb.contrastive_ebp(probe, 0, 1)
    # with_bias / uint8 post-processing (ebp_version 11 / 5)
    wb = Whitebox(WhiteboxSTResnet(ref_net(layers, seed)), ebp_version=11)
    wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_non)
    G['ebp_mwp_awp_withbias'] = wb.ebp(probe, onehot(2, 0), mwp=True)
    G['cebp_v11_u8'] = wb.contrastive_ebp(probe, 0, 1)
    # layerwise (elementwise prior) at a handful of firing indices, three modes
    ks = [1, 2, 3, 4, 5, 6, 7, 8, 12, 14, 15, 16] if do_subtree else [1, 2, 3, 4, 5, 14, 27, 28, 100, 200, 372, 374]
    for mode in ('affineonly_with_prior', 'all', 'norelu'):
        tag = {'affineonly_with_prior': 'awp'}.get(mode, mode)
        wb = Whitebox(WhiteboxSTResnet(ref_net(layers, seed)), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_non)
        wb.ebp(probe, onehot(2, 0), mwp=True)
        Pm = [p.detach().clone() for p in wb.P]
        ks_ok = [k for k in ks if k < len(Pm) - 1]
        els = [int(torch.argmax(Pm[k].flatten())) for k in ks_ok]
        G['lw_k'] = np.array(ks_ok)
        G['lw_el_%s' % tag] = np.array(els)
        G['lw_%s' % tag] = np.stack([wb.layerwise_ebp(probe, k_layer=k, k_element=e, mode='elementwise',
                                                     k_poschannel=0, mwp=True) for k, e in zip(ks_ok, els)])
        print('  layerwise %s done (%.0fs)' % (mode, time.time() - t0))
    if do_subtree:
        for (tag, ctor_mode, sub_mode, gating, mx, ver, scale) in (
                ('ws_demo', 'affineonly_with_prior', 'all', True, False, 5, 1.0 / 2500.0),   # demo/test_whitebox.py:173-199
                ('ws_eval', 'norelu', 'all', False, False, None, 1.0),                        # generate_whitebox_saliency.py:143
                ('ws_norelu_max', 'norelu', 'norelu', True, True, None, 1.0 / 2500.0)):
            wb = Whitebox(WhiteboxSTResnet(ref_net(layers, seed)), ebp_version=ver, ebp_subtree_mode=ctor_mode)
            xm = x_mate / torch.norm(x_mate) if scale == 1.0 else scale * x_mate
            xn = x_non / torch.norm(x_non) if scale == 1.0 else scale * x_non
            wb.net.set_triplet_classifier(xm, xn)
            smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(
                probe, 0, 1, topk=8, verbose=False, do_max_subtree=mx,
                do_mated_similarity_gating=gating, subtree_mode=sub_mode)
            G[tag + '_smap'] = np.asarray(smap)
            G[tag + '_scores'] = np.asarray(P_sub, dtype=np.float64)
            G[tag + '_k'] = np.asarray([int(k) for k in k_sub])
            G[tag + '_maps'] = np.stack([np.asarray(p) for p in P_img])
            print('  subtree %s done (%.0fs)' % (tag, time.time() - t0))
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays, %.0f KB, %.0fs)' % (out, len(G), os.path.getsize(out) / 1024, time.time() - t0))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    gold = os.path.join(ROOT, 'tests', 'golden')
    os.makedirs(gold, exist_ok=True)
    if a.only in (None, 'stresnet1111_seed0'):
        run_set((1, 1, 1, 1), 0, os.path.join(gold, 'stresnet1111_seed0.npz'), do_subtree=True)
    if a.only in (None, 'stresnet101_seed0'):
        run_set((3, 4, 23, 3), 0, os.path.join(gold, 'stresnet101_seed0.npz'), do_subtree=False)
