"""Further ResNet-101 golden vectors from the UNMODIFIED reference (TEST INFRASTRUCTURE; build container only):

    python oracle/gen_golden_r101_extra.py [--only wellcond|subtree|bighead]

  stresnet101_wellcond_seed0.npz
      The seeded random-weight network encodes every image almost identically (cos(mate, non-mate) = 0.9999), so the
      contrastive map of tests/golden/stresnet101_seed0.npz subtracts two maps that agree to three digits and cannot tell a
      0.5 % kernel error from a 50 % one.  Here the classifier rows are WELL SEPARATED: the mate row is the (1/2500-scaled)
      encoding of a mate image, the non-mate row has the same norm and cos(mate, non-mate) = 0.3 (the mate direction mixed
      with a seeded random direction orthogonal to it) - like the bundled real VGGFace2 triplet (cos -0.04 / 0.76).
      contrastive_ebp / truncated_contrastive_ebp / ebp of the reference in two subtree modes, two probes.
  stresnet101_subtree_seed0.npz
      Whitebox.weighted_subtree_ebp on the full [3,4,23,3] net with the settings of the evaluation flow
      (generate_whitebox_saliency.py:122-205: ctor mode 'norelu', subtree_mode 'all', topk 32, no mated-similarity gating,
      unit-norm rows): ~760 hooked ebp() calls, ~10 minutes on 8 cores.  Pins the 378-firing case incl. np.argsort ties.
  stresnet101_bighead_seed0.npz
      The network's own hooked 65,359-class fc2 head (33 M parameters): mean-EBP prior of the blackbox
      (uniform prior over all classes) and the demo's contrastive_ebp(x, 0, 100).
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
from xfr_b200 import synth  # noqa: E402
from gen_golden import ref_net, onehot  # noqa: E402

LAYERS = (3, 4, 23, 3)


def wellcond_rows(x_mate, cos=0.3, seed=11):
    """(mate row, non-mate row): |non-mate| = |mate|, cos(mate, non-mate) = `cos`.  Also used by the tests."""
    m = x_mate.reshape(-1).double()
    g = torch.Generator().manual_seed(seed)
    r = torch.randn(m.numel(), generator=g, dtype=torch.float64)
    mh = m / m.norm()
    r = r - (r @ mh) * mh
    r = r / r.norm()
    n = (cos * mh + (1.0 - cos * cos) ** 0.5 * r) * m.norm()
    return x_mate.reshape(1, -1).float(), n.float().reshape(1, -1)


def run_wellcond(out):
    t0 = time.time()
    imgs = synth.smooth_probes(3, seed=1)
    noise = synth.synthetic_probes(1, seed=2)
    probe, im_mate = imgs[0:1], imgs[1:2]
    G = {}
    wb = Whitebox(WhiteboxSTResnet(ref_net(LAYERS, 0)))
    with torch.no_grad():
        x_mate = wb.net.encode(im_mate).detach()
    xm, xn = wellcond_rows((1.0 / 2500.0) * x_mate)
    G['row_mate'], G['row_nonmate'] = xm.numpy(), xn.numpy()
    G['cos'] = np.array(float((xm @ xn.t()) / (xm.norm() * xn.norm())))
    for mode in ('affineonly_with_prior', 'all'):
        tag = {'affineonly_with_prior': 'awp'}.get(mode, mode)
        wb = Whitebox(WhiteboxSTResnet(ref_net(LAYERS, 0)), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier(xm, xn)
        for pname, x in (('smooth', probe), ('noise', noise)):
            G['cebp_%s_%s' % (tag, pname)] = wb.contrastive_ebp(x, 0, 1)
            G['tcebp20_%s_%s' % (tag, pname)] = wb.truncated_contrastive_ebp(x, 0, 1, percentile=20)
            G['ebp1_%s_%s' % (tag, pname)] = wb.ebp(x, onehot(2, 1))
            G['ebp1_mwp_%s_%s' % (tag, pname)] = wb.ebp(x, onehot(2, 1), mwp=True)
        print('  wellcond %s done (%.0fs)' % (mode, time.time() - t0), flush=True)
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays, %.0f KB, %.0fs)' % (out, len(G), os.path.getsize(out) / 1024, time.time() - t0))


def run_subtree(out):
    t0 = time.time()
    imgs = synth.smooth_probes(3, seed=1)
    probe, im_mate, im_non = imgs[0:1], imgs[1:2], imgs[2:3]
    G = {}
    wb = Whitebox(WhiteboxSTResnet(ref_net(LAYERS, 0)), ebp_subtree_mode='norelu')
    with torch.no_grad():
        x_mate = wb.net.encode(im_mate).detach()
    # unit-norm rows as the evaluation flow sets them; the non-mate row well separated from the mate row (cos 0.3)
    xm, xn = wellcond_rows(x_mate / torch.norm(x_mate))
    wb.net.set_triplet_classifier(xm, xn)
    G['row_mate'], G['row_nonmate'] = xm.numpy(), xn.numpy()
    smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(probe, 0, 1, topk=32, verbose=False, do_max_subtree=False,
                                                        do_mated_similarity_gating=False, subtree_mode='all')
    G['ws_smap'] = np.asarray(smap)
    G['ws_scores'] = np.asarray(P_sub, dtype=np.float64)
    G['ws_k'] = np.asarray([int(k) for k in k_sub])
    G['ws_maps'] = np.stack([np.asarray(p) for p in P_img]).astype(np.float16)      # 32 maps: kept in half precision (size)
    G['ws_maps_max'] = np.array([float(np.max(p)) for p in P_img])
    G['ws_maps_argmax'] = np.array([int(np.argmax(p)) for p in P_img])
    G['n_firings'] = np.array(len(wb.P_layername))
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays, %.0f KB, %.0fs)' % (out, len(G), os.path.getsize(out) / 1024, time.time() - t0))


def run_bighead(out, num_classes=65359):
    """The STR network's own 65,359-class fc2 (no triplet classifier: hooked, W+ in the backward): STRise.mean_ebp_prior
    (blackbox.py:280-294: ebp with a uniform prior over all classes) and the demo's contrastive_ebp(x, 0, 100)
    (demo/test_whitebox.py:92-99)."""
    t0 = time.time()
    probe = synth.smooth_probes(3, seed=1)[0:1]
    G = {'num_classes': np.array(num_classes)}
    wb = Whitebox(WhiteboxSTResnet(ref_net(LAYERS, 0, num_classes)))
    assert wb.net.num_classes() == num_classes
    P = torch.ones((1, wb.net.num_classes()))
    G['mean_ebp'] = wb.ebp(probe, P)
    G['mean_ebp_mwp'] = wb.ebp(probe, P, mwp=True)
    G['n_firings'] = np.array(len(wb.P_layername))
    G['cebp_0_100'] = wb.contrastive_ebp(probe, k_poschannel=0, k_negchannel=100)
    P1 = torch.zeros((1, num_classes))
    P1[0][100] = 1.0
    G['ebp_100_mwp'] = wb.ebp(probe, P1, mwp=True)
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays, %.0f KB, %.0fs)' % (out, len(G), os.path.getsize(out) / 1024, time.time() - t0))


if __name__ == '__main__':
    ap = argparse.ArgumentParser()
    ap.add_argument('--only', default=None)
    a = ap.parse_args()
    torch.set_num_threads(os.cpu_count())
    gold = os.path.join(ROOT, 'tests', 'golden')
    if a.only in (None, 'wellcond'):
        run_wellcond(os.path.join(gold, 'stresnet101_wellcond_seed0.npz'))
    if a.only in (None, 'subtree'):
        run_subtree(os.path.join(gold, 'stresnet101_subtree_seed0.npz'))
    if a.only in (None, 'bighead'):
        run_bighead(os.path.join(gold, 'stresnet101_bighead_seed0.npz'))
