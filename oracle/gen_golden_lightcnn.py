"""Generate tests/golden/lightcnn29v2_seed0.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Build container only (/root/reference present):   python oracle/gen_golden_lightcnn.py [--check]

Loads the seeded synthetic state_dict of xfr_b200/synth.py (the reference bundles no Light-CNN weights) into the
reference's own `LightCNN_29Layers_v2` (lightcnn.py:293-296), wraps it in the reference's `WhiteboxLightCNN` /
`Whitebox` (whitebox.py:113-159, 261-304) and records what the hook-based implementation returns for the BASELINE
config-5 inputs (U[0,1] 1x128x128 probes).  --check additionally compares oracle/lightcnn_oracle.py firing by firing.
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

from xfr.models.whitebox import Whitebox, WhiteboxLightCNN  # noqa: E402  (the reference)
from xfr.models.lightcnn import LightCNN_29Layers_v2  # noqa: E402  (the reference)
from xfr_b200 import synth  # noqa: E402

NUM_CLASSES = 64        # the network's own (hooked) fc2 for the non-triplet pin; the real net has 80,013


def ref_net(seed=0):
    net = LightCNN_29Layers_v2(num_classes=NUM_CLASSES)
    net.load_state_dict(synth.lightcnn_state_dict(seed, NUM_CLASSES))
    net.eval()
    return net


def main(check):
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    imgs = synth.lightcnn_probes(3, seed=1)
    noise = synth.lightcnn_probes(1, seed=2, smooth=False)
    probe, im_mate, im_non = imgs[0:1], imgs[1:2], imgs[2:3]
    G = {}
    wb = Whitebox(WhiteboxLightCNN(ref_net()))
    with torch.no_grad():
        x_mate = wb.net.encode(im_mate).detach().clone()
        x_non = wb.net.encode(im_non).detach().clone()
    G['enc_mate'], G['enc_nonmate'] = x_mate.numpy(), x_non.numpy()
    P0 = torch.zeros(1, 2)
    P0[0, 0] = 1.0
    if check:
        from oracle import lightcnn_oracle as O
        sd = synth.lightcnn_state_dict(0, NUM_CLASSES)
        W2 = torch.cat((x_mate, x_non), 0)
    for mode in ('affineonly_with_prior', 'all', 'affineonly', 'norelu'):
        tag = {'affineonly_with_prior': 'awp'}.get(mode, mode)
        wb = Whitebox(WhiteboxLightCNN(ref_net()), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier(x_mate, x_non)
        for pname, x in (('smooth', probe), ('noise', noise)):
            G['ebp_mwp_%s_%s' % (tag, pname)] = wb.ebp(x, P0, mwp=True)
            G['Psum_%s_%s' % (tag, pname)] = np.array([float(p.double().sum()) for p in wb.P])
            G['Pmax_%s_%s' % (tag, pname)] = np.array([float(p.max()) for p in wb.P])
            if pname == 'smooth' and mode == 'affineonly_with_prior':
                G['P_kinds'] = np.array([n.split('(')[0] for n in wb.P_layername])
                G['P_numel'] = np.array([p.numel() for p in wb.P])
            if check:
                P, names = O.ebp_mwp(sd, x, P0, W2, mode=mode)
                assert names == [n.split('(')[0] for n in wb.P_layername], (names, wb.P_layername)
                worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(P, wb.P))
                print('check %-22s %-6s %d firings, worst max-abs/max over all P: %.3g' % (mode, pname, len(P), worst))
            G['ebp_%s_%s' % (tag, pname)] = wb.ebp(x, P0)
            G['cebp_%s_%s' % (tag, pname)] = wb.contrastive_ebp(x, 0, 1)
            G['tcebp20_%s_%s' % (tag, pname)] = wb.truncated_contrastive_ebp(x, 0, 1, percentile=20)
        print(mode, 'ebp max %.6e argmax %d | cebp max %.6e argmax %d' % (
            G['ebp_%s_smooth' % tag].max(), G['ebp_%s_smooth' % tag].argmax(), G['cebp_%s_smooth' % tag].max(),
            G['cebp_%s_smooth' % tag].argmax()))
    # the network's own fc2 as the (hooked) classifier: one more leading Linear firing (88)
    wb = Whitebox(WhiteboxLightCNN(ref_net()))
    Pk = torch.zeros(1, NUM_CLASSES)
    Pk[0, 5] = 1.0
    G['ebp_mwp_awp_fc2head'] = wb.ebp(probe, Pk, mwp=True)
    G['Psum_awp_fc2head'] = np.array([float(p.double().sum()) for p in wb.P])
    if check:
        P, names = O.ebp_mwp(sd, probe, Pk, None)
        worst = max(float((a - b).abs().max() / b.abs().max().clamp_min(1e-30)) for a, b in zip(P, wb.P))
        print('check fc2head %d firings worst %.3g' % (len(P), worst))
    G['cebp_awp_fc2head'] = wb.contrastive_ebp(probe, 5, 9)
    # layerwise_ebp (elementwise prior) at a few firings, default mode (whitebox.py:561-581)
    wb = Whitebox(WhiteboxLightCNN(ref_net()))
    wb.net.set_triplet_classifier(x_mate, x_non)
    ks = (3, 12, 30, 53, 60, 77)
    G['lw_k'] = np.array(ks)
    for k in ks:
        G['lw_elem_%d' % k] = wb.layerwise_ebp(probe, k_layer=k, mode='argmax', mwp=True)
    # layerwise_ebp(mode='elementwise'): the reference indexes the flattened [1,C,H,W] MWP (whitebox.py:572-577)
    wb.ebp(probe, P0)
    el = [int(torch.argmax(wb.P[k].flatten())) for k in ks]
    G['lw_el_idx'] = np.array(el)
    for k, e in zip(ks, el):
        G['lw_el_%d' % k] = wb.layerwise_ebp(probe, k_layer=k, mode='elementwise', k_element=e, mwp=True)
    # weighted_subtree_ebp (whitebox.py:647-737) in two configurations
    for tag, kw in (('a', dict(topk=8, do_max_subtree=False, do_mated_similarity_gating=True, subtree_mode='affineonly_with_prior')),
                    ('b', dict(topk=4, do_max_subtree=True, do_mated_similarity_gating=False, subtree_mode='all'))):
        wbs = Whitebox(WhiteboxLightCNN(ref_net()))
        wbs.net.set_triplet_classifier(x_mate, x_non)
        smap, P_img, P_sub, k_sub = wbs.weighted_subtree_ebp(probe, 0, 1, verbose=False, **kw)
        G['ws_%s_smap' % tag] = smap
        G['ws_%s_scores' % tag] = np.array(P_sub, dtype=np.float64)
        G['ws_%s_k' % tag] = np.array(k_sub)
        G['ws_%s_first' % tag] = P_img[-1]
        print('weighted_subtree', tag, 'k', list(k_sub), 'scores', ['%.3g' % v for v in P_sub])
    out = os.path.join(ROOT, 'tests', 'golden', 'lightcnn29v2_seed0.npz')
    np.savez_compressed(out, **G)
    print('wrote', out, os.path.getsize(out) // 1024, 'KB')


if __name__ == '__main__':
    main('--check' in sys.argv)
