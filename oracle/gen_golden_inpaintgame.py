"""Generate tests/golden/inpaintgame_seed0.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists:

    python oracle/gen_golden_inpaintgame.py

It imports the reference's scoring module python/xfr/inpainting_game/inpainting_game.py and its Whitebox /
WhiteboxSTResnet / ResNet classes (missing third-party imports satisfied by oracle/shim), loads this repo's seeded
synthetic [1,1,1,1] state_dict into the reference's ResNet and records what `create_threshold_masks` and
`classified_as_inpainted_twin` return for seeded inputs (tests/inpaintgame_fixture.py rebuilds the inputs from the seeds on
the GPU box).  Masks are stored as per-percentile populations plus a SHA-256 of the packed bits.
"""
import hashlib
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
from xfr.inpainting_game import inpainting_game as REF  # noqa: E402  (the reference)
from xfr.inpainting_game import generate_whitebox_saliency as REFGEN  # noqa: E402  (the reference)
from xfr_b200 import synth  # noqa: E402

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from inpaintgame_fixture import scoring_fixture, jobs, images, PCT_DENSITY, PCT_PIXELS  # noqa: E402


def mask_digest(masks):
    return hashlib.sha256(np.packbits(masks.astype(bool)).tobytes()).hexdigest()


def fresh_wb(**kw):
    """A new reference ResNet [1,1,1,1] with the seeded weights under a new reference Whitebox (set_triplet_classifier
    replaces net.fc2 and Whitebox.__init__ registers hooks on the module objects: nothing is shared between cases)."""
    layers = (1, 1, 1, 1)
    net = ResNet(Bottleneck, list(layers), mode='encode', num_classes=2)
    net.load_state_dict(synth.stresnet_state_dict(0, layers, 2))
    net.eval()
    return Whitebox(WhiteboxSTResnet(net), **kw)


def main():
    torch.manual_seed(0)
    torch.set_num_threads(os.cpu_count())
    snet = fresh_wb()
    F = scoring_fixture()
    G = {}
    with torch.no_grad():
        gal_o = snet.embeddings([F['orig']])
        gal_p = snet.embeddings([F['inp']])
    G['gal_orig'], G['gal_inp'] = gal_o, gal_p
    # --- masks (inpainting_game.py:12-80)
    for tag, kw in (('density', dict(threshold_method='percent-density', percentiles=PCT_DENSITY, seed=0)),
                    ('density_nozero', dict(threshold_method='percent-density', percentiles=PCT_DENSITY, seed=3,
                                            include_zero_elements=False)),
                    ('pixels', dict(threshold_method='percent-pixels', percentiles=PCT_PIXELS, seed=1)),
                    ('thresholds', dict(threshold_method='thresholds', thresholds=np.array([0.5, 1e-4, 2e-5, 1e-5, 0.0]),
                                        percentiles=None, seed=2))):
        smap = F['smap_sparse'] if tag == 'density_nozero' else F['smap']
        m = REF.create_threshold_masks(smap, **kw)
        G['masks_%s_count' % tag] = m.reshape(m.shape[0], -1).sum(1).astype(np.int64)
        G['masks_%s_sha256' % tag] = np.array(mask_digest(m))
    mb = REF.create_threshold_masks(F['smap'], 'percent-density', percentiles=PCT_DENSITY[::10], seed=0, blur_sigma=4)
    G['masks_blur_sum'] = mb.reshape(mb.shape[0], -1).astype(np.float64).sum(1)
    G['masks_blur_row'] = mb[:, 100, :].astype(np.float64)
    # --- scoring (inpainting_game.py:83-146)
    with torch.no_grad():
        cls, pg, pr = REF.classified_as_inpainted_twin(snet, F['orig'], F['inp'], gal_o, gal_p, F['smap'],
                                                       mask_threshold_method='percent-density', percentiles=PCT_DENSITY, seed=0)
        G['cls'], G['pg_dist'], G['pr_dist'] = cls, pg, pr
        cls, pg, pr = REF.classified_as_inpainted_twin(snet, F['orig'], F['inp'], gal_o, gal_p, F['smap'],
                                                       mask_threshold_method='percent-density', percentiles=PCT_DENSITY[::10],
                                                       seed=0, mask_blur_sigma=4)
        G['cls_blur'], G['pg_dist_blur'], G['pr_dist_blur'] = cls, pg, pr
    # --- job functions (generate_whitebox_saliency.py:81-118, 122-205, 207-215), batch 1 as the reference runs them
    import contextlib
    import io
    cpu = torch.device('cpu')
    with contextlib.redirect_stdout(io.StringIO()):
        for i, (im_mates, im_nonmates, probe_im) in enumerate(jobs()):
            for pct in (None, 20):
                wb = fresh_wb()
                G['job%d_contrastive%s' % (i, '' if pct is None else '_pct%d' % pct)] = REFGEN.run_contrastive_triplet_ebp(
                    wb, im_mates, im_nonmates, probe_im, net_name='resnetv4_pytorch', ebp_version=6, truncate_percent=pct,
                    device=cpu)
        im = images(4, seed=21)
        wb = fresh_wb(ebp_subtree_mode='norelu')          # create_wbnet.py: the eval flow's ctor mode
        G['ws_eval_smap'] = REFGEN.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch',
                                                                    subtree_mode_weighted='all', ebp_version=6, device=cpu, topk=4)
        wb = fresh_wb(ebp_subtree_mode='affineonly_with_prior')
        G['ws_v7_smap'] = REFGEN.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch',
                                                                  subtree_mode_weighted='affineonly_with_prior', ebp_version=7,
                                                                  device=cpu, topk=4)
        wb = fresh_wb()
        G['mean_ebp'] = REFGEN.mean_ebp(wb, im[3], net_name='resnetv4_pytorch', ebp_version=6, device=cpu)
    out = os.path.join(ROOT, 'tests', 'golden', 'inpaintgame_seed0.npz')
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays, %.0f KB)' % (out, len(G), os.path.getsize(out) / 1024))
    print('transitions: first twin blend at percentile', PCT_DENSITY[np.argmax(G['cls'])], '; pg', G['pg_dist'][[0, 50, 100]],
          'pr', G['pr_dist'][[0, 50, 100]])


if __name__ == '__main__':
    main()
