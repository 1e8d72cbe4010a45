"""TEST INFRASTRUCTURE.  Runs the reference's UNMODIFIED caller code on top of this repo's classes through the drop-in
overlay (dropin/xfr): the job functions of python/xfr/inpainting_game/generate_whitebox_saliency.py:81-215, the scoring code of
python/xfr/inpainting_game/inpainting_game.py:83-146 and the Whitebox call sequence of demo/test_whitebox.py:77-145, 173-199,
with `from xfr.models import whitebox` resolving to the B200 engine exactly as eval/create_wbnet.py:4-7 imports it.

Launched by tests/test_dropin.py in a fresh interpreter with
    PYTHONPATH = <repo>/dropin : <repo>/oracle/shim : /root/reference/python
(the shim only satisfies the reference's imports of skimage / imageio, absent from this image).  Needs /root/reference, so it
runs in the build container; there the kernels are the torch emulation (--backend emul), on a B200 the CUDA library
(--backend cuda).  Prints one JSON line with the deviations from the reference-generated goldens."""
import contextlib
import io
import json
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [os.path.join(ROOT, 'tests')]

import numpy as np  # noqa: E402
import torch  # noqa: E402

import xfr  # noqa: E402                       (the overlay, which runs the reference's own __init__)
from xfr import xfr_root  # noqa: E402         create_wbnet.py:5
from xfr.models import whitebox  # noqa: E402  create_wbnet.py:6  -> dropin/xfr/models/whitebox.py
from xfr.models import resnet  # noqa: E402    create_wbnet.py:7  -> the reference's model definition
from xfr.inpainting_game import generate_whitebox_saliency as REFGEN  # noqa: E402   (reference, unmodified)
from xfr.inpainting_game import inpainting_game as REFGAME  # noqa: E402             (reference, unmodified)

from xfr_b200 import synth  # noqa: E402
import xfr_b200.whitebox  # noqa: E402

backend = sys.argv[sys.argv.index('--backend') + 1] if '--backend' in sys.argv else 'emul'
out = {'whitebox_module': whitebox.__file__, 'xfr_root': xfr_root, 'refgen': REFGEN.__file__, 'resnet': resnet.__file__}
assert whitebox.Whitebox is xfr_b200.whitebox.Whitebox and os.path.samefile(os.path.dirname(whitebox.__file__), os.path.join(ROOT, 'dropin', 'xfr', 'models'))
assert REFGEN.__file__.startswith('/root/reference/') and resnet.__file__.startswith('/root/reference/') and xfr_root == '/root/reference'

if backend == 'emul':       # no GPU in the build container: the kernel set is the torch emulation (tests/emul_backend.py)
    from emul_backend import EmulBackend
    from xfr_b200.engine import StResnetEngine

    def _engine(self, with_bias=False):
        if self._engine is None or self._engine.with_bias != with_bias:
            self._engine = StResnetEngine(self._sd, EmulBackend(), self._layers, with_bias=with_bias)
        return self._engine
    xfr_b200.whitebox.WhiteboxSTResnet._device = lambda self: torch.device('cpu')
    xfr_b200.whitebox.WhiteboxSTResnet.engine = _engine
    torch.set_num_threads(os.cpu_count())
device = torch.device('cuda:0') if backend == 'cuda' else torch.device('cpu')
LAYERS = (1, 1, 1, 1)


def fresh_wb(on_device, **kw):
    """eval/create_wbnet.py:33-45, with the seeded synthetic weights in the reference's own ResNet class (the real STR weights are
    git-LFS pointers): model -> WhiteboxSTResnet -> Whitebox(...).to(device), attributes set as the factory sets them."""
    model = resnet.ResNet(resnet.Bottleneck, list(LAYERS), mode='encode', num_classes=2)
    model.load_state_dict(synth.stresnet_state_dict(0, LAYERS, 2))
    if on_device:
        model.to(device)
    wbnet = whitebox.WhiteboxSTResnet(model)
    net = whitebox.Whitebox(wbnet, **kw).to(device)
    net.match_threshold = 0.9636
    net.platts_scaling = 15.05
    return net


def rel(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


from inpaintgame_fixture import PCT_DENSITY, images, jobs, scoring_fixture  # noqa: E402
G = np.load(os.path.join(ROOT, 'tests', 'golden', 'inpaintgame_seed0.npz'))
G1 = np.load(os.path.join(ROOT, 'tests', 'golden', 'stresnet1111_seed0.npz'))
err = {}
with contextlib.redirect_stdout(io.StringIO()):
    # ---- generate_whitebox_saliency.py job functions, batch 1, one fresh Whitebox per job like the eval driver
    for i, (im_mates, im_nonmates, probe_im) in enumerate(jobs()):
        for pct in (None, 20):
            wb = fresh_wb(True)
            m = REFGEN.run_contrastive_triplet_ebp(wb, im_mates, im_nonmates, probe_im, net_name='resnetv4_pytorch', ebp_version=6,
                                                   truncate_percent=pct, device=device)
            key = 'job%d_contrastive%s' % (i, '' if pct is None else '_pct%d' % pct)
            assert m.shape == (112, 112) and m.dtype == np.float32
            err[key] = [float(np.abs(m - G[key]).max()), rel(m, G[key])]
    im = images(4, seed=21)
    wb = fresh_wb(True, ebp_subtree_mode='norelu')
    m = REFGEN.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch', subtree_mode_weighted='all',
                                                ebp_version=6, device=device, topk=4)
    err['ws_eval_smap'] = [float(np.abs(m - G['ws_eval_smap']).max()), rel(m, G['ws_eval_smap'])]
    wb = fresh_wb(True, ebp_subtree_mode='affineonly_with_prior')
    m = REFGEN.run_weighted_subtree_triplet_ebp(wb, im[0:2], im[2:3], im[3], net_name='resnetv4_pytorch',
                                                subtree_mode_weighted='affineonly_with_prior', ebp_version=7, device=device, topk=4)
    err['ws_v7_smap'] = [float(np.abs(m - G['ws_v7_smap']).max()), rel(m, G['ws_v7_smap'])]
    wb = fresh_wb(True)
    m = REFGEN.mean_ebp(wb, im[3], net_name='resnetv4_pytorch', ebp_version=6, device=device)
    err['mean_ebp'] = [float(np.abs(m - G['mean_ebp']).max()), rel(m, G['mean_ebp'])]
    # ---- inpainting_game.py scoring over Whitebox.embeddings
    F = scoring_fixture()
    snet = fresh_wb(True)
    gal_o, gal_p = snet.embeddings([F['orig']]), snet.embeddings([F['inp']])
    cls, pg, pr = REFGAME.classified_as_inpainted_twin(snet, F['orig'], F['inp'], gal_o, gal_p, F['smap'],
                                                       mask_threshold_method='percent-density', percentiles=PCT_DENSITY, seed=0)
    err['twin_cls_mismatches'] = [int((np.asarray(cls) != G['cls']).sum()), 0.0]
    err['twin_pg_dist'] = [float(np.abs(pg - G['pg_dist']).max()), rel(pg, G['pg_dist'])]
    # ---- demo/test_whitebox.py:77-145, 173-199, 257-280: a CPU-resident network, CPU tensors in, numpy maps out
    probe = synth.smooth_probes(3, seed=1)
    wb = fresh_wb(False)                                        # Whitebox(WhiteboxSTResnet(stresnet101(...))): never moved to a device
    P = torch.zeros((1, wb.net.num_classes()))
    P[0][1] = 1.0
    m = wb.ebp(probe[0:1], P, mwp=True)                          # the network's own (hooked) fc2 head
    err['demo_ebp_fc2head'] = [float(np.abs(m - G1['ebp_mwp_awp_fc2head']).max()), rel(m, G1['ebp_mwp_awp_fc2head'])]
    m = wb.contrastive_ebp(probe[0:1], k_poschannel=0, k_negchannel=1)
    err['demo_cebp_fc2head'] = [float(np.abs(m - G1['cebp_awp_fc2head']).max()), rel(m, G1['cebp_awp_fc2head'])]
    x_mate, x_nonmate = wb.net.encode(probe[1:2]).detach(), wb.net.encode(probe[2:3]).detach()
    assert not x_mate.is_cuda
    err['demo_encode'] = [float(np.abs(x_mate.numpy() - G1['enc_mate']).max()), rel(x_mate.numpy(), G1['enc_mate'])]
    wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_nonmate)
    P = torch.zeros((1, wb.net.num_classes()))
    P[0][0] = 1.0
    for key, m in (('ebp_awp_smooth', wb.ebp(probe[0:1], P)), ('cebp_awp_smooth', wb.contrastive_ebp(probe[0:1], k_poschannel=0, k_negchannel=1)),
                   ('tcebp20_awp_smooth', wb.truncated_contrastive_ebp(probe[0:1], k_poschannel=0, k_negchannel=1, percentile=20))):
        err['demo_' + key] = [float(np.abs(m - G1[key]).max()), rel(m, G1[key])]
    wb = fresh_wb(False, ebp_version=5, ebp_subtree_mode='affineonly_with_prior')      # demo weighted_subtree_ebp (:173-199)
    wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_nonmate)
    smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(probe[0:1], 0, 1, topk=8, verbose=False, do_max_subtree=False,
                                                        do_mated_similarity_gating=True, subtree_mode='all')
    d = np.abs(smap.astype(int) - G1['ws_demo_smap'].astype(int))
    err['demo_ws_u8_pixels_off_by_more_than_2'] = [int((d > 2).sum()), float(d.max())]
    err['demo_ws_k_set_diff'] = [len(set(int(k) for k in k_sub) ^ set(int(k) for k in G1['ws_demo_k'])), 0.0]
    # ---- attributes callers read
    wb = fresh_wb(False)
    wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_nonmate)
    wb.record_P = True
    wb.ebp(probe[0:1], P, mwp=True)
    out['P_len'], out['P_names_ok'] = len(wb.P), wb.P_layername == [str(s) for s in G1['P_kinds']]
    out['P_sums_rel'] = max(abs(float(p.double().sum()) / s - 1.0) for p, s in zip(wb.P[:-1], G1['Psum_awp_smooth'][:-1]) if s > 0)
    out['layerlist'] = len(wb.layerlist)
    out['layerlist_names_match_reference_visit'] = [d['name'] for d in wb.layerlist][:3]
out['err'] = err
print(json.dumps(out))
