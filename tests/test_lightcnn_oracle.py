"""CPU: oracle/lightcnn_oracle.py reproduces what the unmodified reference produced for Light-CNN-29v2
(tests/golden/lightcnn29v2_seed0.npz, written by oracle/gen_golden_lightcnn.py)."""
import os

import numpy as np
import pytest
import torch

from helpers import GOLD, rel_err
from oracle import lightcnn_oracle as O
from xfr_b200 import synth

MODES = (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly'), ('norelu', 'norelu'))
NUM_CLASSES = 64


def lc_setup():
    G = np.load(os.path.join(GOLD, 'lightcnn29v2_seed0.npz'))
    sd = synth.lightcnn_state_dict(0, NUM_CLASSES)
    imgs = synth.lightcnn_probes(3, seed=1)
    noise = synth.lightcnn_probes(1, seed=2, smooth=False)
    fc2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float()
    return G, sd, imgs, noise, fc2


def test_encode_matches_reference():
    G, sd, imgs, _, _ = lc_setup()
    assert rel_err(O.encode(sd, imgs[1:2]).numpy(), G['enc_mate']) < 1e-6
    assert rel_err(O.encode(sd, imgs[2:3]).numpy(), G['enc_nonmate']) < 1e-6


@pytest.mark.parametrize('mode,tag', MODES)
def test_all_modes(mode, tag):
    G, sd, imgs, noise, fc2 = lc_setup()
    P0 = O._onehot(1, 2, 0)
    for pname, x in (('smooth', imgs[0:1]), ('noise', noise)):
        P, kinds = O.ebp_mwp(sd, x, P0, fc2, mode=mode)
        assert len(P) == 87                                        # 87 firings in triplet mode
        if tag == 'awp' and pname == 'smooth':
            assert kinds == [str(k) for k in G['P_kinds']]         # firing order = k_layer indexing
            assert [p.numel() for p in P] == list(G['P_numel'])
        sums = np.array([float(p.double().sum()) for p in P])
        gs = G['Psum_%s_%s' % (tag, pname)]
        assert np.max(np.abs(sums - gs) / (np.abs(gs) + 1e-30)) < 1e-6
        assert rel_err(O.ebp(sd, x, P0, fc2, mwp=True, mode=mode)[0], G['ebp_mwp_%s_%s' % (tag, pname)]) < 1e-6
        assert rel_err(O.ebp(sd, x, P0, fc2, mode=mode)[0], G['ebp_%s_%s' % (tag, pname)]) < 1e-5
        assert rel_err(O.contrastive_ebp(sd, x, fc2, mode=mode)[0], G['cebp_%s_%s' % (tag, pname)]) < 1e-4
        assert rel_err(O.contrastive_ebp(sd, x, fc2, percentile=20, mode=mode)[0], G['tcebp20_%s_%s' % (tag, pname)]) < 1e-4


def test_hooked_fc2_head_and_layerwise_prior():
    G, sd, imgs, _, fc2 = lc_setup()
    x = imgs[0:1]
    Pk = O._onehot(1, NUM_CLASSES, 5)
    P, kinds = O.ebp_mwp(sd, x, Pk, None)
    assert len(P) == 88 and kinds[0] == 'Linear'
    gs = G['Psum_awp_fc2head']
    assert np.max(np.abs(np.array([float(p.double().sum()) for p in P]) - gs) / (np.abs(gs) + 1e-30)) < 1e-6
    assert rel_err(O.contrastive_ebp(sd, x, None, 5, 9, num_classes=NUM_CLASSES)[0], G['cebp_awp_fc2head']) < 1e-4
    # layerwise_ebp(mode='argmax') (whitebox.py:570-571): the prior keeps only the arg-max element of P_mate[k]
    Pm, _ = O.ebp_mwp(sd, x, O._onehot(1, 2, 0), fc2)
    for k in G['lw_k']:
        pm = Pm[int(k)]
        pr = pm * (1.0 - (pm != pm.max()).float())
        P, _ = O.ebp_mwp(sd, x, 0.0 * O._onehot(1, 2, 0), fc2, prior={int(k): pr}, stop_at_stem=True)
        assert rel_err(P[-2].sum(1)[0].numpy(), G['lw_elem_%d' % k]) < 1e-5, k
