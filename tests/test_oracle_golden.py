"""CPU: the oracle restatement reproduces what the unmodified reference produced (tests/golden)."""
import numpy as np
import pytest
import torch

from helpers import L101, L1111, golden, rel_err
from oracle import stresnet_oracle as O
from xfr_b200 import synth

MODES = (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly'), ('norelu', 'norelu'))


def _setup(layers):
    G = golden(layers)
    sd = synth.stresnet_state_dict(0, layers, 2)
    imgs = synth.smooth_probes(3, seed=1)
    noise = synth.synthetic_probes(1, seed=2)
    fc2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float() / 2500.0
    return G, sd, imgs, noise, fc2


def test_encode_matches_reference():
    G, sd, imgs, _, _ = _setup(L1111)
    assert rel_err(O.encode(sd, imgs[1:2], L1111).numpy(), G['enc_mate']) < 1e-6
    assert rel_err(O.encode(sd, imgs[2:3], L1111).numpy(), G['enc_nonmate']) < 1e-6


@pytest.mark.parametrize('mode,tag', MODES)
def test_small_net_all_modes(mode, tag):
    G, sd, imgs, noise, fc2 = _setup(L1111)
    for pname, x in (('smooth', imgs[0:1]), ('noise', noise)):
        P, kinds = O.ebp_mwp(sd, x, O._onehot(1, 2, 0), fc2, mode=mode, layers=L1111)
        assert len(P) == len(G['Psum_%s_%s' % (tag, pname)])
        if tag == 'awp' and pname == 'smooth':
            assert [k for k in kinds] == [str(k) for k in G['P_kinds']]       # firing order = k_layer indexing
            assert [p.numel() for p in P] == list(G['P_numel'])
        sums = np.array([float(p.double().sum()) for p in P])
        gs = G['Psum_%s_%s' % (tag, pname)]
        assert np.max(np.abs(sums - gs) / (np.abs(gs) + 1e-30)) < 2e-5
        kw = dict(mode=mode, layers=L1111)
        assert rel_err(O.ebp(sd, x, O._onehot(1, 2, 0), fc2, mwp=True, **kw)[0], G['ebp_mwp_%s_%s' % (tag, pname)]) < 1e-5
        assert rel_err(O.ebp(sd, x, O._onehot(1, 2, 0), fc2, **kw)[0], G['ebp_%s_%s' % (tag, pname)]) < 1e-5
        # contrastive maps subtract two nearly equal MWPs (synthetic encodings have cos ~ 0.998): fp32 noise is amplified
        assert rel_err(O.contrastive_ebp(sd, x, fc2, **kw)[0], G['cebp_%s_%s' % (tag, pname)]) < 5e-4
        assert rel_err(O.contrastive_ebp(sd, x, fc2, percentile=20, **kw)[0], G['tcebp20_%s_%s' % (tag, pname)]) < 5e-4
    assert rel_err(O.ebp(sd, imgs[0:1], torch.ones(1, 2), fc2, mode=mode, layers=L1111)[0], G['meanebp_%s_smooth' % tag]) < 1e-5


def test_small_net_hooked_fc2_and_with_bias():
    G, sd, imgs, _, fc2 = _setup(L1111)
    x = imgs[0:1]
    assert rel_err(O.ebp(sd, x, O._onehot(1, 2, 1), None, mwp=True, layers=L1111)[0], G['ebp_mwp_awp_fc2head']) < 1e-5
    assert rel_err(O.contrastive_ebp(sd, x, None, layers=L1111)[0], G['cebp_awp_fc2head']) < 5e-4
    assert rel_err(O.ebp(sd, x, O._onehot(1, 2, 0), fc2, mwp=True, with_bias=True, layers=L1111)[0],
                   G['ebp_mwp_awp_withbias']) < 1e-5


@pytest.mark.parametrize('mode,tag', (('affineonly_with_prior', 'awp'), ('all', 'all'), ('norelu', 'norelu')))
def test_small_net_layerwise_prior(mode, tag):
    G, sd, imgs, _, fc2 = _setup(L1111)
    x = imgs[0:1]
    Pm, _ = O.ebp_mwp(sd, x, O._onehot(1, 2, 0), fc2, mode=mode, layers=L1111)
    for i, (k, e) in enumerate(zip(G['lw_k'], G['lw_el_%s' % tag])):
        pr = torch.zeros_like(Pm[k]).flatten()
        pr[e] = Pm[k].flatten()[e]
        P, _ = O.ebp_mwp(sd, x, 0.0 * O._onehot(1, 2, 0), fc2, mode=mode, prior={int(k): pr}, layers=L1111,
                         stop_at_stem=True)
        got = P[-2].sum(1)[0].numpy()
        assert rel_err(got, G['lw_%s' % tag][i]) < 2e-5, (mode, k)


def test_resnet101_default_mode():
    G, sd, imgs, noise, fc2 = _setup(L101)
    x = imgs[0:1]
    P, _ = O.ebp_mwp(sd, x, O._onehot(1, 2, 0), fc2, layers=L101)
    assert len(P) == 378                                  # triplet mode has 378 firings
    sums = np.array([float(p.double().sum()) for p in P])
    gs = G['Psum_awp_smooth']
    assert np.max(np.abs(sums - gs) / (np.abs(gs) + 1e-30)) < 2e-5
    got = O.mwp_to_saliency(P[-2].sum(1)[0].numpy().astype(np.float32))
    assert rel_err(got, G['ebp_awp_smooth']) < 1e-5
    assert np.abs(got - G['ebp_awp_smooth']).max() < 1e-4   # the north-star bar (vacuous on a sum-normalised map)
