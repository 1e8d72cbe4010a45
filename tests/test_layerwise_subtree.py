"""layerwise_ebp / layerwise_contrastive_ebp / weighted_subtree_ebp through the drop-in Whitebox API against the
reference's outputs (tests/golden).  CPU: the host logic over the kernel emulation; GPU: the CUDA kernels."""
import os

import numpy as np
import pytest
import torch

from emul_backend import EmulBackend
from helpers import L101, L1111, golden, rel_err
from xfr_b200 import synth, whitebox
from xfr_b200.engine import StResnetEngine

SUBTREE = (  # tag, ctor mode, subtree_mode, gating, do_max, ebp_version, classifier scale (1.0 = unit-norm rows)
    ('ws_demo', 'affineonly_with_prior', 'all', True, False, 5, 1.0 / 2500.0),     # demo/test_whitebox.py:173-199
    ('ws_eval', 'norelu', 'all', False, False, None, 1.0),                          # generate_whitebox_saliency.py:143
    ('ws_norelu_max', 'norelu', 'norelu', True, True, None, 1.0 / 2500.0))


class _EmulNet(whitebox.WhiteboxSTResnet):
    """Test double: the same plugin class with the engine built on the torch emulation of the kernels."""

    def _device(self):
        return torch.device('cpu')

    def engine(self, with_bias=False):
        if self._engine is None or self._engine.with_bias != with_bias:
            self._engine = StResnetEngine(self._sd, EmulBackend(), self._layers, with_bias=with_bias)
        return self._engine


def _net(layers, gpu):
    sd = synth.stresnet_state_dict(0, layers, 2)
    if gpu:
        return whitebox.WhiteboxSTResnet({k: v.cuda() for k, v in sd.items()}, layers=layers)
    return _EmulNet(sd, layers=layers)


def _check_subtree(gpu):
    G = golden(L1111)
    probe = synth.smooth_probes(3, seed=1)[0:1]
    xm, xn = torch.from_numpy(G['enc_mate']), torch.from_numpy(G['enc_nonmate'])
    for tag, ctor_mode, sub_mode, gating, mx, ver, scale in SUBTREE:
        wb = whitebox.Whitebox(_net(L1111, gpu), ebp_version=ver, ebp_subtree_mode=ctor_mode)
        a = xm / torch.norm(xm) if scale == 1.0 else scale * xm
        b = xn / torch.norm(xn) if scale == 1.0 else scale * xn
        wb.net.set_triplet_classifier(a, b)
        smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(probe, 0, 1, topk=8, verbose=False, do_max_subtree=mx,
                                                            do_mated_similarity_gating=gating, subtree_mode=sub_mode)
        assert wb.ebp_subtree_mode() == sub_mode                           # the reference's side effect (whitebox.py:651)
        assert len(P_img) == len(P_sub) == len(k_sub) == 8
        # firings chained on one tensor tie exactly; np.argsort may order a tie differently, the maps are the same
        assert sorted(P_sub) == pytest.approx(sorted(G[tag + '_scores']), rel=1e-3)
        assert len(set(int(k) for k in k_sub) ^ set(int(k) for k in G[tag + '_k'])) <= 2
        ref = G[tag + '_smap']
        if ver is None:
            assert smap.dtype == np.float32 and rel_err(smap, ref) < 2e-3
        else:
            d = np.abs(smap.astype(int) - ref.astype(int))          # 8-bit quantisation edges: a 1e-5 change can move a pixel
            assert smap.dtype == np.uint8 and int((d > 2).sum()) <= 5 and d.max() <= 16


def _check_layerwise(layers, gpu, tol):
    G = golden(layers)
    probe = synth.smooth_probes(3, seed=1)[0:1]
    xm, xn = torch.from_numpy(G['enc_mate']), torch.from_numpy(G['enc_nonmate'])
    for mode, tag in (('affineonly_with_prior', 'awp'), ('all', 'all'), ('norelu', 'norelu')):
        wb = whitebox.Whitebox(_net(layers, gpu), ebp_subtree_mode=mode)
        wb.net.set_triplet_classifier(xm / 2500.0, xn / 2500.0)
        ks = list(zip(G['lw_k'], G['lw_el_%s' % tag]))
        for i, (k, e) in list(enumerate(ks))[::(1 if layers == L1111 else 4)]:
            got = wb.layerwise_ebp(probe, k_layer=int(k), k_element=int(e), mode='elementwise', k_poschannel=0, mwp=True)
            assert got.shape == (112, 112)
            assert rel_err(got, G['lw_%s' % tag][i]) < tol, (mode, k)
        assert len(wb.P) == len(G['P_kinds']) and wb.P_layername == [str(s) for s in G['P_kinds']]
        assert tuple(wb.P[3].shape[1:]) == (2048, 7, 7)                     # self.P entries look like the reference's [1,C,H,W]


def test_weighted_subtree_emulated():
    _check_subtree(False)


def test_layerwise_emulated():
    _check_layerwise(L1111, False, 2e-5)


def _check_layer_sweep(gpu, tol):
    """BASELINE configs[2]: the percentile layer sweep of one triplet against the reference's own layerwise_contrastive_ebp run
    firing by firing (tests/golden/layersweep1111_seed0.npz, oracle/gen_golden_layersweep.py)."""
    import os
    from helpers import GOLD
    G, R = golden(L1111), np.load(os.path.join(GOLD, 'layersweep1111_seed0.npz'))
    n = int(R['n_firings'])
    wb = whitebox.Whitebox(_net(L1111, gpu))
    wb.net.set_triplet_classifier(torch.from_numpy(G['enc_mate']) / 2500.0, torch.from_numpy(G['enc_nonmate']) / 2500.0)
    probe = synth.smooth_probes(3, seed=1)[0:1]
    maps = wb.layerwise_contrastive_ebp_sweep(probe, 0, 1, list(range(n)), mode='percentile', percentile=20, rows_per_sweep=16)
    assert maps.shape == (n, 112, 112) and maps.dtype == np.float32
    assert wb.P_layername == [str(s).split('(')[0] for s in R['names']]          # module kinds in firing order
    live = R['map_max'] > 0
    assert live.sum() == 15 and not live[n - 1]
    for k in range(n):
        if not live[k]:
            assert not maps[k].any(), k                                    # non-affine firings / empty contrast: all-zero map
        else:
            assert abs(float(maps[k].astype(np.float64).sum()) - R['map_sum'][k]) < 1e-4
            assert abs(float(maps[k].max()) / R['map_max'][k] - 1.0) < tol, k
    for k, ref in zip(R['full_k'], R['full']):
        if live[k]:
            assert rel_err(maps[k], ref) < tol and np.abs(maps[k] - ref).max() < 1e-4, k
    # the batched sweep is the per-layer operator, layer by layer
    for k in (3, 29, -2):
        with pytest.warns(UserWarning):
            one = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode='percentile', percentile=20)
        assert rel_err(maps[k % n], one) < 1e-6
    sub = wb.layerwise_contrastive_ebp_sweep(probe, 0, 1, [29, 3, 29, n - 1], mode='copy', mwp=True)      # order, repeats, last firing
    assert sub.shape == (4, 112, 112) and np.array_equal(sub[0], sub[2]) and not sub[3].any() and sub[1].max() > 0
    with pytest.raises(ValueError):
        wb.layerwise_contrastive_ebp_sweep(probe, 0, 1, [3], mode='elementwise')
    with pytest.warns(UserWarning):
        assert not wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=n - 1, mode='percentile', percentile=20).any()
    assert not wb.layerwise_ebp(probe, k_layer=-1, mode='argmax').any()


def _check_other_modes(gpu, tol):
    """The remaining modes of layerwise_contrastive_ebp, layerwise_ebp's default 'argmax' mode and other truncation percentiles
    against the reference's own outputs (same golden file)."""
    import os
    from helpers import GOLD
    G, R = golden(L1111), np.load(os.path.join(GOLD, 'layersweep1111_seed0.npz'))
    wb = whitebox.Whitebox(_net(L1111, gpu))
    wb.net.set_triplet_classifier(torch.from_numpy(G['enc_mate']) / 2500.0, torch.from_numpy(G['enc_nonmate']) / 2500.0)
    probe = synth.smooth_probes(3, seed=1)[0:1]
    for mode in ('copy', 'mean', 'product', 'argmax', 'argmax_product', 'percentile_argmax'):
        for k in (7, 29):
            with pytest.warns(UserWarning):
                m = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode=mode, percentile=20)
            assert rel_err(m, R['lc_%s_%d' % (mode, k)]) < tol, (mode, k)
        sweep = wb.layerwise_contrastive_ebp_sweep(probe, 0, 1, [7, 29], mode=mode, percentile=20)
        assert rel_err(sweep[1], R['lc_%s_29' % mode]) < tol and rel_err(sweep[0], R['lc_%s_7' % mode]) < tol
    for k in (3, 15, 29, 46, 57):
        m = wb.layerwise_ebp(probe, k_layer=k, mode='argmax', k_poschannel=0, mwp=True)
        assert rel_err(m, R['lw_argmax_%d' % k]) < tol, k
    assert rel_err(wb.layerwise_ebp(probe, k_layer=-2), R['lw_argmax_57']) < tol                    # defaults: argmax, mwp
    for pct in (0, 50, 80, 100):
        t = wb.truncated_contrastive_ebp(probe, 0, 1, percentile=pct)
        ref = R['trunc_pct%d' % pct]
        assert np.abs(t - ref).max() < 1e-4
        assert rel_err(t, ref) < 10 * tol if ref.max() > 0 else not t.any(), pct


def test_other_modes_emulated():
    _check_other_modes(False, 2e-3)


def test_layer_sweep_emulated():
    _check_layer_sweep(False, 2e-3)


@pytest.mark.gpu
def test_layer_sweep_gpu():
    _check_layer_sweep(True, 2e-2)


def test_layerwise_contrastive_modes_emulated():
    """Every mode of the deprecated layerwise_contrastive_ebp runs and obeys its definition (whitebox.py:606-642):
    'copy' with the contrastive difference of the LAST-but-one firing as prior reproduces... itself at that firing."""
    wb = whitebox.Whitebox(_net(L1111, False))
    G = golden(L1111)
    wb.net.set_triplet_classifier(torch.from_numpy(G['enc_mate']) / 2500.0, torch.from_numpy(G['enc_nonmate']) / 2500.0)
    probe = synth.smooth_probes(3, seed=1)[0:1]
    k = 7        # (unnormalised non-mate MWPs dominate the mate ones below firing 8 on this synthetic triplet)
    with pytest.warns(UserWarning):
        base = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode='copy', mwp=True)
    assert base.shape == (112, 112) and np.isfinite(base).all() and base.max() > 0
    prior_copy = wb.P[k].clone()
    for mode in ('mean', 'product', 'argmax', 'argmax_product', 'percentile', 'percentile_argmax'):
        with pytest.warns(UserWarning):
            m = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode=mode, percentile=20, mwp=True)
        assert np.isfinite(m).all()
        if 'argmax' in mode:
            assert int((wb.P[k] > 0).sum()) <= 2                            # a single (tied) node seeds the sub-tree
    with pytest.warns(UserWarning):
        pm = wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode='percentile', percentile=0, mwp=True)
    assert rel_err(pm, base) < 1e-6                                         # percentile 0 keeps everything == 'copy'
    assert torch.equal(wb.P[k], prior_copy)
    with pytest.raises(ValueError):
        wb.layerwise_contrastive_ebp(probe, 0, 1, k_layer=k, mode='nope')


@pytest.mark.gpu
def test_weighted_subtree_gpu():
    _check_subtree(True)


def _check_subtree_resnet101(gpu):
    """weighted_subtree_ebp on the full [3,4,23,3] net with the evaluation flow's settings (generate_whitebox_saliency.py:122-205)
    against one run of the unmodified reference (tests/golden/stresnet101_subtree_seed0.npz, ~760 hooked ebp() calls, 499 s on 8
    cores).  The 32 selected firings must be the SAME SET; their order may differ only inside groups of firings chained on one
    tensor, whose scores tie exactly (np.argsort is not stable across summation orders) and whose sub-tree maps coincide."""
    from helpers import GOLD
    G = np.load(os.path.join(GOLD, 'stresnet101_subtree_seed0.npz'))
    probe = synth.smooth_probes(3, seed=1)[0:1]
    wb = whitebox.Whitebox(_net(L101, gpu), ebp_subtree_mode='norelu')
    wb.net.set_triplet_classifier(torch.from_numpy(G['row_mate']), torch.from_numpy(G['row_nonmate']))
    smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(probe, 0, 1, topk=32, verbose=False, do_max_subtree=False,
                                                        do_mated_similarity_gating=False, subtree_mode='all')
    assert len(wb.P_layername) == int(G['n_firings']) == 378
    k_sub, k_ref = [int(k) for k in k_sub], [int(k) for k in G['ws_k']]
    assert set(k_sub) == set(k_ref)
    tol = 2e-2 if gpu else 1e-4          # scores are maxima of true-gradient products: measured <= 1.2e-2 on the tensor-core plans
    assert sorted(P_sub) == pytest.approx(sorted(G['ws_scores']), rel=tol)
    score = dict(zip(k_ref, G['ws_scores']))
    for pos, (a, b) in enumerate(zip(k_sub, k_ref)):
        assert a == b or abs(score[a] - score[b]) <= (2 * tol if gpu else 1e-6) * abs(score[b]), (pos, a, b)      # reordered only inside (near-)ties
    ref_max = dict(zip(k_ref, G['ws_maps_max']))
    off = [k for k, m in zip(k_sub, P_img) if abs(float(np.max(m)) / ref_max[k] - 1.0) > (5e-2 if gpu else 1e-3)]
    r = rel_err(smap, G['ws_smap'])
    print('weighted_subtree_ebp ResNet-101: %d / 32 sub-tree maps off in their maximum %s, smap max-abs/max(ref) %.3g' % (len(off), off, r))
    # a sub-tree is seeded at the arg-max ELEMENT of its firing's gated true gradient (whitebox.py:687-696): on the tensor-core plans a
    # near-tie between two elements may pick the other one (measured: firing 333), which changes that one map, not the selected set
    assert len(off) <= (2 if gpu else 0)
    # (one re-seeded sub-tree of 32 moves the merged map by its weight: measured 0.11 of the maximum with firing 333 re-seeded)
    assert smap.dtype == np.float32 and r < ((0.25 if off else 5e-2) if gpu else 1e-3)
    assert np.abs(smap - G['ws_smap']).max() < (1e-3 if off else 1e-4)


@pytest.mark.gpu
def test_weighted_subtree_resnet101_gpu():
    _check_subtree_resnet101(True)


@pytest.mark.skipif(not os.environ.get('XFRB_SLOW'), reason='200 s of kernel emulation on 8 cores (set XFRB_SLOW=1); the GPU twin runs in the -m gpu suite')
def test_weighted_subtree_resnet101_emulated():
    _check_subtree_resnet101(False)


@pytest.mark.gpu
def test_other_modes_gpu():
    """the six other layerwise_contrastive_ebp modes, layerwise_ebp 'argmax', truncation percentiles 0 / 50 / 80 / 100 on the CUDA
    kernels against the reference's outputs"""
    _check_other_modes(True, 2e-2)


@pytest.mark.gpu
def test_layerwise_gpu():
    _check_layerwise(L1111, True, 2e-3)
    _check_layerwise(L101, True, 2e-2)
