"""VGGFace2 ResNet-50-128d with the REAL bundled weights on the REAL bundled triplet (the north star's "bundled demo
triplets"): oracle and host schedule on the CPU, CUDA kernels on the GPU, all against the reference's own outputs
(tests/golden/resnet50_128_real.npz, made by oracle/gen_golden_r50.py).  Skipped when oracle/_ref/resnet50_128.pth
(95 MB, git-ignored, extracted from the reference's tarball by the generator) is not in the working tree."""
import numpy as np
import pytest
import torch

from emul_backend import EmulBackend
from helpers import ShadowBackend, pack_map, r50_inputs, rel_err
from xfr_b200.engine import Resnet50_128Engine

MODES = (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly'), ('norelu', 'norelu'))
R50 = r50_inputs()
needs_weights = pytest.mark.skipif(R50 is None, reason='oracle/_ref/resnet50_128.pth not present')


@needs_weights
def test_oracle_matches_reference_fingerprints():
    from oracle import resnet50_128_oracle as O
    sd, G, X = R50
    assert rel_err(O.encode(sd, X['mate']).numpy(), G['enc_mate']) < 1e-6
    fc1 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float()
    P0 = torch.zeros(1, 2)
    P0[0, 0] = 1
    P, kinds = O.ebp_mwp(sd, X['probe'], P0, fc1)
    assert kinds == [str(k) for k in G['P_kinds']] and len(P) == 158
    sums = np.array([float(p.double().sum()) for p in P])
    assert np.max(np.abs(sums - G['Psum_awp_probe']) / (np.abs(G['Psum_awp_probe']) + 1e-30)) < 1e-5
    c = O.contrastive_ebp(sd, X['probe'], fc1)[0]
    # SURVEY 8c fingerprints of the shimmed reference: max 1.649949e-03 at pixel 5553
    assert abs(float(G['cebp_awp_probe'].max()) - 1.649949e-03) < 1e-8 and int(G['cebp_awp_probe'].argmax()) == 5553
    assert int(c.argmax()) == 5553 and rel_err(c, G['cebp_awp_probe']) < 1e-4
    t = O.contrastive_ebp(sd, X['probe'], fc1, percentile=20)[0]
    assert rel_err(t, G['tcebp20_awp_probe']) < 1e-4


@needs_weights
@pytest.mark.parametrize('mode,tag', MODES)
def test_schedule_emulation(mode, tag):
    sd, G, X = R50
    eng = Resnet50_128Engine(sd, EmulBackend(impl_name='tf32x3'))
    x = torch.cat([X['probe'], X['demo']]).permute(0, 2, 3, 1).contiguous()
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float().unsqueeze(0).repeat(2, 1, 1)
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    s = eng.ebp(x, P1, W2, mode).clone().numpy()
    c = eng.contrastive(x, W2, mode=mode).clone().numpy()
    t = eng.contrastive(x, W2, mode=mode, percentile=20).clone().numpy()
    # 'tf32x3' packs: the emulation multiplies relu(W) rounded to TF32 in the W+ GEMMs (the product's two-pass plan), which
    # moves the real-weights maps by 2e-5 (EBP) / 1.5e-4 (contrastive) of their maximum
    for i, p in enumerate(('probe', 'demo')):
        assert rel_err(s[i], G['ebp_%s_%s' % (tag, p)]) < 1e-4
        assert rel_err(c[i], G['cebp_%s_%s' % (tag, p)]) < 1e-3
        assert rel_err(t[i], G['tcebp20_%s_%s' % (tag, p)]) < 1e-3


@needs_weights
@pytest.mark.gpu
@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
def test_gpu_kernels_shadow(impl):
    from xfr_b200.kernels import CudaBackend
    sd, G, X = R50
    dev = torch.device('cuda:0')
    be = CudaBackend(dev, impl=impl)
    eng_cpu = Resnet50_128Engine(sd, EmulBackend(impl_name=impl))
    eng = Resnet50_128Engine(sd, be, device=dev)
    sh = ShadowBackend(be, EmulBackend(), pack_map(eng, eng_cpu))
    eng.be = sh
    x = torch.cat([X['probe'], X['demo']]).permute(0, 2, 3, 1).contiguous().to(dev)
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float().unsqueeze(0).repeat(2, 1, 1).to(dev)
    for mode in ('affineonly_with_prior', 'all'):
        eng.contrastive(x, W2, mode=mode)
    tol = 2e-5 if impl == 'fp32' else 2e-4
    print('\n'.join('%-20s %.3g' % kv for kv in sorted(sh.errors.items())))
    bad = {k: v for k, v in sh.errors.items() if v > tol}
    assert not bad, bad


@needs_weights
@pytest.mark.gpu
@pytest.mark.parametrize('impl', ['fp32', 'tf32x3', 'tf32'])
def test_gpu_real_weights_real_triplet_vs_reference(impl):
    """North-star parity: maps of the CUDA path vs the reference's CPU output on the bundled triplet, <= 1e-4 max-abs."""
    from xfr_b200 import whitebox
    sd, G, X = R50
    dev = torch.device('cuda:0')
    sdd = {k: v.to(dev) for k, v in sd.items()}
    report = {}
    for mode, tag in MODES:
        wb = whitebox.Whitebox(whitebox.Whitebox_resnet50_128(sdd, impl=impl), ebp_subtree_mode=mode)
        x_mate, x_non = wb.net.encode(X['mate']), wb.net.encode(X['nonmate'])
        assert rel_err(x_mate.cpu().numpy(), G['enc_mate']) < (1e-4 if impl != 'tf32' else 2e-2)
        wb.net.set_triplet_classifier(x_mate, x_non)
        P0 = torch.zeros(1, 2)
        P0[0][0] = 1.0
        for p in ('probe', 'demo'):
            for key, got in (('ebp', wb.ebp(X[p], P0)), ('cebp', wb.contrastive_ebp(X[p], 0, 1)),
                             ('tcebp20', wb.truncated_contrastive_ebp(X[p], 0, 1, percentile=20))):
                ref = G['%s_%s_%s' % (key, tag, p)]
                report['%s_%s_%s' % (key, tag, p)] = (float(np.abs(got - ref).max()), rel_err(got, ref))
    print('\n'.join('%-26s max-abs %.3g   max-abs/max(ref) %.3g' % (k, v[0], v[1]) for k, v in sorted(report.items())))
    rel_tol = {'fp32': 1e-3, 'tf32x3': 1e-2, 'tf32': 1.0}[impl]
    for k, (a, r) in report.items():
        assert a < 1e-4, (k, a)
        assert r < rel_tol, (k, r)


@needs_weights
@pytest.mark.parametrize('mode,tag', (('affineonly_with_prior', 'awp'), ('all', 'all')))
def test_generic_sweep_emulation(mode, tag):
    """The firing-by-firing sweep (xfr_b200.generic.R50Sweep) on the emulated kernel set: all 158 recorded MWPs and the
    layerwise priors reproduce the reference (whitebox.py:561-581)."""
    sd, G, X = R50
    eng = Resnet50_128Engine(sd, EmulBackend())
    x = X['probe'].permute(0, 2, 3, 1).contiguous()
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float().unsqueeze(0)
    P1 = torch.zeros(1, 2)
    P1[0, 0] = 1
    eng.forward(x)
    gs = eng.sweep()
    P, names, P2 = gs.run(P1, W2, mode, record=True)
    assert names == [str(k) for k in G['P_kinds']] and len(P) == 158
    sums = np.array([float(p.double().sum()) for p in P[:-1]])
    gsum = G['Psum_%s_probe' % tag][:-1]
    assert np.max(np.abs(sums - gsum) / (np.abs(gsum) + 1e-30)) < 2e-5
    assert rel_err(P2.sum(-1)[0].numpy(), G['ebp_mwp_%s_probe' % tag]) < 1e-5
    if tag == 'awp':
        Pm = [None if p is None else p.clone() for p in P]
        ks = [int(k) for k in G['lw_k']]
        priors = {k: (r, (Pm[k] * (Pm[k] == Pm[k].max())).reshape(-1).contiguous()) for r, k in enumerate(ks)}
        _, _, P2 = gs.run(torch.zeros(len(ks), 2), W2, mode, priors=priors)
        maps = P2.sum(-1).numpy()
        for r, k in enumerate(ks):
            assert rel_err(maps[r], G['lw_argmax_%d' % k]) < 2e-5, k
        for k, e in zip(ks, G['lw_el_idx']):
            ed = gs.elem_index(k, int(e), Pm[k].shape)
            _, _, P2 = gs.run(torch.zeros(1, 2), W2, mode, priors={k: (0, ed, float(Pm[k].reshape(-1)[ed]))})
            assert rel_err(P2.sum(-1)[0].numpy(), G['lw_el_%d' % k]) < 2e-5, k


@needs_weights
@pytest.mark.gpu
def test_gpu_layerwise_and_weighted_subtree_vs_reference():
    """layerwise_ebp / weighted_subtree_ebp of the ResNet-50-128d plugin on the real weights and triplet."""
    from xfr_b200 import whitebox
    sd, G, X = R50
    dev = torch.device('cuda:0')
    sdd = {k: v.to(dev) for k, v in sd.items()}
    wb = whitebox.Whitebox(whitebox.Whitebox_resnet50_128(sdd))
    x_mate, x_non = wb.net.encode(X['mate']), wb.net.encode(X['nonmate'])
    wb.net.set_triplet_classifier(x_mate, x_non)
    for k, e in zip(G['lw_k'], G['lw_el_idx']):
        m = wb.layerwise_ebp(X['probe'], k_layer=int(k), mode='argmax', mwp=True)
        assert rel_err(m, G['lw_argmax_%d' % k]) < 1e-2, k
        m = wb.layerwise_ebp(X['probe'], k_layer=int(k), mode='elementwise', k_element=int(e), mwp=True)
        assert rel_err(m, G['lw_el_%d' % k]) < 1e-2, k
    assert len(wb.P) == 158 and [str(n) for n in wb.P_layername] == [str(n) for n in G['P_kinds']]
    smap, P_img, P_sub, k_sub = wb.weighted_subtree_ebp(X['probe'], 0, 1, topk=16, verbose=False, do_max_subtree=False,
                                                        do_mated_similarity_gating=True, subtree_mode='affineonly_with_prior')
    assert [int(k) for k in k_sub] == [int(k) for k in G['ws_k']]
    assert np.allclose(P_sub, G['ws_scores'], rtol=1e-2)
    assert rel_err(P_img[-1], G['ws_first']) < 2e-2
    assert np.abs(smap - G['ws_smap']).max() < 1e-4 and rel_err(smap, G['ws_smap']) < 5e-2


@needs_weights
def test_inpaintgame_rows_emulated():
    """SURVEY 8(f) rows 1 and 3 through the ResNet-50-128d plugin (un-normalised 128-d encodings, head on the wrapper) with the
    REAL weights: batched front end == per-job flow, scoring == the numpy expression of inpainting_game.py:124-146."""
    from xfr_b200 import inpaintgame as IG
    from xfr_b200 import whitebox

    class _EmulR50(whitebox.Whitebox_resnet50_128):
        def _device(self):
            return torch.device('cpu')

        def engine(self, with_bias=False):
            if self._engine is None:
                self._engine = Resnet50_128Engine(self._sd, EmulBackend(), with_bias=with_bias)
            return self._engine

    sd, G, X = R50
    wb = whitebox.Whitebox(_EmulR50(sd))
    crops = {k: G['crop_' + k] for k in ('probe', 'mate', 'nonmate', 'demo')}           # 224x224x3 uint8
    jobs = [([crops['mate']], [crops['nonmate']], crops['probe']), ([crops['mate'], crops['probe']], [crops['nonmate']], crops['demo'])]
    got = IG.run_contrastive_triplet_ebp_batch(wb, jobs)
    assert got.shape == (2, 112, 112)
    # job 0 is the golden triplet up to the classifier scale: unit-norm rows / 2500 instead of the raw encodings
    xm, xn = wb.encode(X['mate']), wb.encode(X['nonmate'])
    wb.net.set_triplet_classifier((xm / torch.norm(xm)) / 2500.0, (xn / torch.norm(xn)) / 2500.0)
    want = wb.contrastive_ebp(X['probe'], 0, 1)
    assert rel_err(got[0], want) < 1e-4 and int(got[0].argmax()) == int(want.argmax())
    # scoring: the probe turned into the non-mate, five percentiles
    orig, inp = X['probe'][0].numpy(), X['nonmate'][0].numpy()
    rng = np.random.RandomState(3)
    smap = (got[0].repeat(2, 0).repeat(2, 1) + 1e-7 * rng.rand(224, 224)).astype(np.float32)
    pct = np.array([0, 25, 50, 75, 100])
    gal_o, gal_p = wb.embeddings([orig]), wb.embeddings([inp])
    cls, pg, pr = IG.classified_as_inpainted_twin(wb, orig, inp, gal_o, gal_p, smap, 'percent-density', percentiles=pct, seed=0)
    masks = IG.create_threshold_masks(smap, 'percent-density', percentiles=pct, seed=0)[:, np.newaxis]
    blends = (1.0 - masks) * orig.astype(np.float64)[np.newaxis] + masks * inp.astype(np.float64)[np.newaxis]
    emb = wb.embeddings(blends)
    emb = emb / np.linalg.norm(emb, axis=1, keepdims=True)
    assert np.abs(pr - np.linalg.norm(emb - gal_o, axis=1)).max() < 1e-5 and np.abs(pg - np.linalg.norm(emb - gal_p, axis=1)).max() < 1e-5
    assert not cls[0] and cls[-1] and np.all(np.diff(pg) <= 1e-3)        # the more salient pixels go, the closer to the twin
