"""GPU: end-to-end parity of the CUDA path with the reference's own outputs (tests/golden, produced by
oracle/gen_golden.py from the unmodified reference) and with the oracle on fresh seeded inputs."""
import os

import numpy as np
import pytest
import torch

from helpers import L101, L1111, golden, golden_inputs, rel_err
from xfr_b200 import synth

pytestmark = pytest.mark.gpu
MODES = (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly'), ('norelu', 'norelu'))


def _engine(layers, impl):
    from xfr_b200.engine import StResnetEngine
    from xfr_b200.kernels import CudaBackend
    dev = torch.device('cuda:0')
    try:
        be = CudaBackend(dev, impl=impl)
    except NotImplementedError:
        pytest.skip('impl %s not built' % impl)
    return StResnetEngine(synth.stresnet_state_dict(0, layers, 2), be, layers, device=dev), dev


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3', 'tf32x3full'])
@pytest.mark.parametrize('mode,tag', MODES)
def test_small_net_vs_reference(impl, mode, tag):
    G = golden(L1111)
    eng, dev = _engine(L1111, impl)
    x, W2, imgs = golden_inputs(G)
    x, W2 = x.to(dev), W2.to(dev)
    xn = eng.forward(imgs.permute(0, 2, 3, 1).contiguous().to(dev))
    assert rel_err(50 * xn[1:2].cpu().numpy(), G['enc_mate']) < 1e-4
    P1 = torch.zeros(2, 2, device=dev)
    P1[:, 0] = 1
    m = eng.ebp(x, P1, W2, mode, saliency=False).cpu().numpy()
    s = eng.ebp(x, P1, W2, mode).cpu().numpy()
    c = eng.contrastive(x, W2, mode=mode).cpu().numpy()
    tol = 1e-4 if impl == 'fp32' else 1e-3
    for i, pname in enumerate(('smooth', 'noise')):
        assert rel_err(m[i], G['ebp_mwp_%s_%s' % (tag, pname)]) < tol
        assert rel_err(s[i], G['ebp_%s_%s' % (tag, pname)]) < tol
        assert np.abs(c[i] - G['cebp_%s_%s' % (tag, pname)]).max() < 1e-4      # north-star bar
        assert rel_err(c[i], G['cebp_%s_%s' % (tag, pname)]) < 50 * tol          # cancellation-amplified


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
def test_resnet101_vs_reference(impl):
    G = golden(L101)
    eng, dev = _engine(L101, impl)
    x, W2, _ = golden_inputs(G)
    x, W2 = x.to(dev), W2.to(dev)
    P1 = torch.zeros(2, 2, device=dev)
    P1[:, 0] = 1
    report = {}
    for mode, tag in (('affineonly_with_prior', 'awp'), ('all', 'all')):
        s = eng.ebp(x, P1, W2, mode).cpu().numpy()
        c = eng.contrastive(x, W2, mode=mode).cpu().numpy()
        for i, pname in enumerate(('smooth', 'noise')):
            report['ebp_%s_%s' % (tag, pname)] = (np.abs(s[i] - G['ebp_%s_%s' % (tag, pname)]).max(),
                                                  rel_err(s[i], G['ebp_%s_%s' % (tag, pname)]))
            report['cebp_%s_%s' % (tag, pname)] = (np.abs(c[i] - G['cebp_%s_%s' % (tag, pname)]).max(),
                                                   rel_err(c[i], G['cebp_%s_%s' % (tag, pname)]))
    print('\n'.join('%-22s max-abs %.3g   max-abs/max(ref) %.3g' % (k, v[0], v[1]) for k, v in sorted(report.items())))
    for k, (a, r) in report.items():
        assert a < 1e-4, (k, a)                       # the north-star bar: <= 1e-4 max-abs per pixel
        if k.startswith('ebp'):
            assert r < (1e-3 if impl == 'fp32' else 5e-3), (k, r)


def test_batch_rows_are_independent():
    """A probe gives the same map alone and inside a batch (J = G*N row bookkeeping)."""
    eng, dev = _engine(L1111, 'fp32')
    x = synth.synthetic_probes(5, seed=7).permute(0, 2, 3, 1).contiguous().to(dev)
    g = torch.Generator().manual_seed(5)
    W2 = (torch.randn(5, 2, 512, generator=g) * 0.02).to(dev)
    full = eng.contrastive(x, W2).cpu().numpy().copy()
    for i in (0, 3):
        one = eng.contrastive(x[i:i + 1].contiguous(), W2[i:i + 1].contiguous()).cpu().numpy()
        assert rel_err(one[0], full[i]) < 1e-5


def test_whitebox_api_vs_reference():
    """The drop-in classes (xfr_b200.whitebox) called the way demo/test_whitebox.py calls the reference's."""
    from xfr_b200 import whitebox
    G = golden(L101)
    dev = torch.device('cuda:0')
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, L101, 2).items()}
    x, W2, imgs = golden_inputs(G)
    probe = imgs[0:1]                                       # NCHW like net.preprocess() returns
    for ver, key_c, key_t in ((None, 'cebp_awp_smooth', 'tcebp20_awp_smooth'), (11, 'cebp_v11_u8', None)):
        wb = whitebox.Whitebox(whitebox.WhiteboxSTResnet(sd), ebp_version=ver)
        x_mate = wb.net.encode(imgs[1:2])
        x_non = wb.net.encode(imgs[2:3])
        assert rel_err(x_mate.cpu().numpy(), G['enc_mate']) < 1e-3
        wb.net.set_triplet_classifier((1.0 / 2500.0) * x_mate, (1.0 / 2500.0) * x_non)
        assert wb.net.num_classes() == 2
        c = wb.contrastive_ebp(probe, k_poschannel=0, k_negchannel=1)
        assert c.shape == (112, 112)
        if ver is None:
            assert c.dtype == np.float32 and np.abs(c - G[key_c]).max() < 1e-4
            t = wb.truncated_contrastive_ebp(probe, 0, 1, percentile=20)
            assert np.abs(t - G[key_t]).max() < 1e-4 and rel_err(t, G[key_t]) < 5e-2
            P = torch.zeros(1, 2)
            P[0][0] = 1.0
            e = wb.ebp(probe, P)
            assert rel_err(e, G['ebp_awp_smooth']) < 5e-3
            em = wb.ebp(probe, P, mwp=True)
            assert rel_err(em, G['ebp_mwp_awp_smooth']) < 5e-3
        else:
            assert c.dtype == np.uint8                       # ebp_version 11: with_bias + uint8/PIL post-processing
            assert np.abs(c.astype(np.int32) - G[key_c].astype(np.int32)).max() <= 40   # contrastive noise amplified by min-max stretch
    with pytest.raises(RuntimeError):
        whitebox.Whitebox(whitebox.WhiteboxSTResnet(sd), ebp_version=3)


def test_full_size_properties():
    """Size-independent checks at a full 64-probe sweep: maps are non-negative, sum to one, permuting the batch permutes
    the maps, and swapping mate/non-mate rows equals swapping k_pos/k_neg."""
    eng, dev = _engine(L101, 'tf32x3')
    N = 64
    x = synth.synthetic_probes(N, seed=11).permute(0, 2, 3, 1).contiguous().to(dev)
    g = torch.Generator().manual_seed(12)
    W2 = (torch.randn(N, 2, 512, generator=g) * 0.02).to(dev)
    a = eng.contrastive(x, W2).clone()
    assert torch.isfinite(a).all() and (a >= 0).all()
    assert torch.allclose(a.sum(dim=(1, 2)), torch.ones(N, device=dev), atol=1e-4)
    perm = torch.randperm(N, generator=g).to(dev)
    b = eng.contrastive(x[perm].contiguous(), W2[perm].contiguous()).clone()
    assert float((b - a[perm]).abs().max() / a.max()) < 1e-4
    c = eng.contrastive(x, W2.flip(1).contiguous(), k_pos=1, k_neg=0).clone()
    assert float((c - a).abs().max() / a.max()) < 1e-4


def test_hooked_fc2_head_api():
    """demo/test_whitebox.py:77-107 style: no triplet classifier, the network's own fc2 is the (hooked) classifier."""
    from xfr_b200 import whitebox
    G = golden(L101)
    dev = torch.device('cuda:0')
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, L101, 2).items()}
    wb = whitebox.Whitebox(whitebox.WhiteboxSTResnet(sd))
    _, _, imgs = golden_inputs(G)
    P = torch.zeros((1, wb.net.num_classes()))
    P[0][1] = 1.0
    m = wb.ebp(imgs[0:1], P, mwp=True)
    assert rel_err(m, G['ebp_mwp_awp_fc2head']) < 5e-3
    c = wb.contrastive_ebp(imgs[0:1], k_poschannel=0, k_negchannel=1)
    assert np.abs(c - G['cebp_awp_fc2head']).max() < 1e-4


# Opt-in plan kernels.HYBRID_IMPLS['tf32x2f'] (two-pass forward dual convs: signed weights rounded to TF32); first run on a B200
# in round 2 (profiles/r2_bench_a_tf32x2f.json: 3,348 maps/s against 3,156 for the default plan).
@pytest.mark.parametrize('layers', [L1111, L101])
def test_two_pass_forward_plan_vs_reference(layers):
    G = golden(layers)
    eng, dev = _engine(layers, 'tf32x2f')
    x, W2, _ = golden_inputs(G)
    x, W2 = x.to(dev), W2.to(dev)
    P1 = torch.zeros(2, 2, device=dev)
    P1[:, 0] = 1
    s = eng.ebp(x, P1, W2).cpu().numpy()
    c = eng.contrastive(x, W2).cpu().numpy()
    t = eng.contrastive(x, W2, percentile=20).cpu().numpy()
    for i, pname in enumerate(('smooth', 'noise')):
        assert rel_err(s[i], G['ebp_awp_%s' % pname]) < 1e-2                     # emulation: 3e-3
        for got, key in ((c, 'cebp_awp_%s'), (t, 'tcebp20_awp_%s')):
            assert np.abs(got[i] - G[key % pname]).max() < 1e-4                   # north-star bar (emulation: 6e-6)
            assert rel_err(got[i], G[key % pname]) < 5e-2


def _wellcond(dev):
    from helpers import GOLD
    Gw = np.load(os.path.join(GOLD, 'stresnet101_wellcond_seed0.npz'))
    W2 = torch.cat((torch.from_numpy(Gw['row_mate']), torch.from_numpy(Gw['row_nonmate']))).unsqueeze(0).repeat(2, 1, 1).contiguous()
    return Gw, W2.to(dev)


@pytest.mark.parametrize('impl,tol', [('fp32', 1e-3), ('bf16x2', 1e-2), ('tf32x3', 1e-2), ('tf32x3full', 1e-2)])
def test_resnet101_wellcond_vs_reference(impl, tol):
    """The benched quantity under a MEANINGFUL bar.  On tests/golden/stresnet101_seed0.npz the classifier rows are encodings of
    random-weight images (cos(mate, non-mate) = 0.9999): the contrastive map subtracts two maps that agree to three digits and even
    two CPU fp32 implementations differ by 2e-3 of its maximum.  stresnet101_wellcond_seed0.npz (oracle/gen_golden_r101_extra.py, the
    unmodified reference) has well-separated rows (cos 0.3); there two CPU fp32 implementations agree to 1.7e-4 / 4.8e-4 (oracle /
    kernel emulation vs the reference).  Measured on a B200 (round 2): the fp32 CUDA-core plan holds 1e-4 .. 2e-4 of the map maximum
    (asserted: 1e-3, SURVEY 8d's aim); EVERY tensor-core plan sits at 3e-3 .. 6e-3 whatever its operand precision - bf16x2 3.8e-3 /
    6.1e-3, tf32x3 2.8e-3 / 6.0e-3, the fp32-equivalent three-pass tf32x3full 3.0e-3 / 5.8e-3 (smooth / noise probe) - because the
    tensor core accumulates in fp32 with truncation (a bias of ~1e-5 per GEMM, tools/bias_probe) and the signed forward sums
    amplify it; asserted: 1e-2, with the north star's 1e-4 max-abs bar (measured <= 4e-6)."""
    G = golden(L101)
    eng, dev = _engine(L101, impl)
    x, _, _ = golden_inputs(G)
    x = x.to(dev)
    Gw, W2 = _wellcond(dev)
    P0 = torch.zeros(2, 2, device=dev)
    P0[:, 1] = 1
    rep = {}
    for mode, tag in (('affineonly_with_prior', 'awp'), ('all', 'all')):
        c = eng.contrastive(x, W2, mode=mode).cpu().numpy()
        t = eng.contrastive(x, W2, mode=mode, percentile=20).cpu().numpy()
        e = eng.ebp(x, P0, W2, mode).cpu().numpy()
        for i, p in enumerate(('smooth', 'noise')):
            for key, got in (('cebp', c), ('tcebp20', t), ('ebp1', e)):
                ref = Gw['%s_%s_%s' % (key, tag, p)]
                rep['%s_%s_%s' % (key, tag, p)] = (float(np.abs(got[i] - ref).max()), rel_err(got[i], ref))
    print('\n'.join('%-24s max-abs %.3g   max-abs/max(ref) %.3g' % (k, v[0], v[1]) for k, v in sorted(rep.items())))
    for k, (a, r) in rep.items():
        assert a < 1e-4, (k, a)
        # mode 'all' divides by X everywhere: the reference's own fp32 noise there is 9.4e-4 on the truncated map (kernel emulation)
        assert r < (tol if '_awp_' in k else 3 * tol), (k, r)


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3', 'bf16x2'])
def test_with_bias_float_golden(impl):
    """with_bias = True (ebp_version 11, whitebox.py:286-289, 321-324) on the CUDA kernels against the reference's float MWP."""
    from xfr_b200.engine import StResnetEngine
    from xfr_b200.kernels import CudaBackend
    G = golden(L101)
    dev = torch.device('cuda:0')
    eng = StResnetEngine(synth.stresnet_state_dict(0, L101, 2), CudaBackend(dev, impl=impl), L101, device=dev, with_bias=True)
    x, W2, _ = golden_inputs(G)
    P1 = torch.zeros(1, 2, device=dev)
    P1[:, 0] = 1
    m = eng.ebp(x[:1].contiguous().to(dev), P1, W2[:1].to(dev), saliency=False).cpu().numpy()
    r = rel_err(m[0], G['ebp_mwp_awp_withbias'])
    print('with_bias MWP (%s): max-abs/max(ref) %.3g' % (impl, r))
    assert r < (1e-4 if impl == 'fp32' else 5e-3)        # un-normalised MWP: measured 1.4e-3 on the split-TF32 plan
    assert rel_err(m[0], G['ebp_mwp_awp_smooth']) > 1e-4             # with_bias really changes the map


def test_eps_reaches_the_kernels():
    """Whitebox(eps=...) (whitebox.py:267): a custom eps with with_bias = True (the reference docstring's own v11 configuration uses
    1e-12) must reach the engine that actually runs, and a second Whitebox on the same plugin gets its own eps back."""
    from xfr_b200 import whitebox
    dev = torch.device('cuda:0')
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, L1111, 2).items()}
    net = whitebox.WhiteboxSTResnet(sd, layers=L1111)
    g = torch.Generator().manual_seed(4)
    W2 = torch.randn(2, 512, generator=g) * 0.02
    net.set_triplet_classifier(W2[0:1], W2[1:2])
    x = synth.smooth_probes(1, seed=3)
    P = torch.zeros(1, 2)
    P[0, 0] = 1
    wa = whitebox.Whitebox(net, ebp_version=11, eps=1e-3)            # huge eps: the map must change measurably
    wb_ = whitebox.Whitebox(net, ebp_version=11)
    a = wa.ebp(x, P, mwp=True)
    assert wa._engine().be.eps == 1e-3 and wa._engine().with_bias
    b = wb_.ebp(x, P, mwp=True)
    assert wb_._engine().be.eps == 1e-16
    assert rel_err(a, b) > 1e-3
    a2 = wa.ebp(x, P, mwp=True)
    assert np.array_equal(a, a2)


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
def test_big_hooked_head_gpu(impl):
    """the 65,359-class hooked head (mean-EBP prior of the blackbox, demo contrastive_ebp(x, 0, 100)) on the CUDA kernels"""
    from test_schedule_emul import _bighead_check
    from xfr_b200.engine import StResnetEngine
    from xfr_b200.kernels import CudaBackend
    dev = torch.device('cuda:0')
    r = _bighead_check(lambda sd: StResnetEngine(sd, CudaBackend(dev, impl=impl), L101, device=dev), lambda t: t.to(dev),
                       1e-4 if impl == 'fp32' else 2e-3)
    print('contrastive_ebp(x, 0, 100) on the 65,359-class head: max-abs/max(ref) %.3g' % r)
    assert r < 5e-2
