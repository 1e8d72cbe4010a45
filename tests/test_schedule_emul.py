"""CPU: the host schedule (xfr_b200/engine.py) driven through the torch emulation of the kernel set
reproduces the reference's maps.  This pins hook order, the Add closure quirk, the un-hooked
triplet fc2, packing layouts and the J = G*N row convention without a GPU."""
import numpy as np
import pytest
import torch

from emul_backend import EmulBackend
from helpers import L101, L1111, golden, golden_inputs, rel_err
from xfr_b200 import packing, synth
from xfr_b200.engine import StResnetEngine

MODES = (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly'), ('norelu', 'norelu'))


def _engine(layers, impl='fp32'):
    return StResnetEngine(synth.stresnet_state_dict(0, layers, 2), EmulBackend(impl_name=impl), layers)


def test_pack_roundtrip():
    g = torch.Generator().manual_seed(3)
    w = torch.randn(128, 64, 3, 3, generator=g)
    b = torch.randn(128, generator=g)
    Bf, bias = packing.pack_dual_fwd(w, b, 128)
    t, p = packing.unpack_dual_cols(Bf.t().contiguous().t().reshape(1, -1, Bf.shape[1]).permute(2, 0, 1).reshape(Bf.shape[1], -1), 128)
    wk = w.permute(0, 2, 3, 1).reshape(128, -1)
    assert torch.equal(t, wk.t()) and torch.equal(p, wk.clamp_min(0).t())
    bt, bp = packing.unpack_dual_cols(bias.view(1, -1), 128)
    assert torch.equal(bt[0], b) and torch.equal(bp[0], b)
    # dgrad pack == conv2d_input
    y = torch.randn(1, 128, 5, 5, generator=g)
    want = torch.nn.grad.conv2d_input((1, 64, 5, 5), w.clamp_min(0), y, 1, 1)
    from emul_backend import im2col_nhwc
    got = im2col_nhwc(y.permute(0, 2, 3, 1).contiguous(), 3, 3, 1) @ packing.pack_dgrad(w).t()
    assert torch.allclose(got.view(1, 5, 5, 64).permute(0, 3, 1, 2), want, atol=1e-4)


def test_split_planes_are_exact():
    g = torch.Generator().manual_seed(9)
    B = torch.randn(64, 96, generator=g) * torch.logspace(-6, 3, 96)
    P = packing.gemm_planes(B, 'tf32x3')
    assert torch.equal(P[0] + P[1], B)                                   # hi + lo reproduces W bit for bit
    assert torch.equal(P[0].view(torch.int32) & 0x1FFF, torch.zeros(64, 96, dtype=torch.int32))   # hi is a TF32 number
    assert float((P[1].abs() / B.abs()).max()) <= 2.0 ** -11 + 1e-9      # |lo| <= half a TF32 ulp


@pytest.mark.parametrize('mode,tag', MODES)
@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
def test_small_net(mode, tag, impl):
    G = golden(L1111)
    eng = _engine(L1111, impl)
    x, W2, imgs = golden_inputs(G)
    xn = eng.forward(imgs.permute(0, 2, 3, 1).contiguous())
    assert rel_err(50 * xn[1:2].numpy(), G['enc_mate']) < 1e-5
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    m = eng.ebp(x, P1, W2, mode, saliency=False).clone().numpy()
    s = eng.ebp(x, P1, W2, mode).clone().numpy()
    c = eng.contrastive(x, W2, mode=mode).clone().numpy()
    # 'tf32x3' packs: the emulation multiplies relu(W) rounded to TF32 in the W+ GEMMs, as the product's two-pass plan does
    tol = 1e-5 if impl == 'fp32' else 1e-4
    for i, pname in enumerate(('smooth', 'noise')):
        assert rel_err(m[i], G['ebp_mwp_%s_%s' % (tag, pname)]) < tol
        assert rel_err(s[i], G['ebp_%s_%s' % (tag, pname)]) < tol
        assert rel_err(c[i], G['cebp_%s_%s' % (tag, pname)]) < 50 * tol


def test_resnet101_default_mode():
    G = golden(L101)
    eng = _engine(L101)
    x, W2, _ = golden_inputs(G)
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    s = eng.ebp(x, P1, W2).clone().numpy()
    assert rel_err(s[0], G['ebp_awp_smooth']) < 1e-5
    assert rel_err(s[1], G['ebp_awp_noise']) < 1e-4
    c = eng.contrastive(x, W2).clone().numpy()
    # ill-conditioned on synthetic encodings (cos(mate, non-mate) = 0.9999): the reference's own fp32 noise
    # (oracle vs hooks, both CPU fp32) is 2e-3 of the map maximum here; see DESIGN.md "parity".
    assert rel_err(c[0], G['cebp_awp_smooth']) < 2e-2
    assert np.abs(c[0] - G['cebp_awp_smooth']).max() < 1e-4


@pytest.mark.parametrize('layers', [L1111, L101])
def test_truncated_and_with_bias(layers):
    G = golden(layers)
    sd = synth.stresnet_state_dict(0, layers, 2)
    x, W2, _ = golden_inputs(G)
    eng = StResnetEngine(sd, EmulBackend(), layers)
    c = eng.contrastive(x, W2, percentile=20).clone().numpy()
    tol = 5e-4 if layers == L1111 else 2e-2
    for i, pname in enumerate(('smooth', 'noise')):
        assert rel_err(c[i], G['tcebp20_awp_%s' % pname]) < tol
        assert np.abs(c[i] - G['tcebp20_awp_%s' % pname]).max() < 1e-4
    engb = StResnetEngine(sd, EmulBackend(), layers, with_bias=True)
    P1 = torch.zeros(1, 2)
    P1[:, 0] = 1
    m = engb.ebp(x[:1], P1, W2[:1], saliency=False).clone().numpy()
    assert rel_err(m[0], G['ebp_mwp_awp_withbias']) < 1e-5
    assert rel_err(m[0], G['ebp_mwp_awp_smooth']) > 1e-4        # with_bias really changes the map


@pytest.mark.parametrize('layers', [L1111, L101])
def test_hooked_fc2_head(layers):
    """No set_triplet_classifier: the network's own fc2 takes part with relu(W) and adds a leading Linear firing."""
    G = golden(layers)
    sd = synth.stresnet_state_dict(0, layers, 2)
    eng = StResnetEngine(sd, EmulBackend(), layers)
    x, _, _ = golden_inputs(G)
    W2 = sd['fc2.weight']
    P1 = torch.zeros(1, 2)
    P1[0, 1] = 1
    m = eng.ebp(x[:1].contiguous(), P1, W2, hooked_fc2=True, saliency=False).clone().numpy()
    assert rel_err(m[0], G['ebp_mwp_awp_fc2head']) < 1e-5
    c = eng.contrastive(x[:1].contiguous(), W2, hooked_fc2=True, num_classes=2).clone().numpy()
    assert np.abs(c[0] - G['cebp_awp_fc2head']).max() < 1e-4 and rel_err(c[0], G['cebp_awp_fc2head']) < (5e-4 if layers == L1111 else 5e-2)
    if layers == L1111:
        from xfr_b200.generic import GenericSweep
        eng.forward(x[:1].contiguous())
        P, names, P2 = GenericSweep(eng).run(P1, W2, 'affineonly_with_prior', record=True, hooked_fc2=True)
        assert len(P) == len(G['P_kinds']) + 1 and names[0] == 'Linear'          # 379 vs 378 firings on the 101
        assert rel_err(P2.sum(-1)[0].numpy(), G['ebp_mwp_awp_fc2head']) < 1e-5


@pytest.mark.parametrize('layers', [L1111, L101])
def test_opt_in_plans_emulated(layers):
    """The parity estimates DESIGN.md section 8 quotes for the opt-in plans of kernels.HYBRID_IMPLS, on the kernel emulation:
    'tf32x2f' (signed forward weights rounded to TF32) keeps every map far inside the 1e-4 bar; 'tf32x3b1' (one TF32 pass in the
    W+ dgrads) does not on the ill-conditioned ResNet-101 golden triplet."""
    G = golden(layers)
    x, W2, _ = golden_inputs(G)
    sd = synth.stresnet_state_dict(0, layers, 2)
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    eng = StResnetEngine(sd, EmulBackend(impl_name='tf32x3', fwd_two_pass=True), layers)
    s = eng.ebp(x, P1, W2).clone().numpy()
    c = eng.contrastive(x, W2).clone().numpy()
    t = eng.contrastive(x, W2, percentile=20).clone().numpy()
    for i, p in enumerate(('smooth', 'noise')):
        assert rel_err(s[i], G['ebp_awp_%s' % p]) < 5e-3
        assert np.abs(c[i] - G['cebp_awp_%s' % p]).max() < 1e-5 and np.abs(t[i] - G['tcebp20_awp_%s' % p]).max() < 1e-5
    if layers == L101:
        eng = StResnetEngine(sd, EmulBackend(impl_name='tf32x3', bwd_single_pass=True), layers)
        c = eng.contrastive(x, W2).clone().numpy()
        assert np.abs(c[0] - G['cebp_awp_smooth']).max() > 1e-4          # why that plan stays opt-in


def _bighead_check(eng_factory, to_dev, tol):
    """the STR network's own 65,359-class fc2 (33 M parameters, hooked: relu(W) in the backward and a leading Linear
    firing) - STRise.mean_ebp_prior (blackbox.py:280-294: uniform prior over all classes) and the demo's contrastive_ebp(x, 0, 100)
    (demo/test_whitebox.py:92-99) against the reference's outputs (tests/golden/stresnet101_bighead_seed0.npz)."""
    import os
    from helpers import GOLD
    G = np.load(os.path.join(GOLD, 'stresnet101_bighead_seed0.npz'))
    C = int(G['num_classes'])
    assert C == 65359
    sd = synth.stresnet_state_dict(0, L101, C)
    eng = eng_factory(sd)
    x = to_dev(synth.smooth_probes(3, seed=1)[0:1].permute(0, 2, 3, 1).contiguous())
    W2 = to_dev(sd['fc2.weight'])
    m = eng.ebp(x, to_dev(torch.ones(1, C)), W2, hooked_fc2=True, saliency=False).cpu().numpy()
    assert rel_err(m[0], G['mean_ebp_mwp']) < tol
    s = eng.ebp(x, to_dev(torch.ones(1, C)), W2, hooked_fc2=True).cpu().numpy()
    assert rel_err(s[0], G['mean_ebp']) < tol
    P1 = torch.zeros(1, C)
    P1[0, 100] = 1
    m = eng.ebp(x, to_dev(P1), W2, hooked_fc2=True, saliency=False).cpu().numpy()
    assert rel_err(m[0], G['ebp_100_mwp']) < tol
    c = eng.contrastive(x, W2, k_pos=0, k_neg=100, hooked_fc2=True, num_classes=C).cpu().numpy()
    assert np.abs(c[0] - G['cebp_0_100']).max() < 1e-4
    return rel_err(c[0], G['cebp_0_100'])


def test_big_hooked_head_emulated():
    r = _bighead_check(lambda sd: StResnetEngine(sd, EmulBackend(), L101), lambda t: t, 1e-5)
    assert r < 5e-2
