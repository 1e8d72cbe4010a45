"""The bf16x2 plan (kernels.HYBRID_IMPLS, include/xfrb.h XFRB_IMPL_BF16X2): pair tensors, bf16 weight packs, the fused
sweep on tcgen05 kind::f16.  CPU: pack / pair formats and the schedule on the kernel emulation against the reference's
outputs; GPU: the kernels against fp64 restatements on decoded operands, then the sweep against the reference's outputs."""
import numpy as np
import pytest
import torch

from emul_backend import EmulBackend, im2col_nhwc
from helpers import GOLD, L101, L1111, golden, golden_inputs, rel_err
from xfr_b200 import packing, synth
from xfr_b200.engine import StResnetEngine


def test_pair_format_roundtrip():
    g = torch.Generator().manual_seed(2)
    x = torch.randn(7, 64, generator=g) * torch.logspace(-20, 10, 64)
    p = packing.to_pair(x)
    assert p.shape == x.shape and p.dtype == torch.float32              # the same bytes per row as fp32
    y = packing.from_pair(p)
    assert float(((y - x).abs() / x.abs()).max()) < 2.0 ** -16           # 16 significant bits
    b = p.view(torch.bfloat16)
    assert torch.equal(b[:, :64], x.bfloat16())                          # [hi half-row | lo half-row]


def test_bf16_packs():
    g = torch.Generator().manual_seed(3)
    sd = {'c.weight': torch.randn(128, 64, 3, 3, generator=g) * 0.05, 'c.bias': torch.randn(128, generator=g),
          'b.weight': torch.rand(128, generator=g), 'b.bias': torch.randn(128, generator=g),
          'b.running_mean': torch.randn(128, generator=g), 'b.running_var': torch.rand(128, generator=g) + 0.5}
    L = packing.ConvBN(sd, 'c', 'b', 'bf16x2')
    L32 = packing.ConvBN(sd, 'c', 'b', 'tf32x3')
    assert L.Bf.dtype == L.Bd.dtype == torch.bfloat16 and L.Bf.shape == (2,) + tuple(L32.Bf.shape[1:]) and L.Bd.shape == L32.Bd.shape[1:]
    W = L32.Bf.sum(0)
    assert float((L.Bf.float().sum(0) - W).abs().max() / W.abs().max()) < 2.0 ** -15     # signed weights: two bf16 terms
    Wp = L32.Bd.sum(0)
    assert float((L.Bd.float() - Wp).abs().max() / Wp.abs().max()) < 2.0 ** -8 and float(L.Bd.float().min()) >= 0.0
    # the fp32-activation sweeps (layerwise / weighted-subtree operators) multiply the SAME bf16-rounded relu(W) as a split-TF32 pack
    assert torch.equal(L.Bd32()[0], L.Bd.float()) and not L.Bd32()[1].any() and torch.equal(L.signed_dgrad(), L32.signed_dgrad())


def _engine(layers):
    return StResnetEngine(synth.stresnet_state_dict(0, layers, 2), EmulBackend(impl_name='tf32x3', pairs=True), layers)


@pytest.mark.parametrize('layers', [L1111, L101])
def test_schedule_emulated(layers):
    """What the plan costs in parity, on the kernel emulation (bit-exact pair tensors, bf16 packs, fp32 accumulation): every
    map stays inside the north star's 1e-4 max-abs bar with >= 5x margin; single EBP maps within 1e-3 of their maximum."""
    G = golden(layers)
    eng = _engine(layers)
    x, W2, _ = golden_inputs(G)
    P1 = torch.zeros(2, 2)
    P1[:, 0] = 1
    s = eng.ebp(x, P1, W2).clone().numpy()
    c = eng.contrastive(x, W2).clone().numpy()
    t = eng.contrastive(x, W2, percentile=20).clone().numpy()
    for i, p in enumerate(('smooth', 'noise')):
        assert rel_err(s[i], G['ebp_awp_%s' % p]) < 2e-3
        assert np.abs(c[i] - G['cebp_awp_%s' % p]).max() < 2.5e-5 and np.abs(t[i] - G['tcebp20_awp_%s' % p]).max() < 2.5e-5
    if layers == L101:      # the well-conditioned triplet (oracle/gen_golden_r101_extra.py): scale-aware
        Gw = np.load(GOLD + '/stresnet101_wellcond_seed0.npz')
        W2w = torch.cat((torch.from_numpy(Gw['row_mate']), torch.from_numpy(Gw['row_nonmate']))).unsqueeze(0).repeat(2, 1, 1).contiguous()
        c = eng.contrastive(x, W2w).clone().numpy()
        for i, p in enumerate(('smooth', 'noise')):
            assert rel_err(c[i], Gw['cebp_awp_%s' % p]) < 5e-3


# ------------------------------------------------------------------------------------------------------------------ GPU
def _cuda_be():
    from xfr_b200.kernels import CudaBackend
    return CudaBackend(torch.device('cuda:0'), impl='bf16x2')


@pytest.mark.gpu
def test_to_pair_gpu():
    be = _cuda_be()
    g = torch.Generator().manual_seed(5)
    x = (torch.randn(3, 9, 9, 128, generator=g) * torch.logspace(-12, 6, 128)).cuda()
    p = torch.empty_like(x)
    be.to_pair(x, p)
    assert torch.equal(p.cpu(), packing.to_pair(x.cpu()))                 # bit-identical to the host statement of the format
    y = torch.empty_like(x)
    be.to_pair(p, y, inverse=True)
    assert torch.equal(y.cpu(), packing.from_pair(p.cpu()))


class _L(object):
    pass


def _rand_layer(cin, cout, R, g):
    """A ConvBN pack with random weights / BatchNorm in the bf16x2 layout."""
    sd = {'c.weight': torch.randn(cout, cin, R, R, generator=g) / (R * cin ** 0.5), 'c.bias': torch.randn(cout, generator=g) * 0.1,
          'b.weight': torch.randn(cout, generator=g) * 0.5 + 0.7, 'b.bias': torch.randn(cout, generator=g) * 0.2,
          'b.running_mean': torch.randn(cout, generator=g) * 0.3, 'b.running_var': torch.rand(cout, generator=g) + 0.5}
    return packing.ConvBN(sd, 'c', 'b', 'bf16x2')


GEMM_CASES = [  # N (images), H, Cin, Cout, R
    (3, 56, 64, 64, 1), (3, 56, 64, 64, 3), (2, 56, 64, 256, 1), (3, 28, 128, 128, 3), (3, 28, 128, 512, 1), (5, 14, 256, 256, 3),
    (5, 14, 1024, 256, 1), (5, 7, 512, 512, 3), (3, 7, 512, 2048, 1),
    # enough tiles for the CTA-pair (cta_group::2) kernels, incl. an odd tile count and a ragged M tail
    (13, 56, 256, 64, 1), (32, 28, 128, 128, 3), (40, 14, 256, 1024, 1), (160, 7, 512, 512, 3), (24, 56, 64, 64, 3),
]


@pytest.mark.gpu
def test_forward_dual_gpu():
    """xfrb_conv_dual under XFRB_IMPL_BF16X2 against the fp64 statement on the decoded operands: o = conv_W(a) + b with W as
    two bf16 terms, xr = relu(conv_W+(a) + b) with relu(W) as one, act = relu(bn(o) + res) as a pair tensor and in fp32."""
    be = _cuda_be()
    g = torch.Generator().manual_seed(7)
    errs = {}
    for (N, H, cin, cout, R) in GEMM_CASES:
        L = _rand_layer(cin, cout, R, g)
        a = torch.relu(torch.randn(N, H, H, cin, generator=g))
        ap = packing.to_pair(a)
        res = torch.randn(N, H, H, cout, generator=g)
        Ld = _L()
        Ld.__dict__.update(L.__dict__)
        for k in ('Bf', 'bias', 'bn'):
            setattr(Ld, k, getattr(L, k).cuda())
        o, xr, act, actf = (torch.full((N, H, H, cout), float('nan'), device='cuda') for _ in range(4))
        be.conv_dual(ap.cuda(), Ld, o, xr, act, res.cuda(), act_f32=actf)
        torch.cuda.synchronize()
        A = im2col_nhwc(packing.from_pair(ap).double(), R, R, R // 2)
        Bf = L.Bf.double()
        t, _ = packing.unpack_dual_cols(A @ Bf.sum(0).t() + L.bias.double(), L.tn)
        _, p = packing.unpack_dual_cols(A @ Bf[0].t() + L.bias.double(), L.tn)
        want_act = torch.relu(t * L.bn[0].double() + L.bn[1].double() + res.double().view(-1, cout))
        e = lambda got, want: float((got.cpu().double().view(-1, cout) - want).abs().max() / want.abs().max())
        errs[(N, H, cin, cout, R)] = (e(o, t), e(xr, torch.relu(p)), e(actf, want_act),
                                      e(packing.from_pair(act.cpu()), want_act))
    print('\n'.join('%-26s o %.2e  xr %.2e  act %.2e  act(pair) %.2e' % ((str(k),) + v) for k, v in errs.items()))
    for k, v in errs.items():
        assert max(v[:3]) < 2e-5 and v[3] < 4e-5, (k, v)          # fp32 accumulation; the pair adds its 2^-17 rounding


@pytest.mark.gpu
def test_dgrads_gpu():
    """xfrb_dgrad_plain / _mid / _join under XFRB_IMPL_BF16X2 against the kernel emulation on the same pair tensors
    (the emulation decodes the pairs and multiplies the bf16 relu(W) in fp32)."""
    be = _cuda_be()
    emul = EmulBackend(impl_name='tf32x3', pairs=True)
    g = torch.Generator().manual_seed(8)
    errs = {}
    for (N, H, cin, cout, R) in GEMM_CASES:
        for J in (N, 2 * N):
            L = _rand_layer(cin, cout, R, g)          # dgrad: A = y [J,H,H,cout] (K), output [J,H,H,cin]
            Ld = _L()
            Ld.__dict__.update(L.__dict__)
            Ld.Bd = L.Bd.cuda()
            y = torch.relu(torch.randn(J, H, H, cout, generator=g)) * 1e-3
            yp = packing.to_pair(y)
            o = torch.randn(N, H, H, cin, generator=g)
            xr = torch.rand(N, H, H, cin, generator=g) + 0.05
            bn = packing.fold_bn({'b.weight': torch.randn(cin, generator=g) * 0.5 + 0.7, 'b.bias': torch.randn(cin, generator=g) * 0.2,
                                  'b.running_mean': torch.randn(cin, generator=g) * 0.3, 'b.running_var': torch.rand(cin, generator=g) + 0.5}, 'b')
            key = (J, H, cin, cout, R)
            # plain
            z = torch.full((J, H, H, cin), float('nan'), device='cuda')
            be.dgrad_plain(yp.cuda(), Ld, z, pair=True)
            zw = torch.empty(J, H, H, cin)
            emul.dgrad_plain(yp, L, zw, pair=True)
            e = [rel_err(z.cpu().numpy(), zw.numpy())]
            for mode in (0, 1, 2):
                yo = torch.full((J, H, H, cin), float('nan'), device='cuda')
                be.dgrad_mid(yp.cuda(), Ld, o.cuda(), xr.cuda(), bn.cuda(), mode, yo)
                yw = torch.empty(J, H, H, cin)
                emul.dgrad_mid(yp, L, o, xr, bn, mode, yw)
                got, want = packing.from_pair(yo.cpu()), packing.from_pair(yw)
                assert torch.isfinite(got).all()
                e.append(rel_err(got.numpy(), want.numpy()))
            if R == 1:
                out = torch.relu(torch.randn(N, H, H, cin, generator=g))
                gres = torch.relu(torch.randn(J, H, H, cin, generator=g)) * 1e-3
                for mode in (0, 2):
                    go = torch.full((J, H, H, cin), float('nan'), device='cuda')
                    y3 = torch.full((J, H, H, cin), float('nan'), device='cuda')
                    be.dgrad_join(yp.cuda(), Ld, gres.cuda(), out.cuda(), o.cuda(), xr.cuda(), bn.cuda(), None, 2, mode, go, y3)
                    gw, y3w = torch.empty(J, H, H, cin), torch.empty(J, H, H, cin)
                    emul.dgrad_join(yp, L, gres, out, o, xr, bn, None, 2, mode, gw, y3w)
                    e.append(rel_err(go.cpu().numpy(), gw.numpy()))
                    e.append(rel_err(packing.from_pair(y3.cpu()).numpy(), packing.from_pair(y3w).numpy()))
            errs[key] = e
    torch.cuda.synchronize()
    print('\n'.join('%-28s %s' % (str(k), ' '.join('%.1e' % x for x in v)) for k, v in errs.items()))
    bad = {k: v for k, v in errs.items() if not max(v) < 2e-4}
    assert not bad, bad


def _gpu_engine(layers, impl='bf16x2'):
    from xfr_b200.kernels import CudaBackend
    dev = torch.device('cuda:0')
    return StResnetEngine(synth.stresnet_state_dict(0, layers, 2), CudaBackend(dev, impl=impl), layers, device=dev), dev


@pytest.mark.gpu
@pytest.mark.parametrize('layers', [L1111, L101])
def test_sweep_vs_reference_gpu(layers):
    """The fused sweep on the bf16x2 plan against the reference's own outputs: the north star's 1e-4 max-abs bar on every map
    (measured margin reported), single EBP maps to 2e-3 of their maximum, the well-conditioned ResNet-101 triplet scale-aware."""
    G = golden(layers)
    eng, dev = _gpu_engine(layers)
    x, W2, _ = golden_inputs(G)
    x, W2 = x.to(dev), W2.to(dev)
    P1 = torch.zeros(2, 2, device=dev)
    P1[:, 0] = 1
    rep = {}
    for mode, tag in (('affineonly_with_prior', 'awp'), ('all', 'all'), ('affineonly', 'affineonly')):
        s = eng.ebp(x, P1, W2, mode).cpu().numpy()
        c = eng.contrastive(x, W2, mode=mode).cpu().numpy()
        t = eng.contrastive(x, W2, mode=mode, percentile=20).cpu().numpy()
        for i, p in enumerate(('smooth', 'noise')):
            rep['ebp_%s_%s' % (tag, p)] = (np.abs(s[i] - G['ebp_%s_%s' % (tag, p)]).max(), rel_err(s[i], G['ebp_%s_%s' % (tag, p)]))
            rep['cebp_%s_%s' % (tag, p)] = (np.abs(c[i] - G['cebp_%s_%s' % (tag, p)]).max(), rel_err(c[i], G['cebp_%s_%s' % (tag, p)]))
            rep['tcebp20_%s_%s' % (tag, p)] = (np.abs(t[i] - G['tcebp20_%s_%s' % (tag, p)]).max(), rel_err(t[i], G['tcebp20_%s_%s' % (tag, p)]))
    if layers == L101:
        Gw = np.load(GOLD + '/stresnet101_wellcond_seed0.npz')
        W2w = torch.cat((torch.from_numpy(Gw['row_mate']), torch.from_numpy(Gw['row_nonmate']))).unsqueeze(0).repeat(2, 1, 1).contiguous().to(dev)
        for mode, tag in (('affineonly_with_prior', 'awp'), ('all', 'all')):
            c = eng.contrastive(x, W2w, mode=mode).cpu().numpy()
            t = eng.contrastive(x, W2w, mode=mode, percentile=20).cpu().numpy()
            for i, p in enumerate(('smooth', 'noise')):
                rep['WELL cebp_%s_%s' % (tag, p)] = (np.abs(c[i] - Gw['cebp_%s_%s' % (tag, p)]).max(), rel_err(c[i], Gw['cebp_%s_%s' % (tag, p)]))
                rep['WELL tcebp20_%s_%s' % (tag, p)] = (np.abs(t[i] - Gw['tcebp20_%s_%s' % (tag, p)]).max(), rel_err(t[i], Gw['tcebp20_%s_%s' % (tag, p)]))
    print('\n'.join('%-28s max-abs %.3g   max-abs/max(ref) %.3g' % (k, v[0], v[1]) for k, v in sorted(rep.items())))
    for k, (a, r) in rep.items():
        assert a < 5e-5, (k, a)                            # north-star bar 1e-4 max-abs, with margin
        if k.startswith('ebp'):
            assert r < 3e-3, (k, r)
        if k.startswith('WELL'):
            assert r < 1e-2, (k, r)
