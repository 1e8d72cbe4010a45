"""GPU: every CUDA kernel of libxfr_b200.so against the torch statement of the same kernel
(tests/emul_backend.py) on the tensors it actually sees inside a full sweep (ShadowBackend)."""
import pytest
import torch

from emul_backend import EmulBackend
from helpers import L1111, ShadowBackend, golden, golden_inputs, pack_map
from xfr_b200 import synth

pytestmark = pytest.mark.gpu

# GEMM-backed stages accumulate in a different order than the CPU BLAS: a few 1e-6; everything else is
# elementwise IEEE arithmetic and must agree to rounding of the reductions.
TOL = {'fp32': 2e-5, 'tf32x3': 2e-4, 'tf32x3full': 2e-4, 'tf32': 3e-2}


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
@pytest.mark.parametrize('mode', ['affineonly_with_prior', 'all', 'affineonly'])
def test_each_kernel_against_emulation(impl, mode):
    from xfr_b200.engine import StResnetEngine
    from xfr_b200.kernels import CudaBackend
    sd = synth.stresnet_state_dict(0, L1111, 2)
    dev = torch.device('cuda:0')
    try:
        cuda_be = CudaBackend(dev, impl=impl)
    except NotImplementedError:
        pytest.skip('impl %s not built' % impl)
    eng_cpu = StResnetEngine(sd, EmulBackend(impl_name=impl), L1111)
    eng = StResnetEngine(sd, cuda_be, L1111, device=dev)
    sh = ShadowBackend(cuda_be, EmulBackend(), pack_map(eng, eng_cpu))
    eng.be = sh
    G = golden(L1111)
    x, W2, _ = golden_inputs(G)
    eng.contrastive(x.to(dev), W2.to(dev), mode=mode)
    torch.cuda.synchronize()
    bad = {k: v for k, v in sh.errors.items() if v > TOL[impl]}
    print('\n'.join('%-18s %.3g' % kv for kv in sorted(sh.errors.items())))
    assert not bad, bad


def _conv_case(be, J, H, C_in, C_out, R, seed=0):
    """dgrad_plain == plain implicit GEMM: y [J,H,H,C_out-as-K] x Bd [C_in, R*R*K] -> [J,H,H,C_in]."""
    from emul_backend import im2col_nhwc
    g = torch.Generator().manual_seed(seed)
    y = torch.randn(J, H, H, C_out, generator=g)
    Bd = torch.randn(C_in, R * R * C_out, generator=g) / (R * (C_out ** 0.5))

    class L(object):
        pass
    from xfr_b200.packing import gemm_planes
    L.Bd, L.cin, L.R = gemm_planes(Bd, be.impl_name).cuda(), C_in, R
    out = torch.full((J, H, H, C_in), float('nan'), device='cuda')
    be.dgrad_plain(y.cuda(), L, out)
    torch.cuda.synchronize()
    want = (im2col_nhwc(y.double(), R, R, R // 2) @ Bd.double().t()).view(J, H, H, C_in)
    got = out.cpu().double()
    assert torch.isfinite(got).all()
    return float((got - want).abs().max() / want.abs().max())


# 'tf32x3' runs dgrad_plain as a W+ GEMM (two passes: weights rounded to TF32, activations exact), so on the SIGNED random
# operands of this test it carries the 2^-12 weight rounding; 'tf32x3full' (three passes) is the fp32-equivalent plan.
@pytest.mark.parametrize('impl,tol', [('fp32', 1e-5), ('tf32x3', 5e-4), ('tf32x3full', 1e-4), ('tf32', 3e-3)])
def test_gemm_shapes(impl, tol):
    """Every (spatial size, tap count, N width) family the ResNet-101 schedule launches, incl. ragged M tails."""
    from xfr_b200.kernels import CudaBackend
    try:
        be = CudaBackend('cuda:0', impl=impl)
    except NotImplementedError:
        pytest.skip('impl %s not built' % impl)
    cases = [  # J, H, C_in (GEMM N), C_out (A channels), R
        (3, 56, 64, 64, 1), (3, 56, 64, 64, 3), (2, 56, 256, 64, 1), (3, 28, 128, 128, 3), (3, 28, 512, 128, 1),
        (5, 14, 256, 256, 3), (5, 14, 1024, 256, 1), (5, 7, 512, 512, 3), (3, 7, 2048, 512, 1), (7, 1, 2048, 512, 1),
        (1, 7, 512, 2048, 1), (300, 1, 128, 64, 1),
        # enough tiles for the CTA-pair (cta_group::2) kernels, incl. an odd tile count (phantom half) and a ragged M tail
        (13, 56, 256, 64, 1), (32, 28, 128, 128, 3), (40, 14, 1024, 256, 1), (160, 7, 512, 512, 3), (24, 56, 64, 64, 3),
    ]
    errs = {}
    for c in cases:
        errs[c] = _conv_case(be, *c)
    print('\n'.join('%-28s %.3g' % (str(k), v) for k, v in errs.items()))
    bad = {k: v for k, v in errs.items() if not v < tol}
    assert not bad, bad


def test_cta_pairs_match_single_cta():
    """The cta_group::2 kernels accumulate in the same order as the single-CTA ones: a 16-probe ResNet-101 contrastive
    sweep (every conv launch large enough to run as CTA pairs) is bit-identical with pairs on and off."""
    from xfr_b200.engine import StResnetEngine
    from xfr_b200.kernels import CudaBackend
    from helpers import L101
    dev = torch.device('cuda:0')
    be = CudaBackend(dev, impl='tf32x3')
    eng = StResnetEngine(synth.stresnet_state_dict(0, L101, 2), be, L101, device=dev)
    N = 16
    x = synth.synthetic_probes(N, seed=31).permute(0, 2, 3, 1).contiguous().to(dev)
    g = torch.Generator().manual_seed(32)
    W2 = (torch.randn(N, 2, 512, generator=g) * 0.02).to(dev)
    prev_mc = be.lib.xfrb_set_multicast_pairs(0)
    prev = be.lib.xfrb_set_cta_pairs(1)
    try:
        a = eng.contrastive(x, W2, saliency=False).clone()
        be.lib.xfrb_set_cta_pairs(0)
        b = eng.contrastive(x, W2, saliency=False).clone()               # single-CTA kernels
        be.lib.xfrb_set_multicast_pairs(1)
        c = eng.contrastive(x, W2, saliency=False).clone()               # multicast pairs (the default)
    finally:
        be.lib.xfrb_set_cta_pairs(prev)
        be.lib.xfrb_set_multicast_pairs(prev_mc)
    assert torch.isfinite(a).all() and float(a.abs().max()) > 0
    assert torch.equal(a, b), float((a - b).abs().max() / b.abs().max())
    assert torch.equal(c, b), float((c - b).abs().max() / b.abs().max())
