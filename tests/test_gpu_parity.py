"""GPU: end-to-end parity of the CUDA path with the reference's own outputs (tests/golden, produced by
oracle/gen_golden.py from the unmodified reference) and with the oracle on fresh seeded inputs."""
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


@pytest.mark.parametrize('impl', ['fp32', 'tf32x3'])
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
