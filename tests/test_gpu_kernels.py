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
TOL = {'fp32': 2e-5, 'tf32x3': 2e-4, 'tf32': 3e-2}


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
    eng_cpu = StResnetEngine(sd, EmulBackend(), L1111)
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
