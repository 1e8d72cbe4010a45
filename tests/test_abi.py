"""CPU: libxfr_b200.so builds, loads and exports every symbol include/xfrb.h declares (no compute calls)."""
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope='module')
def lib():
    from xfr_b200 import build, kernels
    build.build()
    return kernels.load_library()


def test_header_symbols_exported(lib):
    hdr = open(os.path.join(ROOT, 'include', 'xfrb.h')).read()
    names = set(re.findall(r'\b(xfrb_[a-z0-9_]+)\s*\(', hdr))
    assert len(names) >= 15
    from xfr_b200 import kernels
    assert names == set(kernels.EXPORTS)
    for n in names:
        assert hasattr(lib, n), n
    assert lib.xfrb_version() >= 1


def test_no_cpu_fallback():
    import torch
    from xfr_b200 import kernels
    if torch.cuda.is_available():
        pytest.skip('GPU present')
    with pytest.raises(RuntimeError):
        kernels.CudaBackend('cpu')
