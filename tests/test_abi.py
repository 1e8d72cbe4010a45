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


def test_tile_geometry_host_logic():
    """The 4-D TMA box search of the tcgen05 conv kernel (host code, runs without a GPU)."""
    import ctypes
    from xfr_b200 import kernels
    lib = kernels.load_library()

    def geo(H, W, N):
        bh, bimg = ctypes.c_int(), ctypes.c_int()
        eff = lib.xfrb_tile_geometry(H, W, N, ctypes.byref(bh), ctypes.byref(bimg))
        assert 1 <= bh.value <= H and bimg.value >= 1 and bh.value * W * bimg.value <= 128
        return bh.value, bimg.value, eff
    assert geo(56, 56, 256)[:2] == (2, 1) and abs(geo(56, 56, 256)[2] - 0.875) < 1e-9
    assert geo(28, 28, 256)[:2] == (4, 1)
    bh, bimg, eff = geo(14, 14, 256)                     # 1 image row of 9 images: 126 of 128 rows, 29 image groups
    assert (bh, bimg) == (1, 9) and abs(eff - (126 / 128) * (256 / 261)) < 1e-9
    bh, bimg, eff = geo(7, 7, 256)
    assert (bh, bimg) == (1, 18) and eff > 0.93
    bh, bimg, eff = geo(14, 14, 1)                       # a single image: two strips either way, 98 of 128 rows on average
    assert bimg == 1 and bh in (7, 9) and abs(eff - 98 / 128) < 1e-9
    assert geo(64, 64, 128)[:2] == (2, 1) and geo(128, 128, 8)[:2] == (1, 1) and geo(16, 16, 128)[:2] == (8, 1)   # Light-CNN maps: 100 %
    assert lib.xfrb_tile_geometry(14, 200, 4, ctypes.byref(ctypes.c_int()), ctypes.byref(ctypes.c_int())) == -1.0
