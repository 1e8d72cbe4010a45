"""net.preprocess() of the three plugins and Whitebox.convert_from_numpy against the reference's own outputs
(tests/golden/preprocess_seed0.npz from oracle/gen_golden_preprocess.py; reference whitebox.py:108-110, 137-139, 235-258,
787-806): bit-exact (SHA-256 of the float32 tensor) - both sides call the same PIL resampling."""
import hashlib
import os

import numpy as np
import pytest
import torch

from helpers import GOLD
from preprocess_fixture import test_images
from xfr_b200 import whitebox


def _digest(t):
    a = np.ascontiguousarray(t.detach().numpy())
    assert a.dtype == np.float32
    return hashlib.sha256(a.tobytes()).hexdigest(), a


@pytest.fixture(scope='module')
def plugins():
    # preprocess() reads no weights: state-dict-less instances (the engines are never built)
    mk = lambda cls: cls.__new__(cls)
    return {'stresnet': mk(whitebox.WhiteboxSTResnet), 'resnet50_128': mk(whitebox.Whitebox_resnet50_128),
            'lightcnn': mk(whitebox.WhiteboxLightCNN), 'senet50_256': mk(whitebox.Whitebox_senet50_256)}


def test_preprocess_bit_exact_vs_reference(plugins):
    G = np.load(os.path.join(GOLD, 'preprocess_seed0.npz'))
    for name, im in test_images().items():
        for plug in ('stresnet', 'resnet50_128', 'lightcnn'):
            key = '%s_%s' % (plug, name)
            d, a = _digest(plugins[plug].preprocess(im))
            assert tuple(a.shape) == tuple(G[key + '_shape']), key
            assert abs(a.astype(np.float64).sum() - G[key + '_stats'][0]) <= 1e-6 * abs(G[key + '_stats'][0]), key
            assert d == str(G[key + '_sha256']), key
        # the SENet plugin shares the VGGFace2 preprocessing (reference whitebox.py:185-208 == 235-258)
        assert _digest(plugins['senet50_256'].preprocess(im))[0] == str(G['resnet50_128_%s_sha256' % name])


def test_convert_from_numpy_bit_exact_vs_reference(plugins):
    G = np.load(os.path.join(GOLD, 'preprocess_seed0.npz'))
    wb = whitebox.Whitebox.__new__(whitebox.Whitebox)
    torch.nn.Module.__init__(wb)
    wb.net = plugins['stresnet']
    im = np.array(test_images()['square224'])
    assert _digest(wb.convert_from_numpy(im))[0] == str(G['from_numpy_u8_square224_sha256'])
    assert _digest(wb.convert_from_numpy(im.astype(np.float32) / 255))[0] == str(G['from_numpy_f32_square224_sha256'])
    assert _digest(wb.convert_from_numpy(im.astype(np.float64)))[0] == str(G['from_numpy_f64_255_square224_sha256'])   # [0,255] floats
    with pytest.raises(ValueError):
        wb.convert_from_numpy(im.astype(np.float32) - 300.0)
