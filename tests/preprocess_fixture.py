"""Seeded PIL test images for the preprocess goldens (TEST INFRASTRUCTURE): shared by oracle/gen_golden_preprocess.py and
tests/test_preprocess.py."""
import numpy as np
import PIL.Image


def _smooth(h, w, seed):
    rng = np.random.RandomState(seed)
    coarse = rng.rand(h // 16 + 2, w // 16 + 2, 3)
    big = np.kron(coarse, np.ones((16, 16, 1)))[:h, :w]
    return np.clip(big * 255 + rng.randn(h, w, 3) * 6, 0, 255).astype(np.uint8)


def test_images():
    """Portrait, landscape, the 224x224 size of the inpainting-game PNGs, and a small image that is upsampled."""
    return {'portrait': PIL.Image.fromarray(_smooth(300, 260, 1)), 'landscape': PIL.Image.fromarray(_smooth(250, 330, 2)),
            'square224': PIL.Image.fromarray(_smooth(224, 224, 3)), 'small': PIL.Image.fromarray(_smooth(96, 120, 4))}


test_images.__test__ = False      # not a pytest test
