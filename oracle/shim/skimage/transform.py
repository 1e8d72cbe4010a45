"""Import shim (TEST INFRASTRUCTURE ONLY): see skimage/__init__.py."""
import numpy as np


def resize(image, output_shape, *a, **k):
    """Only the identity case: Whitebox.convert_from_numpy (reference whitebox.py:802) resizes every image to 224x224 with
    preserve_range=True, a no-op on the 224x224 images of the inpainting-game flow in every scikit-image version (the pixel
    centres coincide).  Real resampling is not restated: parity is unpinned there."""
    image = np.asarray(image)
    if tuple(image.shape[:2]) != tuple(output_shape[:2]) or not k.get('preserve_range', False):
        raise NotImplementedError('oracle shim: skimage.transform.resize is only the identity here (%s -> %s)'
                                  % (image.shape, tuple(output_shape)))
    return image.astype(np.float64)
