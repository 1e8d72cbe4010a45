"""Import shim (TEST INFRASTRUCTURE ONLY): see skimage/__init__.py."""
import numpy as np
import scipy.ndimage as ndi


def gaussian(image, sigma=1, mode='nearest', cval=0, truncate=4.0, **kw):
    """skimage.filters.gaussian == scipy.ndimage.gaussian_filter on a float image."""
    image = np.asarray(image)
    if image.dtype.kind != 'f':
        image = image.astype(np.float64)
    return ndi.gaussian_filter(image, sigma, mode=mode, cval=cval, truncate=truncate)
