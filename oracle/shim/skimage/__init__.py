"""Import shim (TEST INFRASTRUCTURE ONLY): scikit-image is absent from this image.

Lets `oracle/gen_golden.py` import the unmodified reference package from
/root/reference/python.  Only `skimage.filters.gaussian` is on the whitebox
path (reference whitebox.py:457); it is restated with scipy.ndimage, which is
what scikit-image (>=0.17.2 per the reference README.md:37) calls underneath.
"""
from . import filters, transform, color, morphology  # noqa: F401
