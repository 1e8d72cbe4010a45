"""Import shim (TEST INFRASTRUCTURE ONLY): see skimage/__init__.py."""
import numpy as np


def rgb2gray(rgb):
    rgb = np.asarray(rgb)
    if rgb.dtype == np.uint8:
        rgb = rgb.astype(np.float64) / 255
    return rgb @ np.array([0.2125, 0.7154, 0.0721])
