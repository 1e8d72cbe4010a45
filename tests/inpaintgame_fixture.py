"""Seeded inputs of the inpainting-game scoring goldens (TEST INFRASTRUCTURE): shared by oracle/gen_golden_inpaintgame.py,
which feeds them to the reference, and tests/test_inpaintgame.py, which feeds them to this package."""
import numpy as np
import torch

from xfr_b200 import synth

# run_inpainting_game_eval.py:127-134: the evaluation's percentile grids
PCT_DENSITY = np.unique(np.sort(np.append(np.arange(0, 100, 1), [0, 100])))                        # 101, the standard
PCT_PIXELS = np.unique(np.sort(np.append(100 * np.exp(-np.arange(0, 15, 0.1)), [0, 100])))         # 152


def scoring_fixture():
    imgs = synth.smooth_probes(2, seed=31).numpy()              # [2,3,224,224] float32, network format (mean-subtracted)
    yy, xx = np.mgrid[0:224, 0:224].astype(np.float64)
    blob = np.exp(-((yy - 90) ** 2 + (xx - 120) ** 2) / (2 * 30.0 ** 2)) + 0.4 * np.exp(-((yy - 170) ** 2 + (xx - 60) ** 2) / (2 * 18.0 ** 2))
    rng = np.random.RandomState(7)
    smap = (blob * (0.5 + rng.rand(224, 224))).astype(np.float32)
    smap /= smap.sum()                                          # plot_inpainting_game.py:1013-1014
    sparse = smap.copy()
    sparse[blob < 0.2] = 0
    sparse /= sparse.sum()
    return {'orig': imgs[0], 'inp': imgs[1], 'smap': smap, 'smap_sparse': sparse}


def images(n, seed):
    """n smooth 224x224x3 uint8 images (the inpainting-game images are 224x224 PNGs)."""
    x = synth.smooth_probes(n, seed=seed) + torch.tensor(synth.MEAN_RGB).view(1, 3, 1, 1)
    return [np.ascontiguousarray(im.permute(1, 2, 0).numpy().astype(np.uint8)) for im in x]


def jobs():
    """Three (im_mates, im_nonmates, probe_im) jobs, ragged: 3 mates / 1 non-mate, 1 mate / 2 non-mates, 1 / 1 with a probe
    that is also job 0's mate."""
    im = images(9, seed=11)
    return [(im[0:3], im[3:4], im[4]), (im[5:6], im[6:8], im[8]), (im[2:3], im[7:8], im[0])]
