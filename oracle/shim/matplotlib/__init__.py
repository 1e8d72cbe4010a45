"""Import shim (TEST INFRASTRUCTURE ONLY): matplotlib is absent from this image.  Lets oracle/gen_golden_inpaintgame.py import
the reference's python/xfr/inpainting_game/generate_whitebox_saliency.py, whose job functions (lines 45-219) never plot."""


def use(*a, **k):
    pass
