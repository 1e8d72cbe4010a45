"""Drop-in overlay of the reference package `xfr` (stresearch/xfr, python/xfr): put this directory IN FRONT of the
reference's python/ directory on PYTHONPATH and every caller - demo/test_whitebox.py, eval/create_wbnet.py,
xfr/inpainting_game/generate_whitebox_saliency.py, eval/generate_inpaintinggame_wb_saliency_maps_multigpu.py - runs
unchanged, with `xfr.models.whitebox` served by the B200 engine (xfr_b200.whitebox) and everything else (xfr.models.resnet,
xfr.utils, xfr.show, xfr.inpainting_game, ...) still coming from the reference tree:

    PYTHONPATH=/path/to/xfr_b200_repo/dropin:/path/to/xfr/python python demo/test_whitebox.py

How: this file runs the reference's own xfr/__init__.py (which computes xfr_root from __path__[0], reference
python/xfr/__init__.py:7) with the reference directory first on the package path, then puts the overlay's models/ directory
in front of xfr.models.__path__, so that `from xfr.models import whitebox` finds dropin/xfr/models/whitebox.py.
No reference source is copied or modified.
"""
import os
import sys

_here = os.path.dirname(os.path.abspath(__file__))
_repo = os.path.dirname(os.path.dirname(_here))
if _repo not in sys.path:
    sys.path.append(_repo)                  # the xfr_b200 package


def _reference_dir():
    for p in sys.path:
        d = os.path.join(os.path.abspath(p) if p else os.getcwd(), 'xfr')
        if d != _here and os.path.isfile(os.path.join(d, '__init__.py')):
            return d
    raise ImportError('xfr drop-in overlay: the reference package (stresearch/xfr, python/xfr) must be on sys.path behind %s'
                      % os.path.dirname(_here))


_ref = _reference_dir()
__path__ = [_ref, _here]                    # reference first: xfr_root and every other submodule resolve there
with open(os.path.join(_ref, '__init__.py')) as _f:
    exec(compile(_f.read(), os.path.join(_ref, '__init__.py'), 'exec'))

import xfr.models as _models  # noqa: E402  (the reference's models package)

_models.__path__.insert(0, os.path.join(_here, 'models'))      # ... except xfr.models.whitebox
