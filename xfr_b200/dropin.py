"""Alternative to the PYTHONPATH overlay (dropin/xfr): `import xfr_b200.dropin` BEFORE the first `import xfr.models.whitebox`
aliases the reference's module name to the B200 engine in sys.modules.  The reference package itself must be importable."""
import importlib
import sys

from . import whitebox as _wb


def install():
    models = importlib.import_module('xfr.models')          # the reference's package (python/xfr/models)
    sys.modules['xfr.models.whitebox'] = _wb
    models.whitebox = _wb
    return _wb


install()
