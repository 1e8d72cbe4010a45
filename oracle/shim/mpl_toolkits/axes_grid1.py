"""Import shim (TEST INFRASTRUCTURE ONLY): see matplotlib/__init__.py."""


def make_axes_locatable(*a, **k):
    raise NotImplementedError('oracle shim: plotting is outside the hot path')
