"""Import shim (TEST INFRASTRUCTURE ONLY): imageio is absent; not on the whitebox hot path."""


def imread(*a, **k):
    raise NotImplementedError('oracle shim')


def imwrite(*a, **k):
    raise NotImplementedError('oracle shim')
