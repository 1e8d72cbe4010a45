"""Import shim (TEST INFRASTRUCTURE ONLY): see matplotlib/__init__.py.  Nothing here is called by the golden generators."""


def __getattr__(name):
    raise NotImplementedError('oracle shim: matplotlib.pyplot.%s is not available (plotting is outside the hot path)' % name)
