def resize(*a, **k):
    raise NotImplementedError('oracle shim: skimage.transform.resize is not on the whitebox hot path')
