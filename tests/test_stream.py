"""Whitebox.contrastive_ebp_stream: the streaming form of contrastive_ebp_batch (copies overlapped with the sweeps) returns what the
batch call returns.  CPU: through the kernel emulation (tests/emul_backend.py, TEST INFRASTRUCTURE); GPU: bit-identical maps."""
import numpy as np
import pytest
import torch

from helpers import L1111
from xfr_b200 import synth, whitebox


def _rows(n, seed):
    g = torch.Generator().manual_seed(seed)
    return torch.randn(n, 512, generator=g) / 50.0, torch.randn(n, 512, generator=g) / 50.0


def test_stream_matches_batch_emulated():
    from emul_backend import EmulBackend
    from xfr_b200.engine import StResnetEngine

    class Net(whitebox.WhiteboxSTResnet):
        def engine(self, with_bias=False):
            if self._engine is None:
                self._engine = StResnetEngine(self._sd, EmulBackend(), self._layers, with_bias=with_bias)
            return self._engine

        def _device(self):
            return torch.device('cpu')
    net = Net(synth.stresnet_state_dict(0, L1111, 2), layers=L1111)
    wb = whitebox.Whitebox(net)
    xa, xb = synth.synthetic_probes(2, seed=5), synth.smooth_probes(1, seed=6)
    ra, rb = _rows(2, 1), _rows(1, 2)
    got = wb.contrastive_ebp_stream([(xa, ra[0], ra[1]), (xb, rb[0], rb[1])])
    net.set_triplet_classifiers(*ra)
    wa = wb.contrastive_ebp_batch(xa)
    net.set_triplet_classifiers(*rb)
    wbm = wb.contrastive_ebp_batch(xb)
    assert len(got) == 2 and got[0].shape == (2, 112, 112) and got[1].shape == (1, 112, 112)
    assert np.array_equal(got[0].numpy(), wa) and np.array_equal(got[1].numpy(), wbm)
    assert wb.contrastive_ebp_stream([]) == []


@pytest.mark.gpu
@pytest.mark.parametrize('percentile', [None, 20])
def test_stream_matches_batch_gpu(percentile):
    dev = torch.device('cuda:0')
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, L1111, 2).items()}
    net = whitebox.WhiteboxSTResnet(sd, layers=L1111)
    wb = whitebox.Whitebox(net)
    old = whitebox._CHUNK
    whitebox._CHUNK = 4                                   # several chunks per batch: both staging buffers and the ragged tail are used
    try:
        xs = [synth.synthetic_probes(10, seed=7).pin_memory(), synth.smooth_probes(3, seed=8).pin_memory(), synth.synthetic_probes(4, seed=9)]
        rows = [_rows(10, 3), _rows(3, 4), _rows(4, 5)]
        want = []
        for x, r in zip(xs, rows):
            net.set_triplet_classifiers(r[0].to(dev), r[1].to(dev))
            want.append(wb.contrastive_ebp_batch(x, percentile=percentile).copy())
        for _ in range(2):                                # the second pass reuses the staging buffers (and their events)
            got = wb.contrastive_ebp_stream([(x, r[0].to(dev), r[1].to(dev)) for x, r in zip(xs, rows)], percentile=percentile)
            for g, w in zip(got, want):
                assert np.array_equal(g.numpy(), w)
        assert np.abs(want[0]).max() > 0
    finally:
        whitebox._CHUNK = old
