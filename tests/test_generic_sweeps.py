"""The machinery under the layer sweeps / weighted_subtree_ebp: device-resident prior tables (generic.PriorTable), hook chains, row
skipping and graph replay of the firing-by-firing sweeps.  CPU: the table's host logic and its layout against include/xfrb.h;
GPU: every shortcut is bit-identical to the plain one-launch-per-firing sweep it replaces."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest
import torch

from helpers import L1111
from xfr_b200 import generic, synth, whitebox

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_prior_entry_layout_matches_header():
    """PRIOR_DTYPE mirrors XfrbPriorEntry field by field (api.cu static_asserts the C++ twin against the header's size)."""
    hdr = open(os.path.join(ROOT, 'include', 'xfrb.h')).read()
    body = re.search(r'typedef struct XfrbPriorEntry \{(.*?)\} XfrbPriorEntry;', hdr, re.S).group(1)
    fields = re.findall(r'^\s*(?:const\s+)?(int|long long|float)\s*\*?\s*(\w+);', body, re.M)
    names = [n for _, n in fields]
    assert names == ['row', 'probe_row', 'elem', 'probe_elem', 'tensor', 'val', 'pad_', 'pad2_']
    dt = generic.PRIOR_DTYPE
    assert dt.itemsize == 48 and [dt.fields[n][1] for n in dt.names] == [0, 4, 8, 16, 24, 32, 36, 40]


def test_prior_table_host_logic():
    t = generic.PriorTable(16, 'cpu')
    assert t.legacy(3) is None and not t.zero_seed
    x = torch.arange(12, dtype=torch.float32)
    t.clear(zero_seed=True)
    t.set_elem(2, 0, 7, 0.5)
    t.set_tensor(5, 1, x)
    t.set_probe(9, 0, 11)
    t.upload()
    assert t.legacy(2) == (0, 7, 0.5) and t.legacy(5)[0] == 1 and t.legacy(5)[1] is x and t.legacy(9) is None
    st = t.start.numpy()
    assert st[0] == 2 and st[1] == 5 and (st[2:] == generic.NEVER).all()           # rows without a prior never start
    raw = np.frombuffer(t.dev.numpy().tobytes(), dtype=generic.PRIOR_DTYPE)
    assert raw['row'][2] == 0 and raw['elem'][2] == 7 and raw['tensor'][5] == x.data_ptr() and raw['probe_elem'][9] == 11
    assert raw['row'][0] == -1 and raw['probe_row'][0] == -1
    t.clear()
    t.set_elem(1, 0, 0, 1.0)
    t.set_elem(2, 0, 0, 1.0)
    t.zero_seed = True
    with pytest.raises(AssertionError):                                              # two priors on one gradient row
        t.upload()


def _setup(dev):
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, L1111, 2).items()}
    net = whitebox.WhiteboxSTResnet(sd, layers=L1111, impl='tf32x3')
    x = synth.smooth_probes(1, seed=1).to(dev)
    g = torch.Generator().manual_seed(3)
    net.set_triplet_classifier(torch.randn(1, 512, generator=g) / 50, torch.randn(1, 512, generator=g) / 50)
    eng = net.engine()
    eng.forward(net._nhwc(x))
    return net, eng, net.triplet_rows(1)


@pytest.mark.gpu
@pytest.mark.parametrize('mode', ['affineonly_with_prior', 'all', 'norelu'])
def test_chains_rows_and_graphs_are_bit_identical(mode):
    """One 11-row zero-seeded prior sweep four ways: (a) one launch per firing, priors as launch arguments (the round-1 form);
    (b) device table, no chains; (c) chains + row skipping; (d) the same replayed from a captured graph.  Plus a recording sweep
    with and without chains."""
    dev = torch.device('cuda:0')
    net, eng, W2 = _setup(dev)
    gs = eng.sweep()
    P0 = torch.zeros(1, 2, device=dev)
    P0[0, 0] = 1
    gs.chain_hooks = False
    P, names, P2ref = gs.run(P0, W2, mode, record=True)
    P2ref = P2ref.clone()
    gs2 = eng.sweep()
    Pc, _, P2c = gs2.run(P0, W2, mode, record=True)                                 # chains on, recording every firing
    assert torch.equal(P2c, P2ref)
    for a, b in zip(P, Pc):
        assert (a is None and b is None) or torch.equal(a, b)
    ks = [k for k in (1, 4, 9, 12, 17, 21, 26, 30, 37, 44, len(P) - 3) if P[k] is not None]     # > 8 rows: the row-walk kernel, a ragged last group
    Z = torch.zeros(len(ks), 2, device=dev)
    pri = {}
    for r, k in enumerate(ks):
        flat = P[k].reshape(-1)
        e = int(torch.argmax(flat))
        pri[k] = (r, e, float(flat[e]))
    gs.chain_hooks = False
    _, _, want = gs.run(Z, W2, mode, priors=pri)                                     # (a)
    want = want.clone()
    assert float(want.abs().max()) > 0
    tab = eng.prior_table('test')
    for chains, zero_seed in ((False, False), (True, False), (True, True)):          # (b), chains alone, (c)
        tab.clear(zero_seed=zero_seed)
        for k, (r, e, v) in pri.items():
            tab.set_elem(k, r, e, v)
        tab.upload()
        g3 = eng.sweep()
        g3.chain_hooks = chains
        _, _, got = g3.run(Z, W2, mode, ptab=tab)
        assert torch.equal(got, want), (chains, zero_seed)
    for _ in range(3):                                                               # (d): eager, capture, replay
        got = eng.generic_call(Z, W2, mode=mode, ptab=tab)['P2']
        assert torch.equal(got, want)
    probe = eng.prior_table('test_probe')                                            # probes: p of one element per firing
    probe.clear()
    for k, (r, e, v) in pri.items():
        probe.set_probe(k, 0, e)
    probe.upload()
    eng.sweep().run(P0, W2, mode, ptab=probe)
    torch.cuda.synchronize()
    for k, (r, e, v) in pri.items():
        assert float(probe.probe[k]) == v


@pytest.mark.gpu
def test_scalar_hook_kernel_matches_vector_kernel():
    """XFRB_HOOK_SCALAR=1 (read once per process) routes every firing through hook_kernel<1, .>, the path of channel counts that
    are not multiples of four: same maps to the last bit."""
    code = ("import sys, torch, numpy as np; sys.path[:0] = ['.', 'tests']\n"
            "from test_generic_sweeps import _setup\n"
            "net, eng, W2 = _setup(torch.device('cuda:0'))\n"
            "P0 = torch.zeros(1, 2, device='cuda'); P0[0, 0] = 1\n"
            "_, _, P2 = eng.sweep().run(P0, W2, 'all')\n"
            "np.save(sys.argv[1], P2.cpu().numpy())\n")
    outs = []
    for flag in ('0', '1'):
        path = os.path.join(ROOT, 'gpurun_out', 'hook_scalar_%s.npy' % flag)
        os.makedirs(os.path.dirname(path), exist_ok=True)
        env = dict(os.environ, XFRB_HOOK_SCALAR=flag)
        subprocess.run([sys.executable, '-c', code, path], cwd=ROOT, env=env, check=True, timeout=300)
        outs.append(np.load(path))
    assert np.array_equal(outs[0], outs[1]) and np.abs(outs[0]).max() > 0
