"""Where a layer sweep / weighted-subtree job spends its time: wall-clock per phase (synchronised) and the kernels of one call by
total device time (torch profiler).  python tools/generic_profile.py [layer_sweep|weighted_subtree]"""
import collections
import sys
import time

import torch

sys.path.insert(0, '.')
from xfr_b200 import synth, whitebox  # noqa: E402

AFFINE = ('Conv', 'Linear', 'AvgPool', 'BatchNorm')


def main():
    what = sys.argv[1] if len(sys.argv) > 1 else 'layer_sweep'
    dev = torch.device('cuda:0')
    if what == 'lightcnn':
        return lightcnn(dev)
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
    net = whitebox.WhiteboxSTResnet(sd, impl=sys.argv[2] if len(sys.argv) > 2 else whitebox.DEFAULT_IMPL)
    wb = whitebox.Whitebox(net, ebp_subtree_mode='affineonly_with_prior' if what == 'layer_sweep' else 'norelu')
    x = synth.synthetic_probes(1, seed=100).to(dev)
    with torch.no_grad():
        enc = net.encode(synth.synthetic_probes(2, seed=1000).to(dev))
    net.set_triplet_classifier(enc[0:1] / 2500.0, enc[1:2] / 2500.0)
    if what == 'layer_sweep':
        wb.layerwise_contrastive_ebp_sweep(x, 0, 1, [3], mode='percentile', percentile=20)
        ks = [k for k, n in enumerate(wb.P_layername[:-1]) if any(a in n for a in AFFINE)]
        call = lambda: wb.layerwise_contrastive_ebp_sweep(x, 0, 1, ks, mode='percentile', percentile=20)
    else:
        call = lambda: wb.weighted_subtree_ebp(x, 0, 1, topk=32, verbose=False, do_mated_similarity_gating=False, subtree_mode='all')
    for _ in range(6):          # eager, capture, ... : the graphs of a call settle after a few calls (a grown workspace buffer drops them once)
        call()
    phases = collections.OrderedDict()
    eng = net.engine(wb._ebp_with_bias)

    def wrap(obj, name, label):
        fn = getattr(obj, name)

        def timed(*a, **kw):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            r = fn(*a, **kw)
            torch.cuda.synchronize()
            phases[label] = phases.get(label, 0.0) + (time.perf_counter() - t0)
            return r
        setattr(obj, name, timed)
    wrap(eng, 'forward', 'forward')
    wrap(eng, 'generic_call', 'sweeps (graph replays)')
    wrap(eng, 'graph_fn', 'priors (graph replay)')
    wrap(wb, '_finish_map', 'finish maps')
    caps = eng.graph_captures
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    tot = time.perf_counter() - t0
    print('%s: %.1f ms per call (%d graph captures during it, %d before)' % (what, tot * 1e3, eng.graph_captures - caps, caps))
    for k, v in phases.items():
        print('  %-28s %7.1f ms' % (k, v * 1e3))
    print('  %-28s %7.1f ms' % ('other (host)', (tot - sum(phases.values())) * 1e3))
    for n in ('forward', 'generic_call', 'graph_fn'):
        delattr(eng, n)
    delattr(wb, '_finish_map')
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        call()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[e.name[:70]]
            a[0] += 1
            a[1] += e.device_time
    total = sum(v[1] for v in agg.values())
    print('device time %.1f ms in %d kernels' % (total / 1e3, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:14]:
        print('  %-70s %5d %9.1f us %5.1f %%' % (k, v[0], v[1], 100 * v[1] / total))


def kernel_table(call):
    from torch.profiler import ProfilerActivity, profile
    with profile(activities=[ProfilerActivity.CUDA]) as prof:
        call()
        torch.cuda.synchronize()
    agg = collections.defaultdict(lambda: [0, 0.0])
    for e in prof.events():
        if e.device_type == torch.autograd.DeviceType.CUDA:
            a = agg[e.name[:70]]
            a[0] += 1
            a[1] += e.device_time
    total = sum(v[1] for v in agg.values())
    print('device time %.1f ms in %d kernels' % (total / 1e3, sum(v[0] for v in agg.values())))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1])[:16]:
        print('  %-70s %5d %9.1f us %5.1f %%' % (k, v[0], v[1], 100 * v[1] / total))


def lightcnn(dev):
    sd = {k: v.to(dev) for k, v in synth.lightcnn_state_dict(0, 2).items()}
    net = whitebox.WhiteboxLightCNN(sd, impl='tf32x3')
    wb = whitebox.Whitebox(net, ebp_subtree_mode='affineonly')
    B = 128
    x = synth.lightcnn_probes(B, seed=3, smooth=False).to(dev)
    W2 = torch.randn(B, 2, 256, generator=torch.Generator().manual_seed(4)).to(dev)
    net.set_triplet_classifiers(W2[:, 0], W2[:, 1])
    P = torch.zeros(1, 2)
    P[0, 0] = 1
    call = lambda: wb.ebp_batch(x, P)
    for _ in range(3):
        call()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    call()
    torch.cuda.synchronize()
    print('lightcnn ebp, %d probes: %.1f ms per call' % (B, 1e3 * (time.perf_counter() - t0)))
    kernel_table(call)


if __name__ == '__main__':
    main()
