#!/usr/bin/env python
"""Benchmark of the whitebox hot path: contrastive-EBP saliency maps/sec, STR ResNet-101, 224x224 probes.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--batch 256] [--gemm tf32x3|tf32|fp32]
  python -m torch.distributed.run --nnodes=1 --nproc-per-node N ... bench.py --gpus N ...
  python bench.py --impl reference ...      the reference algorithm on the host cores (oracle port)

One step = one pass of forward + mate/non-mate excitation backprop + contrastive combine + saliency
post-filter over `batch` synthetic (probe, mate, non-mate) triplets per GPU (BASELINE.json configs[1]).
Prints ONE JSON line (rank 0).  See DESIGN.md "measurement" for what each key means.
"""
import argparse
import json
import os
import subprocess
import sys
import threading
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
if '--impl' in sys.argv and 'reference' in sys.argv:
    # torchrun exports OMP_NUM_THREADS=1 to every rank; the reference arm is a CPU measurement on all host cores (rank 0 only)
    for _k in ('OMP_NUM_THREADS', 'MKL_NUM_THREADS'):
        os.environ.pop(_k, None)

import numpy as np  # noqa: E402
import torch  # noqa: E402

# algorithmic work per ResNet-101 contrastive map: fp32, mate + non-mate
EBP_BWD_BYTES_PER_MAP = 336e6     # 4*[2*S_o + 2*(S_o+S_i)],  S_i = 13.65 M, S_o = 14.20 M
FLOP_PER_MAP = 57.2e9             # 28.8 forward (true + positive) + 28.4 two W+ dgrad sweeps
METRIC = 'contrastive-EBP saliency maps/sec (ResNet-101, 224x224)'


def measured_peaks():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        d = json.load(open(p))
        return float(d['hbm_gbs']), 'measured (MEASURED_PEAKS.json)'
    return 6650.0, 'fallback (B200_PROFILING.md)'


class ClockSampler(threading.Thread):
    """nvidia-smi clocks/throttle reasons sampled every 200 ms while the timed region runs."""
    Q = ('clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,'
         'clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap')

    def __init__(self, index):
        super(ClockSampler, self).__init__(daemon=True)
        self.index, self.rows, self.stop_flag = index, [], False

    def run(self):
        while not self.stop_flag:
            try:
                out = subprocess.run(['nvidia-smi', '-i', str(self.index), '--query-gpu=' + self.Q, '--format=csv,noheader,nounits'],
                                     capture_output=True, text=True, timeout=5).stdout.strip()
                if out:
                    self.rows.append([c.strip() for c in out.split(',')])
            except Exception:
                pass
            time.sleep(0.2)

    def summary(self):
        self.stop_flag = True
        if not self.rows:
            return {'sm_mhz': None, 'sm_max_mhz': None, 'reasons': ['nvidia-smi unavailable']}
        sm = sorted(float(r[0]) for r in self.rows)
        names = ('hw_slowdown', 'hw_thermal_slowdown', 'sw_thermal_slowdown', 'sw_power_cap')
        reasons = [n for i, n in enumerate(names) if any(r[2 + i].lower().startswith('active') for r in self.rows)]
        return {'sm_mhz': sm[len(sm) // 2], 'sm_max_mhz': float(self.rows[0][1]), 'reasons': reasons, 'samples': len(sm)}


# DRAM bytes (read + write) of one layer3 launch (128-probe sweep) of each kernel family from `ncu --set full`, bf16x2 plan
# (profiles/r2_ncu_full_bf16x2.csv: JOIN dgrad 574 + 378 MB, 3x3 MID dgrad 107 + 25 MB, conv3 forward 66 + 152 MB - the 126 MB L2
# holds part of a launch's writes back past its end, so per-launch DRAM writes under-count)
NCU_TRAFFIC = {'dgrad_join': 952e6, 'dgrad_mid': 132e6, 'conv_dual': 218e6}
# tcgen05 issue peaks of this chip measured by tools/mma_peak.cu (profiles/r2_mma_peak.jsonl; operands resident in shared memory, M = 128
# per CTA, N = 256): kind::tf32 1,116 TFLOP/s, kind::f16 (bf16 operands) 2,232 TFLOP/s at 1.84 GHz - TF32 is exactly half of bf16
MMA_PEAK_TFLOPS = {'tf32': 1116.0, 'bf16': 2232.0}


def kernel_families(eng, step_fn, gemm='bf16x2'):
    """One extra (untimed) step with CUDA events around every GEMM-backed launch: per kernel family the launch count,
    average device time, algorithmic bytes / flops per launch and what that is of the HBM / tensor peak."""
    be = eng.be
    rec = []

    def rows(t):
        return t.numel() // t.shape[-1]

    def meta(name, a):
        if name == 'dgrad_join':      # y1, L, g_res, out, o3, xr3, bn3, res, hooks, mode, g_out, y3_out
            M, K, C, Ms = rows(a[0]), a[0].shape[-1], a[10].shape[-1], rows(a[3])
            return 4.0 * (3 * Ms * C + 3 * M * C + M * K), 2.0 * M * C * K * a[1].R ** 2
        if name == 'dgrad_mid':       # y, L, o, xr, bn, mode, y_out
            M, K, C, Ms = rows(a[0]), a[0].shape[-1], a[6].shape[-1], rows(a[2])
            return 4.0 * (2 * Ms * C + M * C + M * K), 2.0 * M * C * K * a[1].R ** 2
        if name == 'conv_dual':       # inp, L, o, xr, act, res
            M, K, C = rows(a[0]), a[0].shape[-1], a[2].shape[-1]
            return 4.0 * (M * K + 3 * M * C + (M * C if len(a) > 5 and a[5] is not None else 0)), 2.0 * 2 * M * C * K * a[1].R ** 2
        return 0.0, 0.0
    saved = {}
    for name in ('dgrad_join', 'dgrad_mid', 'conv_dual'):
        orig = getattr(be, name)
        saved[name] = orig

        def wrapped(*a, _orig=orig, _name=name, **k):
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            _orig(*a, **k)
            e1.record()
            rec.append((_name, e0, e1) + meta(_name, a) + (list(a[0].shape) + [a[1].cin, a[1].cout, a[1].R],))
        setattr(be, name, wrapped)
    try:
        step_fn(False, gather=False)          # rank 0 only: no collective in here
        torch.cuda.synchronize()
    finally:
        for name in saved:
            delattr(be, name)
    hbm, _ = measured_peaks()
    bf16 = gemm == 'bf16x2'
    mma_peak = MMA_PEAK_TFLOPS['bf16' if bf16 else 'tf32']
    dump = os.environ.get('XFRB_BENCH_LAUNCHES')          # per-launch rows (name, A shape, us, bytes, flops) for roofline studies
    if dump:
        with open(dump, 'w') as f:
            for r in rec:
                f.write(json.dumps({'k': r[0], 'us': 1e3 * r[1].elapsed_time(r[2]), 'bytes': r[3], 'flops': r[4], 'shape': r[5]}) + '\n')
    out = []
    sp = (4, 5) if bf16 else (2, 3)
    label = {'dgrad_join': 'W+ dgrad + JOIN hook chain (conv_tc_kernel<BN,%d,JOIN>)' % sp[0], 'dgrad_mid': 'W+ dgrad + MID hook chain (conv_tc_kernel<BN,%d,MID>)' % sp[0],
             'conv_dual': 'forward dual conv: o, xr, act (conv_tc_kernel<BN,%d,FWD_DUAL>)' % sp[1]}
    passes = {'dgrad_join': 2.0, 'dgrad_mid': 2.0, 'conv_dual': 2.5}      # MMA passes per useful flop (bf16x2: bf16 passes; split-TF32: TF32 passes)
    for name in ('dgrad_join', 'dgrad_mid', 'conv_dual'):
        rr = [r for r in rec if r[0] == name]
        if not rr:
            continue
        t = sum(r[1].elapsed_time(r[2]) for r in rr) * 1e-3
        by, fl = sum(r[3] for r in rr), sum(r[4] for r in rr)
        out.append({'kernel': label[name], 'launches': len(rr), 'avg_us': 1e6 * t / len(rr),
                    'alg_bytes_per_launch': by / len(rr), 'achieved_gbs': by / t / 1e9, 'hbm_frac': by / t / 1e9 / hbm,
                    'useful_tflops': fl / t / 1e12, 'issued_mma_tflops': passes[name] * fl / t / 1e12, 'mma_kind': 'bf16' if bf16 else 'tf32',
                    'mma_frac_of_measured_issue_peak': passes[name] * fl / t / 1e12 / mma_peak, 'ncu_dram_bytes_layer3_launch': NCU_TRAFFIC[name]})
    return out


def tensor_peak():
    p = os.path.join(ROOT, 'MEASURED_PEAKS.json')
    if os.path.exists(p):
        return float(json.load(open(p)).get('bf16_tflops', 1650.0))
    return 1650.0


def cpu_port(n_triplets, threads):
    """The reference algorithm restated on the CPU (oracle port, torch CPU fp32), batch 1 like the reference."""
    from oracle import stresnet_oracle as O      # cpu_baseline leg only
    from xfr_b200 import synth
    torch.set_num_threads(threads)
    sd = synth.stresnet_state_dict(0)
    x = synth.synthetic_probes(n_triplets + 1, seed=1)
    g = torch.Generator().manual_seed(5)
    W2 = torch.randn(n_triplets + 1, 2, 512, generator=g) * 0.02
    O.contrastive_ebp(sd, x[:1], W2[:1])                     # warm-up
    t0 = time.time()
    for i in range(1, n_triplets + 1):
        O.contrastive_ebp(sd, x[i:i + 1], W2[i:i + 1])
    return n_triplets / (time.time() - t0)


def gpu_library_baseline(dev, n_triplets=64, batch=16, allow_tf32=False):
    """SURVEY.md section 2a "the Blackwell kernel to beat": stock PyTorch eager (cuDNN / cuBLAS library kernels) running the
    same algorithm - the oracle restatement, batched, one shared forward + mate / non-mate sweeps - on the same B200.
    fp32 with TF32 off is the parity-grade setting; allow_tf32 shows what the library does with single-pass TF32.
    Outside every timed region of the product; the host-side Gaussian post-filter is left out (it favours this arm)."""
    from oracle import stresnet_oracle as O      # comparator leg only
    from xfr_b200 import synth
    old = (torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32)
    torch.backends.cudnn.allow_tf32 = torch.backends.cuda.matmul.allow_tf32 = bool(allow_tf32)
    try:
        sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
        x = synth.synthetic_probes(batch, seed=1).to(dev)
        g = torch.Generator().manual_seed(5)
        W2 = (torch.randn(batch, 2, 512, generator=g) * 0.02).to(dev)
        O.contrastive_mwp(sd, x, W2)                 # warm-up (cuDNN autotune, allocator)
        torch.cuda.synchronize()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(max(1, n_triplets // batch)):
            O.contrastive_mwp(sd, x, W2)
        e1.record()
        torch.cuda.synchronize()
        n = max(1, n_triplets // batch) * batch
        return {'value': n / (e0.elapsed_time(e1) * 1e-3), 'unit': 'maps/s', 'kind': 'torch %s eager (cuDNN/cuBLAS), oracle restatement batched' % torch.__version__,
                'batch': batch, 'triplets': n, 'precision': 'tf32 (library single pass)' if allow_tf32 else 'fp32 (TF32 off)'}
    finally:
        torch.backends.cudnn.allow_tf32, torch.backends.cuda.matmul.allow_tf32 = old
        torch.cuda.empty_cache()


def batch1_latency(wb, net, dev, x_host, enc):
    """Every reference caller is batch 1 (whitebox.py:482-527, generate_whitebox_saliency.py:81-205): wall-clock latency of one
    Whitebox.contrastive_ebp / truncated_contrastive_ebp / weighted_subtree_ebp(topk=32) call through the public API with a
    host-resident probe (H2D of the probe and D2H of the map inside), median of several calls after warm-up."""
    from xfr_b200 import whitebox
    res = {}
    net.set_triplet_classifier(enc[0:1] / 2500.0, enc[1:2] / 2500.0)
    x1 = x_host[0:1]

    def med(fn, n):
        fn()
        fn()
        ts = []
        for _ in range(n):
            torch.cuda.synchronize()
            t0 = time.perf_counter()
            fn()
            torch.cuda.synchronize()
            ts.append(1e3 * (time.perf_counter() - t0))
        ts.sort()
        return ts[len(ts) // 2]
    l0 = net.engine().be.launches
    wb.contrastive_ebp(x1, 0, 1)
    res['contrastive_ebp_launches'] = net.engine().be.launches - l0
    res['contrastive_ebp'] = med(lambda: wb.contrastive_ebp(x1, 0, 1), 15)
    res['truncated_contrastive_ebp'] = med(lambda: wb.truncated_contrastive_ebp(x1, 0, 1, percentile=20), 15)
    P = torch.zeros(1, 2)
    P[0, 0] = 1
    res['ebp'] = med(lambda: wb.ebp(x1, P), 15)
    wbs = whitebox.Whitebox(net, ebp_subtree_mode='norelu')
    unit = enc[0:2] / torch.norm(enc[0:2], dim=1, keepdim=True)
    net.set_triplet_classifier(unit[0:1], unit[1:2])
    res['weighted_subtree_ebp_topk32'] = med(lambda: wbs.weighted_subtree_ebp(x1, 0, 1, topk=32, verbose=False, do_max_subtree=False,
                                                                              do_mated_similarity_gating=False, subtree_mode='all'), 3)
    res['note'] = 'ms per call, public batch-1 API, host probe in / numpy map out'
    return res


def run_reference(args, rank):
    """--impl reference: the reference's CPU implementation of the path on this box's host cores.  The reference is
    pure Python on torch and cannot travel to the GPU box, so this is the oracle port (oracle/stresnet_oracle.py, pinned to
    the reference's outputs by tests/golden) with every host thread."""
    if rank != 0:
        return
    cores = os.cpu_count()
    per_step = 2
    vals = []
    from oracle import stresnet_oracle as O
    from xfr_b200 import synth
    torch.set_num_threads(cores)
    sd = synth.stresnet_state_dict(0)
    n = per_step * (args.steps + args.warmup)
    x = synth.synthetic_probes(n, seed=1)
    g = torch.Generator().manual_seed(5)
    W2 = torch.randn(n, 2, 512, generator=g) * 0.02
    k = 0
    t_steps = []
    for s in range(args.warmup + args.steps):
        t0 = time.time()
        for _ in range(per_step):
            O.contrastive_ebp(sd, x[k:k + 1], W2[k:k + 1])
            k += 1
        if s >= args.warmup:
            t_steps.append(time.time() - t0)
    tot = sum(t_steps)
    v = per_step * args.steps / tot
    sample = '%d triplets per step, batch 1, torch CPU fp32, %d threads' % (per_step, cores)
    emit(json.dumps({
        'impl': 'reference', 'metric': METRIC, 'value': v, 'unit': 'maps/s', 'n_gpus': args.gpus, 'steps': args.steps,
        'warmup': args.warmup, 'ms_per_step': 1e3 * tot / args.steps, 'higher_is_better': True, 'scaling': 'weak',
        'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
        'config': {'workload': 'contrastive triplet EBP, ResNet-101, synthetic 224x224 (bounded sample of configs[1])',
                   'mode': 'affineonly_with_prior', 'sample': sample},
        'cpu_baseline': {'value': v, 'unit': 'maps/s', 'cores': cores, 'kind': 'port', 'sample': sample},
        'e2e': {'value': v, 'unit': 'maps/s', 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0},
    }))


# ---------------------------------------------------------------------------------------------------------------------------
# The other BASELINE.json workloads (configs[2], [3], [4]); `--workload contrastive` (configs[1]) is the headline above.
WORKLOADS = {
    'layer_sweep': ('truncated contrastive EBP layer-sweep: layerwise_contrastive_ebp(mode=percentile, 20 %) at every affine firing, '
                    'ResNet-101 (BASELINE configs[2])', 'layer-sweep saliency maps/sec (ResNet-101, 224x224)', 'maps/s'),
    'weighted_subtree': ('weighted subtree triplet EBP, ResNet-101, eval-flow settings (ctor norelu, subtree_mode all, topk 32, no gating; '
                         'BASELINE configs[3]: the 541-job inpainting-game set, synthetic pixels)', 'weighted-subtree saliency jobs/sec (ResNet-101, 224x224)', 'jobs/s'),
    'lightcnn': ('Light-CNN-29v2 MFM-layer EBP, 128x128 grayscale, batch 512 per GPU (BASELINE configs[4])',
                 'EBP saliency maps/sec (Light-CNN-29v2, 128x128)', 'maps/s'),
}
AFFINE = ('Conv', 'Linear', 'AvgPool', 'BatchNorm')
# algorithmic HBM bytes (fp32), SURVEY.md section 8d element counts: ResNet-101 S_i 13.65 M / S_o 14.20 M, Light-CNN S_i 2.11 M / S_o 6.32 M
R101_ROW_BYTES = 4.0 * (14.20e6 + 13.65e6)        # one gradient row of an EBP backward sweep: gradient read + write per layer
R101_SAVED_BYTES = 4.0 * 2 * 14.20e6              # saved o / xr, read once per sweep of rows over the same probe
R101_FWD_BYTES = 4.0 * (13.65e6 + 2 * 14.20e6)
LC_MAP_BYTES = 4.0 * ((2.11e6 + 6.32e6) + (6.32e6 + (6.32e6 + 2.11e6)))     # forward (in + conv out) + backward (saved c + gradient read / write)


def _synthetic_jobs(n, seed):
    """n (im_mates, im_nonmates, probe_im) jobs of 224x224x3 uint8 images: 2 mates, 2 non-mates, 1 probe each - the shape of an
    inpainting-game job (generate_whitebox_saliency.py:222-416); the game's images are git-LFS pointers, hence synthetic pixels"""
    from xfr_b200 import synth
    x = synth.smooth_probes(5 * n, seed=seed) + torch.tensor(synth.MEAN_RGB).view(1, 3, 1, 1)
    im = [np.ascontiguousarray(t.permute(1, 2, 0).numpy().astype(np.uint8)) for t in x]
    return [(im[5 * i:5 * i + 2], im[5 * i + 2:5 * i + 4], im[5 * i + 4]) for i in range(n)]


def cpu_port_extra(workload, budget_s, threads):
    """CPU baseline of the extra workloads on a bounded sample: the reference algorithm restated (oracle/, torch CPU fp32)."""
    from xfr_b200 import synth
    torch.set_num_threads(threads)
    if workload == 'lightcnn':
        from oracle import lightcnn_oracle as LO
        sd = synth.lightcnn_state_dict(0, 2)
        x = synth.lightcnn_probes(64, seed=3, smooth=False)
        W2 = torch.randn(64, 2, 256, generator=torch.Generator().manual_seed(4))
        P = torch.zeros(1, 2)
        P[0, 0] = 1
        LO.ebp(sd, x[:1], P, W2[:1], mode='affineonly')
        t0, n = time.time(), 0
        while time.time() - t0 < budget_s and n < 63:
            n += 1
            LO.ebp(sd, x[n:n + 1], P, W2[n:n + 1], mode='affineonly')
        return n / (time.time() - t0), '%d probes, batch 1, ebp in mode affineonly' % n
    from oracle import stresnet_oracle as O
    sd = synth.stresnet_state_dict(0)
    x = synth.synthetic_probes(1, seed=1)
    W2 = torch.randn(1, 2, 512, generator=torch.Generator().manual_seed(5)) * 0.02
    P = torch.zeros(1, 2)
    P[0, 0] = 1
    mode = 'affineonly_with_prior' if workload == 'layer_sweep' else 'all'
    O.ebp_mwp(sd, x, P, W2, mode=mode, stop_at_stem=True)
    t0, n = time.time(), 0
    while time.time() - t0 < budget_s:
        n += 1
        O.ebp_mwp(sd, x, P, W2, mode=mode, stop_at_stem=True)          # one forward + one hooked backward = one ebp() of the reference
    t_ebp = (time.time() - t0) / n
    if workload == 'layer_sweep':
        # the reference runs three ebp() per (triplet, layer) (whitebox.py:584-644): mate, non-mate, the prior-restarted third pass
        return 1.0 / (3 * t_ebp), '%d ebp() passes timed, 3 per layer map as the reference runs them (whitebox.py:584-644)' % n
    # weighted_subtree_ebp: 3 true-gradient passes + 1 + 2 per firing (layerwise_ebp = 2 ebp()) = 2*377 + 4 passes per job (whitebox.py:647-737)
    passes = 2 * 377 + 4
    return 1.0 / (passes * t_ebp), ('%d ebp() passes timed and extrapolated to the %d passes one job makes (whitebox.py:647-737; the '
                                    'unmodified reference measured 499 s per job on 8 cores of the build container)' % (n, passes))


def run_extra(args, rank, local, world):
    """--workload layer_sweep | weighted_subtree | lightcnn: the same JSON schema as the headline workload."""
    import torch.distributed as dist
    from xfr_b200 import inpaintgame as IG
    from xfr_b200 import synth, whitebox
    label, metric, unit = WORKLOADS[args.workload]
    if args.impl == 'reference':
        if rank != 0:
            return
        cores = os.cpu_count()
        for k in range(args.warmup):
            cpu_port_extra(args.workload, 2.0, cores)
        vals = [cpu_port_extra(args.workload, 6.0, cores) for _ in range(max(1, args.steps))]
        v = sum(x[0] for x in vals) / len(vals)
        emit(json.dumps({'impl': 'reference', 'metric': metric, 'value': v, 'unit': unit, 'n_gpus': args.gpus, 'steps': args.steps, 'warmup': args.warmup,
                          'ms_per_step': 6000.0, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None, 'dtype': 'f32', 'data': 'synthetic',
                          'config': {'workload': label + ' (bounded sample)', 'sample': vals[-1][1]},
                          'cpu_baseline': {'value': v, 'unit': unit, 'cores': cores, 'kind': 'port', 'sample': vals[-1][1]},
                          'e2e': {'value': v, 'unit': unit, 'h2d_bytes_per_step': 0, 'd2h_bytes_per_step': 0}}))
        return
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)
    # Light-CNN runs on the split-TF32 kernels; the ResNet-101 sweeps take the plan as given (default bf16x2: forward and W+ dgrads on
    # pair tensors / kind::f16, true-gradient passes and the head on split-TF32)
    gemm = 'tf32x3' if (args.workload == 'lightcnn' and args.gemm == 'bf16x2') else args.gemm
    ev = lambda: torch.cuda.Event(enable_timing=True)
    cfg = {'workload': label, 'gemm': gemm, 'weights': 'seeded synthetic'}
    if args.workload == 'lightcnn':
        B = args.batch if args.batch != 256 else 512
        sd = {k: v.to(dev) for k, v in synth.lightcnn_state_dict(0, 2).items()}
        net = whitebox.WhiteboxLightCNN(sd, impl=gemm)
        wb = whitebox.Whitebox(net, ebp_subtree_mode='affineonly')               # demo/test_whitebox.py:232-254
        whitebox._CHUNK = 128
        x_host = synth.lightcnn_probes(B, seed=3 + rank, smooth=False).pin_memory()
        W2 = torch.randn(B, 2, 256, generator=torch.Generator().manual_seed(4)).to(dev)
        net.set_triplet_classifiers(W2[:, 0], W2[:, 1])
        P = torch.zeros(1, 2)
        P[0, 0] = 1
        x_dev = x_host.to(dev)
        units = B
        step_e2e = lambda: wb.ebp_batch(x_host, P)
        step_res = lambda: wb.ebp_batch(x_dev, P)
        h2d, d2h = x_host.numel() * 4, B * 128 * 128 * 4
        alg_bytes = LC_MAP_BYTES * B
        cfg.update(batch=B, mode='affineonly', chunk=128, l2='%.0f MB of probes + saved tensors per sweep, larger than L2' % (x_host.numel() * 4 / 1e6))
    else:
        sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
        T = args.batch if args.batch != 256 else 2                            # triplets / jobs per GPU per step
        if args.workload == 'layer_sweep':
            net = whitebox.WhiteboxSTResnet(sd, impl=gemm)
            wb = whitebox.Whitebox(net)
            x_host = synth.synthetic_probes(T, seed=100 + rank).pin_memory()
            with torch.no_grad():
                enc = net.encode(synth.synthetic_probes(2 * T, seed=1000 + rank).to(dev))
            rows = [(enc[i:i + 1] / 2500.0, enc[T + i:T + i + 1] / 2500.0) for i in range(T)]
            net.set_triplet_classifier(*rows[0])
            wb.layerwise_contrastive_ebp_sweep(x_host[0:1], 0, 1, [3], mode='percentile', percentile=20)
            names = wb.P_layername
            ks = [k for k, n in enumerate(names[:-1]) if any(a in n for a in AFFINE)]     # non-affine firings give all-zero maps in this mode
            x_dev = x_host.to(dev)
            units = T * len(ks)

            def sweep(x):
                out = []
                for i in range(T):
                    net.set_triplet_classifier(*rows[i])
                    out.append(wb.layerwise_contrastive_ebp_sweep(x[i:i + 1], 0, 1, ks, mode='percentile', percentile=20, rows_per_sweep=args.rows))
                return out
            step_e2e = lambda: sweep(x_host)
            step_res = lambda: sweep(x_dev)
            h2d, d2h = x_host.numel() * 4, units * 112 * 112 * 4
            nsweeps = -(-len(ks) // args.rows)
            alg_bytes = T * (R101_FWD_BYTES + (len(ks) + 2) * R101_ROW_BYTES + (nsweeps + 1) * R101_SAVED_BYTES)
            cfg.update(triplets_per_gpu_per_step=T, layers_per_triplet=len(ks), firings=len(names), rows_per_sweep=args.rows,
                       note='BASELINE configs[2] runs 1,024 triplets on 8 GPUs: 1024 x %d layer maps' % len(ks),
                       l2='every sweep streams > 1 GB of saved tensors and gradients (larger than L2)')
        else:
            net = whitebox.WhiteboxSTResnet(sd, impl=gemm)
            wb = whitebox.Whitebox(net, ebp_subtree_mode='norelu')                 # create_wbnet.py:26-27
            # the GLOBAL job list, identical on every rank: the sharded driver gives rank r its contiguous slice of T jobs.  (Until the
            # end of round 2 every rank built its OWN T jobs and handed them to the sharded driver, which then ran T jobs in total
            # (one per rank on the first T ranks) while world * T were counted: the 8-GPU figures of this workload in profiles/r2m_*
            # and r2w_bench8_weighted_subtree.json are inflated by the factor T = 2.)
            jobs = _synthetic_jobs(world * T, seed=200)
            units = T
            step_e2e = lambda: IG.run_weighted_subtree_triplet_ebp_sharded(wb, jobs, subtree_mode_weighted='all', ebp_version=None, device=dev, topk=32)
            step_res = step_e2e               # the job API takes host images (numpy): there is no device-resident variant of a job
            h2d, d2h = T * 5 * 224 * 224 * 3 * 4, T * 112 * 112 * 4
            nrows = 3 + 1 + 377
            alg_bytes = T * (5 * R101_FWD_BYTES + nrows * R101_ROW_BYTES + (2 + -(-377 // 48)) * R101_SAVED_BYTES)
            cfg.update(jobs_per_gpu_per_step=T, topk=32, subtree_mode='all', ctor_mode='norelu', gating=False,
                       note='BASELINE configs[3]: the 541 resnetv4 probe jobs of filtered_masks_threshold-resnetv4_pytorch.csv; jobs are '
                            'independent, so the set takes 541 / value seconds',
                       l2='every sweep streams > 1 GB of saved tensors and gradients (larger than L2)')
    eng = net.engine(wb._ebp_with_bias)

    def timed(fn, steps):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = ev(), ev()
        s.record()
        for _ in range(steps):
            fn()
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())
    for _ in range(args.warmup):
        step_res()
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.be.launches
    ms = timed(step_res, args.steps)
    launches = eng.be.launches - l0
    ms_e2e = timed(step_e2e, args.steps) if step_e2e is not step_res else ms
    clocks = sampler.summary()
    total = world * units * args.steps
    peak, peak_src = measured_peaks()
    ach = alg_bytes * args.steps / (ms / 1e3) / 1e9
    out = {'metric': metric, 'value': total / (ms / 1e3), 'unit': unit, 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
           'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
           'dtype': ('bf16 terms on tcgen05 kind::f16 (activations / gradients 2 terms, relu(W) 1, signed W 2) for the forward and the W+ dgrads, '
                     'split-TF32 for the true-gradient passes and the head; fp32 accumulate and fp32 hook algebra') if gemm == 'bf16x2'
           else 'f32 (split-TF32 tcgen05: 3 passes on signed weights, 2 on W+; fp32 accumulate)', 'data': 'synthetic', 'config': cfg,
           'e2e': {'value': total / (ms_e2e / 1e3), 'unit': unit, 'h2d_bytes_per_step': int(h2d), 'd2h_bytes_per_step': int(d2h),
                   'ms_per_step': ms_e2e / args.steps},
           'gpu_launches': int(launches), 'clocks': clocks,
           'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak, 'traffic': None,
                        'kernel': 'whole step (firing-by-firing sweep: xfrb_hook + xfrb_dgrad_plain per firing)' if args.workload != 'lightcnn'
                        else 'whole step (Light-CNN sweep: conv_bias GEMMs, mfm / pool kernels, xfrb_hook per firing)',
                        'alg_bytes_per_step': alg_bytes, 'peak_source': peak_src}}
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        v, sample = cpu_port_extra(args.workload, 12.0, cores)
        out['cpu_baseline'] = {'value': v, 'unit': unit, 'cores': cores, 'kind': 'port', 'sample': sample}
    if rank == 0:
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_RESULT_FD = None


def emit(line):
    """The ONE line this program puts on its standard output."""
    if _RESULT_FD is None:
        print(line, flush=True)
    else:
        os.write(_RESULT_FD, (line + '\n').encode())


def main():
    # stdout carries ONE JSON line.  Libraries write there too (the boxes set NCCL_DEBUG=VERSION: NCCL prints its version banner on
    # stdout at communicator set-up), so file descriptor 1 is pointed at stderr for the run and the result goes to the saved one.
    global _RESULT_FD
    sys.stdout.flush()
    _RESULT_FD = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument('--gpus', type=int, default=1)
    ap.add_argument('--steps', type=int, default=5)
    ap.add_argument('--warmup', type=int, default=3)
    ap.add_argument('--batch', type=int, default=256, help='triplets per GPU per step')
    ap.add_argument('--chunk', type=int, default=256, help='probes per engine sweep (256: 64 GB workspace, +4 %% over 128)')
    ap.add_argument('--gemm', default='bf16x2', choices=['tf32x3', 'tf32x3full', 'tf32', 'fp32', 'tf32x2f', 'tf32x3b1', 'bf16x2'],
                    help="bf16x2 (default, = xfr_b200.whitebox.DEFAULT_IMPL); tf32x3 / tf32x3full: split-TF32 plans; fp32: CUDA cores")
    ap.add_argument('--mode', default='affineonly_with_prior')
    ap.add_argument('--impl', default='b200', choices=['b200', 'reference'])
    ap.add_argument('--no-cpu-baseline', action='store_true')
    ap.add_argument('--workload', default='contrastive', choices=['contrastive'] + sorted(WORKLOADS),
                    help='contrastive = BASELINE configs[1] (the headline); the others are configs[2], [3], [4]')
    ap.add_argument('--rows', type=int, default=48, help='gradient rows per firing-by-firing sweep (layer_sweep)')
    ap.add_argument('--no-extras', action='store_true', help='skip the batch-1 latency and library-comparator legs')
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3) if args.impl == 'b200' else args.warmup

    world = int(os.environ.get('WORLD_SIZE', '1'))
    rank = int(os.environ.get('RANK', '0'))
    local = int(os.environ.get('LOCAL_RANK', '0'))
    if args.workload != 'contrastive':
        return run_extra(args, rank, local, world)
    if args.impl == 'reference':
        return run_reference(args, rank)

    import torch.distributed as dist
    from xfr_b200 import synth, whitebox
    from xfr_b200.engine import StResnetEngine  # noqa: F401
    dev = torch.device('cuda', local)
    torch.cuda.set_device(dev)
    if world > 1:
        dist.init_process_group('nccl', device_id=dev)

    # ---------------- setup (untimed): weights, probes, classifier rows from encodings of 2*B further images
    B = args.batch
    sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0).items()}
    net = whitebox.WhiteboxSTResnet(sd, impl=args.gemm)
    wb = whitebox.Whitebox(net, ebp_subtree_mode=args.mode)
    whitebox._CHUNK = args.chunk
    x_host = synth.synthetic_probes(B, seed=100 + rank).pin_memory()            # [B,3,224,224] fp32, 154 MB at B=256
    with torch.no_grad():
        enc = torch.cat([net.encode(synth.synthetic_probes(64, seed=1000 + 16 * rank + i).to(dev)) for i in range(2 * B // 64)]) \
            if B >= 64 else net.encode(synth.synthetic_probes(2 * B, seed=1000 + rank).to(dev))
    net.set_triplet_classifiers(enc[:B] / 2500.0, enc[B:2 * B] / 2500.0)       # generate_whitebox_saliency.py:103
    eng = net.engine()
    W2 = net.triplet_rows(B)
    x_dev = x_host.to(dev).permute(0, 2, 3, 1).contiguous()                    # resident NHWC copy for the kernel-side number
    maps_dev = torch.empty(B, 112, 112, device=dev)
    maps_host = torch.empty(B, 112, 112).pin_memory()
    from xfr_b200.shard import gather_maps

    ev = lambda: torch.cuda.Event(enable_timing=True)
    bwd_events = []

    def step_resident(record, gather=True):
        for i in range(0, B, args.chunk):
            xs, ws = x_dev[i:i + args.chunk], W2[i:i + args.chunk].contiguous()
            n = xs.shape[0]
            eng.forward(xs)
            if record:
                e0, e1 = ev(), ev()
                e0.record()
            P2, _, sums = eng.ebp_backward(eng.priors_contrastive(n, 2, 0, 1), ws, args.mode)
            if record:
                e1.record()
                bwd_events.append((e0, e1))
            mwp = eng.buf('cmwp', n, 112, 112)
            eng.be.contrast(P2, sums, n, mwp)
            eng.be.saliency_post(mwp, maps_dev[i:i + n])
        if world > 1 and gather:
            gather_maps(maps_dev, world * B, dst=0)          # the only collective on the data path (NCCL gather)

    def step_e2e(nsteps=1):
        # the public streaming call: every step copies its probes from pinned host memory (the copy of step i+1 overlaps the sweep
        # of step i on a copy stream) and its maps back to pinned host memory
        wb.contrastive_ebp_stream([x_host] * nsteps, 0, 1, outs=[maps_host] * nsteps)
        if world > 1:
            dist.barrier()

    def timed(fn, steps, **kw):
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        s, e = ev(), ev()
        s.record()
        for _ in range(steps):
            fn(**kw)
        e.record()
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        ms = torch.tensor([s.elapsed_time(e)], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        return float(ms.item())

    for _ in range(args.warmup):
        step_resident(False)
    sampler = ClockSampler(local)
    sampler.start()
    l0 = eng.be.launches
    ms = timed(step_resident, args.steps, record=True)
    launches = eng.be.launches - l0
    torch.cuda.synchronize()
    bwd_ms = sum(a.elapsed_time(b) for a, b in bwd_events)
    for _ in range(2):
        step_e2e()
    ms_e2e = timed(step_e2e, 1, nsteps=args.steps)
    clocks = sampler.summary()
    families = kernel_families(eng, step_resident, args.gemm) if rank == 0 else []      # after the timed regions: events around launches

    if world > 1:
        lt = torch.tensor([float(launches), bwd_ms], device=dev)
        dist.all_reduce(lt, op=dist.ReduceOp.MAX)
        bwd_ms = float(lt[1].item())
    total_maps = world * B * args.steps
    value = total_maps / (ms / 1e3)
    e2e = total_maps / (ms_e2e / 1e3)
    peak, peak_src = measured_peaks()
    # EBP-backward stage of ONE GPU: algorithmic bytes of the maps it swept / device time of its backward sweeps
    ach = EBP_BWD_BYTES_PER_MAP * B * args.steps / (bwd_ms / 1e3) / 1e9
    out = {
        'metric': METRIC, 'value': value, 'unit': 'maps/s', 'n_gpus': world, 'steps': args.steps, 'warmup': args.warmup,
        'ms_per_step': ms / args.steps, 'higher_is_better': True, 'scaling': 'weak', 'vs_baseline': None,
        'dtype': {'tf32x3': 'f32 (split-TF32 tcgen05: 3 passes on signed weights, 2 on W+; fp32 accumulate)',
                  'tf32x3full': 'f32 (3xTF32 tcgen05 in every GEMM, fp32 accumulate)', 'tf32': 'tf32', 'fp32': 'f32',
                  'tf32x2f': 'f32 (opt-in hybrid: two-pass split-TF32 forward with TF32-rounded weights, default W+ dgrads; fp32 accumulate)',
                  'bf16x2': 'bf16 terms on tcgen05 kind::f16 (activations / gradients 2 terms = 16 bits, relu(W) 1, signed W 2), fp32 accumulate and fp32 epilogues',
                  'tf32x3b1': 'f32 forward (split-TF32 tcgen05) / tf32 single-pass W+ dgrads (opt-in hybrid, not the parity-grade default)'}[args.gemm], 'data': 'synthetic',
        'config': {'workload': 'contrastive triplet EBP, ResNet-101, batch %d synthetic 224x224 per GPU (BASELINE configs[1])' % B,
                   'mode': args.mode, 'ebp_version': 6, 'chunk': args.chunk, 'gemm': args.gemm,
                   'l2': 'inputs larger than L2 (%.0f MB probes, %.1f GB workspace per sweep)' % (x_dev.numel() * 4 / 1e6, eng.workspace_bytes() / 1e9),
                   'weights': 'seeded synthetic (real STR weights are git-LFS pointers)'},
        'e2e': {'value': e2e, 'unit': 'maps/s', 'h2d_bytes_per_step': int(x_host.numel() * 4), 'd2h_bytes_per_step': int(maps_host.numel() * 4),
                'ms_per_step': ms_e2e / args.steps},
        'gpu_launches': int(launches),
        'clocks': clocks,
        'roofline': {'bound': 'hbm', 'achieved': ach, 'peak': peak, 'unit': 'GB/s', 'frac': ach / peak,
                     'traffic': NCU_TRAFFIC['dgrad_join'],
                     'traffic_note': 'DRAM read+write of one layer3 JOIN launch at a 128-probe sweep (ncu --set full, profiles/r2_ncu_full_bf16x2.csv); '
                                     'its algorithmic bytes are 976e6',
                     'mma_issue_peaks_measured_tflops': MMA_PEAK_TFLOPS,
                     'kernel': 'EBP backward sweep (conv_tc_kernel dgrad + fused hook epilogues, join/stem kernels)',
                     'peak_source': peak_src, 'bwd_ms_per_step': bwd_ms / args.steps,
                     'tensor_tflops_whole_step': FLOP_PER_MAP * B * args.steps / (ms / 1e3) / 1e12 / 1.0,
                     'kernels': families},
    }
    if rank == 0 and world == 1 and not args.no_extras:
        out['latency_ms_batch1'] = batch1_latency(wb, net, dev, x_host, enc)
        torch.cuda.empty_cache()
        out['gpu_library_baseline'] = gpu_library_baseline(dev, 64, 32)
        out['gpu_library_baseline']['speedup_of_this_repo'] = value / out['gpu_library_baseline']['value']
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        cores = os.cpu_count()
        n = 40                                   # ~10 s of host work at the ~4 maps/s measured on the GPU box's 16 cores
        out['cpu_baseline'] = {'value': cpu_port(n, cores), 'unit': 'maps/s', 'cores': cores, 'kind': 'port',
                               'sample': '%d triplets of the same workload, batch 1, torch CPU fp32 restatement (oracle/)' % n,
                               'note': 'the port shares one forward between the mate and the non-mate sweep (2 forwards + 2 backwards per map); the '
                                       'unmodified reference (hook-based, 6 forwards + 2 backwards) measured 0.73 maps/s on 8 cores of the build '
                                       'container (BASELINE.md section 2) and cannot travel to the GPU box',
                               'reference_measured_maps_per_s_8_cores': 0.73}
    if rank == 0:
        emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


if __name__ == '__main__':
    main()
