"""Per-launch roofline of a bench run: XFRB_BENCH_LAUNCHES=file python bench.py ... ; python tools/launch_roofline.py file [tensor TFLOP/s issued peak] [passes]"""
import json
import sys

rows = [json.loads(l) for l in open(sys.argv[1])]
hbm = 6553.3e9
peak = float(sys.argv[2]) * 1e12 if len(sys.argv) > 2 else 1115e12       # measured kind::tf32 issue peak (tools/mma_peak.cu)
passes = {'dgrad_join': 2.0, 'dgrad_mid': 2.0, 'conv_dual': 2.5}
if len(sys.argv) > 3:
    passes = {k: float(sys.argv[3]) * v / 2.0 for k, v in passes.items()}
tot = {}
print('%-11s %-28s %8s %8s %8s %6s' % ('kind', 'A shape + cin,cout,R', 'us', 'hbm_us', 'mma_us', 'eff'))
for r in rows:
    t_h = r['bytes'] / hbm * 1e6
    t_m = passes[r['k']] * r['flops'] / peak * 1e6
    b = max(t_h, t_m)
    key = (r['k'], tuple(r['shape']))
    a = tot.setdefault(key, [0, 0.0, 0.0, t_h, t_m])
    a[0] += 1; a[1] += r['us']; a[2] += b
for (k, shp), (n, us, b, t_h, t_m) in sorted(tot.items(), key=lambda kv: -kv[1][1]):
    print('%-11s %-28s x%-3d %7.0f %8.0f %8.0f %6.2f   total %.2f ms (bound %.2f)' % (k, str(list(shp)), n, us / n, t_h, t_m, b / us, us / 1e3, b / 1e3))
print('sum %.1f ms, sum of bounds %.1f ms' % (sum(v[1] for v in tot.values()) / 1e3, sum(v[2] for v in tot.values()) / 1e3))
