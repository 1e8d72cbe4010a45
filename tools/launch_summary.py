"""Turn an ncu launch list into the per-sweep CSV kept under profiles/ plus a per-kernel share table.

    ncu --metrics gpu__time_duration.sum --clock-control none -s 1000 -c 480 --csv --log-file gpurun_out/launches.csv \\
        python bench.py --steps 1 --warmup 3 --batch 128 --chunk 128 --no-cpu-baseline
    python tools/launch_summary.py gpurun_out/launches.csv profiles/rN_launches_<build>.csv

The capture window (-s / -c) only has to CONTAIN one complete sweep (225 launches at chunk 128): the first stretch from one
stem_conv_kernel launch to the next is extracted.  Times under ncu are cold-cache and serialised: compare shares, not sums."""
import collections
import csv
import re
import sys


def read(path):
    lines = open(path, errors='ignore').readlines()
    start = [i for i, l in enumerate(lines) if l.startswith('"ID"')][0]
    rows = []
    for r in csv.DictReader(lines[start:]):
        unit = r['Metric Unit']
        v = float(r['Metric Value'].replace(',', ''))
        us = v / 1e3 if unit in ('ns', 'nsecond') else v * 1e3 if unit in ('ms', 'msecond') else v
        rows.append((r['Kernel Name'], r['Grid Size'], r['Block Size'], us))
    return rows


def main():
    rows = read(sys.argv[1])
    stems = [i for i, r in enumerate(rows) if 'stem_conv_kernel' in r[0]]
    if not stems:
        sys.exit('no stem_conv_kernel launch inside the capture window: move -s')
    lo, hi = stems[0], (stems[1] if len(stems) > 1 else len(rows))
    sweep = rows[lo:hi]
    if len(stems) < 2:
        sys.stderr.write('warning: only one stem launch in the window, the sweep may be cut short\n')
    if len(sys.argv) > 2:
        with open(sys.argv[2], 'w') as f:
            f.write('launch,kernel,grid,block,gpu__time_duration.sum [us]\n')
            for i, r in enumerate(sweep):
                f.write('%d,"%s","%s","%s",%.1f\n' % (i, re.sub(r'\(CUtensorMap_st.*', '', r[0]), r[1], r[2], r[3]))
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in sweep:
        name = r[0].replace('void ', '')
        name = name[:name.index('>(') + 1] if '>(' in name else name.split('(')[0]
        agg[name][0] += 1
        agg[name][1] += r[3]
    tot = sum(v[1] for v in agg.values())
    print('%d launches, %.2f ms' % (len(sweep), tot / 1e3))
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print('%-64s %4d %10.1f us %5.1f %%' % (k[:64], v[0], v[1], 100 * v[1] / tot))


if __name__ == '__main__':
    main()
