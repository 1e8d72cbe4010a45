#!/bin/bash
# GPU call 26: K >= 1,024 MID dgrads on 128-wide tiles (A) against the default 256 (B); pairs off for them as well (C)
mkdir -p gpurun_out
for v in A B C; do
  unset XFRB_MID_BN XFRB_PAIR_KINDS
  if [ $v = A ]; then export XFRB_MID_BN=128; fi
  if [ $v = C ]; then export XFRB_PAIR_KINDS=1; fi
  XFRB_BENCH_LAUNCHES=gpurun_out/r2ac_launches_$v.jsonl timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2ac_bench_$v.json 2> gpurun_out/r2ac_bench_$v.err
  echo "VARIANT $v"; python tools/launch_roofline.py gpurun_out/r2ac_launches_$v.jsonl 2232 2 2>/dev/null | grep "dgrad_mid" | head -4; cut -c1-120 gpurun_out/r2ac_bench_$v.json
done
