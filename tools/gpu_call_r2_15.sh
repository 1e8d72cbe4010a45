#!/bin/bash
# GPU call 15: phase / kernel profile of the graph-replayed generic sweeps; streaming e2e; stream test; ncu of the elementwise kernels
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_stream.py -m gpu -q > gpurun_out/r2p_stream_test.log 2>&1; echo "rc $?" >> gpurun_out/r2p_stream_test.log
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2p_profile_layer_sweep.log 2>&1
timeout 300 python tools/generic_profile.py weighted_subtree > gpurun_out/r2p_profile_weighted_subtree.log 2>&1
timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2p_bench.json 2> gpurun_out/r2p_bench.err
NCU="ncu --set full --clock-control none --import-source on --kernel-name-base demangled"
B="python bench.py --no-cpu-baseline --no-extras --batch 128 --chunk 128 --steps 1 --warmup 3"
timeout 400 $NCU -k 'regex:stem_bwd_b_kernel|contrast_kernel|join_kernel|stem_conv_kernel' -s 21 -c 7 -o gpurun_out/r2p_ncu_elementwise -f $B > gpurun_out/r2p_ncu_elementwise.log 2>&1
tail -n 3 gpurun_out/r2p_stream_test.log
cat gpurun_out/r2p_profile_layer_sweep.log | tail -n 25
cat gpurun_out/r2p_profile_weighted_subtree.log | tail -n 25
python - <<'PY'
import json
d=json.load(open('gpurun_out/r2p_bench.json'))
print(d['value'], d['e2e'], d['roofline']['frac'], d['roofline']['bwd_ms_per_step'])
PY
tail -n 2 gpurun_out/r2p_bench.err
