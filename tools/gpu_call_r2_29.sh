#!/bin/bash
# GPU call 29: do the layer-sweep graphs reach a steady state? (capture counter)
mkdir -p gpurun_out
timeout 300 python tools/generic_profile.py layer_sweep > gpurun_out/r2af_profile_layer_sweep.log 2>&1
grep -v Warn gpurun_out/r2af_profile_layer_sweep.log | head -12 | cut -c1-170
