#!/bin/bash
# GPU call 34: quad-per-thread stem backward (A) against the per-pixel kernel (B = XFRB_STEM_QUAD=0); parity suites (STR pad 1, VGGFace2 pad 0)
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_parity.py tests/test_bf16x2.py tests/test_resnet50_128.py tests/test_gpu_kernels.py tests/test_stream.py -m gpu -q -x > gpurun_out/r2am_tests.log 2>&1; echo "rc $?" >> gpurun_out/r2am_tests.log
cat > /tmp/stem_ab.py <<'PY'
import sys, torch, numpy as np
sys.path.insert(0, '.')
from xfr_b200 import synth, whitebox
dev = torch.device('cuda:0')
sd = {k: v.to(dev) for k, v in synth.stresnet_state_dict(0, (1, 1, 1, 1), 2).items()}
net = whitebox.WhiteboxSTResnet(sd, layers=(1, 1, 1, 1))
wb = whitebox.Whitebox(net)
g = torch.Generator().manual_seed(3)
net.set_triplet_classifiers(torch.randn(8, 512, generator=g).to(dev) / 50, torch.randn(8, 512, generator=g).to(dev) / 50)
m = wb.contrastive_ebp_batch(synth.synthetic_probes(8, seed=5), 0, 1)
np.save(sys.argv[1], m)
PY
XFRB_STEM_QUAD=1 python /tmp/stem_ab.py gpurun_out/stem_quad_1.npy; XFRB_STEM_QUAD=0 python /tmp/stem_ab.py gpurun_out/stem_quad_0.npy
python -c "import numpy as np; a=np.load('gpurun_out/stem_quad_1.npy'); b=np.load('gpurun_out/stem_quad_0.npy'); print('quad vs per-pixel: equal', np.array_equal(a,b), 'max-abs', float(np.abs(a-b).max()), 'max', float(a.max()))"
for v in A B A B; do
  if [ $v = B ]; then export XFRB_STEM_QUAD=0; else unset XFRB_STEM_QUAD; fi
  timeout 300 python bench.py --no-cpu-baseline --no-extras > gpurun_out/r2am_bench_$v.json 2> gpurun_out/r2am_bench_$v.err
  python - <<PY
import json
d = json.load(open('gpurun_out/r2am_bench_$v.json'))
print('variant $v', round(d['value']), 'e2e', round(d['e2e']['value']), 'bwd', round(d['roofline']['bwd_ms_per_step'], 2), d['clocks']['sm_mhz'])
PY
done
tail -n 3 gpurun_out/r2am_tests.log | cut -c1-300
