"""Stock-PyTorch (cuDNN / cuBLAS) comparator on the same B200: python tools/torch_eager_baseline.py [batch] [triplets]"""
import json
import os
import sys

sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), '..'))
import torch  # noqa: E402

import bench  # noqa: E402

if __name__ == '__main__':
    batch = int(sys.argv[1]) if len(sys.argv) > 1 else 16
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 64
    dev = torch.device('cuda:0')
    for tf32 in (False, True):
        print(json.dumps(bench.gpu_library_baseline(dev, n, batch, tf32)))
