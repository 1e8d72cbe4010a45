"""Generate tests/golden/preprocess_seed0.npz by running the UNMODIFIED reference (TEST INFRASTRUCTURE).

Runs only in the build container, where /root/reference exists:   python oracle/gen_golden_preprocess.py

`net.preprocess(PIL image)` of the reference's three plugins - WhiteboxSTResnet (whitebox.py:108-110, resnet.py:25-37),
Whitebox_resnet50_128 (whitebox.py:235-258), WhiteboxLightCNN (whitebox.py:137-139, lightcnn.py:19-31; torchvision
transforms) - and Whitebox.convert_from_numpy (whitebox.py:787-806, 224x224 inputs) on seeded synthetic images
(tests/preprocess_fixture.py).  Stored: SHA-256 of the float32 tensor bytes plus shape / sum / min / max.
"""
import hashlib
import os
import sys
import warnings

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
sys.path[:0] = [os.path.join(HERE, 'shim'), '/root/reference/python', '/root/reference/models/resnet50_128_pytorch', ROOT]
warnings.filterwarnings('ignore')

import numpy as np  # noqa: E402
import torch  # noqa: E402

from xfr.models import whitebox as RW  # noqa: E402  (the reference)
from xfr.models.resnet import ResNet, Bottleneck  # noqa: E402  (the reference)
from xfr.models.lightcnn import LightCNN_29Layers_v2  # noqa: E402  (the reference)

sys.path.insert(0, os.path.join(ROOT, 'tests'))
from preprocess_fixture import test_images  # noqa: E402


def record(G, key, t):
    a = np.ascontiguousarray(t.detach().numpy())
    assert a.dtype == np.float32
    G[key + '_sha256'] = np.array(hashlib.sha256(a.tobytes()).hexdigest())
    G[key + '_shape'] = np.array(a.shape)
    G[key + '_stats'] = np.array([a.astype(np.float64).sum(), a.min(), a.max()])


def main():
    torch.manual_seed(0)
    G = {}
    str_net = RW.WhiteboxSTResnet(ResNet(Bottleneck, [1, 1, 1, 1], mode='encode', num_classes=2))
    lc_net = RW.WhiteboxLightCNN(LightCNN_29Layers_v2(num_classes=10))
    r50 = RW.Whitebox_resnet50_128.__new__(RW.Whitebox_resnet50_128)       # preprocess() reads no state; the ctor wants weights
    wb = RW.Whitebox(str_net)
    for name, im in test_images().items():
        record(G, 'stresnet_' + name, str_net.preprocess(im))
        record(G, 'resnet50_128_' + name, RW.Whitebox_resnet50_128.preprocess(r50, im))
        record(G, 'lightcnn_' + name, lc_net.preprocess(im))
        if im.size == (224, 224):
            record(G, 'from_numpy_u8_' + name, wb.convert_from_numpy(np.array(im)))
            record(G, 'from_numpy_f32_' + name, wb.convert_from_numpy(np.array(im).astype(np.float32) / 255))
            record(G, 'from_numpy_f64_255_' + name, wb.convert_from_numpy(np.array(im).astype(np.float64)))
    out = os.path.join(ROOT, 'tests', 'golden', 'preprocess_seed0.npz')
    np.savez_compressed(out, **G)
    print('wrote %s (%d arrays)' % (out, len(G)))
    for k in sorted(G):
        if k.endswith('_shape'):
            print(k, G[k], G[k.replace('_shape', '_stats')])


if __name__ == '__main__':
    main()
