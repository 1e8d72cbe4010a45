"""Seeded synthetic weights and inputs (no datasets / checkpoints can be fetched).

The real STR-Janus ResNet-101 weights are git-LFS pointers in the reference
(models/resnet101v4_28NOV17_train.pth), so parity and throughput runs use a
deterministic state_dict with the reference's own initialisation statistics
(reference resnet.py:191-198: conv ~ N(0, sqrt(2/(k*k*Cout))), bias same) plus
randomised BatchNorm affine/statistics so that the gamma+ / beta / mu paths of
excitation backprop are exercised.

The same state_dict loads into the reference's `xfr.models.resnet.ResNet`
(oracle/gen_golden.py) and into this package's engine, which is what makes the
committed golden maps reproducible on the GPU box from the seed alone.
"""
import math

import numpy as np
import torch

STRESNET101_LAYERS = (3, 4, 23, 3)
MEAN_RGB = (122.782, 117.001, 104.298)  # reference resnet.py:23


def _bn(sd, prefix, c, g):
    sd[prefix + '.weight'] = torch.randn(c, generator=g) * 0.5 + 0.7
    sd[prefix + '.bias'] = torch.randn(c, generator=g) * 0.2
    sd[prefix + '.running_mean'] = torch.randn(c, generator=g) * 0.3
    sd[prefix + '.running_var'] = torch.rand(c, generator=g) + 0.5
    sd[prefix + '.num_batches_tracked'] = torch.zeros((), dtype=torch.long)


def _conv(sd, prefix, cout, cin, k, g, bias=True):
    std = math.sqrt(2.0 / (k * k * cout))
    sd[prefix + '.weight'] = torch.randn(cout, cin, k, k, generator=g) * std
    if bias:
        sd[prefix + '.bias'] = torch.randn(cout, generator=g) * std


def _linear(sd, prefix, cout, cin, g, bias=True):
    bound = 1.0 / math.sqrt(cin)
    sd[prefix + '.weight'] = (torch.rand(cout, cin, generator=g) * 2 - 1) * bound
    if bias:
        sd[prefix + '.bias'] = (torch.rand(cout, generator=g) * 2 - 1) * bound


def stresnet_state_dict(seed=0, layers=STRESNET101_LAYERS, num_classes=2):
    """state_dict with the key layout of the reference STR ResNet (resnet.py:168-221)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    _conv(sd, 'conv1', 64, 3, 7, g)
    _bn(sd, 'bn1', 64, g)
    inplanes = 64
    for li, (planes, nblocks) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for bi in range(nblocks):
            p = 'layer%d.%d' % (li, bi)
            _conv(sd, p + '.conv1', planes, inplanes, 1, g)
            _bn(sd, p + '.bn1', planes, g)
            _conv(sd, p + '.conv2', planes, planes, 3, g)
            _bn(sd, p + '.bn2', planes, g)
            _conv(sd, p + '.conv3', planes * 4, planes, 1, g)
            _bn(sd, p + '.bn3', planes * 4, g)
            inplanes = planes * 4
    _linear(sd, 'fc1', 512, inplanes, g)
    _linear(sd, 'fc2', num_classes, 512, g)
    return sd


def synthetic_probes(n, seed=1, kind='uint8'):
    """n synthetic probe tensors [n,3,224,224] fp32, already mean-subtracted
    (reference resnet.py:25-37 / whitebox.py:108-110)."""
    g = torch.Generator().manual_seed(seed)
    if kind == 'uint8':
        img = torch.randint(0, 256, (n, 224, 224, 3), generator=g, dtype=torch.int32).float()
        img = img - torch.tensor(MEAN_RGB, dtype=torch.float32)
        return img.permute(0, 3, 1, 2).contiguous()
    if kind == 'normal':
        return torch.randn(n, 3, 224, 224, generator=g) * 60.0
    raise ValueError(kind)


def smooth_probes(n, seed=1):
    """Low-frequency synthetic 'face-like' probes: a smooth random field so that
    neighbouring pixels correlate as in a photograph (used by parity tests so the
    saliency is not white noise)."""
    g = torch.Generator().manual_seed(seed)
    coarse = torch.rand(n, 3, 14, 14, generator=g) * 255.0
    img = torch.nn.functional.interpolate(coarse, size=(224, 224), mode='bicubic', align_corners=False)
    img = img.clamp(0, 255).round()
    mean = torch.tensor(MEAN_RGB, dtype=torch.float32).view(1, 3, 1, 1)
    return (img - mean).contiguous()


# ---------------------------------------------------------------- Light-CNN-29v2 (reference lightcnn.py:216-275)
LIGHTCNN_LAYERS = (1, 2, 3, 4)


def lightcnn_convs(layers=LIGHTCNN_LAYERS):
    """(prefix, cin, cout, k) of every mfm of network_29layers_v2 in forward order (the Conv2d has 2*cout outputs)."""
    out = [('conv1', 1, 48, 5)]
    for bi, (n, cin, cout) in enumerate(zip(layers, (48, 96, 192, 128), (96, 192, 128, 128)), start=1):
        for i in range(n):
            out.append(('block%d.%d.conv1' % (bi, i), cin, cin, 3))
            out.append(('block%d.%d.conv2' % (bi, i), cin, cin, 3))
        out.append(('group%d.conv_a' % bi, cin, cin, 1))
        out.append(('group%d.conv' % bi, cin, cout, 3))
    return out


def lightcnn_state_dict(seed=0, num_classes=2, layers=LIGHTCNN_LAYERS):
    """Seeded state_dict with the key layout of the reference network_29layers_v2 and torch's default
    Conv2d / Linear initialisation statistics, U(-1/sqrt(fan_in), 1/sqrt(fan_in)) for weights and biases
    (no Light-CNN weights are bundled with the reference)."""
    g = torch.Generator().manual_seed(seed)
    sd = {}
    for prefix, cin, cout, k in lightcnn_convs(layers):
        bound = 1.0 / math.sqrt(cin * k * k)
        sd[prefix + '.filter.weight'] = (torch.rand(2 * cout, cin, k, k, generator=g) * 2 - 1) * bound
        sd[prefix + '.filter.bias'] = (torch.rand(2 * cout, generator=g) * 2 - 1) * bound
    _linear(sd, 'fc', 256, 8 * 8 * 128, g)
    _linear(sd, 'fc2', num_classes, 256, g, bias=False)
    return sd


def lightcnn_probes(n, seed=1, smooth=True):
    """n synthetic Light-CNN inputs [n,1,128,128] in [0,1] (reference lightcnn.py:19-31: grayscale / 255)."""
    g = torch.Generator().manual_seed(seed)
    if smooth:
        coarse = torch.rand(n, 1, 16, 16, generator=g) * 255.0
        img = torch.nn.functional.interpolate(coarse, size=(128, 128), mode='bicubic', align_corners=False)
        return (img.clamp(0, 255).round() / 255.0).contiguous()
    return (torch.randint(0, 256, (n, 1, 128, 128), generator=g, dtype=torch.int32).float() / 255.0).contiguous()
