"""Shared test helpers (TEST INFRASTRUCTURE)."""
import os

import numpy as np
import torch

from xfr_b200 import synth

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), 'golden')
L101 = (3, 4, 23, 3)
L1111 = (1, 1, 1, 1)


def golden(layers):
    tag = '101' if tuple(layers) == L101 else ''.join(map(str, layers))
    return np.load(os.path.join(GOLD, 'stresnet%s_seed0.npz' % tag))


def rel_err(a, b):
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    return float(np.abs(a - b).max() / (np.abs(b).max() + 1e-300))


def golden_inputs(G):
    """The probes / classifier rows oracle/gen_golden.py used: x [2,224,224,3] NHWC (smooth, noise),
    W2 [2,2,512] = (1/2500) * (enc_mate, enc_nonmate) for both probes."""
    imgs = synth.smooth_probes(3, seed=1)
    noise = synth.synthetic_probes(1, seed=2)
    x = torch.cat((imgs[0:1], noise)).permute(0, 2, 3, 1).contiguous()
    W2 = torch.from_numpy(np.concatenate((G['enc_mate'], G['enc_nonmate']))).float() / 2500.0
    return x, W2.unsqueeze(0).repeat(2, 1, 1).contiguous(), imgs


class ShadowBackend(object):
    """Runs every kernel call on the CUDA backend and, with the SAME inputs copied to the host,
    on the torch emulation; records the worst relative error per kernel name."""
    OUTPUTS = {  # positional indices of output tensors per method
        'stem_fwd': (2, 3), 'subsample2': (1,), 'avgpool2': (1,), 'conv_dual': (2, 3, 4),
        'head_fwd': (2, 3, 4, 5, 6), 'head_bwd': (8,), 'dgrad_mid': (6,), 'dgrad_plain': (2,),
        'dgrad_join': (10, 11), 'join': (11, 12), 'ds_res': (3,), 'stem_bwd': (6, 7, 8),
        'contrast': (3,), 'saliency_post': (1,), 'trunc_threshold': (4,), 'bn_hook': (4,), 'head_fwd_linear': (2, 3),
        'head_bwd_linear': (5,),
    }

    def __init__(self, cuda_be, emul_be, pack_map):
        self.cuda, self.emul, self.pack_map = cuda_be, emul_be, pack_map
        self.errors = {}
        self.name = 'shadow'

    def __getattr__(self, name):
        if name not in self.OUTPUTS:
            raise AttributeError(name)

        def call(*args, **kw):
            outs = self.OUTPUTS[name]
            torch.cuda.synchronize()
            cargs = []          # host copies taken BEFORE the CUDA call (some outputs are also read: accumulate)
            for i, a in enumerate(args):
                if torch.is_tensor(a):
                    cargs.append(a.detach().cpu().clone())
                elif id(a) in self.pack_map:
                    cargs.append(self.pack_map[id(a)])
                else:
                    cargs.append(a)
            ckw = {k: (v.detach().cpu().clone() if torch.is_tensor(v) else v) for k, v in kw.items()}
            getattr(self.cuda, name)(*args, **kw)
            torch.cuda.synchronize()
            getattr(self.emul, name)(*cargs, **ckw)
            for i in outs:
                if i >= len(args) or args[i] is None:
                    continue
                got, want = args[i].detach().cpu().double(), cargs[i].double()
                fin = torch.isfinite(want)
                assert bool((torch.isfinite(got) == fin).all()), '%s: finiteness differs' % name
                err = float((got[fin] - want[fin]).abs().max() / (want[fin].abs().max() + 1e-300)) if fin.any() else 0.0
                key = '%s[%d]' % (name, i)
                self.errors[key] = max(self.errors.get(key, 0.0), err)
        return call


def pack_map(eng_gpu, eng_cpu):
    m = {id(eng_gpu.stem): eng_cpu.stem, id(eng_gpu.head): eng_cpu.head}
    for bg, bc in zip(eng_gpu.blocks, eng_cpu.blocks):
        for k in ('c1', 'c2', 'c3', 'cp'):
            if getattr(bg, k, None) is not None:
                m[id(getattr(bg, k))] = getattr(bc, k)
    return m


R50_WEIGHTS = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), 'oracle', '_ref', 'resnet50_128.pth')
R50_MEAN = (131.0912, 103.8827, 91.4953)


def r50_inputs():
    """Real VGGFace2 ResNet-50-128d weights + the bundled real triplet / demo face (crops stored in the golden file).
    Returns (state_dict, G, X dict of [1,3,224,224] tensors) or None when the 95 MB weight file did not travel."""
    if not os.path.exists(R50_WEIGHTS):
        return None
    G = np.load(os.path.join(GOLD, 'resnet50_128_real.npz'))
    sd = torch.load(R50_WEIGHTS)
    X = {k: torch.from_numpy((G['crop_' + k].astype(np.float64) - np.array(R50_MEAN)).transpose(2, 0, 1).astype(np.float32)).unsqueeze(0)
         for k in ('probe', 'mate', 'nonmate', 'demo')}
    return sd, G, X
