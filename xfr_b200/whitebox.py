"""Drop-in mirror of the reference plugin API `xfr.models.whitebox` (reference
python/xfr/models/whitebox.py:25-110, 261-304, 482-558, 739-785) on top of the B200 engine.

Same class names, constructor arguments, method names, argument meaning, return types
(numpy float32 maps) and error behaviour as the reference, so that callers such as
demo/test_whitebox.py and xfr/inpainting_game/generate_whitebox_saliency.py keep working;
underneath, no torch hook / autograd runs: every call goes to the hand-written sm_100a
kernels through xfr_b200.kernels (and raises if they are unavailable).

Additions (not in the reference, which is batch-1 only): `Whitebox.ebp_batch`,
`Whitebox.contrastive_ebp_batch`, `WhiteboxSTResnet.set_triplet_classifiers` (one
(mate, non-mate) pair per probe) — the batched entry points bench.py measures.

Status of the reference surface in this round:
  done   encode / classify / set_triplet_classifier / num_classes / preprocess / clear,
         ebp, contrastive_ebp, truncated_contrastive_ebp (all four ebp_subtree_mode values, ebp_version 6 post-processing),
         embeddings, ebp_subtree_mode, with_bias / ebp_version 11, ebp_version != 6 (uint8 + PIL blur on the host)
  next   layerwise_ebp, layerwise_contrastive_ebp, weighted_subtree_ebp,
         the hooked (non-triplet) fc2 head
"""
import numpy as np
import torch
import torch.nn as nn

from . import synth
from .engine import MODE_IDS, Resnet50_128Engine, StResnetEngine

_CHUNK = 128    # probes per engine sweep (workspace = ~210 MB per probe)


class WhiteboxNetwork(object):
    """reference whitebox.py:25-85"""

    def __init__(self, net):
        self.net = net
        self.net.eval()

    def encode(self, x):
        raise NotImplementedError

    def classify(self, x):
        raise NotImplementedError

    def clear(self):
        """Reference zeroes parameter grads (whitebox.py:66-71); nothing accumulates here."""
        return None

    def set_triplet_classifier(self, x_mate, x_nonmate):
        raise NotImplementedError

    def num_classes(self):
        raise NotImplementedError

    def preprocess(self, im):
        raise NotImplementedError


class WhiteboxSTResnet(WhiteboxNetwork):
    """STR-Janus ResNet-101 plugin (reference whitebox.py:87-110).

    `net` is a torch module (or a plain state_dict) with the reference's parameter names
    (reference resnet.py:168-221: conv1, bn1, layer{1-4}.{i}.conv{1,2,3}/bn{1,2,3}, fc1, fc2).
    The reference's own `xfr.models.resnet.ResNet` instance works unchanged."""

    def __init__(self, net, layers=None, impl='tf32x3'):
        if isinstance(net, dict):
            self._sd = net
            self.net = _StateDictModule(net)
        else:
            self.net = net
            self.net.eval()
            self._sd = net.state_dict()
        self._layers = tuple(layers) if layers is not None else _infer_layers(self._sd)
        self._impl = impl
        self._engine = None
        self._W2 = None            # [n, C, 512] un-hooked triplet rows (whitebox.py:93-96), or None = the net's own fc2
        self._ncls = int(self._sd['fc2.weight'].shape[0]) if 'fc2.weight' in self._sd else 0

    # -- engine plumbing
    def _device(self):
        for v in self._sd.values():
            if v.is_cuda:
                return v.device
        raise RuntimeError('xfr_b200: the network must live on a CUDA device (.to("cuda")); there is no CPU path')

    def engine(self, with_bias=False):
        if self._engine is None or self._engine.with_bias != with_bias:
            from .kernels import CudaBackend
            dev = self._device()
            self._engine = StResnetEngine(self._sd, CudaBackend(dev, impl=self._impl), self._layers, device=dev,
                                          with_bias=with_bias)
        return self._engine

    def _nhwc(self, x):
        dev = self._device()
        x = x.detach().to(dev, dtype=torch.float32, non_blocking=True)
        return x.permute(0, 2, 3, 1).contiguous()

    # -- reference API
    def set_triplet_classifier(self, x_mate, x_nonmate):
        """whitebox.py:93-96: fc2 <- un-hooked Linear(512, 2) with rows (x_mate, x_nonmate)."""
        self._W2 = torch.cat((x_mate.detach().reshape(1, -1), x_nonmate.detach().reshape(1, -1)), dim=0).float().unsqueeze(0)
        self._ncls = 2

    def set_triplet_classifiers(self, x_mates, x_nonmates):
        """Batched extension: probe i is scored against rows (x_mates[i], x_nonmates[i])."""
        self._W2 = torch.stack((x_mates.detach().float(), x_nonmates.detach().float()), dim=1).contiguous()
        self._ncls = 2

    def triplet_rows(self, n):
        if self._W2 is None:
            raise NotImplementedError('xfr_b200: the hooked fc2 head (no set_triplet_classifier) is not on the CUDA path yet')
        W2 = self._W2.to(self._device())
        if W2.shape[0] == 1 and n > 1:
            W2 = W2.expand(n, -1, -1)
        if W2.shape[0] != n:
            raise ValueError('%d probes but %d triplet classifiers' % (n, W2.shape[0]))
        return W2.contiguous()

    def encode(self, x):
        """whitebox.py:98-100: 50 * L2-normalised fc1 output, [N,512] on the input's device."""
        eng = self.engine()
        out = []
        for i in range(0, x.shape[0], _CHUNK):
            out.append(50.0 * eng.forward(self._nhwc(x[i:i + _CHUNK])).clone())
        return torch.cat(out)

    def classify(self, x):
        enc = self.encode(x)
        if self._W2 is not None:
            return torch.einsum('nd,ncd->nc', enc, self.triplet_rows(enc.shape[0]))
        return enc @ self._sd['fc2.weight'].to(enc.device).t() + self._sd['fc2.bias'].to(enc.device)

    def num_classes(self):
        return self._ncls

    def preprocess(self, im):
        """PIL image -> [1,3,224,224] (whitebox.py:108-110, resnet.py:25-37)."""
        img = np.asarray(im.resize((224, 224)).convert('RGB'), dtype=np.float64) - np.array(synth.MEAN_RGB)
        return torch.from_numpy(np.moveaxis(img, 2, 0)).float().unsqueeze(0)


class Whitebox_resnet50_128(WhiteboxSTResnet):
    """VGGFace2 ResNet-50-128d plugin (reference whitebox.py:210-258): 128-d encoding = feat_extract output, classifier =
    an un-hooked Linear(128, 2) held by the wrapper (whitebox.py:216-230).  `net` is the reference's
    resnet50_128.Resnet50_128 module or its state_dict."""

    def __init__(self, net, impl='tf32x3'):
        if isinstance(net, dict):
            self._sd = net
            self.net = _StateDictModule(net)
        else:
            self.net = net
            self.net.eval()
            self._sd = net.state_dict()
        self._impl = impl
        self._engine = None
        # reference: self.fc1 = nn.Linear(128, 2, bias=False) with default (random) init; here the rows must be set
        self._W2 = None
        self._ncls = 2

    def engine(self, with_bias=False):
        if self._engine is None or self._engine.with_bias != with_bias:
            from .kernels import CudaBackend
            dev = self._device()
            self._engine = Resnet50_128Engine(self._sd, CudaBackend(dev, impl=self._impl), device=dev, with_bias=with_bias)
        return self._engine

    def encode(self, x):
        """whitebox.py:222-224: self.net(x)[0], the 128-d feat_extract output (not normalised)."""
        eng = self.engine()
        return torch.cat([eng.forward(self._nhwc(x[i:i + _CHUNK])).clone() for i in range(0, x.shape[0], _CHUNK)])

    def classify(self, x):
        enc = self.encode(x)
        return torch.einsum('nd,ncd->nc', enc, self.triplet_rows(enc.shape[0]))

    def num_classes(self):
        return 2

    def preprocess(self, img):
        """whitebox.py:235-258: shorter side -> 224 (bilinear), centre crop 224, subtract the VGGFace2 mean."""
        import PIL.Image
        mean = (131.0912, 103.8827, 91.4953)
        im_shape = np.array(img.size)
        img = img.convert('RGB')
        ratio = 224.0 / np.min(im_shape)
        img = img.resize(size=(int(np.ceil(im_shape[0] * ratio)), int(np.ceil(im_shape[1] * ratio))), resample=PIL.Image.BILINEAR)
        x = np.array(img)
        h0, w0 = (x.shape[0] - 224) // 2, (x.shape[1] - 224) // 2
        x = x[h0:h0 + 224, w0:w0 + 224] - mean
        return torch.from_numpy(x.transpose(2, 0, 1).astype(np.float32)).unsqueeze(0)


class _StateDictModule(nn.Module):
    def __init__(self, sd):
        super(_StateDictModule, self).__init__()
        self._sd = sd

    def state_dict(self, *a, **k):
        return self._sd

    def parameters(self, recurse=True):
        return iter([v for v in self._sd.values() if v.is_floating_point()])


def _infer_layers(sd):
    return tuple(len({k.split('.')[1] for k in sd if k.startswith('layer%d.' % li)}) for li in (1, 2, 3, 4))


class Whitebox(nn.Module):
    """reference whitebox.py:261-304"""

    def __init__(self, net, ebp_version=None, with_bias=None, eps=1E-16, ebp_subtree_mode='affineonly_with_prior'):
        super(Whitebox, self).__init__()
        assert isinstance(net, WhiteboxNetwork)
        self.net = net
        self.eps = eps
        self.layerlist = None
        self.ebp_ver = ebp_version
        if self.ebp_ver is None:
            self.ebp_ver = 6
        elif self.ebp_ver < 4:
            raise RuntimeError('ebp version, if set, must be at least 4')
        self.convert_saliency_uint8 = (self.ebp_ver != 6)
        self._ebp_with_bias = with_bias if with_bias is not None else self.ebp_ver == 11
        self.dA, self.A, self.X, self.P, self.P_prior, self.P_layername = [], [], [], [], [], []
        self.batch_size = 32
        self._ebp_mode = 'disable'
        if ebp_subtree_mode not in MODE_IDS:
            raise ValueError('Invalid subtree mode "%s"' % ebp_subtree_mode)
        self._ebp_subtree_mode = ebp_subtree_mode

    # ---------------------------------------------------------------- helpers
    def _engine(self):
        if self.eps != 1E-16 and abs(self.eps - self.net.engine().be.eps) > 0:
            self.net.engine().be.eps = float(self.eps)
        return self.net.engine(self._ebp_with_bias)

    def _float32_to_uint8(self, img):
        """whitebox.py:439-441"""
        return np.uint8(255 * ((img - np.min(img)) / (self.eps + (np.max(img) - np.min(img)))))

    def _mwp_to_saliency_uint8(self, P, blur_radius=2):
        """ebp_version != 6 branch of _mwp_to_saliency (whitebox.py:451-454): 8-bit quantise, PIL GaussianBlur,
        re-quantise.  PIL's filter is the reference's own dependency; it runs on the host on a 112x112 uint8 image."""
        import PIL.Image
        import PIL.ImageFilter
        img = self._float32_to_uint8(P)
        img = np.array(PIL.Image.fromarray(img).filter(PIL.ImageFilter.GaussianBlur(radius=blur_radius)))
        return self._float32_to_uint8(img)

    # ---------------------------------------------------------------- batched entry points
    def ebp_batch(self, x, Pn, mwp=False):
        """x [N,3,224,224], Pn [N,C] (or [1,C]) -> float32 [N,112,112] maps (numpy)."""
        eng = self._engine()
        outs = []
        N = x.shape[0]
        W2 = self.net.triplet_rows(N)
        Pn = Pn.to(W2.device, dtype=torch.float32)
        if Pn.shape[0] == 1 and N > 1:
            Pn = Pn.expand(N, -1)
        for i in range(0, N, _CHUNK):
            m = eng.ebp(self.net._nhwc(x[i:i + _CHUNK]), Pn[i:i + _CHUNK].contiguous(), W2[i:i + _CHUNK].contiguous(),
                        self._ebp_subtree_mode, saliency=not (mwp or self.convert_saliency_uint8))
            outs.append(m.cpu())
        maps = torch.cat(outs).numpy()
        if self.convert_saliency_uint8 and not mwp:
            maps = np.stack([self._mwp_to_saliency_uint8(m) for m in maps])
        return maps

    def contrastive_ebp_batch(self, x, k_poschannel=0, k_negchannel=1, out=None, percentile=None):
        """N probes, each against its own (mate, non-mate) rows -> float32 [N,112,112] (numpy, or `out` pinned tensor)."""
        eng = self._engine()
        N = x.shape[0]
        W2 = self.net.triplet_rows(N)
        res = out if out is not None else torch.empty(N, 112, 112)
        for i in range(0, N, _CHUNK):
            m = eng.contrastive(self.net._nhwc(x[i:i + _CHUNK]), W2[i:i + _CHUNK].contiguous(), k_poschannel, k_negchannel,
                                self._ebp_subtree_mode, percentile=percentile, saliency=not self.convert_saliency_uint8)
            res[i:i + m.shape[0]].copy_(m, non_blocking=True)
        torch.cuda.current_stream(W2.device).synchronize()
        if self.convert_saliency_uint8:
            return np.stack([self._mwp_to_saliency_uint8(m) for m in res.numpy()])
        return res if out is not None else res.numpy()

    # ---------------------------------------------------------------- reference API (batch 1)
    def ebp(self, x, Pn, mwp=False):
        """whitebox.py:482-504"""
        return self.ebp_batch(x, Pn, mwp)[0]

    def contrastive_ebp(self, img_probe, k_poschannel, k_negchannel):
        """whitebox.py:506-527"""
        assert(k_poschannel >= 0 and k_poschannel < self.net.num_classes())
        assert(k_negchannel >= 0 and k_negchannel < self.net.num_classes())
        return self.contrastive_ebp_batch(img_probe, k_poschannel, k_negchannel)[0]

    def truncated_contrastive_ebp(self, img_probe, k_poschannel, k_negchannel, percentile=20):
        """whitebox.py:529-558"""
        assert(k_poschannel >= 0 and k_poschannel < self.net.num_classes())
        assert(k_negchannel >= 0 and k_negchannel < self.net.num_classes())
        return self.contrastive_ebp_batch(img_probe, k_poschannel, k_negchannel, percentile=percentile)[0]

    def layerwise_ebp(self, img_probe, k_layer, mode='argmax', k_element=None, k_poschannel=0, mwp=True):
        raise NotImplementedError('xfr_b200: layerwise_ebp is scheduled next (DESIGN.md)')

    def layerwise_contrastive_ebp(self, *a, **k):
        raise NotImplementedError('xfr_b200: layerwise_contrastive_ebp is scheduled next (DESIGN.md)')

    def weighted_subtree_ebp(self, *a, **k):
        raise NotImplementedError('xfr_b200: weighted_subtree_ebp is scheduled next (DESIGN.md)')

    def ebp_subtree_mode(self):
        return self._ebp_subtree_mode

    def encode(self, x):
        return self.net.encode(x)

    def embeddings(self, images, norm=True):
        """whitebox.py:747-785 for tensors / arrays already in network format."""
        if isinstance(images[0], torch.Tensor):
            imagesT = torch.stack(list(images)) if not isinstance(images, torch.Tensor) else images
        else:
            imagesT = torch.stack([torch.from_numpy(np.asarray(im)).float() for im in images])
        embeds = self.encode(imagesT).detach().cpu().numpy()
        if norm:
            embeds = embeds / np.linalg.norm(embeds.reshape(embeds.shape[0], -1), axis=1, keepdims=True)
        return embeds
