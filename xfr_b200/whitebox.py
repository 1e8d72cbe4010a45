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
  done   layerwise_ebp, layerwise_contrastive_ebp, weighted_subtree_ebp on the STR ResNet plugin (xfr_b200/generic.py)
  done   the hooked (non-triplet) fc2 head for ebp / contrastive_ebp (e.g. the 65,359-class STR head, blackbox.py:280-294)
  done   Light-CNN-29v2 plugin (WhiteboxLightCNN): every operator above incl. layerwise / weighted-subtree (xfr_b200/lightcnn.py)
  next   layerwise operators for the ResNet-50-128d plugin
"""
import numpy as np
import torch
import torch.nn as nn

from . import synth
from .engine import MODE_IDS, Resnet50_128Engine, StResnetEngine
from .lightcnn import LightCNNEngine

_CHUNK = 128    # probes per engine sweep (workspace = ~210 MB per probe)
_MAX_ROWS = 64  # gradient rows of one firing-by-firing sweep (generic.MAX_ROWS)
# The GEMM plan of the fused sweep (xfr_b200/kernels.py).  'bf16x2' (tcgen05 kind::f16 on bf16 terms, pair-tensor operands) is the
# default because on the B200 it is as close to the reference as the three-pass TF32 plan: what separates every tensor-core plan
# from the reference is the tensor core's truncating fp32 accumulation, not the operand precision (on the
# well-conditioned ResNet-101 triplet bf16x2 / tf32x3 / tf32x3full sit at 3.8e-3 / 2.8e-3 / 3.0e-3 of the contrastive map's
# maximum, all <= 4e-6 max-abs against the 1e-4 bar).  'fp32' (CUDA cores, exact products, round-to-nearest accumulation) is the
# reference-grade plan: 1e-4 of the maximum, at a seventh of the speed.
DEFAULT_IMPL = 'bf16x2'


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

    _CONTAINERS = ('Sequential', 'Bottleneck')

    def _layer_visitor(self, f_forward=None, f_preforward=None, net=None):
        """reference whitebox.py:34-56 (Light-CNN override 142-159): the modules of the torch network in definition order,
        containers expanded, as [{'name': str(layer), 'hooks': [...]}].  The B200 engine runs a static schedule instead of
        hooks, so nothing depends on this list; hook functions, if given, are still registered on the torch modules (they fire
        only if the caller runs the torch network itself).  A plain state_dict has no modules: one entry per parameter owner."""
        net = self.net if net is None else net
        if isinstance(net, _StateDictModule):
            owners = []
            for k in net.state_dict():
                o = k.rsplit('.', 1)[0]
                if o not in owners:
                    owners.append(o)
            return [{'name': o, 'hooks': []} for o in owners]
        layerlist = []
        for name, layer in net._modules.items():
            if type(layer).__name__ in self._CONTAINERS:
                layerlist.append({'name': str(layer), 'hooks': [None]})
                layerlist = layerlist + self._layer_visitor(f_forward, f_preforward, layer)
            else:
                hooks = []
                if f_forward is not None:
                    hooks.append(layer.register_forward_hook(f_forward))
                if f_preforward is not None:
                    hooks.append(layer.register_forward_pre_hook(f_preforward))
                layerlist.append({'name': str(layer), 'hooks': hooks})
        return layerlist


class WhiteboxSTResnet(WhiteboxNetwork):
    """STR-Janus ResNet-101 plugin (reference whitebox.py:87-110).

    `net` is a torch module (or a plain state_dict) with the reference's parameter names
    (reference resnet.py:168-221: conv1, bn1, layer{1-4}.{i}.conv{1,2,3}/bn{1,2,3}, fc1, fc2).
    The reference's own `xfr.models.resnet.ResNet` instance works unchanged."""

    def __init__(self, net, layers=None, impl=DEFAULT_IMPL):
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
        """Where the engine runs: the network's CUDA device, or - for a CPU-resident network, which is how demo/test_whitebox.py
        builds its Whitebox objects (:77-107, 257-280) - the current CUDA device: the weights are packed onto the GPU once and
        every kernel runs there (this is not a CPU path: without a B200 it raises)."""
        for v in self._sd.values():
            if v.is_cuda:
                return v.device
        if torch.cuda.is_available():
            return torch.device('cuda', torch.cuda.current_device())
        raise RuntimeError('xfr_b200: no CUDA device - the whitebox engine runs on a B200 only (there is no CPU path)')

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

    def hooked(self):
        """True when no triplet classifier was set: the network's own (hooked) fc2 is the classifier."""
        return self._W2 is None

    def triplet_rows(self, n):
        if self._W2 is None:                      # the network's own fc2 [C,512]: takes part in EBP with relu(W)
            w = self._sd['fc2.weight']
            key = (w.data_ptr(), w._version, str(self._device()))
            if getattr(self, '_fc2_dev', (None, None))[0] != key:         # one device copy (134 MB for the 65,359-class STR head)
                self._fc2_dev = (key, w.detach().to(self._device()).float().contiguous())
            return self._fc2_dev[1]
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
        return torch.cat(out).to(x.device)

    _ENC_SCALE = 50.0      # encode() = 50 * L2-normalised fc1 (whitebox.py:98-100); the other two plugins return the raw feature

    def encode_nhwc(self, x_nhwc):
        """encode() for probes that are already on the device in the engine's NHWC layout (xfr_b200/inpaintgame.py)."""
        eng = self.engine()
        out = [eng.forward(x_nhwc[i:i + _CHUNK]).clone() for i in range(0, x_nhwc.shape[0], _CHUNK)]
        enc = torch.cat(out)
        return enc * self._ENC_SCALE if self._ENC_SCALE != 1.0 else enc

    def classify(self, x):
        enc = self.encode(x)
        if self._W2 is not None:
            return torch.einsum('nd,ncd->nc', enc, self.triplet_rows(enc.shape[0]).to(enc.device))
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

    _ENC_SCALE = 1.0

    def __init__(self, net, impl=DEFAULT_IMPL):
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
        return torch.cat([eng.forward(self._nhwc(x[i:i + _CHUNK])).clone() for i in range(0, x.shape[0], _CHUNK)]).to(x.device)

    def classify(self, x):
        enc = self.encode(x)
        return torch.einsum('nd,ncd->nc', enc, self.triplet_rows(enc.shape[0]).to(enc.device))

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


class Whitebox_senet50_256(Whitebox_resnet50_128):
    """VGGFace2 SENet-50-256d plugin (reference whitebox.py:163-208).  Its squeeze-and-excitation gates are Sigmoid modules,
    for which the reference's own EBP hooks raise ValueError ('... is a special case ... and is not yet supported',
    whitebox.py:404-405, 413-414, 421-422): no excitation-backprop operator of the reference runs on this net, so there is
    no hot path to accelerate.  The class exists so that code which names it imports; preprocess() is the reference's
    (identical to the ResNet-50-128d plugin's, whitebox.py:185-208)."""

    def __init__(self, net, impl=DEFAULT_IMPL):
        self.net = net
        self._sd = {}
        self._impl = impl
        self._engine = None
        self._W2 = None
        self._ncls = 2

    def engine(self, with_bias=False):
        raise ValueError('Whitebox_senet50_256: Sigmoid layers are a special case of excitation backprop that the reference '
                         'does not support either (whitebox.py:404); no EBP operator is available for this network')

    def encode(self, x):
        raise NotImplementedError('Whitebox_senet50_256.encode: the SENet forward is not part of the B200 hot path')


class WhiteboxLightCNN(WhiteboxSTResnet):
    """Light-CNN-29v2 plugin (reference whitebox.py:113-159).  `net` is the reference's
    xfr.models.lightcnn.network_29layers_v2 module (lightcnn.py:216-275) or its state_dict.  The classifier is the
    network's fc2 (hooked, W+ in the backward) until set_triplet_classifier replaces it by an un-hooked Linear(256, 2)
    (whitebox.py:120-123).  Saliency maps are 128x128 (P[-2] is the first Split input, 96x128x128)."""
    _ENC_SCALE = 1.0
    _CONTAINERS = ('Sequential', 'Bottleneck', 'mfm', 'group', 'resblock')       # whitebox.py:149

    def __init__(self, net, impl=DEFAULT_IMPL):
        if isinstance(net, dict):
            self._sd = net
            self.net = _StateDictModule(net)
        else:
            self.net = net
            self.net.eval()
            self._sd = net.state_dict()
        self._impl = impl
        self._engine = None
        self._W2 = None
        self._ncls = int(self._sd['fc2.weight'].shape[0]) if 'fc2.weight' in self._sd else 0

    def engine(self, with_bias=False):
        if self._engine is None or self._engine.with_bias != with_bias:
            from .kernels import CudaBackend
            dev = self._device()
            self._engine = LightCNNEngine(self._sd, CudaBackend(dev, impl=self._impl), device=dev, with_bias=with_bias)
        return self._engine

    def encode(self, x):
        """whitebox.py:125-128: the 256-d fc output (`features` of net(x))."""
        eng = self.engine()
        return torch.cat([eng.forward(self._nhwc(x[i:i + _CHUNK])).clone() for i in range(0, x.shape[0], _CHUNK)]).to(x.device)

    def classify(self, x):
        """whitebox.py:130-132: fc2(dropout(fc)) in eval mode; fc2 has no bias (lightcnn.py:228)."""
        enc = self.encode(x)
        if self._W2 is not None:
            return torch.einsum('nd,ncd->nc', enc, self.triplet_rows(enc.shape[0]).to(enc.device))
        return enc @ self._sd['fc2.weight'].to(enc.device).t()

    def preprocess(self, im):
        """whitebox.py:137-139 / lightcnn.py:19-31: Resize(144) (shorter side, bilinear), CenterCrop(128), luminance in [0,1]
        with skimage.color.rgb2gray's weights -> [1,1,128,128]."""
        import PIL.Image
        w, h = im.size
        if w <= h:
            nw, nh = 144, int(144 * h / w)
        else:
            nw, nh = int(144 * w / h), 144
        im = im.resize((nw, nh), PIL.Image.BILINEAR)
        l, t = int(round((nw - 128) / 2.0)), int(round((nh - 128) / 2.0))
        a = np.array(im.crop((l, t, l + 128, t + 128)))
        if a.ndim == 3:
            a = (a[..., :3].astype(np.float64) / 255.0) @ np.array([0.2125, 0.7154, 0.0721])
        return torch.from_numpy(np.asarray(a)).float().unsqueeze(0).unsqueeze(0)


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
        # the reference registers its hooks here (whitebox.py:303); the engine needs none - the list of layers is kept for callers
        self.layerlist = self.net._layer_visitor()
        # record_P = True: ebp() also fills self.P / self.P_layername with the MWP of every hook firing, as the reference's hooks
        # do on every call (whitebox.py:388-395); off by default - the fused sweep never materialises them (layerwise_ebp,
        # layerwise_contrastive_ebp and weighted_subtree_ebp always record what they need)
        self.record_P = False
        if ebp_subtree_mode not in MODE_IDS:
            raise ValueError('Invalid subtree mode "%s"' % ebp_subtree_mode)
        self._ebp_subtree_mode = ebp_subtree_mode

    # ---------------------------------------------------------------- helpers
    def _engine(self):
        """The plugin's engine for this object's with_bias setting, with this object's eps (whitebox.py:267, 325-327): fetched
        once, eps set unconditionally - several Whitebox objects may share one plugin."""
        eng = self.net.engine(self._ebp_with_bias)
        eng.be.eps = eng.eps = float(self.eps)
        return eng

    def _logits(self, eng, W2):
        """classify() of probe 0 after eng.forward(): the un-hooked triplet rows, or the network's own fc2 (with its bias)."""
        if not self.net.hooked():
            return eng.logits(W2)
        y = eng.hooked_logits(W2)
        b2 = self.net._sd.get('fc2.bias')
        return y if b2 is None else y + b2.to(y.device).float()

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
        hk = self.net.hooked()
        # the maps of every chunk go to ONE pinned host tensor with asynchronous copies (a pageable .cpu() per chunk cost 3 ms per
        # 128 Light-CNN maps: 9 % of the sweep)
        res = torch.empty(N, eng.map_hw, eng.map_hw, pin_memory=bool(W2.is_cuda))      # torch's caching host allocator: no cudaHostAlloc per call
        for i in range(0, N, _CHUNK):
            m = eng.graph_call('ebp', (self.net._nhwc(x[i:i + _CHUNK]), Pn[i:i + _CHUNK].contiguous(), W2 if hk else W2[i:i + _CHUNK].contiguous()),
                               mode=self._ebp_subtree_mode, hooked_fc2=hk, saliency=not (mwp or self.convert_saliency_uint8))
            res[i:i + m.shape[0]].copy_(m, non_blocking=True)
        if W2.is_cuda:
            torch.cuda.current_stream(W2.device).synchronize()
        maps = res.numpy()
        if self.convert_saliency_uint8 and not mwp:
            maps = np.stack([self._mwp_to_saliency_uint8(m) for m in maps])
        return maps

    def contrastive_ebp_batch(self, x, k_poschannel=0, k_negchannel=1, out=None, percentile=None):
        """N probes, each against its own (mate, non-mate) rows -> float32 [N,112,112] (numpy, or `out` pinned tensor)."""
        eng = self._engine()
        N = x.shape[0]
        W2 = self.net.triplet_rows(N)
        res = out if out is not None else torch.empty(N, eng.map_hw, eng.map_hw)
        hk = self.net.hooked()
        for i in range(0, N, _CHUNK):
            m = eng.graph_call('contrastive', (self.net._nhwc(x[i:i + _CHUNK]), W2 if hk else W2[i:i + _CHUNK].contiguous()),
                               k_pos=k_poschannel, k_neg=k_negchannel, mode=self._ebp_subtree_mode, hooked_fc2=hk, percentile=percentile,
                               saliency=not self.convert_saliency_uint8, num_classes=self.net.num_classes())
            res[i:i + m.shape[0]].copy_(m, non_blocking=True)
        if W2.is_cuda:
            torch.cuda.current_stream(W2.device).synchronize()
        if self.convert_saliency_uint8:
            return np.stack([self._mwp_to_saliency_uint8(m) for m in res.numpy()])
        return res if out is not None else res.numpy()

    def contrastive_ebp_stream(self, batches, k_poschannel=0, k_negchannel=1, outs=None, percentile=None):
        """Streaming form of contrastive_ebp_batch for host-resident probes: `batches` is a sequence of [n,3,H,W] host tensors
        (pinned memory makes the copies asynchronous), each swept against the classifier rows set for its n probes
        (set_triplet_classifiers; an item may also be (x, x_mates, x_nonmates) with its own rows).  The host-to-device copy of
        chunk i+1 runs on a copy stream while chunk i is swept, and the finished maps go back with an asynchronous copy, so the
        copies stay off the critical path (3.4 ms of a 69 ms step at 256 probes otherwise).  Returns the list of [n,h,w] float32
        maps (pinned host tensors, `outs` if given); results are identical to contrastive_ebp_batch."""
        eng = self._engine()
        dev = eng.device
        if getattr(eng.be, 'name', '') != 'cuda' or self.convert_saliency_uint8:
            res = []
            for b, item in enumerate(batches):
                x = item[0] if isinstance(item, (tuple, list)) else item
                if isinstance(item, (tuple, list)):
                    self.net.set_triplet_classifiers(item[1], item[2])
                m = self.contrastive_ebp_batch(x, k_poschannel, k_negchannel, percentile=percentile)
                m = torch.from_numpy(np.ascontiguousarray(m))
                if outs is not None:
                    outs[b].copy_(m)
                    m = outs[b]
                res.append(m)
            return res
        hk = self.net.hooked()
        work, res = [], []
        for b, item in enumerate(batches):
            x = item[0] if isinstance(item, (tuple, list)) else item
            if isinstance(item, (tuple, list)):
                self.net.set_triplet_classifiers(item[1], item[2])
            N = x.shape[0]
            W2 = self.net.triplet_rows(N)
            out = outs[b] if outs is not None else torch.empty(N, eng.map_hw, eng.map_hw).pin_memory()
            res.append(out)
            for i in range(0, N, _CHUNK):
                work.append((x[i:i + _CHUNK], W2 if hk else W2[i:i + _CHUNK].contiguous(), out[i:i + _CHUNK]))
        if not work:
            return res
        st = getattr(self, '_stream_state', None)
        shape = (max(w[0].shape[0] for w in work),) + tuple(work[0][0].shape[1:])        # staging capacity: the largest chunk
        if any(tuple(w[0].shape[1:]) != shape[1:] for w in work):
            raise ValueError('contrastive_ebp_stream: every batch must have the same image shape')
        if st is None or st['shape'][1:] != shape[1:] or st['shape'][0] < shape[0] or st['dev'] != dev:
            st = self._stream_state = {'shape': shape, 'dev': dev, 'copy': torch.cuda.Stream(dev),
                                       'stage': [torch.empty(shape, dtype=torch.float32, device=dev) for _ in range(2)],
                                       'landed': [torch.cuda.Event() for _ in range(2)], 'free': [torch.cuda.Event() for _ in range(2)]}
        main = torch.cuda.current_stream(dev)
        for e in st['free']:
            e.record(main)

        def prefetch(i):
            xs, s = work[i][0], i % 2
            with torch.cuda.stream(st['copy']):
                st['copy'].wait_event(st['free'][s])                   # the sweep that read this staging buffer is past its repack
                st['stage'][s][:xs.shape[0]].copy_(xs, non_blocking=True)
                st['landed'][s].record(st['copy'])
        prefetch(0)
        for i, (xs, ws, out) in enumerate(work):
            s, n = i % 2, xs.shape[0]
            main.wait_event(st['landed'][s])
            x_nhwc = st['stage'][s][:n].permute(0, 2, 3, 1).contiguous()
            st['free'][s].record(main)
            if i + 1 < len(work):
                prefetch(i + 1)
            m = eng.graph_call('contrastive', (x_nhwc, ws), k_pos=k_poschannel, k_neg=k_negchannel, mode=self._ebp_subtree_mode,
                               hooked_fc2=hk, percentile=percentile, saliency=True, num_classes=self.net.num_classes())
            out.copy_(m, non_blocking=True)
        main.synchronize()
        return res

    # ---------------------------------------------------------------- reference API (batch 1)
    def ebp(self, x, Pn, mwp=False):
        """whitebox.py:482-504"""
        if self.record_P and x.shape[0] == 1:
            return self._ebp_recorded(x, Pn, mwp)
        return self.ebp_batch(x, Pn, mwp)[0]

    def _ebp_recorded(self, x, Pn, mwp):
        """ebp() through the firing-by-firing sweep, keeping self.P / self.P_layername like the reference (record_P)."""
        gs, W2 = self._generic(x)
        P, names, P2 = gs.run(Pn.to(W2.device, dtype=torch.float32), W2, self._ebp_subtree_mode, record=True, hooked_fc2=self.net.hooked())
        self._set_P(gs, P, names)
        return self._finish_map(P2, mwp)[0]

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

    # ---------------------------------------------------------------- layerwise / sub-tree operators (generic sweep)
    def _generic(self, img_probe):
        """Forward once, return (firing-by-firing sweep of the engine, W2 rows)."""
        eng = self._engine()
        if not hasattr(eng, 'sweep'):
            raise NotImplementedError('xfr_b200: layerwise / weighted-subtree EBP is implemented for the STR ResNet and '
                                      'Light-CNN plugins')
        if img_probe.shape[0] != 1:
            raise ValueError('layerwise operators take one probe (as in the reference)')
        eng.forward(self.net._nhwc(img_probe))
        return eng.sweep(), self.net.triplet_rows(1)

    def _set_P(self, gs, P, names):
        """Expose the recorded MWPs like the reference's self.P / self.P_layername: [1,C,H,W] views (vectors: [1,C])."""
        self.P = [gs.to_reference(k, p) for k, p in enumerate(P)]
        self.P_layername = list(names)

    def _finish_map(self, P2, mwp):
        chansum = P2.sum(-1)                                      # [J,112,112] pool over channels (whitebox.py:499)
        if mwp:
            return chansum.cpu().numpy().astype(np.float32)
        out = torch.empty_like(chansum)
        self.net.engine(self._ebp_with_bias).be.saliency_post(chansum.contiguous(), out)
        maps = out.cpu().numpy()
        if self.convert_saliency_uint8:
            maps = np.stack([self._mwp_to_saliency_uint8(m) for m in chansum.cpu().numpy()])
        return maps

    def _zero_map(self, mwp):
        """A prior at the last firing (the Conv2d hook on the image, not recorded) cannot reach P[-2]: the reference's third
        pass starts from a zero seed and returns an all-zero map."""
        hw = self.net.engine(self._ebp_with_bias).map_hw
        return np.zeros((hw, hw), dtype=np.uint8 if (self.convert_saliency_uint8 and not mwp) else np.float32)

    def _onehot(self, k, dev):
        P0 = torch.zeros((1, self.net.num_classes()), device=dev)
        P0[0][k] = 1.0
        return P0

    def layerwise_ebp(self, img_probe, k_layer, mode='argmax', k_element=None, k_poschannel=0, mwp=True):
        """whitebox.py:561-581: EBP restarted from one node (or the arg-max nodes) of firing k_layer."""
        assert(k_poschannel >= 0 and k_poschannel < self.net.num_classes())
        gs, W2 = self._generic(img_probe)
        hk = self.net.hooked()            # no triplet classifier: the network's own fc2 fires (and is indexed) first
        P0 = self._onehot(k_poschannel, W2.device)
        P_mate, names, _ = gs.run(P0, W2, self._ebp_subtree_mode, record=True, hooked_fc2=hk)
        k_layer = int(k_layer) % len(P_mate)                    # the reference indexes Python lists: negative k_layer counts from the image
        if P_mate[k_layer] is None:
            return self._zero_map(mwp)
        Pk = P_mate[k_layer]
        if mode == 'argmax':
            prior = (0, (Pk * (Pk == Pk.max())).reshape(-1).contiguous())
        elif mode == 'elementwise':
            assert(k_element is not None)
            e = gs.elem_index(k_layer, k_element, Pk.shape)
            prior = (0, e, float(Pk.reshape(-1)[e]))
        else:
            raise ValueError('invalid layerwise EBP mode "%s"' % mode)
        P, names, P2 = gs.run(0.0 * P0, W2, self._ebp_subtree_mode, priors={int(k_layer): prior}, record=True, hooked_fc2=hk)
        self._set_P(gs, P, names)
        return self._finish_map(P2, mwp)[0]

    def _contrastive_prior(self, gs, Pm, Pq, k_layer, mode, percentile, k_element):
        """The prior of whitebox.py:603-636 at one firing from the mate / non-mate MWPs recorded there ([1,...] each)."""
        dev = Pm.device
        C = torch.clamp_min(Pm - Pq, 0)
        argmax_only = lambda t: t * (t == t.max())
        if mode == 'copy':
            prior = C
        elif mode == 'mean':
            prior = 0.5 * (Pm + C)
        elif mode == 'product':
            prior = torch.sqrt(Pm.double() * C.double()).float()
        elif mode == 'argmax':
            prior = argmax_only(C)
        elif mode == 'argmax_product':
            prior = argmax_only(torch.sqrt(Pm.double() * C.double()).float())
        elif mode in ('percentile', 'percentile_argmax'):
            assert(percentile >= 0 and percentile <= 100)
            be = gs.be
            thr = torch.empty(1, device=dev)
            sums = Pm.double().sum().reshape(1)
            be.trunc_threshold(Pm.reshape(1, 1, 1, -1).contiguous(), sums, 1, percentile, thr)
            prior = (Pm >= thr) * C
            if mode == 'percentile_argmax':
                prior = argmax_only(prior)
        elif mode == 'elementwise':
            e = gs.elem_index(k_layer, k_element, Pm.shape)
            prior = torch.zeros_like(C).reshape(-1)
            prior[e] = C.reshape(-1)[e]
        else:
            raise ValueError('unknown contrastive ebp mode "%s"' % mode)
        return prior

    def layerwise_contrastive_ebp_sweep(self, img_probe, k_poschannel, k_negchannel, k_layers, mode='percentile', percentile=20,
                                        mwp=False, rows_per_sweep=48):
        """Batched extension for layer sweeps (BASELINE.json configs[2]): layerwise_contrastive_ebp(..., k_layer=k) for every k in
        k_layers of ONE probe -> [len(k_layers), h, w] maps.  The reference runs three ebp() per (probe, layer); the mate /
        non-mate MWPs are the same for every layer, so they are recorded once (one 2-row sweep) and the per-layer third passes,
        each with its prior at a different firing, are batched as gradient rows (`rows_per_sweep` at a time)."""
        assert(k_poschannel >= 0 and k_poschannel < self.net.num_classes())
        assert(k_negchannel >= 0 and k_negchannel < self.net.num_classes())
        if mode == 'elementwise':
            raise ValueError('layerwise_contrastive_ebp_sweep: mode "elementwise" needs a per-layer k_element; call '
                             'layerwise_contrastive_ebp per layer')
        gs, W2 = self._generic(img_probe)
        hk = self.net.hooked()            # no triplet classifier: the network's own fc2 fires (and is indexed) first
        dev = W2.device
        eng = gs.eng
        Pn = torch.cat((self._onehot(k_poschannel, dev), self._onehot(k_negchannel, dev)))
        # every sweep below goes through the engine's graph table: the ~500 launches of a sweep are captured once and replayed,
        # the per-layer priors travel in a device table (generic.PriorTable) instead of launch arguments
        rec = eng.generic_call(Pn, W2, mode=self._ebp_subtree_mode, record=True, hooked_fc2=hk)
        P, names = rec['P'], rec['names']
        k_layers = [int(k) % len(P) for k in k_layers]                 # negative indices as Python lists take them
        # the last firing (the Conv2d hook on the image) is not recorded: a prior there cannot reach P[-2], the map is all zero
        ks = tuple(sorted(k for k in set(k_layers) if P[k] is not None))

        def make_priors():
            return {k: self._contrastive_prior(gs, P[k][0:1], P[k][1:2], k, mode, percentile, None).reshape(-1).contiguous() for k in ks}
        # ~10 small launches per layer: replayed as one graph once the recorded MWPs sit at static addresses (the record sweep's graph)
        priors = eng.graph_fn(('layer_priors', mode, float(percentile), ks, tuple(P[k].data_ptr() for k in ks)), make_priors)
        zero = self._zero_map(mwp)
        del P, rec
        self.P_layername = list(names)
        todo = sorted(priors)                                             # one gradient row per distinct firing
        maps = {}
        tab = eng.prior_table('rows')
        rows_per_sweep = max(1, min(int(rows_per_sweep), _MAX_ROWS))                                    # a prior table describes up to 64 rows
        Z = torch.zeros(min(rows_per_sweep, max(len(todo), 1)), self.net.num_classes(), device=dev)     # fixed row count: one graph
        for i in range(0, len(todo), Z.shape[0]):
            chunk = todo[i:i + Z.shape[0]]
            tab.clear(zero_seed=True)                                     # zero class priors: row r is all zero before its firing
            for r, k in enumerate(chunk):
                tab.set_tensor(k, r, priors[k])
            tab.upload()
            P2 = eng.generic_call(Z, W2, mode=self._ebp_subtree_mode, hooked_fc2=hk, ptab=tab)['P2']
            for k, m in zip(chunk, self._finish_map(P2[:len(chunk)], mwp)):
                maps[k] = m
        return np.stack([maps.get(k, zero) for k in k_layers])

    def layerwise_contrastive_ebp(self, img_probe, k_poschannel, k_negchannel, k_layer, mode='copy', percentile=80, k_element=None,
                                  gradlayer=None, mwp=False):
        """whitebox.py:584-644 (deprecated in the reference in favour of weighted_subtree_ebp, kept for the layer sweeps)."""
        import warnings
        warnings.warn("layerwise_contrastive_ebp is deprecated, use weighted_subtree_ebp instead")
        assert(k_poschannel >= 0 and k_poschannel < self.net.num_classes())
        assert(k_negchannel >= 0 and k_negchannel < self.net.num_classes())
        gs, W2 = self._generic(img_probe)
        hk = self.net.hooked()            # no triplet classifier: the network's own fc2 fires (and is indexed) first
        dev = W2.device
        Pn = torch.cat((self._onehot(k_poschannel, dev), self._onehot(k_negchannel, dev)))
        P, names, _ = gs.run(Pn, W2, self._ebp_subtree_mode, record=True, hooked_fc2=hk)         # mate and non-mate as two gradient rows
        k_layer = int(k_layer) % len(P)
        if P[k_layer] is None:
            return self._zero_map(mwp)
        prior = self._contrastive_prior(gs, P[k_layer][0:1], P[k_layer][1:2], k_layer, mode, percentile, k_element)
        P0 = self._onehot(k_poschannel, dev)
        P, names, P2 = gs.run(0.0 * P0, W2, self._ebp_subtree_mode, priors={int(k_layer): (0, prior.reshape(-1).contiguous())}, record=True,
                                hooked_fc2=hk)
        self._set_P(gs, P, names)
        return self._finish_map(P2, mwp)[0]

    def weighted_subtree_ebp(self, img_probe, k_poschannel, k_negchannel, topk=1, verbose=True, do_max_subtree=False,
                             do_mated_similarity_gating=True, subtree_mode='norelu', do_mwp_to_saliency=True, rows_per_sweep=48):
        """whitebox.py:647-737.  One forward; the three true-gradient passes (cross-entropy, mate logit, non-mate logit)
        are one 3-row sweep without hooks; the per-layer scores come from the xfrb_subtree_score kernel; the ~n one-element
        sub-tree EBPs are batched as gradient rows (`rows_per_sweep` at a time) instead of 2n separate ebp() calls."""
        self._ebp_subtree_mode = subtree_mode                                       # the reference's side effect (651)
        gs, W2 = self._generic(img_probe)
        hk = self.net.hooked()            # no triplet classifier: the network's own fc2 fires (and is indexed) first
        eng, be, dev = gs.eng, gs.be, W2.device
        # true gradients dA of: cross-entropy(y, 0), y[0], y[1]
        y = self._logits(eng, W2)                                                        # classify(): logits of the triplet head
        sm = torch.softmax(y, dim=1)
        Pn = torch.zeros(3, self.net.num_classes(), device=dev)
        Pn[0] = sm[0]
        Pn[0, 0] -= 1.0
        Pn[1, 0] = 1.0
        Pn[2, 1] = 1.0
        # every sweep goes through the engine's graph table (captured once, replayed): the 3-row true-gradient sweep with the
        # per-firing score / arg-max kernels (not including the image layer, 684), then P_mate with a probe at every arg-max
        # node instead of a recording of all of P, then the one-element priors as batched gradient rows - priors and probes
        # travel in device tables (generic.PriorTable), the only thing that changes between replays
        tg = eng.generic_call(Pn, W2, mode=subtree_mode, record=True, true_grad=True, hooked_fc2=hk,
                              gating=bool(do_mated_similarity_gating))
        names = tg['names']
        P_subtree = [float(v) for v in tg['score'].cpu().numpy()]
        P_subtree_idx = tg['arg'].cpu().numpy()
        n_layers = len(P_subtree) + 1
        del tg
        k_subtree = np.argsort(np.array(P_subtree))                                 # ascending, one per layer (697)
        # layerwise EBP for every sub-tree: P_mate once, then one-element priors as batched gradient rows
        P0 = self._onehot(k_poschannel, dev)
        probe = eng.prior_table('probe')
        probe.clear()
        for k in range(n_layers - 1):
            probe.set_probe(k, 0, int(P_subtree_idx[k]))
        probe.upload()
        eng.generic_call(P0, W2, mode=subtree_mode, hooked_fc2=hk, ptab=probe)
        p_at = probe.probe[:n_layers - 1].cpu().numpy()
        seeds = [(int(k), int(P_subtree_idx[k]), float(p_at[k])) for k in k_subtree]
        P_img = []
        tab = eng.prior_table('rows')
        rows_per_sweep = max(1, min(int(rows_per_sweep), _MAX_ROWS))                             # a prior table describes up to 64 rows
        Z = torch.zeros(min(rows_per_sweep, len(seeds)), self.net.num_classes(), device=dev)     # fixed row count: one graph
        for i in range(0, len(seeds), Z.shape[0]):
            chunk = seeds[i:i + Z.shape[0]]
            tab.clear(zero_seed=True)                                     # zero class priors: row r is all zero before its firing
            for r, (k, e, v) in enumerate(chunk):
                tab.set_elem(k, r, e, v)
            tab.upload()
            P2 = eng.generic_call(Z, W2, mode=subtree_mode, hooked_fc2=hk, ptab=tab)['P2']
            P_img += list(P2[:len(chunk)].sum(-1).cpu().numpy().astype(np.float32))             # layerwise_ebp(..., mwp=True) maps
        self.P_layername = list(names)
        if verbose:
            for k in k_subtree:
                print('[weighted_subtree_ebp][%d]: layername=%s, grad=%f' % (k, names[k], P_subtree[k]))
        # merge (whitebox.py:705-737)
        k_valid = [np.max(P) > 0 for P in P_img]
        k_subtree_valid = [k for (k, v) in zip(k_subtree, k_valid) if v == True and k != 1][-topk:]     # noqa: E712
        if len(k_subtree_valid) == 0:
            raise RuntimeError(
                'Failed to calculate valid subtrees. The ebp subtree mode '
                '(%s) may not support by this type of network. You may want '
                'to try the "affineonly_with_prior" ebp subtree mode.' %
                self._ebp_subtree_mode
            )
        P_img_valid = [p for (p, k, v) in zip(P_img, k_subtree, k_valid) if v == True and k != 1][-topk:]   # noqa: E712
        P_subtree_valid = [P_subtree[k] for k in k_subtree_valid]
        norm = self._scale_normalized(P_subtree_valid)
        P_subtree_valid_norm = norm if not np.sum(norm) == 0 else np.ones_like(P_subtree_valid)
        stack = np.dstack([float(w) * np.array(P) * (1.0 / (np.max(P) + 1E-12)) for (w, P) in zip(P_subtree_valid_norm, P_img_valid)])
        smap = np.max(stack, axis=2) if do_max_subtree else np.sum(stack, axis=2)
        if self.convert_saliency_uint8:
            smap = self._float32_to_uint8(smap)
        else:
            smap /= max(smap.sum(), self.eps)
        if do_mwp_to_saliency:                      # the merged map and the topk sub-tree maps through ONE post-filter call
            sal = self._mwp_to_saliency_many([smap] + list(P_img_valid))
            smap, P_img_valid = sal[0], sal[1:]
        return (smap, P_img_valid, P_subtree_valid, k_subtree_valid)

    def _scale_normalized(self, img):
        """whitebox.py:443-446"""
        img = np.float32(img)
        return (img - np.min(img)) / (self.eps + (np.max(img) - np.min(img)))

    def _mwp_to_saliency(self, P, blur_radius=2):
        """whitebox.py:448-460 on one host map (the batched operators use the xfrb_saliency_post kernel instead)."""
        if self.convert_saliency_uint8:
            return self._mwp_to_saliency_uint8(P, blur_radius)
        t = torch.from_numpy(np.ascontiguousarray(P, dtype=np.float32)).unsqueeze(0)
        be = self.net.engine(self._ebp_with_bias).be
        dev = self.net._device()
        out = torch.empty(t.shape, device=dev)
        be.saliency_post(t.to(dev), out)
        return out[0].cpu().numpy()

    def _mwp_to_saliency_many(self, maps, blur_radius=2):
        """_mwp_to_saliency over a list of equally sized host maps: one copy in, one kernel, one copy out."""
        if self.convert_saliency_uint8 or len(maps) == 0:
            return [self._mwp_to_saliency(P, blur_radius) for P in maps]
        t = torch.from_numpy(np.ascontiguousarray(np.stack(maps), dtype=np.float32))
        dev = self.net._device()
        out = torch.empty(t.shape, device=dev)
        self.net.engine(self._ebp_with_bias).be.saliency_post(t.to(dev), out)
        return list(out.cpu().numpy())

    def ebp_subtree_mode(self):
        return self._ebp_subtree_mode

    def encode(self, x):
        return self.net.encode(x)

    def convert_from_numpy(self, img):
        """whitebox.py:787-806: float RGB image (HxWx3) in [0,1] or uint8 image -> network input tensor via net.preprocess.
        The reference resamples to 224x224 with skimage.transform.resize (absent here, and the identity for the 224x224
        images of the inpainting-game flow); other sizes go through PIL's bilinear resize."""
        import PIL.Image
        img = np.asarray(img)
        if img.dtype == np.uint8:
            img = img.astype(np.float32) / 255
        if img.max() > 1 + 1e-6 and img.min() > 0 - 1e-6:
            img = img / 255
        if img.max() > 1 + 1e-6 or img.min() < 0 - 1e-6:
            raise ValueError('convert_from_numpy: image range outside [0, 1] / [0, 255] (the reference drops into pdb here)')
        im8 = (img * 255).astype(np.uint8)
        pil = PIL.Image.fromarray(im8).convert('RGB')
        if pil.size != (224, 224):
            pil = pil.resize((224, 224), PIL.Image.BILINEAR)
        return self.net.preprocess(pil)

    def preprocess_loader(self, images, returnImageIndex=False, repeats=1):
        """whitebox.py:808-824 for in-memory HxWx3 arrays: yields (displayable image, [C,H,W] tensor, file name = None).
        (File lists / DataFrames go through xfr.utils.image_loader in the reference - host glue outside this package.)"""
        for i, im in enumerate(images):
            if not isinstance(im, np.ndarray):
                raise NotImplementedError('preprocess_loader: only in-memory numpy images are handled here')
            assert im.ndim == 3 and im.shape[2] == 3
            imT = self.convert_from_numpy(im)
            base = (im, imT[0]) + ((i,) if returnImageIndex else ()) + (None,)
            if repeats == 1:
                yield base
            else:
                for r in range(repeats):
                    yield base + (r,)

    def embeddings(self, images, norm=True):
        """whitebox.py:747-785: tensors / arrays already in network format ([C,H,W]), or HxWx3 numpy images."""
        if isinstance(images[0], torch.Tensor):
            imagesT = torch.stack(list(images)) if not isinstance(images, torch.Tensor) else images
        elif isinstance(images[0], np.ndarray) and images[0].ndim == 3 and images[0].shape[0] not in (1, 3):
            imagesT = torch.stack([self.convert_from_numpy(im)[0] for im in images])
        else:
            imagesT = torch.stack([torch.from_numpy(np.asarray(im)).float() for im in images])
        embeds = self.encode(imagesT).detach().cpu().numpy()
        if norm:
            embeds = embeds / np.linalg.norm(embeds.reshape(embeds.shape[0], -1), axis=1, keepdims=True)
        return embeds
