"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Hook-free CPU restatement (torch fp32 on CPU, no autograd, no register_hook) of
the reference's excitation-backprop path for the STR-Janus ResNet:

    Whitebox.ebp                        reference whitebox.py:482-504
    Whitebox._forward_hook/_backward_ebp reference whitebox.py:351-437
    Whitebox._preforward_hook            reference whitebox.py:306-349
    contrastive / truncated contrastive  reference whitebox.py:506-558
    layerwise_ebp                        reference whitebox.py:561-581
    weighted_subtree_ebp                 reference whitebox.py:647-737
    _mwp_to_saliency                     reference whitebox.py:448-460
    network topology                     reference resnet.py:104-265

Parity status: PINNED by outputs of the reference itself.  `oracle/gen_golden.py`
imports the unmodified reference from /root/reference/python (with the import
shim in oracle/shim), runs it on seeded inputs and commits its outputs to
tests/golden/; tests/test_oracle_golden.py checks this file against them.
The reference ships no tests / golden vectors of its own.

All conv / BN / pooling / linear arithmetic is torch's CPU fp32 (the reference's
own arithmetic dependency, un-vendored: reference README.md:18 "PyTorch 1.3";
here torch 2.11).  The Gaussian post-filter is scikit-image's (README.md:37),
restated with scipy.ndimage (see oracle/shim/skimage/filters.py).

Definitions.  For every hooked tensor t (an input of a
leaf-module call) the reference keeps A_t = relu(value of t in the true forward)
and X_t = relu(value of t in the "positive" forward, where every module ran with
relu(weight) on the A inputs, bias/beta/running stats unchanged unless
with_bias).  On the way back each hook sees the incoming gradient z and does
    zh = relu(z); p = A*zh; [p <- prior]; record p; return f(mode, kind, ...)
Hooks that share one tensor run in forward-registration order, each fed the
previous one's return value.  The Add module's two hooks both close over the
(A, X) of its LAST input (late-binding closure, whitebox.py:379-432).
"""
import numpy as np
import torch
import torch.nn.functional as F

AFFINE_KINDS = ('Conv', 'Linear', 'AvgPool', 'BatchNorm')  # whitebox.py:399,409 substring test
BN_EPS = 1e-5
LAYERS101 = (3, 4, 23, 3)


def _relu(t):
    return torch.clamp_min(t, 0)


class _Hooks(object):
    """State of one backward sweep: records P and applies priors in firing order."""

    def __init__(self, mode, eps, prior=None):
        self.mode = mode
        self.eps = eps
        self.prior = prior or {}
        self.P = []
        self.names = []

    def fire(self, kind, a, x, z):
        k = len(self.P)
        prior = self.prior.get(k)
        zh = _relu(z)
        p = a * zh
        if prior is not None:
            p = prior.reshape(p.shape).to(p.dtype).clone()
        self.P.append(p)
        self.names.append(kind)
        affine = any(s in kind for s in AFFINE_KINDS)
        mode = self.mode
        if mode == 'affineonly':
            return p / (x + self.eps) if affine else z
        if mode == 'affineonly_with_prior':
            if prior is not None:
                zh = (prior.reshape(p.shape) > 0) * z
                p = (prior.reshape(p.shape) > 0) * p
            return p / (x + self.eps) if affine else zh
        if mode == 'norelu':
            if prior is not None and ('MaxPool' in kind or 'ReLU' in kind):
                return z
            return p / (x + self.eps)
        if mode == 'all':
            return p / (x + self.eps)
        raise ValueError('Invalid subtree mode "%s"' % mode)


def block_names(layers=LAYERS101):
    return ['layer%d.%d' % (li, bi) for li, n in enumerate(layers, start=1) for bi in range(n)]


def _bn_params(sd, name, positive=False, with_bias=False):
    g = sd[name + '.weight']
    b = sd[name + '.bias']
    if positive:
        g = _relu(g)
        if with_bias:
            b = _relu(b)
    inv = 1.0 / torch.sqrt(sd[name + '.running_var'] + BN_EPS)
    return g, b, sd[name + '.running_mean'], inv


def _bn(sd, name, t, positive=False, with_bias=False):
    g, b, _, _ = _bn_params(sd, name, positive, with_bias)
    return F.batch_norm(t, sd[name + '.running_mean'], sd[name + '.running_var'], g, b, False, 0.0, BN_EPS)


def _bn_bwd_pos(sd, name, gr):
    g, _, _, inv = _bn_params(sd, name, positive=True)
    return gr * (g / torch.sqrt(sd[name + '.running_var'] + BN_EPS)).view(1, -1, 1, 1)


def _bn_bwd_true(sd, name, gr):
    return gr * (sd[name + '.weight'] / torch.sqrt(sd[name + '.running_var'] + BN_EPS)).view(1, -1, 1, 1)


def _conv(sd, name, t, stride, pad, positive=False, with_bias=False):
    w = sd[name + '.weight']
    b = sd.get(name + '.bias')
    if positive:
        w = _relu(w)
        if with_bias and b is not None:
            b = _relu(b)
    return F.conv2d(t, w, b, stride, pad)


def _conv_bwd(sd, name, gr, in_shape, stride, pad, positive=True):
    w = sd[name + '.weight']
    if positive:
        w = _relu(w)
    return torch.nn.grad.conv2d_input(in_shape, w, gr, stride, pad)


def forward(sd, x, layers=LAYERS101):
    """True forward (reference resnet.py:224-265), keeping every tensor the backward needs."""
    T = {}
    c1 = _conv(sd, 'conv1', x, 2, 3)
    r1 = _relu(_bn(sd, 'bn1', c1))
    mp, mp_idx = F.max_pool2d(r1, 3, 2, 1, return_indices=True)
    T.update(x=x, c1=c1, r1=r1, mp=mp, mp_idx=mp_idx)
    u = mp
    blocks = []
    inplanes = 64
    for li, (planes, n) in enumerate(zip((64, 128, 256, 512), layers), start=1):
        for bi in range(n):
            name = 'layer%d.%d' % (li, bi)
            stride = 2 if (bi == 0 and li > 1) else 1
            has_ds = bi == 0
            o1 = _conv(sd, name + '.conv1', u, stride, 0)
            a1 = _relu(_bn(sd, name + '.bn1', o1))
            o2 = _conv(sd, name + '.conv2', a1, 1, 1)
            a2 = _relu(_bn(sd, name + '.bn2', o2))
            o3 = _conv(sd, name + '.conv3', a2, 1, 0)
            n3 = _bn(sd, name + '.bn3', o3)
            if has_ds:
                ap = F.avg_pool2d(u, stride, stride)  # resnet.py:211 (kernel tied to stride)
                rep = planes * 4 // inplanes - 1      # resnet.py:212 ConcatChannels
                res = torch.cat((ap, torch.zeros_like(ap).repeat(1, rep, 1, 1)), dim=1)
            else:
                ap = None
                res = u
            out = _relu(n3 + res)
            blocks.append(dict(name=name, stride=stride, has_ds=has_ds, u=u, o1=o1, a1=a1, o2=o2, a2=a2,
                               o3=o3, n3=n3, ap=ap, res=res, out=out))
            u = out
            inplanes = planes * 4
    pool = F.avg_pool2d(u, 7, 7)
    v = pool.flatten(1)
    f1 = F.linear(v, sd['fc1.weight'], sd['fc1.bias'])
    nrm = f1.norm(dim=1, keepdim=True).clamp_min(1e-12)  # F.normalize eps
    xn = f1 / nrm
    T.update(blocks=blocks, pool=pool, v=v, f1=f1, nrm=nrm, xn=xn, enc=xn * 50.0)
    return T


def encode(sd, x, layers=LAYERS101):
    """WhiteboxSTResnet.encode (whitebox.py:98-100): 50 * L2-normalised fc1 output."""
    with torch.no_grad():
        return forward(sd, x, layers)['enc']


def ebp_mwp(sd, x, Pn, fc2=None, mode='affineonly_with_prior', prior=None, with_bias=False, eps=1e-16,
            layers=LAYERS101, T=None, stop_at_stem=False):
    """One excitation-backprop sweep.  Returns (P list, kind list).

    x    [N,3,224,224]; Pn [N,C] prior over classes.
    fc2  None -> the network's own hooked fc2 ('fc2.weight' in sd, W+ in the backward,
         one extra leading 'Linear' firing); a tensor [N,2,512] (or [2,512]) -> the
         *un-hooked* replacement classifier of set_triplet_classifier
         (whitebox.py:93-96): signed weights, no P entry.
    prior  {firing index: tensor} (layerwise modes, whitebox.py:570-577).
    """
    r = _relu
    with torch.no_grad():
        if T is None:
            T = forward(sd, x, layers)
        H = _Hooks(mode, eps, prior)
        N = x.shape[0]
        xn, v, nrm = T['xn'], T['v'], T['nrm']
        W1p = r(sd['fc1.weight'])
        if fc2 is not None:
            W2 = fc2 if fc2.dim() == 3 else fc2.unsqueeze(0).expand(N, -1, -1)
            gr = torch.einsum('nc,ncd->nd', Pn, W2)
        else:
            gr = Pn @ r(sd['fc2.weight'])
            gr = H.fire('Linear', r(xn * 50.0), r(50.0 * r(xn)), gr)
        gr = gr * 50.0                                                  # Multiply backward (resnet.py:160-165)
        b1 = sd['fc1.bias']
        f1p = F.linear(r(v), W1p, r(b1) if with_bias else b1)
        gr = H.fire('Multiply', r(xn), r(F.normalize(f1p, p=2, dim=1)), gr)
        gr = (gr - xn * (xn * gr).sum(1, keepdim=True)) / nrm           # true Jacobian of F.normalize (resnet.py:250)
        gr = gr @ W1p
        blocks = T['blocks']
        u_last = blocks[-1]['out']
        gr = H.fire('Linear', r(v), r(F.avg_pool2d(r(u_last), 7, 7).flatten(1)), gr)
        gr = gr.view(T['pool'].shape)
        gr = F.interpolate(gr, scale_factor=7, mode='nearest') / 49.0   # AvgPool2d(7) backward
        for i in range(len(blocks) - 1, -1, -1):
            S = blocks[i]
            name, out = S['name'], S['out']
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            # hooks chained on the in-place-ReLU'd block output, in forward registration order
            gr = H.fire('ReLU', out, r(r(S['n3']) + r(S['res'])), gr)
            if nxt is None:
                gr = H.fire('AvgPool2d', out, out, gr)
            else:
                gr = H.fire('Conv2d', out, out, gr)
                gr = H.fire('AvgPool2d' if nxt['has_ds'] else 'Add', out, out, gr)
            gr = gr * (out > 0)                                         # ReLU backward
            g_res = gr
            if S['has_ds']:
                g_res = H.fire('Add', r(S['res']), r(S['res']), g_res)  # slot 1 (residual) fires first
                C = S['ap'].shape[1]
                g_res = g_res[:, :C]
                g_res = H.fire('ConcatChannels', r(S['ap']), r(S['ap']), g_res)
                k = S['stride']
                if k > 1:
                    g_res = F.interpolate(g_res, scale_factor=k, mode='nearest') / (k * k)
            g = H.fire('Add', r(S['res']), r(S['res']), gr)             # slot 0: closure bug -> residual's (A, X)
            g = _bn_bwd_pos(sd, name + '.bn3', g)
            g = H.fire('BatchNorm2d', r(S['o3']),
                       r(_conv(sd, name + '.conv3', S['a2'], 1, 0, True, with_bias)), g)
            g = _conv_bwd(sd, name + '.conv3', g, S['a2'].shape, 1, 0)
            g = H.fire('ReLU', S['a2'], r(_bn(sd, name + '.bn2', r(S['o2']), True, with_bias)), g)
            g = H.fire('Conv2d', S['a2'], S['a2'], g)
            g = g * (S['a2'] > 0)
            g = _bn_bwd_pos(sd, name + '.bn2', g)
            g = H.fire('BatchNorm2d', r(S['o2']),
                       r(_conv(sd, name + '.conv2', S['a1'], 1, 1, True, with_bias)), g)
            g = _conv_bwd(sd, name + '.conv2', g, S['a1'].shape, 1, 1)
            g = H.fire('ReLU', S['a1'], r(_bn(sd, name + '.bn1', r(S['o1']), True, with_bias)), g)
            g = H.fire('Conv2d', S['a1'], S['a1'], g)
            g = g * (S['a1'] > 0)
            g = _bn_bwd_pos(sd, name + '.bn1', g)
            g = H.fire('BatchNorm2d', r(S['o1']),
                       r(_conv(sd, name + '.conv1', r(S['u']), S['stride'], 0, True, with_bias)), g)
            g = _conv_bwd(sd, name + '.conv1', g, S['u'].shape, S['stride'], 0)
            gr = g + g_res
        mp, r1, c1 = T['mp'], T['r1'], T['c1']
        gr = H.fire('Conv2d', r(mp), r(mp), gr)        # layer1.0.conv1
        gr = H.fire('AvgPool2d', r(mp), r(mp), gr)     # layer1.0.downsample[0] (kernel 1)
        gr = torch.zeros_like(r1).flatten(2).scatter_add_(2, T['mp_idx'].flatten(2), gr.flatten(2)).view_as(r1)
        gr = H.fire('ReLU', r1, r(_bn(sd, 'bn1', r(c1), True, with_bias)), gr)
        gr = H.fire('MaxPool2d', r1, r1, gr)
        gr = gr * (r1 > 0)
        gr = _bn_bwd_pos(sd, 'bn1', gr)
        gr = H.fire('BatchNorm2d', r(c1), r(_conv(sd, 'conv1', r(T['x']), 2, 3, True, with_bias)), gr)
        if not stop_at_stem:
            gr = _conv_bwd(sd, 'conv1', gr, T['x'].shape, 2, 3)
            gr = H.fire('Conv2d', r(T['x']), r(T['x']), gr)
        else:
            H.P.append(None)
            H.names.append('Conv2d')
    return H.P, H.names


# ---------------------------------------------------------------- post-processing

def gaussian_blur_sigma2(img):
    """skimage.filters.gaussian(img, 2) == scipy gaussian_filter(sigma=2, mode='nearest', truncate=4)."""
    import scipy.ndimage as ndi
    return ndi.gaussian_filter(np.asarray(img), 2, mode='nearest', truncate=4.0)


def mwp_to_saliency(P, eps=1e-16):
    """_mwp_to_saliency, ebp_version 6 branch (whitebox.py:455-460)."""
    img = gaussian_blur_sigma2(P)
    img = np.maximum(0, img)
    img /= max(img.sum(), eps)
    return img


def mwp_to_saliency_uint8(P, eps=1e-16, blur_radius=2):
    """_mwp_to_saliency, ebp_version != 6 branch (whitebox.py:451-454)."""
    import PIL.Image
    import PIL.ImageFilter
    img = np.uint8(255 * ((P - np.min(P)) / (eps + (np.max(P) - np.min(P)))))
    img = np.array(PIL.Image.fromarray(img).filter(PIL.ImageFilter.GaussianBlur(radius=blur_radius)))
    img = np.uint8(255 * ((img - np.min(img)) / (eps + (np.max(img) - np.min(img)))))
    return img


def ebp(sd, x, Pn, fc2=None, mwp=False, **kw):
    """Whitebox.ebp (whitebox.py:482-504) -> [N,112,112] float32."""
    P, _ = ebp_mwp(sd, x, Pn, fc2, stop_at_stem=True, **kw)
    m = P[-2].sum(1).cpu().numpy().astype(np.float32)
    if mwp:
        return m
    return np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])


def _onehot(n, c, k, device=None):
    p = torch.zeros(n, c, device=device)
    p[:, k] = 1.0
    return p


def contrastive_mwp(sd, x, fc2, k_pos=0, k_neg=1, percentile=None, num_classes=2, **kw):
    """contrastive_ebp / truncated_contrastive_ebp up to (not including) _mwp_to_saliency
    (whitebox.py:506-527, 529-558).  Normalisation and percentile mask are per sample."""
    N = x.shape[0]
    T = forward(sd, x, kw.get('layers', LAYERS101))
    Pm, _ = ebp_mwp(sd, x, _onehot(N, num_classes, k_pos, x.device), fc2, T=T, stop_at_stem=True, **kw)
    Pn, _ = ebp_mwp(sd, x, _onehot(N, num_classes, k_neg, x.device), fc2, T=T, stop_at_stem=True, **kw)
    pm, pn = Pm[-2], Pn[-2]
    mm = pm / pm.sum(dim=(1, 2, 3), keepdim=True)
    mn = pn / pn.sum(dim=(1, 2, 3), keepdim=True)
    if percentile is not None:
        out = []
        for i in range(N):
            flat = mm[i].flatten()
            srt, idx = torch.sort(flat.clone())
            cs = torch.cumsum(srt, 0)
            mask = torch.zeros_like(srt)
            mask[idx] = (cs >= (percentile / 100.0) * cs[-1]).float()
            mask = mask.view_as(mm[i])
            out.append(_relu(mask * mm[i] - mask * mn[i]).sum(0))
        return torch.stack(out).cpu().numpy().astype(np.float32)
    return _relu(mm - mn).sum(1).cpu().numpy().astype(np.float32)


def contrastive_ebp(sd, x, fc2, k_pos=0, k_neg=1, percentile=None, **kw):
    m = contrastive_mwp(sd, x, fc2, k_pos, k_neg, percentile, **kw)
    return np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])
