"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Hook-free CPU restatement (torch fp32 on CPU) of the reference's excitation backprop for the VGGFace2
ResNet-50-128d plugin:

    Whitebox_resnet50_128                reference whitebox.py:210-258 (un-hooked fc1 head on the wrapper, 216, 226-230)
    network                              reference models/resnet50_128_pytorch/resnet50_128.py:6-348
    hooks / ebp / contrastive            reference whitebox.py:306-437, 482-558 (shared with oracle/stresnet_oracle.py)

Parity status: PINNED by outputs of the reference itself run on the bundled real weights
(models/resnet50_128_pytorch/resnet50_128_pytorch.tar.gz) and the bundled VGGFace2 triplet / demo face
(oracle/gen_golden_r50.py -> tests/golden/resnet50_128_real.npz; SURVEY.md section 8c fingerprints).

Topology differences from the STR ResNet that matter to EBP:
  * convs have no bias; shortcuts of the first block of each stage are conv1x1(stride s) + BatchNorm ("proj");
  * the residual sum is the function torch.add, not a module: no Add hooks, and in the positive pass its
    operands are NOT overridden by A (so X of a block's ReLU is relu(BN+(relu(o3)) + X_res));
  * ReLUs are in-place modules: the ReLU hook and the consumers' Conv hooks chain on one tensor;
  * MaxPool2d(3, 2, padding 0, ceil_mode=True); AvgPool2d(7) then a 1x1 conv 2048->128 ("feat_extract").
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle.stresnet_oracle import _Hooks, _relu, BN_EPS, mwp_to_saliency, _onehot  # noqa: F401

STAGES = ((2, 3, 64), (3, 4, 128), (4, 6, 256), (5, 3, 512))   # (stage id, blocks, planes)


def block_list():
    out = []
    for s, n, planes in STAGES:
        for i in range(1, n + 1):
            out.append(dict(name='conv%d_%d' % (s, i), planes=planes, proj=(i == 1), stride=2 if (i == 1 and s > 2) else 1))
    return out


def _bn(sd, name, t, positive=False, with_bias=False):
    g, b = sd[name + '.weight'], sd[name + '.bias']
    if positive:
        g = _relu(g)
        if with_bias:
            b = _relu(b)
    return F.batch_norm(t, sd[name + '.running_mean'], sd[name + '.running_var'], g, b, False, 0.0, BN_EPS)


def _bn_bwd_pos(sd, name, gr):
    return gr * (_relu(sd[name + '.weight']) / torch.sqrt(sd[name + '.running_var'] + BN_EPS)).view(1, -1, 1, 1)


def _conv(sd, name, t, stride, pad, positive=False):
    w = sd[name + '.weight']
    return F.conv2d(t, _relu(w) if positive else w, None, stride, pad)


def _conv_bwd(sd, name, gr, in_shape, stride, pad):
    return torch.nn.grad.conv2d_input(in_shape, _relu(sd[name + '.weight']), gr, stride, pad)


def forward(sd, x):
    T = dict(x=x)
    c1 = _conv(sd, 'conv1_7x7_s2', x, 2, 3)
    r1 = _relu(_bn(sd, 'conv1_7x7_s2_bn', c1))
    mp, mp_idx = F.max_pool2d(r1, 3, 2, 0, ceil_mode=True, return_indices=True)
    T.update(c1=c1, r1=r1, mp=mp, mp_idx=mp_idx)
    u = mp
    blocks = []
    for B in block_list():
        n, s = B['name'], B['stride']
        o1 = _conv(sd, n + '_1x1_reduce', u, s, 0)
        a1 = _relu(_bn(sd, n + '_1x1_reduce_bn', o1))
        o2 = _conv(sd, n + '_3x3', a1, 1, 1)
        a2 = _relu(_bn(sd, n + '_3x3_bn', o2))
        o3 = _conv(sd, n + '_1x1_increase', a2, 1, 0)
        n3 = _bn(sd, n + '_1x1_increase_bn', o3)
        if B['proj']:
            op = _conv(sd, n + '_1x1_proj', u, s, 0)
            res = _bn(sd, n + '_1x1_proj_bn', op)
        else:
            op, res = None, u
        out = _relu(res + n3)
        blocks.append(dict(B, u=u, o1=o1, a1=a1, o2=o2, a2=a2, o3=o3, n3=n3, op=op, res=res, out=out))
        u = out
    pool = F.avg_pool2d(u, 7, 1)
    fe = _conv(sd, 'feat_extract', pool, 1, 0)
    T.update(blocks=blocks, pool=pool, enc=fe.flatten(1))
    return T


def encode(sd, x):
    """Whitebox_resnet50_128.encode (whitebox.py:222-224): the 128-d feat_extract output."""
    with torch.no_grad():
        return forward(sd, x)['enc']


def ebp_mwp(sd, x, Pn, fc1, mode='affineonly_with_prior', prior=None, with_bias=False, eps=1e-16, T=None,
            stop_at_stem=False):
    """fc1: [N,2,128] (or [2,128]) rows of the wrapper's un-hooked classifier (whitebox.py:216-220)."""
    r = _relu
    with torch.no_grad():
        if T is None:
            T = forward(sd, x)
        H = _Hooks(mode, eps, prior)
        N = x.shape[0]
        W = fc1 if fc1.dim() == 3 else fc1.unsqueeze(0).expand(N, -1, -1)
        gr = torch.einsum('nc,ncd->nd', Pn, W).view(N, -1, 1, 1)             # un-hooked fc1: signed weights, no P entry
        pool = T['pool']
        gr = _conv_bwd(sd, 'feat_extract', gr, pool.shape, 1, 0)
        blocks = T['blocks']
        u_last = blocks[-1]['out']
        gr = H.fire('Conv2d', r(pool), r(F.avg_pool2d(r(u_last), 7, 1)), gr)  # feat_extract input
        gr = gr.expand(-1, -1, 7, 7) / 49.0                                   # AvgPool2d(7) backward
        for i in range(len(blocks) - 1, -1, -1):
            S = blocks[i]
            n, s, out = S['name'], S['stride'], S['out']
            nxt = blocks[i + 1] if i + 1 < len(blocks) else None
            # X of the block ReLU: the add is a function, so it sums positive-pass values: BN+(relu(o3)) and the
            # positive-pass shortcut (identity: the block input itself; proj: BN+(relu(op)))
            xres = _bn(sd, n + '_1x1_proj_bn', r(S['op']), True, with_bias) if S['proj'] else S['u']
            xblk = r(_bn(sd, n + '_1x1_increase_bn', r(S['o3']), True, with_bias) + xres)
            gr = H.fire('ReLU', out, xblk, gr)
            if nxt is None:
                gr = H.fire('AvgPool2d', out, out, gr)
            else:
                gr = H.fire('Conv2d', out, out, gr)                 # next reduce
                if nxt['proj']:
                    gr = H.fire('Conv2d', out, out, gr)             # next proj
            gr = gr * (out > 0)
            g_skip = gr
            if S['proj']:
                # proj_bn was created after increase_bn: autograd runs it (and fires its hook) first
                gp = _bn_bwd_pos(sd, n + '_1x1_proj_bn', gr)
                gp = H.fire('BatchNorm2d', r(S['op']), r(_conv(sd, n + '_1x1_proj', r(S['u']), s, 0, True)), gp)
                g_skip = _conv_bwd(sd, n + '_1x1_proj', gp, S['u'].shape, s, 0)
            g = _bn_bwd_pos(sd, n + '_1x1_increase_bn', gr)
            g = H.fire('BatchNorm2d', r(S['o3']), r(_conv(sd, n + '_1x1_increase', S['a2'], 1, 0, True)), g)
            g = _conv_bwd(sd, n + '_1x1_increase', g, S['a2'].shape, 1, 0)
            g = H.fire('ReLU', S['a2'], r(_bn(sd, n + '_3x3_bn', r(S['o2']), True, with_bias)), g)
            g = H.fire('Conv2d', S['a2'], S['a2'], g)
            g = g * (S['a2'] > 0)
            g = _bn_bwd_pos(sd, n + '_3x3_bn', g)
            g = H.fire('BatchNorm2d', r(S['o2']), r(_conv(sd, n + '_3x3', S['a1'], 1, 1, True)), g)
            g = _conv_bwd(sd, n + '_3x3', g, S['a1'].shape, 1, 1)
            g = H.fire('ReLU', S['a1'], r(_bn(sd, n + '_1x1_reduce_bn', r(S['o1']), True, with_bias)), g)
            g = H.fire('Conv2d', S['a1'], S['a1'], g)
            g = g * (S['a1'] > 0)
            g = _bn_bwd_pos(sd, n + '_1x1_reduce_bn', g)
            g = H.fire('BatchNorm2d', r(S['o1']), r(_conv(sd, n + '_1x1_reduce', r(S['u']), s, 0, True)), g)
            g = _conv_bwd(sd, n + '_1x1_reduce', g, S['u'].shape, s, 0)
            gr = g + g_skip
        mp, r1, c1 = T['mp'], T['r1'], T['c1']
        gr = H.fire('Conv2d', r(mp), r(mp), gr)            # conv2_1 reduce
        gr = H.fire('Conv2d', r(mp), r(mp), gr)            # conv2_1 proj
        gr = torch.zeros_like(r1).flatten(2).scatter_add_(2, T['mp_idx'].flatten(2), gr.flatten(2)).view_as(r1)
        gr = H.fire('ReLU', r1, r(_bn(sd, 'conv1_7x7_s2_bn', r(c1), True, with_bias)), gr)
        gr = H.fire('MaxPool2d', r1, r1, gr)
        gr = gr * (r1 > 0)
        gr = _bn_bwd_pos(sd, 'conv1_7x7_s2_bn', gr)
        gr = H.fire('BatchNorm2d', r(c1), r(_conv(sd, 'conv1_7x7_s2', r(T['x']), 2, 3, True)), gr)
        if not stop_at_stem:
            gr = _conv_bwd(sd, 'conv1_7x7_s2', gr, T['x'].shape, 2, 3)
            gr = H.fire('Conv2d', r(T['x']), r(T['x']), gr)
        else:
            H.P.append(None)
            H.names.append('Conv2d')
    return H.P, H.names


def ebp(sd, x, Pn, fc1, mwp=False, **kw):
    P, _ = ebp_mwp(sd, x, Pn, fc1, stop_at_stem=True, **kw)
    m = P[-2].sum(1).numpy().astype(np.float32)
    return m if mwp else np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])


def contrastive_mwp(sd, x, fc1, k_pos=0, k_neg=1, percentile=None, **kw):
    N = x.shape[0]
    T = forward(sd, x)
    Pm, _ = ebp_mwp(sd, x, _onehot(N, 2, k_pos), fc1, T=T, stop_at_stem=True, **kw)
    Pn, _ = ebp_mwp(sd, x, _onehot(N, 2, k_neg), fc1, T=T, stop_at_stem=True, **kw)
    pm, pn = Pm[-2], Pn[-2]
    mm = pm / pm.sum(dim=(1, 2, 3), keepdim=True)
    mn = pn / pn.sum(dim=(1, 2, 3), keepdim=True)
    if percentile is None:
        return _relu(mm - mn).sum(1).numpy().astype(np.float32)
    out = []
    for i in range(N):
        srt, idx = torch.sort(mm[i].flatten().clone())
        cs = torch.cumsum(srt, 0)
        mask = torch.zeros_like(srt)
        mask[idx] = (cs >= (percentile / 100.0) * cs[-1]).float()
        mask = mask.view_as(mm[i])
        out.append(_relu(mask * mm[i] - mask * mn[i]).sum(0))
    return torch.stack(out).numpy().astype(np.float32)


def contrastive_ebp(sd, x, fc1, k_pos=0, k_neg=1, percentile=None, **kw):
    m = contrastive_mwp(sd, x, fc1, k_pos, k_neg, percentile, **kw)
    return np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])
