"""ORACLE — TEST INFRASTRUCTURE ONLY (never imported by the product path).

Hook-free CPU restatement (torch fp32 on CPU) of the reference's excitation backprop for the Light-CNN-29v2 plugin:

    WhiteboxLightCNN                     reference whitebox.py:113-159 (fc2 replaced by an un-hooked 2-row Linear, 120-123)
    network_29layers_v2, mfm, group,     reference lightcnn.py:216-275, 48-62, 64-73, 76-89
      resblock, Split, Add
    hooks / ebp / contrastive            reference whitebox.py:306-437, 482-558 (hook algebra shared with stresnet_oracle)

Parity status: PINNED by outputs of the reference itself (oracle/gen_golden_lightcnn.py runs the unmodified reference on
the seeded state_dict of xfr_b200/synth.py -> tests/golden/lightcnn29v2_seed0.npz).  The reference bundles no Light-CNN
weights and no tests.

What is specific to this net.  There are no ReLU / BatchNorm modules: activations are signed, so
A = relu(true value) and X = relu(positive-pass value) differ wherever an UN-HOOKED op (torch.max of the MFM, the `+` of
max-pool and avg-pool, the residual Add's inputs) sits between two hooked modules:
  * a direct MFM output m:            positive pass value = max(relu(c_a), relu(c_b)) = relu(m)        -> X = A
  * the pooled sum p = maxpool(m) + avgpool(m):   p+ = maxpool(relu(m)) + avgpool(relu(m))             -> X = p+, A = relu(p)
  * a resblock output y = out + res:  the Add module ran on its A inputs, y+ = relu(out) + relu(res)   -> X = y+, A = relu(y)
  * a Split input c = conv(u):        c+ = conv_{relu(W)}(relu(u)) + b                                  -> X = relu(c+), A = relu(c)
  * both hooks of an Add close over the (A, X) of its LAST input, the residual (whitebox.py:379-432).
Hooks chained on one tensor fire in forward-registration order: [MaxPool2d, AvgPool2d] on a pooled MFM output,
[Conv2d, Add(slot 1)] on the input of a resblock.  torch.max(a, b) backward sends the gradient to the larger branch and
half to each on exact ties.
"""
import numpy as np
import torch
import torch.nn.functional as F

from oracle.stresnet_oracle import _Hooks, _relu, mwp_to_saliency, _onehot  # noqa: F401

LAYERS = (1, 2, 3, 4)


def _mfm(sd, name, u, positive=False, with_bias=False):
    """mfm.forward (lightcnn.py:58-62) -> (c, m);  positive: relu(W) (and relu(b) iff with_bias) on the given input."""
    w, b = sd[name + '.filter.weight'], sd[name + '.filter.bias']
    if positive:
        w = _relu(w)
        if with_bias:
            b = _relu(b)
    c = F.conv2d(u, w, b, 1, w.shape[-1] // 2)
    h = c.shape[1] // 2
    return c, torch.max(c[:, :h], c[:, h:])


def _pool(m):
    """maxpool2(m) + avgpool2(m)  (lightcnn.py:252)"""
    return F.max_pool2d(m, 2) + F.avg_pool2d(m, 2)


def forward(sd, x, layers=LAYERS):
    """True forward of network_29layers_v2 (lightcnn.py:249-275).  Records the mfm sites in forward order."""
    T = {'x': x, 'sites': []}

    def mfm(name, u):
        c, m = _mfm(sd, name, u)
        T['sites'].append(dict(name=name, u=u, c=c, m=m))
        return m
    t = mfm('conv1', x)
    stages = []
    chans = (48, 96, 192, 128)
    for bi, n in enumerate(layers, start=1):
        st = dict(blocks=[], pooled=None)
        if bi in (1, 2, 3):
            st['pool_in'] = t
            t = _pool(t)
            st['pooled'] = t
        for i in range(n):
            res = t
            ma = mfm('block%d.%d.conv1' % (bi, i), t)
            mb = mfm('block%d.%d.conv2' % (bi, i), ma)
            t = mb + res
            st['blocks'].append(dict(res=res, out=mb, y=t))
        mga = mfm('group%d.conv_a' % bi, t)
        t = mfm('group%d.conv' % bi, mga)
        st['gout'] = t
        stages.append(st)
    T['pool4_in'] = t
    p4 = _pool(t)
    v = p4.flatten(1)
    fc = F.linear(v, sd['fc.weight'], sd['fc.bias'])
    T.update(stages=stages, p4=p4, v=v, fc=fc)
    return T


def encode(sd, x, layers=LAYERS):
    """WhiteboxLightCNN.encode (whitebox.py:125-128): the 256-d fc output."""
    with torch.no_grad():
        return forward(sd, x, layers)['fc']


def _mfm_bwd(c, g):
    """autograd of torch.max(a, b) followed by the cat of Split's backward."""
    h = c.shape[1] // 2
    a, b = c[:, :h], c[:, h:]
    ga = torch.where(a == b, g / 2, g).masked_fill(a < b, 0)
    gb = torch.where(a == b, g / 2, g).masked_fill(b < a, 0)
    return torch.cat((ga, gb), 1)


def _pool_bwd(m, g):
    _, idx = F.max_pool2d(m, 2, return_indices=True)
    gm = torch.zeros_like(m).flatten(2).scatter_add_(2, idx.flatten(2), g.flatten(2)).view_as(m)
    return gm + F.interpolate(g, scale_factor=2, mode='nearest') / 4.0


def ebp_mwp(sd, x, Pn, fc2=None, mode='affineonly_with_prior', prior=None, with_bias=False, eps=1e-16, layers=LAYERS,
            T=None, stop_at_stem=False):
    """One excitation-backprop sweep -> (P list, kind list).
    fc2 None: the network's own hooked fc2 (one extra leading 'Linear' firing); a tensor [N,2,256] / [2,256]: the
    un-hooked replacement of set_triplet_classifier (whitebox.py:120-123): signed weights, no P entry."""
    r = _relu
    with torch.no_grad():
        if T is None:
            T = forward(sd, x, layers)
        H = _Hooks(mode, eps, prior)
        N = x.shape[0]
        sites = {s['name']: s for s in T['sites']}
        Wfc_p = r(sd['fc.weight'])
        bfc = sd['fc.bias']
        p4_in = T['pool4_in']
        p4_pos = _pool(r(p4_in))
        if fc2 is not None:
            W2 = fc2 if fc2.dim() == 3 else fc2.unsqueeze(0).expand(N, -1, -1)
            gr = torch.einsum('nc,ncd->nd', Pn, W2)
        else:
            gr = Pn @ r(sd['fc2.weight'])
            fc_pos = F.linear(r(T['v']), Wfc_p, r(bfc) if with_bias else bfc)
            gr = H.fire('Linear', r(T['fc']), r(fc_pos), gr)            # fc2 input = eval-mode dropout(fc) = fc itself
        gr = gr @ Wfc_p
        gr = H.fire('Linear', r(T['v']), p4_pos.flatten(1), gr)
        gr = gr.view_as(T['p4'])

        def through_mfm(name, g_m, last=False):
            """gradient at an mfm output (hooks on it already applied) -> gradient at the mfm input (before its hooks)."""
            s = sites[name]
            gc = _mfm_bwd(s['c'], g_m)
            cpos, _ = _mfm(sd, name, r(s['u']), True, with_bias)
            gc = H.fire('Split', r(s['c']), r(cpos), gc)
            if last and stop_at_stem:
                return None
            w = r(sd[name + '.filter.weight'])
            return torch.nn.grad.conv2d_input(s['u'].shape, w, gc, 1, w.shape[-1] // 2)

        def pooled_site(m, g_p):
            g = _pool_bwd(m, g_p)
            g = H.fire('MaxPool2d', r(m), r(m), g)
            return H.fire('AvgPool2d', r(m), r(m), g)

        gr = pooled_site(p4_in, gr)
        stages = T['stages']
        for bi in range(len(stages), 0, -1):
            st = stages[bi - 1]
            gr = through_mfm('group%d.conv' % bi, gr)
            mga = sites['group%d.conv_a' % bi]['m']
            gr = H.fire('Conv2d', r(mga), r(mga), gr)
            gr = through_mfm('group%d.conv_a' % bi, gr)
            blocks = st['blocks']
            for i in range(len(blocks) - 1, -1, -1):
                B = blocks[i]
                y, out, res = B['y'], B['out'], B['res']
                ypos = r(out) + r(res)
                gr = H.fire('Conv2d', r(y), ypos, gr)                 # next consumer's filter (conv_a or next conv1)
                if i + 1 < len(blocks):
                    gr = H.fire('Add', r(y), ypos, gr)                # next resblock's Add, slot 1
                g_res = gr
                # (A, X) of the residual input of THIS block's Add
                if i > 0:
                    pb = blocks[i - 1]
                    xres = r(pb['out']) + r(pb['res'])
                elif st['pooled'] is not None:
                    xres = _pool(r(st['pool_in']))
                else:
                    xres = r(res)                                       # block4.0: the residual is group3's MFM output
                gr = H.fire('Add', r(res), xres, gr)                  # slot 0 on `out`, residual's (A, X)
                gr = through_mfm('block%d.%d.conv2' % (bi, i), gr)
                ma = sites['block%d.%d.conv1' % (bi, i)]['m']
                gr = H.fire('Conv2d', r(ma), r(ma), gr)
                gr = through_mfm('block%d.%d.conv1' % (bi, i), gr)
                gr = gr + g_res
                if i == 0:
                    # hooks on the stage input: conv1.filter then Add slot 1
                    gr = H.fire('Conv2d', r(res), xres, gr)
                    gr = H.fire('Add', r(res), xres, gr)
            if st['pooled'] is not None:
                gr = pooled_site(st['pool_in'], gr)
        # block4's input is group3's output: its [Conv2d, Add] hooks were fired above with x = a; nothing pools there
        gr = through_mfm('conv1', gr, last=True)
        if not stop_at_stem:
            H.fire('Conv2d', r(T['x']), r(T['x']), gr)
        else:
            H.P.append(None)
            H.names.append('Conv2d')
    return H.P, H.names


def ebp(sd, x, Pn, fc2=None, mwp=False, **kw):
    """Whitebox.ebp (whitebox.py:482-504) -> [N,128,128] float32."""
    P, _ = ebp_mwp(sd, x, Pn, fc2, stop_at_stem=True, **kw)
    m = P[-2].sum(1).numpy().astype(np.float32)
    if mwp:
        return m
    return np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])


def contrastive_mwp(sd, x, fc2, k_pos=0, k_neg=1, percentile=None, num_classes=2, **kw):
    """contrastive_ebp / truncated_contrastive_ebp before _mwp_to_saliency (whitebox.py:506-558)."""
    N = x.shape[0]
    T = forward(sd, x, kw.get('layers', LAYERS))
    Pm, _ = ebp_mwp(sd, x, _onehot(N, num_classes, k_pos), fc2, T=T, stop_at_stem=True, **kw)
    Pn, _ = ebp_mwp(sd, x, _onehot(N, num_classes, k_neg), fc2, T=T, stop_at_stem=True, **kw)
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
        return torch.stack(out).numpy().astype(np.float32)
    return _relu(mm - mn).sum(1).numpy().astype(np.float32)


def contrastive_ebp(sd, x, fc2, k_pos=0, k_neg=1, percentile=None, **kw):
    m = contrastive_mwp(sd, x, fc2, k_pos, k_neg, percentile, **kw)
    return np.stack([mwp_to_saliency(mi, kw.get('eps', 1e-16)) for mi in m])
