"""TEST INFRASTRUCTURE: torch (CPU) emulation of the C-ABI kernel set of libxfr_b200.so.

Every method has the name, argument order and semantics of the kernel wrapper of
the same name in xfr_b200/kernels.py (include/xfrb.h documents each).  It exists so
that (1) the host-side schedule in xfr_b200/engine.py can be checked against the
oracle on a machine without a GPU, and (2) each CUDA kernel can be checked on the
GPU against a readable statement of what it must compute.  The product never
imports this module; xfr_b200.kernels raises if the CUDA library is missing.
"""
import numpy as np
import torch
import torch.nn.functional as F

from xfr_b200.packing import from_pair, to_pair, unpack_dual_cols

MODE_AWP, MODE_ALL, MODE_AFFINEONLY = 0, 1, 2
EPS = 1e-16


def relu(t):
    return torch.clamp_min(t, 0)


def hook(affine, a, x, z, mode, eps=EPS):
    """One _backward_ebp firing without a prior (reference whitebox.py:381-430)."""
    zh = relu(z)
    p = a * zh
    if mode == MODE_AWP:
        return p / (x + eps) if affine else zh
    if mode == MODE_ALL:
        return p / (x + eps)
    if mode == MODE_AFFINEONLY:
        return p / (x + eps) if affine else z
    raise ValueError(mode)


def _rows(t, J):
    """Saved tensors hold N samples; gradient tensors hold J = G*N rows (sample = j % N)."""
    n = t.shape[0]
    if n == J:
        return t
    assert J % n == 0
    return t.repeat((J // n,) + (1,) * (t.dim() - 1))


def im2col_nhwc(x, R, S, pad):
    """[N,H,W,C] -> [N*H*W, R*S*C], K ordered (r, s, c); stride 1."""
    N, H, W, C = x.shape
    if R == 1 and S == 1:
        return x.reshape(N * H * W, C)
    xp = F.pad(x, (0, 0, pad, pad, pad, pad))
    cols = [xp[:, r:r + H, s:s + W, :] for r in range(R) for s in range(S)]
    return torch.stack(cols, dim=3).reshape(N * H * W, R * S * C)


def _w(B):
    """Weight operand -> plain matrix: the 3xTF32 pack is two planes (hi, lo) with hi + lo == W exactly."""
    B = B.float()
    return B.sum(0) if B.dim() == 3 else B


def _wpos(B):
    """W+ operand as the split-TF32 plan of impl 'tf32x3' multiplies it (conv_tc.cu, SPLIT 2 / 3): activations exact,
    relu(W) rounded to TF32 = the hi plane alone.  Single-plane packs (fp32 / tf32) are used as they are."""
    B = B.float()
    return B[0] if B.dim() == 3 else B


class EmulBackend(object):
    name = 'emul'

    def __init__(self, eps=EPS, impl_name='fp32', bwd_single_pass=False, fwd_two_pass=False, pairs=False):
        self.eps = eps
        self.impl_name = impl_name      # which weight packing the engine should build (the arithmetic here is fp32)
        # kernels.HYBRID_IMPLS['bf16x2']: GEMM operands of the fused sweep are pair tensors (bf16 hi | lo rows, 16 significant
        # bits), relu(W) one bf16 term, signed W two; fp32 accumulation.  Pair tensors are bit-exact images of the device format.
        self.pairs = pairs
        self.conv_pack = 'bf16x2' if pairs else impl_name
        # kernels.HYBRID_IMPLS['tf32x3b1']: the W+ dgrads as ONE TF32 pass - the tensor core truncates the fp32 activation
        # operand to TF32 (tools/trunc_probe.py) and multiplies the hi weight plane
        self.bwd_single_pass = bwd_single_pass
        # kernels.HYBRID_IMPLS['tf32x2f']: the forward dual convs multiply the hi plane of the signed weights too (two passes)
        self.fwd_two_pass = fwd_two_pass

    # ------------------------------------------------------------ forward
    def stem_fwd(self, x, stem, o, mp, mp_arg=None):
        """x [N,224,224,3]; o = conv7x7/2(x)+b [N,112,112,64]; mp = maxpool3x3/2(relu(bn(o)))."""
        w = stem.W.view(7, 7, 3, stem.cout).permute(3, 2, 0, 1)
        oo = F.conv2d(x.permute(0, 3, 1, 2), w, stem.b, 2, 3)
        r1 = relu(oo * stem.bn[0].view(1, -1, 1, 1) + stem.bn[1].view(1, -1, 1, 1))
        o.copy_(oo.permute(0, 2, 3, 1))
        pad = stem.pool_pad
        mp.copy_(F.max_pool2d(r1, 3, 2, pad, ceil_mode=(pad == 0)).permute(0, 2, 3, 1))

    def subsample2(self, u, out):
        out.copy_(u[:, ::2, ::2, :])

    def avgpool2(self, u, out):
        out.copy_(F.avg_pool2d(u.permute(0, 3, 1, 2), 2, 2).permute(0, 2, 3, 1))

    def cubic_zoom(self, maps, out, normalize=True):
        """show.processSaliency's normalisation + the cubic resize as scikit-image >= 0.19 evaluates it (scipy.ndimage.zoom)"""
        import scipy.ndimage
        for i, m in enumerate(maps.numpy()):
            lo, hi = m.min(), m.max()
            if normalize:
                m = (m - lo) / np.float32(np.float32(hi - lo) + np.float32(1e-9))
            z = scipy.ndimage.zoom(m, (out.shape[1] / m.shape[0], out.shape[2] / m.shape[1]), order=3, mode='grid-constant', cval=0.0, grid_mode=True)
            out[i].copy_(torch.from_numpy(np.clip(z, m.min(), m.max()).astype(np.float32)))

    def to_pair(self, x, out, inverse=False):
        out.copy_(from_pair(x) if inverse else to_pair(x))

    def conv_dual(self, inp, L, o, xr, act, res=None, relu_act=True, act_f32=None):
        """o = conv(inp)+b ; xr = relu(conv_{W+}(inp)+b') ; act = relu(o*alpha+beta [+ res, zero-padded channels])."""
        if self.pairs:
            inp = from_pair(inp)
        A = im2col_nhwc(inp, L.R, L.S, L.R // 2)
        t, _ = unpack_dual_cols(A @ (_wpos(L.Bf) if self.fwd_two_pass else _w(L.Bf)).t() + L.bias, L.tn)
        _, p = unpack_dual_cols(A @ _wpos(L.Bf).t() + L.bias, L.tn)
        o.view(-1, L.cout).copy_(t)
        xr.view(-1, L.cout).copy_(relu(p))
        a = t * L.bn[0] + L.bn[1]
        if res is not None:
            rc = res.shape[-1]
            a[:, :rc] += res.reshape(-1, rc)
        a = relu(a) if relu_act else a
        if self.pairs:
            if act is not None:
                act.view(-1, L.cout).copy_(to_pair(a))
            if act_f32 is not None:
                act_f32.view(-1, L.cout).copy_(a)
        else:
            act.view(-1, L.cout).copy_(a)

    def head_fwd(self, u, head, v, f1, f1p, xn, nrm, xmul=None):
        """v = avgpool7(u); (f1 | f1p) = one dual GEMM v @ [W1 ; relu(W1)]^T + (b | b'); xn = f1/|f1|."""
        vv = F.avg_pool2d(u.permute(0, 3, 1, 2), 7, 7).flatten(1)
        ff, fp = unpack_dual_cols(vv @ _w(head.B1).t() + head.bias1, head.tn)
        nn_ = ff.norm(dim=1).clamp_min(1e-12)
        v.copy_(vv)
        f1.copy_(ff)
        f1p.copy_(fp)
        if xmul is not None:
            xmul.copy_(relu(F.normalize(fp, p=2, dim=1)))
        nrm.copy_(nn_)
        xn.copy_(ff / nn_.unsqueeze(1))

    # ------------------------------------------------------------ backward
    def head_bwd(self, Pn, W2, head, v, f1p, xn, nrm, mode, g_out, hooked_fc2=False):
        """Pn [J,C]; W2 [N,C,512] per-sample un-hooked classifier rows (signed) or, when
        hooked_fc2, the network's own [C,512] fc2 (W+ and a Linear hook).
        g_out [J,7,7,2048] = gradient w.r.t. the last block output (after AvgPool backward)."""
        J = Pn.shape[0]
        xn_, v_, nrm_ = _rows(xn, J), _rows(v, J), _rows(nrm, J)
        if W2 is None:
            gr = Pn                       # already the gradient at the fc2 input (hooked fc2 head)
        else:
            gr = torch.einsum('jc,jcd->jd', Pn, _rows(W2, J))
        gr = gr * head.scale
        Xmul = relu(F.normalize(_rows(f1p, J), p=2, dim=1))
        gr = hook(False, relu(xn_), Xmul, gr, mode, self.eps)
        gr = (gr - xn_ * (xn_ * gr).sum(1, keepdim=True)) / nrm_.unsqueeze(1)
        gr = gr @ _wpos(head.W1pT).t()
        gr = hook(True, relu(v_), relu(v_), gr, mode, self.eps)     # X = relu(avgpool(relu(u))) = v since u >= 0
        g_out.copy_((gr / 49.0).view(J, 1, 1, -1).expand(-1, 7, 7, -1))

    def _mid_chain(self, z, o, xr, bn, mode):
        alpha, beta, sp, tp = bn[0], bn[1], bn[2], bn[3]
        a = relu(o * alpha + beta)
        xrelu = relu(relu(o) * sp + tp)
        z = hook(False, a, xrelu, z, mode, self.eps)     # ReLU module hook
        z = hook(True, a, a, z, mode, self.eps)          # consumer Conv2d hook
        z = z * (a > 0)                                  # ReLU backward
        z = z * sp                                       # BatchNorm backward with gamma+
        return hook(True, relu(o), xr, z, mode, self.eps)  # BatchNorm hook

    def _dgrad(self, y, Bd, R, signed=False):
        """y [J,H,W,Cout] -> [J*H*W, Cin]"""
        if self.bwd_single_pass and not signed:
            y = (y.contiguous().view(torch.int32) & ~0x1FFF).view(torch.float32)
        return im2col_nhwc(y, R, R, R // 2) @ (_w(Bd) if signed else _wpos(Bd)).t()

    def dgrad_mid(self, y, L, o, xr, bn, mode, y_out):
        """z = W+^T y (conv L), then the hook chain at the activation a = relu(bn(o)) that fed conv L."""
        J = y.shape[0]
        z = self._dgrad(from_pair(y) if self.pairs else y, L.Bd, L.R)
        r = self._mid_chain(z, _rows(o, J).reshape(-1, L.cin), _rows(xr, J).reshape(-1, L.cin), bn, mode)
        y_out.view(-1, L.cin).copy_(to_pair(r) if self.pairs else r)

    def dgrad_plain(self, y, L, z_out, signed=False, accumulate=False, pair=False):
        if pair:
            y, B = from_pair(y), L.Bd
        else:
            B = L.signed_dgrad() if signed else (L.Bd32() if getattr(L, 'pair_pack', False) else L.Bd)
        z = self._dgrad(y, B, L.R, signed)
        if accumulate:
            z = z + z_out.reshape(-1, L.cin)
        z_out.view(-1, L.cin).copy_(z)

    def _join_chain(self, z, out, o3, xr3, bn3, res, hooks, mode):
        """Hook chain on a block output `out` (hooks: 1 = [affine], 2 = [affine, non-affine Add],
        3 = [affine, affine]) followed by the start of that block's main path.
        res = the block's residual input zero-padded to C channels (only read in MODE_ALL)."""
        alpha, beta, sp, tp = bn3[0], bn3[1], bn3[2], bn3[3]
        fn_add, chain = bool(hooks & 4), hooks & 3      # hooks & 4: torch.add instead of an Add module (ResNet-50-128d)
        if mode != MODE_ALL:      # non-affine hooks ignore (a, x) in these modes
            xblk = out
            rres = out
        elif fn_add:
            rres = None
            xblk = relu((relu(o3) * sp + tp) + res)
        else:
            rres = relu(res)
            xblk = relu(relu(o3 * alpha + beta) + rres)
        z = hook(False, out, xblk, z, mode, self.eps)
        z = hook(True, out, out, z, mode, self.eps)
        if chain == 2:
            z = hook(False, out, out, z, mode, self.eps)
        elif chain == 3:
            z = hook(True, out, out, z, mode, self.eps)
        g = z * (out > 0)
        zz = g if fn_add else hook(False, rres, rres, g, mode, self.eps)      # Add slot 0 with the residual's (A, X)
        zz = zz * sp
        y3 = hook(True, relu(o3), xr3, zz, mode, self.eps)
        return g, y3

    def dgrad_join(self, y1, L, g_res, out, o3, xr3, bn3, res, hooks, mode, g_out, y3_out):
        """z = W+^T y1 (1x1 conv L, stride 1) + g_res, then the join chain."""
        J = y1.shape[0]
        C = L.cin
        z = self._dgrad(from_pair(y1) if self.pairs else y1, L.Bd, 1) + g_res.reshape(-1, C)
        r = None if res is None else _rows(res, J)
        if r is not None:
            r = (F.pad(r, (0, C - r.shape[-1])) if r.shape[-1] < C else r).reshape(-1, C)
        g, y3 = self._join_chain(z, _rows(out, J).reshape(-1, C), _rows(o3, J).reshape(-1, C),
                                 _rows(xr3, J).reshape(-1, C), bn3, r, hooks, mode)
        g_out.view(-1, C).copy_(g)
        y3_out.view(-1, C).copy_(to_pair(y3) if self.pairs else y3)

    def join(self, zmain, up, gres_lo, k, out, o3, xr3, bn3, res, hooks, mode, g_out, y3_out, y3_pair=False):
        """Unfused block boundary: z[j,h,w,c] = zmain (at stride `up`: even pixels only) +
        avgpool_k backward of gres_lo (first Cr channels), then the join chain."""
        J, H, W, C = g_out.shape
        z = torch.zeros(J, H, W, C)
        z[:, ::up, ::up, :] += zmain
        if gres_lo is not None:
            cr = gres_lo.shape[-1]
            z[..., :cr] += gres_lo.repeat_interleave(k, 1).repeat_interleave(k, 2) / float(k * k)
        r = None if res is None else _rows(res, J)
        if r is not None and r.shape[-1] < C:
            r = F.pad(r, (0, C - r.shape[-1]))
        g, y3 = self._join_chain(z.view(-1, C), _rows(out, J).reshape(-1, C), _rows(o3, J).reshape(-1, C),
                                 _rows(xr3, J).reshape(-1, C), bn3, None if r is None else r.reshape(-1, C),
                                 hooks, mode)
        g_out.view(-1, C).copy_(g)
        y3_out.view(-1, C).copy_(to_pair(y3) if y3_pair else y3)

    def ds_res(self, g, ap, mode, gres_lo):
        """Residual branch of a downsample block, up to (not including) AvgPool backward:
        Add slot-1 hook -> channel slice -> ConcatChannels hook (both non-affine, a = x = relu(ap))."""
        J = g.shape[0]
        cr = ap.shape[-1]
        a = relu(_rows(ap, J))
        z = hook(False, a, a, g[..., :cr], mode, self.eps)
        gres_lo.copy_(hook(False, a, a, z, mode, self.eps))

    # ------------------------------------------------------------ generic single-hook path
    def hook(self, z_in, z_out, shape, recipe, affine, mode, s0=None, s1=None, s2=None, bn=None, up=1, zc=None, z_in2=None, k2=1,
             pre_scale=1.0, prior=None, P_out=None, relu_or_maxpool=0, post_mask=False, post_scale_row=-1, N=None,
             pre_scale_row=-1, mfm_c=None, out_pair=False):
        """One _backward_ebp firing with optional prior / recording (reference whitebox.py:381-430); include/xfrb.h xfrb_hook."""
        J, H, W, C = shape
        if mfm_c is not None:                            # the MFM backward fused into the Split firing
            zc = torch.empty(J, H, W, C)
            self.mfm_bwd(z_in, mfm_c, zc)
            z_in = zc
        z = torch.zeros(J, H, W, C)
        if z_in is not None:
            z[:, ::up, ::up, :] += z_in.reshape(J, H // up, W // up, -1)[..., :C]
        if z_in2 is not None:
            c2 = z_in2.shape[-1]
            z[..., :c2] += z_in2.reshape(J, H // k2, W // k2, c2).repeat_interleave(k2, 1).repeat_interleave(k2, 2) / float(k2 * k2)
        z = z * pre_scale
        if pre_scale_row >= 0:
            z = z * bn[pre_scale_row]

        def src(t, width=C):
            if t is None:
                return None
            t = _rows(t.reshape(t.shape[0], H, W, -1), J)
            return F.pad(t, (0, width - t.shape[-1])) if t.shape[-1] < width else t
        v0, v1, v2 = src(s0), src(s1), src(s2)
        if bn is not None:
            alpha, beta, sp, tp = bn[0], bn[1], bn[2], bn[3]
        if recipe == 0:
            a = x = relu(v0)
        elif recipe == 1:
            a = relu(v0 * alpha + beta)
            x = relu(relu(v0) * sp + tp)
        elif recipe == 2:
            a = x = relu(v0 * alpha + beta)
        elif recipe == 4:
            a = relu(v0)
            x = relu(relu(v1 * alpha + beta) + (relu(v2) if v2 is not None else 0))
        elif recipe == 6:
            a, x = relu(v0), relu(v1) + relu(v2)
        elif recipe == 7:
            a, x = relu(v0), relu(v1)
        elif recipe == 8:
            a, x = relu(v0), relu(relu(v1) * sp + tp + (v2 if v2 is not None else 0))
        else:
            a, x = relu(v0), v1
        if mode == 3:                                    # XFRB_MODE_NONE: true gradient, dA recording
            if P_out is not None:
                P_out.copy_(z.view_as(P_out))
            ret = z
        else:
            zh = relu(z)
            p = a * zh
            pr = None
            if prior is not None:
                row = prior[0]
                pr = torch.zeros(H * W * C)
                if len(prior) == 2:
                    pr = prior[1].reshape(-1).float().clone()
                else:
                    pr[prior[1]] = prior[2]
                pr = pr.view(H, W, C)
                p = p.clone()
                p[row] = pr
            if P_out is not None:
                P_out.copy_(p.view_as(P_out))
            quo = p / (x + self.eps)
            if mode == MODE_ALL:
                ret = quo
            elif mode == MODE_AFFINEONLY:
                ret = quo if affine else z
            else:
                ret = quo if affine else zh
            if pr is not None:
                ret = ret.clone()
                if mode == MODE_AWP:
                    ret[row] = ((pr > 0) * quo[row]) if affine else ((pr > 0) * z[row])
                if mode == MODE_ALL and relu_or_maxpool == 2:
                    ret[row] = z[row]
        if post_mask:
            ret = ret * (a > 0)
        if post_scale_row >= 0:
            ret = ret * bn[post_scale_row]
        if z_out is not None:
            z_out.copy_(to_pair(ret.reshape(-1, C)).view_as(z_out) if out_pair else ret.view_as(z_out))

    def head_seed(self, Pn, W2, seed):
        seed.copy_(torch.einsum('jc,jcd->jd', Pn, _rows(W2, Pn.shape[0])))

    def normalize_bwd(self, gin, xn, nrm, gout):
        J = gin.shape[0]
        x, n = _rows(xn, J), _rows(nrm, J)
        gout.copy_((gin - x * (x * gin).sum(1, keepdim=True)) / n.unsqueeze(1))

    def maxpool_bwd(self, g, o, bn, out, pool_pad=1, mp_arg=None):
        J = g.shape[0]
        r1 = relu(_rows(o, J) * bn[0] + bn[1]).permute(0, 3, 1, 2)
        _, idx = F.max_pool2d(r1, 3, 2, pool_pad, ceil_mode=(pool_pad == 0), return_indices=True)
        zz = torch.zeros_like(r1).flatten(2).scatter_add_(2, idx.flatten(2), g.permute(0, 3, 1, 2).flatten(2))
        out.copy_(zz.view_as(r1).permute(0, 2, 3, 1))

    def subtree_score(self, gate, gneg, gate_ge0, score, arg):
        m = (gate >= 0) if gate_ge0 else (gate < 0)
        v = (m * (-gneg)).flatten()
        score.fill_(float(v.max()))
        arg.fill_(int(torch.argmax(v)))

    def bn_hook(self, g, o, xr, bn, y, kind, mode):
        """kind 0: BatchNorm backward (gamma+) + BatchNorm hook of a conv output; kind 1: relu(o)*sp + tp."""
        sp, tp = bn[2], bn[3]
        if kind == 1:
            y.copy_(relu(o) * sp + tp)
        else:
            J = g.shape[0]
            y.copy_(hook(True, relu(_rows(o, J)), _rows(xr, J), g * sp, mode, self.eps))

    def head_fwd_linear(self, u, head, v, enc):
        vv = F.avg_pool2d(u.permute(0, 3, 1, 2), 7, 7).flatten(1)
        v.copy_(vv)
        enc.copy_(vv @ _w(head.Bfe).t())

    def head_bwd_linear(self, Pn, W2, head, v, mode, g_out):
        J = Pn.shape[0]
        v_ = _rows(v, J)
        gr = torch.einsum('jc,jcd->jd', Pn, _rows(W2, J)) @ _wpos(head.BfeT).t()
        gr = hook(True, relu(v_), relu(v_), gr, mode, self.eps)
        g_out.copy_((gr / 49.0).view(J, 1, 1, -1).expand(-1, 7, 7, -1))

    def stem_bwd(self, zmain, gres, o, mp, bn, mode, P2, chansum, sums, pool_pad=1, mp_arg=None):
        """Chain at the max-pool output (Conv2d + AvgPool2d(k=1) hooks, both affine), MaxPool
        backward, ReLU / MaxPool2d hooks, ReLU + BN backward, BN hook -> P[-2] = relu(o)*relu(z)."""
        J = zmain.shape[0]
        alpha, beta, sp, tp = bn[0], bn[1], bn[2], bn[3]
        mp_ = _rows(mp, J)
        o_ = _rows(o, J)
        z = zmain if gres is None else zmain + gres
        z = hook(True, mp_, mp_, z, mode, self.eps)
        z = hook(True, mp_, mp_, z, mode, self.eps)
        r1 = relu(o_ * alpha + beta).permute(0, 3, 1, 2)
        _, idx = F.max_pool2d(r1, 3, 2, pool_pad, ceil_mode=(pool_pad == 0), return_indices=True)
        zz = torch.zeros_like(r1).flatten(2).scatter_add_(2, idx.flatten(2), z.permute(0, 3, 1, 2).flatten(2))
        zz = zz.view_as(r1).permute(0, 2, 3, 1)
        r1 = r1.permute(0, 2, 3, 1)
        xrelu = relu(relu(o_) * sp + tp)
        zz = hook(False, r1, xrelu, zz, mode, self.eps)
        zz = hook(False, r1, r1, zz, mode, self.eps)
        zz = zz * (r1 > 0) * sp
        p = relu(o_) * relu(zz)
        P2.copy_(p)
        chansum.copy_(p.sum(-1))
        sums.copy_(p.double().sum(dim=(1, 2, 3)))

    def contrast(self, P2, sums, N, out, thr=None):
        """out[n] = sum_c relu(k*P2[n]/sums[n] - k*P2[N+n]/sums[N+n]), k = (P2[n] >= thr[n]) or 1
        (reference whitebox.py:524-526, 556)."""
        sf = sums.float()
        pm = P2[:N] / sf[:N].view(N, 1, 1, 1)
        pn = P2[N:2 * N] / sf[N:2 * N].view(N, 1, 1, 1)
        d = relu(pm - pn)
        if thr is not None:
            d = d * (P2[:N] >= thr.view(N, 1, 1, 1))
        out.copy_(d.sum(-1))

    def trunc_threshold(self, P2, sums, N, percentile, thr):
        """thr[n] = smallest value of P2[n] whose ascending cumulative sum reaches percentile% of the total
        (reference whitebox.py:550-554, evaluated in double)."""
        for n in range(N):
            srt, _ = torch.sort(P2[n].flatten().double())
            cs = torch.cumsum(srt, 0)
            k = int(torch.searchsorted(cs, torch.tensor(percentile / 100.0 * float(sums[n]), dtype=torch.float64)))
            thr[n] = 0.0 if percentile <= 0 else float(srt[min(k, srt.numel() - 1)])

    def saliency_post(self, mwp, out):
        """gaussian(sigma=2, nearest, truncate 4) -> max(0,.) -> / max(sum, eps)  (whitebox.py:455-460)."""
        import scipy.ndimage as ndi
        for i in range(mwp.shape[0]):
            img = ndi.gaussian_filter(mwp[i].numpy(), 2, mode='nearest', truncate=4.0)
            img = np.maximum(0, img)
            img = img / max(img.sum(), self.eps)
            out[i].copy_(torch.from_numpy(img))

    def twin_blends(self, orig, inp, value, thr, masks, out, mask_f32=False):
        """xfrb_twin_blends: float64 blend of two CHW images under K masks, rounded once to fp32, written NHWC."""
        m = masks if masks is not None else (value.unsqueeze(0) > thr.view(-1, 1, 1)).double()          # [K,H,W]
        m = m.unsqueeze(1)
        w = (1.0 - m.float()).double() if mask_f32 else 1.0 - m
        b = w * orig.double().unsqueeze(0) + m * inp.double().unsqueeze(0)                                 # [K,C,H,W]
        out.copy_(b.float().permute(0, 2, 3, 1))

    # ------------------------------------------------------------ Light-CNN-29v2 pieces (include/xfrb.h)
    def conv_bias(self, inp, B, bias, out, R, positive=False):
        out.view(-1, out.shape[-1]).copy_(im2col_nhwc(inp, R, R, R // 2) @ (_wpos(B) if positive else _w(B)).t() + bias)

    def lc_conv1(self, x, Wt, b, bpos, c, cpos=None):
        C2 = Wt.shape[1]
        w = Wt.t().reshape(C2, 1, 5, 5)
        c.copy_(F.conv2d(x.unsqueeze(1), w, b, 1, 2).permute(0, 2, 3, 1))
        if cpos is not None:
            cpos.copy_(F.conv2d(relu(x).unsqueeze(1), relu(w), bpos, 1, 2).permute(0, 2, 3, 1))

    def mfm_fwd(self, c, m, res=None, y=None, relu_out=None):
        cp = m.shape[-1]
        v = torch.max(c[..., :cp], c[..., cp:])
        m.copy_(v)
        if y is not None:
            v = v + res
            y.copy_(v)
        if relu_out is not None:
            relu_out.copy_(relu(v))

    def mfm_bwd(self, g, c, z):
        cp = g.shape[-1]
        cc = _rows(c, g.shape[0])
        a, b = cc[..., :cp], cc[..., cp:]
        t = torch.where(a == b, g / 2, g)
        z[..., :cp] = t.masked_fill(a < b, 0)
        z[..., cp:] = t.masked_fill(b < a, 0)

    def pool2_fwd(self, m, p, ppos=None):
        f = lambda t: (F.max_pool2d(t.permute(0, 3, 1, 2), 2) + F.avg_pool2d(t.permute(0, 3, 1, 2), 2)).permute(0, 2, 3, 1)
        p.copy_(f(m))
        if ppos is not None:
            ppos.copy_(f(relu(m)))

    def pool2_bwd(self, g, m, gm):
        mm = _rows(m, g.shape[0]).permute(0, 3, 1, 2).contiguous()
        gg = g.permute(0, 3, 1, 2).contiguous()
        _, idx = F.max_pool2d(mm, 2, return_indices=True)
        out = torch.zeros_like(mm).flatten(2).scatter_add_(2, idx.flatten(2), gg.flatten(2)).view_as(mm)
        out = out + F.interpolate(gg, scale_factor=2, mode='nearest') / 4.0
        gm.copy_(out.permute(0, 2, 3, 1))

    def relu(self, inp, out):
        out.copy_(relu(inp))

    def chansum(self, P2, chansum, sums):
        chansum.copy_(P2.sum(-1))
        sums.copy_(P2.double().sum(dim=(1, 2, 3)))
