"""Light-CNN-29v2 whitebox engine (reference python/xfr/models/lightcnn.py:216-275, plugin whitebox.py:113-159).

Forward: every `mfm` (Conv2d(in, 2*out) -> Split -> torch.max, lightcnn.py:48-62) is one implicit-GEMM conv with bias
(xfrb_conv_bias; the 5x5 single-channel stem has its own kernel) whose output c is kept as [rows][2*Cp] (first Split half
in columns [0,Cp), second in [Cp,2*Cp)), followed by the MFM max (+ the resblock Add) and, after conv1 / group1-3 / group4,
the maxpool2 + avgpool2 sum.  Channel counts are padded to the GEMM tile granularity (48 -> 64, 96 -> 128); padded
columns carry zero weights and biases and stay exact zeros in both directions.

Backward: the sweep is issued hook firing by hook firing (87 in triplet mode) through the generic
xfrb_hook kernel, so priors, P recording and true-gradient passes (layerwise_ebp, weighted_subtree_ebp) come for free.
(A, X) of each hooked tensor (no ReLU / BatchNorm modules in this net, so X != A wherever an un-hooked op sits between
two hooked modules):
    direct MFM output m              a = x = relu(m)                               recipe 0
    pooled sum p                     a = relu(p), x = p+ = pool(relu(m))           recipe 3
    resblock output y = out + res    a = relu(y), x = relu(out) + relu(res)        recipe 6
    Split input c = conv(u)          a = relu(c), x = relu(conv_{W+}(relu(u)) + b) recipe 7 (x only read in 'all'/'norelu':
                                     the W+ forward GEMMs run lazily, only for those modes)
    Add slot 0 (on `out`)            (a, x) of the residual (late-binding closure, whitebox.py:379-432)
"""
import torch

from . import packing
from .engine import MODE_IDS, _Engine
from .generic import _FlushBE, _PriorMixin

AFFINE = ('Conv', 'Linear', 'AvgPool', 'BatchNorm')
MODE_NONE = 3
LAYERS = (1, 2, 3, 4)


class _Site(object):
    """One mfm: its pack and the tensors the forward saved."""
    pass


class _FC(object):
    def __init__(self, Bd, cin, signed=None):
        self.Bd, self.cin, self.R = Bd, cin, 1
        self._signed = signed

    def signed_dgrad(self):
        return self._signed


class LightCNNEngine(_Engine):
    map_hw = 128

    def __init__(self, state_dict, backend, layers=LAYERS, device='cpu', with_bias=False, eps=1e-16):
        impl = self._init_base(backend, device, with_bias, eps)
        self.layers = tuple(layers)
        sd = {k: v.detach().cpu().float() for k, v in state_dict.items()}
        self.stem = packing.MfmStem(sd, 'conv1', with_bias).to(self.device)
        self.packs = {}
        chans = ((48, 96), (96, 192), (192, 128), (128, 128))
        for bi, (n, (cin, cout)) in enumerate(zip(self.layers, chans), start=1):
            names = []
            for i in range(n):
                names += ['block%d.%d.conv1' % (bi, i), 'block%d.%d.conv2' % (bi, i)]
            names += ['group%d.conv_a' % bi, 'group%d.conv' % bi]
            for nm in names:
                self.packs[nm] = packing.MfmConv(sd, nm, impl, with_bias).to(self.device)
        self.head = packing.LcHead(sd, impl, with_bias).to(self.device)
        self.fc_pack = _FC(self.head.BfcT_pos, 8192, self.head.BfcT_signed)      # fc seen as a 1x1 'conv' by dgrad_plain
        self.enc_dim = 256

    # ------------------------------------------------------------ forward
    def _mfm(self, S, name, u, hw, res=None):
        """c = conv(u) + b; m = max of the Split halves; y = m + res for the second mfm of a resblock."""
        be, L = self.be, self.packs[name]
        N = u.shape[0]
        st = _Site()
        st.name, st.L, st.u, st.hw = name, L, u, hw
        st.c = self.buf('c:' + name, N, hw, hw, 2 * L.cp)
        st.cpos = None
        be.conv_bias(u, L.Bf, L.bias, st.c, L.R)
        st.m = self.buf('m:' + name, N, hw, hw, L.cp)
        st.y = self.buf('y:' + name, N, hw, hw, L.cp) if res is not None else None
        be.mfm_fwd(st.c, st.m, res, st.y)
        S['sites'][name] = st
        return st

    def forward(self, x_nhwc):
        """x_nhwc [N,128,128,1] in [0,1] (lightcnn.py:19-31).  Fills the saved tensors; returns fc [N,256]
        (WhiteboxLightCNN.encode, whitebox.py:125-128)."""
        be = self.be
        N = x_nhwc.shape[0]
        x = x_nhwc.reshape(N, 128, 128)
        S = {'N': N, 'sites': {}, 'x': x, 'pos': False}
        st = _Site()
        st.name, st.L, st.u, st.hw = 'conv1', self.stem, None, 128
        st.c = self.buf('c:conv1', N, 128, 128, 2 * self.stem.cp)
        st.cpos = None
        be.lc_conv1(x, self.stem.Wt, self.stem.b, self.stem.bpos, st.c, None)
        st.m = self.buf('m:conv1', N, 128, 128, self.stem.cp)
        st.y = None
        be.mfm_fwd(st.c, st.m)
        S['sites']['conv1'] = st
        t, hw = st.m, 128
        stages = []
        for bi, n in enumerate(self.layers, start=1):
            sg = {'blocks': [], 'pool_in': None}
            if bi <= 3:
                sg['pool_in'] = t
                hw //= 2
                p = self.buf('p%d' % bi, N, hw, hw, t.shape[-1])
                pp = self.buf('pp%d' % bi, N, hw, hw, t.shape[-1])
                be.pool2_fwd(t, p, pp)
                sg['p'], sg['ppos'] = p, pp
                t = p
            for i in range(n):
                res = t
                a = self._mfm(S, 'block%d.%d.conv1' % (bi, i), t, hw)
                b = self._mfm(S, 'block%d.%d.conv2' % (bi, i), a.m, hw, res=res)
                sg['blocks'].append({'res': res, 'out': b.m, 'y': b.y})
                t = b.y
            ga = self._mfm(S, 'group%d.conv_a' % bi, t, hw)
            g = self._mfm(S, 'group%d.conv' % bi, ga.m, hw)
            t = g.m
            sg['hw'] = hw
            stages.append(sg)
        S['stages'] = stages
        S['pool4_in'] = t
        S['p4'] = self.buf('p4', N, 8, 8, 128)
        S['p4pos'] = self.buf('pp4', N, 8, 8, 128)
        be.pool2_fwd(t, S['p4'], S['p4pos'])
        S['fc'] = self.buf('fc', N, 256)
        be.conv_bias(S['p4'].view(N, 1, 1, 8192), self.head.Bfc, self.head.bfc, S['fc'].view(N, 1, 1, 256), 1)
        self.saved = S
        return S['fc']

    def ensure_positive(self):
        """The 'positive_activation' pass (whitebox.py:317-330) for the Split hooks' X: c+ = conv_{relu(W)}(relu(u)) + b'.
        Only 'all' / 'norelu' read it (Split is not an affine kind), so it runs on demand."""
        S, be = self.saved, self.be
        if S['pos']:
            return
        N = S['N']
        for name, st in S['sites'].items():
            st.cpos = self.buf('cpos:' + name, *st.c.shape)
            if name == 'conv1':
                be.lc_conv1(S['x'], self.stem.Wt, self.stem.b, self.stem.bpos, st.c, st.cpos)
                continue
            ru = self.buf('relu_u', *st.u.shape)
            be.relu(st.u, ru)
            be.conv_bias(ru, st.L.Bfp, st.L.bias_pos, st.cpos, st.L.R, positive=True)
        S['fcpos'] = self.buf('fcpos', N, 256)
        rv = self.buf('relu_v', N, 1, 1, 8192)
        be.relu(S['p4'].view(N, 1, 1, 8192), rv)
        be.conv_bias(rv, self.head.Bfc_pos, self.head.bfc_pos, S['fcpos'].view(N, 1, 1, 256), 1, positive=True)
        S['pos'] = True

    # ------------------------------------------------------------ operators
    def sweep(self):
        return LightCNNSweep(self)

    def logits(self, W2):
        """classify() of the triplet head for probe 0 (whitebox.py:130-132): fc @ W2^T"""
        return self.saved['fc'][0:1] @ W2[0].t()

    def hooked_logits(self, W2):
        """classify() with the network's own fc2 [C,256] for probe 0 (no bias: lightcnn.py:228)"""
        return self.saved['fc'][0:1] @ W2.t()

    def ebp_backward(self, Pn, W2, mode='affineonly_with_prior', hooked_fc2=False):
        """-> (P2 [J,128,128,2*64] = P[-2] in the padded Split layout, chansum [J,128,128], sums [J])"""
        _, _, P2 = self.sweep().run(Pn, W2, mode, hooked_fc2=hooked_fc2)
        J = Pn.shape[0]
        chansum = self.buf('chansum', J, 128, 128)
        sums = self.buf('sums', J, dtype=torch.float64)
        self.be.chansum(P2, chansum, sums)
        return P2, chansum, sums


class LightCNNSweep(_PriorMixin):
    """Firing-by-firing backward sweep; same interface as xfr_b200.generic.GenericSweep."""

    def __init__(self, engine):
        self.eng = engine
        self.be = engine.be

    def run(self, Pn, W2, mode, priors=None, record=False, true_grad=False, hooked_fc2=False, ptab=None):
        """priors: {firing k: (row, elem, value) | (row, tensor)} in the DEVICE layout (NHWC, padded Split columns), or ptab: a
        generic.PriorTable (device-resident, graph-replayable);
        record: keep p (true_grad: the incoming gradient dA) of every firing; returns (P list | None, names, P2)."""
        eng, be, S = self.eng, _FlushBE(self), self.eng.saved
        N, J = S['N'], Pn.shape[0]
        assert J % N == 0
        self._k = 0
        self._P = [] if record else None
        self._names = []
        self._layout = []                            # per firing: (real channels, padded channels, is_split)
        self._priors = priors or {}
        self._ptab = ptab
        self._norelu = (mode == 'norelu')
        self._k_fcvec = -1
        self._m = m = MODE_NONE if true_grad else MODE_IDS[mode]
        need_x = (not true_grad) and mode in ('all', 'norelu')
        if need_x or (hooked_fc2 and not true_grad):
            eng.ensure_positive()
        buf, head, sites = eng.buf, eng.head, S['sites']

        def dgrad(y, L, out):
            be.dgrad_plain(y, L, out, signed=true_grad)
            return out

        # ---- head: fc2 -> [Linear hook on fc] -> fc (W+) -> Linear hook on v
        seed = buf('lc_seed', J, 1, 1, 256)
        if hooked_fc2:
            W2p = eng.fc2_rows(W2, signed=true_grad)      # relu(W2) (whitebox.py:371-374); the signed rows for true gradients
            raw = buf('lc_seed_raw', J, 1, 1, 256)
            be.head_seed(Pn, W2p, raw.view(J, 256))
            self.fire('Linear', 7, raw, (J, 1, 1, 256), s0=S['fc'], s1=S['fc'] if true_grad else S['fcpos'], out='lc_seed')
        else:
            be.head_seed(Pn, W2, seed.view(J, 256))
        z = dgrad(seed, eng.fc_pack, buf('lc_gv', J, 1, 1, 8192)).view(J, 8, 8, 128)
        self._k_fcvec = self._k
        z = self.fire('Linear', 3, z, (J, 8, 8, 128), s0=S['p4'], s1=S['p4pos'], out='lc_g8')

        def pooled_site(mt, g_p, tag, rc):
            """hooks [MaxPool2d, AvgPool2d] on the tensor that feeds a pooled sum"""
            n_, h_, w_, c_ = mt.shape
            gm = buf('lc_pb' + tag, J, h_, w_, c_)
            be.pool2_bwd(g_p, mt, gm)
            zz = self.fire('MaxPool2d', 0, gm, (J, h_, w_, c_), s0=mt, out='lc_pa' + tag, rc=rc, keep=False)
            return self.fire('AvgPool2d', 0, zz, (J, h_, w_, c_), s0=mt, out='lc_pb' + tag, rc=rc)

        def through_mfm(name, g_m, P_force=None):
            """gradient at an mfm output (its hooks applied) -> gradient at the mfm input (before its hooks)"""
            st = sites[name]
            hw, cp = st.hw, st.L.cp
            last = name == 'conv1'
            # the MFM backward (route the gradient to the larger Split half) is the first thing the Split firing does: no [J,hw,hw,2cp]
            # tensor is written and read back in between
            yc = self.fire('Split', 7, g_m, (J, hw, hw, 2 * cp), s0=st.c, s1=st.cpos if st.cpos is not None else st.c,
                           out=None if last else 'lc_yc', P_force=P_force, rc=st.L.c, split=True, mfm_c=st.c, zc=cp)
            if last:
                return None
            return dgrad(yc, st.L, buf('lc_zu', J, hw, hw, st.L.cin))

        z = pooled_site(S['pool4_in'], z, '4', 128)
        stages = S['stages']
        for bi in range(len(stages), 0, -1):
            sg = stages[bi - 1]
            hw = sg['hw']
            z = through_mfm('group%d.conv' % bi, z)
            mga = sites['group%d.conv_a' % bi].m
            shp = (J, hw, hw, mga.shape[-1])
            rc = sites['group%d.conv_a' % bi].L.c       # real channel count of this stage's block tensors
            z = self.fire('Conv2d', 0, z, shp, s0=mga, out='lc_a', rc=rc)
            z = through_mfm('group%d.conv_a' % bi, z)
            blocks = sg['blocks']
            z2 = None                                # residual-path gradient waiting to be summed at the next firing
            for i in range(len(blocks) - 1, -1, -1):
                B = blocks[i]
                shp = (J, hw, hw, B['y'].shape[-1])
                yrec = dict(s0=B['y'], s1=B['out'], s2=B['res'], rc=rc)
                gname = 'lc_gres%d' % (i % 2)          # Add backward: the same gradient goes to `out` and to `res`
                if i + 1 < len(blocks):
                    z = self.fire('Conv2d', 6, z, shp, z_in2=z2, out='lc_a', keep=False, **yrec)
                    g_res = self.fire('Add', 6, z, shp, out=gname, **yrec)
                else:
                    g_res = self.fire('Conv2d', 6, z, shp, z_in2=z2, out=gname, **yrec)
                if i > 0:
                    pb = blocks[i - 1]
                    rres, rrec = 6, dict(s0=pb['y'], s1=pb['out'], s2=pb['res'], rc=rc)
                elif sg['pool_in'] is not None:
                    rres, rrec = 3, dict(s0=sg['p'], s1=sg['ppos'], rc=rc)
                else:
                    rres, rrec = 0, dict(s0=B['res'], rc=rc)          # block4.0: the residual is group3's MFM output
                z = self.fire('Add', rres, g_res, shp, out='lc_a', **rrec)             # slot 0: the residual's (A, X)
                z = through_mfm('block%d.%d.conv2' % (bi, i), z)
                ma = sites['block%d.%d.conv1' % (bi, i)].m
                z = self.fire('Conv2d', 0, z, shp, s0=ma, out='lc_a', rc=rc)
                z = through_mfm('block%d.%d.conv1' % (bi, i), z)
                z2 = g_res
                if i == 0:
                    z = self.fire('Conv2d', rres, z, shp, z_in2=z2, out='lc_a', keep=False, **rrec)
                    z = self.fire('Add', rres, z, shp, out='lc_b', **rrec)
                    z2 = None
            if sg['pool_in'] is not None:
                z = pooled_site(sg['pool_in'], z, str(bi), rc)
        P2 = buf('lc_P2', J, 128, 128, 2 * eng.stem.cp)
        through_mfm('conv1', z, P_force=P2)
        if self._P is not None:
            self._P.append(None)                     # Conv2d hook on the image: never read by any output
        self._names.append('Conv2d')
        self._layout.append((1, 1, False))
        self._flush()
        return self._P, self._names, P2

    def elem_index(self, k, e, shape):
        """flat index into the reference's [1,C,H,W] MWP of firing k -> flat index into the device tensor (NHWC, padded)"""
        c_real, cp, split = self._layout[k]
        _, H, W, Cdev = shape
        c, hw = divmod(int(e), H * W)
        h, w = divmod(hw, W)
        if split and c >= c_real:
            c = cp + (c - c_real)
        return (h * W + w) * Cdev + c

    def to_reference(self, k, p):
        """Recorded MWP of firing k in the reference's layout: [J,C,H,W] without channel padding (fc vectors: [J,C])."""
        if p is None:
            return None
        c_real, cp, split = self._layout[k]
        if split:
            p = torch.cat((p[..., :c_real], p[..., cp:cp + c_real]), -1)
        else:
            p = p[..., :c_real]
        if k == self._k_fcvec:
            return p.permute(0, 3, 1, 2).reshape(p.shape[0], -1)          # the Linear hook sees the NCHW-flattened vector
        return p.permute(0, 3, 1, 2) if p.shape[1] * p.shape[2] > 1 else p.reshape(p.shape[0], -1)

    def fire(self, kind, recipe, z_in, shape, out, P_force=None, z_in2=None, rc=None, split=False, keep=True, **kw):
        k = self._k
        self._k += 1
        self._names.append(kind)
        cdev = shape[-1]
        self._layout.append((rc if rc is not None else (cdev // 2 if split else cdev), cdev // 2 if split else cdev, split))
        affine = any(s in kind for s in AFFINE)
        P_out = P_force
        if self._P is not None:
            if P_out is None:
                P_out = torch.empty(shape, dtype=torch.float32, device=self.eng.device)
            self._P.append(P_out)
        z_out = self.eng.buf(out, *shape) if out is not None else None
        flag = 2 if (self._norelu and ('MaxPool' in kind or 'ReLU' in kind)) else 0
        self._hook(k, z_in, z_out, shape, recipe, affine, P_out=P_out, keep=keep, relu_or_maxpool=flag, N=self.eng.saved['N'],
                   z_in2=z_in2, k2=1, **kw)
        return z_out
