"""Generic (one kernel per hook firing) backward sweep of the STR ResNet, for the parts of the reference API that need
what the fused sweep never materialises: the MWP of EVERY firing (`Whitebox.P`, indexed by hook-firing order = the
`k_layer` of reference whitebox.py:561-737), a prior that overrides p at one firing (whitebox.py:390-392), and the
TRUE gradient at every hooked tensor (`self.dA`, whitebox.py:353-358, used by weighted_subtree_ebp).

The schedule below is the firing order of SURVEY.md appendix A (validated against the reference hook for hook); every
firing is one launch of the `xfrb_hook` kernel, convolutions are `xfrb_dgrad_plain` GEMMs.  Gradient rows may carry
different priors: row j's prior sits at firing k_j, which is how weighted_subtree_ebp batches its per-layer sub-trees.
"""
import numpy as np
import torch

from .engine import MODE_IDS

AFFINE = ('Conv', 'Linear', 'AvgPool', 'BatchNorm')       # substring test of reference whitebox.py:399,409
MODE_NONE = 3

# One entry per hook firing, mirrored by XfrbPriorEntry (include/xfrb.h) / PriorEntry (csrc/common.cuh): 48 bytes
PRIOR_DTYPE = np.dtype([('row', '<i4'), ('probe_row', '<i4'), ('elem', '<i8'), ('probe_elem', '<i8'), ('tensor', '<u8'),
                        ('val', '<f4'), ('pad', '<f4'), ('pad2', '<i8')])
assert PRIOR_DTYPE.itemsize == 48


class PriorRef(object):
    """What CudaBackend.hook passes to xfrb_hook for firing k: the address of its table entry and of its probe slot."""
    __slots__ = ('entry_ptr', 'probe_ptr')

    def __init__(self, entry_ptr, probe_ptr):
        self.entry_ptr, self.probe_ptr = entry_ptr, probe_ptr


MAX_ROWS = 64          # gradient rows a PriorTable describes (engine.graph_generic_rows)
NEVER = 2 ** 31 - 1
MAX_CHAIN = 6          # XFRB_MAX_CHAIN (csrc/common.cuh)


class PriorTable(object):
    """The priors of ONE sweep as device data (entry k = hook firing k) instead of launch arguments, so that a sweep captured
    into a CUDA graph is replayed with other priors by rewriting this table: the layer sweeps and weighted_subtree_ebp
    (whitebox.py:584-737) issue the same ~500 launches per sweep with nothing but the priors changing.  An entry may also
    carry a probe: p of one element of one row is written to probe[k] (P_mate at a sub-tree's arg-max node, whitebox.py:699,
    without recording all of P).  On the CPU emulation backend the same table is resolved on the host."""

    def __init__(self, n, device):
        self.n = int(n)
        self.device = torch.device(device)
        self.host = np.zeros(self.n, PRIOR_DTYPE)
        self.dev = torch.zeros(self.n * PRIOR_DTYPE.itemsize, dtype=torch.uint8, device=self.device)
        self.probe = torch.zeros(self.n, dtype=torch.float32, device=self.device)
        self._stage = torch.zeros(self.n * PRIOR_DTYPE.itemsize, dtype=torch.uint8)
        if self.device.type == 'cuda':
            self._stage = self._stage.pin_memory()
        self._keep = {}
        self._copied = None
        # zero-seeded sweeps (every gradient row starts from a zero class prior and comes to life at the firing that carries its
        # prior): start[j] = that firing; the hook kernels skip row j before it and take its incoming gradient as zero there
        self.zero_seed = False
        self.start = torch.full((MAX_ROWS,), NEVER, dtype=torch.int32, device=self.device)
        self._start_stage = torch.full((MAX_ROWS,), NEVER, dtype=torch.int32)
        if self.device.type == 'cuda':
            self._start_stage = self._start_stage.pin_memory()
        self.clear()

    def clear(self, zero_seed=False):
        self.host[:] = 0
        self.host['row'] = -1
        self.host['probe_row'] = -1
        self.host['probe_elem'] = -1
        self._keep.clear()
        self.zero_seed = bool(zero_seed)

    def set_elem(self, k, row, elem, val):
        e = self.host[k]
        e['row'], e['elem'], e['val'], e['tensor'] = row, elem, val, 0

    def set_tensor(self, k, row, tensor):
        assert tensor.is_contiguous() and tensor.dtype == torch.float32 and tensor.device == self.device
        self._keep[k] = tensor                                  # the entry holds a raw address
        e = self.host[k]
        e['row'], e['tensor'] = row, tensor.data_ptr()

    def set_probe(self, k, row, elem):
        e = self.host[k]
        e['probe_row'], e['probe_elem'] = row, elem

    def upload(self):
        if self._copied is not None:
            self._copied.synchronize()         # the previous upload's asynchronous copies have read the pinned staging buffers
        self._stage.copy_(torch.from_numpy(self.host.view(np.uint8)))
        self.dev.copy_(self._stage, non_blocking=True)
        self._upload_start()
        if self.device.type == 'cuda':
            self._copied = torch.cuda.Event()
            self._copied.record(torch.cuda.current_stream(self.device))

    def _upload_start(self):
        if self.zero_seed:
            st = np.full(MAX_ROWS, NEVER, dtype=np.int32)
            ks = np.nonzero(self.host['row'] >= 0)[0]
            rows = self.host['row'][ks]
            assert rows.max(initial=0) < MAX_ROWS and len(set(rows.tolist())) == len(rows), 'one prior per gradient row'
            st[rows] = ks
            self._start_stage.copy_(torch.from_numpy(st))
            self.start.copy_(self._start_stage, non_blocking=True)

    def start_ptr(self):
        return self.start.data_ptr() if self.zero_seed else None

    def ref(self, k):
        return PriorRef(self.dev.data_ptr() + PRIOR_DTYPE.itemsize * k, self.probe.data_ptr() + 4 * k)

    def legacy(self, k):
        """entry k as the (row, tensor) | (row, elem, val) | None argument of the host-resolved path"""
        e = self.host[k]
        if e['row'] < 0:
            return None
        if e['tensor']:
            return (int(e['row']), self._keep[k])
        return (int(e['row']), int(e['elem']), float(e['val']))


class _FlushBE(object):
    """The backend as the sweeps see it: any call other than hook() first launches the pending chain of firings."""

    def __init__(self, sweep):
        self._sweep, self._be = sweep, sweep.be

    def __getattr__(self, name):
        attr = getattr(self._be, name)
        if not callable(attr) or name == 'hook':
            return attr
        sweep = self._sweep

        def call(*a, **kw):
            sweep._flush()
            return attr(*a, **kw)
        return call


class _PriorMixin(object):
    """Where a firing's prior comes from - the {k: prior} dict of run(priors=...) or the PriorTable of run(ptab=...) - and how
    firings reach the backend: on the CUDA backend consecutive firings on the same tensor are queued and launched as ONE
    xfrb_hook chain (link l + 1 takes link l's return value from registers; an intermediate the sweep marked keep=False is
    never written), everything else - the emulation backend, firings with launch-argument priors - goes out one by one."""
    _ptab = None
    _pending = None
    chain_hooks = True

    def _prior_of(self, k):
        """-> (prior argument of backend.hook, probe (row, elem) to resolve on the host or None)"""
        t = self._ptab
        if t is None:
            return self._priors.get(k), None
        assert k < t.n, 'PriorTable shorter than the sweep'
        if getattr(self.be, 'name', '') == 'cuda':
            return t.ref(k), None
        e = t.host[k]
        return t.legacy(k), ((int(e['probe_row']), int(e['probe_elem'])) if e['probe_row'] >= 0 else None)

    def _hook(self, k, z_in, z_out, shape, recipe, affine, P_out=None, keep=True, **kw):
        prior, probe = self._prior_of(k)
        if getattr(self.be, 'name', '') != 'cuda':
            if probe is not None and P_out is None:              # host-resolved probe (emulation backend): needs p itself
                P_out = torch.empty(shape, dtype=torch.float32, device=self.eng.device)
            self.be.hook(z_in, z_out, shape, recipe, affine, self._m, prior=prior, P_out=P_out, **kw)
            if probe is not None:
                self._ptab.probe[k] = P_out[probe[0]].reshape(-1)[probe[1]]
            return
        if self._pending is None:
            self._pending = []
        pend = self._pending
        legacy = prior is not None and not hasattr(prior, 'entry_ptr')       # a launch-argument prior: never chained
        C = shape[-1]
        vec = C % 4 == 0 and all(t is None or t.shape[-1] % 4 == 0 for t in (kw.get('s0'), kw.get('s2')))
        ok = (self.chain_hooks and pend and not legacy and vec and len(pend) < MAX_CHAIN and z_in is not None
              and pend[-1]['z_out'] is not None and pend[-1]['z_out'].data_ptr() == z_in.data_ptr()
              and tuple(pend[-1]['shape']) == tuple(shape) and kw.get('up', 1) == 1 and kw.get('z_in2') is None
              and kw.get('zc') in (None, C) and pend[-1]['chainable'])
        if not ok:
            self._flush()
        pend.append(dict(k=k, z_in=z_in, z_out=z_out, shape=shape, recipe=recipe, affine=affine, prior=prior, P_out=P_out,
                         keep=keep, kw=kw, chainable=vec and not legacy))

    def _flush(self):
        pend = self._pending
        if not pend:
            return
        rs = self._ptab.start_ptr() if self._ptab is not None else None
        n = len(pend)
        for i, L in enumerate(pend):
            last = i == n - 1
            self.be.hook(L['z_in'], L['z_out'] if (last or L['keep']) else None, L['shape'], L['recipe'], L['affine'], self._m,
                         prior=L['prior'], P_out=L['P_out'], chain=0 if n == 1 else (2 if last else 1), row_start=rs, k=L['k'], **L['kw'])
        del pend[:]


class _FC(object):
    """fc1 seen as a 1x1 'conv' by dgrad_plain."""

    def __init__(self, B, cin):
        self.Bd, self.cin, self.R = B, cin, 1

    def signed_dgrad(self):
        return self.Bd


class GenericSweep(_PriorMixin):
    pair_dgrads = True       # bf16x2 plan: pair-tensor dgrads in the W+ sweeps (False: fp32 activations on the split-TF32 kernels)

    def __init__(self, engine):
        self.eng = engine
        self.be = engine.be

    # -------------------------------------------------------------------------------------------------
    @staticmethod
    def elem_index(k, e, shape):
        """The reference indexes flattened [1,C,H,W] tensors (whitebox.py:575-577); device tensors here are [1,H,W,C]."""
        _, H, W, C = shape
        c, hw = divmod(int(e), H * W)
        h, w = divmod(hw, W)
        return (h * W + w) * C + c

    @staticmethod
    def to_reference(k, p):
        """Recorded MWP of firing k as the reference exposes it in self.P: [J,C,H,W] (vectors: [J,C])."""
        if p is None:
            return None
        return p.permute(0, 3, 1, 2) if p.shape[1] * p.shape[2] > 1 else p.reshape(p.shape[0], -1)

    def run(self, Pn, W2, mode, priors=None, record=False, true_grad=False, hooked_fc2=False, ptab=None):
        """One sweep over J = Pn.shape[0] gradient rows.
        priors: {firing k: (row, elem, value) | (row, tensor)}, or ptab: a PriorTable (device-resident, graph-replayable); record: keep p of every firing (list of [J,H,W,C]
        device tensors, NHWC; entry -1, the Conv2d hook on the image, is None: nothing reads it);
        true_grad: no hooks, signed weights, true BatchNorm backward; the recorded tensors are then the gradients dA.
        Returns (P list or None, names, P2 [J,112,112,64] = P[-2])."""
        eng, be, S = self.eng, _FlushBE(self), self.eng.saved
        N, J = S['N'], Pn.shape[0]
        self._k = 0
        self._P = [] if record else None
        self._names = []
        self._priors = priors or {}
        self._ptab = ptab
        self._norelu = (mode == 'norelu')
        m = MODE_NONE if true_grad else MODE_IDS[mode]
        self._m = m
        srow = 0 if true_grad else 2                     # BatchNorm backward: gamma/sigma (true) or gamma+/sigma
        buf = eng.buf
        head = eng.head

        # bf16x2 plan: the W+ dgrads of the block convs run on the pair-tensor kernels (tcgen05 kind::f16, as in the fused sweep):
        # the BatchNorm2d firing that feeds one writes its return value as a pair tensor.  True-gradient sweeps (signed weights,
        # three TF32 passes) and the head's fc dgrad keep fp32 activations.
        pairs = self._pairs = bool(getattr(self.be, 'pairs', False)) and not true_grad and self.pair_dgrads

        def dgrad(y, L, out):
            be.dgrad_plain(y, L, out, signed=true_grad, pair=pairs and getattr(L, 'pair_pack', False))
            return out

        # ---- head: fc2 (un-hooked triplet rows) -> x50 -> Multiply hook -> normalize' -> fc1 -> Linear hook -> AvgPool'
        if hooked_fc2:          # the network's own fc2: W+ and one more (leading) Linear firing
            k = self._k
            self._k += 1
            self._names.append('Linear')
            P_out = torch.empty(J, 1, 1, 512, device=eng.device) if record else None
            if record:
                self._P.append(P_out)
            prior, probe = self._prior_of(k)
            if probe is not None and P_out is None:          # host-resolved probe (emulation backend)
                P_out = torch.empty(J, 1, 1, 512, device=eng.device)
            seed = eng.hooked_fc2_seed(Pn, W2, m, prior=prior, P_out=P_out, signed=true_grad).view(J, 1, 1, 512)
            if probe is not None:
                self._ptab.probe[k] = P_out[probe[0]].reshape(-1)[probe[1]]
        else:
            seed = buf('gs_seed', J, 1, 1, 512)
            be.head_seed(Pn, W2, seed.view(J, 512))
        z = self.fire('Multiply', 5, seed, (J, 1, 1, 512), s0=S['xn'], s1=S['xmul'], pre_scale=50.0, out='gs_mul')
        zn = buf('gs_nb', J, 1, 1, 512)
        be.normalize_bwd(z.view(J, 512), S['xn'], S['nrm'], zn.view(J, 512))
        W1 = head.W1T_signed() if true_grad else head.W1pT
        z = dgrad(zn, _FC(W1, 2048), buf('gs_fc1', J, 1, 1, 2048))
        z = self.fire('Linear', 0, z, (J, 1, 1, 2048), s0=S['v'], out='gs_lin')
        nb = len(eng.blocks)
        last = eng.blocks[-1]
        # AvgPool2d(7) backward rides on the first hook of the last block (z_in2 with k2 = 7)
        zin, zin_up, zin2, k2 = None, 1, z, 7
        for i in range(nb - 1, -1, -1):
            b, t = eng.blocks[i], S[i]
            h, C = b.hw, b.cout
            shp = (J, h, h, C)
            nxt = eng.blocks[i + 1] if i + 1 < nb else None
            res = t['res']
            z = self.fire('ReLU', 4, zin, shp, s0=t['out'], s1=t['o3'], s2=res, bn=b.c3.bn, up=zin_up, z_in2=zin2, k2=k2, out='gs_a', keep=False)
            if nxt is None:
                z = self.fire('AvgPool2d', 0, z, shp, s0=t['out'], post_mask=True, out='gs_g%d' % (i % 2))
            else:
                z = self.fire('Conv2d', 0, z, shp, s0=t['out'], out='gs_b', keep=False)
                z = self.fire('AvgPool2d' if nxt.has_ds else 'Add', 0, z, shp, s0=t['out'], post_mask=True, out='gs_g%d' % (i % 2))
            g = z                                                            # gradient at the block output after ReLU backward
            gres, gres_k = g, 1
            if b.has_ds:
                Cr = b.cin
                zr = self.fire('Add', 0, g, shp, s0=res, out='gs_r1', keep=False)   # slot 1 (residual) fires first
                gres = self.fire('ConcatChannels', 0, zr, (J, h, h, Cr), s0=t['ap'], zc=C, out='gs_r2')
                gres_k = b.stride
            z = self.fire('Add', 0, g, shp, s0=res, post_scale_row=srow, bn=b.c3.bn, out='gs_a', keep=False)   # slot 0: residual's (A, X)
            y3 = self.fire('BatchNorm2d', 3, z, shp, s0=t['o3'], s1=t['xr3'], out='gs_y3', out_pair=pairs)
            z = dgrad(y3, b.c3, buf('gs_z2', J, h, h, b.planes))
            shp2 = (J, h, h, b.planes)
            z = self.fire('ReLU', 1, z, shp2, s0=t['o2'], bn=b.c2.bn, out='gs_c', keep=False)
            z = self.fire('Conv2d', 2, z, shp2, s0=t['o2'], bn=b.c2.bn, post_mask=True, post_scale_row=srow, out='gs_d', keep=False)
            y2 = self.fire('BatchNorm2d', 3, z, shp2, s0=t['o2'], s1=t['xr2'], out='gs_y2', out_pair=pairs)
            z = dgrad(y2, b.c2, buf('gs_z1', J, h, h, b.planes))
            z = self.fire('ReLU', 1, z, shp2, s0=t['o1'], bn=b.c1.bn, out='gs_c', keep=False)
            z = self.fire('Conv2d', 2, z, shp2, s0=t['o1'], bn=b.c1.bn, post_mask=True, post_scale_row=srow, out='gs_d', keep=False)
            y1 = self.fire('BatchNorm2d', 3, z, shp2, s0=t['o1'], s1=t['xr1'], out='gs_y1', out_pair=pairs)
            zlo = dgrad(y1, b.c1, buf('gs_zlo', J, h, h, b.cin))
            # the sum g_main + g_res is taken by the next firing: z_in (stride-2 scatter) + z_in2 (AvgPool backward)
            zin, zin_up, zin2, k2 = zlo, b.stride, gres, gres_k
        # ---- stem
        shp = (J, 56, 56, 64)
        z = self.fire('Conv2d', 0, zin, shp, s0=S['mp'], up=zin_up, z_in2=zin2, k2=k2, out='gs_a', keep=False)     # layer1.0.conv1
        z = self.fire('AvgPool2d', 0, z, shp, s0=S['mp'], out='gs_b')                                   # layer1.0 shortcut (k = 1)
        zz = buf('gs_mp', J, 112, 112, 64)
        be.maxpool_bwd(z, S['o_s'], eng.stem.bn, zz, eng.stem.pool_pad, mp_arg=S.get('mp_arg'))
        shp = (J, 112, 112, 64)
        z = self.fire('ReLU', 1, zz, shp, s0=S['o_s'], bn=eng.stem.bn, out='gs_s1', keep=False)
        z = self.fire('MaxPool2d', 2, z, shp, s0=S['o_s'], bn=eng.stem.bn, post_mask=True, post_scale_row=srow, out='gs_s2', keep=False)
        P2 = buf('gs_P2', *shp)
        self.fire('BatchNorm2d', 3, z, shp, s0=S['o_s'], s1=S['o_s'], out=None, P_force=P2)   # X only shapes the unused return
        if self._P is not None:
            self._P.append(None)                            # Conv2d hook on the image: never read by any output
        self._names.append('Conv2d')
        self._flush()
        return self._P, self._names, P2

    # -------------------------------------------------------------------------------------------------
    def fire(self, kind, recipe, z_in, shape, out, P_force=None, keep=True, **kw):
        """Firing number self._k of the sweep.  keep=False: the return value is read by the NEXT firing only (a chained launch
        then never stores it)."""
        k = self._k
        self._k += 1
        self._names.append(kind)
        affine = any(s in kind for s in AFFINE)
        P_out = P_force
        if self._P is not None:
            if P_out is None:
                P_out = torch.empty(shape, dtype=torch.float32, device=self.eng.device)
            self._P.append(P_out)
        z_out = self.eng.buf(out, *shape) if out is not None else None
        flag = 2 if (self._norelu and ('MaxPool' in kind or 'ReLU' in kind)) else 0
        self._hook(k, z_in, z_out, shape, recipe, affine, P_out=P_out, keep=keep, relu_or_maxpool=flag, N=self.eng.saved['N'], **kw)
        return z_out


class R50Sweep(GenericSweep):
    """The same firing-by-firing sweep for the VGGFace2 ResNet-50-128d (reference resnet50_128.py, plugin whitebox.py:210-258):
    158 firings in triplet mode.  Differences from the STR net: the residual sum is a function (no Add
    hooks; the block ReLU's X sums positive-pass values), projection shortcuts are conv + BatchNorm whose hook fires before
    the main path's, un-hooked fc1 head on the wrapper, 1x1 feat_extract conv after the 7x7 average pool."""

    def run(self, Pn, W2, mode, priors=None, record=False, true_grad=False, hooked_fc2=False, ptab=None):
        assert not hooked_fc2, 'the VGGFace2 plugin has no hooked classifier (whitebox.py:216)'
        eng, be, S = self.eng, _FlushBE(self), self.eng.saved
        N, J = S['N'], Pn.shape[0]
        self._k = 0
        self._P = [] if record else None
        self._names = []
        self._priors = priors or {}
        self._ptab = ptab
        self._norelu = (mode == 'norelu')
        m = MODE_NONE if true_grad else MODE_IDS[mode]
        self._m = m
        srow = 0 if true_grad else 2                     # BatchNorm backward: gamma/sigma (true) or gamma+/sigma
        need_x = (not true_grad) and mode in ('all', 'norelu')
        buf, head = eng.buf, eng.head
        D, C = head.dim, head.cin

        pairs = self._pairs = bool(getattr(self.be, 'pairs', False)) and not true_grad and self.pair_dgrads      # see GenericSweep.run

        def dgrad(y, L, out, accumulate=False):
            be.dgrad_plain(y, L, out, signed=true_grad, accumulate=accumulate, pair=pairs and getattr(L, 'pair_pack', False))
            return out

        seed = buf('gs_seed', J, 1, 1, D)
        be.head_seed(Pn, W2, seed.view(J, D))
        fe = _FC(head.BfeT_signed() if true_grad else head.BfeT, C)
        z = dgrad(seed, fe, buf('gs_fe', J, 1, 1, C))
        z = self.fire('Conv2d', 0, z, (J, 1, 1, C), s0=S['v'], out='gs_lin')            # feat_extract input: x = a = avgpool7(out)
        nb = len(eng.blocks)
        zin, zin_up, zin2, k2 = None, 1, z, 7                                            # AvgPool2d(7) backward rides on the next firing
        for i in range(nb - 1, -1, -1):
            b, t = eng.blocks[i], S[i]
            h, Co = b.hw, b.cout
            shp = (J, h, h, Co)
            nxt = eng.blocks[i + 1] if i + 1 < nb else None
            xres = eng._xres(i, MODE_IDS['all']) if need_x else t['u']      # positive-pass shortcut (only X of the ReLU hook reads it)
            if b.proj and not need_x:
                xres = None
            z = self.fire('ReLU', 8, zin, shp, s0=t['out'], s1=t['o3'], s2=xres, bn=b.c3.bn, up=zin_up, z_in2=zin2, k2=k2, out='gs_a', keep=False)
            if nxt is None:
                g = self.fire('AvgPool2d', 0, z, shp, s0=t['out'], post_mask=True, out='gs_g%d' % (i % 2))
            elif nxt.proj:
                z = self.fire('Conv2d', 0, z, shp, s0=t['out'], out='gs_b', keep=False)
                g = self.fire('Conv2d', 0, z, shp, s0=t['out'], post_mask=True, out='gs_g%d' % (i % 2))
            else:
                g = self.fire('Conv2d', 0, z, shp, s0=t['out'], post_mask=True, out='gs_g%d' % (i % 2))
            zlo = buf('gs_zlo', J, h, h, b.cin)
            if b.proj:       # proj_bn's hook fires before the main path's (it was created later: autograd order)
                yp = self.fire('BatchNorm2d', 3, g, shp, s0=t['op'], s1=t['xrp'], bn=b.cp.bn, pre_scale_row=srow, out='gs_yp', out_pair=pairs)
                dgrad(yp, b.cp, zlo)
            y3 = self.fire('BatchNorm2d', 3, g, shp, s0=t['o3'], s1=t['xr3'], bn=b.c3.bn, pre_scale_row=srow, out='gs_y3', out_pair=pairs)
            z = dgrad(y3, b.c3, buf('gs_z2', J, h, h, b.planes))
            shp2 = (J, h, h, b.planes)
            z = self.fire('ReLU', 1, z, shp2, s0=t['o2'], bn=b.c2.bn, out='gs_c', keep=False)
            z = self.fire('Conv2d', 2, z, shp2, s0=t['o2'], bn=b.c2.bn, post_mask=True, post_scale_row=srow, out='gs_d', keep=False)
            y2 = self.fire('BatchNorm2d', 3, z, shp2, s0=t['o2'], s1=t['xr2'], out='gs_y2', out_pair=pairs)
            z = dgrad(y2, b.c2, buf('gs_z1', J, h, h, b.planes))
            z = self.fire('ReLU', 1, z, shp2, s0=t['o1'], bn=b.c1.bn, out='gs_c', keep=False)
            z = self.fire('Conv2d', 2, z, shp2, s0=t['o1'], bn=b.c1.bn, post_mask=True, post_scale_row=srow, out='gs_d', keep=False)
            y1 = self.fire('BatchNorm2d', 3, z, shp2, s0=t['o1'], s1=t['xr1'], out='gs_y1', out_pair=pairs)
            dgrad(y1, b.c1, zlo, accumulate=b.proj)
            if b.proj:
                zin, zin_up, zin2, k2 = zlo, b.stride, None, 1          # both dgrads landed on the (sub-sampled) block input
            else:
                zin, zin_up, zin2, k2 = zlo, 1, g, 1                    # identity shortcut: the block-output gradient itself
        # ---- stem: conv2_1's reduce and proj Conv2d hooks on the max-pool output, then as the STR net (pool pad 0, ceil)
        shp = (J, 56, 56, 64)
        z = self.fire('Conv2d', 0, zin, shp, s0=S['mp'], up=zin_up, z_in2=zin2, k2=k2, out='gs_a', keep=False)
        z = self.fire('Conv2d', 0, z, shp, s0=S['mp'], out='gs_b')
        zz = buf('gs_mp', J, 112, 112, 64)
        be.maxpool_bwd(z, S['o_s'], eng.stem.bn, zz, eng.stem.pool_pad, mp_arg=S.get('mp_arg'))
        shp = (J, 112, 112, 64)
        z = self.fire('ReLU', 1, zz, shp, s0=S['o_s'], bn=eng.stem.bn, out='gs_s1', keep=False)
        z = self.fire('MaxPool2d', 2, z, shp, s0=S['o_s'], bn=eng.stem.bn, post_mask=True, post_scale_row=srow, out='gs_s2', keep=False)
        P2 = buf('gs_P2', *shp)
        self.fire('BatchNorm2d', 3, z, shp, s0=S['o_s'], s1=S['o_s'], out=None, P_force=P2)
        if self._P is not None:
            self._P.append(None)
        self._names.append('Conv2d')
        self._flush()
        return self._P, self._names, P2
