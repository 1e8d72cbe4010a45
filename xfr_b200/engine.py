"""Host-side schedule of the batched forward + excitation-backprop sweep for the
STR-Janus ResNet (reference resnet.py:168-265, whitebox.py:482-527).

The engine owns the packed weights and a workspace of NHWC fp32 buffers and issues
one kernel per fused stage through a backend object:

  xfr_b200.kernels.CudaBackend  - the product path: ctypes calls into libxfr_b200.so
                                  (hand-written sm_100a kernels), launched on torch's
                                  current CUDA stream;
  tests/emul_backend.EmulBackend - test double used only by tests/ to check this
                                  schedule against the oracle on a CPU-only machine.

What is stored by the forward sweep, per conv c:
  o_c  = conv_c(a_in) + b                (true pre-BN output)
  xr_c = relu(conv_{W+}(a_in) + b)       (the X of the BatchNorm hook, whitebox.py:327)
and per block its output `out` (needed for the hooks chained on it and as the
residual of the next block).  Everything else (a = relu(bn(o)), the gamma+ forward,
ReLU masks) is recomputed inside the backward epilogues from o_c.

Backward rows: the gradient batch holds J = G*N rows (G seed groups, e.g. mate and
non-mate, over the same N probes); row j reads the saved tensors of sample j % N, so
the two contrastive sweeps share one forward and run as one batch.
"""
import os

import torch

from . import packing

MODE_IDS = {'affineonly_with_prior': 0, 'all': 1, 'norelu': 1, 'affineonly': 2}
STRESNET101 = (3, 4, 23, 3)


class _Block(object):
    pass


class _Engine(object):
    """Workspace + the operators composed from forward() / ebp_backward() of a concrete topology."""
    map_hw = 112

    def _init_base(self, backend, device, with_bias, eps):
        self.be = backend
        self.device = torch.device(device)
        self.with_bias = with_bias
        self.eps = eps
        self._ws = {}
        self._graphs = {}             # CUDA graphs of small-batch sweeps (graph_call); dropped whenever a workspace buffer moves
        self.graph_captures = 0       # captures made so far (a steady-state caller should see this stop growing)
        # batch sizes whose ebp / contrastive sweeps are replayed from a captured graph (0: never; XFRB_GRAPH_MAX_N: A/B probe)
        self.graph_max_n = int(os.environ.get('XFRB_GRAPH_MAX_N', '2'))
        # bf16x2 plan: the GEMM operands of the fused sweep (block inputs, inner activations, y1 / y2 / y3) are PAIR tensors -
        # rows of [C bf16 hi | C bf16 lo], the same bytes as fp32 (include/xfrb.h XFRB_IMPL_BF16X2) - written by the producing
        # kernel; block outputs are kept in fp32 as well (residuals and hook chains read them)
        self.pairs = bool(getattr(backend, 'pairs', False))
        impl = getattr(backend, 'impl_name', 'fp32')     # decides the weight planes / tile widths of the packs
        self.conv_impl = getattr(backend, 'conv_pack', impl)
        return impl

    # ------------------------------------------------------------ workspace
    def buf(self, name, *shape, **kw):
        """Named buffer [rows, ...]: one allocation per (name, trailing shape) with a row CAPACITY - a smaller batch (the
        ragged last chunk, a different job count) is a view of the same storage instead of a second complete workspace
        (~210 MB per probe); addresses stay static while the capacity is not exceeded (graph capture)."""
        key = (name,) + tuple(shape[1:])
        t = self._ws.get(key)
        if t is None or t.shape[0] < shape[0]:
            t = torch.empty(shape, dtype=kw.get('dtype', torch.float32), device=self.device)
            self._ws[key] = t
            self._graphs.clear()      # captured graphs hold raw addresses of the buffers they were recorded with
        return t if t.shape[0] == shape[0] else t[:shape[0]]

    # ------------------------------------------------------------ batch-1 latency: one CUDA graph per (operator, shapes, options)
    def graph_call(self, op, tensors, max_n=None, **opts):
        """Every caller of the reference is batch 1 (whitebox.py:482-527): a ResNet-101 sweep is ~225 launches, each with a
        ctypes call and two or three tensor-map encodes on the host, i.e. host-bound at ~9 ms per map.  For batches up to
        graph_max_n the whole sweep is therefore captured once into a CUDA graph (the workspace addresses are static, the tensor
        maps are by-value kernel parameters) and replayed: first call eager (it allocates the workspace), second call captured,
        later calls copy the inputs into the captured input buffers and replay.  Falls back to the eager sweep when capture is
        impossible (CPU emulation backend, a capture already in progress)."""
        fn = getattr(self, op)
        x = tensors[0]
        if (self.graph_max_n <= 0 or x.shape[0] > (self.graph_max_n if max_n is None else max_n) or not x.is_cuda
                or getattr(self.be, 'name', '') != 'cuda' or torch.cuda.is_current_stream_capturing()):
            return fn(*tensors, **opts)
        # tensors flagged static (the network's own fc2, 134 MB for the STR head) are baked in by address instead of copied
        static = tuple(t.data_ptr() if t.dim() == 2 and t.shape[0] > 1024 else None for t in tensors)
        key = (op, tuple((tuple(t.shape), t.dtype) for t in tensors), static, tuple(sorted(opts.items())), float(self.be.eps))
        ent = self._graphs.get(key)
        if ent is None:                                   # first call: eager, allocates every buffer the sweep touches
            out = fn(*tensors, **opts)
            if key not in self._graphs:                   # (buf() may have cleared the table while allocating)
                self._graphs[key] = {'graph': None}
            return out
        if ent['graph'] is None:                          # second call: capture
            ins = [t if p is not None else torch.empty_like(t) for t, p in zip(tensors, static)]
            for a, t, p in zip(ins, tensors, static):
                if p is None:
                    a.copy_(t)
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = getattr(self.be, 'launches', 0)
            with torch.cuda.graph(g):
                out = fn(*ins, **opts)
            if key not in self._graphs:                   # the capture itself allocated (should not happen): stay eager
                return fn(*tensors, **opts)
            ent.update(graph=g, ins=ins, out=out, launches=getattr(self.be, 'launches', 0) - l0)
            self.be.launches = l0                         # nothing ran yet: the replay below counts them
            self.graph_captures += 1
        else:
            for a, t, p in zip(ent['ins'], tensors, static):
                if p is None:
                    a.copy_(t, non_blocking=True)
        ent['graph'].replay()
        self.be.launches += ent['launches']               # kernels of this library the replay enqueued
        return ent['out']

    # ------------------------------------------------------------ firing-by-firing sweeps as replayable graphs
    graph_generic_rows = 64           # gradient-row counts up to which generic_run is replayed from a captured graph (0: never)
    prior_table_len = 512             # >= hook firings of any plugin's sweep (STR ResNet-101: 379, ResNet-50-128d: 158, Light-CNN: 88)

    def prior_table(self, tag):
        """A PriorTable of this engine, one per purpose: a captured graph holds its addresses."""
        from .generic import PriorTable
        key = ('ptab', tag)
        t = self._ws.get(key)
        if t is None:
            t = self._ws[key] = PriorTable(self.prior_table_len, self.device)
        return t

    def generic_run(self, Pn, W2, mode='affineonly_with_prior', record=False, true_grad=False, hooked_fc2=False, ptab=None,
                    gating=None, n_saved=None, zero_seed=None):
        """One firing-by-firing sweep (generic.GenericSweep.run and its twins) as an operator graph_call can capture: the layer
        sweeps and weighted_subtree_ebp repeat the same ~500 launches with nothing but the device-resident priors (ptab)
        changing.  gating (None | bool): also score every firing of a 3-row true-gradient sweep (whitebox.py:684-696, rows =
        cross-entropy / mate / non-mate; True: do_mated_similarity_gating).  n_saved / zero_seed only key the graph table (a
        captured sweep holds the table's row-start pointer or none).
        -> {'gs': the sweep object (layout of the recorded tensors), 'P', 'names', 'P2'[, 'score', 'arg']}"""
        gs = self.sweep()
        P, names, P2 = gs.run(Pn, W2, mode, record=record, true_grad=true_grad, hooked_fc2=hooked_fc2, ptab=ptab)
        out = {'gs': gs, 'P': P, 'names': names, 'P2': P2}
        if gating is not None:
            n = len(P) - 1                                                   # not including the image layer
            score = torch.empty(n, device=self.device)
            arg = torch.empty(n, dtype=torch.int64, device=self.device)
            for k in range(n):
                gate = P[k][1] if gating else P[k][0]
                self.be.subtree_score(gate.contiguous(), P[k][2].contiguous(), bool(gating), score[k:k + 1], arg[k:k + 1])
            out['score'], out['arg'] = score, arg
        return out

    def generic_call(self, Pn, W2, **opts):
        """generic_run through the graph table (CUDA backend) or directly (emulation backend, row counts above graph_generic_rows)"""
        if opts.get('ptab') is not None:
            opts['zero_seed'] = opts['ptab'].zero_seed
        return self.graph_call('generic_run', (Pn.contiguous(), W2), max_n=self.graph_generic_rows, n_saved=self.saved['N'], **opts)

    def graph_fn(self, key, fn):
        """fn() - device work on static addresses, no host synchronisation - through the graph table: eager the first time `key` is
        seen, captured the second time, replayed afterwards (the returned tensors are then the captured ones, overwritten by every
        replay).  Used for the ~2,000 small launches that turn the recorded MWPs of a layer sweep into its per-layer priors."""
        if (self.graph_max_n <= 0 or self.device.type != 'cuda' or getattr(self.be, 'name', '') != 'cuda'
                or torch.cuda.is_current_stream_capturing()):
            return fn()
        ent = self._graphs.get(key)
        if ent is None:
            out = fn()
            self._graphs.setdefault(key, {'graph': None})
            return out
        if ent['graph'] is None:
            torch.cuda.current_stream(self.device).synchronize()
            g = torch.cuda.CUDAGraph()
            l0 = getattr(self.be, 'launches', 0)
            with torch.cuda.graph(g):
                out = fn()
            if key not in self._graphs:
                return fn()
            ent.update(graph=g, out=out, launches=getattr(self.be, 'launches', 0) - l0)
            self.be.launches = l0
            self.graph_captures += 1
        ent['graph'].replay()
        self.be.launches += ent['launches']
        return ent['out']

    def workspace_bytes(self):
        return sum(t.numel() * t.element_size() for t in self._ws.values() if torch.is_tensor(t))

    # ------------------------------------------------------------ composed operators
    def _onehot(self, J, C, cols):
        P = torch.zeros(J, C, dtype=torch.float32)
        for (lo, hi), c in cols:
            P[lo:hi, c] = 1.0
        return P.to(self.device)

    def ebp(self, x_nhwc, Pn, W2, mode='affineonly_with_prior', hooked_fc2=False, saliency=True):
        """Whitebox.ebp over a batch (reference whitebox.py:482-504) -> [N,112,112] device tensor."""
        self.forward(x_nhwc)
        _, chansum, _ = self.ebp_backward(Pn, W2, mode, hooked_fc2)
        if not saliency:
            return chansum
        out = self.buf('sal', chansum.shape[0], self.map_hw, self.map_hw)
        self.be.saliency_post(chansum, out)
        return out

    def contrastive(self, x_nhwc, W2, k_pos=0, k_neg=1, mode='affineonly_with_prior', hooked_fc2=False,
                    saliency=True, num_classes=None, percentile=None):
        """Whitebox.contrastive_ebp / truncated_contrastive_ebp over a batch (reference whitebox.py:506-558):
        one shared forward, mate and non-mate sweeps as one gradient batch of 2N rows."""
        N = x_nhwc.shape[0]
        C = num_classes if num_classes is not None else W2.shape[-2]
        self.forward(x_nhwc)
        Pn = self.priors_contrastive(N, C, k_pos, k_neg)
        P2, _, sums = self.ebp_backward(Pn, W2, mode, hooked_fc2)
        mwp = self.buf('cmwp', N, self.map_hw, self.map_hw)
        thr = None
        if percentile is not None:      # truncated_contrastive_ebp (reference whitebox.py:529-558)
            thr = self.buf('thr', N)
            self.be.trunc_threshold(P2, sums, N, percentile, thr)
        self.be.contrast(P2, sums, N, mwp, thr)
        if not saliency:
            return mwp
        out = self.buf('sal', N, self.map_hw, self.map_hw)
        self.be.saliency_post(mwp, out)
        return out

    def fc2_rows(self, W2, signed=False):
        """The network's own (hooked) fc2 [C,D] as the [1,C,D] operand of head_seed: relu(W) for excitation backprop
        (whitebox.py:371-374), the signed weights for the true-gradient sweeps.  Cached per (storage, version): an in-place
        update of fc2 or a recycled address must not return a stale relu(W)."""
        if signed:
            return W2.unsqueeze(0).contiguous()
        key = ('W2p', W2.data_ptr(), W2._version, tuple(W2.shape))
        W2p = self._ws.get(key)
        if W2p is None:
            for k in [k for k in self._ws if isinstance(k, tuple) and k and k[0] == 'W2p']:
                del self._ws[k]
            W2p = torch.clamp_min(W2, 0).unsqueeze(0).contiguous()
            self._ws[key] = W2p
        return W2p

    def priors_contrastive(self, N, C, k_pos, k_neg):
        key = ('prior', N, C, k_pos, k_neg)
        P = self._ws.get(key)
        if P is None:
            P = self._onehot(2 * N, C, (((0, N), k_pos), ((N, 2 * N), k_neg)))
            self._ws[key] = P
        return P


class StResnetEngine(_Engine):
    """Batched whitebox engine for the STR ResNet topology [3,4,23,3] (or any `layers`)."""

    def __init__(self, state_dict, backend, layers=STRESNET101, device='cpu', with_bias=False, eps=1e-16):
        impl = self._init_base(backend, device, with_bias, eps)
        self.layers = tuple(layers)
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.stem = packing.Stem(sd, with_bias=with_bias).to(self.device)
        self.head = packing.Head(sd, impl, with_bias=with_bias).to(self.device)
        self.blocks = []
        inplanes, hw = 64, 56
        for li, (planes, n) in enumerate(zip((64, 128, 256, 512), self.layers), start=1):
            for bi in range(n):
                b = _Block()
                b.name = 'layer%d.%d' % (li, bi)
                b.stride = 2 if (bi == 0 and li > 1) else 1
                b.has_ds = bi == 0
                b.cin, b.planes, b.cout = inplanes, planes, planes * 4
                b.hw_in = hw
                hw = hw // b.stride
                b.hw = hw
                b.c1 = packing.ConvBN(sd, b.name + '.conv1', b.name + '.bn1', self.conv_impl, with_bias).to(self.device)
                b.c2 = packing.ConvBN(sd, b.name + '.conv2', b.name + '.bn2', self.conv_impl, with_bias).to(self.device)
                b.c3 = packing.ConvBN(sd, b.name + '.conv3', b.name + '.bn3', self.conv_impl, with_bias).to(self.device)
                self.blocks.append(b)
                inplanes = planes * 4
        self.enc_dim = 512

    # ------------------------------------------------------------ forward
    def forward(self, x_nhwc):
        """x_nhwc [N,224,224,3] (mean-subtracted, reference whitebox.py:108-110).  Fills the saved
        tensors and returns xn [N,512] (the unit-norm encoding; encode() = 50*xn)."""
        be = self.be
        N = x_nhwc.shape[0]
        S = {'N': N}
        S['o_s'] = self.buf('o_s', N, 112, 112, 64)
        S['mp'] = self.buf('mp', N, 56, 56, 64)
        S['mp_arg'] = self.buf('mp_arg', N, 56, 56, 64, dtype=torch.uint8)      # which window position won each max-pool
        be.stem_fwd(x_nhwc, self.stem, S['o_s'], S['mp'], S['mp_arg'])
        u = S['mp']
        pairs = self.pairs
        upair = None                        # pair twin of u: the A operand of the next block's conv1
        if pairs:
            upair = self.buf('mp_pair', N, 56, 56, 64)
            be.to_pair(u, upair)
        for i, b in enumerate(self.blocks):
            h = b.hw
            t = {}
            if b.has_ds:
                if b.stride == 2:
                    t['ap'] = self.buf('ap%d' % i, N, h, h, b.cin)
                    be.avgpool2(u, t['ap'])
                    t['us'] = self.buf('us%d' % i, N, h, h, b.cin)
                    be.subsample2(u, t['us'])
                    cin1 = t['us']
                    if pairs:
                        cin1 = self.buf('us_pair', N, h, h, b.cin)
                        be.to_pair(t['us'], cin1)
                else:
                    t['ap'] = u
                    cin1 = upair if pairs else u
                res = t['ap']
            else:
                cin1 = upair if pairs else u
                res = u
            for k, c in (('1', b.planes), ('2', b.planes), ('3', b.cout)):
                t['o' + k] = self.buf('o%s_%d' % (k, i), N, h, h, c)
                t['xr' + k] = self.buf('xr%s_%d' % (k, i), N, h, h, c)
            a1 = self.buf('a1', N, h, h, b.planes)
            a2 = self.buf('a2', N, h, h, b.planes)
            t['out'] = self.buf('out%d' % i, N, h, h, b.cout)
            t['u'] = u
            t['res'] = res
            be.conv_dual(cin1, b.c1, t['o1'], t['xr1'], a1)
            be.conv_dual(a1, b.c2, t['o2'], t['xr2'], a2)
            if pairs:
                upair = self.buf('out_pair%d' % (i % 2), N, h, h, b.cout)
                be.conv_dual(a2, b.c3, t['o3'], t['xr3'], upair, res, act_f32=t['out'])
            else:
                be.conv_dual(a2, b.c3, t['o3'], t['xr3'], t['out'], res)
            S[i] = t
            u = t['out']
        S['v'] = self.buf('v', N, 2048)
        S['f1'] = self.buf('f1', N, 512)
        S['f1p'] = self.buf('f1p', N, 512)
        S['xn'] = self.buf('xn', N, 512)
        S['nrm'] = self.buf('nrm', N)
        S['xmul'] = self.buf('xmul', N, 512)
        be.head_fwd(u, self.head, S['v'], S['f1'], S['f1p'], S['xn'], S['nrm'], S['xmul'])
        self.saved = S
        return S['xn']

    # ------------------------------------------------------------ generic (firing-by-firing) operators
    def sweep(self):
        from .generic import GenericSweep
        return GenericSweep(self)

    def logits(self, W2):
        """classify() of the triplet head for probe 0: (50 * xn) @ W2^T  (resnet.py:252-258, whitebox.py:93-96)"""
        return (50.0 * self.saved['xn'][0:1]) @ W2[0].t()

    # ------------------------------------------------------------ backward
    def hooked_logits(self, W2):
        """classify() with the network's own fc2 [C,512] for probe 0, without its bias (resnet.py:252-258)"""
        return (50.0 * self.saved['xn'][0:1]) @ W2.t()

    def hooked_fc2_seed(self, Pn, W2, m, prior=None, P_out=None, signed=False):
        """The network's own fc2 takes part in EBP (no set_triplet_classifier): gradient Pn @ relu(W2) at the fc2 input,
        then the Linear hook with a = relu(50*xn), x = relu(50*relu(xn)) (reference whitebox.py:371-374, 381-430;
        SURVEY.md appendix A).  W2 [C,512] signed.  Returns the [J,512] gradient after the hook.  signed: the true-gradient
        sweep (m = MODE_NONE): Pn @ W2, the hook only records."""
        S, J = self.saved, Pn.shape[0]
        W2p = self.fc2_rows(W2, signed)
        seed = self.buf('fc2_seed', J, 1, 1, 512)
        self.be.head_seed(Pn, W2p, seed.view(J, 512))
        a50 = self.buf('fc2_a', S['N'], 512)
        x50 = self.buf('fc2_x', S['N'], 512)
        torch.mul(S['xn'], 50.0, out=a50)                       # 512-float glue per probe
        torch.mul(torch.clamp_min(S['xn'], 0), 50.0, out=x50)
        out = self.buf('fc2_hooked', J, 1, 1, 512)
        self.be.hook(seed, out, (J, 1, 1, 512), 5, True, m, s0=a50, s1=x50, prior=prior, P_out=P_out, N=S['N'])
        return out.view(J, 512)

    def ebp_backward(self, Pn, W2, mode='affineonly_with_prior', hooked_fc2=False):
        """One excitation-backprop sweep over J = Pn.shape[0] gradient rows (J % N == 0).
        Pn [J,C] class priors; W2 [N,C,512] per-sample classifier rows (set_triplet_classifier,
        un-hooked, signed) or the network's hooked fc2 [C,512].
        Returns (P2 [J,112,112,64] = P[-2], chansum [J,112,112], sums [J])."""
        be, S = self.be, self.saved
        N = S['N']
        J = Pn.shape[0]
        assert J % N == 0
        m = MODE_IDS[mode]
        nb = len(self.blocks)
        last = self.blocks[-1]
        g = self.buf('g_head', J, 7, 7, last.cout)
        if hooked_fc2:
            Pn, W2 = self.hooked_fc2_seed(Pn, W2, m), None
        be.head_bwd(Pn, W2, self.head, S['v'], S['f1p'], S['xn'], S['nrm'], m, g)
        # chain on the last block output (ReLU + AvgPool2d hooks) and the start of its main path
        t = S[nb - 1]
        gb = self.buf('g%d' % ((nb - 1) % 2), J, last.hw, last.hw, last.cout)
        y3 = self.buf('y3', J, last.hw, last.hw, last.cout)
        pk = dict(y3_pair=True) if self.pairs else {}           # pairs: y1 / y2 / y3 are pair tensors, the g's stay fp32
        pp = dict(pair=True) if self.pairs else {}
        be.join(g, 1, None, 1, t['out'], t['o3'], t['xr3'], last.c3.bn, t['res'], 1, m, gb, y3, **pk)
        for i in range(nb - 1, -1, -1):
            b, t = self.blocks[i], S[i]
            h = b.hw
            y2 = self.buf('y2', J, h, h, b.planes)
            y1 = self.buf('y1', J, h, h, b.planes)
            be.dgrad_mid(y3, b.c3, t['o2'], t['xr2'], b.c2.bn, m, y2)
            be.dgrad_mid(y2, b.c2, t['o1'], t['xr1'], b.c1.bn, m, y1)
            if not b.has_ds:
                p, tp = self.blocks[i - 1], S[i - 1]
                gp = self.buf('g%d' % ((i - 1) % 2), J, h, h, b.cin)
                y3 = self.buf('y3', J, h, h, b.cin)
                be.dgrad_join(y1, b.c1, gb, tp['out'], tp['o3'], tp['xr3'], p.c3.bn, tp['res'], 2, m, gp, y3)
                gb = gp
                continue
            zlo = self.buf('zlo', J, h, h, b.cin)
            be.dgrad_plain(y1, b.c1, zlo, **pp)
            gres = self.buf('gres', J, h, h, b.cin)
            be.ds_res(gb, t['ap'], m, gres)
            if i == 0:
                P2 = self.buf('P2', J, 112, 112, 64)
                chansum = self.buf('chansum', J, 112, 112)
                sums = self.buf('sums', J, dtype=torch.float64)
                be.stem_bwd(zlo, gres, S['o_s'], S['mp'], self.stem.bn, m, P2, chansum, sums, mp_arg=S['mp_arg'])
                return P2, chansum, sums
            p, tp = self.blocks[i - 1], S[i - 1]
            hp = b.hw_in
            gp = self.buf('g%d' % ((i - 1) % 2), J, hp, hp, b.cin)
            y3 = self.buf('y3', J, hp, hp, b.cin)
            be.join(zlo, b.stride, gres, b.stride, tp['out'], tp['o3'], tp['xr3'], p.c3.bn, tp['res'], 3, m,
                    gp, y3, **pk)
            gb = gp


class Resnet50_128Engine(_Engine):
    """Batched whitebox engine for the VGGFace2 ResNet-50-128d (reference models/resnet50_128_pytorch/resnet50_128.py,
    plugin whitebox.py:210-258).  Same kernels as the STR net; differences: conv + BN projection
    shortcuts, the residual sum is a function (no Add hooks), MaxPool2d(3,2,0,ceil), avgpool7 + 1x1 conv 2048->128 head
    with the wrapper's un-hooked fc1 (set_triplet_classifier rows [N,2,128])."""
    STAGES = ((2, 3, 64), (3, 4, 128), (4, 6, 256), (5, 3, 512))

    def __init__(self, state_dict, backend, device='cpu', with_bias=False, eps=1e-16):
        impl = self._init_base(backend, device, with_bias, eps)
        sd = {k: v.detach().cpu() for k, v in state_dict.items()}
        self.stem = packing.Stem(sd, 'conv1_7x7_s2', 'conv1_7x7_s2_bn', with_bias, pool_pad=0).to(self.device)
        self.head = packing.LinearHead(sd, 'feat_extract', impl).to(self.device)
        self.blocks = []
        inplanes, hw = 64, 56
        for st, n, planes in self.STAGES:
            for i in range(1, n + 1):
                b = _Block()
                b.name = 'conv%d_%d' % (st, i)
                b.stride = 2 if (i == 1 and st > 2) else 1
                b.proj = i == 1
                b.cin, b.planes, b.cout = inplanes, planes, planes * 4
                b.hw_in = hw
                hw = hw // b.stride
                b.hw = hw
                mk = lambda c: packing.ConvBN(sd, b.name + c, b.name + c + '_bn', self.conv_impl, with_bias).to(self.device)
                b.c1, b.c2, b.c3 = mk('_1x1_reduce'), mk('_3x3'), mk('_1x1_increase')
                b.cp = mk('_1x1_proj') if b.proj else None
                self.blocks.append(b)
                inplanes = planes * 4
        self.enc_dim = self.head.dim

    def forward(self, x_nhwc):
        """Fills the saved tensors; returns the 128-d encoding [N,128] (Whitebox_resnet50_128.encode, whitebox.py:222-224)."""
        be = self.be
        N = x_nhwc.shape[0]
        S = {'N': N}
        S['o_s'] = self.buf('o_s', N, 112, 112, 64)
        S['mp'] = self.buf('mp', N, 56, 56, 64)
        S['mp_arg'] = self.buf('mp_arg', N, 56, 56, 64, dtype=torch.uint8)      # which window position won each max-pool
        be.stem_fwd(x_nhwc, self.stem, S['o_s'], S['mp'], S['mp_arg'])
        u = S['mp']
        pairs = self.pairs
        upair = None
        if pairs:
            upair = self.buf('mp_pair', N, 56, 56, 64)
            be.to_pair(u, upair)
        for i, b in enumerate(self.blocks):
            h = b.hw
            t = {'u': u}
            cin1 = upair if pairs else u
            if b.stride == 2:
                cin1 = self.buf('us%d' % i, N, h, h, b.cin)
                be.subsample2(u, cin1)
                if pairs:
                    us = cin1
                    cin1 = self.buf('us_pair', N, h, h, b.cin)
                    be.to_pair(us, cin1)
            for k, c in (('1', b.planes), ('2', b.planes), ('3', b.cout)):
                t['o' + k] = self.buf('o%s_%d' % (k, i), N, h, h, c)
                t['xr' + k] = self.buf('xr%s_%d' % (k, i), N, h, h, c)
            a1 = self.buf('a1', N, h, h, b.planes)
            a2 = self.buf('a2', N, h, h, b.planes)
            t['out'] = self.buf('out%d' % i, N, h, h, b.cout)
            be.conv_dual(cin1, b.c1, t['o1'], t['xr1'], a1)
            be.conv_dual(a1, b.c2, t['o2'], t['xr2'], a2)
            if b.proj:
                t['op'] = self.buf('op%d' % i, N, h, h, b.cout)
                t['xrp'] = self.buf('xrp%d' % i, N, h, h, b.cout)
                res = self.buf('resp', N, h, h, b.cout)
                if pairs:       # res is read as a residual only: fp32, no pair twin
                    be.conv_dual(cin1, b.cp, t['op'], t['xrp'], None, relu_act=False, act_f32=res)
                else:
                    be.conv_dual(cin1, b.cp, t['op'], t['xrp'], res, relu_act=False)     # res = bn_p(conv_p(u)), no ReLU
            else:
                res = u
            t['res'] = res if not b.proj else None          # identity: the X of the shortcut operand is u itself
            if pairs:
                upair = self.buf('out_pair%d' % (i % 2), N, h, h, b.cout)
                be.conv_dual(a2, b.c3, t['o3'], t['xr3'], upair, res, act_f32=t['out'])
            else:
                be.conv_dual(a2, b.c3, t['o3'], t['xr3'], t['out'], res)
            S[i] = t
            u = t['out']
        S['v'] = self.buf('v', N, u.shape[-1])
        S['enc'] = self.buf('enc', N, self.head.dim)
        be.head_fwd_linear(u, self.head, S['v'], S['enc'])
        self.saved = S
        return S['enc']

    def sweep(self):
        from .generic import R50Sweep
        return R50Sweep(self)

    def logits(self, W2):
        """classify() for probe 0 (whitebox.py:226-230): fc1(net(x)[0]) with the wrapper's un-hooked rows"""
        return self.saved['enc'][0:1] @ W2[0].t()

    def _xres(self, i, m):
        """Positive-pass value of block i's shortcut operand (only read in mode 'all'): identity blocks -> the block
        input; projection blocks -> BN+(relu(op))."""
        b, t = self.blocks[i], self.saved[i]
        if m != MODE_IDS['all']:
            return None
        if not b.proj:
            return t['u']
        x = self.buf('xresp', *t['op'].shape)
        self.be.bn_hook(None, t['op'], None, b.cp.bn, x, 1, m)
        return x

    def ebp_backward(self, Pn, W2, mode='affineonly_with_prior', hooked_fc2=False):
        """W2 [N,2,128]: rows of the wrapper's un-hooked fc1.  Returns (P2 [J,112,112,64], chansum, sums)."""
        be, S = self.be, self.saved
        N = S['N']
        J = Pn.shape[0]
        assert J % N == 0 and not hooked_fc2
        m = MODE_IDS[mode]
        nb = len(self.blocks)
        last = self.blocks[-1]
        g = self.buf('g_head', J, 7, 7, last.cout)
        be.head_bwd_linear(Pn, W2, self.head, S['v'], m, g)
        t = S[nb - 1]
        gb = self.buf('g%d' % ((nb - 1) % 2), J, last.hw, last.hw, last.cout)
        y3 = self.buf('y3', J, last.hw, last.hw, last.cout)
        pk = dict(y3_pair=True) if self.pairs else {}
        pp = dict(pair=True) if self.pairs else {}
        be.join(g, 1, None, 1, t['out'], t['o3'], t['xr3'], last.c3.bn, self._xres(nb - 1, m), 1 | 4, m, gb, y3, **pk)
        for i in range(nb - 1, -1, -1):
            b, t = self.blocks[i], S[i]
            h = b.hw
            y2 = self.buf('y2', J, h, h, b.planes)
            y1 = self.buf('y1', J, h, h, b.planes)
            be.dgrad_mid(y3, b.c3, t['o2'], t['xr2'], b.c2.bn, m, y2)
            be.dgrad_mid(y2, b.c2, t['o1'], t['xr1'], b.c1.bn, m, y1)
            if not b.proj:
                p, tp = self.blocks[i - 1], S[i - 1]
                gp = self.buf('g%d' % ((i - 1) % 2), J, h, h, b.cin)
                y3 = self.buf('y3', J, h, h, b.cin)
                be.dgrad_join(y1, b.c1, gb, tp['out'], tp['o3'], tp['xr3'], p.c3.bn, self._xres(i - 1, m), 1 | 4, m, gp, y3)
                gb = gp
                continue
            # projection block: both dgrads land on the (sub-sampled) block input
            zlo = self.buf('zlo', J, h, h, b.cin)
            be.dgrad_plain(y1, b.c1, zlo, **pp)
            yp = self.buf('yp', J, h, h, b.cout)
            be.bn_hook(gb, t['op'], t['xrp'], b.cp.bn, yp, 0, m)          # BN backward (gamma+) + BatchNorm hook of proj_bn
            if self.pairs:
                ypp = self.buf('yp_pair', J, h, h, b.cout)
                be.to_pair(yp, ypp)
                yp = ypp
            be.dgrad_plain(yp, b.cp, zlo, accumulate=True, **pp)
            if i == 0:
                P2 = self.buf('P2', J, 112, 112, 64)
                chansum = self.buf('chansum', J, 112, 112)
                sums = self.buf('sums', J, dtype=torch.float64)
                be.stem_bwd(zlo, None, S['o_s'], S['mp'], self.stem.bn, m, P2, chansum, sums, 0, mp_arg=S['mp_arg'])
                return P2, chansum, sums
            p, tp = self.blocks[i - 1], S[i - 1]
            hp = b.hw_in
            gp = self.buf('g%d' % ((i - 1) % 2), J, hp, hp, b.cin)
            y3 = self.buf('y3', J, hp, hp, b.cin)
            be.join(zlo, b.stride, None, 1, tp['out'], tp['o3'], tp['xr3'], p.c3.bn, self._xres(i - 1, m), 3 | 4, m, gp, y3, **pk)
            gb = gp
