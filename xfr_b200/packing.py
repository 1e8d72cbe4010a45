"""Weight packing for the sm_100a kernels (host side, pure torch, runs once per handle).

Activations live in HBM as NHWC fp32; every convolution is an implicit GEMM
    D[m, n] = sum_k A[m, k] * B[n, k]
with m = (image, y, x) pixel, k = (tap, input channel) and B stored K-major
("[N][K]", the layout both the tcgen05 shared-memory descriptors and the SIMT
kernel read).

Forward "dual" pack (one GEMM gives the true and the positive pre-activation,
reference whitebox.py:317-330 'positive_activation' pass): the N dimension holds
W and relu(W) side by side, arranged per N-tile of width `tn` as
    [tn/2 rows of W | the same tn/2 rows of relu(W)]
so that one thread of the epilogue owns conv(x) and conv+(x) of the same channel.

Dgrad pack: Z = W+^T Y (reference whitebox.py:371-374) is the same implicit GEMM
with the roles of Cin/Cout swapped and the taps mirrored:
    Bd[ci, (r', s', co)] = relu(W)[co, ci, R-1-r', S-1-s'].
"""
import torch

BN_EPS = 1e-5


def rna_tf32(t):
    """cvt.rna.tf32.f32 on the host: round to nearest (ties away from zero) to a 10-bit mantissa."""
    i = t.contiguous().view(torch.int32)
    return ((i + 0x1000) & ~0x1FFF).view(torch.float32)


def gemm_planes(B, impl):
    """Weight operand as the GEMM implementation wants it (include/xfrb.h):
    'fp32' -> B ; 'tf32' -> rna_tf32(B) ; 'tf32x3' / 'tf32x3full' -> [2, rows, K] = (hi, lo) with hi + lo == B exactly
    (the two-pass W+ GEMMs of 'tf32x3' read only the hi plane)."""
    B = B.float().contiguous()
    if impl == 'fp32':
        return B
    hi = rna_tf32(B)
    if impl == 'tf32':
        return hi.contiguous()
    if impl in ('tf32x3', 'tf32x3full'):
        return torch.stack((hi, B - hi)).contiguous()
    raise ValueError(impl)


def to_pair(x):
    """fp32 [..., C] -> pair tensor of the bf16x2 plan (include/xfrb.h XFRB_IMPL_BF16X2): each row stored as
    [C bf16 hi | C bf16 lo], hi = bf16(x), lo = bf16(x - hi), viewed as float32 [..., C] (the same bytes per row)."""
    x = x.float().contiguous()
    hi = x.bfloat16()
    lo = (x - hi.float()).bfloat16()
    return torch.cat((hi, lo), dim=-1).contiguous().view(torch.float32)


def from_pair(p):
    """pair tensor [..., C] (float32 view) -> fp32 hi + lo (exact: both terms lie inside one 24-bit window)"""
    b = p.contiguous().view(torch.bfloat16)
    C = b.shape[-1] // 2
    return b[..., :C].float() + b[..., C:].float()


def bf16_planes(B, two):
    """Weight operand of the bf16x2 plan: one bf16 plane (relu(W): excitation backprop tolerates coarse W+, the same
    rounded W+ is used for X and for the dgrad), or (hi, lo) bf16 planes [2, rows, K] for signed weights."""
    B = B.float().contiguous()
    hi = B.bfloat16()
    if not two:
        return hi.contiguous()
    return torch.stack((hi, (B - hi.float()).bfloat16())).contiguous()


def dual_tile_width(cout, impl):
    """Tile width of the forward dual pack: the tcgen05 kernel runs N = 256 tiles when the layer is wide enough."""
    return 256 if (impl != 'fp32' and cout % 128 == 0) else 128


def fold_bn(sd, name, with_bias=False):
    """Eval-mode BatchNorm as y = x*alpha + beta (what torch's CPU kernel computes),
    plus the positive-pass / backward constants of excitation backprop:
      sp = relu(gamma)/sqrt(var+eps)   (BN backward with gamma+, and the gamma+ forward scale)
      tp = beta' - mu*sp               (gamma+ forward shift; beta' = relu(beta) iff with_bias)
    Returns a [4, C] tensor (alpha, beta, sp, tp)."""
    g, b = sd[name + '.weight'].float(), sd[name + '.bias'].float()
    mu, var = sd[name + '.running_mean'].float(), sd[name + '.running_var'].float()
    inv = 1.0 / torch.sqrt(var + BN_EPS)
    alpha = g * inv
    beta = b - mu * alpha
    sp = torch.clamp_min(g, 0) * inv
    bb = torch.clamp_min(b, 0) if with_bias else b
    tp = bb - mu * sp
    return torch.stack((alpha, beta, sp, tp)).contiguous()


def pack_dual_fwd(w, b, tn, with_bias=False):
    """w [Cout,Cin,R,S], b [Cout] or None -> (Bf [2*Cout, R*S*Cin], bias [2*Cout]) in tile order."""
    cout, cin, R, S = w.shape
    half = tn // 2
    assert cout % half == 0, (cout, tn)
    wk = w.permute(0, 2, 3, 1).reshape(cout, R * S * cin).float()   # [Cout][(r,s,ci)]
    wp = torch.clamp_min(wk, 0)
    if b is None:
        b = torch.zeros(cout)
    b = b.float()
    bp = torch.clamp_min(b, 0) if with_bias else b
    nt = cout // half
    Bf = torch.stack((wk.view(nt, half, -1), wp.view(nt, half, -1)), dim=1).reshape(2 * cout, -1)
    bias = torch.stack((b.view(nt, half), bp.view(nt, half)), dim=1).reshape(2 * cout)
    return Bf.contiguous(), bias.contiguous()


def pack_dgrad(w, positive=True):
    """w [Cout,Cin,R,S] -> Bd [Cin, R*S*Cout] (taps mirrored), relu'd when positive."""
    cout, cin, R, S = w.shape
    wf = torch.flip(w.float(), dims=(2, 3))
    if positive:
        wf = torch.clamp_min(wf, 0)
    return wf.permute(1, 2, 3, 0).reshape(cin, R * S * cout).contiguous()


def unpack_dual_cols(D, tn):
    """[M, 2*Cout] tile-ordered GEMM result -> (true [M,Cout], pos [M,Cout])."""
    M, n2 = D.shape
    half = tn // 2
    nt = n2 // tn
    D = D.view(M, nt, 2, half)
    return D[:, :, 0].reshape(M, nt * half), D[:, :, 1].reshape(M, nt * half)


class ConvBN(object):
    """One conv + its BatchNorm, packed.  Tensors move with .to(device)."""

    def __init__(self, sd, conv, bn, impl='fp32', with_bias=False):
        w = sd[conv + '.weight']
        b = sd.get(conv + '.bias')
        self.name = conv
        self.impl = impl
        self.cout, self.cin, self.R, self.S = w.shape
        self.tn = dual_tile_width(self.cout, impl)
        Bf, self.bias = pack_dual_fwd(w, b, self.tn, with_bias)
        self.pair_pack = impl == 'bf16x2'       # Bf / Bd are bf16 operands of the pair-tensor GEMMs (fused sweep)
        if self.pair_pack:
            self.Bf = bf16_planes(Bf, two=True)
            self.Bd = bf16_planes(pack_dgrad(w, positive=True), two=False)
        else:
            self.Bf = gemm_planes(Bf, impl)
            self.Bd = gemm_planes(pack_dgrad(w, positive=True), impl)
        self.Bd_signed = None       # true-gradient passes (weighted subtree) pack this lazily
        self._Bd32 = None           # bf16x2 plan: split-TF32 pack of relu(W) for the firing-by-firing sweeps (fp32 activations)
        self.bn = fold_bn(sd, bn, with_bias)
        self._w = w                 # kept for lazy packs only

    def signed_dgrad(self):
        if self.Bd_signed is None:
            impl = 'tf32x3' if self.pair_pack else self.impl
            self.Bd_signed = gemm_planes(pack_dgrad(self._w, positive=False), impl).to(self.Bd.device)
        return self.Bd_signed

    def Bd32(self):
        """relu(W) for the fp32-activation dgrads of the firing-by-firing sweeps under the bf16x2 plan: the SAME bf16-rounded
        values the forward used for X (excitation backprop stays mass-conserving only if X and the dgrad share one W+), as a
        split-TF32 pack - a bf16 number is a TF32 number, so the hi plane holds it exactly and the lo plane is zero."""
        if self._Bd32 is None:
            self._Bd32 = gemm_planes(self.Bd.float().cpu(), 'tf32x3').to(self.Bd.device)
        return self._Bd32

    def to(self, device):
        for k in ('Bf', 'bias', 'Bd', 'bn'):
            setattr(self, k, getattr(self, k).to(device))
        if self.Bd_signed is not None:
            self.Bd_signed = self.Bd_signed.to(device)
        return self


class Stem(object):
    """7x7/2 conv, Cin=3 (reference resnet.py:177-181).  K = 147 is padded to 148."""

    def __init__(self, sd, conv='conv1', bn='bn1', with_bias=False, pool_pad=1):
        self.pool_pad = pool_pad                    # MaxPool2d(3,2,1) for the STR net; (3,2,0,ceil_mode) for VGGFace2
        w = sd[conv + '.weight'].float()            # [64,3,7,7]
        b = sd.get(conv + '.bias')
        b = torch.zeros(w.shape[0]) if b is None else b.float()
        self.cout = w.shape[0]
        wk = w.permute(2, 3, 1, 0).reshape(147, self.cout)   # [(r,s,ci)][co]
        self.W = wk.contiguous()
        self.Wp = torch.clamp_min(wk, 0).contiguous()
        self.b = b.contiguous()
        self.bp = (torch.clamp_min(b, 0) if with_bias else b).contiguous()
        self.bn = fold_bn(sd, bn, with_bias)

    def to(self, device):
        for k in ('W', 'Wp', 'b', 'bp', 'bn'):
            setattr(self, k, getattr(self, k).to(device))
        return self


class Head(object):
    """avgpool7 -> fc1 -> L2 normalise -> x50 -> fc2 (reference resnet.py:235-258)."""

    def __init__(self, sd, impl='fp32', with_bias=False):
        self.tn = tn = 128
        self.W1 = sd['fc1.weight'].float().contiguous()              # [512, 2048]
        self.b1 = sd['fc1.bias'].float().contiguous()
        self.W1p = torch.clamp_min(self.W1, 0).contiguous()
        self.b1p = (torch.clamp_min(self.b1, 0) if with_bias else self.b1).contiguous()
        self.W1pT = gemm_planes(self.W1p.t(), impl)                  # [2048, 512]: B operand of the fc1 dgrad
        B1, self.bias1 = pack_dual_fwd(self.W1.view(512, -1, 1, 1), self.b1, tn, with_bias)
        self.B1 = gemm_planes(B1, impl)
        self.W2 = sd['fc2.weight'].float().contiguous() if 'fc2.weight' in sd else None
        self.scale = 50.0
        self.impl = impl
        self._W1T_signed = None

    def W1T_signed(self):
        """[2048][512] signed fc1^T: the dgrad operand of the true-gradient sweeps (weighted_subtree_ebp)."""
        if self._W1T_signed is None:
            self._W1T_signed = gemm_planes(self.W1.t(), self.impl).to(self.W1.device)
        return self._W1T_signed

    def to(self, device):
        for k in ('W1', 'b1', 'W1p', 'b1p', 'W1pT', 'W2', 'B1', 'bias1'):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device))
        return self


class LinearHead(object):
    """avgpool7 -> 1x1 conv C->D without bias (VGGFace2 ResNet-50-128d: pool5_7x7_s1 + feat_extract,
    reference resnet50_128.py:169-170, 345-347)."""

    def __init__(self, sd, name='feat_extract', impl='fp32'):
        w = sd[name + '.weight'].float()
        self.dim, self.cin = w.shape[0], w.shape[1]
        W = w.reshape(self.dim, self.cin)
        self.Bfe = gemm_planes(W, impl)                                   # [D][C]
        self.BfeT = gemm_planes(torch.clamp_min(W, 0).t(), impl)          # [C][D]: relu(W)^T, the dgrad operand
        self._W, self._impl, self._BfeT_signed = W, impl, None

    def BfeT_signed(self):
        """[C][D] signed feat_extract^T: the dgrad operand of the true-gradient sweeps (weighted_subtree_ebp)."""
        if self._BfeT_signed is None:
            self._BfeT_signed = gemm_planes(self._W.t(), self._impl).to(self.BfeT.device)
        return self._BfeT_signed

    def to(self, device):
        self.Bfe, self.BfeT = self.Bfe.to(device), self.BfeT.to(device)
        return self


# ---------------------------------------------------------------- Light-CNN-29v2 (reference lightcnn.py:48-62, 216-275)
def lc_pad(c):
    """Channel count padded to the GEMM tile granularity (multiples of 64: 48 -> 64, 96 -> 128)."""
    return ((c + 63) // 64) * 64


def _mfm_padded(w, b):
    """w [2C, cin, R, S], b [2C] -> zero-padded (w [2Cp, cin_p, R, S], b [2Cp]) with Split half h at rows [h*Cp, h*Cp + C)."""
    c2, cin, R, S = w.shape
    C = c2 // 2
    cp, cin_p = lc_pad(C), (lc_pad(cin) if cin > 1 else 1)
    wp = torch.zeros(2 * cp, cin_p, R, S)
    bp = torch.zeros(2 * cp)
    for h in (0, 1):
        wp[h * cp:h * cp + C, :cin] = w[h * C:(h + 1) * C]
        bp[h * cp:h * cp + C] = b[h * C:(h + 1) * C]
    return wp, bp, C, cp, cin, cin_p


class MfmConv(object):
    """One mfm's Conv2d(in, 2*out, k, 1, k//2) packed for the forward GEMM (true and relu(W) twins) and the W+ dgrad."""

    def __init__(self, sd, name, impl='fp32', with_bias=False):
        w, b = sd[name + '.filter.weight'].float(), sd[name + '.filter.bias'].float()
        wp, bp, self.c, self.cp, self.cin_real, self.cin = _mfm_padded(w, b)      # cin: padded width (GEMM N of the dgrad)
        self.name, self.impl = name, impl
        self.R = self.S = w.shape[-1]
        flat = lambda t: t.permute(0, 2, 3, 1).reshape(t.shape[0], -1)            # [2Cp][(r,s,ci)]
        self.Bf = gemm_planes(flat(wp), impl)
        self.Bfp = gemm_planes(flat(torch.clamp_min(wp, 0)), impl)
        self.bias = bp.contiguous()
        self.bias_pos = (torch.clamp_min(bp, 0) if with_bias else bp).contiguous()
        self.Bd = gemm_planes(pack_dgrad(wp, positive=True), impl)                  # [cin_p][R*S*2Cp]
        self.Bd_signed = None
        self._w = wp

    def signed_dgrad(self):
        if self.Bd_signed is None:
            self.Bd_signed = gemm_planes(pack_dgrad(self._w, positive=False), self.impl).to(self.Bd.device)
        return self.Bd_signed

    def to(self, device):
        for k in ('Bf', 'Bfp', 'bias', 'bias_pos', 'Bd'):
            setattr(self, k, getattr(self, k).to(device))
        return self


class MfmStem(object):
    """conv1 = mfm(1, 48, 5, 1, 2) (lightcnn.py:219): tap-major weights [25][2*Cp] for the direct kernel."""

    def __init__(self, sd, name='conv1', with_bias=False):
        w, b = sd[name + '.filter.weight'].float(), sd[name + '.filter.bias'].float()
        wp, bp, self.c, self.cp, _, _ = _mfm_padded(w, b)
        self.Wt = wp.reshape(2 * self.cp, 25).t().contiguous()
        self.b = bp.contiguous()
        self.bpos = (torch.clamp_min(bp, 0) if with_bias else bp).contiguous()

    def to(self, device):
        for k in ('Wt', 'b', 'bpos'):
            setattr(self, k, getattr(self, k).to(device))
        return self


class LcHead(object):
    """fc = Linear(8*8*128, 256) on the NCHW-flattened pool4 output (lightcnn.py:271-272); columns re-ordered to the
    NHWC flattening used on the device.  fc2 (if present) is the network's own classifier [C,256]."""

    def __init__(self, sd, impl='fp32', with_bias=False):
        W = sd['fc.weight'].float()
        b = sd['fc.bias'].float()
        Wn = W.view(256, 128, 8, 8).permute(0, 2, 3, 1).reshape(256, 8192).contiguous()
        self.Bfc = gemm_planes(Wn, impl)
        self.Bfc_pos = gemm_planes(torch.clamp_min(Wn, 0), impl)
        self.bfc = b.contiguous()
        self.bfc_pos = (torch.clamp_min(b, 0) if with_bias else b).contiguous()
        self.BfcT_pos = gemm_planes(torch.clamp_min(Wn, 0).t(), impl)              # [8192][256]: dgrad operand
        self.BfcT_signed = gemm_planes(Wn.t(), impl)
        self.W2 = sd['fc2.weight'].float().contiguous() if 'fc2.weight' in sd else None

    def to(self, device):
        for k in ('Bfc', 'Bfc_pos', 'bfc', 'bfc_pos', 'BfcT_pos', 'BfcT_signed', 'W2'):
            v = getattr(self, k)
            if v is not None:
                setattr(self, k, v.to(device))
        return self
