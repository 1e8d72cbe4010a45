/* libxfr_b200.so — C ABI of the B200 (sm_100a) whitebox saliency kernels.
 *
 * Drop-in boundary for the excitation-backprop hot path of stresearch/xfr
 * (python/xfr/models/whitebox.py).  The reference has no FFI of its own: its hot
 * path is torch autograd + Python hooks.  Each entry point below replaces one
 * fused stage of that path and cites the reference lines it stands for.  The
 * Python host (xfr_b200/kernels.py, ctypes) is the only caller; INTEGRATION.md
 * shows the binding.
 *
 * Conventions
 *  - every pointer is a DEVICE pointer to contiguous fp32 (double where stated),
 *    activations are NHWC, caller-owned; nothing is allocated or freed here;
 *  - `stream` is a cudaStream_t (as void*); all work is enqueued on it, nothing
 *    synchronises;
 *  - return value 0 = ok, otherwise an error code; xfrb_last_error() gives text;
 *  - "gradient rows": a backward tensor holds J = G*N images (G seed groups over
 *    the same N probes); row j reads the forward tensors of sample j % N;
 *  - `mode`: 0 = affineonly_with_prior, 1 = all (== norelu without a prior),
 *    2 = affineonly   (reference whitebox.py:397-430, no prior set);
 *  - `impl`: 0 = fp32 CUDA-core implicit GEMM; 1 = tcgen05 split-TF32: three passes (fp32-equivalent) wherever a
 *    weight is signed (true forward), two passes (activation split exactly, relu(W) rounded to TF32) for the W+ GEMMs,
 *    whose products are all non-negative; 2 = tcgen05 single-pass TF32; 3 = tcgen05 3xTF32 in every GEMM
 *    (xfrb_dgrad_plain with signed weights - the true-gradient sweep of weighted_subtree_ebp - must be called with 3);
 *    4 = opt-in: two passes in every GEMM kind, the weight operand rounded to TF32 even where it is signed (XFRB_IMPL_TF32X2;
 *    the host's 'tf32x2f' plan passes it to xfrb_conv_dual only);
 *  - weight operands (`Bf`, `Bd`, `B1`, `W1pT`) are K-major [rows][K] fp32.  For impl 1 they hold TWO planes
 *    [2][rows][K]: hi = rna_tf32(W) and lo = W - hi (the host does the 3xTF32 split of the static operand once,
 *    xfr_b200/packing.py); for impl 2 one plane of rna_tf32(W); for impl 0 one plane of W;
 *  - `bn` is [4][C]: alpha, beta (eval BatchNorm y = x*alpha+beta), sp = relu(gamma)/sigma,
 *    tp = beta' - mu*sp (the gamma+ forward of the 'positive_activation' pass).
 */
#ifndef XFRB_H_
#define XFRB_H_

#ifdef __cplusplus
extern "C" {
#endif

#define XFRB_MODE_AWP 0
#define XFRB_MODE_ALL 1
#define XFRB_MODE_AFFINEONLY 2

#define XFRB_IMPL_FP32 0
#define XFRB_IMPL_TF32X3 1
#define XFRB_IMPL_TF32 2
#define XFRB_IMPL_TF32X3_FULL 3
#define XFRB_IMPL_TF32X2 4   /* two passes in every GEMM: activations exact (hi + lo), weights rounded to TF32 (plane 0 of the
                              two-plane packs).  Opt-in (kernels.HYBRID_IMPLS 'tf32x2f' uses it for xfrb_conv_dual only) */
#define XFRB_IMPL_BF16X2 5   /* the bf16x2 plan of the fused sweep (xfrb_conv_dual, xfrb_dgrad_mid / _join / _plain): tcgen05
                              kind::f16 with bf16 terms - activations / gradients 2 terms, relu(W) 1 term, signed W 2 terms.
                              The activation operand (`inp`, `y`, `y1`) and every output that feeds the next GEMM (`act`,
                              `y_out`, `y3_out`) are then PAIR tensors: each row of C values is stored as [C bf16 hi | C bf16 lo]
                              (hi = bf16(x), lo = bf16(x - hi); the same 4*C bytes per row as fp32; xfrb_to_pair converts).
                              Weight operands are bf16 K-major: `Bd` one plane of relu(W) [rows][K]; `Bf` two planes
                              [2][rows][K] (hi, lo) of the dual pack.  Channel counts of the A operand must be multiples of 64 */

int xfrb_version(void);
const char* xfrb_last_error(void);
/* 1 when the device under the current context is compute capability 10.x */
int xfrb_device_ok(void);
/* 1 when GEMM implementation `impl` (XFRB_IMPL_*) is compiled into this library */
int xfrb_impl_available(int impl);
/* tcgen05 kernels as CTA pairs (cta_group::2, one 256-row tile per TPC: each SM stages and reads half of the weight tile) for
 * the forward dual conv and the MID dgrads where a launch has enough tiles: on by default (XFRB_CTA2=0 in the environment
 * disables); returns the previous setting.  Results are bit-identical either way. */
int xfrb_set_cta_pairs(int on);
/* tcgen05 kernels as clusters of two CTAs that TMA-multicast the weight tiles to each other (half the weight traffic out of
 * L2 per CTA; MMAs, TMEM and epilogue stay private): measured neutral, off by default (XFRB_MC=1 enables);
 * returns the previous setting.  Results are bit-identical either way. */
int xfrb_set_multicast_pairs(int on);
/* Host-only helper (no GPU needed): the 4-D TMA box the tcgen05 kernel uses for the activation tiles of a 3x3 conv over
 * Nimg maps of H x W pixels - bimg images x bh image rows x W pixels <= 128 GEMM rows.  Returns the fraction of the 128 rows
 * that carry pixels (-1 on bad arguments). */
double xfrb_tile_geometry(int H, int W, int Nimg, int* bh, int* bimg);

/* ---- forward ("activation" + "positive_activation" passes, whitebox.py:490-493) ---- */

/* STR ResNet stem: o = conv7x7/2(x)+b  [N,112,112,64];  mp = maxpool3x3/2(relu(bn(o))) [N,56,56,64]
 * (reference resnet.py:225-228).  W is [147][64] ((r,s,ci) major), x is [N,224,224,3].  pool_pad = 1 (STR net) or 0
 * (VGGFace2 ResNet-50: MaxPool2d(3,2,0,ceil_mode=True), resnet50_128.py:16; windows clipped at the border).
 * mp_arg (may be NULL) [N,56,56,64] bytes: window position r*3+s of the first maximum, consumed by xfrb_stem_bwd. */
int xfrb_stem_fwd(const float* x, const float* W, const float* b, const float* bn,
                  float* o, float* mp, unsigned char* mp_arg, int N, int pool_pad, void* stream);

/* fp32 [rows,C] -> pair tensor [rows][C bf16 hi | C bf16 lo] (inverse = 0), or back to fp32 hi + lo (inverse = 1); C % 4 == 0.
 * The bf16x2 plan's GEMM operand format (XFRB_IMPL_BF16X2); used where a non-GEMM kernel produces a GEMM input. */
int xfrb_to_pair(const float* in, float* out, long long rows, int C, int inverse, void* stream);

/* u[N,H,W,C] -> out[N,H/2,W/2,C]: even pixels (input of a stride-2 1x1 conv, resnet.py:116) */
int xfrb_subsample2(const float* u, float* out, int N, int H, int W, int C, void* stream);
/* u[N,H,W,C] -> out[N,H/2,W/2,C]: AvgPool2d(2,2) of the shortcut (resnet.py:211) */
int xfrb_avgpool2(const float* u, float* out, int N, int H, int W, int C, void* stream);

/* One conv of a Bottleneck in both passes at once (resnet.py:132-149; whitebox.py:317-330):
 *   o   = conv_W(inp) + b                     [N,H,W,Cout]   (true pre-BN output)
 *   xr  = relu(conv_relu(W)(inp) + b')        [N,H,W,Cout]   (X of the BatchNorm hook)
 *   act = relu(o*alpha + beta + res)          [N,H,W,Cout]   (res: [N,H,W,res_c], zero beyond res_c; may be NULL)
 * Bf/bias: dual pack of xfr_b200/packing.py (tile width tn).  R = 1 or 3, stride 1, pad R/2.
 * relu_act = 0 emits act = o*alpha + beta + res without the ReLU (conv + BN projection shortcut, resnet50_128.py:185-187).
 * impl XFRB_IMPL_BF16X2: inp and act are pair tensors (act may be NULL) and act_f32 (may be NULL) receives act in fp32 as
 * well (block outputs are read again by epilogues: residual, hooks); every other impl: act_f32 must be NULL. */
int xfrb_conv_dual(const float* inp, const float* Bf, const float* bias, const float* bn,
                   const float* res, int res_c, float* o, float* xr, float* act, float* act_f32,
                   int N, int H, int W, int Cin, int Cout, int R, int tn, int relu_act, int impl, void* stream);

/* avgpool7 -> fc1 (+ the W+ twin) -> L2 normalise (resnet.py:235-252).
 * B1 = dual pack of fc1 [1024][2048], bias1 [1024] (tile width tn); scratch [N,1024].
 * v [N,2048], f1 [N,512], f1p [N,512] (fc1 with relu(W) on relu(v)), xn [N,512], nrm [N],
 * xmul [N,512] = relu(normalize(f1p)) (X of the Multiply hook; may be NULL). */
int xfrb_head_fwd(const float* u, const float* B1, const float* bias1, int tn, float* scratch,
                  float* v, float* f1, float* f1p, float* xn, float* nrm, float* xmul, int N, int impl, void* stream);

/* ---- backward (the 'ebp' pass + Xn.backward(Pn), whitebox.py:496-498) ---- */

/* Head: seed = Pn @ W2 (un-hooked per-sample triplet classifier, whitebox.py:93-96), x50,
 * Multiply hook, Jacobian of F.normalize, fc1 with relu(W) (W1pT [2048][512]), Linear hook,
 * AvgPool2d(7) backward.  Pn [J,C], W2 [N,C,512]; scratch [J,2560]; g_out [J,7,7,2048].
 * W2 == NULL: Pn is already the [J,512] gradient at the fc2 input (the network's own hooked fc2: xfrb_head_seed with
 * relu(fc2.weight) followed by the Linear hook through xfrb_hook, whitebox.py:371-374). */
int xfrb_head_bwd(const float* Pn, const float* W2, int C, const float* W1pT,
                  const float* v, const float* f1p, const float* xn, const float* nrm,
                  float* scratch, float* g_out, int J, int N, int mode, float eps, int impl, void* stream);

/* z = relu(W)^T y through conv (R = 1|3), then the hook chain at the activation
 * a = relu(bn(o)) feeding that conv: ReLU hook, Conv2d hook, ReLU backward, BatchNorm
 * backward with gamma+, BatchNorm hook (whitebox.py:381-430).  y [J,H,W,Cout] -> y_out [J,H,W,Cin];
 * o, xr [N,H,W,Cin] are the saved tensors of the conv that PRODUCED a; bn is its BatchNorm. */
int xfrb_dgrad_mid(const float* y, const float* Bd, const float* o, const float* xr, const float* bn,
                   float* y_out, int J, int N, int H, int W, int Cin, int Cout, int R,
                   int mode, float eps, int impl, void* stream);

/* z_out (+)= B^T y only (downsample / projection blocks; true-gradient passes of weighted_subtree_ebp).
 * accumulate = 1 adds onto z_out (the second of the two dgrads that meet at a projection block's input). */
int xfrb_dgrad_plain(const float* y, const float* Bd, float* z_out,
                     int J, int H, int W, int Cin, int Cout, int R, int accumulate, int impl, void* stream);

/* Identity-block boundary: z = relu(W1)^T y1 + g_res, hooks chained on the previous block's
 * output `out` (ReLU; Conv2d; then `hooks & 3`: 1 none, 2 Add(non-affine), 3 a second affine hook; `hooks & 4`: the
 * residual sum is torch.add, not a module - no Add hook, other X for the block ReLU: VGGFace2 ResNet-50),
 * ReLU backward -> g_out; then Add slot-0 hook (residual's (A,X): the late-binding closure of
 * whitebox.py:379-432), BatchNorm backward, BatchNorm hook -> y3_out.  All [.,H,W,C], C = Cin.
 * Bits 8-10 of `hooks` are profiling switches of the tcgen05 epilogue (1: no global loads, 2: no stores, 4: no hook math;
 * tools/epi_probe.py) and must be 0 in production calls. */
int xfrb_dgrad_join(const float* y1, const float* Bd, const float* g_res,
                    const float* out, const float* o3, const float* xr3, const float* bn3,
                    const float* res, int res_c, float* g_out, float* y3_out,
                    int J, int N, int H, int W, int Cin, int Cout,
                    int hooks, int mode, float eps, int impl, void* stream);

/* Same boundary without the GEMM (head -> last block, and below a downsample block):
 * z[j,h,w,c] = zmain[j,h/up,w/up,c] on pixels divisible by `up` (else 0)
 *            + gres_lo[j,h/k,w/k,c]/(k*k) for c < gres_c   (AvgPool2d(k) backward).
 * y3_pair != 0: y3_out is written as a pair tensor (the A operand of the next dgrad under XFRB_IMPL_BF16X2). */
int xfrb_join(const float* zmain, int up, const float* gres_lo, int gres_c, int k,
              const float* out, const float* o3, const float* xr3, const float* bn3,
              const float* res, int res_c, float* g_out, float* y3_out,
              int J, int N, int H, int W, int C, int hooks, int mode, float eps, int y3_pair, void* stream);

/* Shortcut branch of a downsample block up to AvgPool backward: Add slot-1 hook, channel
 * slice, ConcatChannels hook (resnet.py:152-157).  g [J,H,W,C], ap [N,H,W,Cr] -> gres_lo [J,H,W,Cr]. */
int xfrb_ds_res(const float* g, const float* ap, float* gres_lo,
                int J, int N, int H, int W, int C, int Cr, int mode, float eps, void* stream);

/* Stem: hooks on the max-pool output, MaxPool backward (first maximum wins, as torch),
 * ReLU / MaxPool2d hooks, ReLU + BatchNorm backward, BatchNorm hook:
 *   P2 = P[-2] = relu(o)*relu(z) [J,112,112,64]; chansum [J,112,112]; sums [J] (double).
 * zc is scratch [J,56,56,64].  mp_arg: the arg-max bytes of xfrb_stem_fwd, or NULL to re-derive them from o. */
int xfrb_stem_bwd(const float* zmain, const float* gres, const float* o, const float* mp, const float* bn,
                  float* zc, float* P2, float* chansum, double* sums, const unsigned char* mp_arg,
                  int J, int N, int mode, float eps, int pool_pad, void* stream);

/* ---- VGGFace2 ResNet-50-128d pieces (reference models/resnet50_128_pytorch/resnet50_128.py, whitebox.py:210-258) ---- */

/* kind 0: y = BatchNorm hook after BatchNorm backward with gamma+, relu(o)*relu(g*sp)/(xr+eps)  (projection shortcut);
 * kind 1: y = relu(o)*sp + tp, the positive-pass BatchNorm output (X of the shortcut operand, mode 'all').
 * g, y [J,HW,C] (kind 1: [N,HW,C]); o, xr [N,HW,C]. */
int xfrb_bn_hook(const float* g, const float* o, const float* xr, const float* bn, float* y,
                 int J, int N, int HW, int C, int kind, int mode, float eps, void* stream);

/* v = avgpool7(u) [N,C]; enc = v @ Wfe^T [N,D]  (pool5_7x7_s1 + feat_extract, resnet50_128.py:345-347) */
int xfrb_head_fwd_linear(const float* u, const float* Bfe, float* v, float* enc, int N, int C, int D, int impl, void* stream);

/* seed = Pn @ W2 (un-hooked fc1 of the wrapper, whitebox.py:216-230), z = seed @ relu(Wfe) (BfeT [C][D]), Conv2d hook
 * with a = x = v, AvgPool backward.  scratch [J,(D+C)]; g_out [J,7,7,C]. */
int xfrb_head_bwd_linear(const float* Pn, const float* W2, int Ccls, const float* BfeT, const float* v, float* scratch,
                         float* g_out, int J, int N, int C, int D, int mode, float eps, int impl, void* stream);

/* out[n] = sum_c relu(k*P2[n]/sums[n] - k*P2[N+n]/sums[N+n])  (whitebox.py:524-526); k = 1, or with thr != NULL the
 * truncation mask k = (P2[n] >= thr[n]) of whitebox.py:550-556 */
int xfrb_contrast(const float* P2, const double* sums, const float* thr, float* out, int N, int HW, int C, void* stream);

/* Truncated contrastive EBP (whitebox.py:550-554): thr[n] = the smallest mate-MWP value whose ascending cumulative sum
 * reaches percentile% of the total, over the first N rows of P2 (per_sample elements each). */
int xfrb_trunc_threshold(const float* P2, const double* sums, float percentile, float* thr, int N, long long per_sample,
                         void* stream);

/* skimage.filters.gaussian(sigma=2) -> max(0,.) -> /max(sum,eps)  (whitebox.py:455-460); [B,H,W], H,W <= 128 */
int xfrb_saliency_post(const float* mwp, float* out, int B, int H, int W, float eps, void* stream);

/* The saliency .npz format's cubic resize (python/xfr/show.py:131-137 processSaliency: attMap -= min; attMap /= (max + 1e-9);
 * skimage.transform.resize(attMap, img.shape[:2], order=3, mode='constant')).  scikit-image >= 0.19 evaluates that resize as
 * scipy.ndimage.zoom(order=3, mode='grid-constant', cval=0, grid_mode=True) + clip to the input range; this kernel is that
 * algorithm (zero pad 12, cubic B-spline prefilter per axis in double, half-pixel-centred sampling), one map per CTA.
 * in [B,h,w] fp32 -> out [B,oh,ow] fp32; h, w <= 144.  normalize != 0 applies the min-shift / max-normalise first. */
int xfrb_cubic_zoom(const float* in, float* out, int B, int h, int w, int oh, int ow, int normalize, void* stream);

/* Inpainting-game blends (python/xfr/inpainting_game/inpainting_game.py:110-132, consumed by Whitebox.embeddings,
 * whitebox.py:747-785): out[k,h,w,c] = fp32((1 - m) * orig[c,h,w] + m * inp[c,h,w]), evaluated in double, with
 * m = (value[h,w] > thr[k]) (inpainting_game.py:66; `masks` = NULL) or m = masks[k,h,w] (blurred masks, lines 69-78).
 * orig / inp [C,H,W] double (network format, as the reference's astype(np.float64)), value [H,W] / thr [K] / masks [K,H,W]
 * double, out [K,H,W,C] fp32 (NHWC: the forward sweep's input).  mask_f32 != 0: the masks were float32 numbers (numpy then
 * evaluates 1 - m in float32).  C = 1 or 3, H*W % 4 == 0, K <= 65535. */
int xfrb_twin_blends(const double* orig, const double* inp, const double* value, const double* thr, const double* masks, float* out,
                     int K, int C, int H, int W, int mask_f32, void* stream);

/* ---- generic single-hook path: priors, P recording, true gradients (whitebox.py:561-737) ---- */

#define XFRB_MODE_NONE 3   /* no hook fires (plain backprop); P_out then records the incoming gradient (self.dA) */

/* One _backward_ebp firing (whitebox.py:381-430) over [J,H,W,C]:
 *   z = z_in[at stride up, first C of zc channels] + z_in2[/k2, c < c2]/(k2*k2);  z *= pre_scale
 *   (a, x) from `recipe` over the saved tensors s0 [N,H,W,c0], s1 [N,H,W,C], s2 [N,H,W,c2s] and bn [4][C]
 *     0: a = x = relu(s0)   1: a = relu(bn(s0)), x = relu(relu(s0)*sp+tp)   2: a = x = relu(bn(s0))
 *     3: a = relu(s0), x = s1   4: a = relu(s0), x = relu(relu(bn(s1)) + relu(s2))   5: a = relu(s0), x = s1
 *     6: a = relu(s0), x = relu(s1) + relu(s2)   7: a = relu(s0), x = relu(s1)      (Light-CNN: resblock output / Split input)
 *     8: a = relu(s0), x = relu(relu(s1)*sp + tp + s2)                              (VGGFace2 ResNet-50 block ReLU hook)
 *   pre_scale_row >= 0: z *= bn[pre_scale_row][c] before the hook (the BatchNorm backward that precedes a BatchNorm hook);
 *   p = a*relu(z), replaced for gradient row `prior_row` by the prior (a full tensor `prior` [H*W*C], or the single element
 *   prior_elem = prior_val); P_out <- p; return value per `mode` (`affine`: Conv/Linear/AvgPool/BatchNorm kinds;
 *   relu_or_maxpool = 2 marks ReLU/MaxPool kinds for the 'norelu' rule whitebox.py:418-419, passed with mode 1);
 *   then optionally masked by (a > 0) and scaled by bn[post_scale_row][c] (ReLU / BatchNorm backward); -> z_out.
 *   prior_entry (device pointer to ONE XfrbPriorEntry, or NULL): the prior of this firing read from device memory instead of the
 *   four prior arguments - a sweep captured into a CUDA graph is then replayed with other priors by rewriting the table, which is
 *   how the layer sweeps and weighted_subtree_ebp (whitebox.py:584-737) run without per-launch host work; with it, probe_out
 *   (device float, or NULL) receives p of element probe_elem of row probe_row (P_mate at the arg-max node, whitebox.py:699).
 *   chain: consecutive firings on the SAME [J,H,W,C] tensor can run as one launch - 1 appends this firing to the calling thread's
 *   pending chain (nothing is launched), 2 appends it and launches the chain, 0 launches this firing alone; from the second link on
 *   z_in / up / zc / z_in2 / k2 / c2 are ignored (the link takes its predecessor's return value from registers) and a NULL z_out
 *   is simply not stored (at most XFRB_MAX_CHAIN = 6 links).  row_start ([J] device ints or NULL) with k = this firing's index
 *   (those of the chain's first link count): gradient row j is skipped by every firing before row_start[j] and enters firing
 *   row_start[j] with a zero gradient - the rows of a zero-seeded sweep whose priors sit at different firings.
 *   mfm_c (or NULL; first link only): z_in is the gradient at a Light-CNN MFM OUTPUT [J,H,W,C/2] and mfm_c the saved Split input
 *   [N,H,W,C]; the gradient is first routed to the larger half, ties half each (backward of torch.max + Split, lightcnn.py:48-62).
 *   out_pair != 0: z_out is written as a pair tensor (rows of [C bf16 hi | C bf16 lo], XFRB_IMPL_BF16X2) - the A operand of the
 *   kind::f16 dgrad that follows (xfrb_dgrad_plain with impl 5). */
typedef struct XfrbPriorEntry {
    int row;               /* gradient row that takes the prior at this firing, -1: none */
    int probe_row;         /* -1: no probe */
    long long elem;        /* one-element prior: flattened [H,W,C] index (when tensor == NULL) */
    long long probe_elem;
    const float* tensor;   /* full prior tensor [H*W*C] or NULL */
    float val;
    float pad_;
    long long pad2_;
} XfrbPriorEntry;          /* 48 bytes */
int xfrb_hook(const float* z_in, int up, int zc, const float* z_in2, int k2, int c2, float pre_scale, const float* s0, int c0,
              const float* s1, const float* s2, int c2s, const float* bn, const float* prior, int prior_row, long long prior_elem,
              float prior_val, float* P_out, float* z_out, int recipe, int affine, int relu_or_maxpool, int mode, int post_mask,
              int post_scale_row, int pre_scale_row, int J, int N, int H, int W, int C, float eps, const void* prior_entry,
              float* probe_out, int chain, const int* row_start, int k, const float* mfm_c, int out_pair, void* stream);
/* seed[j,:] = Pn[j,:] @ W2[j % N]  (Pn [J,Ccls], W2 [N,Ccls,D]) */
int xfrb_head_seed(const float* Pn, const float* W2, int Ccls, int D, int J, int N, float* seed, void* stream);
/* Jacobian of F.normalize (resnet.py:250): gout = (gin - xn*<xn,gin>)/nrm, rows of length D <= 1024 */
int xfrb_normalize_bwd(const float* gin, const float* xn, const float* nrm, float* gout, int J, int N, int D, void* stream);
/* MaxPool2d(3,2,pool_pad) backward alone: g [J,56,56,64] -> out [J,112,112,64]; arg-max from the bytes xfrb_stem_fwd recorded
 * (mp_arg [N,56,56,64]) or, when mp_arg is NULL, recomputed from relu(bn(o)) */
int xfrb_maxpool_bwd(const float* g, const float* o, const float* bn, float* out, const unsigned char* mp_arg, int J, int N, int pool_pad,
                     void* stream);
/* weighted_subtree_ebp layer score (whitebox.py:687-696): max / first argmax over n elements of m*(-gneg),
 * m = (gate >= 0) if gate_ge0 else (gate < 0) */
int xfrb_subtree_score(const float* gate, const float* gneg, int gate_ge0, long long n, float* score, long long* arg, void* stream);

/* ---- Light-CNN-29v2 (reference python/xfr/models/lightcnn.py:216-275, plugin whitebox.py:113-159) ----
 * The net has no ReLU / BatchNorm: every `mfm` is Conv2d(in, 2*out) -> Split -> torch.max (lightcnn.py:48-62).  A conv output
 * is stored as c [rows][2*Cp]: first Split half in columns [0,Cp), second in [Cp,2*Cp); Cp = `out` padded to the GEMM tile
 * granularity (48 -> 64, 96 -> 128), padded columns are exact zeros.  The backward sweep is issued firing by firing through
 * xfrb_hook (recipes 0, 3, 6, 7) with xfrb_dgrad_plain for the W+ transposed convs (xfr_b200/lightcnn.py). */

/* out [N,H,W,Cout] = conv_B(inp [N,H,W,Cin]) + bias.  B [Cout][R*R*Cin] K-major ((r,s,ci) order, weight planes per `impl`),
 * R = 1|3, stride 1, pad R/2; also the fc layer (H = W = 1).  positive = 1: B is relu(W) and inp >= 0 (two-pass plan under
 * impl 1).  Replaces mfm.filter's forward (lightcnn.py:59) in the 'activation' / 'positive_activation' passes. */
int xfrb_conv_bias(const float* inp, const float* B, const float* bias, float* out, int N, int H, int W, int Cin, int Cout, int R,
                   int positive, int impl, void* stream);
/* conv1 = Conv2d(1, 96, 5, 1, 2) (lightcnn.py:219): x [N,H,W] -> c [N,H,W,C2]; Wt [25][C2] tap-major in the padded column
 * layout; cpos (may be NULL) = conv_relu(W)(relu(x)) + bpos, the positive-pass twin (X of the Split hook = P[-2]'s X). */
int xfrb_lc_conv1(const float* x, const float* Wt, const float* b, const float* bpos, float* c, float* cpos, int N, int H, int W,
                  int C2, void* stream);
/* m = max(c[:, :Cp], c[:, Cp:]) [rows,Cp] (mfm.forward, lightcnn.py:58-62); y (may be NULL) = m + res (resblock Add,
 * lightcnn.py:84-88); relu_out (may be NULL) = relu(y if y else m). */
int xfrb_mfm_fwd(const float* c, const float* res, float* m, float* y, float* relu_out, long long rows, int Cp, void* stream);
/* autograd of torch.max(a, b) + Split: g [rows,Cp], c [rows_saved,2*Cp] (row r reads r % rows_saved) -> z [rows,2*Cp];
 * the larger branch takes g, exact ties take g/2 each. */
int xfrb_mfm_bwd(const float* g, const float* c, float* z, long long rows, long long rows_saved, int Cp, void* stream);
/* p = maxpool2(m) + avgpool2(m) (lightcnn.py:252-269); ppos (may be NULL) = the same on relu(m): the positive-pass value */
int xfrb_pool2_fwd(const float* m, float* p, float* ppos, int N, int H, int W, int C, void* stream);
/* its backward: gm [J,H,W,C] = MaxPool2d backward (first maximum wins, as torch) + AvgPool2d backward of g [J,H/2,W/2,C] */
int xfrb_pool2_bwd(const float* g, const float* m, float* gm, int J, int N, int H, int W, int C, void* stream);
/* out = relu(in), n floats (A operand of a positive-pass GEMM) */
int xfrb_relu(const float* in, float* out, long long n, void* stream);
/* chansum [J,HW] = sum_c P2 [J,HW,C]; sums [J] (double) = total per row  (whitebox.py:499, 524-525) */
int xfrb_chansum(const float* P2, float* chansum, double* sums, int J, int HW, int C, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* XFRB_H_ */
