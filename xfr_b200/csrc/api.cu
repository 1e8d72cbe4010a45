// extern "C" surface of libxfr_b200.so (see include/xfrb.h for the contract of every call).
#include "../../include/xfrb.h"
#include "common.cuh"
#include <stdio.h>
#include <string.h>
#include <stdlib.h>

namespace xfrb {

static thread_local char g_err[512] = "";

void set_error(const char* what, cudaError_t e) {
    snprintf(g_err, sizeof(g_err), "%s: %s", what, e == cudaSuccess ? "invalid argument" : cudaGetErrorString(e));
}

// forward declarations of the stage launchers (stages.cu)
cudaError_t launch_stem_fwd(const float*, const float*, const float*, const float*, float*, float*, unsigned char*, int, int,
                            cudaStream_t);
cudaError_t launch_bn_hook(const float*, const float*, const float*, const float*, float*, size_t, size_t, int, int, int, float,
                           cudaStream_t);
cudaError_t launch_normalize_bwd(const float*, const float*, const float*, float*, int, int, int, cudaStream_t);
cudaError_t launch_maxpool_bwd(const float*, const float*, const float*, float*, const unsigned char*, int, int, int, cudaStream_t);
cudaError_t launch_subtree_score(const float*, const float*, int, size_t, float*, long long*, cudaStream_t);
cudaError_t launch_head_seed(const float*, const float*, int, int, int, int, float*, cudaStream_t);
cudaError_t launch_subsample2(const float*, float*, int, int, int, int, cudaStream_t);
cudaError_t launch_to_pair(const float*, float*, size_t, int, int, cudaStream_t);
cudaError_t launch_avgpool2(const float*, float*, int, int, int, int, cudaStream_t);
cudaError_t launch_avgpool7(const float*, float*, int, int, cudaStream_t);
cudaError_t launch_head_norm(const float*, int, float*, float*, float*, float*, float*, int, cudaStream_t);
cudaError_t launch_head_bwd_a(const float*, const float*, int, const float*, const float*, const float*, float*, int, int,
                              int, float, cudaStream_t);
cudaError_t launch_head_bwd_b(const float*, const float*, float*, int, int, int, int, float, cudaStream_t);
cudaError_t launch_join(const JoinArgs&, cudaStream_t);
cudaError_t launch_ds_res(const float*, const float*, float*, int, int, int, int, int, int, int, float, cudaStream_t);
cudaError_t launch_stem_bwd(const float*, const float*, const float*, const float*, const float*, float*, float*, float*,
                            double*, const unsigned char*, int, int, int, float, int, cudaStream_t);
cudaError_t launch_contrast(const float*, const double*, const float*, float*, int, int, int, cudaStream_t);
cudaError_t launch_trunc_threshold(const float*, const double*, float, float*, int, size_t, cudaStream_t);
cudaError_t launch_saliency_post(const float*, float*, int, int, int, float, cudaStream_t);
cudaError_t launch_cubic_zoom(const float*, float*, int, int, int, int, int, int, cudaStream_t);
cudaError_t launch_twin_blends(const double*, const double*, const double*, const double*, const double*, float*, int, int, int, int,
                               int, cudaStream_t);
// lightcnn.cu
cudaError_t launch_lc_conv1(const float*, const float*, const float*, const float*, float*, float*, int, int, int, int, cudaStream_t);
cudaError_t launch_mfm_fwd(const float*, const float*, float*, float*, float*, size_t, int, cudaStream_t);
cudaError_t launch_mfm_bwd(const float*, const float*, float*, size_t, size_t, int, cudaStream_t);
cudaError_t launch_pool2_fwd(const float*, float*, float*, int, int, int, int, cudaStream_t);
cudaError_t launch_pool2_bwd(const float*, const float*, float*, int, int, int, int, int, cudaStream_t);
cudaError_t launch_relu(const float*, float*, size_t, cudaStream_t);
cudaError_t launch_chansum(const float*, float*, double*, int, int, int, cudaStream_t);

static int finish(const char* what, cudaError_t e) {
    if (e != cudaSuccess) {
        set_error(what, e);
        cudaGetLastError();          // a rejected launch / attribute call must not linger as the caller's next CUDA error
        return (int)e;
    }
    return 0;
}

// positive_weights: the B operand is relu(W) and the A operand is non-negative (a W+ GEMM of excitation backprop): under
// XFRB_IMPL_TF32X3 it runs the two-pass plan (see conv_tc.cu); signed GEMMs keep all three passes.
static cudaError_t run_gemm(const float* A, const float* B, const ConvGeom& g, const EpiParams& ep, int impl, cudaStream_t st,
                            int tn = 0, bool positive_weights = false) {
    if (impl == XFRB_IMPL_FP32) return launch_conv_simt(A, B, g, ep, st);
    if (impl == XFRB_IMPL_BF16X2) return launch_conv_tc_pair(A, B, g, ep, tn, st);     // A: pair tensor, B: bf16 planes
    int split;
    if (impl == XFRB_IMPL_TF32) split = 0;
    else if (impl == XFRB_IMPL_TF32X3_FULL) split = 1;
    else if (impl == XFRB_IMPL_TF32X2) split = 2;
    else if (impl == XFRB_IMPL_TF32X3) split = ep.kind == EPI_FWD_DUAL ? 3 : (ep.kind == EPI_MID || ep.kind == EPI_JOIN || positive_weights) ? 2 : 1;
    else return cudaErrorInvalidValue;
    return launch_conv_tc(A, B, g, ep, split, tn, st);
}

}  // namespace xfrb

using namespace xfrb;

extern "C" {

int xfrb_version(void) { return 1; }
const char* xfrb_last_error(void) { return g_err; }

int xfrb_device_ok(void) {
    int dev = 0;
    cudaDeviceProp p;
    if (cudaGetDevice(&dev) != cudaSuccess || cudaGetDeviceProperties(&p, dev) != cudaSuccess) return 0;
    return p.major == 10 ? 1 : 0;
}

int xfrb_impl_available(int impl) {
    if (impl == XFRB_IMPL_FP32) return 1;
    if (impl == XFRB_IMPL_TF32X3 || impl == XFRB_IMPL_TF32 || impl == XFRB_IMPL_TF32X3_FULL || impl == XFRB_IMPL_TF32X2 ||
        impl == XFRB_IMPL_BF16X2)
        return conv_tc_available() ? 1 : 0;
    return 0;
}

double xfrb_tile_geometry(int H, int W, int Nimg, int* bh, int* bimg) {
    if (H < 1 || W < 1 || W > 128 || Nimg < 1 || bh == nullptr || bimg == nullptr) return -1.0;
    return conv_tc_tile_geometry(H, W, Nimg, bh, bimg);
}
int xfrb_set_cta_pairs(int on) { return conv_tc_set_cta2(on); }
int xfrb_set_multicast_pairs(int on) { return conv_tc_set_mc(on); }

int xfrb_stem_fwd(const float* x, const float* W, const float* b, const float* bn, float* o, float* mp, unsigned char* mp_arg,
                  int N, int pool_pad, void* stream) {
    if (pool_pad != 0 && pool_pad != 1) return finish("xfrb_stem_fwd", cudaErrorInvalidValue);
    return finish("xfrb_stem_fwd", launch_stem_fwd(x, W, b, bn, o, mp, mp_arg, N, pool_pad, (cudaStream_t)stream));
}

int xfrb_subsample2(const float* u, float* out, int N, int H, int W, int C, void* stream) {
    if ((H | W) & 1 || C % 4) return finish("xfrb_subsample2", cudaErrorInvalidValue);
    return finish("xfrb_subsample2", launch_subsample2(u, out, N, H, W, C, (cudaStream_t)stream));
}

int xfrb_avgpool2(const float* u, float* out, int N, int H, int W, int C, void* stream) {
    if ((H | W) & 1 || C % 4) return finish("xfrb_avgpool2", cudaErrorInvalidValue);
    return finish("xfrb_avgpool2", launch_avgpool2(u, out, N, H, W, C, (cudaStream_t)stream));
}

int xfrb_to_pair(const float* in, float* out, long long rows, int C, int inverse, void* stream) {
    if (C % 4 || rows < 0) return finish("xfrb_to_pair", cudaErrorInvalidValue);
    return finish("xfrb_to_pair", launch_to_pair(in, out, (size_t)rows, C, inverse, (cudaStream_t)stream));
}

int xfrb_conv_dual(const float* inp, const float* Bf, const float* bias, const float* bn, const float* res, int res_c,
                   float* o, float* xr, float* act, float* act_f32, int N, int H, int W, int Cin, int Cout, int R, int tn,
                   int relu_act, int impl, void* stream) {
    if ((R != 1 && R != 3) || (tn != 128 && tn != 256) || (impl == XFRB_IMPL_FP32 && tn != 128) || (2 * Cout) % tn ||
        (res && res_c % 4) || (impl != XFRB_IMPL_BF16X2 && (act_f32 != nullptr || act == nullptr)))
        return finish("xfrb_conv_dual", cudaErrorInvalidValue);
    ConvGeom g{H, W, Cin, R, R * R * Cin, 2 * Cout};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_FWD_DUAL;
    ep.M = ep.Ms = N * H * W;
    ep.C = Cout;
    ep.bias = bias; ep.bn = bn; ep.res = res; ep.res_c = res_c;
    ep.hooks = relu_act ? 0 : 1;
    ep.out0 = o; ep.out1 = xr; ep.out2 = act; ep.out3 = act_f32;
    return finish("xfrb_conv_dual", run_gemm(inp, Bf, g, ep, impl, (cudaStream_t)stream, tn));
}

int xfrb_head_fwd(const float* u, const float* B1, const float* bias1, int tn, float* scratch, float* v, float* f1, float* f1p,
                  float* xn, float* nrm, float* xmul, int N, int impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_avgpool7(u, v, N, 2048, st);
    if (e != cudaSuccess) return finish("xfrb_head_fwd/avgpool", e);
    ConvGeom g{1, 1, 2048, 1, 2048, 1024};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = N;
    ep.C = 1024;
    ep.bias = bias1;
    ep.out0 = scratch;
    e = run_gemm(v, B1, g, ep, impl, st, tn);
    if (e != cudaSuccess) return finish("xfrb_head_fwd/fc1", e);
    return finish("xfrb_head_fwd/norm", launch_head_norm(scratch, tn, f1, f1p, xn, nrm, xmul, N, st));
}

int xfrb_head_bwd(const float* Pn, const float* W2, int C, const float* W1pT, const float* v, const float* f1p, const float* xn,
                  const float* nrm, float* scratch, float* g_out, int J, int N, int mode, float eps, int impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_head_bwd_a(Pn, W2, C, f1p, xn, nrm, scratch, J, N, mode, eps, st);
    if (e != cudaSuccess) return finish("xfrb_head_bwd/a", e);
    // z [J,2048] = scratch [J,512] @ relu(W1) ; staged in the tail of g_out, then expanded in place order-safely:
    // g_out is [J,49,2048]; z is written to a separate region at the END of g_out (last J*2048 floats) only if J*2048
    // <= room; to stay simple and race-free we use a dedicated slice: scratch + J*512 must hold J*2048 floats.
    float* z = scratch + (size_t)J * 512;
    ConvGeom g{1, 1, 512, 1, 512, 2048};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = J;
    ep.C = 2048;
    ep.out0 = z;
    e = run_gemm(scratch, W1pT, g, ep, impl, st, 0, true);
    if (e != cudaSuccess) return finish("xfrb_head_bwd/fc1", e);
    return finish("xfrb_head_bwd/b", launch_head_bwd_b(z, v, g_out, J, N, 2048, mode, eps, st));
}

int xfrb_dgrad_mid(const float* y, const float* Bd, const float* o, const float* xr, const float* bn, float* y_out, int J, int N,
                   int H, int W, int Cin, int Cout, int R, int mode, float eps, int impl, void* stream) {
    if (R != 1 && R != 3) return finish("xfrb_dgrad_mid", cudaErrorInvalidValue);
    ConvGeom g{H, W, Cout, R, R * R * Cout, Cin};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_MID;
    ep.M = J * H * W; ep.Ms = N * H * W;
    ep.C = Cin;
    ep.mode = mode; ep.eps = eps;
    ep.bn = bn; ep.o = o; ep.xr = xr;
    ep.out0 = y_out;
    return finish("xfrb_dgrad_mid", run_gemm(y, Bd, g, ep, impl, (cudaStream_t)stream));
}

int xfrb_dgrad_plain(const float* y, const float* Bd, float* z_out, int J, int H, int W, int Cin, int Cout, int R, int accumulate,
                     int impl, void* stream) {
    if (R != 1 && R != 3) return finish("xfrb_dgrad_plain", cudaErrorInvalidValue);
    ConvGeom g{H, W, Cout, R, R * R * Cout, Cin};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = J * H * W;
    ep.C = Cin;
    ep.out0 = z_out;
    ep.g_res = accumulate ? z_out : nullptr;
    return finish("xfrb_dgrad_plain", run_gemm(y, Bd, g, ep, impl, (cudaStream_t)stream, 0, true));
}

int xfrb_dgrad_join(const float* y1, const float* Bd, const float* g_res, const float* out, const float* o3, const float* xr3,
                    const float* bn3, const float* res, int res_c, float* g_out, float* y3_out, int J, int N, int H, int W,
                    int Cin, int Cout, int hooks, int mode, float eps, int impl, void* stream) {
    ConvGeom g{H, W, Cout, 1, Cout, Cin};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_JOIN;
    ep.M = J * H * W; ep.Ms = N * H * W;
    ep.C = Cin;
    ep.mode = mode; ep.hooks = hooks; ep.eps = eps;
    ep.bn = bn3; ep.o = o3; ep.xr = xr3; ep.outp = out; ep.g_res = g_res; ep.res = res; ep.res_c = res_c;
    ep.out0 = g_out; ep.out1 = y3_out;
    return finish("xfrb_dgrad_join", run_gemm(y1, Bd, g, ep, impl, (cudaStream_t)stream));
}

int xfrb_join(const float* zmain, int up, const float* gres_lo, int gres_c, int k, const float* out, const float* o3,
              const float* xr3, const float* bn3, const float* res, int res_c, float* g_out, float* y3_out, int J, int N, int H,
              int W, int C, int hooks, int mode, float eps, int y3_pair, void* stream) {
    if (C % 4 || (gres_lo && gres_c % 4) || (res && res_c % 4) || up < 1 || k < 1) return finish("xfrb_join", cudaErrorInvalidValue);
    JoinArgs a{zmain, up, gres_lo, gres_c, k, out, o3, xr3, bn3, res, res_c, g_out, y3_out, J, N, H, W, C, hooks, mode, eps, y3_pair};
    return finish("xfrb_join", launch_join(a, (cudaStream_t)stream));
}

int xfrb_ds_res(const float* g, const float* ap, float* gres_lo, int J, int N, int H, int W, int C, int Cr, int mode, float eps,
                void* stream) {
    if (C % 4 || Cr % 4) return finish("xfrb_ds_res", cudaErrorInvalidValue);
    return finish("xfrb_ds_res", launch_ds_res(g, ap, gres_lo, J, N, H, W, C, Cr, mode, eps, (cudaStream_t)stream));
}

int xfrb_stem_bwd(const float* zmain, const float* gres, const float* o, const float* mp, const float* bn, float* zc, float* P2,
                  float* chansum, double* sums, const unsigned char* mp_arg, int J, int N, int mode, float eps, int pool_pad,
                  void* stream) {
    if (pool_pad != 0 && pool_pad != 1) return finish("xfrb_stem_bwd", cudaErrorInvalidValue);
    return finish("xfrb_stem_bwd",
                  launch_stem_bwd(zmain, gres, o, mp, bn, zc, P2, chansum, sums, mp_arg, J, N, mode, eps, pool_pad,
                                  (cudaStream_t)stream));
}

int xfrb_bn_hook(const float* g, const float* o, const float* xr, const float* bn, float* y, int J, int N, int HW, int C, int kind,
                 int mode, float eps, void* stream) {
    if (C % 4) return finish("xfrb_bn_hook", cudaErrorInvalidValue);
    return finish("xfrb_bn_hook", launch_bn_hook(g, o, xr, bn, y, (size_t)(kind == 1 ? N : J) * HW, (size_t)N * HW, C, kind, mode, eps,
                                                 (cudaStream_t)stream));
}

int xfrb_hook(const float* z_in, int up, int zc, const float* z_in2, int k2, int c2, float pre_scale, const float* s0, int c0,
              const float* s1, const float* s2, int c2s, const float* bn, const float* prior, int prior_row, long long prior_elem,
              float prior_val, float* P_out, float* z_out, int recipe, int affine, int relu_or_maxpool, int mode, int post_mask,
              int post_scale_row, int pre_scale_row, int J, int N, int H, int W, int C, float eps, const void* prior_entry,
              float* probe_out, int chain, const int* row_start, int k, const float* mfm_c, int out_pair, void* stream) {
    static thread_local HookChain pending = {};      // links appended with chain = 1, launched by the call that passes chain = 2
    if (chain == 0 && pending.n != 0) { pending.n = 0; return finish("xfrb_hook", cudaErrorInvalidValue); }   // an unfinished chain
    if (chain < 0 || chain > 2 || pending.n >= XFRB_MAX_CHAIN) { pending.n = 0; return finish("xfrb_hook", cudaErrorInvalidValue); }
    if (up < 1 || k2 < 1 || recipe < 0 || recipe > 8 || ((post_scale_row >= 0 || pre_scale_row >= 0) && bn == nullptr))
        return finish("xfrb_hook", cudaErrorInvalidValue);
    HookArgs a;
    a.z_in = z_in; a.up = up; a.zc = zc; a.z_in2 = z_in2; a.k2 = k2; a.c2 = c2; a.pre_scale = pre_scale;
    a.s0 = s0; a.s1 = s1; a.s2 = s2; a.c0 = c0; a.c2s = c2s; a.bn = bn; a.prior = prior; a.P_out = P_out; a.z_out = z_out;
    a.recipe = recipe; a.affine = affine; a.relu_or_maxpool = relu_or_maxpool; a.mode = mode; a.post_mask = post_mask;
    a.post_scale_row = post_scale_row; a.J = J; a.N = N; a.H = H; a.W = W; a.C = C; a.eps = eps;
    a.prior_row = prior_row; a.prior_elem = prior_elem; a.prior_val = prior_val;
    a.pre_scale_row = pre_scale_row;
    a.ptab = static_cast<const PriorEntry*>(prior_entry); a.probe_out = probe_out; a.mfm_c = mfm_c; a.out_pair = out_pair;
    static_assert(sizeof(PriorEntry) == sizeof(XfrbPriorEntry), "include/xfrb.h XfrbPriorEntry mirrors PriorEntry");
    if (pending.n == 0) { pending.k0 = k; pending.row_start = row_start; }
    pending.a[pending.n++] = a;
    if (chain == 1) return 0;                        // deferred: nothing launched yet
    HookChain ch = pending;
    pending.n = 0;
    return finish("xfrb_hook", launch_hook_chain(ch, (cudaStream_t)stream));
}

int xfrb_head_seed(const float* Pn, const float* W2, int Ccls, int D, int J, int N, float* seed, void* stream) {
    return finish("xfrb_head_seed", launch_head_seed(Pn, W2, Ccls, D, J, N, seed, (cudaStream_t)stream));
}

int xfrb_normalize_bwd(const float* gin, const float* xn, const float* nrm, float* gout, int J, int N, int D, void* stream) {
    if (D > 1024) return finish("xfrb_normalize_bwd", cudaErrorInvalidValue);
    return finish("xfrb_normalize_bwd", launch_normalize_bwd(gin, xn, nrm, gout, J, N, D, (cudaStream_t)stream));
}

int xfrb_maxpool_bwd(const float* g, const float* o, const float* bn, float* out, const unsigned char* mp_arg, int J, int N, int pool_pad,
                     void* stream) {
    return finish("xfrb_maxpool_bwd", launch_maxpool_bwd(g, o, bn, out, mp_arg, J, N, pool_pad, (cudaStream_t)stream));
}

int xfrb_subtree_score(const float* gate, const float* gneg, int gate_ge0, long long n, float* score, long long* arg, void* stream) {
    return finish("xfrb_subtree_score", launch_subtree_score(gate, gneg, gate_ge0, (size_t)n, score, arg, (cudaStream_t)stream));
}

int xfrb_head_fwd_linear(const float* u, const float* Bfe, float* v, float* enc, int N, int C, int D, int impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_avgpool7(u, v, N, C, st);
    if (e != cudaSuccess) return finish("xfrb_head_fwd_linear/avgpool", e);
    ConvGeom g{1, 1, C, 1, C, D};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = N;
    ep.C = D;
    ep.out0 = enc;
    return finish("xfrb_head_fwd_linear/gemm", run_gemm(v, Bfe, g, ep, impl, st));
}

int xfrb_head_bwd_linear(const float* Pn, const float* W2, int Ccls, const float* BfeT, const float* v, float* scratch, float* g_out,
                         int J, int N, int C, int D, int mode, float eps, int impl, void* stream) {
    cudaStream_t st = (cudaStream_t)stream;
    cudaError_t e = launch_head_seed(Pn, W2, Ccls, D, J, N, scratch, st);
    if (e != cudaSuccess) return finish("xfrb_head_bwd_linear/seed", e);
    float* z = scratch + (size_t)J * D;
    ConvGeom g{1, 1, D, 1, D, C};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = J;
    ep.C = C;
    ep.out0 = z;
    e = run_gemm(scratch, BfeT, g, ep, impl, st, 0, true);
    if (e != cudaSuccess) return finish("xfrb_head_bwd_linear/gemm", e);
    return finish("xfrb_head_bwd_linear/b", launch_head_bwd_b(z, v, g_out, J, N, C, mode, eps, st));
}

int xfrb_contrast(const float* P2, const double* sums, const float* thr, float* out, int N, int HW, int C, void* stream) {
    return finish("xfrb_contrast", launch_contrast(P2, sums, thr, out, N, HW, C, (cudaStream_t)stream));
}

int xfrb_trunc_threshold(const float* P2, const double* sums, float percentile, float* thr, int N, long long per_sample,
                         void* stream) {
    return finish("xfrb_trunc_threshold", launch_trunc_threshold(P2, sums, percentile, thr, N, (size_t)per_sample, (cudaStream_t)stream));
}

int xfrb_saliency_post(const float* mwp, float* out, int B, int H, int W, float eps, void* stream) {
    if (H > 128 || W > 128) return finish("xfrb_saliency_post", cudaErrorInvalidValue);
    return finish("xfrb_saliency_post", launch_saliency_post(mwp, out, B, H, W, eps, (cudaStream_t)stream));
}

int xfrb_cubic_zoom(const float* in, float* out, int B, int h, int w, int oh, int ow, int normalize, void* stream) {
    if (h < 4 || w < 4 || h > 144 || w > 144 || oh < 1 || ow < 1) return finish("xfrb_cubic_zoom", cudaErrorInvalidValue);
    return finish("xfrb_cubic_zoom", launch_cubic_zoom(in, out, B, h, w, oh, ow, normalize, (cudaStream_t)stream));
}

int xfrb_twin_blends(const double* orig, const double* inp, const double* value, const double* thr, const double* masks, float* out,
                     int K, int C, int H, int W, int mask_f32, void* stream) {
    return finish("xfrb_twin_blends", launch_twin_blends(orig, inp, value, thr, masks, out, K, C, H, W, mask_f32, (cudaStream_t)stream));
}

/* ---- Light-CNN-29v2 ---- */

int xfrb_conv_bias(const float* inp, const float* B, const float* bias, float* out, int N, int H, int W, int Cin, int Cout, int R,
                   int positive, int impl, void* stream) {
    if (R != 1 && R != 3) return finish("xfrb_conv_bias", cudaErrorInvalidValue);
    ConvGeom g{H, W, Cin, R, R * R * Cin, Cout};
    EpiParams ep;
    memset(&ep, 0, sizeof(ep));
    ep.kind = EPI_PLAIN;
    ep.M = ep.Ms = N * H * W;
    ep.C = Cout;
    ep.bias = bias;
    ep.out0 = out;
    return finish("xfrb_conv_bias", run_gemm(inp, B, g, ep, impl, (cudaStream_t)stream, 0, positive != 0));
}

int xfrb_lc_conv1(const float* x, const float* Wt, const float* b, const float* bpos, float* c, float* cpos, int N, int H, int W,
                  int C2, void* stream) {
    return finish("xfrb_lc_conv1", launch_lc_conv1(x, Wt, b, bpos, c, cpos, N, H, W, C2, (cudaStream_t)stream));
}

int xfrb_mfm_fwd(const float* c, const float* res, float* m, float* y, float* relu_out, long long rows, int Cp, void* stream) {
    if (Cp % 4 || (y != nullptr && res == nullptr)) return finish("xfrb_mfm_fwd", cudaErrorInvalidValue);
    return finish("xfrb_mfm_fwd", launch_mfm_fwd(c, res, m, y, relu_out, (size_t)rows, Cp, (cudaStream_t)stream));
}

int xfrb_mfm_bwd(const float* g, const float* c, float* z, long long rows, long long rows_saved, int Cp, void* stream) {
    if (Cp % 4 || rows_saved <= 0 || rows % rows_saved) return finish("xfrb_mfm_bwd", cudaErrorInvalidValue);
    return finish("xfrb_mfm_bwd", launch_mfm_bwd(g, c, z, (size_t)rows, (size_t)rows_saved, Cp, (cudaStream_t)stream));
}

int xfrb_pool2_fwd(const float* m, float* p, float* ppos, int N, int H, int W, int C, void* stream) {
    if ((H | W) & 1 || C % 4) return finish("xfrb_pool2_fwd", cudaErrorInvalidValue);
    return finish("xfrb_pool2_fwd", launch_pool2_fwd(m, p, ppos, N, H, W, C, (cudaStream_t)stream));
}

int xfrb_pool2_bwd(const float* g, const float* m, float* gm, int J, int N, int H, int W, int C, void* stream) {
    if ((H | W) & 1 || C % 4 || N <= 0 || J % N) return finish("xfrb_pool2_bwd", cudaErrorInvalidValue);
    return finish("xfrb_pool2_bwd", launch_pool2_bwd(g, m, gm, J, N, H, W, C, (cudaStream_t)stream));
}

int xfrb_relu(const float* in, float* out, long long n, void* stream) {
    if (n % 4) return finish("xfrb_relu", cudaErrorInvalidValue);
    return finish("xfrb_relu", launch_relu(in, out, (size_t)n, (cudaStream_t)stream));
}

int xfrb_chansum(const float* P2, float* chansum, double* sums, int J, int HW, int C, void* stream) {
    return finish("xfrb_chansum", launch_chansum(P2, chansum, sums, J, HW, C, (cudaStream_t)stream));
}

}  // extern "C"
