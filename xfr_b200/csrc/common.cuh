// Shared device helpers: the excitation-backprop hook algebra used by every epilogue.
// Mirrors reference python/xfr/models/whitebox.py:381-430 (_backward_ebp, no prior set).
#pragma once
#include <cuda_bf16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#define XFRB_MODE_AWP 0
#define XFRB_MODE_ALL 1
#define XFRB_MODE_AFFINEONLY 2
#define XFRB_MODE_NONE 3        /* no hook fires: plain (true-gradient) backprop; used by weighted_subtree_ebp */

namespace xfrb {

// cudaFuncSetAttribute / occupancy answers are per DEVICE: "done once" caches are indexed by the current device ordinal
// (one process may drive several GPUs, e.g. the reference's multi-GPU eval driver with one Whitebox per cuda:%d).
constexpr int XFRB_MAX_DEV = 64;
inline int current_device_slot() {
    int d = 0;
    cudaGetDevice(&d);
    return (d >= 0 && d < XFRB_MAX_DEV) ? d : 0;
}

void set_error(const char* what, cudaError_t e);
int check_launch(const char* what);

__device__ __forceinline__ float relu(float v) { return fmaxf(v, 0.f); }

// One hook firing: zh = relu(z); p = a*zh; return p/(x+eps) | zh | z  depending on mode/kind.
template <bool AFFINE>
__device__ __forceinline__ float hook(float a, float x, float z, int mode, float eps) {
    if (mode == XFRB_MODE_NONE) return z;
    float zh = fmaxf(z, 0.f);
    // p/(x+eps): x + eps >= 1e-16 is a normal positive float, so the 2-ulp reciprocal path (MUFU.RCP + FMUL) is safe;
    // an IEEE division costs ~10 issue slots per hook and the epilogues are issue-bound (profiles/r1_notes.md).
    if (AFFINE || mode == XFRB_MODE_ALL) return __fdividef(__fmul_rn(a, zh), __fadd_rn(x, eps));
    return mode == XFRB_MODE_AWP ? zh : z;
}

struct BnC {  // per-channel BatchNorm constants (see include/xfrb.h)
    float alpha, beta, sp, tp;
};

__device__ __forceinline__ float bn_act(float o, const BnC& b) {  // a = relu(bn(o))
    return fmaxf(__fadd_rn(__fmul_rn(o, b.alpha), b.beta), 0.f);
}

// Chain at an inner activation a = relu(bn(o)) (ReLU hook, Conv2d hook, ReLU bwd, BN bwd, BN hook).
__device__ __forceinline__ float mid_chain(float z, float o, float xr, const BnC& b, int mode, float eps) {
    float a = bn_act(o, b);
    float ro = fmaxf(o, 0.f);
    float xrelu = fmaxf(__fadd_rn(__fmul_rn(ro, b.sp), b.tp), 0.f);
    z = hook<false>(a, xrelu, z, mode, eps);
    z = hook<true>(a, a, z, mode, eps);
    z = a > 0.f ? z : 0.f;
    z = __fmul_rn(z, b.sp);
    return hook<true>(ro, xr, z, mode, eps);
}

// Chain on a block output `out` + start of that block's main path.
// hooks & 3: 1 = [affine], 2 = [affine, non-affine], 3 = [affine, affine]   (hooks that follow the block's ReLU hook)
// hooks & 4: the residual sum is a plain function (VGGFace2 ResNet-50, resnet50_128.py:187): no Add hook, and the X of the
//            block ReLU sums positive-pass values, relu(BN+(relu(o3)) + res) with res = the positive-pass shortcut;
//            otherwise (STR ResNet, resnet.py:104-149) the Add module's inputs were overridden by A and its slot-0 hook
//            carries the residual's (A, X) (late-binding closure, whitebox.py:379-432).
__device__ __forceinline__ void join_chain(float z, float out, float o3, float xr3, float res, const BnC& b,
                                           int hooks, int mode, float eps, float& g, float& y3) {
    const bool fn_add = (hooks & 4) != 0;
    const int chain = hooks & 3;
    float rres = fmaxf(res, 0.f);
    float xblk = out;
    if (mode == XFRB_MODE_ALL) {
        if (fn_add) xblk = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(fmaxf(o3, 0.f), b.sp), b.tp), res), 0.f);
        else xblk = fmaxf(__fadd_rn(fmaxf(__fadd_rn(__fmul_rn(o3, b.alpha), b.beta), 0.f), rres), 0.f);
    }
    z = hook<false>(out, xblk, z, mode, eps);
    z = hook<true>(out, out, z, mode, eps);
    if (chain == 2) z = hook<false>(out, out, z, mode, eps);
    else if (chain == 3) z = hook<true>(out, out, z, mode, eps);
    g = out > 0.f ? z : 0.f;
    float zz = fn_add ? g : hook<false>(rres, rres, g, mode, eps);
    zz = __fmul_rn(zz, b.sp);
    y3 = hook<true>(fmaxf(o3, 0.f), xr3, zz, mode, eps);
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }

// ---------------- "pair" tensors of the bf16x2 plan ----------------
// A tensor whose rows of C fp32 values x are stored as [C bf16 hi | C bf16 lo] with hi = bf16(x), lo = bf16(x - hi): the same
// 4*C bytes per row, 16 significant bits, and both halves are ready-made tcgen05 kind::f16 operands (one TMA box each).
__device__ __forceinline__ uint32_t bf16x2_bits(float a, float b) {
    __nv_bfloat162 h = __floats2bfloat162_rn(a, b);           // .x (low 16 bits, lower address) = a
    return *reinterpret_cast<uint32_t*>(&h);
}
__device__ __forceinline__ void st_pair4(float* t, size_t row, int C, int c, const float v[4]) {
    float hf[4], lf[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        hf[e] = __bfloat162float(__float2bfloat16_rn(v[e]));
        lf[e] = v[e] - hf[e];                                  // exact in fp32
    }
    uint8_t* p = reinterpret_cast<uint8_t*>(t) + row * (size_t)C * 4;
    *reinterpret_cast<uint2*>(p + 2 * c) = make_uint2(bf16x2_bits(hf[0], hf[1]), bf16x2_bits(hf[2], hf[3]));
    *reinterpret_cast<uint2*>(p + 2 * (size_t)C + 2 * c) = make_uint2(bf16x2_bits(lf[0], lf[1]), bf16x2_bits(lf[2], lf[3]));
}
__device__ __forceinline__ void ld_pair4(const float* t, size_t row, int C, int c, float v[4]) {
    const uint8_t* p = reinterpret_cast<const uint8_t*>(t) + row * (size_t)C * 4;
    const uint2 h = *reinterpret_cast<const uint2*>(p + 2 * c), l = *reinterpret_cast<const uint2*>(p + 2 * (size_t)C + 2 * c);
    v[0] = __uint_as_float(h.x << 16) + __uint_as_float(l.x << 16);
    v[1] = __uint_as_float(h.x & 0xFFFF0000u) + __uint_as_float(l.x & 0xFFFF0000u);
    v[2] = __uint_as_float(h.y << 16) + __uint_as_float(l.y << 16);
    v[3] = __uint_as_float(h.y & 0xFFFF0000u) + __uint_as_float(l.y & 0xFFFF0000u);
}

// ---------------- epilogue parameter block shared by the SIMT and tcgen05 GEMMs ----------------
enum EpiKind { EPI_PLAIN = 0, EPI_FWD_DUAL = 1, EPI_MID = 2, EPI_JOIN = 3 };

struct EpiParams {
    int kind;
    int M;        // rows of the GEMM (J*H*W or N*H*W)
    int Ms;       // rows of the saved tensors (N*H*W); row m reads m % Ms
    int C;        // channels of the output tensors (Cout for FWD_DUAL, GEMM N otherwise)
    int mode, hooks;
    float eps;
    const float* bias;   // FWD_DUAL: [2C] tile order ; PLAIN: optional [N] bias
    const float* bn;     // [4][C]
    const float* res;    // FWD_DUAL: residual [Ms, res_c] ; JOIN: block residual (MODE_ALL only)
    int res_c;
    const float* o;      // MID: o ; JOIN: o3
    const float* xr;     // MID: xr ; JOIN: xr3
    const float* outp;   // JOIN: previous block output
    const float* g_res;  // JOIN: residual-path gradient [M, C] ; PLAIN: optional tensor to accumulate onto
    float* out0;         // PLAIN: z ; FWD_DUAL: o ; MID: y_out ; JOIN: g_out
    float* out1;         // FWD_DUAL: xr ; JOIN: y3_out
    float* out2;         // FWD_DUAL: act
    float* out3;         // FWD_DUAL, bf16x2 plan: optional fp32 copy of act (act itself is then a bf16 pair tensor)
};

// 4 consecutive channels (c..c+3) of one row.  acc = true/only accumulator, accp = positive twin (FWD_DUAL).
__device__ __forceinline__ void epilogue4(const EpiParams& P, int m, int c, float4 acc, float4 accp,
                                          float4 bias_t, float4 bias_p) {
    const size_t off = (size_t)m * P.C + c;
    if (P.kind == EPI_PLAIN) {
        if (P.g_res != nullptr) {      // accumulate into an existing tensor (second dgrad of a projection block)
            float4 t = ld4(P.g_res + off);
            acc.x += t.x; acc.y += t.y; acc.z += t.z; acc.w += t.w;
        }
        st4(P.out0 + off, acc);
        return;
    }
    const float* bnp = P.bn + c;
    float4 al = ld4(bnp), be = ld4(bnp + P.C), sp = ld4(bnp + 2 * P.C), tp = ld4(bnp + 3 * P.C);
    BnC b[4] = {{al.x, be.x, sp.x, tp.x}, {al.y, be.y, sp.y, tp.y}, {al.z, be.z, sp.z, tp.z}, {al.w, be.w, sp.w, tp.w}};
    float a[4] = {acc.x, acc.y, acc.z, acc.w};
    if (P.kind == EPI_FWD_DUAL) {
        float ap[4] = {accp.x, accp.y, accp.z, accp.w};
        float bt[4] = {bias_t.x, bias_t.y, bias_t.z, bias_t.w};
        float bp[4] = {bias_p.x, bias_p.y, bias_p.z, bias_p.w};
        float r[4] = {0.f, 0.f, 0.f, 0.f};
        if (P.res != nullptr && c < P.res_c) {
            float4 rv = ld4(P.res + (size_t)m * P.res_c + c);
            r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
        }
        float o[4], x[4], act[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            o[i] = __fadd_rn(a[i], bt[i]);
            x[i] = fmaxf(__fadd_rn(ap[i], bp[i]), 0.f);
            act[i] = __fadd_rn(__fadd_rn(__fmul_rn(o[i], b[i].alpha), b[i].beta), r[i]);
            if (!(P.hooks & 1)) act[i] = fmaxf(act[i], 0.f);      // hooks bit 0: emit bn(o) without the ReLU (projection shortcut)
        }
        st4(P.out0 + off, make_float4(o[0], o[1], o[2], o[3]));
        st4(P.out1 + off, make_float4(x[0], x[1], x[2], x[3]));
        st4(P.out2 + off, make_float4(act[0], act[1], act[2], act[3]));
        return;
    }
    const int ms = m % P.Ms;
    const size_t offs = (size_t)ms * P.C + c;
    float4 ov = ld4(P.o + offs), xv = ld4(P.xr + offs);
    float o[4] = {ov.x, ov.y, ov.z, ov.w}, x[4] = {xv.x, xv.y, xv.z, xv.w};
    if (P.kind == EPI_MID) {
        float y[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) y[i] = mid_chain(a[i], o[i], x[i], b[i], P.mode, P.eps);
        st4(P.out0 + off, make_float4(y[0], y[1], y[2], y[3]));
        return;
    }
    // EPI_JOIN
    float4 gv = ld4(P.g_res + off), uv = ld4(P.outp + offs);
    float gr[4] = {gv.x, gv.y, gv.z, gv.w}, u[4] = {uv.x, uv.y, uv.z, uv.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (P.mode == XFRB_MODE_ALL && P.res != nullptr && c < P.res_c) {
        float4 rv = ld4(P.res + (size_t)ms * P.res_c + c);
        r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
    }
    float g[4], y3[4];
#pragma unroll
    for (int i = 0; i < 4; ++i)
        join_chain(__fadd_rn(a[i], gr[i]), u[i], o[i], x[i], r[i], b[i], P.hooks, P.mode, P.eps, g[i], y3[i]);
    st4(P.out0 + off, make_float4(g[0], g[1], g[2], g[3]));
    st4(P.out1 + off, make_float4(y3[0], y3[1], y3[2], y3[3]));
}

// geometry of the implicit GEMM:  A[m, (r,s,ci)] = in[n, h+r-pad, w+s-pad, ci]
struct ConvGeom {
    int H, W, Cin;   // spatial size (stride 1, same in/out) and channels of the A operand
    int R;           // 1 or 3
    int K;           // R*R*Cin
    int Nn;          // GEMM N (rows of B)
};

struct PriorEntry {                        // 48 bytes, mirrored by xfr_b200/generic.py PRIOR_DTYPE
    int row;                               // gradient row that takes the prior at this firing, -1: none
    int probe_row;                         // -1: no probe
    long long elem;                        // one-element prior: flattened [H,W,C] index (used when tensor == nullptr)
    long long probe_elem;
    const float* tensor;                   // full prior tensor [H*W*C] or nullptr
    float val;
    float pad_;
    long long pad2_;
};
// arguments of the generic single-hook kernel (xfrb_hook): one _backward_ebp firing with optional prior and P recording
struct HookArgs {
    const float* z_in;  int up;            // incoming gradient [J,H/up,W/up,zc]; lands on pixels divisible by `up`
    int zc;                                // channels of z_in (>= C: only the first C are read; ConcatChannels slice)
    const float* z_in2; int k2, c2;        // optional addend [J,H/k2,W/k2,c2], replicated over k2 x k2 and divided (AvgPool bwd)
    float pre_scale;                       // z *= pre_scale before the hook (Multiply backward)
    const float* s0; const float* s1; const float* s2;   // recipe sources (saved tensors, sample = j % N)
    int c0, c2s;                           // channel counts of s0 / s2 when narrower than C (zero beyond)
    const float* bn;                       // [4][C] of the recipe's BatchNorm
    const float* prior;                    // full prior tensor of row prior_row ([H*W*C]) or null
    float* P_out;                          // records p (or, in MODE_NONE, the incoming gradient) [J,H,W,C]; may be null
    float* z_out;                          // hook return value after the post ops [J,H,W,C]; may be null
    int recipe, affine, relu_or_maxpool, mode, post_mask, post_scale_row;
    int J, N, H, W, C;
    float eps;
    int prior_row; long long prior_elem; float prior_val;   // one-element prior (layerwise 'elementwise'); prior_row < 0: none
    int pre_scale_row;                     // >= 0: z *= bn[pre_scale_row][c] before the hook (BatchNorm backward)
    // device-resident prior of this firing (one PriorEntry, include/xfrb.h XfrbPriorEntry): when set it REPLACES prior / prior_row /
    // prior_elem / prior_val above, so a captured CUDA graph of a sweep can be replayed with different priors
    const PriorEntry* ptab;
    float* probe_out;                      // with ptab: p of element ptab->probe_elem of row ptab->probe_row is also written here
    // Light-CNN: z_in is the gradient at an MFM OUTPUT [J,H,W,C/2] and mfm_c the saved Split input [N,H,W,C]: the firing (a Split
    // hook on c) first routes the gradient to the larger half (ties: half each) - the backward of torch.max + Split, lightcnn.py:48-62
    const float* mfm_c;
    int out_pair;                          // z_out is a pair tensor (bf16 hi | lo rows): the A operand of a kind::f16 dgrad (bf16x2 plan)
};
constexpr int XFRB_MAX_CHAIN = 6;
struct HookChain {                         // consecutive firings on one [J,H,W,C] tensor fused into one launch (stages.cu hook_kernel)
    HookArgs a[XFRB_MAX_CHAIN];
    int n;
    int k0;                                // firing index of link 0 (compared with row_start)
    const int* row_start;                  // [J] or null: first firing each gradient row takes part in (zero-seeded prior sweeps)
};
cudaError_t launch_hook(const HookArgs& a, cudaStream_t st);
cudaError_t launch_hook_chain(const HookChain& ch, cudaStream_t st);

// arguments of the unfused block-boundary kernel (xfrb_join)
struct JoinArgs {
    const float* zmain; int up;
    const float* gres_lo; int gres_c, k;
    const float* out; const float* o3; const float* xr3; const float* bn3;
    const float* res; int res_c;
    float* g_out; float* y3_out;
    int J, N, H, W, C, hooks, mode; float eps;
    int y3_pair;                           // bf16x2 plan: y3_out is a pair tensor
};

cudaError_t launch_conv_simt(const float* A, const float* B, const ConvGeom& g, const EpiParams& ep, cudaStream_t st);
bool conv_tc_available();
double conv_tc_tile_geometry(int H, int W, int Nimg, int* bh, int* bimg);   // host: 4-D TMA box of a 3x3 conv (conv_tc.cu)
int conv_tc_set_cta2(int on);     // CTA-pair (cta_group::2) kernels on/off; returns the previous setting
int conv_tc_set_mc(int on);       // multicast-pair kernels (shared weight loads) on/off; returns the previous setting
// B: [planes][Nn][K] with the 3xTF32 (hi, lo) planes when split != 0 (pass plans 0-3: conv_tc.cu); tn: dual-pack tile
// width (FWD_DUAL) or a cap on BN
cudaError_t launch_conv_tc(const float* A, const float* B, const ConvGeom& g, const EpiParams& ep, int split, int tn,
                           cudaStream_t st);
bool conv_tc_cta2_enabled();
// bf16x2 plan (conv_tc_pair.cu): A is a pair tensor [M][Cin bf16 hi | Cin bf16 lo]; B bf16 [planes][Nn][K]: one plane of
// relu(W) for the W+ GEMMs (MID / JOIN / PLAIN), (hi, lo) planes of the dual pack for FWD_DUAL (the lo plane is read for the
// W half of each tile only).  Outputs that feed the next GEMM (act / y_out / y3_out) are written as pair tensors.
cudaError_t launch_conv_tc_pair(const float* A, const void* B, const ConvGeom& g, const EpiParams& ep, int tn, cudaStream_t st);

}  // namespace xfrb
