// Light-CNN-29v2 stages that are not GEMMs (reference python/xfr/models/lightcnn.py:48-62, 216-275): the 5x5 single-channel
// stem conv, the max-feature-map (MFM) split/max and its backward, the maxpool2 + avgpool2 sum and its backward, and the
// channel sum of P[-2].  All tensors NHWC fp32; an MFM conv output `c` is stored as [rows][2*Cp] with the first Split
// half in columns [0, Cp) and the second in [Cp, 2*Cp) (Cp = channel count padded to a GEMM-friendly width; padded
// columns are exact zeros end to end because their weights and biases are zero).
#include "common.cuh"

namespace xfrb {

// ------------------------------------------------------------------ stem: Conv2d(1, 96, 5, 1, 2)
// x [N,H,W] (one channel) -> c [N,H,W,C2] = conv_W(x) + b ; cpos (optional) = conv_relu(W)(relu(x)) + bpos.
// Wt [25][C2] tap-major.  One thread = one pixel x 4 consecutive channels: the 25 window values sit in registers, the weights
// come from shared memory as float4 (one copy per CTA), so a tap costs one shared load per 4 (or 8, with the twin) FMAs.  The
// thread-per-(pixel, channel) form it replaces issued two global loads per FMA and ran 20x above its store bound (5.7 ms for 128
// probes, profiles/r2_notes.md).  Same accumulation order (taps row-major, fmaf chain), so the results are bit-identical.
template <bool TWIN>
__global__ void __launch_bounds__(256) lc_conv1_kernel(const float* __restrict__ x, const float* __restrict__ Wt,
                                                       const float* __restrict__ b, const float* __restrict__ bpos,
                                                       float* __restrict__ c, float* __restrict__ cpos, int H, int W, int C2,
                                                       unsigned total4) {
    extern __shared__ float4 w_s[];                                 // [25][C2/4]
    const int C4 = C2 / 4;
    for (int k = threadIdx.x; k < 25 * C4; k += blockDim.x) w_s[k] = reinterpret_cast<const float4*>(Wt)[k];
    __syncthreads();
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const int c4 = (int)(i % (unsigned)C4);
    unsigned p = i / (unsigned)C4;
    const int w = (int)(p % (unsigned)W); p /= (unsigned)W;
    const int h = (int)(p % (unsigned)H);
    const size_t n = p / (unsigned)H;
    const float* xi = x + n * H * W;
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, accp[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int r = 0; r < 5; ++r) {
        const int hh = h + r - 2;
        if (hh < 0 || hh >= H) continue;
#pragma unroll
        for (int s = 0; s < 5; ++s) {
            const int ww = w + s - 2;
            if (ww < 0 || ww >= W) continue;
            const float xv = __ldg(xi + (size_t)hh * W + ww);
            const float4 wv = w_s[(r * 5 + s) * C4 + c4];
            const float wq[4] = {wv.x, wv.y, wv.z, wv.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                acc[q] = fmaf(xv, wq[q], acc[q]);
                if (TWIN) accp[q] = fmaf(fmaxf(xv, 0.f), fmaxf(wq[q], 0.f), accp[q]);
            }
        }
    }
    const float4 bv = reinterpret_cast<const float4*>(b)[c4];
    reinterpret_cast<float4*>(c)[i] = make_float4(__fadd_rn(acc[0], bv.x), __fadd_rn(acc[1], bv.y), __fadd_rn(acc[2], bv.z), __fadd_rn(acc[3], bv.w));
    if (TWIN) {
        const float4 bq = reinterpret_cast<const float4*>(bpos)[c4];
        reinterpret_cast<float4*>(cpos)[i] = make_float4(__fadd_rn(accp[0], bq.x), __fadd_rn(accp[1], bq.y), __fadd_rn(accp[2], bq.z), __fadd_rn(accp[3], bq.w));
    }
}

cudaError_t launch_lc_conv1(const float* x, const float* Wt, const float* b, const float* bpos, float* c, float* cpos, int N,
                            int H, int W, int C2, cudaStream_t st) {
    size_t total4 = (size_t)N * H * W * (C2 / 4);
    if (C2 % 4 || total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // float4 channels, 32-bit index arithmetic
    const size_t smem = (size_t)25 * C2 * sizeof(float);
    if (smem > 48 * 1024) return cudaErrorInvalidValue;
    const unsigned grid = (unsigned)((total4 + 255) / 256);
    if (cpos != nullptr) lc_conv1_kernel<true><<<grid, 256, smem, st>>>(x, Wt, b, bpos, c, cpos, H, W, C2, (unsigned)total4);
    else lc_conv1_kernel<false><<<grid, 256, smem, st>>>(x, Wt, b, bpos, c, cpos, H, W, C2, (unsigned)total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ MFM forward: m = max(c[:, :Cp], c[:, Cp:]) (+ res)
// c [M, 2*Cp] -> m [M, Cp]; y (optional) = m + res (the resblock Add, lightcnn.py:84-88); relu_out (optional) = relu(y or m),
// the A operand of the positive-pass GEMM of the next conv.
__global__ void mfm_fwd_kernel(const float4* __restrict__ c, const float4* __restrict__ res, float4* __restrict__ m,
                               float4* __restrict__ y, float4* __restrict__ relu_out, int Cp4, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const size_t row = i / (unsigned)Cp4;
    const int c4 = (int)(i - row * Cp4);
    const float4 a = c[row * 2 * Cp4 + c4], b = c[row * 2 * Cp4 + Cp4 + c4];
    float4 v = make_float4(fmaxf(a.x, b.x), fmaxf(a.y, b.y), fmaxf(a.z, b.z), fmaxf(a.w, b.w));
    m[i] = v;
    if (y != nullptr) {
        const float4 r = res[i];
        v = make_float4(__fadd_rn(v.x, r.x), __fadd_rn(v.y, r.y), __fadd_rn(v.z, r.z), __fadd_rn(v.w, r.w));
        y[i] = v;
    }
    if (relu_out != nullptr) relu_out[i] = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

cudaError_t launch_mfm_fwd(const float* c, const float* res, float* m, float* y, float* relu_out, size_t rows, int Cp,
                           cudaStream_t st) {
    size_t total4 = rows * (Cp / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    mfm_fwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(
        reinterpret_cast<const float4*>(c), reinterpret_cast<const float4*>(res), reinterpret_cast<float4*>(m),
        reinterpret_cast<float4*>(y), reinterpret_cast<float4*>(relu_out), Cp / 4, total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ MFM backward (autograd of torch.max(a, b) + Split)
// g [J*HW, Cp], c [N*HW, 2*Cp] (row m reads m % Ms) -> z [J*HW, 2*Cp]: the larger branch takes g, exact ties take g/2 each.
__global__ void mfm_bwd_kernel(const float4* __restrict__ g, const float4* __restrict__ c, float4* __restrict__ z, int Cp4,
                               size_t rows_saved, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const size_t row = i / (unsigned)Cp4;
    const int c4 = (int)(i - row * Cp4);
    const size_t rs = (unsigned)row % (unsigned)rows_saved;
    const float4 a = c[rs * 2 * Cp4 + c4], b = c[rs * 2 * Cp4 + Cp4 + c4];
    const float4 gv = g[i];
    const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w}, gg[4] = {gv.x, gv.y, gv.z, gv.w};
    float za[4], zb[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float t = av[e] == bv[e] ? __fmul_rn(gg[e], 0.5f) : gg[e];
        za[e] = av[e] < bv[e] ? 0.f : t;
        zb[e] = bv[e] < av[e] ? 0.f : t;
    }
    z[row * 2 * Cp4 + c4] = make_float4(za[0], za[1], za[2], za[3]);
    z[row * 2 * Cp4 + Cp4 + c4] = make_float4(zb[0], zb[1], zb[2], zb[3]);
}

cudaError_t launch_mfm_bwd(const float* g, const float* c, float* z, size_t rows, size_t rows_saved, int Cp, cudaStream_t st) {
    size_t total4 = rows * (Cp / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    mfm_bwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(g),
                                                                     reinterpret_cast<const float4*>(c),
                                                                     reinterpret_cast<float4*>(z), Cp / 4, rows_saved, total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ p = maxpool2(m) + avgpool2(m)  (lightcnn.py:252-269)
// m [N,H,W,C] -> p [N,H/2,W/2,C]; ppos = maxpool2(relu(m)) + avgpool2(relu(m)): the positive-pass value (X of p's consumers).
__global__ void pool2_fwd_kernel(const float4* __restrict__ m, float4* __restrict__ p, float4* __restrict__ ppos, int H, int W,
                                 int C4, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const int c4 = (int)(i % (unsigned)C4);
    unsigned q = i / (unsigned)C4;
    const int Wo = W / 2, Ho = H / 2;
    const int wo = (int)(q % (unsigned)Wo); q /= (unsigned)Wo;
    const int ho = (int)(q % (unsigned)Ho);
    const size_t n = q / (unsigned)Ho;
    const float4* base = m + ((n * H + 2 * ho) * W + 2 * wo) * C4 + c4;
    const float4 v[4] = {base[0], base[C4], base[(size_t)W * C4], base[(size_t)W * C4 + C4]};
    float out[4], outp[4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float t[4] = {(&v[0].x)[e], (&v[1].x)[e], (&v[2].x)[e], (&v[3].x)[e]};
        const float mx = fmaxf(fmaxf(t[0], t[1]), fmaxf(t[2], t[3]));
        // torch's avg_pool2d accumulates the window in scan order and divides by the window size
        const float av = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(t[0], t[1]), t[2]), t[3]), 4.f);
        out[e] = __fadd_rn(mx, av);
        const float r[4] = {fmaxf(t[0], 0.f), fmaxf(t[1], 0.f), fmaxf(t[2], 0.f), fmaxf(t[3], 0.f)};
        const float mxr = fmaxf(fmaxf(r[0], r[1]), fmaxf(r[2], r[3]));
        const float avr = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(r[0], r[1]), r[2]), r[3]), 4.f);
        outp[e] = __fadd_rn(mxr, avr);
    }
    p[i] = make_float4(out[0], out[1], out[2], out[3]);
    if (ppos != nullptr) ppos[i] = make_float4(outp[0], outp[1], outp[2], outp[3]);
}

cudaError_t launch_pool2_fwd(const float* m, float* p, float* ppos, int N, int H, int W, int C, cudaStream_t st) {
    size_t total4 = (size_t)N * (H / 2) * (W / 2) * (C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    pool2_fwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(m),
                                                                       reinterpret_cast<float4*>(p),
                                                                       reinterpret_cast<float4*>(ppos), H, W, C / 4, total4);
    return cudaGetLastError();
}

// backward of the pooled sum: g_m = MaxPool2d backward (first maximum of the 2x2 window in scan order, as torch) +
// AvgPool2d backward (g/4 on every pixel).  g [J,H/2,W/2,C], m [N,H,W,C] -> gm [J,H,W,C].
__global__ void pool2_bwd_kernel(const float4* __restrict__ g, const float4* __restrict__ m, float4* __restrict__ gm, int H, int W,
                                 int C4, int N, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const int c4 = (int)(i % (unsigned)C4);
    unsigned q = i / (unsigned)C4;
    const int Wo = W / 2, Ho = H / 2;
    const int wo = (int)(q % (unsigned)Wo); q /= (unsigned)Wo;
    const int ho = (int)(q % (unsigned)Ho);
    const size_t j = q / (unsigned)Ho;
    const size_t n = (unsigned)j % (unsigned)N;
    const float4* base = m + ((n * H + 2 * ho) * W + 2 * wo) * C4 + c4;
    const float4 v[4] = {base[0], base[C4], base[(size_t)W * C4], base[(size_t)W * C4 + C4]};
    const float4 gv = g[i];
    float o[4][4];
#pragma unroll
    for (int e = 0; e < 4; ++e) {
        const float t[4] = {(&v[0].x)[e], (&v[1].x)[e], (&v[2].x)[e], (&v[3].x)[e]};
        int arg = 0;
        float best = t[0];
#pragma unroll
        for (int k = 1; k < 4; ++k)
            if (t[k] > best) { best = t[k]; arg = k; }
        const float ge = (&gv.x)[e];
        const float q4 = __fdiv_rn(ge, 4.f);
#pragma unroll
        for (int k = 0; k < 4; ++k) o[k][e] = k == arg ? __fadd_rn(ge, q4) : q4;
    }
    float4* ob = gm + ((j * H + 2 * ho) * W + 2 * wo) * C4 + c4;
    ob[0] = make_float4(o[0][0], o[0][1], o[0][2], o[0][3]);
    ob[C4] = make_float4(o[1][0], o[1][1], o[1][2], o[1][3]);
    ob[(size_t)W * C4] = make_float4(o[2][0], o[2][1], o[2][2], o[2][3]);
    ob[(size_t)W * C4 + C4] = make_float4(o[3][0], o[3][1], o[3][2], o[3][3]);
}

cudaError_t launch_pool2_bwd(const float* g, const float* m, float* gm, int J, int N, int H, int W, int C, cudaStream_t st) {
    size_t total4 = (size_t)J * (H / 2) * (W / 2) * (C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    pool2_bwd_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(g),
                                                                       reinterpret_cast<const float4*>(m),
                                                                       reinterpret_cast<float4*>(gm), H, W, C / 4, N, total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ relu copy (A operand of a positive-pass GEMM)
__global__ void relu_kernel(const float4* __restrict__ in, float4* __restrict__ out, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;      // 32-bit index arithmetic (launchers reject totals >= 2^31)
    if (i >= total4) return;
    const float4 v = in[i];
    out[i] = make_float4(fmaxf(v.x, 0.f), fmaxf(v.y, 0.f), fmaxf(v.z, 0.f), fmaxf(v.w, 0.f));
}

cudaError_t launch_relu(const float* in, float* out, size_t n, cudaStream_t st) {
    size_t total4 = n / 4;
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    relu_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(reinterpret_cast<const float4*>(in), reinterpret_cast<float4*>(out),
                                                                  total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ channel sum of P[-2] + per-row total (whitebox.py:499, 524)
// P2 [J,HW,C] -> chansum [J,HW] ; sums [J] (double, accumulated with atomics; zeroed by the launcher)
__global__ void __launch_bounds__(256) chansum_kernel(const float* __restrict__ P2, float* __restrict__ chansum,
                                                      double* __restrict__ sums, int HW, int C) {
    // one warp per pixel
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const size_t pix = (size_t)blockIdx.x * 8 + warp;
    const int j = blockIdx.y;
    __shared__ double part[8];
    float s = 0.f;
    if (pix < (size_t)HW) {
        const float* p = P2 + ((size_t)j * HW + pix) * C;
        for (int c = lane; c < C; c += 32) s += p[c];
    }
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (lane == 0) {
        if (pix < (size_t)HW) chansum[(size_t)j * HW + pix] = s;
        part[warp] = pix < (size_t)HW ? (double)s : 0.0;
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += part[k];
        atomicAdd(sums + j, t);
    }
}

cudaError_t launch_chansum(const float* P2, float* chansum, double* sums, int J, int HW, int C, cudaStream_t st) {
    cudaError_t e = cudaMemsetAsync(sums, 0, sizeof(double) * J, st);
    if (e != cudaSuccess) return e;
    dim3 grid((HW + 7) / 8, J);
    chansum_kernel<<<grid, 256, 0, st>>>(P2, chansum, sums, HW, C);
    return cudaGetLastError();
}

}  // namespace xfrb
