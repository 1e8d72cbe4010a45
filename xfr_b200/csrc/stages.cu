// HBM-bound stages of the whitebox path that are not GEMMs: stem, pooling, head glue, block
// boundaries of downsample blocks, MaxPool backward, contrastive combine, saliency post-filter.
// All tensors NHWC fp32, 128-bit accesses along channels.
#include "common.cuh"
#include <math.h>

namespace xfrb {

// ------------------------------------------------------------------ stem forward
// o[n,oh,ow,co] = sum_{r,s,ci} x[n,2oh+r-3,2ow+s-3,ci] * W[(r,s,ci),co] + b[co]      (resnet.py:177,225)
// block = 16x16 output pixels x 64 channels; 256 threads = 64 quads of 4 consecutive pixels x 4 channel groups of 16:
// 64 accumulators per thread, 8 FMAs per shared-memory load (the 8x8-pixel / 16-accumulator version was LSU-bound at 12 %
// of the FMA peak and re-read the 37 KB of weights once per 64 pixels).
constexpr int ST_T = 16;                  // output tile edge
constexpr int ST_P = ST_T * 2 + 5;        // input patch edge (37)
__global__ void __launch_bounds__(256, 2) stem_conv_kernel(const float* __restrict__ x, const float* __restrict__ W,
                                                           const float* __restrict__ b, float* __restrict__ o, int N) {
    extern __shared__ __align__(16) float sm[];
    float* sW = sm;                       // [147][64]
    float* sx = sm + 147 * 64;            // [37][37][3]
    const int tid = threadIdx.x;
    const int n = blockIdx.z, oh0 = blockIdx.y * ST_T, ow0 = blockIdx.x * ST_T;
    for (int i = tid; i < 147 * 64 / 4; i += 256) reinterpret_cast<float4*>(sW)[i] = __ldg(reinterpret_cast<const float4*>(W) + i);
    const int ih0 = oh0 * 2 - 3, iw0 = ow0 * 2 - 3;
    for (int i = tid; i < ST_P * ST_P * 3; i += 256) {
        int ci = i % 3, p = i / 3;
        int pw = p % ST_P, ph = p / ST_P;
        int ih = ih0 + ph, iw = iw0 + pw;
        float v = 0.f;
        if (ih >= 0 && ih < 224 && iw >= 0 && iw < 224) v = __ldg(x + (((size_t)n * 224 + ih) * 224 + iw) * 3 + ci);
        sx[i] = v;
    }
    __syncthreads();
    const int quad = tid >> 2, cg = tid & 3;
    const int py = quad >> 2, px0 = (quad & 3) * 4;
    float acc[4][16];
#pragma unroll
    for (int p = 0; p < 4; ++p)
#pragma unroll
        for (int j = 0; j < 16; ++j) acc[p][j] = 0.f;
    for (int r = 0; r < 7; ++r) {
        const float* xr = sx + ((py * 2 + r) * ST_P + px0 * 2) * 3;
        const float* wr = sW + (r * 21) * 64 + cg * 16;
#pragma unroll 7
        for (int sc = 0; sc < 21; ++sc) {
            const float xv[4] = {xr[sc], xr[sc + 6], xr[sc + 12], xr[sc + 18]};     // pixels 2 columns (6 floats) apart
            const float4* w4 = reinterpret_cast<const float4*>(wr + sc * 64);
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float4 w = w4[q];
#pragma unroll
                for (int p = 0; p < 4; ++p) {
                    acc[p][q * 4 + 0] = fmaf(xv[p], w.x, acc[p][q * 4 + 0]);
                    acc[p][q * 4 + 1] = fmaf(xv[p], w.y, acc[p][q * 4 + 1]);
                    acc[p][q * 4 + 2] = fmaf(xv[p], w.z, acc[p][q * 4 + 2]);
                    acc[p][q * 4 + 3] = fmaf(xv[p], w.w, acc[p][q * 4 + 3]);
                }
            }
        }
    }
    float4 bb[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) bb[q] = __ldg(reinterpret_cast<const float4*>(b + cg * 16) + q);
#pragma unroll
    for (int p = 0; p < 4; ++p) {
        float* op = o + (((size_t)n * 112 + oh0 + py) * 112 + ow0 + px0 + p) * 64 + cg * 16;
#pragma unroll
        for (int q = 0; q < 4; ++q)
            st4(op + q * 4, make_float4(acc[p][q * 4] + bb[q].x, acc[p][q * 4 + 1] + bb[q].y, acc[p][q * 4 + 2] + bb[q].z,
                                        acc[p][q * 4 + 3] + bb[q].w));
    }
}

// mp = maxpool3x3/2 pad 1 of relu(bn(o))   (resnet.py:226-228)
// arg (may be null): position r*3+s of the FIRST maximum of each window (torch's scan order), one byte per pooled element -
// what MaxPool2d backward routes by, recorded here so the backward sweep does not re-derive it per gradient row.
__global__ void stem_pool_kernel(const float* __restrict__ o, const float* __restrict__ bn, float* __restrict__ mp,
                                 unsigned char* __restrict__ arg, int total4, int pad) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    int c = (i & 15) * 4;
    int p = i >> 4;
    int pw = p % 56; p /= 56;
    int ph = p % 56;
    int n = p / 56;
    float4 al = ld4(bn + c), be = ld4(bn + 64 + c);
    float m[4] = {-INFINITY, -INFINITY, -INFINITY, -INFINITY};
    unsigned char am[4] = {0, 0, 0, 0};
    const float alv[4] = {al.x, al.y, al.z, al.w}, bev[4] = {be.x, be.y, be.z, be.w};
    for (int r = 0; r < 3; ++r) {
        int h = ph * 2 - pad + r;
        if (h < 0 || h >= 112) continue;
        for (int s = 0; s < 3; ++s) {
            int w = pw * 2 - pad + s;
            if (w < 0 || w >= 112) continue;
            float4 v = ld4(o + (((size_t)n * 112 + h) * 112 + w) * 64 + c);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float rv = fmaxf(__fadd_rn(__fmul_rn(vv[q], alv[q]), bev[q]), 0.f);
                if (rv > m[q]) { m[q] = rv; am[q] = (unsigned char)(r * 3 + s); }
            }
        }
    }
    st4(mp + (size_t)i * 4, make_float4(m[0], m[1], m[2], m[3]));
    if (arg != nullptr) reinterpret_cast<uchar4*>(arg)[i] = make_uchar4(am[0], am[1], am[2], am[3]);
}

cudaError_t launch_stem_fwd(const float* x, const float* W, const float* b, const float* bn, float* o, float* mp,
                            unsigned char* mp_arg, int N, int pool_pad, cudaStream_t st) {
    size_t smem = (147 * 64 + ST_P * ST_P * 3) * sizeof(float);
    static bool attr[XFRB_MAX_DEV] = {};
    const int dev = current_device_slot();
    if (!attr[dev]) {
        cudaFuncSetAttribute(stem_conv_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        attr[dev] = true;
    }
    stem_conv_kernel<<<dim3(112 / ST_T, 112 / ST_T, N), 256, smem, st>>>(x, W, b, o, N);
    int total4 = N * 56 * 56 * 16;
    stem_pool_kernel<<<(total4 + 255) / 256, 256, 0, st>>>(o, bn, mp, mp_arg, total4, pool_pad);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ shortcut helpers
__global__ void subsample2_kernel(const float* __restrict__ u, float* __restrict__ out, int H, int W, int C4, size_t total4) {
    // 32-bit index arithmetic (launchers reject totals >= 2^31): a 64-bit division by a run-time divisor costs ~100 instructions,
    // which made these bandwidth kernels issue-bound (ncu r2: join_kernel 78 % issue-active at 2.7 TB/s)
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned c = i % (unsigned)C4;
    unsigned p = i / (unsigned)C4;
    const unsigned w = p % (unsigned)(W / 2); p /= (unsigned)(W / 2);
    const unsigned h = p % (unsigned)(H / 2);
    const size_t n = p / (unsigned)(H / 2);
    reinterpret_cast<float4*>(out)[i] = __ldg(reinterpret_cast<const float4*>(u) + ((n * H + 2 * h) * W + 2 * w) * C4 + c);
}

__global__ void avgpool2_kernel(const float* __restrict__ u, float* __restrict__ out, int H, int W, int C4, size_t total4) {
    // 32-bit index arithmetic (launchers reject totals >= 2^31): a 64-bit division by a run-time divisor costs ~100 instructions,
    // which made these bandwidth kernels issue-bound (ncu r2: join_kernel 78 % issue-active at 2.7 TB/s)
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned c = i % (unsigned)C4;
    unsigned p = i / (unsigned)C4;
    const unsigned w = p % (unsigned)(W / 2); p /= (unsigned)(W / 2);
    const unsigned h = p % (unsigned)(H / 2);
    const size_t n = p / (unsigned)(H / 2);
    const float4* b = reinterpret_cast<const float4*>(u) + ((n * H + 2 * h) * W + 2 * w) * C4 + c;
    float4 a = __ldg(b), b1 = __ldg(b + C4), c0 = __ldg(b + (size_t)W * C4), c1 = __ldg(b + (size_t)W * C4 + C4);
    // torch avg_pool2d: sum in window order, then divide by the pool size
    float4 r;
    r.x = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.x, b1.x), c0.x), c1.x), 4.f);
    r.y = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.y, b1.y), c0.y), c1.y), 4.f);
    r.z = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.z, b1.z), c0.z), c1.z), 4.f);
    r.w = __fdiv_rn(__fadd_rn(__fadd_rn(__fadd_rn(a.w, b1.w), c0.w), c1.w), 4.f);
    reinterpret_cast<float4*>(out)[i] = r;
}

// ------------------------------------------------------------------ pair tensors (bf16x2 plan, common.cuh)
__global__ void to_pair_kernel(const float* __restrict__ in, float* __restrict__ out, int C, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned C4 = C / 4;
    const float4 t = __ldg(reinterpret_cast<const float4*>(in) + i);
    const float v[4] = {t.x, t.y, t.z, t.w};
    st_pair4(out, i / C4, C, (int)(i % C4) * 4, v);
}
__global__ void from_pair_kernel(const float* __restrict__ in, float* __restrict__ out, int C, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned C4 = C / 4;
    float v[4];
    ld_pair4(in, i / C4, C, (int)(i % C4) * 4, v);
    reinterpret_cast<float4*>(out)[i] = make_float4(v[0], v[1], v[2], v[3]);
}
cudaError_t launch_to_pair(const float* in, float* out, size_t rows, int C, int inverse, cudaStream_t st) {
    size_t total4 = rows * (size_t)(C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    if (total4 == 0) return cudaSuccess;
    if (inverse) from_pair_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(in, out, C, total4);
    else to_pair_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(in, out, C, total4);
    return cudaGetLastError();
}

cudaError_t launch_subsample2(const float* u, float* out, int N, int H, int W, int C, cudaStream_t st) {
    size_t total4 = (size_t)N * (H / 2) * (W / 2) * (C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    subsample2_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(u, out, H, W, C / 4, total4);
    return cudaGetLastError();
}
cudaError_t launch_avgpool2(const float* u, float* out, int N, int H, int W, int C, cudaStream_t st) {
    size_t total4 = (size_t)N * (H / 2) * (W / 2) * (C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    avgpool2_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(u, out, H, W, C / 4, total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ head forward glue
// v[n,c] = mean over 7x7 of u[n,:,:,c]   (resnet.py:235)
__global__ void avgpool7_kernel(const float* __restrict__ u, float* __restrict__ v, int C4, int total4) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    int c = i % C4, n = i / C4;
    const float4* b = reinterpret_cast<const float4*>(u) + (size_t)n * 49 * C4 + c;
    float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
    for (int p = 0; p < 49; ++p) {
        float4 t = __ldg(b + (size_t)p * C4);
        s.x += t.x; s.y += t.y; s.z += t.z; s.w += t.w;
    }
    reinterpret_cast<float4*>(v)[i] = make_float4(s.x / 49.f, s.y / 49.f, s.z / 49.f, s.w / 49.f);
}

// scratch [N, 1024] (dual tile order, bias already added) -> f1, f1p, nrm, xn.   512 threads per row.
__global__ void __launch_bounds__(512) head_norm_kernel(const float* __restrict__ scratch, int tn, float* __restrict__ f1,
                                                        float* __restrict__ f1p, float* __restrict__ xn, float* __restrict__ nrm,
                                                        float* __restrict__ xmul) {
    __shared__ float red[16];
    __shared__ float red2[16];
    const int n = blockIdx.x, c = threadIdx.x;
    const int half = tn / 2;
    const int t = c / half, j = c % half;
    float a = scratch[(size_t)n * 1024 + t * tn + j];
    float p = scratch[(size_t)n * 1024 + t * tn + half + j];
    f1[(size_t)n * 512 + c] = a;
    f1p[(size_t)n * 512 + c] = p;
    float s = a * a, s2 = p * p;
    for (int o = 16; o > 0; o >>= 1) { s += __shfl_xor_sync(0xffffffffu, s, o); s2 += __shfl_xor_sync(0xffffffffu, s2, o); }
    if ((c & 31) == 0) { red[c >> 5] = s; red2[c >> 5] = s2; }
    __syncthreads();
    float tot = 0.f, tot2 = 0.f;
    for (int i = 0; i < 16; ++i) { tot += red[i]; tot2 += red2[i]; }
    float nn = fmaxf(sqrtf(tot), 1e-12f);      // F.normalize eps (resnet.py:250)
    if (c == 0) nrm[n] = nn;
    xn[(size_t)n * 512 + c] = __fdiv_rn(a, nn);
    // X of the Multiply hook: relu(normalize(fc1 with relu(W) on relu(v)))  (positive pass, whitebox.py:327)
    if (xmul != nullptr) xmul[(size_t)n * 512 + c] = fmaxf(__fdiv_rn(p, fmaxf(sqrtf(tot2), 1e-12f)), 0.f);
}

cudaError_t launch_avgpool7(const float* u, float* v, int N, int C, cudaStream_t st) {
    int total4 = N * C / 4;
    avgpool7_kernel<<<(total4 + 255) / 256, 256, 0, st>>>(u, v, C / 4, total4);
    return cudaGetLastError();
}
cudaError_t launch_head_norm(const float* scratch, int tn, float* f1, float* f1p, float* xn, float* nrm, float* xmul, int N,
                             cudaStream_t st) {
    head_norm_kernel<<<N, 512, 0, st>>>(scratch, tn, f1, f1p, xn, nrm, xmul);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ head backward glue
// row j: seed = Pn[j,:] @ W2[j%N] ; x50 ; Multiply hook ; Jacobian of normalize -> scratch[j, 512]
__global__ void __launch_bounds__(512) head_bwd_a_kernel(const float* __restrict__ Pn, const float* __restrict__ W2, int C,
                                                         const float* __restrict__ f1p, const float* __restrict__ xn,
                                                         const float* __restrict__ nrm, float* __restrict__ scratch,
                                                         int N, int mode, float eps) {
    __shared__ float red[16];
    __shared__ float bc;
    const int j = blockIdx.x, n = j % N, d = threadIdx.x;
    auto block_sum = [&](float s) {
        for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
        __syncthreads();
        if ((d & 31) == 0) red[d >> 5] = s;
        __syncthreads();
        if (d == 0) {
            float t = 0.f;
            for (int i = 0; i < 16; ++i) t += red[i];
            bc = t;
        }
        __syncthreads();
        return bc;
    };
    float g = 0.f;
    if (W2 == nullptr) g = Pn[(size_t)j * 512 + d];               // Pn already is the seed (hooked fc2 head, xfrb_head_seed + xfrb_hook)
    else
        for (int c = 0; c < C; ++c) g = fmaf(Pn[(size_t)j * C + c], W2[((size_t)n * C + c) * 512 + d], g);
    g = g * 50.f;                                                 // Multiply backward (resnet.py:160-165)
    float x = xn[(size_t)n * 512 + d];
    float xmul = 0.f;
    if (mode == XFRB_MODE_ALL) {
        float fp = f1p[(size_t)n * 512 + d];
        float nn = fmaxf(sqrtf(block_sum(fp * fp)), 1e-12f);
        xmul = fmaxf(__fdiv_rn(fp, nn), 0.f);
    }
    g = hook<false>(fmaxf(x, 0.f), xmul, g, mode, eps);
    float dot = block_sum(x * g);
    scratch[(size_t)j * 512 + d] = __fdiv_rn(g - x * dot, nrm[n]);
}

// z [J,2048] (fc1 dgrad) -> Linear hook with a = x = v, /49, broadcast over 7x7
__global__ void head_bwd_b_kernel(const float* __restrict__ z, const float* __restrict__ v, float* __restrict__ g_out,
                                  int N, int C4, int total4, int mode, float eps) {
    int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    int c = i % C4, j = i / C4, n = j % N;
    float4 zz = reinterpret_cast<const float4*>(z)[i];
    float4 vv = __ldg(reinterpret_cast<const float4*>(v) + (size_t)n * C4 + c);
    float4 r;
    r.x = hook<true>(fmaxf(vv.x, 0.f), fmaxf(vv.x, 0.f), zz.x, mode, eps) / 49.f;
    r.y = hook<true>(fmaxf(vv.y, 0.f), fmaxf(vv.y, 0.f), zz.y, mode, eps) / 49.f;
    r.z = hook<true>(fmaxf(vv.z, 0.f), fmaxf(vv.z, 0.f), zz.z, mode, eps) / 49.f;
    r.w = hook<true>(fmaxf(vv.w, 0.f), fmaxf(vv.w, 0.f), zz.w, mode, eps) / 49.f;
    float4* o = reinterpret_cast<float4*>(g_out) + (size_t)j * 49 * C4 + c;
    for (int p = 0; p < 49; ++p) o[(size_t)p * C4] = r;
}

cudaError_t launch_head_bwd_a(const float* Pn, const float* W2, int C, const float* f1p, const float* xn, const float* nrm,
                              float* scratch, int J, int N, int mode, float eps, cudaStream_t st) {
    head_bwd_a_kernel<<<J, 512, 0, st>>>(Pn, W2, C, f1p, xn, nrm, scratch, N, mode, eps);
    return cudaGetLastError();
}
cudaError_t launch_head_bwd_b(const float* z, const float* v, float* g_out, int J, int N, int C, int mode, float eps,
                              cudaStream_t st) {
    int total4 = J * C / 4;
    head_bwd_b_kernel<<<(total4 + 255) / 256, 256, 0, st>>>(z, v, g_out, N, C / 4, total4, mode, eps);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ unfused block boundary

// MODE >= 0: the hook mode as a compile-time constant (drops the BatchNorm constants it does not use).
// FAST: `up`, `k` and C/4 are powers of two (every ResNet boundary: 1 or 2, 64 .. 512) - shifts and masks replace the runtime
// integer divisions, a block handles a stretch of ONE image row (grid.x = row of the [N*H] sample rows, grid.y = stretch), and the
// AvgPool backward's division by k*k is a multiplication by its exact reciprocal.  The form it replaces was issue-bound (ncu r2p:
// 79 % issue-active at 3.1 TB/s, ~1,160 instructions per warp, about a third of them integer division); results are bit-identical.
template <int MINB, int MODE, bool FAST>
__global__ void __launch_bounds__(256, MINB) join_kernel(JoinArgs a, size_t total4, int c4_shift, int up_shift, int k_shift) {
    const int mode = MODE >= 0 ? MODE : a.mode;
    // one thread = 4 channels of one pixel of one SAMPLE; it walks the gradient-row groups (mate, non-mate, ...) of that sample so
    // that the saved tensors (out, o3, xr3, res) and the BatchNorm constants are read once, not once per group
    int c, w, h, n;
    if (FAST) {
        const unsigned idx = blockIdx.y * blockDim.x + threadIdx.x;       // (w, c4) inside the image row
        if (idx >= ((unsigned)a.W << c4_shift)) return;
        c = (int)(idx & ((1u << c4_shift) - 1u)) * 4;
        w = (int)(idx >> c4_shift);
        n = (int)(blockIdx.x / (unsigned)a.H);                            // uniform per block
        h = (int)(blockIdx.x - (unsigned)n * a.H);
    } else {
        const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
        if (i >= total4) return;
        const unsigned C4 = a.C / 4;
        c = (int)(i % C4) * 4;
        unsigned p = i / C4;
        w = p % (unsigned)a.W; p /= (unsigned)a.W;
        h = p % (unsigned)a.H;
        n = (int)(p / (unsigned)a.H);
    }
    const size_t ms = ((size_t)n * a.H + h) * a.W + w;
    float4 uv = ld4(a.out + ms * a.C + c), ov = ld4(a.o3 + ms * a.C + c), xv = ld4(a.xr3 + ms * a.C + c);
    float u[4] = {uv.x, uv.y, uv.z, uv.w}, o[4] = {ov.x, ov.y, ov.z, ov.w}, x[4] = {xv.x, xv.y, xv.z, xv.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (mode == XFRB_MODE_ALL && a.res != nullptr && c < a.res_c) {
        float4 rv = ld4(a.res + ms * a.res_c + c);
        r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
    }
    float4 al = ld4(a.bn3 + c), be = ld4(a.bn3 + a.C + c), sp = ld4(a.bn3 + 2 * a.C + c), tp = ld4(a.bn3 + 3 * a.C + c);
    BnC b[4] = {{al.x, be.x, sp.x, tp.x}, {al.y, be.y, sp.y, tp.y}, {al.z, be.z, sp.z, tp.z}, {al.w, be.w, sp.w, tp.w}};
    const bool on_main = FAST ? (((h | w) & ((1 << up_shift) - 1)) == 0) : (h % a.up == 0 && w % a.up == 0);
    const bool on_res = (a.gres_lo != nullptr && c < a.gres_c);
    const int Hm = FAST ? a.H >> up_shift : a.H / a.up, Wm = FAST ? a.W >> up_shift : a.W / a.up;
    const int Hr = FAST ? a.H >> k_shift : a.H / a.k, Wr = FAST ? a.W >> k_shift : a.W / a.k;
    const int hm = FAST ? h >> up_shift : h / a.up, wm = FAST ? w >> up_shift : w / a.up;
    const int hr = FAST ? h >> k_shift : h / a.k, wr = FAST ? w >> k_shift : w / a.k;
    const float kk = (float)(a.k * a.k);
    const float rkk = 1.f / kk;                    // FAST: k*k is a power of two, x * (1/kk) == x / kk to the last bit
    // two gradient-row groups per trip: both groups' gradient loads are in flight before the first hook chain starts
    for (int j0 = n; j0 < a.J; j0 += 2 * a.N) {
        const int ng = (j0 + a.N < a.J) ? 2 : 1;
        float z[2][4] = {{0.f, 0.f, 0.f, 0.f}, {0.f, 0.f, 0.f, 0.f}};
        float4 tm[2], tr[2];
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            const int j = j0 + s * a.N;
            if (s < ng && on_main) tm[s] = ld4(a.zmain + (((size_t)j * Hm + hm) * Wm + wm) * a.C + c);
            if (s < ng && on_res) tr[s] = ld4(a.gres_lo + (((size_t)j * Hr + hr) * Wr + wr) * a.gres_c + c);
        }
#pragma unroll
        for (int s = 0; s < 2; ++s) {
            if (s >= ng) break;
            const int j = j0 + s * a.N;
            if (on_main) { z[s][0] = tm[s].x; z[s][1] = tm[s].y; z[s][2] = tm[s].z; z[s][3] = tm[s].w; }
            if (on_res) {
                if (FAST) {
                    z[s][0] = __fadd_rn(z[s][0], __fmul_rn(tr[s].x, rkk)); z[s][1] = __fadd_rn(z[s][1], __fmul_rn(tr[s].y, rkk));
                    z[s][2] = __fadd_rn(z[s][2], __fmul_rn(tr[s].z, rkk)); z[s][3] = __fadd_rn(z[s][3], __fmul_rn(tr[s].w, rkk));
                } else {
                    z[s][0] = __fadd_rn(z[s][0], __fdiv_rn(tr[s].x, kk)); z[s][1] = __fadd_rn(z[s][1], __fdiv_rn(tr[s].y, kk));
                    z[s][2] = __fadd_rn(z[s][2], __fdiv_rn(tr[s].z, kk)); z[s][3] = __fadd_rn(z[s][3], __fdiv_rn(tr[s].w, kk));
                }
            }
            float g[4], y3[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) join_chain(z[s][q], u[q], o[q], x[q], r[q], b[q], a.hooks, mode, a.eps, g[q], y3[q]);
            const size_t row = ((size_t)j * a.H + h) * a.W + w;
            const size_t off = (row * a.C + c) / 4;
            reinterpret_cast<float4*>(a.g_out)[off] = make_float4(g[0], g[1], g[2], g[3]);
            if (a.y3_pair) st_pair4(a.y3_out, row, a.C, c, y3);      // bf16x2 plan: y3 is the A operand of the next dgrad
            else reinterpret_cast<float4*>(a.y3_out)[off] = make_float4(y3[0], y3[1], y3[2], y3[3]);
        }
    }
}

// A/B twin of join_kernel: one thread per (gradient row, pixel, 4 channels); saved tensors are re-read per gradient-row group
__global__ void join_rows_kernel(JoinArgs a, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned C4 = a.C / 4;
    const int c = (int)(i % C4) * 4;
    unsigned p = i / C4;
    const int w = p % (unsigned)a.W; p /= (unsigned)a.W;
    const int h = p % (unsigned)a.H;
    const int j = (int)(p / (unsigned)a.H);
    int n = j % a.N;
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    if (h % a.up == 0 && w % a.up == 0) {
        int Hm = a.H / a.up, Wm = a.W / a.up;
        float4 t = ld4(a.zmain + (((size_t)j * Hm + h / a.up) * Wm + w / a.up) * a.C + c);
        z[0] = t.x; z[1] = t.y; z[2] = t.z; z[3] = t.w;
    }
    if (a.gres_lo != nullptr && c < a.gres_c) {
        int Hr = a.H / a.k, Wr = a.W / a.k;
        float4 t = ld4(a.gres_lo + (((size_t)j * Hr + h / a.k) * Wr + w / a.k) * a.gres_c + c);
        float kk = (float)(a.k * a.k);
        z[0] = __fadd_rn(z[0], __fdiv_rn(t.x, kk)); z[1] = __fadd_rn(z[1], __fdiv_rn(t.y, kk));
        z[2] = __fadd_rn(z[2], __fdiv_rn(t.z, kk)); z[3] = __fadd_rn(z[3], __fdiv_rn(t.w, kk));
    }
    size_t ms = ((size_t)n * a.H + h) * a.W + w;
    float4 uv = ld4(a.out + ms * a.C + c), ov = ld4(a.o3 + ms * a.C + c), xv = ld4(a.xr3 + ms * a.C + c);
    float u[4] = {uv.x, uv.y, uv.z, uv.w}, o[4] = {ov.x, ov.y, ov.z, ov.w}, x[4] = {xv.x, xv.y, xv.z, xv.w};
    float r[4] = {0.f, 0.f, 0.f, 0.f};
    if (a.mode == XFRB_MODE_ALL && a.res != nullptr && c < a.res_c) {
        float4 rv = ld4(a.res + ms * a.res_c + c);
        r[0] = rv.x; r[1] = rv.y; r[2] = rv.z; r[3] = rv.w;
    }
    float4 al = ld4(a.bn3 + c), be = ld4(a.bn3 + a.C + c), sp = ld4(a.bn3 + 2 * a.C + c), tp = ld4(a.bn3 + 3 * a.C + c);
    BnC b[4] = {{al.x, be.x, sp.x, tp.x}, {al.y, be.y, sp.y, tp.y}, {al.z, be.z, sp.z, tp.z}, {al.w, be.w, sp.w, tp.w}};
    float g[4], y3[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) join_chain(z[q], u[q], o[q], x[q], r[q], b[q], a.hooks, a.mode, a.eps, g[q], y3[q]);
    reinterpret_cast<float4*>(a.g_out)[i] = make_float4(g[0], g[1], g[2], g[3]);
    if (a.y3_pair) st_pair4(a.y3_out, i / C4, a.C, c, y3);
    else reinterpret_cast<float4*>(a.y3_out)[i] = make_float4(y3[0], y3[1], y3[2], y3[3]);
}

cudaError_t launch_join(const JoinArgs& a, cudaStream_t st) {
    if (a.N <= 0 || a.J % a.N != 0) return cudaErrorInvalidValue;
    static const int variant = [] { const char* e = getenv("XFRB_JOIN"); return e ? atoi(e) : 0; }();   // A/B probe: 1 per-row twin, 2 runtime hook mode, 5 / 6 min CTAs per SM
    if (variant == 1) {
        size_t total4 = (size_t)a.J * a.H * a.W * (a.C / 4);
        join_rows_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(a, total4);
        return cudaGetLastError();
    }
    size_t total4 = (size_t)a.N * a.H * a.W * (a.C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    // occupancy decides this kernel (tools/join_probe.py, 128 probes x 2 groups, the four boundaries of a ResNet-101 sweep, all variants
    // bit-identical): per-row twin (46 registers) 2,024 us; per sample with a runtime hook mode: 104 registers (2 CTAs per SM) 2,393 us,
    // 80 (3) 1,846 us, 64 (4) 1,667 us; hook mode as a template constant (fewer BatchNorm constants live): 4 CTAs 1,416 us (default),
    // 5 CTAs 1,364 us, 6 CTAs (192 B spilled) 1,451 us
    auto log2_of = [](int v) { int s = 0; while ((1 << s) < v) ++s; return (v > 0 && (1 << s) == v) ? s : -1; };
    const int c4s = log2_of(a.C / 4), ups = log2_of(a.up), ks = log2_of(a.k);
    const size_t rows = (size_t)a.N * a.H;
    if (variant == 0 && c4s >= 0 && ups >= 0 && ks >= 0 && a.C % 4 == 0 && rows < 0x7FFFFFFFull) {
        const dim3 grid((unsigned)rows, (unsigned)((((size_t)a.W << c4s) + 255) / 256));
        if (a.mode == XFRB_MODE_AWP) join_kernel<4, XFRB_MODE_AWP, true><<<grid, 256, 0, st>>>(a, total4, c4s, ups, ks);
        else if (a.mode == XFRB_MODE_ALL) join_kernel<4, XFRB_MODE_ALL, true><<<grid, 256, 0, st>>>(a, total4, c4s, ups, ks);
        else if (a.mode == XFRB_MODE_AFFINEONLY) join_kernel<4, XFRB_MODE_AFFINEONLY, true><<<grid, 256, 0, st>>>(a, total4, c4s, ups, ks);
        else return cudaErrorInvalidValue;
        return cudaGetLastError();
    }
    const unsigned grid = (unsigned)((total4 + 255) / 256);
    if (variant == 2) join_kernel<4, -1, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);                   // A/B probe: runtime hook mode
    else if (variant == 5 && a.mode == XFRB_MODE_AWP) join_kernel<5, XFRB_MODE_AWP, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);
    else if (variant == 6 && a.mode == XFRB_MODE_AWP) join_kernel<6, XFRB_MODE_AWP, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);
    else if (a.mode == XFRB_MODE_AWP) join_kernel<4, XFRB_MODE_AWP, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);
    else if (a.mode == XFRB_MODE_ALL) join_kernel<4, XFRB_MODE_ALL, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);
    else if (a.mode == XFRB_MODE_AFFINEONLY) join_kernel<4, XFRB_MODE_AFFINEONLY, false><<<grid, 256, 0, st>>>(a, total4, 0, 0, 0);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

__global__ void ds_res_kernel(const float* __restrict__ g, const float* __restrict__ ap, float* __restrict__ gres_lo,
                              int N, size_t HW, int C, int Cr, int mode, float eps, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned C4 = Cr / 4;
    const int c = (int)(i % C4) * 4;
    const size_t p = i / C4;              // j*HW + pixel
    const unsigned j = (unsigned)p / (unsigned)HW, pix = (unsigned)p % (unsigned)HW;
    const size_t n = j % (unsigned)N;
    float4 gv = ld4(g + p * C + c);
    float4 av = ld4(ap + (n * HW + pix) * Cr + c);
    float gg[4] = {gv.x, gv.y, gv.z, gv.w}, aa[4] = {av.x, av.y, av.z, av.w};
    float r[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float a = fmaxf(aa[q], 0.f);
        float z = hook<false>(a, a, gg[q], mode, eps);     // Add, slot 1
        r[q] = hook<false>(a, a, z, mode, eps);            // ConcatChannels
    }
    reinterpret_cast<float4*>(gres_lo)[i] = make_float4(r[0], r[1], r[2], r[3]);
}

cudaError_t launch_ds_res(const float* g, const float* ap, float* gres_lo, int J, int N, int H, int W, int C, int Cr,
                          int mode, float eps, cudaStream_t st) {
    size_t total4 = (size_t)J * H * W * (Cr / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    ds_res_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(g, ap, gres_lo, N, (size_t)H * W, C, Cr, mode, eps, total4);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ stem backward
// zc = two affine hooks (Conv2d of layer1.0.conv1, AvgPool2d(k=1) of its shortcut) on z = zmain + gres, a = x = mp
__global__ void stem_bwd_a_kernel(const float* __restrict__ zmain, const float* __restrict__ gres, const float* __restrict__ mp,
                                  float* __restrict__ zc, size_t per_sample4, int N, int mode, float eps, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned j = i / (unsigned)per_sample4, r = i % (unsigned)per_sample4;
    const size_t n = j % (unsigned)N;
    float4 a = reinterpret_cast<const float4*>(zmain)[i];
    float4 b = gres ? reinterpret_cast<const float4*>(gres)[i] : make_float4(0.f, 0.f, 0.f, 0.f);
    float4 m = __ldg(reinterpret_cast<const float4*>(mp) + n * per_sample4 + r);
    float z[4] = {__fadd_rn(a.x, b.x), __fadd_rn(a.y, b.y), __fadd_rn(a.z, b.z), __fadd_rn(a.w, b.w)};
    float mm[4] = {fmaxf(m.x, 0.f), fmaxf(m.y, 0.f), fmaxf(m.z, 0.f), fmaxf(m.w, 0.f)};
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        z[q] = hook<true>(mm[q], mm[q], z[q], mode, eps);
        z[q] = hook<true>(mm[q], mm[q], z[q], mode, eps);
    }
    reinterpret_cast<float4*>(zc)[i] = make_float4(z[0], z[1], z[2], z[3]);
}

// one thread = one (j, h, w, 4 channels) element of the 112x112x64 stem activation
__global__ void __launch_bounds__(256) stem_bwd_b_kernel(const float* __restrict__ zc, const float* __restrict__ o,
                                                         const float* __restrict__ bn, float* __restrict__ P2,
                                                         float* __restrict__ chansum, double* __restrict__ sums,
                                                         const unsigned char* __restrict__ mp_arg,
                                                         int N, int mode, float eps, int pad) {
    // grid: (112*112*16/256, J)
    __shared__ double red[8];
    const int j = blockIdx.y, n = j % N;
    const int i = blockIdx.x * 256 + threadIdx.x;      // < 112*112*16
    const int c = (i & 15) * 4;
    const int pix = i >> 4;
    const int h = pix / 112, w = pix % 112;
    float4 al = ld4(bn + c), be = ld4(bn + 64 + c), sp = ld4(bn + 128 + c), tp = ld4(bn + 192 + c);
    const float alv[4] = {al.x, al.y, al.z, al.w}, bev[4] = {be.x, be.y, be.z, be.w};
    const float spv[4] = {sp.x, sp.y, sp.z, sp.w}, tpv[4] = {tp.x, tp.y, tp.z, tp.w};
    const float* ob = o + (size_t)n * 112 * 112 * 64 + c;
    float ov[4], me[4];
    {
        float4 v = ld4(ob + ((size_t)h * 112 + w) * 64);
        ov[0] = v.x; ov[1] = v.y; ov[2] = v.z; ov[3] = v.w;
#pragma unroll
        for (int q = 0; q < 4; ++q) me[q] = fmaxf(__fadd_rn(__fmul_rn(ov[q], alv[q]), bev[q]), 0.f);
    }
    float z[4] = {0.f, 0.f, 0.f, 0.f};
    // MaxPool2d(3, 2, pad) backward as a gather: pooled row ph covers input rows 2ph-pad..2ph-pad+2 (clipped: pad 1 for
    // the STR net, pad 0 + ceil_mode for VGGFace2); (h,w) receives the gradient of every window whose FIRST maximum
    // (row-major scan, as torch's CPU kernel) it is.
    if (mp_arg != nullptr) {
        // the forward pass recorded which position won each window: the (up to) four windows' arg-max bytes and gradients are
        // loaded together, from clamped addresses, before any of them is used - the loop form below waited for each load in turn
        // (ncu r2p: 7.9 warps stalled on the long scoreboard per issue, 1.9 TB/s).  Same additions in the same order.
        uchar4 am[4];
        float4 gz[4];
        bool ok[4];
        int me_idx[4];
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            const int a = t >> 1, b2 = t & 1;
            const int ph = (h + pad) / 2 - a, pw = (w + pad) / 2 - b2;
            ok[t] = !(ph < 0 || ph >= 56 || 2 * ph - pad > h || h > 2 * ph - pad + 2) && !(pw < 0 || pw >= 56 || 2 * pw - pad > w || w > 2 * pw - pad + 2);
            const int phc = ok[t] ? ph : 0, pwc = ok[t] ? pw : 0;
            me_idx[t] = (h - (2 * phc - pad)) * 3 + (w - (2 * pwc - pad));
            am[t] = __ldg(reinterpret_cast<const uchar4*>(mp_arg) + (((size_t)n * 56 + phc) * 56 + pwc) * 16 + (c >> 2));
            gz[t] = ld4(zc + (((size_t)j * 56 + phc) * 56 + pwc) * 64 + c);
        }
#pragma unroll
        for (int t = 0; t < 4; ++t) {
            if (!ok[t]) continue;
            if (am[t].x == me_idx[t]) z[0] = __fadd_rn(z[0], gz[t].x);
            if (am[t].y == me_idx[t]) z[1] = __fadd_rn(z[1], gz[t].y);
            if (am[t].z == me_idx[t]) z[2] = __fadd_rn(z[2], gz[t].z);
            if (am[t].w == me_idx[t]) z[3] = __fadd_rn(z[3], gz[t].w);
        }
    } else
#pragma unroll
    for (int a = 0; a < 2; ++a) {
        int ph = (h + pad) / 2 - a;
        if (ph < 0 || ph >= 56 || 2 * ph - pad > h || h > 2 * ph - pad + 2) continue;
#pragma unroll
        for (int b2 = 0; b2 < 2; ++b2) {
            int pw = (w + pad) / 2 - b2;
            if (pw < 0 || pw >= 56 || 2 * pw - pad > w || w > 2 * pw - pad + 2) continue;
            bool win[4] = {true, true, true, true};
            const int my = h - (2 * ph - pad), mx = w - (2 * pw - pad);     // my position inside the window
            if (mp_arg != nullptr) {        // the forward pass recorded which position won each window
                const uchar4 am = __ldg(reinterpret_cast<const uchar4*>(mp_arg) + (((size_t)n * 56 + ph) * 56 + pw) * 16 + (c >> 2));
                const int me_idx = my * 3 + mx;
                win[0] = am.x == me_idx; win[1] = am.y == me_idx; win[2] = am.z == me_idx; win[3] = am.w == me_idx;
            } else
#pragma unroll
            for (int r = 0; r < 3; ++r)
#pragma unroll
                for (int s2 = 0; s2 < 3; ++s2) {
                    int hh = 2 * ph - pad + r, ww = 2 * pw - pad + s2;
                    if ((r == my && s2 == mx) || hh < 0 || hh >= 112 || ww < 0 || ww >= 112) continue;
                    float4 v = ld4(ob + ((size_t)hh * 112 + ww) * 64);
                    const float vv[4] = {v.x, v.y, v.z, v.w};
                    const bool before = (r < my) || (r == my && s2 < mx);
#pragma unroll
                    for (int q = 0; q < 4; ++q) {
                        float rv = fmaxf(__fadd_rn(__fmul_rn(vv[q], alv[q]), bev[q]), 0.f);
                        win[q] = win[q] && (before ? (rv < me[q]) : (rv <= me[q]));
                    }
                }
            float4 gz = ld4(zc + (((size_t)j * 56 + ph) * 56 + pw) * 64 + c);
            const float gzv[4] = {gz.x, gz.y, gz.z, gz.w};
#pragma unroll
            for (int q = 0; q < 4; ++q)
                if (win[q]) z[q] = __fadd_rn(z[q], gzv[q]);
        }
    }
    float p[4];
    float csum = 0.f;
#pragma unroll
    for (int q = 0; q < 4; ++q) {
        float r = me[q];
        float ro = fmaxf(ov[q], 0.f);
        float xrelu = fmaxf(__fadd_rn(__fmul_rn(ro, spv[q]), tpv[q]), 0.f);
        float zz = hook<false>(r, xrelu, z[q], mode, eps);   // ReLU hook
        zz = hook<false>(r, r, zz, mode, eps);               // MaxPool2d hook
        zz = r > 0.f ? zz : 0.f;
        zz = __fmul_rn(zz, spv[q]);
        p[q] = __fmul_rn(ro, fmaxf(zz, 0.f));                // BatchNorm hook records P[-2]
        csum += p[q];
    }
    st4(P2 + ((size_t)j * 112 * 112 + pix) * 64 + c, make_float4(p[0], p[1], p[2], p[3]));
    // channel sum across the 16 lanes that share a pixel
    for (int o2 = 8; o2 > 0; o2 >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o2);
    if ((threadIdx.x & 15) == 0) chansum[(size_t)j * 112 * 112 + pix] = csum;
    double ds = (threadIdx.x & 15) == 0 ? (double)csum : 0.0;
    for (int o2 = 16; o2 > 0; o2 >>= 1) ds += __shfl_xor_sync(0xffffffffu, ds, o2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = ds;
    __syncthreads();
    if (threadIdx.x == 0) {
        double t = 0.0;
        for (int k = 0; k < 8; ++k) t += red[k];
        atomicAdd(sums + j, t);
    }
}

// The same stage with one thread per 2x2 block of stem pixels x 4 channels (needs the arg-max bytes of the forward sweep).  The four
// pixels of a quad (2i + dh, 2j + dw) only ever belong to the pooling windows ph in {i - 1 + PAD, i + PAD} x pw likewise: four
// (arg byte, gradient) pairs are loaded for four pixels instead of nine, and with PAD a template constant which windows a pixel
// belongs to, and its position inside each, are compile-time facts - the per-pixel kernel above spent a third of its ~380
// instructions per thread evaluating them at run time (ncu r2ah: 61 % issue-active at 2.2 TB/s).  Same additions in the same order
// (windows in (a, b2) order), same channel-sum tree: bit-identical.
template <int PAD, int MODE>
__global__ void __launch_bounds__(256) stem_bwd_quad_kernel(const float* __restrict__ zc, const float* __restrict__ o,
                                                            const float* __restrict__ bn, float* __restrict__ P2,
                                                            float* __restrict__ chansum, double* __restrict__ sums,
                                                            const unsigned char* __restrict__ mp_arg, int N, float eps) {
    // grid: (56*56*16/256, J)
    __shared__ double red[8];
    const int j = blockIdx.y, n = j % N;
    const int t = blockIdx.x * 256 + threadIdx.x;      // < 56*56*16
    const int c = (t & 15) * 4;
    const int quad = t >> 4;
    const int qi = quad / 56, qj = quad % 56;
    float4 al = ld4(bn + c), be = ld4(bn + 64 + c), sp = ld4(bn + 128 + c), tp = ld4(bn + 192 + c);
    const float alv[4] = {al.x, al.y, al.z, al.w}, bev[4] = {be.x, be.y, be.z, be.w};
    const float spv[4] = {sp.x, sp.y, sp.z, sp.w}, tpv[4] = {tp.x, tp.y, tp.z, tp.w};
    // the four candidate windows [wy][wx]: ph = qi - 1 + PAD + wy, pw = qj - 1 + PAD + wx
    uchar4 am[2][2];
    float4 gz[2][2];
    bool ok[2][2];
#pragma unroll
    for (int wy = 0; wy < 2; ++wy)
#pragma unroll
        for (int wx = 0; wx < 2; ++wx) {
            const int ph = qi - 1 + PAD + wy, pw = qj - 1 + PAD + wx;
            ok[wy][wx] = ph >= 0 && ph < 56 && pw >= 0 && pw < 56;
            const int phc = ok[wy][wx] ? ph : 0, pwc = ok[wy][wx] ? pw : 0;
            am[wy][wx] = __ldg(reinterpret_cast<const uchar4*>(mp_arg) + (((size_t)n * 56 + phc) * 56 + pwc) * 16 + (c >> 2));
            gz[wy][wx] = ld4(zc + (((size_t)j * 56 + phc) * 56 + pwc) * 64 + c);
        }
    const float* ob = o + (size_t)n * 112 * 112 * 64 + c;
    float4 ov4[2][2];
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) ov4[dh][dw] = ld4(ob + ((size_t)(2 * qi + dh) * 112 + (2 * qj + dw)) * 64);
    double dsum = 0.0;
#pragma unroll
    for (int dh = 0; dh < 2; ++dh)
#pragma unroll
        for (int dw = 0; dw < 2; ++dw) {
            const int pix = (2 * qi + dh) * 112 + (2 * qj + dw);
            const float ov[4] = {ov4[dh][dw].x, ov4[dh][dw].y, ov4[dh][dw].z, ov4[dh][dw].w};
            float z[4] = {0.f, 0.f, 0.f, 0.f};
            // windows in the order of the per-pixel kernel: a = 0, 1 (ph = (h + PAD) / 2 - a), b2 = 0, 1
#pragma unroll
            for (int a = 0; a < 2; ++a) {
                constexpr int dummy = 0; (void)dummy;
                const int ky = (dh + PAD) / 2 - a;                         // ph = qi + ky
                if (!(2 * ky - PAD <= dh && dh <= 2 * ky - PAD + 2)) continue;
                const int wy = ky + 1 - PAD;                               // index into the loaded windows
                if (wy < 0 || wy > 1) continue;
                const int my = dh - 2 * ky + PAD;                          // row inside the window
#pragma unroll
                for (int b2 = 0; b2 < 2; ++b2) {
                    const int kx = (dw + PAD) / 2 - b2;
                    if (!(2 * kx - PAD <= dw && dw <= 2 * kx - PAD + 2)) continue;
                    const int wx = kx + 1 - PAD;
                    if (wx < 0 || wx > 1) continue;
                    const int me_idx = my * 3 + (dw - 2 * kx + PAD);
                    if (!ok[wy][wx]) continue;
                    if (am[wy][wx].x == me_idx) z[0] = __fadd_rn(z[0], gz[wy][wx].x);
                    if (am[wy][wx].y == me_idx) z[1] = __fadd_rn(z[1], gz[wy][wx].y);
                    if (am[wy][wx].z == me_idx) z[2] = __fadd_rn(z[2], gz[wy][wx].z);
                    if (am[wy][wx].w == me_idx) z[3] = __fadd_rn(z[3], gz[wy][wx].w);
                }
            }
            float p[4];
            float csum = 0.f;
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const float r = fmaxf(__fadd_rn(__fmul_rn(ov[q], alv[q]), bev[q]), 0.f);
                const float ro = fmaxf(ov[q], 0.f);
                const float xrelu = fmaxf(__fadd_rn(__fmul_rn(ro, spv[q]), tpv[q]), 0.f);
                float zz = hook<false>(r, xrelu, z[q], MODE, eps);   // ReLU hook
                zz = hook<false>(r, r, zz, MODE, eps);               // MaxPool2d hook
                zz = r > 0.f ? zz : 0.f;
                zz = __fmul_rn(zz, spv[q]);
                p[q] = __fmul_rn(ro, fmaxf(zz, 0.f));                // BatchNorm hook records P[-2]
                csum += p[q];
            }
            st4(P2 + ((size_t)j * 112 * 112 + pix) * 64 + c, make_float4(p[0], p[1], p[2], p[3]));
            for (int o2 = 8; o2 > 0; o2 >>= 1) csum += __shfl_xor_sync(0xffffffffu, csum, o2);    // the 16 lanes that share a pixel
            if ((threadIdx.x & 15) == 0) {
                chansum[(size_t)j * 112 * 112 + pix] = csum;
                dsum += (double)csum;
            }
        }
    for (int o2 = 16; o2 > 0; o2 >>= 1) dsum += __shfl_xor_sync(0xffffffffu, dsum, o2);
    if ((threadIdx.x & 31) == 0) red[threadIdx.x >> 5] = dsum;
    __syncthreads();
    if (threadIdx.x == 0) {
        double tsum = 0.0;
        for (int k = 0; k < 8; ++k) tsum += red[k];
        atomicAdd(sums + j, tsum);
    }
}

cudaError_t launch_stem_bwd(const float* zmain, const float* gres, const float* o, const float* mp, const float* bn,
                            float* zc, float* P2, float* chansum, double* sums, const unsigned char* mp_arg, int J, int N,
                            int mode, float eps, int pool_pad, cudaStream_t st) {
    size_t per4 = (size_t)56 * 56 * 16, total4 = per4 * J;
    cudaMemsetAsync(sums, 0, sizeof(double) * J, st);
    stem_bwd_a_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(zmain, gres, mp, zc, per4, N, mode, eps, total4);
    static const int quad = [] { const char* e = getenv("XFRB_STEM_QUAD"); return e ? atoi(e) : 1; }();       // A/B probe: 0 = the per-pixel kernel
    if (quad && mp_arg != nullptr && (pool_pad == 0 || pool_pad == 1) && mode >= 0 && mode <= 2 && J <= 65535) {
        const dim3 grid(56 * 56 * 16 / 256, J);
#define XFRB_QUAD(PAD_)                                                                                                          \
        switch (mode) {                                                                                                          \
            case XFRB_MODE_AWP: stem_bwd_quad_kernel<PAD_, XFRB_MODE_AWP><<<grid, 256, 0, st>>>(zc, o, bn, P2, chansum, sums, mp_arg, N, eps); break;   \
            case XFRB_MODE_ALL: stem_bwd_quad_kernel<PAD_, XFRB_MODE_ALL><<<grid, 256, 0, st>>>(zc, o, bn, P2, chansum, sums, mp_arg, N, eps); break;   \
            default: stem_bwd_quad_kernel<PAD_, XFRB_MODE_AFFINEONLY><<<grid, 256, 0, st>>>(zc, o, bn, P2, chansum, sums, mp_arg, N, eps); break;     \
        }
        if (pool_pad == 1) { XFRB_QUAD(1) } else { XFRB_QUAD(0) }
#undef XFRB_QUAD
        return cudaGetLastError();
    }
    stem_bwd_b_kernel<<<dim3(112 * 112 * 16 / 256, J), 256, 0, st>>>(zc, o, bn, P2, chansum, sums, mp_arg, N, mode, eps, pool_pad);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ projection-shortcut helpers (VGGFace2 ResNet-50)
// BatchNorm backward with gamma+ then the BatchNorm hook of a conv output: y = relu(o) * relu(g*sp) / (xr + eps).
// g [J,HW,C]; o, xr [N,HW,C].  kind 1 instead emits the positive-pass BN, relu(o)*sp + tp (X of the shortcut, MODE_ALL).
__global__ void bn_hook_kernel(const float* __restrict__ g, const float* __restrict__ o, const float* __restrict__ xr,
                               const float* __restrict__ bn, float* __restrict__ y, size_t rows_saved, int C, int kind, int mode,
                               float eps, size_t total4) {
    const unsigned i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total4) return;
    const unsigned C4 = C / 4;
    const int c = (int)(i % C4) * 4;
    const size_t row = i / C4;
    const size_t rs = (unsigned)row % (unsigned)rows_saved;
    float4 ov = ld4(o + rs * C + c);
    float4 sp = ld4(bn + 2 * C + c), tp = ld4(bn + 3 * C + c);
    const float oo[4] = {ov.x, ov.y, ov.z, ov.w}, spv[4] = {sp.x, sp.y, sp.z, sp.w}, tpv[4] = {tp.x, tp.y, tp.z, tp.w};
    float r[4];
    if (kind == 1) {
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = __fadd_rn(__fmul_rn(fmaxf(oo[q], 0.f), spv[q]), tpv[q]);
    } else {
        float4 gv = ld4(g + row * C + c), xv = ld4(xr + rs * C + c);
        const float gg[4] = {gv.x, gv.y, gv.z, gv.w}, xx[4] = {xv.x, xv.y, xv.z, xv.w};
#pragma unroll
        for (int q = 0; q < 4; ++q) r[q] = hook<true>(fmaxf(oo[q], 0.f), xx[q], __fmul_rn(gg[q], spv[q]), mode, eps);
    }
    reinterpret_cast<float4*>(y)[i] = make_float4(r[0], r[1], r[2], r[3]);
}

cudaError_t launch_bn_hook(const float* g, const float* o, const float* xr, const float* bn, float* y, size_t rows, size_t rows_saved,
                           int C, int kind, int mode, float eps, cudaStream_t st) {
    size_t total4 = rows * (C / 4);
    if (total4 >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // kernels index with 32-bit arithmetic
    bn_hook_kernel<<<(unsigned)((total4 + 255) / 256), 256, 0, st>>>(g, o, xr, bn, y, rows_saved, C, kind, mode, eps, total4);
    return cudaGetLastError();
}

// seed[j,:] = Pn[j,:] @ W2[j % N]   (un-hooked classifier rows, signed)
__global__ void head_seed_kernel(const float* __restrict__ Pn, const float* __restrict__ W2, int C, int D, int N,
                                 float* __restrict__ seed) {
    const int j = blockIdx.x, n = j % N;
    for (int d = threadIdx.x; d < D; d += blockDim.x) {
        float g = 0.f;
        for (int c = 0; c < C; ++c) g = fmaf(Pn[(size_t)j * C + c], W2[((size_t)n * C + c) * D + d], g);
        seed[(size_t)j * D + d] = g;
    }
}

cudaError_t launch_head_seed(const float* Pn, const float* W2, int C, int D, int J, int N, float* seed, cudaStream_t st) {
    head_seed_kernel<<<J, 128, 0, st>>>(Pn, W2, C, D, N, seed);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ generic single hook (priors, P recording, true gradients)
// One _backward_ebp firing (whitebox.py:381-430) as its own kernel: used by layerwise_ebp / layerwise_contrastive_ebp /
// weighted_subtree_ebp, where a prior may replace p at one firing (whitebox.py:390-392, 570-577), every p is recorded
// (self.P) and true gradients are needed (self.dA, whitebox.py:353-358).  (a, x) come from a recipe over saved tensors:
//   0: a = x = relu(s0)                                   (block outputs / inputs, pooled tensors, vectors)
//   1: a = relu(bn(s0)), x = relu(relu(s0)*sp + tp)       (ReLU hook at an inner activation)
//   2: a = x = relu(bn(s0))                               (consumer Conv2d / MaxPool2d hook at an inner activation)
//   3: a = relu(s0), x = s1                               (BatchNorm hook: s0 = o, s1 = xr)
//   4: a = relu(s0), x = relu(relu(bn(s1)) + relu(s2))    (STR block ReLU hook: s0 = out, s1 = o3, s2 = residual)
//   5: a = relu(s0), x = s1                               (materialised pair, e.g. the Multiply hook)
//   6: a = relu(s0), x = relu(s1) + relu(s2)              (Light-CNN resblock output: s0 = out + res, s1 = out, s2 = res)
//   7: a = relu(s0), x = relu(s1)                         (Light-CNN Split hook: s0 = conv output, s1 = its positive twin)
//   8: a = relu(s0), x = relu(relu(s1)*sp + tp + s2)      (VGGFace2 ResNet-50 block ReLU: s0 = out, s1 = o3, s2 = positive shortcut)
struct HookPrior { int row; long long elem; float val; const float* tensor; int probe_row; long long probe_elem; };

// the prior of this firing: from the device table entry when there is one (graph replay), else from the launch arguments.  Only
// the two row numbers are read by every thread; the rest is fetched by the threads of the rows they name.
__device__ __forceinline__ HookPrior hook_prior(const HookArgs& A, int j) {
    HookPrior P;
    P.elem = 0; P.val = 0.f; P.tensor = nullptr; P.probe_elem = -1;
    if (A.ptab != nullptr) {
        P.row = A.ptab->row;
        P.probe_row = A.probe_out != nullptr ? A.ptab->probe_row : -1;
        if (j == P.row) { P.elem = A.ptab->elem; P.val = A.ptab->val; P.tensor = A.ptab->tensor; }
        if (j == P.probe_row) P.probe_elem = A.ptab->probe_elem;
    } else {
        P.row = A.prior_row; P.elem = A.prior_elem; P.val = A.prior_val; P.tensor = A.prior; P.probe_row = -1;
    }
    return P;
}

// (a, x) of one element from the recipe sources
__device__ __forceinline__ void hook_ax(int recipe, float v0, float v1, float v2, const BnC& b, float& a, float& x) {
    switch (recipe) {
        case 0: a = x = fmaxf(v0, 0.f); break;
        case 1: a = bn_act(v0, b); x = fmaxf(__fadd_rn(__fmul_rn(fmaxf(v0, 0.f), b.sp), b.tp), 0.f); break;
        case 2: a = x = bn_act(v0, b); break;
        case 4: a = fmaxf(v0, 0.f); x = fmaxf(__fadd_rn(bn_act(v1, b), fmaxf(v2, 0.f)), 0.f); break;
        case 6: a = fmaxf(v0, 0.f); x = __fadd_rn(fmaxf(v1, 0.f), fmaxf(v2, 0.f)); break;
        case 7: a = fmaxf(v0, 0.f); x = fmaxf(v1, 0.f); break;
        case 8: a = fmaxf(v0, 0.f); x = fmaxf(__fadd_rn(__fadd_rn(__fmul_rn(fmaxf(v1, 0.f), b.sp), b.tp), v2), 0.f); break;
        default: a = fmaxf(v0, 0.f); x = v1; break;          // 3, 5
    }
}

// the firing itself for one element: z (after the pre-scales) -> return value (after the post ops); pv = what P_out records
template <int MODE>      // MODE >= 0: the hook mode as a compile-time constant, -1: A.mode
__device__ __forceinline__ float hook_fire(const HookArgs& A, const HookPrior& P, bool has_prior, size_t e, float pscale, float z, float a,
                                           float x, float& pv) {
    const int mode = MODE >= 0 ? MODE : A.mode;
    float ret;
    if (mode == XFRB_MODE_NONE) {
        pv = z;                                                // dA: the true gradient at this hooked tensor
        ret = z;
    } else {
        const float zh = fmaxf(z, 0.f);
        pv = __fmul_rn(a, zh);
        float pr = 0.f;
        if (has_prior) {
            pr = P.tensor != nullptr ? P.tensor[e] : ((long long)e == P.elem ? P.val : 0.f);
            pv = pr;                                           // p.data.copy_(p_prior)
        }
        const float quo = __fdividef(pv, __fadd_rn(x, A.eps));
        if (mode == XFRB_MODE_ALL) ret = quo;
        else if (mode == XFRB_MODE_AFFINEONLY) ret = A.affine ? quo : z;
        else if (mode == XFRB_MODE_AWP) {
            if (has_prior) ret = A.affine ? (pr > 0.f ? quo : 0.f) : (pr > 0.f ? z : 0.f);
            else ret = A.affine ? quo : zh;
        } else ret = z;
        if (mode == XFRB_MODE_ALL && has_prior && A.relu_or_maxpool == 2) ret = z;    // 'norelu' (mode id ALL + flag 2)
    }
    if (A.post_mask) ret = a > 0.f ? ret : 0.f;
    if (A.post_scale_row >= 0) ret = __fmul_rn(ret, pscale);          // pscale = bn[post_scale_row][c]
    return ret;
}

// VEC = 4: one thread = 4 consecutive channels of one (row, pixel) - every channel count involved is a multiple of 4 and every
// tensor 16-byte aligned (launch_hook_chain checks) - so each tensor moves as float4 and the index arithmetic is paid once per four
// elements (the scalar form ran at 1.4 TB/s of its streams, issue-bound: profiles/r2_notes.md).  VEC = 1: any shape.
//
// CHAIN: ch.n consecutive firings on the same [J,H,W,C] tensor in one launch - link l + 1 takes link l's return value (after its post
// ops) from registers instead of from memory; a link's z_out / P_out are stored only where they are non-null.  The sweeps of
// generic.py / lightcnn.py fire 3 - 5 hooks in a row between two GEMMs (ReLU, Conv2d, Add, Add, BatchNorm2d on a block output):
// as separate launches every one re-read and re-wrote the gradient tensor.
//
// Row skipping: with ch.row_start (zero-seeded prior sweeps) gradient row j is all zero until the firing that carries its prior,
// row_start[j]: earlier firings load nothing for it and store zeros where a tensor is kept (the GEMMs in between may turn those
// zeros into anything they like - rows never mix), and AT that firing the incoming gradient is taken as zero.
struct HookIdx {            // host-prepared index arithmetic: grid.y = gradient row, grid.x covers one row's H*W*C/VEC threads
    unsigned per_row;       // H * W * (C / VEC)
    int cv_shift;           // log2(C / VEC) or -1 (divide)
    unsigned w_magic;       // ceil(2^32 / W) when floor(p / W) == umulhi(p, w_magic) for every pixel p of the tensor, else 0 (divide)
    int up_shift;           // log2(up) or -1 (divide)
};
template <int VEC, int MODE>
__global__ void __launch_bounds__(256) hook_kernel(HookChain ch, HookIdx ix) {
    const HookArgs& A0 = ch.a[0];
    const unsigned il = blockIdx.x * blockDim.x + threadIdx.x;         // position inside the gradient row
    if (il >= ix.per_row) return;
    const int j = blockIdx.y;
    const unsigned i = (unsigned)j * ix.per_row + il;                   // 32-bit index arithmetic (the launcher rejects totals >= 2^32)
    const unsigned Cv = (unsigned)A0.C / VEC;
    const unsigned pix = ix.cv_shift >= 0 ? il >> ix.cv_shift : il / Cv;
    const int c = (int)(il - pix * Cv) * VEC;
    const int h = (int)(ix.w_magic ? __umulhi(pix, ix.w_magic) : pix / (unsigned)A0.W);
    const int w = (int)(pix - (unsigned)h * (unsigned)A0.W);
    const int n = j % A0.N;
    int first = 0;                                                       // first link this row takes part in
    bool zero_in = false;
    if (ch.row_start != nullptr) {
        const int s0 = ch.row_start[j] - ch.k0;
        if (s0 >= 0) { first = s0 < ch.n ? s0 : ch.n; zero_in = true; }
    }
    const size_t ms = ((size_t)n * A0.H + h) * A0.W + w;
    auto ldv = [&](const float* q, float* v) {
        if constexpr (VEC == 4) { const float4 t = *reinterpret_cast<const float4*>(q); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; }
        else v[0] = *q;
    };
    auto stv = [&](float* q, const float* v) {
        if constexpr (VEC == 4) *reinterpret_cast<float4*>(q) = make_float4(v[0], v[1], v[2], v[3]);
        else *q = v[0];
    };
    float z[VEC], t[VEC];
#pragma unroll
    for (int q = 0; q < VEC; ++q) z[q] = 0.f;
    if (!zero_in && A0.mfm_c != nullptr) {                               // MFM backward fused into the Split firing (mfm_bwd_kernel)
        const int Cp = A0.C / 2;
        const int ch = c < Cp ? c : c - Cp;                              // a 4-channel group never straddles the halves (Cp % 4 == 0)
        float g[VEC], a[VEC], b[VEC];
        ldv(A0.z_in + (((size_t)j * A0.H + h) * A0.W + w) * Cp + ch, g);
        ldv(A0.mfm_c + ms * A0.C + ch, a);
        ldv(A0.mfm_c + ms * A0.C + Cp + ch, b);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            const float tq = a[q] == b[q] ? __fmul_rn(g[q], 0.5f) : g[q];
            z[q] = c < Cp ? (a[q] < b[q] ? 0.f : tq) : (b[q] < a[q] ? 0.f : tq);
        }
    } else if (!zero_in) {                                               // link 0's gather of the incoming gradient
        const int us = ix.up_shift;
        const bool on = us >= 0 ? (((h | w) & ((1 << us) - 1)) == 0) : (h % A0.up == 0 && w % A0.up == 0);
        if (A0.z_in != nullptr && on) {
            const int Hm = us >= 0 ? A0.H >> us : A0.H / A0.up, Wm = us >= 0 ? A0.W >> us : A0.W / A0.up;
            const int hm = us >= 0 ? h >> us : h / A0.up, wm = us >= 0 ? w >> us : w / A0.up;
            ldv(A0.z_in + (((size_t)j * Hm + hm) * Wm + wm) * A0.zc + c, z);
        }
        if (A0.z_in2 != nullptr && c < A0.c2) {
            const int Hr = A0.H / A0.k2, Wr = A0.W / A0.k2;
            ldv(A0.z_in2 + (((size_t)j * Hr + h / A0.k2) * Wr + w / A0.k2) * A0.c2 + c, t);
            const float kk = (float)(A0.k2 * A0.k2);
#pragma unroll
            for (int q = 0; q < VEC; ++q) z[q] = __fadd_rn(z[q], __fdiv_rn(t[q], kk));
        }
    }
    const size_t e0 = ((size_t)h * A0.W + w) * A0.C + c;
    const size_t off = (size_t)i * VEC;
    // firings before the row's start: its gradient and p are zero there - the tensors that are kept (residual-path gradients are
    // read again after the start) get their zeros, nothing is loaded
    for (int l = 0; l < first; ++l) {                                    // (a pair row of zeros is all zero bits as well)
        if (ch.a[l].P_out != nullptr) stv(ch.a[l].P_out + off, z);
        if (ch.a[l].z_out != nullptr) stv(ch.a[l].z_out + off, z);
    }
    for (int l = first; l < ch.n; ++l) {
        const HookArgs& A = ch.a[l];
        if (!(zero_in && l == first)) {
#pragma unroll
            for (int q = 0; q < VEC; ++q) z[q] = __fmul_rn(z[q], A.pre_scale);
            if (A.pre_scale_row >= 0) {                                  // BatchNorm backward ahead of its hook
                ldv(A.bn + A.pre_scale_row * A.C + c, t);
#pragma unroll
                for (int q = 0; q < VEC; ++q) z[q] = __fmul_rn(z[q], t[q]);
            }
        }
        BnC b[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) b[q] = {0.f, 0.f, 0.f, 0.f};
        if (A.bn != nullptr) {
            float al[VEC], be[VEC], sp[VEC], tp[VEC];
            ldv(A.bn + c, al); ldv(A.bn + A.C + c, be); ldv(A.bn + 2 * A.C + c, sp); ldv(A.bn + 3 * A.C + c, tp);
#pragma unroll
            for (int q = 0; q < VEC; ++q) b[q] = {al[q], be[q], sp[q], tp[q]};
        }
        float v0[VEC], v1[VEC], v2[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) v0[q] = v1[q] = v2[q] = 0.f;
        const int rc = A.recipe;
        if (A.s0 != nullptr && c < A.c0) ldv(A.s0 + ms * A.c0 + c, v0);
        if (rc >= 3) ldv(A.s1 + ms * A.C + c, v1);                       // recipes 3 - 8 read s1 [N,H,W,C]
        if ((rc == 4 || rc == 8) && A.s2 != nullptr && c < A.c2s) ldv(A.s2 + ms * A.c2s + c, v2);
        if (rc == 6) ldv(A.s2 + ms * A.C + c, v2);
        const HookPrior P = hook_prior(A, j);
        const bool has_prior = ((MODE >= 0 ? MODE : A.mode) != XFRB_MODE_NONE) && (j == P.row);
        float pv[VEC], ps[VEC];
#pragma unroll
        for (int q = 0; q < VEC; ++q) ps[q] = 1.f;
        if (A.post_scale_row >= 0) ldv(A.bn + A.post_scale_row * A.C + c, ps);
#pragma unroll
        for (int q = 0; q < VEC; ++q) {
            float a, x;
            hook_ax(rc, v0[q], v1[q], v2[q], b[q], a, x);
            z[q] = hook_fire<MODE>(A, P, has_prior, e0 + q, ps[q], z[q], a, x, pv[q]);
        }
        if (A.P_out != nullptr) stv(A.P_out + off, pv);
        if (A.z_out != nullptr) {
            if constexpr (VEC == 4) {
                if (A.out_pair) st_pair4(A.z_out, ((size_t)j * A0.H + h) * A0.W + w, A0.C, c, z);
                else stv(A.z_out + off, z);
            } else stv(A.z_out + off, z);
        }
        if (A.probe_out != nullptr && j == P.probe_row && P.probe_elem >= (long long)e0 && P.probe_elem < (long long)e0 + VEC)
            *A.probe_out = pv[(int)(P.probe_elem - (long long)e0)];
    }
}

// Row walk: the sweeps of ONE probe (N = 1: layer sweeps, weighted_subtree_ebp) push up to 48 gradient rows through the same saved
// tensors.  Here a thread owns 4 channels of one pixel and walks R gradient rows: (a, x), the BatchNorm constants and the pre / post
// scales of every link are computed once and kept in registers (NL, the chain length, is a template constant so that they stay
// there), and a row costs its gradient load, the hook arithmetic and its stores - about half of what the row-per-thread kernel
// above spends per (row, link, element) (ncu r2u: that kernel is issue-bound at ~210 instructions per thread of a five-link chain).
// Same operations on every element in the same order: bit-identical.
template <int MODE, int NL, int R>
__global__ void __launch_bounds__(256) hook_rows_kernel(HookChain ch, HookIdx ix) {
    const HookArgs& A0 = ch.a[0];
    const unsigned il = blockIdx.x * blockDim.x + threadIdx.x;         // position inside a gradient row
    if (il >= ix.per_row) return;
    const unsigned Cv = (unsigned)A0.C / 4;
    const unsigned pix = ix.cv_shift >= 0 ? il >> ix.cv_shift : il / Cv;
    const int c = (int)(il - pix * Cv) * 4;
    const int h = (int)(ix.w_magic ? __umulhi(pix, ix.w_magic) : pix / (unsigned)A0.W);
    const int w = (int)(pix - (unsigned)h * (unsigned)A0.W);
    const size_t ms = pix;                                              // N == 1: the saved tensors of the one probe
    auto ld4v = [&](const float* q, float* v) { const float4 t = *reinterpret_cast<const float4*>(q); v[0] = t.x; v[1] = t.y; v[2] = t.z; v[3] = t.w; };
    auto st4v = [&](float* q, const float* v) { *reinterpret_cast<float4*>(q) = make_float4(v[0], v[1], v[2], v[3]); };
    // ---- per link, once per thread
    float a[NL][4], x[NL][4], ps[NL][4], pre[NL][4];
    int prow[NL], qrow[NL];
#pragma unroll
    for (int l = 0; l < NL; ++l) {
        const HookArgs& A = ch.a[l];
        BnC b[4];
#pragma unroll
        for (int q = 0; q < 4; ++q) { b[q] = {0.f, 0.f, 0.f, 0.f}; ps[l][q] = 1.f; pre[l][q] = 1.f; }
        if (A.bn != nullptr) {
            float al[4], be[4], sp[4], tp[4];
            ld4v(A.bn + c, al); ld4v(A.bn + A.C + c, be); ld4v(A.bn + 2 * A.C + c, sp); ld4v(A.bn + 3 * A.C + c, tp);
#pragma unroll
            for (int q = 0; q < 4; ++q) b[q] = {al[q], be[q], sp[q], tp[q]};
        }
        if (A.post_scale_row >= 0) ld4v(A.bn + A.post_scale_row * A.C + c, ps[l]);
        if (A.pre_scale_row >= 0) ld4v(A.bn + A.pre_scale_row * A.C + c, pre[l]);
        float v0[4] = {0.f, 0.f, 0.f, 0.f}, v1[4] = {0.f, 0.f, 0.f, 0.f}, v2[4] = {0.f, 0.f, 0.f, 0.f};
        const int rc = A.recipe;
        if (A.s0 != nullptr && c < A.c0) ld4v(A.s0 + ms * A.c0 + c, v0);
        if (rc >= 3) ld4v(A.s1 + ms * A.C + c, v1);
        if ((rc == 4 || rc == 8) && A.s2 != nullptr && c < A.c2s) ld4v(A.s2 + ms * A.c2s + c, v2);
        if (rc == 6) ld4v(A.s2 + ms * A.C + c, v2);
#pragma unroll
        for (int q = 0; q < 4; ++q) hook_ax(rc, v0[q], v1[q], v2[q], b[q], a[l][q], x[l][q]);
        prow[l] = A.ptab != nullptr ? A.ptab->row : A.prior_row;
        qrow[l] = (A.ptab != nullptr && A.probe_out != nullptr) ? A.ptab->probe_row : -1;
    }
    const int us = ix.up_shift;
    const bool on = us >= 0 ? (((h | w) & ((1 << us) - 1)) == 0) : (h % A0.up == 0 && w % A0.up == 0);
    const int Hm = us >= 0 ? A0.H >> us : A0.H / A0.up, Wm = us >= 0 ? A0.W >> us : A0.W / A0.up;
    const int hm = us >= 0 ? h >> us : h / A0.up, wm = us >= 0 ? w >> us : w / A0.up;
    const int Hr = A0.H / A0.k2, Wr = A0.W / A0.k2, hr = h / A0.k2, wr = w / A0.k2;
    const float kk = (float)(A0.k2 * A0.k2);
    const size_t e0 = ((size_t)h * A0.W + w) * A0.C + c;
    // ---- the gradient rows
    const int jend = min(A0.J, (int)(blockIdx.y + 1) * R);
    for (int j = blockIdx.y * R; j < jend; ++j) {
        int first = 0;
        bool zero_in = false;
        if (ch.row_start != nullptr) {
            const int s0 = ch.row_start[j] - ch.k0;
            if (s0 >= 0) { first = s0 < NL ? s0 : NL; zero_in = true; }
        }
        float z[4] = {0.f, 0.f, 0.f, 0.f};
        const size_t off = ((size_t)j * ix.per_row + il) * 4;
        if (!zero_in) {
            if (A0.z_in != nullptr && on) ld4v(A0.z_in + (((size_t)j * Hm + hm) * Wm + wm) * A0.zc + c, z);
            if (A0.z_in2 != nullptr && c < A0.c2) {
                float t[4];
                ld4v(A0.z_in2 + (((size_t)j * Hr + hr) * Wr + wr) * A0.c2 + c, t);
#pragma unroll
                for (int q = 0; q < 4; ++q) z[q] = __fadd_rn(z[q], __fdiv_rn(t[q], kk));
            }
        }
#pragma unroll
        for (int l = 0; l < NL; ++l) {
            const HookArgs& A = ch.a[l];
            if (l < first) {                                             // before the row's start: zeros where a tensor is kept
                if (A.P_out != nullptr) st4v(A.P_out + off, z);
                if (A.z_out != nullptr) st4v(A.z_out + off, z);
                continue;
            }
            if (!(zero_in && l == first)) {
#pragma unroll
                for (int q = 0; q < 4; ++q) z[q] = __fmul_rn(z[q], A.pre_scale);
                if (A.pre_scale_row >= 0) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) z[q] = __fmul_rn(z[q], pre[l][q]);
                }
            }
            HookPrior P;
            P.row = prow[l]; P.probe_row = qrow[l]; P.elem = 0; P.val = 0.f; P.tensor = nullptr; P.probe_elem = -1;
            const bool has_prior = ((MODE >= 0 ? MODE : A.mode) != XFRB_MODE_NONE) && (j == P.row);
            if (has_prior) {
                if (A.ptab != nullptr) { P.elem = A.ptab->elem; P.val = A.ptab->val; P.tensor = A.ptab->tensor; }
                else { P.elem = A.prior_elem; P.val = A.prior_val; P.tensor = A.prior; }
            }
            float pv[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) z[q] = hook_fire<MODE>(A, P, has_prior, e0 + q, ps[l][q], z[q], a[l][q], x[l][q], pv[q]);
            if (A.P_out != nullptr) st4v(A.P_out + off, pv);
            if (A.z_out != nullptr) {
                if (A.out_pair) st_pair4(A.z_out, (size_t)j * ((size_t)A0.H * A0.W) + pix, A0.C, c, z);
                else st4v(A.z_out + off, z);
            }
            if (j == P.probe_row) {
                const long long pe = A.ptab->probe_elem;
                if (pe >= (long long)e0 && pe < (long long)e0 + 4) *A.probe_out = pv[(int)(pe - (long long)e0)];
            }
        }
    }
}

cudaError_t launch_hook_chain(const HookChain& ch, cudaStream_t st) {
    const HookArgs& a0 = ch.a[0];
    if (ch.n < 1 || ch.n > XFRB_MAX_CHAIN) return cudaErrorInvalidValue;
    if ((size_t)a0.J * a0.H * a0.W * a0.C >= 0xFFFFFF00ull) return cudaErrorInvalidValue;          // the kernel indexes with 32-bit arithmetic
    auto al16 = [](const void* p) { return (reinterpret_cast<uintptr_t>(p) & 15u) == 0; };
    if (a0.mfm_c != nullptr && (a0.z_in == nullptr || a0.z_in2 != nullptr || a0.up != 1 || a0.C % 2)) return cudaErrorInvalidValue;
    bool vec = a0.C % 4 == 0 && a0.zc % 4 == 0 && (a0.z_in2 == nullptr || a0.c2 % 4 == 0) && al16(a0.z_in) && al16(a0.z_in2) &&
               (a0.mfm_c == nullptr || (a0.C % 8 == 0 && al16(a0.mfm_c)));
    for (int l = 0; l < ch.n; ++l) {
        const HookArgs& a = ch.a[l];
        if (a.J != a0.J || a.N != a0.N || a.H != a0.H || a.W != a0.W || a.C != a0.C) return cudaErrorInvalidValue;
        vec = vec && (a.s0 == nullptr || a.c0 % 4 == 0) && (a.s2 == nullptr || a.c2s % 4 == 0 || a.recipe == 6) && al16(a.s0) && al16(a.s1) &&
              al16(a.s2) && al16(a.bn) && al16(a.P_out) && al16(a.z_out);
    }
    static const int force_scalar = [] { const char* e = getenv("XFRB_HOOK_SCALAR"); return e ? atoi(e) : 0; }();   // A/B probe
    if (a0.J > 65535) return cudaErrorInvalidValue;                     // grid.y = gradient row
    const int V = (vec && !force_scalar) ? 4 : 1;
    for (int l = 0; l < ch.n; ++l)
        if (ch.a[l].out_pair && V != 4) return cudaErrorInvalidValue;   // pair rows are written four channels at a time
    auto log2_of = [](unsigned v) { int sh = 0; while ((1u << sh) < v) ++sh; return (v > 0 && (1u << sh) == v) ? sh : -1; };
    HookIdx ix;
    const unsigned Cv = (unsigned)a0.C / V, npix = (unsigned)a0.H * a0.W;
    ix.per_row = npix * Cv;
    ix.cv_shift = log2_of(Cv);
    ix.up_shift = log2_of((unsigned)a0.up);
    ix.w_magic = 0;
    if (npix <= 65536 && a0.W > 1 && a0.W <= 256) {                     // reciprocal verified on the host, once per width
        static unsigned char known[257] = {};                            // 0: not checked, 1: exact for every p < 65536, 2: not exact
        const unsigned m = (unsigned)((0x100000000ull + a0.W - 1) / a0.W);
        if (known[a0.W] == 0) {
            bool exact = true;
            for (unsigned p = 0; p < 65536 && exact; ++p) exact = (unsigned)(((unsigned long long)p * m) >> 32) == p / (unsigned)a0.W;
            known[a0.W] = exact ? 1 : 2;
        }
        if (known[a0.W] == 1) ix.w_magic = m;
    }
    int mode = a0.mode;
    for (int l = 1; l < ch.n; ++l)
        if (ch.a[l].mode != mode) mode = -1;
    // row walk (hook_rows_kernel): sweeps of one probe with enough gradient rows, vector path, a chain length it is instantiated for
    static const int rows_on = [] { const char* e = getenv("XFRB_HOOK_ROWS"); return e ? atoi(e) : 1; }();       // A/B probe: 0 = off
    if (rows_on && V == 4 && a0.N == 1 && a0.J >= 8 && a0.mfm_c == nullptr && mode >= 0 && (ch.n == 1 || ch.n == 2 || ch.n == 3 || ch.n == 5)) {
        constexpr int R = 8;
        const dim3 grid_r((ix.per_row + 255) / 256, (unsigned)((a0.J + R - 1) / R));
#define XFRB_ROWS_MODE(NL_)                                                                                  \
        switch (mode) {                                                                                      \
            case XFRB_MODE_AWP: hook_rows_kernel<XFRB_MODE_AWP, NL_, R><<<grid_r, 256, 0, st>>>(ch, ix); break;            \
            case XFRB_MODE_ALL: hook_rows_kernel<XFRB_MODE_ALL, NL_, R><<<grid_r, 256, 0, st>>>(ch, ix); break;            \
            case XFRB_MODE_AFFINEONLY: hook_rows_kernel<XFRB_MODE_AFFINEONLY, NL_, R><<<grid_r, 256, 0, st>>>(ch, ix); break; \
            default: hook_rows_kernel<XFRB_MODE_NONE, NL_, R><<<grid_r, 256, 0, st>>>(ch, ix); break;        \
        }
        if (ch.n == 1) { XFRB_ROWS_MODE(1) } else if (ch.n == 2) { XFRB_ROWS_MODE(2) } else if (ch.n == 3) { XFRB_ROWS_MODE(3) } else { XFRB_ROWS_MODE(5) }
#undef XFRB_ROWS_MODE
        return cudaGetLastError();
    }
    const dim3 grid((ix.per_row + 255) / 256, (unsigned)a0.J);
#define XFRB_HOOK_LAUNCH(V_)                                                                   \
    switch (mode) {                                                                            \
        case XFRB_MODE_AWP: hook_kernel<V_, XFRB_MODE_AWP><<<grid, 256, 0, st>>>(ch, ix); break;            \
        case XFRB_MODE_ALL: hook_kernel<V_, XFRB_MODE_ALL><<<grid, 256, 0, st>>>(ch, ix); break;            \
        case XFRB_MODE_AFFINEONLY: hook_kernel<V_, XFRB_MODE_AFFINEONLY><<<grid, 256, 0, st>>>(ch, ix); break; \
        case XFRB_MODE_NONE: hook_kernel<V_, XFRB_MODE_NONE><<<grid, 256, 0, st>>>(ch, ix); break;          \
        default: hook_kernel<V_, -1><<<grid, 256, 0, st>>>(ch, ix); break;                     \
    }
    if (V == 4) { XFRB_HOOK_LAUNCH(4) } else { XFRB_HOOK_LAUNCH(1) }
#undef XFRB_HOOK_LAUNCH
    return cudaGetLastError();
}

cudaError_t launch_hook(const HookArgs& a, cudaStream_t st) {
    HookChain ch;
    ch.a[0] = a;
    ch.n = 1;
    ch.row_start = nullptr;
    ch.k0 = 0;
    return launch_hook_chain(ch, st);
}

// g = (g - xn*<xn,g>)/nrm per row: the Jacobian of F.normalize (resnet.py:250).  One block per row, D <= 1024.
__global__ void normalize_bwd_kernel(const float* __restrict__ gin, const float* __restrict__ xn, const float* __restrict__ nrm,
                                     float* __restrict__ gout, int N, int D) {
    __shared__ float red[32];
    __shared__ float bc;
    const int j = blockIdx.x, n = j % N, d = threadIdx.x;
    float x = d < D ? xn[(size_t)n * D + d] : 0.f, g = d < D ? gin[(size_t)j * D + d] : 0.f;
    float s = x * g;
    for (int o = 16; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if ((d & 31) == 0) red[d >> 5] = s;
    __syncthreads();
    if (d == 0) {
        float t = 0.f;
        for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) t += red[i];
        bc = t;
    }
    __syncthreads();
    if (d < D) gout[(size_t)j * D + d] = __fdiv_rn(g - x * bc, nrm[n]);
}
cudaError_t launch_normalize_bwd(const float* gin, const float* xn, const float* nrm, float* gout, int J, int N, int D, cudaStream_t st) {
    normalize_bwd_kernel<<<J, ((D + 31) / 32) * 32, 0, st>>>(gin, xn, nrm, gout, N, D);
    return cudaGetLastError();
}

// MaxPool2d(3,2,pad) backward alone (gather form, first maximum wins): g [J,56,56,64] -> out [J,112,112,64]; r1 = relu(bn(o))
__global__ void maxpool_bwd_kernel(const float* __restrict__ g, const float* __restrict__ o, const float* __restrict__ bn,
                                   float* __restrict__ out, int N, int pad, size_t total) {
    size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c = (int)(i & 63);
    size_t p = i >> 6;
    const int w = (int)(p % 112); p /= 112;
    const int h = (int)(p % 112);
    const int j = (int)(p / 112), n = j % N;
    const float al = bn[c], be = bn[64 + c];
    const float* ob = o + (size_t)n * 112 * 112 * 64 + c;
    const float me = fmaxf(__fadd_rn(__fmul_rn(ob[((size_t)h * 112 + w) * 64], al), be), 0.f);
    float z = 0.f;
    for (int a = 0; a < 2; ++a) {
        int ph = (h + pad) / 2 - a;
        if (ph < 0 || ph >= 56 || 2 * ph - pad > h || h > 2 * ph - pad + 2) continue;
        for (int b2 = 0; b2 < 2; ++b2) {
            int pw = (w + pad) / 2 - b2;
            if (pw < 0 || pw >= 56 || 2 * pw - pad > w || w > 2 * pw - pad + 2) continue;
            bool win = true;
            const int my = h - (2 * ph - pad), mx = w - (2 * pw - pad);
            for (int r = 0; r < 3; ++r)
                for (int s2 = 0; s2 < 3; ++s2) {
                    int hh = 2 * ph - pad + r, ww = 2 * pw - pad + s2;
                    if ((r == my && s2 == mx) || hh < 0 || hh >= 112 || ww < 0 || ww >= 112) continue;
                    float rv = fmaxf(__fadd_rn(__fmul_rn(ob[((size_t)hh * 112 + ww) * 64], al), be), 0.f);
                    bool before = (r < my) || (r == my && s2 < mx);
                    win = win && (before ? (rv < me) : (rv <= me));
                }
            if (win) z = __fadd_rn(z, g[(((size_t)j * 56 + ph) * 56 + pw) * 64 + c]);
        }
    }
    out[i] = z;
}
// the same gather from the window arg-max bytes the forward sweep recorded (stem_pool_kernel: 1 byte per pooled element, the
// position 3*r + s of the first maximum): one thread = 4 channels, no window scans (the scan form: 845 us per 48-row launch)
__global__ void __launch_bounds__(256) maxpool_bwd_arg_kernel(const float* __restrict__ g, const unsigned char* __restrict__ mp_arg,
                                                              float* __restrict__ out, int N, int pad) {
    // grid: (112*112*16/256, J)
    const int j = blockIdx.y, n = j % N;
    const int i = blockIdx.x * 256 + threadIdx.x;
    const int c = (i & 15) * 4;
    const int pix = i >> 4;
    const int h = pix / 112, w = pix % 112;
    uchar4 am[4];
    float4 gz[4];
    bool ok[4];
    int me_idx[4];
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        const int a = t >> 1, b2 = t & 1;
        const int ph = (h + pad) / 2 - a, pw = (w + pad) / 2 - b2;
        ok[t] = !(ph < 0 || ph >= 56 || 2 * ph - pad > h || h > 2 * ph - pad + 2) && !(pw < 0 || pw >= 56 || 2 * pw - pad > w || w > 2 * pw - pad + 2);
        const int phc = ok[t] ? ph : 0, pwc = ok[t] ? pw : 0;
        me_idx[t] = (h - (2 * phc - pad)) * 3 + (w - (2 * pwc - pad));
        am[t] = __ldg(reinterpret_cast<const uchar4*>(mp_arg) + (((size_t)n * 56 + phc) * 56 + pwc) * 16 + (c >> 2));
        gz[t] = ld4(g + (((size_t)j * 56 + phc) * 56 + pwc) * 64 + c);
    }
    float z[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
    for (int t = 0; t < 4; ++t) {
        if (!ok[t]) continue;
        if (am[t].x == me_idx[t]) z[0] = __fadd_rn(z[0], gz[t].x);
        if (am[t].y == me_idx[t]) z[1] = __fadd_rn(z[1], gz[t].y);
        if (am[t].z == me_idx[t]) z[2] = __fadd_rn(z[2], gz[t].z);
        if (am[t].w == me_idx[t]) z[3] = __fadd_rn(z[3], gz[t].w);
    }
    st4(out + ((size_t)j * 112 * 112 + pix) * 64 + c, make_float4(z[0], z[1], z[2], z[3]));
}
cudaError_t launch_maxpool_bwd(const float* g, const float* o, const float* bn, float* out, const unsigned char* mp_arg, int J, int N,
                               int pad, cudaStream_t st) {
    if (mp_arg != nullptr && J <= 65535) {
        maxpool_bwd_arg_kernel<<<dim3(112 * 112 * 16 / 256, J), 256, 0, st>>>(g, mp_arg, out, N, pad);
        return cudaGetLastError();
    }
    size_t total = (size_t)J * 112 * 112 * 64;
    maxpool_bwd_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(g, o, bn, out, N, pad, total);
    return cudaGetLastError();
}

// Zero-initialised device scratch of the multi-block reductions below (one allocation per device, made on the first - eager - call:
// cudaMalloc is not capturable; every user leaves it zeroed again, and the users of one device run on one stream).
struct ReduceScratch {
    unsigned long long key;            // subtree_score: running (value, ~index) maximum
    unsigned int ticket;               // blocks of the current launch that have finished
    unsigned int pad;
};
static void* device_scratch(int which, size_t bytes) {
    static void* ptr[2][XFRB_MAX_DEV] = {};
    void*& p = ptr[which][current_device_slot()];
    if (p == nullptr) {
        if (cudaMalloc(&p, bytes) != cudaSuccess) { p = nullptr; return nullptr; }
        cudaMemset(p, 0, bytes);
    }
    return p;
}

// per firing: score = max_e m(e) * (-gn[e]), arg = first argmax; m = (gm >= 0) (mated-similarity gating) or (gce < 0)
// (whitebox.py:687-696).  n elements, many blocks: every block folds its stretch into one 64-bit key - the value's bits made
// monotonic in the high word, ~index in the low word, so that the largest key is the largest value at its FIRST position - and
// atomicMax-es it into the scratch; the last block to finish decodes it.  (The one-block form took 91 us per firing on the
// 56x56x256 tensors: 34 ms of a 243 ms weighted-subtree job, profiles/r2_notes.md.)
__device__ __forceinline__ unsigned long long score_key(float v, unsigned int e) {
    v = __fadd_rn(v, 0.f);                                              // -0 -> +0: the two compare equal, the first position wins
    const unsigned int u = __float_as_uint(v);
    const unsigned int ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
    return ((unsigned long long)ord << 32) | (unsigned long long)(0xFFFFFFFFu - e);
}
__global__ void __launch_bounds__(256) subtree_score_kernel(const float* __restrict__ gate, const float* __restrict__ gneg, int gate_ge0,
                                                            unsigned int n, float* __restrict__ score, long long* __restrict__ arg,
                                                            ReduceScratch* __restrict__ sc) {
    __shared__ unsigned long long sk[8];
    unsigned long long best = 0ull;                                     // below every real key (ord >= 1 for any non-NaN value)
    for (unsigned int e = blockIdx.x * blockDim.x + threadIdx.x; e < n; e += gridDim.x * blockDim.x) {
        const float g = gate[e];
        const float m = gate_ge0 ? (g >= 0.f ? 1.f : 0.f) : (g < 0.f ? 1.f : 0.f);
        const unsigned long long k = score_key(__fmul_rn(m, -gneg[e]), e);
        best = k > best ? k : best;
    }
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long ok = __shfl_xor_sync(0xffffffffu, best, o);
        best = ok > best ? ok : best;
    }
    if ((threadIdx.x & 31) == 0) sk[threadIdx.x >> 5] = best;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int k = 1; k < (int)(blockDim.x >> 5); ++k) best = sk[k] > best ? sk[k] : best;
        atomicMax(&sc->key, best);
        __threadfence();
        if (atomicAdd(&sc->ticket, 1u) == gridDim.x - 1) {              // last block: every key has been folded in
            const unsigned long long k = atomicMax(&sc->key, 0ull);
            const unsigned int ord = (unsigned int)(k >> 32);
            const unsigned int u = (ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord;
            *score = n ? __uint_as_float(u) : -INFINITY;
            *arg = n ? (long long)(0xFFFFFFFFu - (unsigned int)(k & 0xFFFFFFFFull)) : 0x7fffffffffffffffLL;
            sc->key = 0ull;
            sc->ticket = 0u;
            __threadfence();
        }
    }
}
cudaError_t launch_subtree_score(const float* gate, const float* gneg, int gate_ge0, size_t n, float* score, long long* arg,
                                 cudaStream_t st) {
    if (n >= 0xFFFFFFFFull) return cudaErrorInvalidValue;                // 32-bit positions in the key
    ReduceScratch* sc = static_cast<ReduceScratch*>(device_scratch(0, sizeof(ReduceScratch)));
    if (sc == nullptr) return cudaErrorMemoryAllocation;
    unsigned grid = (unsigned)((n + 2047) / 2048);                       // >= 8 elements per thread
    grid = grid < 1 ? 1 : (grid > 592 ? 592 : grid);
    subtree_score_kernel<<<grid, 256, 0, st>>>(gate, gneg, gate_ge0, (unsigned int)n, score, arg, sc);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ truncation threshold
// truncated_contrastive_ebp (whitebox.py:550-554): ascending sort of the mate MWP, cumulative sum, keep the elements whose
// running sum has reached percentile% of the total.  Values are >= 0, so their IEEE bit patterns are ordered like the
// values and the first kept element v* is found by a 3-level radix descent (11 + 11 + 10 bits) over per-bin VALUE SUMS
// accumulated in double: thr[n] = v*, mask = (P >= v*).  One block per sample; three streaming passes over 3.2 MB.
__global__ void __launch_bounds__(1024) trunc_threshold_kernel(const float* __restrict__ P2, const double* __restrict__ sums,
                                                               float pct, float* __restrict__ thr, size_t per_sample) {
    __shared__ double hist[2048];
    __shared__ unsigned int s_prefix;
    __shared__ double s_below;
    const int n = blockIdx.x, tid = threadIdx.x;
    const float* p = P2 + (size_t)n * per_sample;
    const double target = (double)(pct / 100.0f) * sums[n];
    if (tid == 0) { s_prefix = 0u; s_below = 0.0; }
    const int shifts[3] = {21, 10, 0};
    const int widths[3] = {11, 11, 10};
    for (int level = 0; level < 3; ++level) {
        const int nb = 1 << widths[level];
        for (int i = tid; i < nb; i += 1024) hist[i] = 0.0;
        __syncthreads();
        const unsigned int prefix = s_prefix;
        const int sh = shifts[level];
        const int hsh = sh + widths[level];            // bits above this level must equal the prefix
        for (size_t i = tid; i < per_sample / 4; i += 1024) {
            float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
            const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
            for (int e = 0; e < 4; ++e) {
                unsigned int u = __float_as_uint(vv[e]);
                if (vv[e] > 0.f && (level == 0 || (u >> hsh) == (prefix >> hsh)))
                    atomicAdd(&hist[(u >> sh) & (nb - 1)], (double)vv[e]);
            }
        }
        __syncthreads();
        if (tid == 0) {
            double run = s_below;
            int b = 0;
            for (; b < nb; ++b) {
                if (run + hist[b] >= target) break;
                run += hist[b];
            }
            if (b == nb) b = nb - 1;       // rounding: target marginally above the total
            s_below = run;
            s_prefix = prefix | ((unsigned int)b << sh);
        }
        __syncthreads();
    }
    if (tid == 0) thr[n] = (pct <= 0.f) ? 0.f : __uint_as_float(s_prefix);
}

// Many blocks per sample (the layer sweeps call this with ONE sample per firing: one block took 296 us, a third of it thread 0
// walking the 2,048 bins): one launch per radix level; every block histograms its stretch in shared memory, adds its non-empty
// bins to the sample's global histogram, and the last block to finish (ticket) scans the bins in parallel, descends one level
// and leaves histogram and ticket zeroed for the next launch.  State per sample: TruncState + 2,048 doubles of scratch.
struct TruncState {
    unsigned int prefix, ticket;
    double below;
};
constexpr int TRUNC_MAX_SAMPLES = 1024;
template <int LEVEL>
__global__ void __launch_bounds__(1024) trunc_level_kernel(const float* __restrict__ P2, const double* __restrict__ sums, float pct,
                                                           float* __restrict__ thr, size_t per_sample, TruncState* __restrict__ state,
                                                           double* __restrict__ ghist_all) {
    constexpr int SH = LEVEL == 0 ? 21 : (LEVEL == 1 ? 10 : 0);
    constexpr int WIDTH = LEVEL == 2 ? 10 : 11;
    constexpr int NB = 1 << WIDTH;
    constexpr int HSH = SH + WIDTH;                     // bits above this level must equal the prefix
    __shared__ double hist[2048];
    __shared__ double wsum[32];
    __shared__ int s_bin;
    __shared__ bool s_last;
    const int n = blockIdx.y, tid = threadIdx.x;
    const float* p = P2 + (size_t)n * per_sample;
    TruncState* stt = state + n;
    double* ghist = ghist_all + (size_t)n * 2048;
    for (int i = tid; i < NB; i += 1024) hist[i] = 0.0;
    __syncthreads();
    const unsigned int prefix = LEVEL == 0 ? 0u : stt->prefix;
    for (size_t i = (size_t)blockIdx.x * 1024 + tid; i < per_sample / 4; i += (size_t)gridDim.x * 1024) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(p) + i);
        const float vv[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
        for (int e = 0; e < 4; ++e) {
            const unsigned int u = __float_as_uint(vv[e]);
            if (vv[e] > 0.f && (LEVEL == 0 || (u >> HSH) == (prefix >> HSH))) atomicAdd(&hist[(u >> SH) & (NB - 1)], (double)vv[e]);
        }
    }
    __syncthreads();
    if (gridDim.x > 1) {
        for (int i = tid; i < NB; i += 1024)
            if (hist[i] != 0.0) atomicAdd(&ghist[i], hist[i]);
        __threadfence();
        __syncthreads();
        if (tid == 0) s_last = atomicAdd(&stt->ticket, 1u) == gridDim.x - 1;
        __syncthreads();
        if (!s_last) return;
        __threadfence();
        for (int i = tid; i < NB; i += 1024) {
            hist[i] = __ldcg(&ghist[i]);
            ghist[i] = 0.0;                             // ready for the next level / the next call
        }
        __syncthreads();
    }
    // parallel form of: run = below; for b: if (run + hist[b] >= target) break; run += hist[b];   (thread t owns bins 2t, 2t + 1)
    const double target = (double)(pct / 100.0f) * sums[n];
    const double below = LEVEL == 0 ? 0.0 : stt->below;
    const int b0 = 2 * tid, b1 = 2 * tid + 1;
    const double h0 = b0 < NB ? hist[b0] : 0.0, h1 = b1 < NB ? hist[b1] : 0.0;
    double inc = h0 + h1;
    for (int o = 1; o < 32; o <<= 1) {
        const double t = __shfl_up_sync(0xffffffffu, inc, o);
        if ((tid & 31) >= o) inc += t;
    }
    if ((tid & 31) == 31) wsum[tid >> 5] = inc;
    if (tid == 0) s_bin = NB;
    __syncthreads();
    if (tid < 32) {
        double w = wsum[tid];
        for (int o = 1; o < 32; o <<= 1) {
            const double t = __shfl_up_sync(0xffffffffu, w, o);
            if (tid >= o) w += t;
        }
        wsum[tid] = w;                                   // inclusive over warps
    }
    __syncthreads();
    const double excl0 = below + ((tid >> 5) ? wsum[(tid >> 5) - 1] : 0.0) + (inc - (h0 + h1));
    const double excl1 = excl0 + h0;
    if (b0 < NB && excl0 + h0 >= target) atomicMin(&s_bin, b0);
    else if (b1 < NB && excl1 + h1 >= target) atomicMin(&s_bin, b1);
    __syncthreads();
    int b = s_bin;
    const bool none = b == NB;                           // rounding: target marginally above the total - every bin is below
    if (none) b = NB - 1;
    if (b == b0 || b == b1) {
        const unsigned int np = prefix | ((unsigned int)b << SH);
        if (LEVEL == 2) {
            thr[n] = (pct <= 0.f) ? 0.f : __uint_as_float(np);
            stt->prefix = 0u;
            stt->below = 0.0;
        } else {
            stt->prefix = np;
            stt->below = (b == b0 ? excl0 : excl1) + (none ? (b == b0 ? h0 : h1) : 0.0);
        }
        stt->ticket = 0u;
    }
}

cudaError_t launch_trunc_threshold(const float* P2, const double* sums, float pct, float* thr, int N, size_t per_sample,
                                   cudaStream_t st) {
    if (per_sample % 4) return cudaErrorInvalidValue;
    static const int single = [] { const char* e = getenv("XFRB_TRUNC_SINGLE"); return e ? atoi(e) : 0; }();      // A/B probe: the one-block form
    constexpr size_t BYTES = TRUNC_MAX_SAMPLES * (sizeof(TruncState) + 2048 * sizeof(double));
    char* sc = (N <= TRUNC_MAX_SAMPLES && !single) ? static_cast<char*>(device_scratch(1, BYTES)) : nullptr;
    if (sc == nullptr) {
        trunc_threshold_kernel<<<N, 1024, 0, st>>>(P2, sums, pct, thr, per_sample);
        return cudaGetLastError();
    }
    TruncState* state = reinterpret_cast<TruncState*>(sc + (size_t)TRUNC_MAX_SAMPLES * 2048 * sizeof(double));
    double* ghist = reinterpret_cast<double*>(sc);
    const size_t iters = (per_sample / 4 + 1023) / 1024;               // float4 trips of one block over the whole sample
    int bps = (int)((iters + 3) / 4);                                   // >= 4 trips per block
    const int cap = 296 / N > 1 ? 296 / N : 1;                          // about two blocks per SM over all samples
    bps = bps < 1 ? 1 : (bps > cap ? cap : bps);
    const dim3 grid((unsigned)bps, (unsigned)N);
    trunc_level_kernel<0><<<grid, 1024, 0, st>>>(P2, sums, pct, thr, per_sample, state, ghist);
    trunc_level_kernel<1><<<grid, 1024, 0, st>>>(P2, sums, pct, thr, per_sample, state, ghist);
    trunc_level_kernel<2><<<grid, 1024, 0, st>>>(P2, sums, pct, thr, per_sample, state, ghist);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ contrastive combine
// out[n,pix] = sum_c relu(k*P2[n,pix,c]/S[n] - k*P2[N+n,pix,c]/S[N+n]),  k = (P2[n,pix,c] >= thr[n]) or 1    (whitebox.py:524-526, 556)
__global__ void contrast_kernel(const float* __restrict__ P2, const double* __restrict__ sums, const float* __restrict__ thr,
                                float* __restrict__ out, int N, int HW, int c4_shift) {
    // C/4 lanes per pixel (C = 64 -> 16 lanes), a power of two <= 32; grid.y = sample: no runtime integer division in the kernel
    const int C4 = 1 << c4_shift;
    const unsigned per = (unsigned)HW << c4_shift;                     // float4 groups per sample
    const unsigned il = blockIdx.x * blockDim.x + threadIdx.x;
    const unsigned n = blockIdx.y;
    float s = 0.f;
    if (il < per) {
        // the reference divides in float by float(sum): a / sm is taken as float(double(a) * (1.0 / sm)) - the double reciprocal is
        // good to 2^-53, so the result is the correctly rounded float quotient unless it lies within 2^-52 of a rounding boundary
        // (one element in 2^28) - three instructions instead of the ~50 of an IEEE float division: the kernel was 91 % issue-active
        // on eight of those per thread (ncu r2ah)
        const double rm = 1.0 / (double)(float)sums[n], rn = 1.0 / (double)(float)sums[N + n];
        const float t = thr ? thr[n] : -1.f;
        const size_t i = (size_t)n * per + il;
        const float4 a = reinterpret_cast<const float4*>(P2)[i];
        const float4 b = reinterpret_cast<const float4*>(P2)[i + (size_t)N * per];
        const float av[4] = {a.x, a.y, a.z, a.w}, bv[4] = {b.x, b.y, b.z, b.w};
#pragma unroll
        for (int e = 0; e < 4; ++e)
            if (av[e] >= t) s += fmaxf(__fsub_rn((float)((double)av[e] * rm), (float)((double)bv[e] * rn)), 0.f);
    }
    for (int o = C4 / 2; o > 0; o >>= 1) s += __shfl_xor_sync(0xffffffffu, s, o);
    if (il < per && (threadIdx.x & (C4 - 1)) == 0) out[(size_t)n * HW + (il >> c4_shift)] = s;
}

cudaError_t launch_contrast(const float* P2, const double* sums, const float* thr, float* out, int N, int HW, int C,
                            cudaStream_t st) {
    int C4 = C / 4;
    if (C4 > 32 || C4 < 1 || (C4 & (C4 - 1)) || N > 65535) return cudaErrorInvalidValue;
    int sh = 0;
    while ((1 << sh) < C4) ++sh;
    const size_t per = (size_t)HW * C4;
    if (per >= 0x7FFFFF00ull) return cudaErrorInvalidValue;      // 32-bit index arithmetic in the kernel
    contrast_kernel<<<dim3((unsigned)((per + 255) / 256), (unsigned)N), 256, 0, st>>>(P2, sums, thr, out, N, HW, sh);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ saliency post-filter
// scipy.ndimage.gaussian_filter(sigma=2, mode='nearest', truncate=4): 17 taps, axis 0 then axis 1, each pass
// accumulated in double and stored as float32; then max(0,.) and division by max(sum, eps)  (whitebox.py:455-460)
__global__ void __launch_bounds__(512) saliency_post_kernel(const float* __restrict__ mwp, float* __restrict__ out, int H, int W,
                                                            float eps) {
    extern __shared__ float sm[];
    float* a = sm;             // [H*W]
    float* b = sm + H * W;     // [H*W]
    __shared__ double wts[17];
    __shared__ double red[16];
    const int tid = threadIdx.x;
    const float* src = mwp + (size_t)blockIdx.x * H * W;
    if (tid == 0) {
        double s = 0.0;
        for (int i = -8; i <= 8; ++i) { wts[i + 8] = exp(-0.5 / 4.0 * (double)(i * i)); s += wts[i + 8]; }
        for (int i = 0; i < 17; ++i) wts[i] /= s;
    }
    for (int i = tid; i < H * W; i += blockDim.x) a[i] = src[i];
    __syncthreads();
    for (int i = tid; i < H * W; i += blockDim.x) {        // axis 0 (rows)
        int h = i / W, w = i % W;
        double s = 0.0;
        for (int t = -8; t <= 8; ++t) {
            int hh = min(max(h + t, 0), H - 1);
            s += wts[t + 8] * (double)a[hh * W + w];
        }
        b[i] = (float)s;
    }
    __syncthreads();
    double part = 0.0;
    for (int i = tid; i < H * W; i += blockDim.x) {        // axis 1 (columns)
        int h = i / W, w = i % W;
        double s = 0.0;
        for (int t = -8; t <= 8; ++t) {
            int ww = min(max(w + t, 0), W - 1);
            s += wts[t + 8] * (double)b[h * W + ww];
        }
        float v = fmaxf((float)s, 0.f);
        a[i] = v;
        part += (double)v;
    }
    for (int o = 16; o > 0; o >>= 1) part += __shfl_xor_sync(0xffffffffu, part, o);
    if ((tid & 31) == 0) red[tid >> 5] = part;
    __syncthreads();
    double tot = 0.0;
    for (int i = 0; i < (int)(blockDim.x >> 5); ++i) tot += red[i];
    float denom = fmaxf((float)tot, eps);
    float* dst = out + (size_t)blockIdx.x * H * W;
    for (int i = tid; i < H * W; i += blockDim.x) dst[i] = __fdiv_rn(a[i], denom);
}

cudaError_t launch_saliency_post(const float* mwp, float* out, int B, int H, int W, float eps, cudaStream_t st) {
    size_t smem = (size_t)2 * H * W * sizeof(float);
    static bool attr[XFRB_MAX_DEV] = {};
    const int dev = current_device_slot();
    if (!attr[dev]) {
        cudaFuncSetAttribute(saliency_post_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 2 * 128 * 128 * 4);
        attr[dev] = true;
    }
    saliency_post_kernel<<<B, 512, smem, st>>>(mwp, out, H, W, eps);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ saliency .npz format
// show.processSaliency (show.py:131-137): attMap -= min; attMap /= (max + 1e-9); skimage.transform.resize(attMap, img.shape[:2],
// order=3, mode='constant') - the cubic resize that turns the 112x112 maps into the 224x224 maps the evaluation stores.
// scikit-image >= 0.19 (skimage/transform/_warps.py, resize(): no anti-aliasing when up-scaling) evaluates it as
//     scipy.ndimage.zoom(attMap, zoom, order=3, mode='grid-constant', cval=0, grid_mode=True)  followed by a clip to the input range,
// i.e. (scipy/ndimage/_interpolation.py zoom(), src/ni_splines.c, src/ni_interpolation.c NI_ZoomShift):
//   1. pad by 12 with zeros (_prepad_for_spline_filter), 2. cubic B-spline prefilter per axis in double - gain (1-z)(1-1/z), pole
//   z = sqrt(3) - 2, mirror initial conditions, axis 0 first -, 3. output pixel i samples at (i + 0.5) * in/out - 0.5 (+ 12) with
//   the four B-spline weights of get_spline_interpolation_weights, 4. cast to the input dtype (float32), clip.
// One CTA per map; the padded coefficient image lives in shared memory as doubles ((h+24) x (w+24) x 8 B <= 227 KB: h, w <= 144).
// tests/test_inpaintgame.py pins this kernel to scipy.ndimage.zoom (7e-15 for the restatement, fp32 rounding for the kernel).
__global__ void __launch_bounds__(512) cubic_zoom_kernel(const float* __restrict__ in, float* __restrict__ out, int h, int w, int oh,
                                                         int ow, int normalize) {
    extern __shared__ double zs[];
    __shared__ float red_lo[16], red_hi[16];
    constexpr int NP = 12;
    const int ph = h + 2 * NP, pw = w + 2 * NP;
    const float* src = in + (size_t)blockIdx.x * h * w;
    const int tid = threadIdx.x, nt = blockDim.x;
    // range of the map (processSaliency's min-shift / max-normalise, and the clip range of resize)
    float lo = INFINITY, hi = -INFINITY;
    for (int i = tid; i < h * w; i += nt) { const float v = src[i]; lo = fminf(lo, v); hi = fmaxf(hi, v); }
    for (int o = 16; o > 0; o >>= 1) { lo = fminf(lo, __shfl_xor_sync(0xffffffffu, lo, o)); hi = fmaxf(hi, __shfl_xor_sync(0xffffffffu, hi, o)); }
    if ((tid & 31) == 0) { red_lo[tid >> 5] = lo; red_hi[tid >> 5] = hi; }
    __syncthreads();
    lo = red_lo[0]; hi = red_hi[0];
    for (int i = 1; i < (nt >> 5); ++i) { lo = fminf(lo, red_lo[i]); hi = fmaxf(hi, red_hi[i]); }
    float scale_den = 1.f, clip_lo = lo, clip_hi = hi;
    if (normalize) {                                  // float32 arithmetic, as numpy does on a float32 map
        scale_den = __fadd_rn(__fsub_rn(hi, lo), 1e-9f);
        clip_lo = 0.f;
        clip_hi = __fdiv_rn(__fsub_rn(hi, lo), scale_den);
    }
    for (int i = tid; i < ph * pw; i += nt) {
        const int y = i / pw - NP, x = i % pw - NP;
        double v = 0.0;
        if (y >= 0 && y < h && x >= 0 && x < w) {
            const float f = src[y * w + x];
            v = (double)(normalize ? __fdiv_rn(__fsub_rn(f, lo), scale_den) : f);
        }
        zs[i] = v;
    }
    __syncthreads();
    const double z = sqrt(3.0) - 2.0, gain = (1.0 - z) * (1.0 - 1.0 / z);
    // prefilter: axis 0 (columns of the padded image, stride pw), then axis 1 (rows, stride 1); one thread per line
    for (int axis = 0; axis < 2; ++axis) {
        const int n = axis == 0 ? ph : pw, lines = axis == 0 ? pw : ph;
        const int step = axis == 0 ? pw : 1, lstep = axis == 0 ? 1 : pw;
        const double z_n_1 = pow(z, (double)(n - 1));
        for (int l = tid; l < lines; l += nt) {
            double* c = zs + (size_t)l * lstep;
            for (int i = 0; i < n; ++i) c[i * step] *= gain;
            double z_i = z, c0 = c[0] + z_n_1 * c[(n - 1) * step];
            for (int i = 1; i < n - 1; ++i) { c0 += z_i * (c[i * step] + z_n_1 * c[(n - 1 - i) * step]); z_i *= z; }
            c[0] = c0 / (1.0 - z_n_1 * z_n_1);
            for (int i = 1; i < n; ++i) c[i * step] += z * c[(i - 1) * step];
            c[(n - 1) * step] = (z / (z * z - 1.0)) * (c[(n - 1) * step] + z * c[(n - 2) * step]);
            for (int i = n - 2; i >= 0; --i) c[i * step] = z * (c[(i + 1) * step] - c[i * step]);
        }
        __syncthreads();
    }
    float* dst = out + (size_t)blockIdx.x * oh * ow;
    for (int i = tid; i < oh * ow; i += nt) {
        const int oy = i / ow, ox = i % ow;
        double wy[4], wx[4];
        int y0, x0;
        {
            const double c = ((double)oy + 0.5) * (double)h / (double)oh - 0.5 + NP, f = floor(c), t = c - f, t1 = 1.0 - t;
            wy[1] = (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0; wy[2] = (t1 * t1 * (t1 - 2.0) * 3.0 + 4.0) / 6.0; wy[0] = t1 * t1 * t1 / 6.0;
            wy[3] = 1.0 - wy[0] - wy[1] - wy[2];
            y0 = (int)f - 1;
        }
        {
            const double c = ((double)ox + 0.5) * (double)w / (double)ow - 0.5 + NP, f = floor(c), t = c - f, t1 = 1.0 - t;
            wx[1] = (t * t * (t - 2.0) * 3.0 + 4.0) / 6.0; wx[2] = (t1 * t1 * (t1 - 2.0) * 3.0 + 4.0) / 6.0; wx[0] = t1 * t1 * t1 / 6.0;
            wx[3] = 1.0 - wx[0] - wx[1] - wx[2];
            x0 = (int)f - 1;
        }
        double acc = 0.0;
#pragma unroll
        for (int a = 0; a < 4; ++a)
#pragma unroll
            for (int b = 0; b < 4; ++b) acc += wy[a] * wx[b] * zs[(y0 + a) * pw + x0 + b];
        dst[i] = fminf(fmaxf((float)acc, clip_lo), clip_hi);
    }
}

cudaError_t launch_cubic_zoom(const float* in, float* out, int B, int h, int w, int oh, int ow, int normalize, cudaStream_t st) {
    const size_t smem = (size_t)(h + 24) * (w + 24) * sizeof(double);
    if (smem > 226 * 1024 || B < 1) return cudaErrorInvalidValue;      // 227 KB per CTA minus the kernel's static arrays
    static size_t attr[XFRB_MAX_DEV] = {};            // largest dynamic shared-memory size opted into so far, per device
    const int dev = current_device_slot();
    if (smem > attr[dev]) {
        cudaError_t e = cudaFuncSetAttribute(cubic_zoom_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
        attr[dev] = smem;
    }
    cubic_zoom_kernel<<<B, 512, smem, st>>>(in, out, h, w, oh, ow, normalize);
    return cudaGetLastError();
}

// ------------------------------------------------------------------ inpainting-game blends
// blend[k,h,w,c] = (1 - m) * orig[c,h,w] + m * inp[c,h,w] evaluated in double and rounded once to fp32 - the float64 blends of
// inpainting_game.py:124-132 followed by the .float() of Whitebox.embeddings (whitebox.py:762) - with m = (value[h,w] > thr[k])
// (inpainting_game.py:66: the K masks are never materialised) or, when `masks` is given, the explicit (blurred) mask
// m = masks[k,h,w].  One thread = 4 consecutive pixels of one blend: both images (double, CHW = the reference's network format,
// 1.2 MB each: L2-resident across the K blends) are read as two double2 per channel plane and the blend is written NHWC, C float4
// per thread, ready for the stem kernel.
template <int C>
__global__ void __launch_bounds__(256) twin_blend_kernel(const double* __restrict__ orig, const double* __restrict__ inp,
                                                         const double* __restrict__ value, const double* __restrict__ thr,
                                                         const double* __restrict__ masks, float* __restrict__ out, int HW4, int mask_f32) {
    const int i = blockIdx.x * 256 + threadIdx.x;        // pixel quad inside the image
    if (i >= HW4) return;
    const int k = blockIdx.y;
    const size_t HW = (size_t)HW4 * 4;
    double m[4];
    if (masks != nullptr) {
        const double2* mp = reinterpret_cast<const double2*>(masks + (size_t)k * HW) + (size_t)i * 2;
        const double2 a = mp[0], b = mp[1];
        m[0] = a.x; m[1] = a.y; m[2] = b.x; m[3] = b.y;
    } else {
        const double t = thr[k];
        const double2* vp = reinterpret_cast<const double2*>(value) + (size_t)i * 2;
        const double2 a = __ldg(vp), b = __ldg(vp + 1);
        m[0] = a.x > t ? 1.0 : 0.0; m[1] = a.y > t ? 1.0 : 0.0; m[2] = b.x > t ? 1.0 : 0.0; m[3] = b.y > t ? 1.0 : 0.0;
    }
    // numpy evaluates `1.0 - masks` in the masks' own dtype: float32 masks (a float32 saliency map, blurred) round it to fp32
    double w[4];
#pragma unroll
    for (int q = 0; q < 4; ++q) w[q] = mask_f32 ? (double)__fsub_rn(1.0f, (float)m[q]) : __dsub_rn(1.0, m[q]);
    float f[4 * C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double2* op = reinterpret_cast<const double2*>(orig + (size_t)c * HW) + (size_t)i * 2;
        const double2* pp = reinterpret_cast<const double2*>(inp + (size_t)c * HW) + (size_t)i * 2;
        const double2 o0 = __ldg(op), o1 = __ldg(op + 1), p0 = __ldg(pp), p1 = __ldg(pp + 1);
        const double ov[4] = {o0.x, o0.y, o1.x, o1.y}, pv[4] = {p0.x, p0.y, p1.x, p1.y};
#pragma unroll
        for (int q = 0; q < 4; ++q)
            f[q * C + c] = __double2float_rn(__dadd_rn(__dmul_rn(w[q], ov[q]), __dmul_rn(m[q], pv[q])));
    }
    float4* dst = reinterpret_cast<float4*>(out + ((size_t)k * HW + (size_t)i * 4) * C);
#pragma unroll
    for (int c = 0; c < C; ++c) dst[c] = make_float4(f[4 * c], f[4 * c + 1], f[4 * c + 2], f[4 * c + 3]);
}

cudaError_t launch_twin_blends(const double* orig, const double* inp, const double* value, const double* thr, const double* masks,
                               float* out, int K, int C, int H, int W, int mask_f32, cudaStream_t st) {
    const size_t HW = (size_t)H * W;
    if (K <= 0 || K > 65535 || HW % 4 != 0 || (masks == nullptr && (value == nullptr || thr == nullptr))) return cudaErrorInvalidValue;
    const int HW4 = (int)(HW / 4);
    dim3 grid((unsigned)((HW4 + 255) / 256), (unsigned)K);
    if (C == 3) twin_blend_kernel<3><<<grid, 256, 0, st>>>(orig, inp, value, thr, masks, out, HW4, mask_f32);
    else if (C == 1) twin_blend_kernel<1><<<grid, 256, 0, st>>>(orig, inp, value, thr, masks, out, HW4, mask_f32);
    else return cudaErrorInvalidValue;
    return cudaGetLastError();
}

}  // namespace xfrb
