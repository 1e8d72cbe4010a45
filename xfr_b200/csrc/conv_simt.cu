// fp32 CUDA-core implicit-GEMM convolution (impl 0): the exact-arithmetic path.
//   D[m, n] = sum_k A[m, k] * B[n, k],   A gathered on the fly from an NHWC tensor (R x R taps,
//   stride 1, zero padding R/2), B stored K-major.  128 x BN x 16 tiles, 256 threads, 8 x (BN/16)
//   register tile per thread, double-buffered shared memory.  Epilogues in common.cuh.
#include "common.cuh"

namespace xfrb {

constexpr int BM = 128;
constexpr int BK = 16;
constexpr int LDA = BM + 4;   // padded leading dimensions (transposed tiles: [k][m])

template <int BN>
__global__ void __launch_bounds__(256, 2)
conv_simt_kernel(const float* __restrict__ A, const float* __restrict__ B, ConvGeom g, EpiParams ep) {
    constexpr int LDB = BN + 4;
    constexpr int TN = BN / 16;          // columns per thread: 8 (BN=128) or 4 (BN=64)
    constexpr int NG = TN / 4;           // float4 column groups per thread
    __shared__ __align__(16) float As[2][BK][LDA];
    __shared__ __align__(16) float Bs[2][BK][LDB];

    const int tid = threadIdx.x;
    const int tx = tid & 15, ty = tid >> 4;
    const int m0 = blockIdx.x * BM;
    const int n0 = blockIdx.y * BN;
    const int HW = g.H * g.W;
    const int pad = g.R >> 1;

    // loader mapping: each thread moves float4s along k; row = tid/4 (+64), kq = tid%4
    const int lrow = tid >> 2, lkq = tid & 3;
    int ah[2], aw[2];
    const float* abase[2];
    bool avalid[2];
#pragma unroll
    for (int i = 0; i < 2; ++i) {
        int m = m0 + lrow + i * 64;
        avalid[i] = m < ep.M;
        int mm = avalid[i] ? m : 0;
        int n = mm / HW, rem = mm - n * HW;
        ah[i] = rem / g.W;
        aw[i] = rem - ah[i] * g.W;
        abase[i] = A + (size_t)mm * g.Cin;
    }
    const bool bvalid0 = (n0 + lrow) < g.Nn, bvalid1 = (n0 + lrow + 64) < g.Nn;
    const float* bbase0 = B + (size_t)(n0 + lrow) * g.K;
    const float* bbase1 = B + (size_t)(n0 + lrow + 64) * g.K;

    float acc[8][TN];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < TN; ++j) acc[i][j] = 0.f;

    float4 ra[2], rb[2];
    auto gload = [&](int k0) {
        int tap = k0 / g.Cin;
        int c0 = k0 - tap * g.Cin + lkq * 4;
        int dr = tap / g.R - pad, ds = tap % g.R - pad;
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            int hh = ah[i] + dr, ww = aw[i] + ds;
            bool ok = avalid[i] && hh >= 0 && hh < g.H && ww >= 0 && ww < g.W;
            ra[i] = ok ? __ldg(reinterpret_cast<const float4*>(abase[i] + ((ptrdiff_t)dr * g.W + ds) * g.Cin + c0))
                       : make_float4(0.f, 0.f, 0.f, 0.f);
        }
        rb[0] = bvalid0 ? __ldg(reinterpret_cast<const float4*>(bbase0 + k0 + lkq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
        if (BN == 128)
            rb[1] = bvalid1 ? __ldg(reinterpret_cast<const float4*>(bbase1 + k0 + lkq * 4)) : make_float4(0.f, 0.f, 0.f, 0.f);
    };
    auto sstore = [&](int buf) {
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            As[buf][lkq * 4 + 0][lrow + i * 64] = ra[i].x;
            As[buf][lkq * 4 + 1][lrow + i * 64] = ra[i].y;
            As[buf][lkq * 4 + 2][lrow + i * 64] = ra[i].z;
            As[buf][lkq * 4 + 3][lrow + i * 64] = ra[i].w;
        }
        Bs[buf][lkq * 4 + 0][lrow] = rb[0].x;
        Bs[buf][lkq * 4 + 1][lrow] = rb[0].y;
        Bs[buf][lkq * 4 + 2][lrow] = rb[0].z;
        Bs[buf][lkq * 4 + 3][lrow] = rb[0].w;
        if (BN == 128) {
            Bs[buf][lkq * 4 + 0][lrow + 64] = rb[1].x;
            Bs[buf][lkq * 4 + 1][lrow + 64] = rb[1].y;
            Bs[buf][lkq * 4 + 2][lrow + 64] = rb[1].z;
            Bs[buf][lkq * 4 + 3][lrow + 64] = rb[1].w;
        }
    };

    const int nk = g.K / BK;
    gload(0);
    sstore(0);
    __syncthreads();
    for (int kt = 0; kt < nk; ++kt) {
        const int buf = kt & 1;
        if (kt + 1 < nk) gload((kt + 1) * BK);
#pragma unroll
        for (int k = 0; k < BK; ++k) {
            float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
            float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
            float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            float bv[TN];
#pragma unroll
            for (int q = 0; q < NG; ++q) {
                float4 b4 = *reinterpret_cast<const float4*>(&Bs[buf][k][q * (BN / 2) + tx * 4]);
                bv[q * 4 + 0] = b4.x; bv[q * 4 + 1] = b4.y; bv[q * 4 + 2] = b4.z; bv[q * 4 + 3] = b4.w;
            }
#pragma unroll
            for (int i = 0; i < 8; ++i)
#pragma unroll
                for (int j = 0; j < TN; ++j) acc[i][j] = fmaf(av[i], bv[j], acc[i][j]);
        }
        if (kt + 1 < nk) {
            sstore(buf ^ 1);
            __syncthreads();
        }
    }

    // ---- epilogue: thread owns rows {ty*4+i, 64+ty*4+i} x column groups {q*(BN/2) + tx*4 .. +3}
    if (ep.kind == EPI_FWD_DUAL) {
        // BN == tn == 128: columns [0,64) = W rows of channels tile*64.., [64,128) = relu(W) twins
        const int c = blockIdx.y * (BN / 2) + tx * 4;
        float4 bt = ld4(ep.bias + n0 + tx * 4);
        float4 bp = ld4(ep.bias + n0 + BN / 2 + tx * 4);
#pragma unroll
        for (int i = 0; i < 8; ++i) {
            int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
            if (m < ep.M)
                epilogue4(ep, m, c, make_float4(acc[i][0], acc[i][1], acc[i][2], acc[i][3]),
                          make_float4(acc[i][TN - 4], acc[i][TN - 3], acc[i][TN - 2], acc[i][TN - 1]), bt, bp);
        }
    } else {
        const float4 z4 = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int q = 0; q < NG; ++q) {
            const int c = n0 + q * (BN / 2) + tx * 4;
            if (c >= g.Nn) continue;
            float4 bt = z4;
            if (ep.kind == EPI_PLAIN && ep.bias != nullptr) bt = ld4(ep.bias + c);
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                int m = m0 + (i >> 2) * 64 + ty * 4 + (i & 3);
                if (m < ep.M) {
                    float4 a4 = make_float4(acc[i][q * 4 + 0] + bt.x, acc[i][q * 4 + 1] + bt.y,
                                            acc[i][q * 4 + 2] + bt.z, acc[i][q * 4 + 3] + bt.w);
                    epilogue4(ep, m, c, a4, z4, z4, z4);
                }
            }
        }
    }
}

cudaError_t launch_conv_simt(const float* A, const float* B, const ConvGeom& g, const EpiParams& ep, cudaStream_t st) {
    if (g.K % BK != 0 || g.Cin % 16 != 0) return cudaErrorInvalidValue;
    dim3 block(256);
    if (ep.kind == EPI_FWD_DUAL || g.Nn % 128 == 0) {
        if (g.Nn % 128 != 0) return cudaErrorInvalidValue;
        dim3 grid((ep.M + BM - 1) / BM, g.Nn / 128);
        conv_simt_kernel<128><<<grid, block, 0, st>>>(A, B, g, ep);
    } else {
        if (g.Nn % 64 != 0) return cudaErrorInvalidValue;
        dim3 grid((ep.M + BM - 1) / BM, g.Nn / 64);
        conv_simt_kernel<64><<<grid, block, 0, st>>>(A, B, g, ep);
    }
    return cudaGetLastError();
}

}  // namespace xfrb
