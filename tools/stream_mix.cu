// What HBM3e sustains on a B200 for the stream MIX of the EBP epilogues: R read streams + W write streams of 256 MB each, float4 per
// thread, fully coalesced, grid-stride - no compute.  The copy figure of MEASURED_PEAKS.json is the (1, 1) corner; the JOIN epilogue is
// (5, 2) (y1 through TMA, o3, xr3, out, g_res / g_out, y3_out), the forward dual epilogue (2, 3) or (2, 4).
//   nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o tools/stream_mix tools/stream_mix.cu && tools/stream_mix
#include <cuda_runtime.h>
#include <stdio.h>

template <int R, int W>
__global__ void __launch_bounds__(256) mix_kernel(const float4* __restrict__ const* in, float4* const* out, size_t n4) {
    const float4* ip[R];
    float4* op[W];
#pragma unroll
    for (int r = 0; r < R; ++r) ip[r] = in[r];
#pragma unroll
    for (int w = 0; w < W; ++w) op[w] = out[w];
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n4; i += (size_t)gridDim.x * blockDim.x) {
        float4 s = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
        for (int r = 0; r < R; ++r) { const float4 v = __ldg(ip[r] + i); s.x += v.x; s.y += v.y; s.z += v.z; s.w += v.w; }
#pragma unroll
        for (int w = 0; w < W; ++w) { op[w][i] = s; s.x += 1.f; }
    }
}

template <int R, int W>
static void run(float4** d_in, float4** d_out, size_t n4, int sms) {
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0);
    cudaEventCreate(&e1);
    float best = 1e30f;
    for (int rep = 0; rep < 5; ++rep) {
        cudaEventRecord(e0);
        mix_kernel<R, W><<<sms * 8, 256>>>((const float4* const*)d_in, d_out, n4);
        cudaEventRecord(e1);
        cudaEventSynchronize(e1);
        float ms;
        cudaEventElapsedTime(&ms, e0, e1);
        if (rep && ms < best) best = ms;
    }
    printf("{\"reads\": %d, \"writes\": %d, \"gbs\": %.0f, \"ms\": %.3f}\n", R, W, (R + W) * n4 * 16.0 / (best * 1e-3) / 1e9, best);
}

int main() {
    const size_t n4 = (size_t)16 << 20;      // 16 Mi float4 = 256 MB per stream
    int sms = 0;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0);
    float4* h_in[6];
    float4* h_out[4];
    for (int i = 0; i < 6; ++i) { cudaMalloc(&h_in[i], n4 * 16); cudaMemset(h_in[i], 0, n4 * 16); }
    for (int i = 0; i < 4; ++i) cudaMalloc(&h_out[i], n4 * 16);
    float4 **d_in, **d_out;
    cudaMalloc(&d_in, sizeof(h_in));
    cudaMalloc(&d_out, sizeof(h_out));
    cudaMemcpy(d_in, h_in, sizeof(h_in), cudaMemcpyHostToDevice);
    cudaMemcpy(d_out, h_out, sizeof(h_out), cudaMemcpyHostToDevice);
    run<1, 1>(d_in, d_out, n4, sms);
    run<2, 1>(d_in, d_out, n4, sms);
    run<4, 1>(d_in, d_out, n4, sms);
    run<6, 1>(d_in, d_out, n4, sms);
    run<1, 2>(d_in, d_out, n4, sms);
    run<1, 3>(d_in, d_out, n4, sms);
    run<2, 3>(d_in, d_out, n4, sms);
    run<2, 4>(d_in, d_out, n4, sms);
    run<3, 2>(d_in, d_out, n4, sms);
    run<5, 2>(d_in, d_out, n4, sms);
    run<4, 2>(d_in, d_out, n4, sms);
    return 0;
}
