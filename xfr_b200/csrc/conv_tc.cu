// tcgen05 implicit-GEMM convolution (impl 1 = 3xTF32, impl 2 = TF32) -- placeholder until the kernel lands.
#include "common.cuh"
namespace xfrb {
bool conv_tc_available() { return false; }
cudaError_t launch_conv_tc(const float*, const float*, const ConvGeom&, const EpiParams&, int, cudaStream_t) {
    return cudaErrorNotSupported;
}
}  // namespace xfrb
