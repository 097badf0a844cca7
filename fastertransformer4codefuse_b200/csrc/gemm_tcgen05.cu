// placeholder until the tcgen05 kernel lands
#include "common.cuh"
namespace ftcf {
bool gemm_tcgen05_supported(int, int, int, int) { return false; }
int gemm_w8a16_tcgen05(const void*, const uint8_t*, const void*, const void*, void*, int, int, int, int, cudaStream_t)
{
    set_error("tcgen05 gemm not built");
    return FTCF_ERR_UNSUPPORTED;
}
int gemm_f16_tcgen05(const void*, const void*, const void*, void*, int, int, int, int, int, int, cudaStream_t)
{
    set_error("tcgen05 gemm not built");
    return FTCF_ERR_UNSUPPORTED;
}
}  // namespace ftcf
