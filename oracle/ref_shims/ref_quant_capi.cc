// TEST INFRASTRUCTURE.  extern "C" doorway into the reference's own object code
// (cutlass_preprocessors.cc compiled where it lies under /root/reference) so the
// numpy oracle and our CUDA/C++ quantiser can be checked against the reference
// itself.  Nothing here restates an algorithm: it only forwards calls.
#include <cuda_fp16.h>
#include <cstring>
#include <exception>
#include <vector>

#include "src/fastertransformer/kernels/cutlass_kernels/cutlass_preprocessors.h"

namespace ft = fastertransformer;

extern "C" {

// fp16 weights [e?, k, n] -> processed int8 (sm80 layout), unprocessed int8, fp16 scales.
int ref_symmetric_quantize_half(int8_t* processed, int8_t* unprocessed, void* scales_fp16,
                                const void* weight_fp16, const size_t* shape, int ndim)
{
    try {
        std::vector<size_t> s(shape, shape + ndim);
        ft::symmetric_quantize<half, half>(processed, unprocessed, reinterpret_cast<half*>(scales_fp16),
                                           reinterpret_cast<const half*>(weight_fp16), s,
                                           ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    } catch (const std::exception&) { return 1; }
}

// fp32 weights, fp16 scales (symmetric_quantize<half, float>).
int ref_symmetric_quantize_float(int8_t* processed, int8_t* unprocessed, void* scales_fp16,
                                 const float* weight, const size_t* shape, int ndim)
{
    try {
        std::vector<size_t> s(shape, shape + ndim);
        ft::symmetric_quantize<half, float>(processed, unprocessed, reinterpret_cast<half*>(scales_fp16), weight, s,
                                            ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    } catch (const std::exception&) { return 1; }
}

int ref_preprocess_weights(int8_t* out, const int8_t* in, const size_t* shape, int ndim)
{
    try {
        std::vector<size_t> s(shape, shape + ndim);
        ft::preprocess_weights_for_mixed_gemm(out, in, s, ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    } catch (const std::exception&) { return 1; }
}

int ref_permute_b_rows(int8_t* out, const int8_t* in, const size_t* shape, int ndim, int arch)
{
    try {
        std::vector<size_t> s(shape, shape + ndim);
        ft::permute_B_rows_for_mixed_gemm(out, in, s, ft::QuantType::INT8_WEIGHT_ONLY, arch);
        return 0;
    } catch (const std::exception&) { return 1; }
}

int ref_subbyte_transpose(int8_t* out, const int8_t* in, const size_t* shape, int ndim)
{
    try {
        std::vector<size_t> s(shape, shape + ndim);
        ft::subbyte_transpose(out, in, s, ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    } catch (const std::exception&) { return 1; }
}

int ref_add_bias_and_interleave(int8_t* inout, size_t n)
{
    try {
        ft::add_bias_and_interleave_quantized_tensor_inplace(inout, n, ft::QuantType::INT8_WEIGHT_ONLY);
        return 0;
    } catch (const std::exception&) { return 1; }
}
}
