// pybind11 module `libth_common`: symmetric_quantize_last_axis_of_batched_matrix_int8 over the C ABI.
// Mirrors src/fastertransformer/th_op/common/WeightOnlyQuantOps.cc:140-233,344-349 (call sites
// examples/pytorch/codefuse/codefuse_example.py:196-197,392-399 and quant_and_save.py:46-47,93-99).
#include <torch/extension.h>

#include <vector>

#include "ftcf.h"

namespace {

int dtype_code(at::ScalarType st)
{
    switch (st) {
        case at::kFloat: return 0;
        case at::kHalf: return 1;
        case at::kBFloat16: return 2;
        default: return -1;
    }
}

// weight [k, n] or [e, k, n] on the CPU -> {int8 tensor of the same shape holding the processed bytes, scales [n] | [e, n]}
std::vector<at::Tensor> symmetric_quantize_int8(at::Tensor weight)
{
    TORCH_CHECK(weight.device().is_cpu(), "weight must be a CPU tensor");                       // CHECK_CPU :146
    TORCH_CHECK(weight.dim() == 2 || weight.dim() == 3, "Invalid dim. The dim of weight should be 2 or 3");   // :149
    const int dt = dtype_code(weight.scalar_type());
    TORCH_CHECK(dt >= 0, "Invalid datatype. Weight must be FP16, BF16 or FP32");               // :216-220
    at::Tensor w = weight.contiguous();
    const size_t e = w.dim() == 2 ? 1 : (size_t)w.size(0);
    const size_t k = (size_t)w.size(-2), n = (size_t)w.size(-1);
    at::Tensor processed = at::empty(w.sizes(), at::TensorOptions().dtype(at::kChar));
    at::Tensor scales = w.dim() == 2 ? at::empty({(int64_t)n}, w.options()) : at::empty({(int64_t)e, (int64_t)n}, w.options());
    const int rc = ftcf_symmetric_quantize_int8_host(w.data_ptr(), dt, e, k, n, static_cast<uint8_t*>(processed.data_ptr()), nullptr,
                                                     scales.data_ptr());
    TORCH_CHECK(rc == FTCF_OK, "libftcf error ", rc, ": ", ftcf_last_error());
    return {processed, scales};
}

}  // namespace

PYBIND11_MODULE(libth_common, module)
{
    module.def("symmetric_quantize_last_axis_of_batched_matrix_int8", &symmetric_quantize_int8);
}
