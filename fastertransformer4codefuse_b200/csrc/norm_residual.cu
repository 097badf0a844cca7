// LayerNorm, residual adds and the embedding gather of the GPT-NeoX layer (HBM/latency-bound glue kernels).
//
// Reference behaviour restated (not ported):
//   * LayerNorm: fp32 sum and sum-of-squares, var = E[x^2] - mean^2 + eps, then the normalisation itself in fp16:
//     fma((x - mean_h) * rstd_h, gamma, beta): half2 sub, mul, then ONE fused multiply-add -- what nvcc makes of the reference's
//     hmul2(hsub2(x, mean), rstd, gamma) + beta (its SASS is HADD2, HMUL2, HFMA2; pinned on the GPU by tests/test_ref_kernels_gpu.py)
//     (kernels/layernorm_kernels.cu:158-286, dispatched by invokeGeneralLayerNorm :1653-1735).
//   * out = (half)(x / tp) + ffn + attn + bias with fp16 adds (kernels/add_residual_kernels.cu:116-176).
//   * embedding row gather (kernels/gpt_kernels.cu:32-105).
// One CTA per row, 128-bit loads, row kept in registers between the statistics pass and the normalisation pass.
#include "common.cuh"

namespace ftcf {

constexpr int LN_MAX_VEC = 4;   // up to 4 x 8 halves per thread: n <= 256 threads * 32 = 8192

// PRE: 0 plain LN(x); 1: r = x + add1 (+ bias) stored to residual_out, then LN(r)
template <int PRE>
__global__ void __launch_bounds__(256)
layernorm_kernel(const __half* __restrict__ x, const __half* __restrict__ add1, const __half* __restrict__ add_bias,
                 __half* __restrict__ residual_out, const __half* __restrict__ gamma, const __half* __restrict__ beta,
                 __half* __restrict__ y, int n, float eps)
{
    __shared__ float red[32];
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    const size_t row = blockIdx.x;
    const int nvec = n >> 3;   // 8 halves per vector
    uint4 v[LN_MAX_VEC];
    float s = 0.f, ss = 0.f;
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        const int vi = threadIdx.x + i * blockDim.x;
        if (vi < nvec) {
            v[i] = *reinterpret_cast<const uint4*>(x + row * n + vi * 8);
            if (PRE == 1) {
                // reference (layernorm_kernels.cu:196-232): fp32 sum bias + residual + input, rounded once; the
                // statistics use the fp32 values, the normalisation the rounded ones.
                const uint4 a = *reinterpret_cast<const uint4*>(add1 + row * n + vi * 8);
                uint4 b = make_uint4(0, 0, 0, 0);
                if (add_bias != nullptr) b = *reinterpret_cast<const uint4*>(add_bias + vi * 8);
                __half2* vh = reinterpret_cast<__half2*>(&v[i]);
                const __half2* ah = reinterpret_cast<const __half2*>(&a);
                const __half2* bh = reinterpret_cast<const __half2*>(&b);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 fx = __half22float2(vh[j]), fa = __half22float2(ah[j]), fb = __half22float2(bh[j]);
                    const float v1 = (fb.x + fx.x) + fa.x, v2 = (fb.y + fx.y) + fa.y;
                    vh[j] = __floats2half2_rn(v1, v2);
                    s += v1 + v2;
                    ss += v1 * v1 + v2 * v2;
                }
                *reinterpret_cast<uint4*>(residual_out + row * n + vi * 8) = v[i];
                continue;
            }
            const __half2* vh = reinterpret_cast<const __half2*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = __half22float2(vh[j]);
                s += f.x + f.y;
                ss += f.x * f.x + f.y * f.y;
            }
        }
    }
    s = block_sum(s, red);
    ss = block_sum(ss, red);
    const float mean = s / n;
    const float var = ss / n - mean * mean + eps;
    const float rstd = rsqrtf(var);
    const __half2 mean_h = __float2half2_rn(mean), rstd_h = __float2half2_rn(rstd);
#pragma unroll
    for (int i = 0; i < LN_MAX_VEC; ++i) {
        const int vi = threadIdx.x + i * blockDim.x;
        if (vi < nvec) {
            const uint4 gq = *reinterpret_cast<const uint4*>(gamma + vi * 8);
            const uint4 bq = *reinterpret_cast<const uint4*>(beta + vi * 8);
            const __half2* gh = reinterpret_cast<const __half2*>(&gq);
            const __half2* bh = reinterpret_cast<const __half2*>(&bq);
            __half2* vh = reinterpret_cast<__half2*>(&v[i]);
#pragma unroll
            for (int j = 0; j < 4; ++j)
                vh[j] = __hfma2(__hmul2_rn(__hsub2_rn(vh[j], mean_h), rstd_h), gh[j], bh[j]);   // as the reference's object code (SASS: HADD2, HMUL2, HFMA2): the last multiply and the add are ONE fused operation
            *reinterpret_cast<uint4*>(y + row * n + vi * 8) = v[i];
        }
    }
    if (threadIdx.x == 0) trc_emit(TRC_LN, trc_t0, trc_t0, trc_t0, n, PRE);
}

// MODE 0: out = ((ffn + attn) + bias) + (tp > 1 ? half(x / tp) : x)      (parallel residual)
// MODE 1: out = (y + x) + bias                                           (sequential residual tail)
template <int MODE>
__global__ void __launch_bounds__(256)
residual_kernel(__half* __restrict__ out, const __half* __restrict__ a, const __half* __restrict__ b,
                const __half* __restrict__ x, const __half* __restrict__ bias, int n, size_t total_vec, float inv_tp)
{
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    const size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= total_vec) return;
    const int col = (int)((vi * 8) % n);
    uint4 av = *reinterpret_cast<const uint4*>(a + vi * 8);
    const uint4 xv = *reinterpret_cast<const uint4*>(x + vi * 8);
    __half2* ah = reinterpret_cast<__half2*>(&av);
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
    uint4 biasv = make_uint4(0, 0, 0, 0);
    if (bias != nullptr) biasv = *reinterpret_cast<const uint4*>(bias + col);
    const __half2* bh = reinterpret_cast<const __half2*>(&biasv);
    if (MODE == 0) {
        const uint4 bv = *reinterpret_cast<const uint4*>(b + vi * 8);
        const __half2* b2 = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __half2 r = __hadd2(ah[j], b2[j]);
            if (bias != nullptr) r = __hadd2(r, bh[j]);
            __half2 xs = xh[j];
            if (inv_tp != 1.f) {
                const float2 f = __half22float2(xs);
                xs = __floats2half2_rn(f.x * inv_tp, f.y * inv_tp);
            }
            ah[j] = __hadd2(r, xs);
        }
    } else {
#pragma unroll
        for (int j = 0; j < 4; ++j) {
            __half2 r = __hadd2(ah[j], xh[j]);   // (in + residual) + bias, add_residual_kernels.cu:44-46
            if (bias != nullptr) r = __hadd2(r, bh[j]);
            ah[j] = r;
        }
    }
    *reinterpret_cast<uint4*>(out + vi * 8) = av;
    if (threadIdx.x == 0) trc_emit(TRC_RESIDUAL, trc_t0, trc_t0, trc_t0, n, MODE);
}

__global__ void __launch_bounds__(256)
embedding_kernel(__half* __restrict__ out, const __half* __restrict__ table, const int32_t* __restrict__ ids, int n, int vocab,
                 size_t total_vec)
{
    const size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= total_vec) return;
    const int nvec = n >> 3;
    const int row = (int)(vi / nvec), c = (int)(vi % nvec);
    int id = ids[row];
    id = min(max(id, 0), vocab - 1);
    *reinterpret_cast<uint4*>(out + (size_t)row * n + c * 8) = ld_ro_16(table + (size_t)id * n + c * 8);
}

// Stand-alone gather side of the tensor-parallel exchange (ftcf_tp_exchange): grid (column slices, rows).
__global__ void __launch_bounds__(256)
tp_gather_residual_kernel(const ftcf_tp_exchange ex, int layer, const __half* __restrict__ x, const __half* __restrict__ bias,
                          __half* __restrict__ x_out)
{
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    pdl_launch_dependents();
    pdl_wait();
    const unsigned long long trc_t1 = trc_now(threadIdx.x == 0);
    const TpIndex ix = tp_index(ex, layer);
    const int b = blockIdx.y, nvec = ex.h >> 3;
    const int vi = blockIdx.x * blockDim.x + threadIdx.x;
    if (vi < nvec) {
        const uint4 xv = *reinterpret_cast<const uint4*>(x + (size_t)b * ex.h + vi * 8);
        *reinterpret_cast<uint4*>(x_out + (size_t)b * ex.h + vi * 8) = tp_gather_vec<4>(ex, ix, b, vi, xv, bias);
    }
    if (threadIdx.x == 0) trc_emit(TRC_RESIDUAL, trc_t0, trc_t1, trc_t1, ex.h, 9);
}

FTCF_TRACE_INSTALLER(trace_install_norm_residual)

static int ln_threads(int n)
{
    const int nvec = n / 8;
    int th = ((ceil_div(nvec, LN_MAX_VEC) + 31) / 32) * 32;
    if (th < 64) th = 64;
    return th;
}

}  // namespace ftcf

using namespace ftcf;

extern "C" int ftcf_layernorm(const void* x, const void* gamma, const void* beta, void* y, int m, int n, float eps,
                              void* stream)
{
    FTCF_REQUIRE(m > 0 && n > 0 && n % 8 == 0 && n <= 256 * 8 * LN_MAX_VEC, FTCF_ERR_UNSUPPORTED,
                 "layernorm: m=%d n=%d (n must be a multiple of 8, <= %d)", m, n, 256 * 8 * LN_MAX_VEC);
    layernorm_kernel<0><<<m, ln_threads(n), 0, as_stream(stream)>>>(
        static_cast<const __half*>(x), nullptr, nullptr, nullptr, static_cast<const __half*>(gamma),
        static_cast<const __half*>(beta), static_cast<__half*>(y), n, eps);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_add_bias_residual_layernorm(const void* x, const void* add1, const void* add_bias, void* residual_out,
                                                const void* gamma, const void* beta, void* y, int m, int n, float eps,
                                                void* stream)
{
    FTCF_REQUIRE(m > 0 && n > 0 && n % 8 == 0 && n <= 256 * 8 * LN_MAX_VEC, FTCF_ERR_UNSUPPORTED,
                 "add_bias_residual_layernorm: m=%d n=%d unsupported", m, n);
    FTCF_REQUIRE(add1 != nullptr && residual_out != nullptr, FTCF_ERR_INVALID, "add_bias_residual_layernorm: null operand");
    layernorm_kernel<1><<<m, ln_threads(n), 0, as_stream(stream)>>>(
        static_cast<const __half*>(x), static_cast<const __half*>(add1), static_cast<const __half*>(add_bias),
        static_cast<__half*>(residual_out), static_cast<const __half*>(gamma), static_cast<const __half*>(beta),
        static_cast<__half*>(y), n, eps);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_add_bias_attn_ffn_residual(void* out, const void* ffn, const void* attn, const void* x, const void* bias,
                                               int m, int n, int tp, void* stream)
{
    FTCF_REQUIRE(m > 0 && n > 0 && n % 8 == 0 && tp >= 1, FTCF_ERR_UNSUPPORTED, "residual: m=%d n=%d tp=%d", m, n, tp);
    const size_t total = (size_t)m * n / 8;
    residual_kernel<0><<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<__half*>(out), static_cast<const __half*>(ffn), static_cast<const __half*>(attn),
        static_cast<const __half*>(x), static_cast<const __half*>(bias), n, total, 1.f / (float)tp);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_add_bias_residual(void* out, const void* y, const void* x, const void* bias, int m, int n, void* stream)
{
    FTCF_REQUIRE(m > 0 && n > 0 && n % 8 == 0, FTCF_ERR_UNSUPPORTED, "residual: m=%d n=%d", m, n);
    const size_t total = (size_t)m * n / 8;
    residual_kernel<1><<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<__half*>(out), static_cast<const __half*>(y), nullptr, static_cast<const __half*>(x),
        static_cast<const __half*>(bias), n, total, 1.f);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_embedding_lookup(void* out, const void* table, const int32_t* ids, int m, int n, int vocab, void* stream)
{
    FTCF_REQUIRE(m > 0 && n > 0 && n % 8 == 0, FTCF_ERR_UNSUPPORTED, "embedding: m=%d n=%d", m, n);
    const size_t total = (size_t)m * n / 8;
    embedding_kernel<<<(unsigned)((total + 255) / 256), 256, 0, as_stream(stream)>>>(
        static_cast<__half*>(out), static_cast<const __half*>(table), ids, n, vocab, total);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_tp_gather_residual(const ftcf_tp_exchange* ex, int layer, const void* x, const void* bias, void* x_out, int m, void* stream)
{
    FTCF_REQUIRE(ex && x && x_out && ex->tp > 1 && ex->tp <= 8 && m >= 1 && m <= ex->m_max && ex->h % 8 == 0 && ex->step, FTCF_ERR_INVALID,
                 "tp_gather_residual: bad argument");
    const cudaError_t err = launch_pdl(tp_gather_residual_kernel, dim3(ceil_div(ex->h >> 3, 64), m), dim3(64), 0, as_stream(stream), *ex, layer,
                                       static_cast<const __half*>(x), static_cast<const __half*>(bias), static_cast<__half*>(x_out));
    FTCF_REQUIRE(err == cudaSuccess, FTCF_ERR_CUDA, "tp_gather_residual launch failed: %s", cudaGetErrorString(err));
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
