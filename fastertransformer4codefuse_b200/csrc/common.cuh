// Shared device/host helpers for the sm_100a kernels of libftcf.
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <cstdarg>
#include <cstdio>

#include "../../include/ftcf.h"

namespace ftcf {

// ---------------------------------------------------------------- error plumbing (host)
void set_error(const char* fmt, ...);
extern std::atomic<long long> g_launch_count;   // kernels launched by this library (bench's gpu_launches)
// Bumped whenever an engine buffer is re-allocated or a process-wide tunable changes: a captured decode graph bakes in
// pointers and launch shapes, so the graph cache key carries the generation it was captured under.
extern std::atomic<long long> g_capture_generation;

#define FTCF_CUDA_CHECK(expr)                                                                         \
    do {                                                                                              \
        cudaError_t _e = (expr);                                                                      \
        if (_e != cudaSuccess) {                                                                      \
            ::ftcf::set_error("%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FTCF_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

#define FTCF_LAUNCH_CHECK()                                                                           \
    do {                                                                                              \
        ::ftcf::g_launch_count.fetch_add(1, std::memory_order_relaxed);                               \
        cudaError_t _e = cudaPeekAtLastError();                                                       \
        if (_e != cudaSuccess) {                                                                      \
            ::ftcf::set_error("kernel launch failed: %s (%s:%d)", cudaGetErrorString(_e), __FILE__, __LINE__); \
            return FTCF_ERR_CUDA;                                                                     \
        }                                                                                             \
    } while (0)

#define FTCF_REQUIRE(cond, code, ...)                                                                 \
    do {                                                                                              \
        if (!(cond)) {                                                                                \
            ::ftcf::set_error(__VA_ARGS__);                                                           \
            return code;                                                                              \
        }                                                                                             \
    } while (0)

// ---------------------------------------------------------------- per-CTA timeline (debug; off unless ftcf_debug_trace_start)
// Every instrumented kernel appends one record per CTA: globaltimer stamps of its start, the end of its dependency wait, its
// first useful work and its end.  Device globals are per translation unit (no -rdc), so each .cu that emits records defines an
// installer with FTCF_TRACE_INSTALLER(name) and capi.cu calls them all.
struct TraceRec {
    unsigned long long t0, t1, t2, t3;
    int kind, cta, ncta, a, b, pad;
};
enum { TRC_GEMM_W8 = 1, TRC_GEMM_F16 = 2, TRC_MMHA = 10, TRC_LN = 20, TRC_RESIDUAL = 21, TRC_EMBED = 22, TRC_SAMPLING = 30 };
#ifdef __CUDACC__
static __device__ TraceRec* g_trc_buf = nullptr;
static __device__ unsigned* g_trc_cnt = nullptr;
static __device__ unsigned g_trc_cap = 0;
// `who`: only the one thread that will emit the record reads the timer, and only while a trace buffer is installed
// (%globaltimer reads are slow; 256 threads x 3 reads per CTA cost several per cent of a decode step)
__device__ __forceinline__ unsigned long long trc_now(bool who = true)
{
    if (!who || g_trc_buf == nullptr) return 0;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void trc_emit(int kind, unsigned long long t0, unsigned long long t1, unsigned long long t2, int a, int b, int extra = 0)
{
    if (g_trc_buf == nullptr) return;
    const unsigned i = atomicAdd(g_trc_cnt, 1u);
    if (i >= g_trc_cap) return;
    TraceRec r;
    r.t0 = t0; r.t1 = t1; r.t2 = t2; r.t3 = trc_now();
    r.kind = kind;
    r.cta = blockIdx.x + gridDim.x * (blockIdx.y + gridDim.y * blockIdx.z);
    r.ncta = gridDim.x * gridDim.y * gridDim.z;
    r.a = a; r.b = b; r.pad = extra;
    g_trc_buf[i] = r;
}
#define FTCF_TRACE_INSTALLER(name)                                                        \
    int name(TraceRec* buf, unsigned* cnt, unsigned cap)                                  \
    {                                                                                     \
        FTCF_CUDA_CHECK(cudaMemcpyToSymbol(g_trc_buf, &buf, sizeof(buf)));                \
        FTCF_CUDA_CHECK(cudaMemcpyToSymbol(g_trc_cnt, &cnt, sizeof(cnt)));                \
        FTCF_CUDA_CHECK(cudaMemcpyToSymbol(g_trc_cap, &cap, sizeof(cap)));                \
        return FTCF_OK;                                                                   \
    }
#endif

inline cudaStream_t as_stream(void* s) { return reinterpret_cast<cudaStream_t>(s); }

// Programmatic dependent launch (PDL): the decode step is a chain of short HBM-bound kernels; with this attribute the next
// kernel's CTAs become resident while the previous kernel drains, run their independent prologue (L2 prefetch of their
// weights / KV rows) and block in `griddepcontrol.wait` until the producer has completed.  Every kernel launched through
// launch_pdl() calls pdl_wait() before it reads or writes anything another kernel touches.
extern std::atomic<int> g_pdl_enabled;
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl_if(bool pdl, void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = block;
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    attr[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = attr;
    cfg.numAttrs = (pdl && g_pdl_enabled.load(std::memory_order_relaxed)) ? 1 : 0;
    return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
template <typename... KArgs, typename... Args>
inline cudaError_t launch_pdl(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t st, Args&&... args)
{
    return launch_pdl_if(true, kernel, grid, block, smem, st, static_cast<Args&&>(args)...);
}
__host__ __device__ constexpr int ceil_div(int a, int b) { return (a + b - 1) / b; }

// Fused prologue of the decode layer (m <= 4 tokens): the CTA builds its own copy of the GEMM input in shared memory,
//   r = ((add_ffn + add_attn) + add_bias) + x       (the PREVIOUS layer's parallel-residual add, add_residual_kernels.cu:116-176;
//                                                    skipped when add_ffn == NULL)
//   a = LayerNorm(r; gamma, beta)                   (layernorm_kernels.cu:158-286: fp32 statistics, half2 normalisation)
// instead of reading what a residual kernel and a LayerNorm kernel wrote: two launches (and their kernel boundaries, ~5 us of
// idle HBM each) leave the critical path of every layer.  Every CTA recomputes the same 10 KB row (L2 hits); CTA 0 stores r.
struct SkPro {
    const __half *x, *add_ffn, *add_attn, *add_bias, *gamma, *beta;
    __half* x_out;
    float eps;
    int cta_hint;
};


// ---------------------------------------------------------------- device helpers
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
// Ask L2 to fetch `bytes` (multiple of 16, 16-byte aligned address) -- no registers, no completion to wait for.
__device__ __forceinline__ void l2_prefetch_bulk(const void* p, uint32_t bytes)
{
    asm volatile("cp.async.bulk.prefetch.L2.global [%0], %1;" ::"l"(p), "r"(bytes) : "memory");
}

__device__ __forceinline__ float warp_sum(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
__device__ __forceinline__ float warp_max(float v)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v = fmaxf(v, __shfl_xor_sync(0xffffffffu, v, o));
    return v;
}

// Block-wide sum of one float; `red` is >= 32 floats of shared memory.  All threads get the result.
__device__ __forceinline__ float block_sum(float v, float* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_sum(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : 0.f;
    return warp_sum(r);
}
__device__ __forceinline__ float block_max(float v, float* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
    v = warp_max(v);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    float r = (lane < nw) ? red[lane] : -INFINITY;
    return warp_max(r);
}

// 128-bit streaming load that does not allocate in L1 (weights / KV are touched once per launch).
__device__ __forceinline__ uint4 ld_stream_16(const void* p)
{
    uint4 r;
    asm volatile("ld.global.nc.L1::no_allocate.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w)
                 : "l"(p));
    return r;
}
// 128-bit cached read-only load (activations re-read by many warps).
__device__ __forceinline__ uint4 ld_ro_16(const void* p) { return __ldg(reinterpret_cast<const uint4*>(p)); }

// four biased bytes (value q + 128) -> (b0-128, b1-128) and (b2-128, b3-128) as half2 bit patterns: PRMT builds 0x64xx
// (= 1024 + b, exact in fp16), one HSUB2 subtracts 1152 -- the same constant trick as the reference's
// interleaved_numeric_conversion.h:69-75, without its interleaved byte order
__device__ __forceinline__ void u8x4_to_h2x2(uint32_t w, uint32_t& lo, uint32_t& hi)
{
    lo = __byte_perm(w, 0x64646464u, 0x4140);
    hi = __byte_perm(w, 0x64646464u, 0x4342);
    const uint32_t magic = 0x64806480u;   // 1152.0 = 1024 + 128, twice
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(lo), "r"(magic));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(hi), "r"(magic));
}
// mma.sync m16n8k16, fp16 operands, fp32 accumulate (the legacy tensor path: streaming fp16 GEMM and prefill attention)
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0,
                                          uint32_t b1)
{
    asm volatile(
        "mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
        : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
        : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}

__device__ __forceinline__ float2 h2_to_f2(uint32_t u)
{
    return __half22float2(*reinterpret_cast<const __half2*>(&u));
}
__device__ __forceinline__ uint32_t f2_to_h2(float a, float b)
{
    __half2 h = __floats2half2_rn(a, b);
    return *reinterpret_cast<uint32_t*>(&h);
}

// ---------------------------------------------------------------- tensor-parallel exchange (ftcf_tp_exchange, include/ftcf.h)
struct TpIndex {
    int slot;
    unsigned epoch;
};
__device__ __forceinline__ TpIndex tp_index(const ftcf_tp_exchange& ex, int layer)
{
    const int g = (*ex.step - ex.step_base) * ex.layer_num + layer;
    return {g & 1, (unsigned)(g >> 1) + 1u};
}
// offset, in 8-byte flagged words, of row 0 of (slot, kind, source rank)
__device__ __forceinline__ size_t tp_word_offset(const ftcf_tp_exchange& ex, int slot, int kind, int src)
{
    return ((size_t)(slot * 2 + kind) * ex.tp + src) * ex.m_max * (size_t)(ex.h >> 1);
}
// push side: one flagged word {fp16 pair, epoch}, a single 8-byte store (data and flag become visible together)
__device__ __forceinline__ void tp_store_word(void* area, size_t word, uint32_t pair, unsigned epoch)
{
    asm volatile("st.volatile.global.v2.u32 [%0], {%1, %2};" ::"l"(static_cast<uint2*>(area) + word), "r"(pair), "r"(epoch) : "memory");
}
// gather side: the flagged words of 8 consecutive columns (4 words = two 16-byte loads) of both kinds from CH source ranks at a
// time -- all 4 * CH loads are issued before the first flag is looked at, so a gather costs one L2 round trip per CH ranks instead
// of one per rank and kind -- polled until every word carries `epoch`.  Bounded: a dead peer traps this kernel instead of hanging
// the GPU for good.
__device__ __forceinline__ void tp_ld_words(const void* p, uint4& a, uint4& b)
{
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(a.x), "=r"(a.y), "=r"(a.z), "=r"(a.w) : "l"(p) : "memory");
    asm volatile("ld.volatile.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(b.x), "=r"(b.y), "=r"(b.z), "=r"(b.w) : "l"(static_cast<const char*>(p) + 16) : "memory");
}
// 8 consecutive elements of row b of the all-reduced residual:  sum_r [((ffn_r + attn_r) + bias) + half(x / tp)]  in fp32, rounded
// once.  Same fp16 adds per rank as residual_kernel<0> (kernels/add_residual_kernels.cu:116-176); ranks are summed in rank order.
template <int CH>
__device__ __forceinline__ uint4 tp_gather_vec(const ftcf_tp_exchange& ex, const TpIndex ix, int b, int vi, uint4 xv, const __half* bias)
{
    const uint2* area = static_cast<const uint2*>(ex.peer_data[ex.rank]);
    const float inv_tp = 1.f / ex.tp;
    __half2 xs[4];
    const __half2* xh = reinterpret_cast<const __half2*>(&xv);
    uint4 bv = make_uint4(0, 0, 0, 0);
    if (bias != nullptr) bv = *reinterpret_cast<const uint4*>(bias + vi * 8);
    const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
    for (int j = 0; j < 4; ++j) {
        const float2 f = __half22float2(xh[j]);
        xs[j] = __floats2half2_rn(f.x * inv_tp, f.y * inv_tp);
    }
    float acc[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
    const size_t row_words = (size_t)b * (ex.h >> 1) + (size_t)vi * 4;
    const size_t kind_words = (size_t)ex.tp * ex.m_max * (size_t)(ex.h >> 1);          // kind 0 -> kind 1 of the same slot
    const size_t rank_words = (size_t)ex.m_max * (size_t)(ex.h >> 1);
    const uint2* base = area + tp_word_offset(ex, ix.slot, 0, 0) + row_words;
    for (int r0 = 0; r0 < ex.tp; r0 += CH) {
        uint4 w[CH][4];        // [rank][attn lo, attn hi, ffn lo, ffn hi]
        for (long long spin = 0;; ++spin) {
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (r0 + i < ex.tp) {
                    const uint2* p = base + (size_t)(r0 + i) * rank_words;
                    tp_ld_words(p, w[i][0], w[i][1]);
                    tp_ld_words(p + kind_words, w[i][2], w[i][3]);
                }
            }
            bool ok = true;
#pragma unroll
            for (int i = 0; i < CH; ++i) {
                if (r0 + i < ex.tp) {
#pragma unroll
                    for (int q = 0; q < 4; ++q) ok = ok && w[i][q].y == ix.epoch && w[i][q].w == ix.epoch;
                }
            }
            if (ok) break;
            if (spin > (1ll << 24)) __trap();
        }
#pragma unroll
        for (int i = 0; i < CH; ++i) {
            if (r0 + i < ex.tp) {
                const uint32_t av[4] = {w[i][0].x, w[i][0].z, w[i][1].x, w[i][1].z};
                const uint32_t fv[4] = {w[i][2].x, w[i][2].z, w[i][3].x, w[i][3].z};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    __half2 o = __hadd2(*reinterpret_cast<const __half2*>(&fv[j]), *reinterpret_cast<const __half2*>(&av[j]));
                    if (bias != nullptr) o = __hadd2(o, bh[j]);
                    o = __hadd2(o, xs[j]);
                    const float2 f = __half22float2(o);
                    acc[2 * j] += f.x;
                    acc[2 * j + 1] += f.y;
                }
            }
        }
    }
    uint4 out;
    __half2* oh = reinterpret_cast<__half2*>(&out);
#pragma unroll
    for (int j = 0; j < 4; ++j) oh[j] = __floats2half2_rn(acc[2 * j], acc[2 * j + 1]);
    return out;
}

// tanh-form GELU in fp32: x * 0.5 * (1 + tanh(0.79788456 * (x + 0.044715 x^3))).  tanh through exp so that the
// result is within a few ulp of the exact function (the reference uses a fast tanh in the int8 epilogue,
// cutlass_extensions/epilogue/thread/ft_fused_activations.h:61-84).
__device__ __forceinline__ float tanh_accurate(float x)
{
    const float ax = fabsf(x);
    const float e = __expf(-2.f * ax);
    const float t = __fdividef(1.f - e, 1.f + e);
    return copysignf(t, x);
}
__device__ __forceinline__ float gelu_tanh_f32(float x)
{
    return x * (0.5f * (1.f + tanh_accurate(0.7978845608028654f * (x + 0.044715f * x * x * x))));
}
// fp16 path of the reference (kernels/activation_kernels.cu:59-72): cube and product rounded to half.
__device__ __forceinline__ __half gelu_tanh_half_ref(__half v)
{
    const __half v2 = __hmul(v, v);
    const __half pow3 = __hmul(v, v2);
    const float vf = __half2float(v);
    const float cdf = 0.5f * (1.f + tanh_accurate(0.7978845608028654f * (vf + 0.044715f * __half2float(pow3))));
    return __hmul(v, __float2half_rn(cdf));
}
#endif  // __CUDACC__

}  // namespace ftcf
