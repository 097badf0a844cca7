// C-ABI glue of libftcf: error reporting, device check, GEMM dispatch, and the CPU-side weight-only INT8 quantiser.
#include <cuda_bf16.h>

#include <algorithm>
#include <cmath>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include "common.cuh"

namespace ftcf {

static thread_local char g_err[1024] = "";
std::atomic<long long> g_launch_count{0};
std::atomic<int> g_pdl_enabled{1};
std::atomic<long long> g_capture_generation{0};

void set_error(const char* fmt, ...)
{
    va_list ap;
    va_start(ap, fmt);
    vsnprintf(g_err, sizeof(g_err), fmt, ap);
    va_end(ap);
}

// implemented in gemm_skinny.cu / gemm_tcgen05.cu
int gemm_w8a16_skinny(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                      int act, const ftcf_launch_hint* hint, cudaStream_t st);
int gemm_f16_skinny(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                    int out_f32, const ftcf_launch_hint* hint, cudaStream_t st);
int gemm_w8a16_skinny_ln(const ftcf_ln_prologue& pro, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                         int act, cudaStream_t st);
int gemm_f16_skinny_ln(const ftcf_ln_prologue& pro, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                       int out_f32, cudaStream_t st);
int gemm_w8a16_tcgen05(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                       int act, cudaStream_t st);
int gemm_f16_tcgen05(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                     int out_f32, cudaStream_t st);
bool gemm_tcgen05_supported(int m, int n, int k, int elem_bytes);
int gemm_w8a16_decode(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k, int act,
                      const SkPro* pro, cudaStream_t st, const ftcf_tp_exchange* push = nullptr, int push_kind = 0, int push_layer = 0,
                      const ftcf_launch_hint* hint = nullptr);
bool gemm_decode_supported(int m, int n, int k);
extern std::atomic<int> g_dg_target_ctas, g_dg_min_kb, g_dg_evict_first, g_dg_max_stages, g_dg_cluster, g_dg_lean;
std::atomic<int> g_decode_impl{3};   // tunable "decode_impl": 3 = tcgen05 decode GEMM for int8 at m <= 32 (default), 1 = round-1 streaming mma.sync kernel
extern std::atomic<int> g_prefill_mma, g_mmha_onepass, g_mmha_splits, g_mmha_bulk, g_mmha_lite;
extern std::atomic<int> g_mmha_pdl, g_mmha_prefetch, g_sk_carveout, g_sk_evict_first, g_tc_ksplit, g_sk_target_ctas;

}  // namespace ftcf

using namespace ftcf;

extern "C" const char* ftcf_last_error(void) { return g_err; }
extern "C" int ftcf_abi_version(void) { return 3; }
extern "C" long long ftcf_launch_count(void) { return g_launch_count.load(); }

extern "C" int ftcf_set_tunable(const char* name, int value)
{
    FTCF_REQUIRE(name != nullptr, FTCF_ERR_INVALID, "set_tunable: null name");
    const std::string n(name);
    g_capture_generation.fetch_add(1, std::memory_order_relaxed);   // captured launches may depend on any of these
    if (n == "pdl") g_pdl_enabled.store(value);
    else if (n == "skinny_target_ctas") { FTCF_REQUIRE(value >= 0, FTCF_ERR_INVALID, "skinny_target_ctas %d", value); g_sk_target_ctas.store(value); }
    else if (n == "skinny_evict_first") g_sk_evict_first.store(value);
    else if (n == "tc_ksplit") g_tc_ksplit.store(value);
    else if (n == "mmha_pdl") g_mmha_pdl.store(value);
    else if (n == "prefill_mma") g_prefill_mma.store(value);
    else if (n == "mmha_onepass") g_mmha_onepass.store(value);
    else if (n == "mmha_splits") g_mmha_splits.store(value);
    else if (n == "mmha_bulk") g_mmha_bulk.store(value);
    else if (n == "mmha_lite") g_mmha_lite.store(value);
    else if (n == "mmha_prefetch") g_mmha_prefetch.store(value);
    else if (n == "skinny_carveout") g_sk_carveout.store(value);
    else if (n == "decode_target_ctas") g_dg_target_ctas.store(value);
    else if (n == "decode_min_kb") g_dg_min_kb.store(value);
    else if (n == "decode_evict_first") g_dg_evict_first.store(value);
    else if (n == "decode_impl") g_decode_impl.store(value);
    else if (n == "decode_max_stages") g_dg_max_stages.store(value);
    else if (n == "decode_cluster") g_dg_cluster.store(value);
    else if (n == "decode_lean") g_dg_lean.store(value);
    else FTCF_REQUIRE(false, FTCF_ERR_INVALID, "set_tunable: unknown tunable %s", name);
    return FTCF_OK;
}

// ---------------------------------------------------------------- per-CTA timeline (debug)
namespace ftcf {
int trace_install_gemm_skinny(TraceRec*, unsigned*, unsigned);
int trace_install_gemm_decode(TraceRec*, unsigned*, unsigned);
int trace_install_attention(TraceRec*, unsigned*, unsigned);
int trace_install_norm_residual(TraceRec*, unsigned*, unsigned);
int trace_install_sampling(TraceRec*, unsigned*, unsigned);
}  // namespace ftcf
static TraceRec* g_trace_dev = nullptr;
static unsigned* g_trace_cnt_dev = nullptr;
static unsigned g_trace_cap = 0;
static int trace_install_all(TraceRec* buf, unsigned* cnt, unsigned cap)
{
    int rc = trace_install_gemm_skinny(buf, cnt, cap);
    if (rc == FTCF_OK) rc = trace_install_gemm_decode(buf, cnt, cap);
    if (rc == FTCF_OK) rc = trace_install_attention(buf, cnt, cap);
    if (rc == FTCF_OK) rc = trace_install_norm_residual(buf, cnt, cap);
    if (rc == FTCF_OK) rc = trace_install_sampling(buf, cnt, cap);
    return rc;
}

namespace ftcf { int decode_probe_install(long long* dev_buf); }
// Debug: install (or remove, NULL) a device buffer of 64 x 8 int64 that CTA (0,0,0) of every decode-GEMM launch fills with
// per-K-step clock stamps (see gemm_decode.cu).
extern "C" int ftcf_debug_decode_probe(void* dev_buf) { return decode_probe_install(static_cast<long long*>(dev_buf)); }

extern "C" int ftcf_debug_trace_start(unsigned capacity)
{
    FTCF_CUDA_CHECK(cudaDeviceSynchronize());
    if (g_trace_dev == nullptr || g_trace_cap < capacity) {
        if (g_trace_dev) cudaFree(g_trace_dev);
        if (g_trace_cnt_dev) cudaFree(g_trace_cnt_dev);
        g_trace_dev = nullptr;
        FTCF_CUDA_CHECK(cudaMalloc(&g_trace_dev, (size_t)capacity * sizeof(TraceRec)));
        FTCF_CUDA_CHECK(cudaMalloc(&g_trace_cnt_dev, sizeof(unsigned)));
        g_trace_cap = capacity;
    }
    FTCF_CUDA_CHECK(cudaMemset(g_trace_cnt_dev, 0, sizeof(unsigned)));
    return trace_install_all(g_trace_dev, g_trace_cnt_dev, g_trace_cap);
}

// stops tracing and copies up to max_records records (sizeof == 56) to host memory; returns the count through *n
extern "C" int ftcf_debug_trace_stop(void* out_host, unsigned max_records, unsigned* n)
{
    FTCF_REQUIRE(n != nullptr, FTCF_ERR_INVALID, "trace_stop: null");
    *n = 0;
    if (g_trace_dev == nullptr) return FTCF_OK;
    FTCF_CUDA_CHECK(cudaDeviceSynchronize());
    unsigned cnt = 0;
    FTCF_CUDA_CHECK(cudaMemcpy(&cnt, g_trace_cnt_dev, sizeof(cnt), cudaMemcpyDeviceToHost));
    cnt = std::min(cnt, std::min(g_trace_cap, max_records));
    if (out_host != nullptr && cnt > 0) FTCF_CUDA_CHECK(cudaMemcpy(out_host, g_trace_dev, (size_t)cnt * sizeof(TraceRec), cudaMemcpyDeviceToHost));
    *n = cnt;
    return trace_install_all(nullptr, nullptr, 0);
}

extern "C" int ftcf_device_check(void)
{
    int dev = 0;
    FTCF_CUDA_CHECK(cudaGetDevice(&dev));
    int major = 0, minor = 0;
    FTCF_CUDA_CHECK(cudaDeviceGetAttribute(&major, cudaDevAttrComputeCapabilityMajor, dev));
    FTCF_CUDA_CHECK(cudaDeviceGetAttribute(&minor, cudaDevAttrComputeCapabilityMinor, dev));
    FTCF_REQUIRE(major == 10, FTCF_ERR_UNSUPPORTED,
                 "libftcf is built for sm_100a only; device %d is compute capability %d.%d (no fallback path exists)", dev, major,
                 minor);
    return FTCF_OK;
}

// Auto mode (impl 0) for the INT8 GEMM: m <= 32 rows -> the tcgen05 decode kernel (gemm_decode.cu), more rows -> the tcgen05
// prefill kernel (gemm_tcgen05.cu).  impl 1 forces round 1's streaming mma.sync kernel (kept for A/B measurements and for
// k that is not a multiple of 128), impl 2 the prefill kernel, impl 3 the decode kernel.
static constexpr int kSkinnyMaxM = 11;   // fp16 weights only: above it the tcgen05 kernel takes over

extern "C" int ftcf_gemm_w8a16_ex(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n,
                                  int k, int act, int impl, const ftcf_launch_hint* hint, void* stream)
{
    FTCF_REQUIRE(x && w_nk && scale && y, FTCF_ERR_INVALID, "gemm_w8a16: null operand");
    FTCF_REQUIRE(act == 0 || act == 1, FTCF_ERR_INVALID, "gemm_w8a16: act %d", act);
    cudaStream_t st = as_stream(stream);
    if (impl == 1) return gemm_w8a16_skinny(x, w_nk, scale, bias, y, m, n, k, act, hint, st);
    if (impl == 2) return gemm_w8a16_tcgen05(x, w_nk, scale, bias, y, m, n, k, act, st);
    if (impl == 3) return gemm_w8a16_decode(x, w_nk, scale, bias, y, m, n, k, act, nullptr, st, nullptr, 0, 0, hint);
    if (g_decode_impl.load(std::memory_order_relaxed) == 3 && gemm_decode_supported(m, n, k))
        return gemm_w8a16_decode(x, w_nk, scale, bias, y, m, n, k, act, nullptr, st, nullptr, 0, 0, hint);
    if (m > kSkinnyMaxM && gemm_tcgen05_supported(m, n, k, 1)) return gemm_w8a16_tcgen05(x, w_nk, scale, bias, y, m, n, k, act, st);
    return gemm_w8a16_skinny(x, w_nk, scale, bias, y, m, n, k, act, hint, st);
}

extern "C" int ftcf_gemm_w8a16(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n,
                               int k, int act, int impl, void* stream)
{
    return ftcf_gemm_w8a16_ex(x, w_nk, scale, bias, y, m, n, k, act, impl, nullptr, stream);
}

extern "C" int ftcf_gemm_w8a16_ln(const ftcf_ln_prologue* pro, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m,
                                  int n, int k, int act, void* stream)
{
    FTCF_REQUIRE(pro && w_nk && scale && y, FTCF_ERR_INVALID, "gemm_w8a16_ln: null operand");
    FTCF_REQUIRE(act == 0 || act == 1, FTCF_ERR_INVALID, "gemm_w8a16_ln: act %d", act);
    FTCF_REQUIRE(pro->x_out == nullptr || pro->x_out != pro->x, FTCF_ERR_INVALID, "gemm_w8a16_ln: x_out must not alias x");
    if (g_decode_impl.load(std::memory_order_relaxed) == 3 && m <= 4 && gemm_decode_supported(m, n, k)) {
        SkPro sp{};
        sp.x = static_cast<const __half*>(pro->x); sp.add_ffn = static_cast<const __half*>(pro->add_ffn);
        sp.add_attn = static_cast<const __half*>(pro->add_attn); sp.add_bias = static_cast<const __half*>(pro->add_bias);
        sp.gamma = static_cast<const __half*>(pro->gamma); sp.beta = static_cast<const __half*>(pro->beta);
        sp.x_out = static_cast<__half*>(pro->x_out); sp.eps = pro->eps; sp.cta_hint = pro->cta_hint;
        return gemm_w8a16_decode(nullptr, w_nk, scale, bias, y, m, n, k, act, &sp, as_stream(stream));
    }
    return gemm_w8a16_skinny_ln(*pro, w_nk, scale, bias, y, m, n, k, act, as_stream(stream));
}

extern "C" int ftcf_gemm_w8a16_tp_push(const void* x, const uint8_t* w_nk, const void* scale, const ftcf_tp_exchange* ex, int kind, int layer,
                                       int m, int n, int k, const ftcf_launch_hint* hint, void* stream)
{
    FTCF_REQUIRE(x && w_nk && scale && ex, FTCF_ERR_INVALID, "gemm_w8a16_tp_push: null operand");
    return gemm_w8a16_decode(x, w_nk, scale, nullptr, nullptr, m, n, k, 0, nullptr, as_stream(stream), ex, kind, layer, hint);
}

extern "C" int ftcf_gemm_f16_ln(const ftcf_ln_prologue* pro, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy,
                                int act, int out_f32, void* stream)
{
    FTCF_REQUIRE(pro && w_nk && y, FTCF_ERR_INVALID, "gemm_f16_ln: null operand");
    FTCF_REQUIRE(act == 0 || act == 1, FTCF_ERR_INVALID, "gemm_f16_ln: act %d", act);
    FTCF_REQUIRE(ldy >= n, FTCF_ERR_INVALID, "gemm_f16_ln: ldy %d < n %d", ldy, n);
    FTCF_REQUIRE(pro->x_out == nullptr || pro->x_out != pro->x, FTCF_ERR_INVALID, "gemm_f16_ln: x_out must not alias x");
    return gemm_f16_skinny_ln(*pro, w_nk, bias, y, m, n, k, ldy, act, out_f32, as_stream(stream));
}

extern "C" int ftcf_gemm_f16_ex(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                                int out_f32, int impl, const ftcf_launch_hint* hint, void* stream)
{
    FTCF_REQUIRE(x && w_nk && y, FTCF_ERR_INVALID, "gemm_f16: null operand");
    FTCF_REQUIRE(act == 0 || act == 1, FTCF_ERR_INVALID, "gemm_f16: act %d", act);
    FTCF_REQUIRE(ldy >= n, FTCF_ERR_INVALID, "gemm_f16: ldy %d < n %d", ldy, n);
    cudaStream_t st = as_stream(stream);
    if (impl == 1) return gemm_f16_skinny(x, w_nk, bias, y, m, n, k, ldy, act, out_f32, hint, st);
    if (impl == 2) return gemm_f16_tcgen05(x, w_nk, bias, y, m, n, k, ldy, act, out_f32, st);
    if (m > kSkinnyMaxM && gemm_tcgen05_supported(m, n, k, 2))
        return gemm_f16_tcgen05(x, w_nk, bias, y, m, n, k, ldy, act, out_f32, st);
    return gemm_f16_skinny(x, w_nk, bias, y, m, n, k, ldy, act, out_f32, hint, st);
}

extern "C" int ftcf_gemm_f16(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                             int out_f32, int impl, void* stream)
{
    return ftcf_gemm_f16_ex(x, w_nk, bias, y, m, n, k, ldy, act, out_f32, impl, nullptr, stream);
}

// ------------------------------------------------------------------------------------------------ quantiser (CPU)
namespace {

inline float load_as_float(const void* base, int dtype, size_t i)
{
    switch (dtype) {
        case 0: return static_cast<const float*>(base)[i];
        case 1: return __half2float(static_cast<const __half*>(base)[i]);
        default: return __bfloat162float(static_cast<const __nv_bfloat16*>(base)[i]);
    }
}
inline void store_scale(void* base, int dtype, size_t i, float v)
{
    switch (dtype) {
        case 0: static_cast<float*>(base)[i] = v; break;
        case 1: static_cast<__half*>(base)[i] = __float2half_rn(v); break;
        default: static_cast<__nv_bfloat16*>(base)[i] = __float2bfloat16_rn(v); break;
    }
}

template <typename F>
void parallel_for(size_t n, F fn)
{
    unsigned nt = std::thread::hardware_concurrency();
    if (nt == 0) nt = 1;
    nt = (unsigned)std::min<size_t>(nt, std::max<size_t>(1, n));
    if (nt == 1) {
        fn(0, n);
        return;
    }
    std::vector<std::thread> th;
    const size_t chunk = (n + nt - 1) / nt;
    for (unsigned t = 0; t < nt; ++t) {
        const size_t lo = t * chunk, hi = std::min(n, lo + chunk);
        if (lo >= hi) break;
        th.emplace_back([=] { fn(lo, hi); });
    }
    for (auto& t : th) t.join();
}

}  // namespace

extern "C" int ftcf_symmetric_quantize_int8_host(const void* weight_host, int dtype, size_t e, size_t k, size_t n,
                                                 uint8_t* processed_host, int8_t* unprocessed_host, void* scales_host)
{
    FTCF_REQUIRE(weight_host && processed_host && scales_host, FTCF_ERR_INVALID, "quantize: null pointer");
    FTCF_REQUIRE(dtype >= 0 && dtype <= 2, FTCF_ERR_INVALID, "quantize: dtype %d (0 fp32, 1 fp16, 2 bf16)", dtype);
    FTCF_REQUIRE(e > 0 && k > 0 && n > 0, FTCF_ERR_INVALID, "quantize: empty matrix");
    for (size_t ei = 0; ei < e; ++ei) {
        const size_t moff = ei * k * n;
        std::vector<float> scale(n);
        // per-column absmax / 128 in fp32 (cutlass_preprocessors.cc:603-625)
        parallel_for(n, [&](size_t lo, size_t hi) {
            std::vector<float> mx(hi - lo, 0.f);
            for (size_t r = 0; r < k; ++r)
                for (size_t c = lo; c < hi; ++c) mx[c - lo] = std::max(mx[c - lo], std::fabs(load_as_float(weight_host, dtype, moff + r * n + c)));
            for (size_t c = lo; c < hi; ++c) {
                scale[c] = mx[c - lo] * (1.0f / 128.0f);
                store_scale(scales_host, dtype, ei * n + c, scale[c]);
            }
        });
        // q = clip(round_half_away(w / scale)), written K-major with +128 bias (and optionally plain [k, n])
        constexpr size_t TILE = 64;
        const size_t col_tiles = (n + TILE - 1) / TILE;
        parallel_for(col_tiles, [&](size_t lo, size_t hi) {
            for (size_t ct = lo; ct < hi; ++ct) {
                const size_t c0 = ct * TILE, c1 = std::min(n, c0 + TILE);
                for (size_t r0 = 0; r0 < k; r0 += TILE) {
                    const size_t r1 = std::min(k, r0 + TILE);
                    for (size_t r = r0; r < r1; ++r)
                        for (size_t c = c0; c < c1; ++c) {
                            const float w = load_as_float(weight_host, dtype, moff + r * n + c);
                            const float s = w / scale[c];
                            const float rounded = std::round(s);                       // halves away from zero
                            const float cl = std::max(-128.f, std::min(127.f, rounded));  // NaN -> 127 as in the reference
                            const int8_t q = (int8_t)cl;
                            processed_host[moff + c * k + r] = (uint8_t)((int)q + 128);
                            if (unprocessed_host) unprocessed_host[moff + r * n + c] = q;
                        }
                }
            }
        });
    }
    return FTCF_OK;
}

extern "C" int ftcf_int8_plain_to_b200_host(const int8_t* q_kn, size_t k, size_t n, uint8_t* out_nk)
{
    FTCF_REQUIRE(q_kn && out_nk && k > 0 && n > 0, FTCF_ERR_INVALID, "plain_to_b200: bad argument");
    parallel_for(n, [&](size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c)
            for (size_t r = 0; r < k; ++r) out_nk[c * k + r] = (uint8_t)((int)q_kn[r * n + c] + 128);
    });
    return FTCF_OK;
}

extern "C" int ftcf_int8_ampere_to_b200_host(const int8_t* processed_ampere, size_t k, size_t n, uint8_t* out_nk)
{
    // Inverse of the reference's sm80 pre-processing, element by element.  Forward chain (cutlass_preprocessors.cc):
    //   rows permuted inside groups of 16 (:133-201, map {0,1,8,9,2,3,10,11,4,5,12,13,6,7,14,15}), transpose to
    //   column-major (:207-348), ColumnMajorTileInterleave<64,2> on 32-bit words (:437-498), then +128 and a swap of
    //   bytes 1 and 2 of every word (:350-370).  The stored byte is already q + 128, which is what our layout holds.
    FTCF_REQUIRE(processed_ampere && out_nk, FTCF_ERR_INVALID, "ampere_to_b200: null pointer");
    FTCF_REQUIRE(k % 64 == 0 && n % 2 == 0, FTCF_ERR_UNSUPPORTED, "ampere_to_b200: k=%zu must be a multiple of 64, n=%zu even", k, n);
    static const int perm[16] = {0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15};
    int inv_perm[16];
    for (int i = 0; i < 16; ++i) inv_perm[perm[i]] = i;
    static const int swap12[4] = {0, 2, 1, 3};
    const uint8_t* src = reinterpret_cast<const uint8_t*>(processed_ampere);
    const size_t vec_rows = k / 4;
    parallel_for(n, [&](size_t lo, size_t hi) {
        for (size_t c = lo; c < hi; ++c)
            for (size_t r = 0; r < k; ++r) {
                const size_t rp = (r / 16) * 16 + inv_perm[r % 16];   // row of the permuted matrix that holds plain row r
                const size_t vr = rp / 4, byte = rp % 4;
                const size_t base = (vr / 16) * 16;
                const size_t wrow = 2 * base + 16 * (c % 2) + vr % 16;
                const size_t wcol = c / 2;
                const size_t word = wcol * (vec_rows * 2) + wrow;
                out_nk[c * k + r] = src[word * 4 + swap12[byte]];
            }
    });
    return FTCF_OK;
}
