// Skinny weight-streaming GEMM for the decode step (m <= 32 tokens per pass): y[m,n] = x[m,k] . W[k,n].
//
// Stands in for the reference's fpA_intB CUTLASS GEMM at decode shapes
// (kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:461-570; tile heuristic cutlass_heuristic.cc:128-212)
// and for cuBLAS on the fp16 / LM-head GEMMs (utils/cublasMMWrapper.cc:154-328, models/gptneox/GptNeoX.cc:869-912).
//
// Roofline: HBM.  At m <= 32 every weight byte is used for <= 64 flop, so the only thing that matters is streaming
// W once at full bandwidth.  Design:
//   * W is stored K-major (W^T, [n][k]); a CTA owns 16*RT output features and ALL of k, its 8 warps stride over k
//     in 128-byte-per-row steps (so every DRAM sector is consumed whole), 4*RT independent 16-byte loads in
//     flight per lane per step, no shared-memory staging of W (it is touched exactly once);
//   * u8 -> fp16 in registers: PRMT builds 0x64xx (1024 + b), one HSUB2 subtracts 1152 -> exact (b - 128);
//   * the 16 x (8*MT) x k product runs on mma.sync.m16n8k16 with the weights as the 16-row operand and the tokens
//     as the 8-column operand; fp32 accumulators; the contraction index is relabelled so that each lane's operand
//     fragments come from ONE contiguous 16-byte piece of a weight row (no shuffles, no transposes);
//   * cross-warp (split-k inside the CTA) reduction through shared memory in a fixed order -> deterministic;
//   * epilogue: per-column scale (fp32), bias, tanh-GELU, fp16 (or fp32 logits) store.
// Since round 2 the INT8 decode GEMMs run on tcgen05 (gemm_decode.cu); this kernel streams the fp16 weights (LM head, int8_mode = 0
// layers) and stays selectable for INT8 (impl = 1) as the A/B baseline.
#include <algorithm>

#include "tma_utils.cuh"

namespace ftcf {

enum { EPI_W8 = 0, EPI_F16 = 1, EPI_F32 = 2 };

__device__ __forceinline__ int cw_of(int warp) { return warp - 1; }   // consumer-warp index (warp 0 is the producer)
__device__ __forceinline__ uint32_t u4_get(const uint4& v, int i)
{
    return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w));
}

// tunables (ftcf_set_tunable)
std::atomic<int> g_sk_target_ctas{0};    // CTAs one launch aims for, all co-resident (0: automatic, see skinny_shape)
std::atomic<int> g_sk_evict_first{1};    // weight tiles are loaded with an L2 evict_first policy
std::atomic<int> g_sk_carveout{1};       // 1: ask for the maximum shared-memory carveout (3 CTAs per SM fit)

namespace sk {
constexpr int ROWS = 32;                 // output features per pass (two 16-row MMA tiles)
constexpr int CHUNK = 512;               // bytes of one weight row per stage (4 k-steps of 128 bytes)
constexpr int SUB_BYTES = ROWS * 128;    // one TMA box: 32 rows x 128 bytes, SWIZZLE_128B
constexpr int STAGE_BYTES = 4 * SUB_BYTES;
constexpr int STAGES = 4;
constexpr int THREADS = 288;             // warp 0: TMA producer, warps 1..8: consumers
using namespace tma;
}  // namespace sk

// Pipelined skinny GEMM.  One CTA owns the contiguous feature rows [r0, r1) and all of k.
//   producer      : one thread issues four TMA boxes (32 rows x 128 bytes, SWIZZLE_128B) per stage, 4 stages = 64 KB in flight
//                   per CTA and no registers; it never waits for the producer KERNEL (weights are constants), so with PDL
//                   the ring is already full when the previous kernel retires;
//   consumer warps: warp (tile r, k-step ks) reads its 16 x 128-byte fragment (4 LDS.128 per lane, conflict-free through the
//                   128-byte swizzle), converts
//                   u8 -> fp16 in registers and issues mma.sync.m16n8k16 against the token fragments (read through L1);
//   end of a pass : the four k-step warps of a tile are summed through shared memory in a fixed order, epilogue, store.
template <typename WT, int MT, int EPI, bool PRO = false>
__global__ void __launch_bounds__(sk::THREADS, (MT <= 2 ? 2 : 1))
gemm_skinny_kernel(const __grid_constant__ CUtensorMap map_w, const SkPro pro, const __half* __restrict__ x, const __half* __restrict__ scale,
                   const __half* __restrict__ bias, void* __restrict__ y, int m, int n, int k, int ldy, int act, int rows_per_cta,
                   int evict_first, int R)
{
    using namespace sk;
    constexpr int EPC = 16 / sizeof(WT);   // k-elements per 16-byte chunk
    constexpr int KSTEP = 8 * EPC;         // k-elements per 128-byte k-step
    constexpr int NMMA = EPC / 4;          // mma per 16-byte chunk
    constexpr int XV = EPC / 4;            // uint4 per lane per token group per k-step
    constexpr int NT = 8 * MT;

    extern __shared__ __align__(1024) uint8_t sk_smem_raw[];
    uint8_t* sk_smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(sk_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[STAGES], bar_empty[STAGES];
    __shared__ float red[8][NT][16 + 4];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int r0 = blockIdx.x * rows_per_cta, r1 = min(n, r0 + rows_per_cta);
    const int m0 = blockIdx.y * NT;
    const int row_bytes = k * (int)sizeof(WT);
    const int chunks = (row_bytes + CHUNK - 1) / CHUNK;
    const int kc0 = 0, kc1 = chunks;
    // R = rows per pass = height of the TMA box (<= ROWS; each box still owns a 32-row slot of the stage so that the
    // 128-byte swizzle pattern stays 1024-byte aligned): lets the host cut n into equal shares for ALL co-resident CTAs
    const int passes = (r1 - r0 + R - 1) / R;
    const int total = passes * chunks;       // stages this CTA streams
    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_empty[s], 8);
        }
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    __syncthreads();

    if (warp == 0) {
        // ================= producer (warp-uniform loop, the elected lane issues: no per-instruction waterfall) =================
        {
            if (elect_one_sync()) prefetch_map(&map_w);
            const uint64_t pol = l2_policy_evict_first();
            for (int i = 0; i < total; ++i) {
                const int s = i % STAGES;
                const uint32_t ph = (i / STAGES) & 1;
                const int pass = i / chunks, kc = kc0 + i % chunks;
                const int nsub = min(4, (row_bytes - kc * CHUNK) / 128);     // 128-byte k-steps in this chunk
                mbar_wait(&bar_empty[s], ph ^ 1);
                if (elect_one_sync()) {
                    mbar_arrive_expect_tx(&bar_full[s], (uint32_t)(nsub * R * 128));
                    if (evict_first) {
                        for (int j = 0; j < nsub; ++j)
                            load_2d_hint(sk_smem + (size_t)s * STAGE_BYTES + j * SUB_BYTES, &map_w, &bar_full[s],
                                         (kc * CHUNK + j * 128) / (int)sizeof(WT), r0 + pass * R, pol);
                    } else {
                        for (int j = 0; j < nsub; ++j)
                            load_2d(sk_smem + (size_t)s * STAGE_BYTES + j * SUB_BYTES, &map_w, &bar_full[s],
                                    (kc * CHUNK + j * 128) / (int)sizeof(WT), r0 + pass * R);
                    }
                }
                __syncwarp();
            }
        }
        return;
    }

    // ================= consumers =================
    const int cw = warp - 1, tile = cw >> 2, ks = cw & 3;
    const int g = lane >> 2, t = lane & 3;
    const bool trc_who = threadIdx.x == 32;
    const unsigned long long trc_t0 = trc_now(trc_who);
    pdl_wait();
    const unsigned long long trc_t1 = trc_now(trc_who);
    unsigned long long trc_t2 = 0;                              // x comes from the previous kernel; y may still be read by it
    const __half* xrow[MT];
    if constexpr (PRO) {
        // ---- fused residual + LayerNorm into shared memory (row pitch k + 8 halves: rows fall into different banks)
        __half* xs = reinterpret_cast<__half*>(sk_smem + (size_t)STAGES * STAGE_BYTES);
        const int pitch = k + 8, nvec = k >> 3, ct = threadIdx.x - 32;
        float* pred = &red[0][0][0];
        const bool writer = pro.x_out != nullptr && blockIdx.x == 0 && blockIdx.y == 0;
        for (int b = 0; b < m; ++b) {
            float s = 0.f, ss = 0.f;
            for (int vi = ct; vi < nvec; vi += 256) {
                uint4 v = *reinterpret_cast<const uint4*>(pro.x + (size_t)b * k + vi * 8);
                if (pro.add_ffn != nullptr) {
                    const uint4 fv = *reinterpret_cast<const uint4*>(pro.add_ffn + (size_t)b * k + vi * 8);
                    const uint4 av = *reinterpret_cast<const uint4*>(pro.add_attn + (size_t)b * k + vi * 8);
                    uint4 bv = make_uint4(0, 0, 0, 0);
                    if (pro.add_bias != nullptr) bv = *reinterpret_cast<const uint4*>(pro.add_bias + vi * 8);
                    __half2* xh = reinterpret_cast<__half2*>(&v);
                    const __half2* fh = reinterpret_cast<const __half2*>(&fv);
                    const __half2* ah = reinterpret_cast<const __half2*>(&av);
                    const __half2* bh = reinterpret_cast<const __half2*>(&bv);
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        __half2 r = __hadd2(fh[j], ah[j]);
                        if (pro.add_bias != nullptr) r = __hadd2(r, bh[j]);
                        xh[j] = __hadd2(r, xh[j]);
                    }
                    if (writer) *reinterpret_cast<uint4*>(pro.x_out + (size_t)b * k + vi * 8) = v;
                }
                *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8) = v;
                const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = __half22float2(vh[j]);
                    s += f.x + f.y;
                    ss += f.x * f.x + f.y * f.y;
                }
            }
            s = warp_sum(s);
            ss = warp_sum(ss);
            if (lane == 0) {
                pred[2 * cw_of(warp)] = s;
                pred[2 * cw_of(warp) + 1] = ss;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            float ts = 0.f, tss = 0.f;
#pragma unroll
            for (int w = 0; w < 8; ++w) {
                ts += pred[2 * w];
                tss += pred[2 * w + 1];
            }
            const float mean = ts / k;
            const float rstd = rsqrtf(tss / k - mean * mean + pro.eps);
            const __half2 mean_h = __float2half2_rn(mean), rstd_h = __float2half2_rn(rstd);
            for (int vi = ct; vi < nvec; vi += 256) {
                uint4 v = *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8);
                const uint4 gq = ld_ro_16(pro.gamma + vi * 8);
                const uint4 bq = ld_ro_16(pro.beta + vi * 8);
                const __half2* gh = reinterpret_cast<const __half2*>(&gq);
                const __half2* bh = reinterpret_cast<const __half2*>(&bq);
                __half2* vh = reinterpret_cast<__half2*>(&v);
#pragma unroll
                for (int j = 0; j < 4; ++j) vh[j] = __hfma2(__hmul2_rn(__hsub2_rn(vh[j], mean_h), rstd_h), gh[j], bh[j]);
                *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8) = v;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
        }
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int tok = min(m0 + mt * 8 + g, m - 1);
            xrow[mt] = xs + (size_t)tok * pitch + ks * KSTEP + t * 2 * EPC;
        }
    } else {
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
            const int tok = min(m0 + mt * 8 + g, m - 1);
            xrow[mt] = x + (size_t)tok * k + ks * KSTEP + t * 2 * EPC;
        }
    }
    int i = 0;
    for (int pass = 0; pass < passes; ++pass) {
        float acc[MT][4], acc1[MT][4];       // two accumulator chains: consecutive MMAs do not wait for each other
#pragma unroll
        for (int mt = 0; mt < MT; ++mt)
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][j] = acc1[mt][j] = 0.f;
        for (int kc = kc0; kc < kc1; ++kc, ++i) {
            const int s = i % STAGES;
            const uint32_t ph = (i / STAGES) & 1;
            const bool active = (kc * CHUNK + ks * 128) < row_bytes;     // last chunk of a row may be short
            uint4 xv[MT][XV];
            if (active) {
#pragma unroll
                for (int mt = 0; mt < MT; ++mt)
#pragma unroll
                    for (int q = 0; q < XV; ++q) {
                        if constexpr (PRO) xv[mt][q] = *reinterpret_cast<const uint4*>(xrow[mt] + (size_t)kc * (CHUNK / (int)sizeof(WT)) + q * 8);
                        else xv[mt][q] = ld_ro_16(xrow[mt] + (size_t)kc * (CHUNK / (int)sizeof(WT)) + q * 8);
                    }
            }
            mbar_wait(&bar_full[s], ph);
            if (i == 0) trc_t2 = trc_now(trc_who);
            if (active) {
                // box layout: row r at r * 128 bytes, 16-byte chunk c stored at chunk (c ^ (r & 7))  (SWIZZLE_128B)
                const int rl = tile * 16 + g;                      // rl & 7 == (rl + 8) & 7 == g & 7
                const uint8_t* st = sk_smem + (size_t)s * STAGE_BYTES + ks * SUB_BYTES + rl * 128;
                uint4 wv[2][2];
#pragma unroll
                for (int hh = 0; hh < 2; ++hh)
#pragma unroll
                    for (int c = 0; c < 2; ++c)
                        wv[hh][c] = *reinterpret_cast<const uint4*>(st + hh * 8 * 128 + (((2 * t + c) ^ (g & 7)) << 4));
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);
#pragma unroll
                for (int c = 0; c < 2; ++c)
#pragma unroll
                    for (int j = 0; j < NMMA; ++j) {
                        const int pi = c * (EPC / 2) + 2 * j;
                        uint32_t a0, a1, a2, a3;
                        if constexpr (sizeof(WT) == 1) {
                            u8x4_to_h2x2(u4_get(wv[0][c], j), a0, a2);
                            u8x4_to_h2x2(u4_get(wv[1][c], j), a1, a3);
                        } else {
                            a0 = u4_get(wv[0][c], 2 * j);
                            a2 = u4_get(wv[0][c], 2 * j + 1);
                            a1 = u4_get(wv[1][c], 2 * j);
                            a3 = u4_get(wv[1][c], 2 * j + 1);
                        }
#pragma unroll
                        for (int mt = 0; mt < MT; ++mt) {
                            if (j & 1) mma_16816(acc1[mt], a0, a1, a2, a3, u4_get(xv[mt][pi / 4], pi % 4), u4_get(xv[mt][(pi + 1) / 4], (pi + 1) % 4));
                            else mma_16816(acc[mt], a0, a1, a2, a3, u4_get(xv[mt][pi / 4], pi % 4), u4_get(xv[mt][(pi + 1) / 4], (pi + 1) % 4));
                        }
                    }
            } else {
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_empty[s]);
            }
        }
        // ---- sum the four k-step warps of each tile (fixed order), epilogue, store
#pragma unroll
        for (int mt = 0; mt < MT; ++mt) {
#pragma unroll
            for (int j = 0; j < 4; ++j) acc[mt][j] += acc1[mt][j];
            const int tok = mt * 8 + 2 * t;
            red[cw][tok][g] = acc[mt][0];
            red[cw][tok + 1][g] = acc[mt][1];
            red[cw][tok][g + 8] = acc[mt][2];
            red[cw][tok + 1][g + 8] = acc[mt][3];
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
        const int p0 = r0 + pass * R;
        const int pr1 = min(r1, p0 + R);          // rows of this pass that belong to this CTA
        for (int o = threadIdx.x - 32; o < ROWS * NT; o += 256) {
            const int f = o % ROWS, tok = o / ROWS;
            const int col = p0 + f, row = m0 + tok;
            if (col >= pr1 || row >= m) continue;
            const int tl = f >> 4, fl = f & 15;
            float v = red[tl * 4 + 0][tok][fl] + red[tl * 4 + 1][tok][fl];
            v += red[tl * 4 + 2][tok][fl];
            v += red[tl * 4 + 3][tok][fl];
            if constexpr (EPI == EPI_W8) {
                v *= __half2float(scale[col]);
                if (bias != nullptr) v += __half2float(bias[col]);
                if (act == 1) v = gelu_tanh_f32(v);
                reinterpret_cast<__half*>(y)[(size_t)row * ldy + col] = __float2half_rn(v);
            } else if constexpr (EPI == EPI_F16) {
                __half hv = __float2half_rn(v);
                if (bias != nullptr) hv = __hadd(hv, bias[col]);
                if (act == 1) hv = gelu_tanh_half_ref(hv);
                reinterpret_cast<__half*>(y)[(size_t)row * ldy + col] = hv;
            } else {
                reinterpret_cast<float*>(y)[(size_t)row * ldy + col] = v;
            }
        }
        asm volatile("bar.sync 1, 256;" ::: "memory");
    }
    if (threadIdx.x == 32) trc_emit(sizeof(WT) == 1 ? TRC_GEMM_W8 : TRC_GEMM_F16, trc_t0, trc_t1, trc_t2, n, k);
}

FTCF_TRACE_INSTALLER(trace_install_gemm_skinny)

// ---- launch shape: how many CTAs, how many 32-row passes each
// A CTA's mma.sync consumer pipeline sustains ~20-25 GB/s (measured: 148 CTAs alone reach 3.7 TB/s, 296 reach 5.7 TB/s), so a
// launch wants two CTAs on every SM, all co-resident, with whole 32-row tiles each.  (Measured and dropped in round 1: three
// CTAs per SM, equal-share box heights, k-splits, L2 prefetch of the next GEMM's weights.)
struct SkShape {
    int rows_per_cta, ctas_x;
};
static SkShape skinny_shape(int n, int m_groups, int hint)
{
    const int forced = hint > 0 ? hint : g_sk_target_ctas.load(std::memory_order_relaxed);
    const int slots = forced > 0 ? forced : 296;
    const int tiles = ceil_div(n, sk::ROWS);
    const int passes = ceil_div(tiles * m_groups, slots);
    SkShape sh;
    sh.rows_per_cta = passes * sk::ROWS;
    sh.ctas_x = ceil_div(n, sh.rows_per_cta);
    return sh;
}

template <typename WT, int EPI>
static int launch_skinny(const void* x, const void* w, const void* scale, const void* bias, void* y, int m, int n, int k,
                         int ldy, int act, const ftcf_launch_hint* hint, cudaStream_t st, const SkPro* pro = nullptr)
{
    constexpr int EPC = 16 / sizeof(WT);
    FTCF_REQUIRE(k % (8 * EPC) == 0, FTCF_ERR_UNSUPPORTED, "skinny gemm: k=%d must be a multiple of %d", k, 8 * EPC);
    FTCF_REQUIRE(m > 0 && n > 0, FTCF_ERR_INVALID, "skinny gemm: empty problem m=%d n=%d", m, n);
    const int mt = m >= 25 ? 4 : ceil_div(m, 8);
    const int m_groups = ceil_div(m, 8 * mt);
    const int want = pro != nullptr ? pro->cta_hint : (hint != nullptr ? hint->target_ctas : 0);
    const SkShape sh = skinny_shape(n, m_groups, want);
    const int rows_per_cta = sh.rows_per_cta;
    const int evict_first = g_sk_evict_first.load(std::memory_order_relaxed);
    const dim3 grid(sh.ctas_x, m_groups, 1);
    const dim3 block(sk::THREADS);
    size_t smem = (size_t)sk::STAGES * sk::STAGE_BYTES + 1024;
    SkPro prov{};
    if (pro != nullptr) {
        FTCF_REQUIRE(m <= 4, FTCF_ERR_UNSUPPORTED, "skinny gemm: the fused LayerNorm prologue takes m <= 4 rows (m=%d)", m);
        FTCF_REQUIRE(pro->x && pro->gamma && pro->beta && (pro->add_ffn == nullptr) == (pro->add_attn == nullptr), FTCF_ERR_INVALID,
                     "skinny gemm: incomplete prologue");
        prov = *pro;
        smem += (size_t)m * (k + 8) * sizeof(__half);
        FTCF_REQUIRE(smem <= 200 * 1024, FTCF_ERR_UNSUPPORTED, "skinny gemm: prologue rows do not fit shared memory (m=%d k=%d)", m, k);
    }
    const __half* xs = static_cast<const __half*>(x);
    CUtensorMap mw;
    {
        const int rc = make_tensor_map_2d(&mw, w, n, k, (int)sizeof(WT), sk::ROWS);
        if (rc != FTCF_OK) return rc;
    }
    const __half* sc = static_cast<const __half*>(scale);
    const __half* bs = static_cast<const __half*>(bias);
    const bool pdl = hint == nullptr || hint->no_pdl == 0;
    cudaError_t err = cudaSuccess;
#define FTCF_SK_(MT_, PRO_)                                                                                             \
    do {                                                                                                                \
        static std::atomic<size_t> configured{0};                                                                       \
        if (configured.load(std::memory_order_relaxed) < smem) {                                                        \
            err = cudaFuncSetAttribute(gemm_skinny_kernel<WT, MT_, EPI, PRO_>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem); \
            if (err == cudaSuccess && g_sk_carveout.load())                                                             \
                err = cudaFuncSetAttribute(gemm_skinny_kernel<WT, MT_, EPI, PRO_>, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared); \
            configured.store(smem, std::memory_order_relaxed);                                                          \
        }                                                                                                               \
        if (err == cudaSuccess)                                                                                         \
            err = launch_pdl_if(pdl, gemm_skinny_kernel<WT, MT_, EPI, PRO_>, grid, block, smem, st, mw, prov, xs, sc, bs, y, m, n, k, ldy, act, rows_per_cta, evict_first, sk::ROWS); \
    } while (0)
#define FTCF_SK(MT_) FTCF_SK_(MT_, false)
    if (pro != nullptr) {
        FTCF_SK_(1, true);
        FTCF_REQUIRE(err == cudaSuccess, FTCF_ERR_CUDA, "skinny gemm (fused prologue) launch failed: %s", cudaGetErrorString(err));
        FTCF_LAUNCH_CHECK();
        return FTCF_OK;
    }
    switch (mt) {
        case 1: FTCF_SK(1); break;
        case 2: FTCF_SK(2); break;
        case 3: FTCF_SK(3); break;
        default: FTCF_SK(4); break;
    }
#undef FTCF_SK
#undef FTCF_SK_
    FTCF_REQUIRE(err == cudaSuccess, FTCF_ERR_CUDA, "skinny gemm launch failed: %s", cudaGetErrorString(err));
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

int gemm_w8a16_skinny(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                      int act, const ftcf_launch_hint* hint, cudaStream_t st)
{
    return launch_skinny<uint8_t, EPI_W8>(x, w_nk, scale, bias, y, m, n, k, n, act, hint, st);
}

int gemm_f16_skinny(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                    int out_f32, const ftcf_launch_hint* hint, cudaStream_t st)
{
    if (out_f32) return launch_skinny<__half, EPI_F32>(x, w_nk, nullptr, bias, y, m, n, k, ldy, act, hint, st);
    return launch_skinny<__half, EPI_F16>(x, w_nk, nullptr, bias, y, m, n, k, ldy, act, hint, st);
}

static SkPro to_skpro(const ftcf_ln_prologue& p)
{
    SkPro r{};
    r.x = static_cast<const __half*>(p.x);
    r.add_ffn = static_cast<const __half*>(p.add_ffn);
    r.add_attn = static_cast<const __half*>(p.add_attn);
    r.add_bias = static_cast<const __half*>(p.add_bias);
    r.gamma = static_cast<const __half*>(p.gamma);
    r.beta = static_cast<const __half*>(p.beta);
    r.x_out = static_cast<__half*>(p.x_out);
    r.eps = p.eps;
    r.cta_hint = p.cta_hint;
    return r;
}
int gemm_w8a16_skinny_ln(const ftcf_ln_prologue& pro, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k,
                         int act, cudaStream_t st)
{
    const SkPro sp = to_skpro(pro);
    return launch_skinny<uint8_t, EPI_W8>(pro.x, w_nk, scale, bias, y, m, n, k, n, act, nullptr, st, &sp);
}
int gemm_f16_skinny_ln(const ftcf_ln_prologue& pro, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                       int out_f32, cudaStream_t st)
{
    const SkPro sp = to_skpro(pro);
    if (out_f32) return launch_skinny<__half, EPI_F32>(pro.x, w_nk, nullptr, bias, y, m, n, k, ldy, act, nullptr, st, &sp);
    return launch_skinny<__half, EPI_F16>(pro.x, w_nk, nullptr, bias, y, m, n, k, ldy, act, nullptr, st, &sp);
}

// ---------------------------------------------------------------- fp16 [k,n] -> [n,k]
__global__ void transpose_f16_kernel(const __half* __restrict__ in, __half* __restrict__ out, int k, int n)
{
    __shared__ __half tile[32][33];
    const int bx = blockIdx.x * 32, by = blockIdx.y * 32;   // bx over n, by over k
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = by + i, c = bx + threadIdx.x;
        if (r < k && c < n) tile[i][threadIdx.x] = in[(size_t)r * n + c];
    }
    __syncthreads();
    for (int i = threadIdx.y; i < 32; i += blockDim.y) {
        const int r = bx + i, c = by + threadIdx.x;          // out row = n index, col = k index
        if (r < n && c < k) out[(size_t)r * k + c] = tile[threadIdx.x][i];
    }
}

}  // namespace ftcf

extern "C" int ftcf_transpose_f16(const void* in_kn, void* out_nk, int k, int n, void* stream)
{
    using namespace ftcf;
    FTCF_REQUIRE(k > 0 && n > 0, FTCF_ERR_INVALID, "transpose: empty matrix");
    transpose_f16_kernel<<<dim3(ceil_div(n, 32), ceil_div(k, 32)), dim3(32, 8), 0, as_stream(stream)>>>(
        static_cast<const __half*>(in_kn), static_cast<__half*>(out_nk), k, n);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
