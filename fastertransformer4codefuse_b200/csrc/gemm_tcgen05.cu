// tcgen05 GEMM for the prefill (and large-batch) shapes: y[m,n] = act(x[m,k] . dequant(W)[k,n] + bias).
//
// Stands in for the reference's CUTLASS 2.x mixed-input GEMM (kernels/cutlass_kernels/fpA_intB_gemm/
// fpA_intB_gemm_template.h:45-197,461-570; mainloop cutlass_extensions/gemm/threadblock/dq_mma_multistage.h:475-522)
// -- which has no path for sm >= 90 -- and for cuBLAS on the fp16 GEMMs.  Nothing is shared with either: this is a
// Blackwell design.
//
//   * swap-AB: the WEIGHTS are the M = 128 operand of tcgen05.mma (one CTA owns 128 output features), the TOKENS are
//     the N operand (16..128 per CTA).  W is stored K-major (W^T, [n][k]) so both operands are K-major.
//   * TMA (cp.async.bulk.tensor, SWIZZLE_128B) stages 128 x 128-byte weight tiles and N x 128-byte activation tiles
//     into a 4-8 deep shared-memory ring, one elected producer thread, mbarrier complete_tx.
//   * INT8 weights: tcgen05 has no int8 x fp16 kind, so eight converter warps read the u8 tile from shared memory
//     (conflict-free thanks to the 128B swizzle), turn each byte into fp16 (PRMT + one HSUB2: 0x64xx - 1152 = b - 128,
//     exact) and write it with tcgen05.st straight into TENSOR MEMORY as the A operand (TS-form MMA).  The fp16 copy of
//     W never touches shared memory: shared-memory traffic per tile is 16 KB (u8 read) instead of 80 KB.
//   * one thread issues tcgen05.mma.cta_group::1.kind::f16 (M128 x N x K16, fp32 accumulators in TMEM); tcgen05.commit
//     releases the shared-memory stage and the TMEM A-stage back to their producers.
//   * epilogue: tcgen05.ld of the accumulator rows, per-feature dequant scale (fp32), bias, tanh-GELU, fp16/fp32 store.
//   * fp16 weights (int8_mode = 0, LM head): same pipeline, A comes from shared memory through a UMMA descriptor.
#include <algorithm>

#include "umma.cuh"

namespace ftcf {

bool splitk_scratch_acquire(cudaStream_t st, size_t part_elems, int tickets_needed, float** part, int** tickets);   // gemm_decode.cu
std::atomic<int> g_tc_ksplit{1};   // tunable "tc_ksplit": k-splits for decode-size launches of the tcgen05 GEMM

namespace tc {

constexpr int kThreads = 320;          // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..9: convert + epilogue
constexpr int kConvWarps = 8;
constexpr int kTileM = 128;            // output features per CTA (UMMA M)
constexpr int kAStages = 4;            // TMEM A-operand stages (u8 path), 64 columns each
constexpr uint32_t kTmemCols = 512;
constexpr int kGroupM = 16;           // token tiles per rasterisation group
using namespace umma;
enum { EPI_W8 = 0, EPI_F16 = 1, EPI_F32 = 2 };

struct Args {
    const __half* scale;
    const __half* bias;
    void* y;
    int m, n, k, ldy, act;
    float* part;      // split-K (gridDim.z > 1): fp32 partial sums [z][m][n]; the last CTA of a tile adds them in the order 0, 1, ...
    int* tickets;     // one self-resetting counter per (feature tile, token tile)
};

// W8 = true : A tile = 128 rows x 128 u8  (BK = 128), converted into TMEM
// W8 = false: A tile = 128 rows x 64 fp16 (BK = 64), consumed from shared memory
template <bool W8, int NT, int STAGES, int EPI, bool SPLIT = false>
__global__ void __launch_bounds__(kThreads, 1)
gemm_tc_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Args args)
{
    constexpr int BK = W8 ? 128 : 64;
    constexpr int XSUB = BK / 64;                         // 64-element (128-byte) activation sub-tiles per stage
    constexpr uint32_t W_BYTES = kTileM * 128;            // 16 KB either way
    constexpr uint32_t X_BYTES = NT * 128 * XSUB;
    constexpr uint32_t STAGE_BYTES = W_BYTES + X_BYTES;
    constexpr uint32_t D_COL = 0, A_COL = 256;            // TMEM columns: accumulators [0, NT), A stages [256, 512)

    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[STAGES], bar_w_empty[STAGES], bar_x_empty[STAGES];
    __shared__ uint64_t bar_a_full[kAStages], bar_a_empty[kAStages], bar_d_full;
    __shared__ uint32_t s_tmem_base;

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    // Grouped rasterisation of the 1-D grid: consecutive CTAs (= the CTAs that run together) cover kGroupM token tiles x a run of
    // feature tiles, so one wave re-uses each weight tile kGroupM times and each activation tile ~148 / kGroupM times out of L2.
    // (Round 1 ran feature-tile-major: every token tile streamed the whole weight matrix from DRAM again -- ncu: 729 MB read for
    // a 105 MB matrix at T = 1024.)
    const int tiles_n = (args.n + kTileM - 1) / kTileM, tiles_m = (args.m + NT - 1) / NT;
    int tile_m, tile_n;
    {
        const int per_group = kGroupM * tiles_n;
        const int g = (int)blockIdx.x / per_group, r = (int)blockIdx.x % per_group;
        const int gm = min(kGroupM, tiles_m - g * kGroupM);      // the last group may be short
        tile_m = g * kGroupM + r % gm;
        tile_n = r / gm;
    }
    const int n0 = tile_n * kTileM, m0 = tile_m * NT;
    // split-K: this CTA contracts k-blocks [kb0, kb0 + num_kb) only (decode-size launches with few feature tiles: n = 5120
    // gives 40 CTAs for 148 SMs; three k-splits stream the same weights with 120)
    const int kb_all = args.k / BK;
    const int kb_per = SPLIT ? (kb_all + (int)gridDim.z - 1) / (int)gridDim.z : kb_all;
    const int kb0 = SPLIT ? (int)blockIdx.z * kb_per : 0;
    const int num_kb = SPLIT ? max(0, min(kb_all, kb0 + kb_per) - kb0) : kb_all;
    __shared__ int s_last;

    if (threadIdx.x == 0) {
        for (int s = 0; s < STAGES; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_w_empty[s], W8 ? kConvWarps : 1);   // one arrival per converter WARP (per-thread arrivals
            // serialise on the barrier word: ~2000 cycles per K block, ncu r2c)
            mbar_init(&bar_x_empty[s], 1);
        }
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&bar_a_full[s], kConvWarps);
            mbar_init(&bar_a_empty[s], 1);
        }
        mbar_init(&bar_d_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    if (warp == 1) {
        asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(&s_tmem_base)), "r"(kTmemCols)
                     : "memory");
        asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem = s_tmem_base;

    // Single-thread instructions with uniform operands (TMA, tcgen05.mma, tcgen05.commit) are issued by the elected lane of a warp
    // that runs its role in warp-uniform control flow; behind `if (lane == 0)` every one of them costs a ~120-cycle waterfall loop.
    if (warp == 0) {
        // ================= TMA producer =================
        if (tma::elect_one_sync()) {
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_w) : "memory");
            asm volatile("prefetch.tensormap [%0];" ::"l"(&map_x) : "memory");
        }
        __syncwarp();
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            mbar_wait(&bar_w_empty[s], ph ^ 1);
            mbar_wait(&bar_x_empty[s], ph ^ 1);
            if (tma::elect_one_sync()) {
                uint8_t* st = smem + (size_t)s * STAGE_BYTES;
                mbar_arrive_expect_tx(&bar_full[s], STAGE_BYTES);
                tma_load_2d(st, &map_w, &bar_full[s], (kb0 + kb) * BK, n0);
#pragma unroll
                for (int xs = 0; xs < XSUB; ++xs)
                    tma_load_2d(st + W_BYTES + xs * NT * 128, &map_x, &bar_full[s], (kb0 + kb) * BK + xs * 64, m0);
            }
            __syncwarp();
        }
    } else if (warp == 1) {
        // ================= MMA issuer =================
        const uint32_t idesc = umma_idesc_f16(NT);
        for (int kb = 0; kb < num_kb; ++kb) {
            const int s = kb % STAGES;
            const uint32_t ph = (kb / STAGES) & 1;
            const uint32_t st_addr = smem_u32(smem + (size_t)s * STAGE_BYTES);
            mbar_wait(&bar_full[s], ph);
            if constexpr (W8) {
                const int as = kb % kAStages;
                const uint32_t aph = (kb / kAStages) & 1;
                mbar_wait(&bar_a_full[as], aph);
                tc_fence_after();
                if (tma::elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint32_t a_t = tmem + A_COL + as * 64 + ks * 8;
                        const uint32_t xb = st_addr + W_BYTES + (ks / 4) * NT * 128 + (ks % 4) * 32;
                        mma_ts(tmem + D_COL, a_t, umma_desc_k128(xb), idesc, (kb | ks) != 0);
                    }
                    tc_commit(&bar_a_empty[as]);
                    tc_commit(&bar_x_empty[s]);
                    if (kb == num_kb - 1) tc_commit(&bar_d_full);
                }
            } else {
                tc_fence_after();
                if (tma::elect_one_sync()) {
#pragma unroll
                    for (int ks = 0; ks < BK / 16; ++ks) {
                        const uint32_t wa = st_addr + ks * 32;
                        const uint32_t xb = st_addr + W_BYTES + ks * 32;
                        mma_ss(tmem + D_COL, umma_desc_k128(wa), umma_desc_k128(xb), idesc, (kb | ks) != 0);
                    }
                    tc_commit(&bar_w_empty[s]);
                    tc_commit(&bar_x_empty[s]);
                    if (kb == num_kb - 1) tc_commit(&bar_d_full);
                }
            }
            __syncwarp();
        }
        if (num_kb == 0) {
            if (tma::elect_one_sync()) tc_commit(&bar_d_full);   // nothing was issued: the commit completes at once
            __syncwarp();
        }
    } else {
        // ================= converter warps (u8 -> fp16 -> TMEM), then epilogue =================
        const int q = warp & 3;                 // TMEM lane quarter this warp may touch
        const int hf = (warp - 2) >> 2;         // which half of the k-bytes (convert) / token columns (epilogue)
        const int row = q * 32 + lane;          // feature row inside the tile == TMEM lane
        if constexpr (W8) {
            for (int kb = 0; kb < num_kb; ++kb) {
                const int s = kb % STAGES;
                const uint32_t ph = (kb / STAGES) & 1;
                const int as = kb % kAStages;
                const uint32_t aph = (kb / kAStages) & 1;
                mbar_wait(&bar_full[s], ph);
                const uint32_t wt = smem_u32(smem + (size_t)s * STAGE_BYTES + row * 128);
                uint4 v[4];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    const int chunk = (hf * 4 + c) ^ (row & 7);      // SWIZZLE_128B: 16-byte chunk index XOR row % 8
                    v[c] = tma::lds_128(wt + chunk * 16);
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_w_empty[s]);         // every lane of this warp has issued its reads of the u8 tile
                uint32_t r[32];
#pragma unroll
                for (int c = 0; c < 4; ++c) {
                    u8x4_to_h2x2(v[c].x, r[c * 8 + 0], r[c * 8 + 1]);
                    u8x4_to_h2x2(v[c].y, r[c * 8 + 2], r[c * 8 + 3]);
                    u8x4_to_h2x2(v[c].z, r[c * 8 + 4], r[c * 8 + 5]);
                    u8x4_to_h2x2(v[c].w, r[c * 8 + 6], r[c * 8 + 7]);
                }
                mbar_wait(&bar_a_empty[as], aph ^ 1);
                tc_fence_after();
                tmem_st_x32(tmem + ((uint32_t)(q * 32) << 16) + A_COL + as * 64 + hf * 32, r);
                asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
                tc_fence_before();
                __syncwarp();
                if (lane == 0) mbar_arrive(&bar_a_full[as]);
            }
        }
        // ---- epilogue: this warp owns accumulator rows [32q, 32q+32) and token columns [hf*NT/2, (hf+1)*NT/2)
        mbar_wait(&bar_d_full, 0);
        tc_fence_after();
        const int col = n0 + row;
        const bool col_ok = col < args.n;
        float sc = 1.f, bs = 0.f;
        if constexpr (EPI == EPI_W8) {
            if (col_ok) sc = __half2float(args.scale[col]);
        }
        if (args.bias != nullptr && col_ok) bs = __half2float(args.bias[col]);
        const int S = SPLIT ? (int)gridDim.z : 1;
        bool finish = true;
        if constexpr (SPLIT) {
            // publish this k-split's partial accumulators, take a ticket; only the last arriver of the tile goes on
#pragma unroll
            for (int c0 = 0; c0 < NT / 2; c0 += 8) {
                uint32_t acc[8];
                const int tcol = hf * (NT / 2) + c0;
                tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + D_COL + tcol, acc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int tok = m0 + tcol + j;
                    if (col_ok && tok < args.m) args.part[((size_t)blockIdx.z * args.m + tok) * args.n + col] = num_kb > 0 ? __uint_as_float(acc[j]) : 0.f;
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight converter / epilogue warps
            if (threadIdx.x == 64) {
                const int tile_id = (int)blockIdx.x;
                const int old = atomicAdd(&args.tickets[tile_id], 1);
                s_last = old == S - 1;
                if (old == S - 1) args.tickets[tile_id] = 0;
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            finish = s_last != 0;
            if (finish) __threadfence();
        }
#pragma unroll
        for (int c0 = 0; c0 < NT / 2; c0 += 8) {
            if (!finish) break;
            uint32_t acc[8];
            const int tcol = hf * (NT / 2) + c0;
            if constexpr (!SPLIT) tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + D_COL + tcol, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tok = m0 + tcol + j;
                if (!col_ok || tok >= args.m) continue;
                float v;
                if constexpr (!SPLIT) {
                    v = __uint_as_float(acc[j]);
                } else {
                    v = __ldcg(&args.part[(size_t)tok * args.n + col]);
                    for (int z = 1; z < S; ++z) v += __ldcg(&args.part[((size_t)z * args.m + tok) * args.n + col]);
                }
                if constexpr (EPI == EPI_W8) {
                    v = v * sc + bs;
                    if (args.act == 1) v = gelu_tanh_f32(v);
                    reinterpret_cast<__half*>(args.y)[(size_t)tok * args.ldy + col] = __float2half_rn(v);
                } else if constexpr (EPI == EPI_F16) {
                    __half hv = __float2half_rn(v);
                    if (args.bias != nullptr) hv = __hadd(hv, __float2half_rn(bs));
                    if (args.act == 1) hv = gelu_tanh_half_ref(hv);
                    reinterpret_cast<__half*>(args.y)[(size_t)tok * args.ldy + col] = hv;
                } else {
                    reinterpret_cast<float*>(args.y)[(size_t)tok * args.ldy + col] = v;
                }
            }
        }
        tc_fence_before();
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

// ------------------------------------------------------------------------------------------------ host side
template <bool W8, int NT, int STAGES, int EPI>
static int launch(const CUtensorMap& mw, const CUtensorMap& mx, const Args& a, cudaStream_t st)
{
    constexpr int BK = W8 ? 128 : 64;
    constexpr size_t smem = (size_t)STAGES * (kTileM * 128 + NT * 128 * (BK / 64)) + 1024;
    auto kern = gemm_tc_kernel<W8, NT, STAGES, EPI, false>;
    auto kern_split = gemm_tc_kernel<W8, NT, STAGES, EPI, true>;
    static bool configured = false;
    if (!configured) {
        FTCF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FTCF_CUDA_CHECK(cudaFuncSetAttribute(kern_split, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        configured = true;
    }
    dim3 grid(ceil_div(a.n, kTileM) * ceil_div(a.m, NT));
    Args aa = a;
    aa.part = nullptr;
    aa.tickets = nullptr;
    const int tiles = (int)grid.x, kb_all = a.k / BK;
    // (measured on the 13B decode step: batch 32 12.2 -> 10.3 ms with the split, batch 16 8.1 -> 9.2 ms: only above 16 rows)
    if (g_tc_ksplit.load(std::memory_order_relaxed) != 0 && a.m > 16 && tiles * 2 <= 148 && kb_all >= 16) {
        int S = std::min(std::min(148 / tiles, 4), kb_all / 8);
        if (S >= 2 && splitk_scratch_acquire(st, (size_t)S * a.m * a.n, tiles, &aa.part, &aa.tickets)) grid.z = S;
    }
    if (grid.z > 1) kern_split<<<grid, kThreads, smem, st>>>(mw, mx, aa);
    else kern<<<grid, kThreads, smem, st>>>(mw, mx, aa);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

template <bool W8, int EPI>
static int dispatch(const void* x, const void* w, const Args& a, cudaStream_t st)
{
    constexpr int elem = W8 ? 1 : 2;
    CUtensorMap mw, mx;
    const int nt = a.m <= 16 ? 16 : (a.m <= 32 ? 32 : (a.m <= 64 ? 64 : 128));
    int rc = make_tensor_map_2d(&mw, w, a.n, a.k, elem, kTileM);
    if (rc != FTCF_OK) return rc;
    rc = make_tensor_map_2d(&mx, x, a.m, a.k, 2, nt);
    if (rc != FTCF_OK) return rc;
    switch (nt) {
        case 16: return launch<W8, 16, 8, EPI>(mw, mx, a, st);
        case 32: return launch<W8, 32, 8, EPI>(mw, mx, a, st);
        case 64: return launch<W8, 64, 6, EPI>(mw, mx, a, st);
        default: return launch<W8, 128, 4, EPI>(mw, mx, a, st);
    }
}

}  // namespace tc

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn get_encode()
{
    static EncodeTiledFn fn = nullptr;
    if (fn) return fn;
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) != cudaSuccess || qres != cudaDriverEntryPointSuccess)
        return nullptr;
    fn = reinterpret_cast<EncodeTiledFn>(p);
    return fn;
}

// 2-D row-major [rows, cols] tensor of `elem` bytes, box = [box_rows, 128 bytes], SWIZZLE_128B, zero fill out of bounds
int make_tensor_map_2d(CUtensorMap* map, const void* base, int rows, int cols, int elem, int box_rows)
{
    EncodeTiledFn enc = get_encode();
    FTCF_REQUIRE(enc != nullptr, FTCF_ERR_CUDA, "cuTensorMapEncodeTiled is not available from the driver");
    const cuuint64_t dims[2] = {(cuuint64_t)cols, (cuuint64_t)rows};
    const cuuint64_t strides[1] = {(cuuint64_t)cols * elem};
    const cuuint32_t box[2] = {(cuuint32_t)(128 / elem), (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    const CUresult r = enc(map, elem == 1 ? CU_TENSOR_MAP_DATA_TYPE_UINT8 : CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<void*>(base),
                           dims, strides, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B,
                           CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    FTCF_REQUIRE(r == CUDA_SUCCESS, FTCF_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d) for a [%d x %d] tensor of %d-byte elements", (int)r,
                 rows, cols, elem);
    return FTCF_OK;
}


bool gemm_tcgen05_supported(int m, int n, int k, int elem_bytes)
{
    const int bk = elem_bytes == 1 ? 128 : 64;
    return m > 0 && n > 0 && k >= bk && k % bk == 0 && (k * elem_bytes) % 16 == 0;
}

int gemm_w8a16_tcgen05(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k, int act,
                       cudaStream_t st)
{
    FTCF_REQUIRE(gemm_tcgen05_supported(m, n, k, 1), FTCF_ERR_UNSUPPORTED, "tcgen05 w8a16 gemm: k=%d must be a multiple of 128", k);
    tc::Args a{static_cast<const __half*>(scale), static_cast<const __half*>(bias), y, m, n, k, n, act, nullptr, nullptr};
    return tc::dispatch<true, tc::EPI_W8>(x, w_nk, a, st);
}

int gemm_f16_tcgen05(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act, int out_f32,
                     cudaStream_t st)
{
    FTCF_REQUIRE(gemm_tcgen05_supported(m, n, k, 2), FTCF_ERR_UNSUPPORTED, "tcgen05 f16 gemm: k=%d must be a multiple of 64", k);
    tc::Args a{nullptr, static_cast<const __half*>(bias), y, m, n, k, ldy, act, nullptr, nullptr};
    if (out_f32) return tc::dispatch<false, tc::EPI_F32>(x, w_nk, a, st);
    return tc::dispatch<false, tc::EPI_F16>(x, w_nk, a, st);
}

}  // namespace ftcf
