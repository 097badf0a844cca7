// Decode-step weight-only INT8 GEMM on tcgen05 (m <= 32 tokens): y[m,n] = act(x[m,k] . dequant(W)[k,n] + bias).
//
// Stands in for the reference's fpA_intB CUTLASS GEMM at decode shapes (kernels/cutlass_kernels/fpA_intB_gemm/
// fpA_intB_gemm_template.h:461-570: m = 1 ... 32, every launch of the BASELINE decode configurations).  Roofline: HBM -- at
// m <= 32 each weight byte is used for <= 64 flop, so the kernel is a weight STREAM; the tensor core is there to make the
// consumer side of that stream cheap enough that one CTA per SM already saturates HBM (round 1's mma.sync consumer sustained
// ~25 GB/s per CTA: PRMT / HSUB2 / LDS / mma.sync issue slots, profiles/r1v_gemm_skinny_ncu_full.txt).
//
//   * swap-AB: 128 output features (weight rows of W^T, [n][k] K-major) are the M operand, the token rows the N = 16 / 32 operand;
//   * a CTA owns one 128-feature tile and a contiguous K range (grid.z = k-splits so that a launch fills the co-resident CTA
//     slots of the GPU); 2 CTAs per SM (<= 113 KB shared memory, 256 TMEM columns each) so that the two branches of a decode
//     layer (QKV -> attention -> O and FFN1 -> FFN2) stream side by side;
//   * warp 0: TMA producer -- 128 x 128-byte weight boxes (SWIZZLE_128B, L2 evict_first) into a 4-5 deep ring (64-80 KB in flight
//     per CTA); weights do not depend on the previous kernel, so with programmatic dependent launch the ring is full before the
//     producer kernel has retired; activation boxes follow after griddepcontrol.wait;
//   * warps 2-9: u8 -> fp16 (PRMT + one HSUB2, exact) straight into TENSOR MEMORY as the A operand (tcgen05.st); the fp16 copy of
//     W never exists in shared memory;
//   * warp 1: one thread issues tcgen05.mma.kind::f16 (TS form, M128 x N x K16, fp32 accumulators in TMEM), tcgen05.commit hands
//     the A stage and the activation slab back;
//   * epilogue from tcgen05.ld: per-feature dequant scale (fp32), bias, tanh-GELU, fp16 store -- or, with tensor parallelism,
//     8-byte flagged stores of the tile into every rank's exchange area (ftcf_tp_exchange: the all-reduce of the layer);
//   * k-splits of a tile: a thread-block cluster whose members add their accumulators into the leader's shared memory over
//     DSMEM in rank order (when the buffer is <= 16 KB), else fp32 partials in a per-stream scratch summed in a fixed order by
//     the last CTA of the tile (ticket) -- both deterministic; never more CTAs than co-resident slots;
//   * single-thread instructions (TMA, tcgen05.mma, tcgen05.commit) are issued by the lane `elect.sync` picks, from warp-uniform
//     control flow: behind `if (lane == 0)` the compiler serialises them per lane (118 vs 25 cycles per MMA, tools/umma_probe.cu);
//   * optional fused prologue (m <= 4): previous layer's residual add + LayerNorm computed by every CTA into shared memory
//     (common.cuh, SkPro), from which the converter warps build the swizzled activation slab of each K step.
#include <algorithm>
#include <mutex>
#include <unordered_map>

#include "umma.cuh"

namespace ftcf {

std::atomic<int> g_dg_target_ctas{240};   // tunable "decode_target_ctas": CTAs a launch aims for (k-splits fill up to it)
std::atomic<int> g_dg_min_kb{8};          // tunable "decode_min_kb": fewest 128-byte K steps a k-split may get
std::atomic<int> g_dg_evict_first{1};     // tunable "decode_evict_first"
std::atomic<int> g_dg_lean{0};            // tunable "decode_lean": 64-register variant of the kernel (3 CTAs per SM by registers)
std::atomic<int> g_dg_cluster{1};         // tunable "decode_cluster": k-splits of a tile as a thread-block cluster (DSMEM reduction)
std::atomic<int> g_dg_max_stages{4};      // tunable "decode_max_stages": cap on the weight-ring depth (16 KB per stage)

// ---- split-K scratch: one (partials, tickets) slot per STREAM.  Launches on one stream are ordered (a PDL-launched successor
// touches its scratch only after griddepcontrol.wait), so one slot per stream is enough, and engines / threads that use
// different streams never share one (round 1 handed out 8 process-wide slots round-robin).
namespace {
struct SplitSlot {
    float* part = nullptr;
    int* tickets = nullptr;
};
constexpr size_t kSlotPartElems = (size_t)4 << 20;   // fp32 partials: k-splits x m x n  (8 x 32 x 16384)
constexpr int kSlotTickets = 4096;
std::mutex g_slot_mu;
std::unordered_map<cudaStream_t, SplitSlot> g_slots;
}  // namespace

int splitk_reserve_for_stream(cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_slot_mu);
    if (g_slots.count(st)) return FTCF_OK;
    SplitSlot s;
    FTCF_CUDA_CHECK(cudaMalloc(&s.part, kSlotPartElems * sizeof(float)));
    FTCF_CUDA_CHECK(cudaMalloc(&s.tickets, kSlotTickets * sizeof(int)));
    FTCF_CUDA_CHECK(cudaMemset(s.tickets, 0, kSlotTickets * sizeof(int)));
    g_slots[st] = s;
    return FTCF_OK;
}
void splitk_release_for_stream(cudaStream_t st)
{
    std::lock_guard<std::mutex> lk(g_slot_mu);
    auto it = g_slots.find(st);
    if (it == g_slots.end()) return;
    cudaFree(it->second.part);
    cudaFree(it->second.tickets);
    g_slots.erase(it);
}
// false when the stream has no slot and none can be made now (stream capture in progress) or the problem does not fit:
// callers then do not split
bool splitk_scratch_acquire(cudaStream_t st, size_t part_elems, int tickets_needed, float** part, int** tickets)
{
    if (part_elems > kSlotPartElems || tickets_needed > kSlotTickets) return false;
    {
        std::lock_guard<std::mutex> lk(g_slot_mu);
        auto it = g_slots.find(st);
        if (it != g_slots.end()) {
            *part = it->second.part;
            *tickets = it->second.tickets;
            return true;
        }
    }
    cudaStreamCaptureStatus cs = cudaStreamCaptureStatusNone;
    if (cudaStreamIsCapturing(st, &cs) != cudaSuccess || cs != cudaStreamCaptureStatusNone) return false;
    if (splitk_reserve_for_stream(st) != FTCF_OK) return false;
    return splitk_scratch_acquire(st, part_elems, tickets_needed, part, tickets);
}

namespace dg {
using namespace umma;

// Debug: per-K-step clock stamps of CTA (0,0,0) -- converter warp 2 [0..4] and the MMA warp [5..7] -- written when a buffer is
// installed with ftcf_debug_decode_probe (tools/decode_gemm_probe.py).  8 x int64 per K step.
__device__ long long* g_probe = nullptr;

constexpr int kThreads = 320;          // warp 0: TMA, warp 1: MMA + TMEM owner, warps 2..9: convert + epilogue
constexpr int kConvWarps = 8;
constexpr int kTileM = 128;            // output features per CTA (UMMA M)
constexpr int kAStages = 3;            // TMEM A-operand stages, 64 columns each
constexpr uint32_t kTmemCols = 256;    // two CTAs per SM share the 512 columns
constexpr uint32_t D_COL = 0, A_COL = 64;
constexpr int kMaxStages = 10;
constexpr uint32_t W_BYTES = kTileM * 128;   // one weight box: 128 features x 128 k (u8)
constexpr int BK = 128;

struct Args {
    const __half* scale;
    const __half* bias;
    __half* y;
    int m, n, k, ldy, act;
    float* part;      // split-K (gridDim.z > 1): fp32 partial sums [z][m][n]
    int* tickets;     // one self-resetting counter per feature tile
    int stages;       // depth of the weight ring (<= kMaxStages)
    int evict_first;
    SkPro pro;        // PRO only
    ftcf_tp_exchange push;   // push.tp > 1: the epilogue stores the output into every rank's exchange area instead of y
    int push_kind, push_layer;
    int cluster;      // 1: the gridDim.z k-splits of a tile form a thread-block cluster and reduce through distributed shared
                      // memory (the leader's `red` buffer at red_off) instead of global partials + ticket
    uint32_t red_off; // byte offset of the reduction buffer [(S-1)][m][128] fp32 inside the aligned dynamic shared memory
};

__device__ __forceinline__ uint32_t cluster_ctarank()
{
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync_all()
{
    asm volatile("barrier.cluster.arrive.release.aligned;" ::: "memory");
    asm volatile("barrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// fp32 store into the shared memory of CTA `rank` of this cluster (same offset as `local_addr` in this CTA's window)
__device__ __forceinline__ void st_cluster_f32(uint32_t local_addr, uint32_t rank, float v)
{
    uint32_t remote;
    asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(remote) : "r"(local_addr), "r"(rank));
    asm volatile("st.shared::cluster.f32 [%0], %1;" ::"r"(remote), "f"(v) : "memory");
}

// PRO = false: activations arrive by TMA (map_x: [m, k] fp16, box NT rows x 128 bytes, rows >= m read as zero)
// PRO = true : activations are built in shared memory by the fused residual + LayerNorm prologue (m <= 4)
// MINB = CTAs per SM the register allocation is bounded for: 2 -> up to 102 registers (90 used), 3 -> 64.  The lean variant leaves
// room in the register file for the attention CTAs beside two resident GEMM CTAs (tunable "decode_lean").
template <int NT, bool PRO, int MINB>
__global__ void __launch_bounds__(kThreads, MINB)
gemm_decode_kernel(const __grid_constant__ CUtensorMap map_w, const __grid_constant__ CUtensorMap map_x, const Args args)
{
    constexpr uint32_t X_BYTES = NT * 128 * 2;            // two 64-element (128-byte) activation sub-slabs per K step
    constexpr uint32_t STAGE_BYTES = W_BYTES + (PRO ? 0u : X_BYTES);

    extern __shared__ __align__(1024) uint8_t dg_smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(dg_smem_raw) + 1023) & ~(uintptr_t)1023);
    __shared__ uint64_t bar_full[kMaxStages], bar_w_empty[kMaxStages], bar_x_empty[kMaxStages];
    __shared__ uint64_t bar_a_full[kAStages], bar_a_empty[kAStages], bar_d_full;
    __shared__ uint32_t s_tmem_base;
    __shared__ int s_last;
    __shared__ float s_stat[2 * kConvWarps];

    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int stages = args.stages;
    const int n0 = blockIdx.x * kTileM;
    const int S = (int)gridDim.z;
    const int kb_all = args.k / BK;
    const int kb_per = (kb_all + S - 1) / S;
    const int kb0 = (int)blockIdx.z * kb_per;
    const int num_kb = max(0, min(kb_all, kb0 + kb_per) - kb0);
    // PRO: [weight ring][activation slabs, one per A stage][xs: m rows of LayerNorm output, pitch k + 8 halves]
    uint8_t* slab0 = smem + (size_t)stages * STAGE_BYTES;
    __half* xs = reinterpret_cast<__half*>(slab0 + (size_t)kAStages * X_BYTES);

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        for (int s = 0; s < stages; ++s) {
            mbar_init(&bar_full[s], 1);
            mbar_init(&bar_w_empty[s], kConvWarps);      // one arrival per converter WARP
            mbar_init(&bar_x_empty[s], 1);
        }
        for (int s = 0; s < kAStages; ++s) {
            mbar_init(&bar_a_full[s], kConvWarps);
            mbar_init(&bar_a_empty[s], 1);
        }
        mbar_init(&bar_d_full, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
    }
    // Bulk-copy / MMA / commit instructions take uniform operands: they are issued by the elected lane of a warp that runs its
    // role in warp-uniform control flow (tma::elect_one_sync), never from behind `if (lane == 0)`.
    auto load_w = [&](int kb, int s) {
        uint8_t* st = smem + (size_t)s * STAGE_BYTES;
        if (args.evict_first) {
            tma::load_2d_hint(st, &map_w, &bar_full[s], (kb0 + kb) * BK, n0, tma::l2_policy_evict_first());
        } else {
            tma::load_2d(st, &map_w, &bar_full[s], (kb0 + kb) * BK, n0);
        }
    };
    const int pre = min(stages, num_kb);
    if (warp == 0) {
        __syncwarp();      // the barrier initialisation above (lane 0) is visible to whichever lane is elected
        // The first ring fill goes out NOW, before this CTA has its tensor memory: weights are constants, so a CTA that was
        // launched early (programmatic dependent launch) streams them while the previous kernel still runs -- even while it
        // waits in tcgen05.alloc for that kernel's CTAs on this SM to release their columns.
        if (tma::elect_one_sync()) {
            tma::prefetch_map(&map_w);
            for (int kb = 0; kb < pre; ++kb) {
                mbar_arrive_expect_tx(&bar_full[kb], STAGE_BYTES);
                load_w(kb, kb);
            }
        }
        __syncwarp();
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
    long long* const probe = (blockIdx.x == 0 && blockIdx.z == 0) ? g_probe : nullptr;
    // converter / epilogue geometry (warps 2..9): TMEM lane quarter, half of the K step / of the token columns, feature row
    const int q = warp & 3, hf = (warp - 2) >> 2, row = q * 32 + lane, ct = threadIdx.x - 64;
    const int col = n0 + row;
    const bool col_ok = col < args.n;
    bool finish = true;                         // this CTA runs the second half of the epilogue for its tile
    unsigned long long trc_t0 = 0, trc_t1 = 0, trc_t2 = 0;
    int trc_pro_ns = 0;

    if (warp == 0) {
        // ================= TMA producer (whole warp in lock step, one elected lane issues) =================
        auto load_x = [&](int kb, int s) {
            uint8_t* st = smem + (size_t)s * STAGE_BYTES + W_BYTES;
            tma::load_2d(st, &map_x, &bar_full[s], (kb0 + kb) * BK, 0);
            tma::load_2d(st + NT * 128, &map_x, &bar_full[s], (kb0 + kb) * BK + 64, 0);
        };
        // (the weights of the first ring fill were requested at kernel entry;) the activations the previous kernel produces
        // are requested after the dependency wait
        pdl_wait();
        if constexpr (!PRO) {
            if (tma::elect_one_sync()) {
                tma::prefetch_map(&map_x);
                for (int kb = 0; kb < pre; ++kb) load_x(kb, kb);
            }
            __syncwarp();
        }
        // ring positions advance with running counters: `stages` is a launch parameter, and kb % stages / kb / stages on a run-time
        // divisor cost ~240 cycles per K step (measured with tools/decode_gemm_probe.py)
        int s = pre == stages ? 0 : pre;
        uint32_t ph = pre == stages ? 1 : 0;
        for (int kb = pre; kb < num_kb; ++kb) {
            mbar_wait(&bar_w_empty[s], ph ^ 1);
            if constexpr (!PRO) mbar_wait(&bar_x_empty[s], ph ^ 1);
            if (tma::elect_one_sync()) {
                mbar_arrive_expect_tx(&bar_full[s], STAGE_BYTES);
                load_w(kb, s);
                if constexpr (!PRO) load_x(kb, s);
            }
            __syncwarp();
            if (++s == stages) { s = 0; ph ^= 1; }
        }
    } else if (warp == 1) {
        // ================= MMA issuer (whole warp waits, one elected lane issues MMAs and commits) =================
        const uint32_t idesc = umma_idesc_f16(NT);
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            uint32_t xaddr;
            if constexpr (PRO) {
                xaddr = smem_u32(slab0 + (size_t)as * X_BYTES);
            } else {
                mbar_wait(&bar_full[s], ph);      // the activation slab of this K step has landed
                xaddr = smem_u32(smem + (size_t)s * STAGE_BYTES + W_BYTES);
            }
            const bool prb = probe != nullptr && lane == 0 && kb < 64;
            if (prb) probe[kb * 8 + 5] = clock64();
            mbar_wait(&bar_a_full[as], aph);
            if (prb) probe[kb * 8 + 6] = clock64();
            tc_fence_after();
            if (tma::elect_one_sync()) {
#pragma unroll
                for (int ks = 0; ks < BK / 16; ++ks) {
                    const uint32_t a_t = tmem + A_COL + as * 64 + ks * 8;
                    const uint32_t xb = xaddr + (ks / 4) * NT * 128 + (ks % 4) * 32;
                    mma_ts(tmem + D_COL, a_t, umma_desc_k128(xb), idesc, (kb | ks) != 0);
                }
                tc_commit(&bar_a_empty[as]);
                if constexpr (!PRO) tc_commit(&bar_x_empty[s]);
                if (kb == num_kb - 1) tc_commit(&bar_d_full);
            }
            __syncwarp();
            if (prb) probe[kb * 8 + 7] = clock64();
            if (++s == stages) { s = 0; ph ^= 1; }
            if (++as == kAStages) { as = 0; aph ^= 1; }
        }
        if (num_kb == 0) {
            if (tma::elect_one_sync()) tc_commit(&bar_d_full);   // nothing was issued: the commit completes at once
            __syncwarp();
        }
    } else {
        // ================= converter warps (u8 -> fp16 -> TMEM), then epilogue =================
        const bool trc_who = ct == 0;
        trc_t0 = trc_now(trc_who);
        trc_t1 = trc_t0;
        if constexpr (PRO) {
            // ---- zero the activation slabs once (rows >= m stay zero for the whole launch)
            for (int i = ct; i < (int)(kAStages * X_BYTES / 16); i += 256) reinterpret_cast<uint4*>(slab0)[i] = make_uint4(0, 0, 0, 0);
            asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // the tensor core reads these rows through the async proxy
            constexpr int PU = 3;                // vectors per thread per batch: 256 threads x 3 x 8 = 6144 columns in one go
            pdl_wait();                         // x / add_ffn / add_attn come from the previous kernels
            trc_t1 = trc_now(trc_who);
            // ---- fused residual + LayerNorm into xs (same arithmetic as ftcf_add_bias_attn_ffn_residual + ftcf_layernorm).
            // A thread owns up to PU 16-byte vectors of the row and issues ALL their loads before it touches any of them (one L2
            // round trip per pass instead of one per vector: the prologue is on the critical path of every layer -- 4.6 us before,
            // profiles/r2m_timeline_a.txt).
            const SkPro& pro = args.pro;
            const int k = args.k, pitch = k + 8, nvec = k >> 3;
            const bool writer = pro.x_out != nullptr && blockIdx.x == 0 && blockIdx.z == 0;
            for (int b = 0; b < args.m; ++b) {
                float sum = 0.f, sq = 0.f;
                for (int base = ct; base < nvec; base += 256 * PU) {
                    uint4 v[PU], fv[PU], av[PU], bv[PU];
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int vi = base + u * 256;
                        v[u] = fv[u] = av[u] = bv[u] = make_uint4(0, 0, 0, 0);
                        if (vi < nvec) {
                            v[u] = *reinterpret_cast<const uint4*>(pro.x + (size_t)b * k + vi * 8);
                            if (pro.add_ffn != nullptr) {
                                fv[u] = *reinterpret_cast<const uint4*>(pro.add_ffn + (size_t)b * k + vi * 8);
                                av[u] = *reinterpret_cast<const uint4*>(pro.add_attn + (size_t)b * k + vi * 8);
                                if (pro.add_bias != nullptr) bv[u] = ld_ro_16(pro.add_bias + vi * 8);
                            }
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int vi = base + u * 256;
                        if (vi >= nvec) continue;
                        if (pro.add_ffn != nullptr) {
                            __half2* xh = reinterpret_cast<__half2*>(&v[u]);
                            const __half2* fh = reinterpret_cast<const __half2*>(&fv[u]);
                            const __half2* ah = reinterpret_cast<const __half2*>(&av[u]);
                            const __half2* bh = reinterpret_cast<const __half2*>(&bv[u]);
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                __half2 r = __hadd2(fh[j], ah[j]);
                                if (pro.add_bias != nullptr) r = __hadd2(r, bh[j]);
                                xh[j] = __hadd2(r, xh[j]);
                            }
                            if (writer) *reinterpret_cast<uint4*>(pro.x_out + (size_t)b * k + vi * 8) = v[u];
                        }
                        *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8) = v[u];
                        const __half2* vh = reinterpret_cast<const __half2*>(&v[u]);
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = __half22float2(vh[j]);
                            sum += f.x + f.y;
                            sq += f.x * f.x + f.y * f.y;
                        }
                    }
                }
                sum = warp_sum(sum);
                sq = warp_sum(sq);
                if (lane == 0) {
                    s_stat[2 * (warp - 2)] = sum;
                    s_stat[2 * (warp - 2) + 1] = sq;
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
                float ts = 0.f, tss = 0.f;
#pragma unroll
                for (int w = 0; w < kConvWarps; ++w) {
                    ts += s_stat[2 * w];
                    tss += s_stat[2 * w + 1];
                }
                const float mean = ts / k;
                const float rstd = rsqrtf(tss / k - mean * mean + pro.eps);
                const __half2 mean_h = __float2half2_rn(mean), rstd_h = __float2half2_rn(rstd);
                for (int base = ct; base < nvec; base += 256 * PU) {
                    uint4 gq[PU], bq[PU];
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int vi = base + u * 256;
                        gq[u] = bq[u] = make_uint4(0, 0, 0, 0);
                        if (vi < nvec) {
                            gq[u] = ld_ro_16(pro.gamma + vi * 8);
                            bq[u] = ld_ro_16(pro.beta + vi * 8);
                        }
                    }
#pragma unroll
                    for (int u = 0; u < PU; ++u) {
                        const int vi = base + u * 256;
                        if (vi >= nvec) continue;
                        uint4 v = *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8);
                        const __half2* gh = reinterpret_cast<const __half2*>(&gq[u]);
                        const __half2* bh = reinterpret_cast<const __half2*>(&bq[u]);
                        __half2* vh = reinterpret_cast<__half2*>(&v);
#pragma unroll
                        for (int j = 0; j < 4; ++j) vh[j] = __hfma2(__hmul2_rn(__hsub2_rn(vh[j], mean_h), rstd_h), gh[j], bh[j]);
                        *reinterpret_cast<uint4*>(xs + (size_t)b * pitch + vi * 8) = v;
                    }
                }
                asm volatile("bar.sync 1, 256;" ::: "memory");
            }
        }
        trc_pro_ns = (int)(trc_now(trc_who) - trc_t1);           // duration of the fused prologue (0 without one)
        int s = 0, as = 0;
        uint32_t ph = 0, aph = 0;
        for (int kb = 0; kb < num_kb; ++kb) {
            const bool prb = probe != nullptr && warp == 2 && lane == 0 && kb < 64;
            if (prb) probe[kb * 8 + 0] = clock64();
            mbar_wait(&bar_full[s], ph);
            if (prb) probe[kb * 8 + 1] = clock64();
            if (kb == 0) trc_t2 = trc_now(trc_who);
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
            if (prb) probe[kb * 8 + 2] = clock64();
            mbar_wait(&bar_a_empty[as], aph ^ 1);
            if (prb) probe[kb * 8 + 3] = clock64();
            tc_fence_after();
            if constexpr (PRO) {
                // activation slab of this K step: m rows x 256 bytes out of xs, 16-byte chunks placed as the UMMA descriptor
                // (K-major, SWIZZLE_128B, 8-row groups of 1024 bytes) expects them
                if (ct < 16 * args.m) {
                    const int b = ct >> 4, sub = (ct >> 3) & 1, c = ct & 7;
                    const uint4 xv = *reinterpret_cast<const uint4*>(xs + (size_t)b * (args.k + 8) + (size_t)(kb0 + kb) * BK + sub * 64 + c * 8);
                    uint8_t* dst = slab0 + (size_t)as * X_BYTES + sub * (NT * 128) + (b >> 3) * 1024 + (b & 7) * 128 + ((c ^ (b & 7)) << 4);
                    *reinterpret_cast<uint4*>(dst) = xv;
                    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");   // generic-proxy store -> read by the tensor core
                }
            }
            tmem_st_x32(tmem + ((uint32_t)(q * 32) << 16) + A_COL + as * 64 + hf * 32, r);
            asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory");
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(&bar_a_full[as]);
            if (prb) probe[kb * 8 + 4] = clock64();
            if (++s == stages) { s = 0; ph ^= 1; }
            if (++as == kAStages) { as = 0; aph ^= 1; }
        }
        // ---- epilogue, first half: wait for the accumulators; k-splits publish their partial sums
        if constexpr (!PRO) pdl_wait();          // y may still be read by the previous kernel
        mbar_wait(&bar_d_full, 0);
        tc_fence_after();
        if (S > 1 && args.cluster) {
            // thread-block cluster of the S k-splits of this tile: ranks 1..S-1 store their accumulators into the leader's
            // shared memory (distributed shared memory), one cluster barrier, the leader adds them in rank order
            const uint32_t crank = cluster_ctarank();
            finish = crank == 0;
            if (crank != 0) {
                const uint32_t red_local = smem_u32(smem + args.red_off);
#pragma unroll
                for (int c0 = 0; c0 < NT / 2; c0 += 8) {
                    const int tcol = hf * (NT / 2) + c0;
                    if (tcol >= args.m) continue;
                    uint32_t acc[8];
                    tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + D_COL + tcol, acc);
#pragma unroll
                    for (int j = 0; j < 8; ++j) {
                        const int tok = tcol + j;
                        if (tok < args.m)
                            st_cluster_f32(red_local + (uint32_t)((((crank - 1) * args.m + tok) * kTileM + row) * 4), 0, __uint_as_float(acc[j]));
                    }
                }
            }
        } else if (S > 1) {
            // publish this k-split's partial accumulators, take a ticket; only the last arriver of the tile goes on
#pragma unroll
            for (int c0 = 0; c0 < NT / 2; c0 += 8) {
                const int tcol = hf * (NT / 2) + c0;
                if (tcol >= args.m) continue;
                uint32_t acc[8];
                tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + D_COL + tcol, acc);
#pragma unroll
                for (int j = 0; j < 8; ++j) {
                    const int tok = tcol + j;
                    if (col_ok && tok < args.m) args.part[((size_t)blockIdx.z * args.m + tok) * args.n + col] = num_kb > 0 ? __uint_as_float(acc[j]) : 0.f;
                }
            }
            __threadfence();
            asm volatile("bar.sync 1, 256;" ::: "memory");          // the eight converter / epilogue warps
            if (ct == 0) {
                const int old = atomicAdd(&args.tickets[blockIdx.x], 1);
                s_last = old == S - 1;
                if (old == S - 1) args.tickets[blockIdx.x] = 0;      // self-resetting for the next launch on this stream
            }
            asm volatile("bar.sync 1, 256;" ::: "memory");
            finish = s_last != 0;
            if (finish) __threadfence();
        }
    }
    if (S > 1 && args.cluster) cluster_sync_all();          // every thread of every CTA of the cluster
    if (warp >= 2 && finish) {
        // ---- epilogue, second half (one CTA per tile): scale, bias, activation, store -- or push to every rank
        const bool clustered = S > 1 && args.cluster;
        const float* red = reinterpret_cast<const float*>(smem + args.red_off);
        float sc = 1.f, bs = 0.f;
        if (col_ok) sc = __half2float(args.scale[col]);
        if (args.bias != nullptr && col_ok) bs = __half2float(args.bias[col]);
        const bool push_on = args.push.tp > 1;
        size_t push_word = 0;
        unsigned push_epoch = 0;
        if (push_on) {
            const TpIndex ix = tp_index(args.push, args.push_layer);
            push_word = tp_word_offset(args.push, ix.slot, args.push_kind, args.push.rank);
            push_epoch = ix.epoch;
        }
#pragma unroll
        for (int c0 = 0; c0 < NT / 2; c0 += 8) {
            const int tcol = hf * (NT / 2) + c0;
            if (tcol >= args.m) continue;
            uint32_t acc[8];
            if (S == 1 || clustered) tmem_ld_x8(tmem + ((uint32_t)(q * 32) << 16) + D_COL + tcol, acc);
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const int tok = tcol + j;
                if (tok >= args.m) continue;                 // (warp-uniform; col_ok is per lane: the push pairs lanes by shuffle)
                float v = 0.f;
                if (col_ok) {
                if (S == 1) {
                    v = __uint_as_float(acc[j]);
                } else if (clustered) {
                    v = __uint_as_float(acc[j]);
                    for (int z = 1; z < S; ++z) v += red[((z - 1) * args.m + tok) * kTileM + row];
                } else {
                    v = __ldcg(&args.part[(size_t)tok * args.n + col]);
                    for (int z = 1; z < S; ++z) v += __ldcg(&args.part[((size_t)z * args.m + tok) * args.n + col]);
                }
                v = v * sc + bs;
                if (args.act == 1) v = gelu_tanh_f32(v);
                }
                const __half hv = __float2half_rn(v);
                if (push_on) {
                    // one-shot exchange: the tile goes straight from the accumulators into every rank's memory over NVLink as
                    // flagged words {two adjacent columns, epoch} -- one 8-byte store each, no fence, no separate flag
                    const uint32_t mine = (uint32_t)__half_as_ushort(hv);
                    const uint32_t other = __shfl_down_sync(0xffffffffu, mine, 1);
                    if ((lane & 1) == 0 && col_ok) {
                        const size_t word = push_word + (size_t)tok * (args.push.h >> 1) + (col >> 1);
                        for (int r = 0; r < args.push.tp; ++r) tp_store_word(args.push.peer_data[r], word, mine | (other << 16), push_epoch);
                    }
                } else if (col_ok) {
                    args.y[(size_t)tok * args.ldy + col] = hv;
                }
            }
        }
    }
    if (warp >= 2) {
        tc_fence_before();
        if (ct == 0) trc_emit(TRC_GEMM_W8, trc_t0, trc_t1, trc_t2, args.n, args.k, trc_pro_ns);
    }
    __syncthreads();
    if (warp == 1) {
        tc_fence_after();
        asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem), "r"(kTmemCols) : "memory");
    }
}

}  // namespace dg

FTCF_TRACE_INSTALLER(trace_install_gemm_decode)

int decode_probe_install(long long* dev_buf)
{
    FTCF_CUDA_CHECK(cudaMemcpyToSymbol(dg::g_probe, &dev_buf, sizeof(dev_buf)));
    return FTCF_OK;
}

bool gemm_decode_supported(int m, int n, int k)
{
    return m >= 1 && m <= 32 && n >= 1 && k >= dg::BK && k % dg::BK == 0;
}

template <int NT, bool PRO, int MINB>
static int launch_decode_v(const CUtensorMap& mw, const CUtensorMap& mx, dg::Args a, dim3 grid, size_t smem, cudaStream_t st, bool pdl)
{
    auto kern = dg::gemm_decode_kernel<NT, PRO, MINB>;
    static std::atomic<size_t> configured{0};      // per instantiation; the attribute is per function (and per device context)
    if (configured.load(std::memory_order_relaxed) < smem) {
        FTCF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
        FTCF_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributePreferredSharedMemoryCarveout, (int)cudaSharedmemCarveoutMaxShared));
        configured.store(smem, std::memory_order_relaxed);
    }
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = grid;
    cfg.blockDim = dim3(dg::kThreads);
    cfg.dynamicSmemBytes = smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[2];
    unsigned na = 0;
    if (pdl && g_pdl_enabled.load(std::memory_order_relaxed)) {
        attr[na].id = cudaLaunchAttributeProgrammaticStreamSerialization;
        attr[na].val.programmaticStreamSerializationAllowed = 1;
        ++na;
    }
    if (a.cluster) {     // the k-splits of a tile are one thread-block cluster (distributed-shared-memory reduction)
        attr[na].id = cudaLaunchAttributeClusterDimension;
        attr[na].val.clusterDim.x = 1;
        attr[na].val.clusterDim.y = 1;
        attr[na].val.clusterDim.z = grid.z;
        ++na;
    }
    cfg.attrs = attr;
    cfg.numAttrs = na;
    const cudaError_t err = cudaLaunchKernelEx(&cfg, kern, mw, mx, a);
    FTCF_REQUIRE(err == cudaSuccess, FTCF_ERR_CUDA, "decode gemm launch failed: %s", cudaGetErrorString(err));
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

template <int NT, bool PRO>
static int launch_decode(const CUtensorMap& mw, const CUtensorMap& mx, dg::Args a, dim3 grid, size_t smem, cudaStream_t st, bool pdl)
{
    if (g_dg_lean.load(std::memory_order_relaxed) != 0) return launch_decode_v<NT, PRO, 3>(mw, mx, a, grid, smem, st, pdl);
    return launch_decode_v<NT, PRO, 2>(mw, mx, a, grid, smem, st, pdl);
}

// x == nullptr selects the fused-prologue variant (pro != nullptr, m <= 4)
int gemm_w8a16_decode(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n, int k, int act,
                      const SkPro* pro, cudaStream_t st, const ftcf_tp_exchange* push, int push_kind, int push_layer, const ftcf_launch_hint* hint)
{
    FTCF_REQUIRE(gemm_decode_supported(m, n, k), FTCF_ERR_UNSUPPORTED, "decode gemm: m=%d (1..32), k=%d (multiple of 128)", m, k);
    const int nt = m <= 16 ? 16 : 32;
    const int tiles = ceil_div(n, dg::kTileM), kb_all = k / dg::BK;
    // k-splits: fill the co-resident CTA slots (2 per SM), but leave every split enough K steps to amortise its pipeline fill
    int target = g_dg_target_ctas.load(std::memory_order_relaxed);
    if (pro != nullptr && pro->cta_hint > 0) target = pro->cta_hint;
    if (hint != nullptr && hint->target_ctas > 0) target = hint->target_ctas;
    const bool pdl = hint == nullptr || hint->no_pdl == 0;
    // (rounded to nearest: 80 tiles against a target of 148 take two splits, not one)
    int S = std::max(1, std::min((target + tiles / 2) / tiles, kb_all / std::max(1, g_dg_min_kb.load(std::memory_order_relaxed))));
    S = std::min(S, 16);
    // never more CTAs than co-resident slots (2 per SM): a second wave of a few CTAs doubles the launch's duration
    // (n = 20480 at target 296: 320 CTAs ran 36.9 us under ncu, profiles/r2_gemm_decode_ncu_full.txt)
    while (S > 1 && tiles * S > 2 * 148) --S;
    dg::Args a{};
    a.scale = static_cast<const __half*>(scale);
    a.bias = static_cast<const __half*>(bias);
    a.y = static_cast<__half*>(y);
    a.m = m; a.n = n; a.k = k; a.ldy = n; a.act = act;
    a.evict_first = g_dg_evict_first.load(std::memory_order_relaxed);
    if (push != nullptr) {
        FTCF_REQUIRE(push->tp > 1 && push->tp <= 8 && push->rank >= 0 && push->rank < push->tp && m <= push->m_max && n == push->h &&
                         (push_kind == 0 || push_kind == 1) && push->step != nullptr,
                     FTCF_ERR_INVALID, "decode gemm: bad tensor-parallel push (tp %d rank %d m %d/%d n %d/%d kind %d)", push->tp, push->rank, m,
                     push->m_max, n, push->h, push_kind);
        a.push = *push;
        a.push_kind = push_kind;
        a.push_layer = push_layer;
    }
    if (S > 1 && !splitk_scratch_acquire(st, (size_t)S * m * n, tiles, &a.part, &a.tickets)) S = 1;
    const int kb_per = ceil_div(kb_all, S);
    S = ceil_div(kb_all, kb_per);                       // no empty splits
    // shared memory: 2 CTAs per SM -> 228 KB / 2 minus the 1 KB the system reserves per CTA and the static barriers
    const size_t budget = 112 * 1024;
    const size_t x_bytes = (size_t)nt * 128 * 2;
    size_t fixed = 1024, per_stage = dg::W_BYTES;
    if (pro != nullptr) {
        FTCF_REQUIRE(m <= 4, FTCF_ERR_UNSUPPORTED, "decode gemm: the fused LayerNorm prologue takes m <= 4 rows (m=%d)", m);
        FTCF_REQUIRE(pro->x && pro->gamma && pro->beta && (pro->add_ffn == nullptr) == (pro->add_attn == nullptr), FTCF_ERR_INVALID,
                     "decode gemm: incomplete prologue");
        fixed += dg::kAStages * x_bytes + (size_t)m * (k + 8) * sizeof(__half);
        a.pro = *pro;
    } else {
        per_stage += x_bytes;
    }
    // k-splits reduce through distributed shared memory when the leader's buffer is small (decode rows); otherwise through
    // global partials + ticket
    const size_t red_bytes = (size_t)(S - 1) * m * dg::kTileM * sizeof(float);
    const bool cluster = S > 1 && S <= 8 && red_bytes <= 16 * 1024 && g_dg_cluster.load(std::memory_order_relaxed) != 0;
    const size_t before_red = fixed - 1024;             // bytes behind the ring, relative to the 1024-byte aligned base
    if (cluster) fixed += red_bytes;
    int fit = fixed < budget ? (int)((budget - fixed) / per_stage) : 0;
    const int want_stages = hint != nullptr ? hint->stages : 0;
    if (fit < 3 || want_stages > fit) fit = (int)((220 * 1024 - fixed) / per_stage);      // one CTA per SM: the prologue rows crowd the
                                                                                          // ring out, or the caller asked for a deeper ring
    FTCF_REQUIRE(fit >= 2, FTCF_ERR_UNSUPPORTED, "decode gemm: m=%d k=%d does not fit shared memory", m, k);
    int cap = std::min(dg::kMaxStages, std::max(2, g_dg_max_stages.load(std::memory_order_relaxed)));
    if (want_stages > 0) cap = std::min(dg::kMaxStages, std::max(2, want_stages));
    const int stages = std::min(std::min(fit, cap), std::max(kb_per, 2));
    a.stages = stages;
    const size_t smem = fixed + (size_t)stages * per_stage;
    if (cluster) {
        a.cluster = 1;
        a.red_off = (uint32_t)((size_t)stages * per_stage + before_red);
    }
    CUtensorMap mw, mx;
    int rc = make_tensor_map_2d(&mw, w_nk, n, k, 1, dg::kTileM);
    if (rc != FTCF_OK) return rc;
    if (pro == nullptr) {
        rc = make_tensor_map_2d(&mx, x, m, k, 2, nt);
        if (rc != FTCF_OK) return rc;
    } else {
        mx = mw;   // unused
    }
    const dim3 grid(tiles, 1, S);
    if (pro != nullptr) return launch_decode<16, true>(mw, mx, a, grid, smem, st, pdl);
    if (nt == 16) return launch_decode<16, false>(mw, mx, a, grid, smem, st, pdl);
    return launch_decode<32, false>(mw, mx, a, grid, smem, st, pdl);
}

}  // namespace ftcf
