// Persistent decode-step kernel: one launch runs every layer of one generated token (plus the final LayerNorm and the LM
// head) for a small batch (B <= 8) on one GPU.
//
// What it stands in for, per token, in the reference (paths relative to src/fastertransformer): the whole of
// GptNeoXDecoder<T>::forward (models/gptneox/GptNeoXDecoder.cc:197-389, parallel-residual branch :301-360) -- LayerNorm
// (kernels/layernorm_kernels.cu:158-286) x2, the four fpA_intB GEMMs (kernels/cutlass_kernels/fpA_intB_gemm/
// fpA_intB_gemm_template.h:461-570), masked_multihead_attention_kernel (kernels/decoder_masked_multihead_attention/
// decoder_masked_multihead_attention_template.hpp:1099-1919), invokeAddBiasAttentionFfnResidual
// (kernels/add_residual_kernels.cu:116-176) -- ~8 launches per layer there, and the embedding lookup, final LayerNorm and
// LM-head GEMM of GptNeoX<T>::forward (models/gptneox/GptNeoX.cc:776-912).
//
// Why one kernel (roofline: HBM; a token has to stream 12.6 GB of INT8 weights + 1 GB of LM head + the KV cache): with one
// launch per operator the weight stream stops at every kernel boundary (drain, launch gap, ramp-up: ~8 us x 330 per token).
// Here every CTA (one per SM) has a PRODUCER warp that walks the complete, statically known schedule of the step -- all
// weight tiles and KV-cache tiles this CTA will need, layer after layer -- and keeps a ~200 KB shared-memory ring full with
// TMA bulk copies (cp.async.bulk + mbarrier complete_tx).  Weights and old KV rows are constants of the step, so the producer
// never waits for a data dependency: HBM keeps streaming while the consumer warps sit in a grid-wide barrier or normalise
// activations.  The eight CONSUMER warps follow the same schedule:
//   phase A  LayerNorm(x) (each CTA redundantly, into shared memory) -> QKV and FFN1 (+bias, tanh-GELU) row tiles
//   phase B  attention over the cache (split-KV work units, last arriver merges) and FFN2 as split-K partial tiles
//   phase C  O-projection row tiles with the residual add fused into their epilogue (x <- x + attn + ffn + bias)
// with one grid barrier (release/acquire counter in global memory) after each phase.  All work units are ~80 KB (16 weight
// rows x 5120 bytes, or 160 keys of K+V), handed out in contiguous, per-phase-rotated ranges, so every SM pulls the same
// number of bytes per layer; a +-1 unit skew inside a phase is absorbed by the ring (an early CTA prefetches the next phase).
//   * weight tile = 16 output features x 1024 bytes of k per stage, rows padded to 1040 bytes in shared memory so that the
//     LDS.128 fragment reads are bank-conflict free; u8 -> fp16 in registers (PRMT + HSUB2, exact), mma.sync.m16n8k16 with
//     the weights as the 16-row operand and the <= 8 tokens as the 8-column operand, the 8 warps split k, fixed-order
//     shared-memory reduction -> deterministic; dequant scale / bias / GELU / residual in the epilogue.
//   * KV tile = 64 keys x dh fp16 per stage; each warp owns 8 keys of a tile and keeps a private online-softmax state, so
//     there is no CTA-wide synchronisation inside a unit.
#include <algorithm>
#include <type_traits>

#include "decode_mega.cuh"
#include "tma_utils.cuh"

namespace ftcf {
namespace mg {

using namespace tma;

// ------------------------------------------------------------------------------------------------ small device helpers
// streamed-once data (weights, KV): L2 evict_first, so that the activations / barrier words the consumers re-read stay resident
__device__ __forceinline__ uint64_t l2_evict_first_policy()
{
    uint64_t pol;
    asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    return pol;
}
__device__ __forceinline__ void bulk_g2s(void* dst, const void* src, uint32_t bytes, uint64_t* bar, uint64_t pol)
{
    if (pol == 0) {
        asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(smem_u32(dst)),
                     "l"(src), "r"(bytes), "r"(smem_u32(bar))
                     : "memory");
        return;
    }
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], [%1], %2, [%3], %4;" ::"r"(
                     smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(smem_u32(bar)), "l"(pol)
                 : "memory");
}
__device__ __forceinline__ void cbar() { asm volatile("bar.sync 1, 256;" ::: "memory"); }   // consumer warps only
__device__ __forceinline__ uint4 ld_cg_16(const void* p)
{
    uint4 r;
    asm volatile("ld.global.cg.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p));
    return r;
}
__device__ __forceinline__ float ld_cg_f32(const float* p)
{
    float r;
    asm volatile("ld.global.cg.f32 %0, [%1];" : "=f"(r) : "l"(p));
    return r;
}
// raw 16-bit loads whose ISSUE point is pinned (asm volatile) and whose VALUE is not touched until the caller converts it:
// used to request epilogue operands a whole unit ahead of their use
__device__ __forceinline__ unsigned short ld_nc_u16_raw(const void* p)
{
    unsigned short r;
    asm volatile("ld.global.nc.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ unsigned short ld_cg_u16_raw(const void* p)
{
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return r;
}
__device__ __forceinline__ __half ld_cg_h(const __half* p)
{
    unsigned short r;
    asm volatile("ld.global.cg.u16 %0, [%1];" : "=h"(r) : "l"(p));
    return __ushort_as_half(r);
}
__device__ __forceinline__ void mma_16816(float (&d)[4], uint32_t a0, uint32_t a1, uint32_t a2, uint32_t a3, uint32_t b0, uint32_t b1)
{
    asm volatile("mma.sync.aligned.m16n8k16.row.col.f32.f16.f16.f32 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, {%0,%1,%2,%3};"
                 : "+f"(d[0]), "+f"(d[1]), "+f"(d[2]), "+f"(d[3])
                 : "r"(a0), "r"(a1), "r"(a2), "r"(a3), "r"(b0), "r"(b1));
}
__device__ __forceinline__ void u8x4_to_h2x2(uint32_t w, uint32_t& lo, uint32_t& hi)
{
    lo = __byte_perm(w, 0x64646464u, 0x4140);
    hi = __byte_perm(w, 0x64646464u, 0x4342);
    const uint32_t magic = 0x64806480u;   // 1152 = 1024 + 128: (1024 + b) - 1152 = b - 128, exact in fp16
    asm("sub.f16x2 %0, %1, %2;" : "=r"(lo) : "r"(lo), "r"(magic));
    asm("sub.f16x2 %0, %1, %2;" : "=r"(hi) : "r"(hi), "r"(magic));
}
__device__ __forceinline__ uint32_t u4_get(const uint4& v, int i) { return i == 0 ? v.x : (i == 1 ? v.y : (i == 2 ? v.z : v.w)); }
// 16-byte chunk j of an operand row lives at chunk swz(j): lanes t and t+2 of a fragment read would otherwise hit the same banks
__device__ __forceinline__ int swz(int j) { return j ^ (((j >> 3) & 1) << 1); }

__device__ __forceinline__ float rotary_angle(int pos, int i, int rot)
{
    return (float)pos / powf(10000.f, (2.f * (float)i) / (float)rot);   // decoder_masked_multihead_attention_utils.h:1325-1329
}
__device__ __forceinline__ __half rotary_neox(__half xd, __half xpartner, int d, int rot, int pos)
{
    const int half_rot = rot >> 1;
    const int i = d < half_rot ? d : d - half_rot;
    float sn, cs;
    sincosf(rotary_angle(pos, i, rot), &sn, &cs);
    const float a = __half2float(xd), b = __half2float(xpartner);
    return __float2half_rn(d < half_rot ? cs * a - sn * b : cs * a + sn * b);
}

// contiguous share [u0, u1) of N units for this CTA; `rot` moves the CTAs that get the remainder from phase to phase
__device__ __forceinline__ void my_range(int N, int rot, int& u0, int& u1)
{
    const int G = gridDim.x;
    const int c = (int)((blockIdx.x + (unsigned)rot) % (unsigned)G);
    u0 = (int)(((long long)N * c) / G);
    u1 = (int)(((long long)N * (c + 1)) / G);
}

struct RingPos {
    int slot;
    uint32_t par;
    __device__ __forceinline__ void next(int ns)
    {
        if (++slot == ns) {
            slot = 0;
            par ^= 1;
        }
    }
};
// producer only: before issuing stage i, stage i - inflight must have LANDED.  The ring still buffers `ns` stages of skew, but
// no more than `inflight` copies per SM queue up in the memory system -- with 148 SMs x 12 x 16 KB outstanding every dependent
// L2 access of the consumers (grid barrier, operand staging) waited microseconds behind the prefetch stream.
struct Smem;
struct Throttle {
    RingPos tail;      // oldest stage that may still be in flight
    int ahead;         // stages issued and not yet known to have landed
    uint64_t pol;
    long long c_empty, c_thr;   // dbg & 64 accounting
};
__device__ __forceinline__ void throttle_before_issue(Throttle& th, Smem& sm, const Params& p);

// per-sequence attention bookkeeping, identical in every CTA (derived from device-resident request state)
struct AttInfo {
    int tlen[MAX_B];     // cache slot of the new token == number of earlier slots
    int inl[MAX_B];      // valid prompt keys: min(input_len, tlen)
    int nvalid[MAX_B];   // cached keys attended (pad gap [input_len, max_in) excluded)
    int nu[MAX_B];       // work units per (sequence, head); 0 for a finished sequence
    int pre[MAX_B + 1];  // unit offsets (x Hl) per sequence
    int pos[MAX_B];      // rotary position of the new token
    int tok[MAX_B];      // previous token id (embedding row)
};

struct Smem {
    uint64_t full[16], empty[16];
    AttInfo att;
    const __half* xrow[MAX_B];
    __half2 mean_h[MAX_B], rstd_h[MAX_B];
    float red[2][CW][8][RED_PITCH];
    float att_red[CW][128 + 2];
    __align__(16) __half q[128];
    __align__(16) __half knew[128];
    __align__(16) __half vnew[128];
    const __half* src[MAX_B];
    long long c_full[CW];
    float s_new;
    int flag;
};

// ------------------------------------------------------------------------------------------------ producer
// one weight unit = consecutive 16-row blocks of the tiled layout: one contiguous bulk copy per stage
__device__ __forceinline__ void throttle_before_issue(Throttle& th, Smem& sm, const Params& p)
{
    if (th.ahead >= p.inflight) {
        const long long t0 = clock64();
        mbar_wait(&sm.full[th.tail.slot], th.tail.par);   // completes when that stage's bytes have arrived
        th.c_thr += clock64() - t0;
        th.tail.next(p.ns);
        --th.ahead;
    }
    ++th.ahead;
}

__device__ __forceinline__ void produce_weight_unit(const Params& p, uint8_t* ring, Smem& sm, RingPos& rp, Throttle& th, const uint8_t* src,
                                                    int kbytes, int lane)
{
    for (int k0 = 0; k0 < kbytes; k0 += STAGE_K) {
        const uint32_t bytes = (uint32_t)(ROWS * min(STAGE_K, kbytes - k0));
        if (lane == 0) {
            throttle_before_issue(th, sm, p);
            const long long t0 = clock64();
            mbar_wait(&sm.empty[rp.slot], rp.par ^ 1);
            th.c_empty += clock64() - t0;
            mbar_arrive_expect_tx(&sm.full[rp.slot], bytes);
            bulk_g2s(ring + (size_t)rp.slot * STAGE_BYTES, src + (size_t)k0 * ROWS, bytes, &sm.full[rp.slot], th.pol);
        }
        rp.next(p.ns);
    }
    __syncwarp();
}

// rows [tv, tv + nk) of the virtual (gap-free) key index space of (b, head) -> one stage
__device__ __forceinline__ void produce_kv_tile(const Params& p, uint8_t* ring, Smem& sm, RingPos& rp, Throttle& th, const __half* cache_bh,
                                                int b, int tv, int nk, int lane)
{
    const int rowb = p.dh * 2;
    const int inl = sm.att.inl[b];
    if (lane == 0) {
        throttle_before_issue(th, sm, p);
        mbar_wait(&sm.empty[rp.slot], rp.par ^ 1);
        mbar_arrive_expect_tx(&sm.full[rp.slot], (uint32_t)(nk * rowb));
        uint8_t* dst = ring + (size_t)rp.slot * STAGE_BYTES;
        int n1 = 0;
        if (tv < inl) {
            n1 = min(nk, inl - tv);
            bulk_g2s(dst, cache_bh + (size_t)tv * p.dh, (uint32_t)(n1 * rowb), &sm.full[rp.slot], th.pol);
        }
        if (n1 < nk) {
            const int pos = (tv + n1) - inl + p.max_in;   // past the pad gap [input_len, max_in)
            bulk_g2s(dst + (size_t)n1 * rowb, cache_bh + (size_t)pos * p.dh, (uint32_t)((nk - n1) * rowb), &sm.full[rp.slot], th.pol);
        }
    }
    __syncwarp();
    rp.next(p.ns);
}

// ------------------------------------------------------------------------------------------------ consumer: GEMM units
// register image of one ring stage for one warp: its 128-byte k-step of the 16 weight rows and the matching activations
template <typename WT>
struct Frag {
    uint4 x[16 / (int)sizeof(WT) / 4];
    uint4 w[2][2];
};

template <typename WT>
__device__ __forceinline__ void frag_load_x(Frag<WT>& f, const uint8_t* opnd_row, int kelem0, int warp, int t)
{
    constexpr int EPC = 16 / (int)sizeof(WT);   // k elements per 16-byte weight chunk
    constexpr int KSTEP = 8 * EPC;              // k elements per 128-byte k-step
    const int j0 = (kelem0 + warp * KSTEP + t * 2 * EPC) >> 3;
#pragma unroll
    for (int q = 0; q < EPC / 4; ++q) f.x[q] = *reinterpret_cast<const uint4*>(opnd_row + (swz(j0 + q) << 4));
}
template <typename WT>
__device__ __forceinline__ void frag_load_w(Frag<WT>& f, const uint8_t* slot, int bytes, int warp, int g, int t)
{
    // block image: row r at r * bytes, 16-byte chunk c at chunk c ^ (r & 7)   ((g + 8) & 7 == g & 7)
    const uint8_t* st = slot + g * bytes + warp * 128;
#pragma unroll
    for (int hh = 0; hh < 2; ++hh)
#pragma unroll
        for (int c = 0; c < 2; ++c) f.w[hh][c] = *reinterpret_cast<const uint4*>(st + hh * 8 * bytes + (((2 * t + c) ^ (g & 7)) << 4));
}
// two independent accumulator chains (one per 16-byte chunk) so that consecutive MMAs do not wait for each other
template <typename WT>
__device__ __forceinline__ void frag_mma(const Frag<WT>& f, float (&acc0)[4], float (&acc1)[4])
{
    constexpr int EPC = 16 / (int)sizeof(WT);
    constexpr int NMMA = EPC / 4;
#pragma unroll
    for (int j = 0; j < NMMA; ++j)
#pragma unroll
        for (int c = 0; c < 2; ++c) {
            const int pi = c * (EPC / 2) + 2 * j;
            uint32_t a0, a1, a2, a3;
            if constexpr (sizeof(WT) == 1) {
                u8x4_to_h2x2(u4_get(f.w[0][c], j), a0, a2);
                u8x4_to_h2x2(u4_get(f.w[1][c], j), a1, a3);
            } else {
                a0 = u4_get(f.w[0][c], 2 * j);
                a2 = u4_get(f.w[0][c], 2 * j + 1);
                a1 = u4_get(f.w[1][c], 2 * j);
                a3 = u4_get(f.w[1][c], 2 * j + 1);
            }
            if (c == 0) mma_16816(acc0, a0, a1, a2, a3, u4_get(f.x[pi / 4], pi % 4), u4_get(f.x[(pi + 1) / 4], (pi + 1) % 4));
            else mma_16816(acc1, a0, a1, a2, a3, u4_get(f.x[pi / 4], pi % 4), u4_get(f.x[(pi + 1) / 4], (pi + 1) % 4));
        }
}

enum { EP_QKV = 0, EP_FFN1 = 1, EP_FFN2 = 2, EP_O = 3, EP_LM = 4 };

// One weight unit: 16 rows [row0, row0 + 16) x kbytes of k (operand elements from 0); leaves the 16 x B results (summed over
// the 8 k-splitting warps, fixed order) with threads 0..127 and runs the epilogue `ep` there.  The stage loop is software
// pipelined (the next stage's barrier wait and shared-memory loads are issued before the current stage's MMAs) and the
// epilogue's global operands are requested before the loop, so neither latency sits on the unit's critical path.
template <typename WT, bool W8>
__device__ __forceinline__ void gemm_unit(const Params& p, const LayerDev* Ld, uint8_t* ring, Smem& sm, RingPos& rp, int& redbuf,
                                          const uint8_t* opnd, int kbytes, int row0, int nrows_total, int ep, int kc, int warp,
                                          int lane)
{
    const int tid = warp * 32 + lane;
    const int g = lane >> 2, t = lane & 3;
    const uint8_t* opnd_row = opnd + (size_t)min(g, p.B - 1) * p.opnd_pitch;

    // ---- epilogue operands, requested early
    const int f = tid & 15, tok = tid >> 4;
    const int col = row0 + f;
    const bool ep_thread = tid < 128 && tok < p.B && col < nrows_total;
    // (raw bits only here: touching a value would stall this warp for a DRAM latency at the START of every unit and, through
    // the unit-end barrier, all the others with it)
    constexpr int MAX_KS_PF = 4;
    unsigned short raw_scale = 0x3c00 /* 1.0 */, raw_b = 0, raw_xs = 0;
    float raw_ffn[MAX_KS_PF] = {0.f, 0.f, 0.f, 0.f};
    if (ep_thread && ep != EP_LM) {
        const int kind = ep == EP_QKV ? 0 : (ep == EP_O ? 1 : (ep == EP_FFN1 ? 2 : 3));
        if constexpr (W8) raw_scale = ld_nc_u16_raw(Ld->scale[kind] + col);
        if (ep == EP_FFN1) raw_b = ld_nc_u16_raw(Ld->ffn1_b + col);
        if (ep == EP_O) {
            raw_b = ld_nc_u16_raw(Ld->res_b + col);
            raw_xs = ld_cg_u16_raw(sm.xrow[tok] + col);
            if (p.ks <= MAX_KS_PF) {
#pragma unroll
                for (int c2 = 0; c2 < MAX_KS_PF; ++c2)
                    if (c2 < p.ks) raw_ffn[c2] = ld_cg_f32(p.ffn_part + ((size_t)c2 * p.B + tok) * p.h + col);
            }
        }
    }

    // ---- pipelined stage loop
    float acc[4] = {0.f, 0.f, 0.f, 0.f}, acc1[4] = {0.f, 0.f, 0.f, 0.f};
    const int nst = (kbytes + STAGE_K - 1) / STAGE_K;
    auto load_stage = [&](Frag<WT>& fr, int s) -> bool {
        const int bytes = min(STAGE_K, kbytes - s * STAGE_K);
        const bool active = warp * 128 < bytes;
        if (active && !(p.dbg & 16)) frag_load_x<WT>(fr, opnd_row, s * (STAGE_K / (int)sizeof(WT)), warp, t);
        if (p.dbg & 64) {
            const long long t0 = clock64();
            mbar_wait(&sm.full[rp.slot], rp.par);
            sm.c_full[warp] += clock64() - t0;
        } else {
            mbar_wait(&sm.full[rp.slot], rp.par);
        }
        if (active && !(p.dbg & 16)) frag_load_w<WT>(fr, ring + (size_t)rp.slot * STAGE_BYTES, bytes, warp, g, t);
        __syncwarp();
        if (lane == 0) mbar_arrive(&sm.empty[rp.slot]);
        rp.next(p.ns);
        return active;
    };
    Frag<WT> f0, f1;
    if (p.dbg & 17) {
#pragma unroll
        for (int q = 0; q < 16 / (int)sizeof(WT) / 4; ++q) f0.x[q] = f1.x[q] = make_uint4(0, 0, 0, 0);
#pragma unroll
        for (int q = 0; q < 4; ++q) f0.w[q >> 1][q & 1] = f1.w[q >> 1][q & 1] = make_uint4(0, 0, 0, 0);
    }
    const bool do_mma = !(p.dbg & 1);
    bool a0 = load_stage(f0, 0), a1 = false;
    for (int s = 0; s < nst; s += 2) {
        a1 = (s + 1 < nst) ? load_stage(f1, s + 1) : false;
        if (a0 && do_mma) frag_mma<WT>(f0, acc, acc1);
        a0 = (s + 2 < nst) ? load_stage(f0, s + 2) : false;
        if (a1 && do_mma) frag_mma<WT>(f1, acc, acc1);
    }
    if (p.dbg & 8) {
        if (acc[0] + acc1[0] + __uint_as_float(f0.w[0][0].x ^ f1.w[0][0].x) == 123.456f) p.logits[0] = 1.f;   // keep the loads alive
        return;
    }
#pragma unroll
    for (int i = 0; i < 4; ++i) acc[i] += acc1[i];

    float(*red)[8][RED_PITCH] = sm.red[redbuf];
    red[warp][2 * t][g] = acc[0];
    red[warp][2 * t + 1][g] = acc[1];
    red[warp][2 * t][g + 8] = acc[2];
    red[warp][2 * t + 1][g + 8] = acc[3];
    cbar();
    redbuf ^= 1;
    if (!ep_thread) return;
    float v = red[0][tok][f];
#pragma unroll
    for (int w = 1; w < CW; ++w) v += red[w][tok][f];
    if (ep == EP_LM) {
        p.logits[(size_t)tok * p.ld_logits + col] = v;
        return;
    }
    const float pf_scale = __half2float(__ushort_as_half(raw_scale));
    const __half pf_b = __ushort_as_half(raw_b), pf_xs = __ushort_as_half(raw_xs);
    v *= pf_scale;
    if (ep == EP_QKV) {
        p.qkv[(size_t)tok * 3 * p.hl + col] = __float2half_rn(v);
    } else if (ep == EP_FFN1) {
        __half o;
        if constexpr (W8) {
            v += __half2float(pf_b);
            o = __float2half_rn(gelu_tanh_f32(v));
        } else {
            o = gelu_tanh_half_ref(__hadd(__float2half_rn(v), pf_b));
        }
        p.inter_buf[(size_t)tok * p.inter + col] = o;
    } else if (ep == EP_FFN2) {
        p.ffn_part[((size_t)kc * p.B + tok) * p.h + col] = v;
    } else {   // EP_O: x <- ((ffn + attn) + bias) + x / tp      (add_residual_kernels.cu:116-152, fp16 adds)
        const __half attn = __float2half_rn(v);
        float pf_ffn = 0.f;
        if (p.ks <= MAX_KS_PF) {
#pragma unroll
            for (int c2 = 0; c2 < MAX_KS_PF; ++c2) pf_ffn += raw_ffn[c2];   // unused slots are 0; fixed order
        } else {
            for (int c2 = 0; c2 < p.ks; ++c2) pf_ffn += ld_cg_f32(p.ffn_part + ((size_t)c2 * p.B + tok) * p.h + col);
        }
        __half r = __hadd(__float2half_rn(pf_ffn), attn);
        r = __hadd(r, pf_b);
        __half xs = pf_xs;
        if (p.tp > 1) xs = __float2half_rn(__half2float(xs) * (1.f / (float)p.tp));
        p.x[(size_t)tok * p.h + col] = __hadd(r, xs);
    }
}

// ------------------------------------------------------------------------------------------------ consumer: operand staging
// raw rows (B x n fp16, row b at src[b]) -> operand buffer, 16-byte chunks swizzled
__device__ __forceinline__ void stage_rows(const Params& p, uint8_t* opnd, const __half* const* src, size_t src_off, int n, int tid)
{
    const int nvec = n >> 3;
    for (int i = tid; i < p.B * nvec; i += CT) {
        const int b = i / nvec, vi = i - b * nvec;
        const uint4 v = ld_cg_16(src[b] + src_off + (size_t)vi * 8);
        *reinterpret_cast<uint4*>(opnd + (size_t)b * p.opnd_pitch + (swz(vi) << 4)) = v;
    }
}
// LayerNorm statistics of the raw rows already in the operand buffer (warp w takes rows w, w + 8, ...)
__device__ __forceinline__ void ln_stats(const Params& p, const uint8_t* opnd, Smem& sm, int warp, int lane)
{
    const int nvec = p.h >> 3;
    for (int b = warp; b < p.B; b += CW) {
        float s = 0.f, ss = 0.f;
        for (int vi = lane; vi < nvec; vi += 32) {
            const uint4 v = *reinterpret_cast<const uint4*>(opnd + (size_t)b * p.opnd_pitch + (swz(vi) << 4));
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
            const float mean = s / p.h;
            const float var = ss / p.h - mean * mean + p.eps;
            sm.mean_h[b] = __float2half2_rn(mean);
            sm.rstd_h[b] = __float2half2_rn(rsqrtf(var));
        }
    }
}
// x rows -> LayerNorm(x; gamma, beta) in the operand buffer (layernorm_kernels.cu:158-286: fp32 statistics, the
// normalisation in half2 with a rounding after every operation).  `have_stats`: the statistics of these rows are in sm.
__device__ __forceinline__ void stage_layernorm(const Params& p, uint8_t* opnd, Smem& sm, const __half* gamma, const __half* beta,
                                                bool have_stats, int warp, int lane)
{
    const int tid = warp * 32 + lane;
    if (p.dbg & 32) return;
    stage_rows(p, opnd, sm.xrow, 0, p.h, tid);
    cbar();
    if (!have_stats) {
        ln_stats(p, opnd, sm, warp, lane);
        cbar();
    }
    const int nvec = p.h >> 3;
    for (int i = tid; i < p.B * nvec; i += CT) {
        const int b = i / nvec, vi = i - b * nvec;
        uint4* slot = reinterpret_cast<uint4*>(opnd + (size_t)b * p.opnd_pitch + (swz(vi) << 4));
        uint4 v = *slot;
        const uint4 gq = __ldg(reinterpret_cast<const uint4*>(gamma + vi * 8));
        const uint4 bq = __ldg(reinterpret_cast<const uint4*>(beta + vi * 8));
        const __half2* gh = reinterpret_cast<const __half2*>(&gq);
        const __half2* bh = reinterpret_cast<const __half2*>(&bq);
        __half2* vh = reinterpret_cast<__half2*>(&v);
        const __half2 mean_h = sm.mean_h[b], rstd_h = sm.rstd_h[b];
#pragma unroll
        for (int j = 0; j < 4; ++j) vh[j] = __hfma2(__hmul2_rn(__hsub2_rn(vh[j], mean_h), rstd_h), gh[j], bh[j]);
        *slot = v;
    }
    cbar();
}

// ------------------------------------------------------------------------------------------------ grid barrier
constexpr int TS_BARS = 130, TS_CTAS = 160;
__device__ unsigned long long g_mega_ts[3][TS_BARS][TS_CTAS];
__device__ long long g_mega_cyc[8][TS_CTAS];   // dbg & 64: per CTA cycles [0] producer empty-wait [1] producer throttle-wait [2] producer total
                                               // [3] consumer warp 0 full-wait [4] consumer total [5] consumer cbar-wait [6] stages   // dbg & 64: globaltimer at [enter, arrive, release] of each barrier
__device__ __forceinline__ unsigned long long gtime()
{
    unsigned long long t;
    asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
    return t;
}
__device__ __forceinline__ void grid_barrier(unsigned* ctr, unsigned target, int tid, int dbg)
{
    if (dbg & 2) {
        cbar();
        return;
    }
    const bool ts = (dbg & 64) && tid == 0 && target / gridDim.x <= TS_BARS && blockIdx.x < TS_CTAS;
    const unsigned bi = target / gridDim.x - 1;
    if (ts) g_mega_ts[0][bi][blockIdx.x] = gtime();
    __threadfence();
    cbar();
    if (ts) g_mega_ts[1][bi][blockIdx.x] = gtime();
    if (tid == 0) {
        asm volatile("red.release.gpu.global.add.u32 [%0], 1;" ::"l"(ctr) : "memory");
        unsigned v = 0;
        long long spins = 0;
        do {
            asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(ctr) : "memory");
            if (v >= target) break;
            __nanosleep(32);
            if (++spins > (1ll << 26)) __trap();   // a scheduling bug must not hang the GPU
        } while (true);
        __threadfence();
    }
    if (ts) g_mega_ts[2][bi][blockIdx.x] = gtime();
    cbar();
}

// ------------------------------------------------------------------------------------------------ consumer: attention units
template <int DH>
__device__ __forceinline__ void attention_unit(const Params& p, const LayerDev* Ld, int layer, uint8_t* ring, Smem& sm, RingPos& rp, int b,
                                               int hh, int j, int warp, int lane)
{
    constexpr int LPR = DH / 8;          // lanes per cache row (16 bytes each)
    constexpr int KPW = 32 / LPR;        // keys per warp-wide load
    constexpr int WAVES = 8 / KPW;       // a warp owns 8 keys of a 64-key tile
    const int tid = warp * 32 + lane;
    const int H = p.Hl;
    const int nu = sm.att.nu[b], tlen = sm.att.tlen[b], nvalid = sm.att.nvalid[b];
    const bool owner = j == nu - 1;
    __half* kc = p.kv + (size_t)(2 * layer) * p.kv_layer_elems + ((size_t)b * H + hh) * (size_t)p.max_len * DH;
    __half* vc = p.kv + (size_t)(2 * layer + 1) * p.kv_layer_elems + ((size_t)b * H + hh) * (size_t)p.max_len * DH;

    // ---- q (every unit), new k / v (owner): bias, NeoX rotary, append to the cache (template.hpp:1204-1398)
    if (tid < DH) {
        const int d = tid;
        const __half* qkv = p.qkv + (size_t)b * 3 * H * DH;
        const __half* bias = Ld->qkv_b;
        const int rot = p.rot, pos = sm.att.pos[b];
        const int qi = hh * DH + d;
        const bool do_rot = d < rot;
        const int dp = d < (rot >> 1) ? d + (rot >> 1) : d - (rot >> 1);
        __half q = ld_cg_h(qkv + qi);
        if (bias) q = __hadd(q, bias[qi]);
        if (do_rot) {
            __half qp = ld_cg_h(qkv + hh * DH + dp);
            if (bias) qp = __hadd(qp, bias[hh * DH + dp]);
            q = rotary_neox(q, qp, d, rot, pos);
        }
        sm.q[d] = q;
        if (owner) {
            const int ki = H * DH + qi, vi = 2 * H * DH + qi;
            __half k = ld_cg_h(qkv + ki), v = ld_cg_h(qkv + vi);
            if (bias) {
                k = __hadd(k, bias[ki]);
                v = __hadd(v, bias[vi]);
            }
            if (do_rot) {
                __half kp = ld_cg_h(qkv + H * DH + hh * DH + dp);
                if (bias) kp = __hadd(kp, bias[H * DH + hh * DH + dp]);
                k = rotary_neox(k, kp, d, rot, pos);
            }
            sm.knew[d] = k;
            sm.vnew[d] = v;
            kc[(size_t)tlen * DH + d] = k;
            vc[(size_t)tlen * DH + d] = v;
        }
    }
    cbar();

    const int dl = lane % LPR, kq = lane / LPR;
    float q[8];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(&sm.q[dl * 8]);
        const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(qh[i]);
            q[2 * i] = f.x;
            q[2 * i + 1] = f.y;
        }
    }
    if (owner && warp == 0) {   // score of the new token itself (its k never comes from the cache)
        float dot = 0.f;
        if (kq == 0) {
            const uint4 kv = *reinterpret_cast<const uint4*>(&sm.knew[dl * 8]);
            const __half2* kh = reinterpret_cast<const __half2*>(&kv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(kh[i]);
                dot = fmaf(q[2 * i], f.x, dot);
                dot = fmaf(q[2 * i + 1], f.y, dot);
            }
        }
        dot = warp_sum(dot);
        if (lane == 0) sm.s_new = dot * p.inv_sqrt_dh;
    }

    // ---- warp-private online softmax over this warp's keys of every tile
    float m_run = -INFINITY, l_run = 0.f, o[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) o[i] = 0.f;
    const int v0 = j * ATT_UNIT, v1 = min(nvalid, v0 + ATT_UNIT);
    for (int tv = v0; tv < v1; tv += ATT_TILE) {
        const int nk = min(ATT_TILE, v1 - tv);
        float pr[WAVES];
        {   // K stage
            const uint8_t* st = ring + (size_t)rp.slot * STAGE_BYTES;
            mbar_wait(&sm.full[rp.slot], rp.par);
            uint4 kv[WAVES];
#pragma unroll
            for (int w = 0; w < WAVES; ++w) {
                const int key = warp * 8 + w * KPW + kq;
                kv[w] = make_uint4(0, 0, 0, 0);
                if (key < nk) kv[w] = *reinterpret_cast<const uint4*>(st + (size_t)key * (DH * 2) + dl * 16);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[rp.slot]);
            rp.next(p.ns);
            float mx = -INFINITY;
#pragma unroll
            for (int w = 0; w < WAVES; ++w) {
                const __half2* kh = reinterpret_cast<const __half2*>(&kv[w]);
                float dot = 0.f;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(kh[i]);
                    dot = fmaf(q[2 * i], f.x, dot);
                    dot = fmaf(q[2 * i + 1], f.y, dot);
                }
#pragma unroll
                for (int off = LPR / 2; off > 0; off >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, off);
                const int key = warp * 8 + w * KPW + kq;
                pr[w] = key < nk ? dot * p.inv_sqrt_dh : -INFINITY;
                mx = fmaxf(mx, pr[w]);
            }
#pragma unroll
            for (int off = LPR; off < 32; off <<= 1) mx = fmaxf(mx, __shfl_xor_sync(0xffffffffu, mx, off));
            const float m_new = fmaxf(m_run, mx);
            if (m_new > -INFINITY) {
                const float corr = m_run == -INFINITY ? 0.f : __expf(m_run - m_new);
                float ps = 0.f;
#pragma unroll
                for (int w = 0; w < WAVES; ++w) {
                    pr[w] = pr[w] == -INFINITY ? 0.f : __expf(pr[w] - m_new);
                    ps += pr[w];
                }
                l_run = l_run * corr + ps;
#pragma unroll
                for (int i = 0; i < 8; ++i) o[i] *= corr;
                m_run = m_new;
            } else {
#pragma unroll
                for (int w = 0; w < WAVES; ++w) pr[w] = 0.f;
            }
        }
        {   // V stage
            const uint8_t* st = ring + (size_t)rp.slot * STAGE_BYTES;
            mbar_wait(&sm.full[rp.slot], rp.par);
            uint4 vv[WAVES];
#pragma unroll
            for (int w = 0; w < WAVES; ++w) {
                const int key = warp * 8 + w * KPW + kq;
                vv[w] = make_uint4(0, 0, 0, 0);
                if (key < nk) vv[w] = *reinterpret_cast<const uint4*>(st + (size_t)key * (DH * 2) + dl * 16);
            }
            __syncwarp();
            if (lane == 0) mbar_arrive(&sm.empty[rp.slot]);
            rp.next(p.ns);
#pragma unroll
            for (int w = 0; w < WAVES; ++w) {
                const __half2* vh = reinterpret_cast<const __half2*>(&vv[w]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(vh[i]);
                    o[2 * i] = fmaf(pr[w], f.x, o[2 * i]);
                    o[2 * i + 1] = fmaf(pr[w], f.y, o[2 * i + 1]);
                }
            }
        }
    }
    // ---- merge the key groups of the warp, then the 8 warps (fixed order), then the new token, then the other units
#pragma unroll
    for (int off = LPR; off < 32; off <<= 1) {
        l_run += __shfl_xor_sync(0xffffffffu, l_run, off);
#pragma unroll
        for (int i = 0; i < 8; ++i) o[i] += __shfl_xor_sync(0xffffffffu, o[i], off);
    }
    if (kq == 0) {
#pragma unroll
        for (int i = 0; i < 8; ++i) sm.att_red[warp][dl * 8 + i] = o[i];
        if (dl == 0) {
            sm.att_red[warp][DH] = m_run;
            sm.att_red[warp][DH + 1] = l_run;
        }
    }
    cbar();
    float O = 0.f, L = 0.f, M = -INFINITY;
    if (tid < DH) {
#pragma unroll
        for (int w = 0; w < CW; ++w) M = fmaxf(M, sm.att_red[w][DH]);
        if (owner) M = fmaxf(M, sm.s_new);
#pragma unroll
        for (int w = 0; w < CW; ++w) {
            const float mw = sm.att_red[w][DH];
            const float e = mw == -INFINITY ? 0.f : __expf(mw - M);
            O = fmaf(sm.att_red[w][tid], e, O);
            L = fmaf(sm.att_red[w][DH + 1], e, L);
        }
        if (owner) {
            const float e = __expf(sm.s_new - M);
            O = fmaf(__half2float(sm.vnew[tid]), e, O);
            L += e;
        }
    }
    __half* ctx = p.ctx + (size_t)b * H * DH + hh * DH;
    if (nu == 1) {
        if (tid < DH) ctx[tid] = __float2half_rn(O * (1.f / (L + 1e-6f)));   // template.hpp:1632: 1 / (sum + 1e-6)
        cbar();
        return;
    }
    float* part = p.att_part + ((size_t)(b * H + hh) * p.att_max_units + j) * (DH + 2);
    if (tid < DH) part[tid] = O;
    if (tid == 0) {
        part[DH] = M;
        part[DH + 1] = L;
    }
    __threadfence();
    cbar();
    if (tid == 0) {
        const int old = atomicAdd(&p.att_cnt[b * H + hh], 1);
        sm.flag = old == nu - 1;
    }
    cbar();
    if (sm.flag) {
        __threadfence();
        const float* all = p.att_part + (size_t)(b * H + hh) * p.att_max_units * (DH + 2);
        if (tid < DH) {
            float Mg = -INFINITY;
            for (int u = 0; u < nu; ++u) Mg = fmaxf(Mg, ld_cg_f32(&all[u * (DH + 2) + DH]));
            float Lg = 0.f, Og = 0.f;
            for (int u = 0; u < nu; ++u) {
                const float mi = ld_cg_f32(&all[u * (DH + 2) + DH]);
                const float wgt = mi == -INFINITY ? 0.f : __expf(mi - Mg);
                Lg = fmaf(ld_cg_f32(&all[u * (DH + 2) + DH + 1]), wgt, Lg);
                Og = fmaf(ld_cg_f32(&all[u * (DH + 2) + tid]), wgt, Og);
            }
            ctx[tid] = __float2half_rn(Og * (1.f / (Lg + 1e-6f)));
        }
        if (tid == 0) p.att_cnt[b * H + hh] = 0;
    }
    cbar();
}

// attention unit index a in [0, NA) -> (sequence, head, unit of the head)
__device__ __forceinline__ void att_decode(const Smem& sm, int B, int a, int& b, int& hh, int& j)
{
    b = 0;
    while (b + 1 < B && a >= sm.att.pre[b + 1]) ++b;
    const int r = a - sm.att.pre[b];
    hh = r / sm.att.nu[b];
    j = r - hh * sm.att.nu[b];
}

// ------------------------------------------------------------------------------------------------ the kernel
template <bool W8, int DH>
__global__ void __launch_bounds__(THREADS, 1) decode_mega_kernel(const __grid_constant__ Params p)
{
    using WT = typename std::conditional<W8, uint8_t, __half>::type;
    // NOTE: no integer round trip on this pointer -- the compiler must keep seeing shared memory (LDS/STS, not generic LD/ST)
    extern __shared__ __align__(128) uint8_t smem_raw[];
    uint8_t* ring = smem_raw;
    uint8_t* opnd = ring + (size_t)p.ns * STAGE_BYTES;
    Smem& sm = *reinterpret_cast<Smem*>(opnd + (size_t)p.B * p.opnd_pitch);

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    constexpr int wsz = (int)sizeof(WT);

    if (tid == 0) {
        for (int s = 0; s < p.ns; ++s) {
            mbar_init(&sm.full[s], 1);
            mbar_init(&sm.empty[s], CW);
        }
        mbar_fence_init();
        // attention bookkeeping (identical in every CTA)
        int acc = 0;
        const int step = *p.step;
        for (int b = 0; b < p.B; ++b) {
            const int tlen = p.seq_len[b];
            const int inl = min(p.input_len[b], tlen);
            const int nvalid = inl + max(0, tlen - p.max_in);
            const bool fin = p.finished != nullptr && p.finished[b] != 0;
            sm.att.tlen[b] = tlen;
            sm.att.inl[b] = inl;
            sm.att.nvalid[b] = nvalid;
            sm.att.nu[b] = fin ? 0 : max(1, (nvalid + ATT_UNIT - 1) / ATT_UNIT);
            sm.att.pre[b] = acc;
            acc += sm.att.nu[b] * p.Hl;
            sm.att.pos[b] = (step - 1) - p.pad_count[b];
            int id = p.embed ? p.out_ids[(size_t)(step - 1) * p.B + b] : 0;
            sm.att.tok[b] = min(max(id, 0), p.vocab - 1);
        }
        sm.att.pre[p.B] = acc;
    }
    __syncthreads();

    const int NA = (p.dbg & 4) ? 0 : sm.att.pre[p.B];
    const int Tq = (3 * p.hl + ROWS - 1) / ROWS, Tf1 = (p.inter + ROWS - 1) / ROWS, Th = (p.h + ROWS - 1) / ROWS;
    const int hB = p.h * wsz, hlB = p.hl * wsz, interB = p.inter * wsz;
    const int KU = p.ks > 1 ? hB : interB;      // FFN2 k bytes per unit
    RingPos rp{0, 0};
    int phase = 0;

    if (warp >= CW) {
        // ======================================= producer =======================================
        // register budget: the CTA owns 384 x 168; the producer warpgroup keeps 40 per thread and the 256 consumer threads may
        // then grow to 168 + 128 x (168 - 40) / 256 = 232 (asking for more would block forever)
        asm volatile("setmaxnreg.dec.sync.aligned.u32 40;");
        if (warp != CW) return;
        Throttle th{RingPos{0, 0}, 0, (p.dbg & 128) ? 0ull : l2_evict_first_policy(), 0, 0};
        const long long t_start = clock64();
        for (int l = p.l0; l < p.l1; ++l) {
            const LayerDev* Ld = p.layers + l;
            int u0, u1;
            // phase A: QKV tiles, then FFN1 tiles (k = h)
            my_range(Tq + Tf1, phase++ * 53, u0, u1);
            for (int u = u0; u < u1; ++u) {
                const bool is_q = u < Tq;
                const int row0 = (is_q ? u : u - Tq) * ROWS;
                const uint8_t* W = static_cast<const uint8_t*>(Ld->w[is_q ? 0 : 2]);
                produce_weight_unit(p, ring, sm, rp, th, W + (size_t)row0 * hB, hB, lane);
            }
            // phase B: attention units, then FFN2 split-k tiles
            my_range(NA + p.ks * Th, phase++ * 53, u0, u1);
            for (int u = u0; u < u1; ++u) {
                if (u < NA) {
                    int b, hh, j;
                    att_decode(sm, p.B, u, b, hh, j);
                    const __half* kc = p.kv + (size_t)(2 * l) * p.kv_layer_elems + ((size_t)b * p.Hl + hh) * (size_t)p.max_len * DH;
                    const __half* vc = p.kv + (size_t)(2 * l + 1) * p.kv_layer_elems + ((size_t)b * p.Hl + hh) * (size_t)p.max_len * DH;
                    const int v0 = j * ATT_UNIT, v1 = min(sm.att.nvalid[b], v0 + ATT_UNIT);
                    for (int tv = v0; tv < v1; tv += ATT_TILE) {
                        const int nk = min(ATT_TILE, v1 - tv);
                        produce_kv_tile(p, ring, sm, rp, th, kc, b, tv, nk, lane);
                        produce_kv_tile(p, ring, sm, rp, th, vc, b, tv, nk, lane);
                    }
                } else {
                    const int kc = (u - NA) / Th, row0 = ((u - NA) % Th) * ROWS;
                    const uint8_t* W = static_cast<const uint8_t*>(Ld->w[3]);
                    produce_weight_unit(p, ring, sm, rp, th, W + (size_t)row0 * interB + (size_t)kc * KU * ROWS, min(KU, interB - kc * KU), lane);
                }
            }
            // phase C: O-projection tiles (k = hl)
            my_range(Th, phase++ * 53, u0, u1);
            for (int u = u0; u < u1; ++u) {
                const int row0 = u * ROWS;
                const uint8_t* W = static_cast<const uint8_t*>(Ld->w[1]);
                produce_weight_unit(p, ring, sm, rp, th, W + (size_t)row0 * hlB, hlB, lane);
            }
        }
        if (p.lm_rows > 0) {
            int u0, u1;
            my_range((p.lm_rows + ROWS - 1) / ROWS, phase++ * 53, u0, u1);
            for (int u = u0; u < u1; ++u) {
                const int row0 = u * ROWS;
                produce_weight_unit(p, ring, sm, rp, th, static_cast<const uint8_t*>(p.lm_head) + (size_t)row0 * p.h * 2, p.h * 2, lane);
            }
        }
        if ((p.dbg & 64) && lane == 0 && blockIdx.x < TS_CTAS) {
            g_mega_cyc[0][blockIdx.x] = th.c_empty;
            g_mega_cyc[1][blockIdx.x] = th.c_thr;
            g_mega_cyc[2][blockIdx.x] = clock64() - t_start;
        }
        return;
    }

    // ======================================= consumers =======================================
    asm volatile("setmaxnreg.inc.sync.aligned.u32 232;");
    int redbuf = 0;
    unsigned nbar = 0;
    const long long t_start_c = clock64();
    if (lane == 0) sm.c_full[warp] = 0;
    for (int l = p.l0; l < p.l1; ++l) {
        const LayerDev* Ld = p.layers + l;
        int u0, u1;
        if (tid < p.B) sm.xrow[tid] = (p.embed && l == 0) ? p.wte + (size_t)sm.att.tok[tid] * p.h : p.x + (size_t)tid * p.h;
        cbar();
        // ---------------- phase A
        my_range(Tq + Tf1, phase++ * 53, u0, u1);
        {
            int staged = -1;   // which LayerNorm the operand buffer holds: 0 = LN1(x), 1 = LN2(x)
            bool have_stats = false;
            for (int u = u0; u < u1; ++u) {
                const bool is_q = u < Tq;
                if (staged != (is_q ? 0 : 1)) {
                    stage_layernorm(p, opnd, sm, is_q ? Ld->ln1_g : Ld->ln2_g, is_q ? Ld->ln1_b : Ld->ln2_b, have_stats, warp, lane);
                    staged = is_q ? 0 : 1;
                    have_stats = true;
                }
                const int row0 = (is_q ? u : u - Tq) * ROWS;
                gemm_unit<WT, W8>(p, Ld, ring, sm, rp, redbuf, opnd, hB, row0, is_q ? 3 * p.hl : p.inter, is_q ? EP_QKV : EP_FFN1, 0, warp, lane);
            }
        }
        grid_barrier(p.gbar, ++nbar * gridDim.x, tid, p.dbg);
        // ---------------- phase B
        my_range(NA + p.ks * Th, phase++ * 53, u0, u1);
        {
            int staged_kc = -1;
            for (int u = u0; u < u1; ++u) {
                if (u < NA) {
                    int b, hh, j;
                    att_decode(sm, p.B, u, b, hh, j);
                    attention_unit<DH>(p, Ld, l, ring, sm, rp, b, hh, j, warp, lane);
                } else {
                    const int kc = (u - NA) / Th, row0 = ((u - NA) % Th) * ROWS;
                    const int kb = min(KU, interB - kc * KU);
                    if (staged_kc != kc) {
                        cbar();   // everyone is done reading the previous operand
                        if (tid < p.B) sm.src[tid] = p.inter_buf + (size_t)tid * p.inter;
                        cbar();
                        stage_rows(p, opnd, sm.src, (size_t)kc * (KU / wsz), kb / wsz, tid);
                        cbar();
                        staged_kc = kc;
                    }
                    gemm_unit<WT, W8>(p, Ld, ring, sm, rp, redbuf, opnd, kb, row0, p.h, EP_FFN2, kc, warp, lane);
                }
            }
        }
        grid_barrier(p.gbar, ++nbar * gridDim.x, tid, p.dbg);
        // ---------------- phase C
        my_range(Th, phase++ * 53, u0, u1);
        if (u1 > u0) {
            if (tid < p.B) sm.src[tid] = p.ctx + (size_t)tid * p.hl;
            cbar();
            stage_rows(p, opnd, sm.src, 0, p.hl, tid);
            cbar();
            for (int u = u0; u < u1; ++u) gemm_unit<WT, W8>(p, Ld, ring, sm, rp, redbuf, opnd, hlB, u * ROWS, p.h, EP_O, 0, warp, lane);
        }
        grid_barrier(p.gbar, ++nbar * gridDim.x, tid, p.dbg);
    }
    if (p.lm_rows > 0) {
        int u0, u1;
        my_range((p.lm_rows + ROWS - 1) / ROWS, phase++ * 53, u0, u1);
        if (tid < p.B) sm.xrow[tid] = p.x + (size_t)tid * p.h;
        cbar();
        if (u1 > u0) stage_layernorm(p, opnd, sm, p.lnf_g, p.lnf_b, false, warp, lane);
        for (int u = u0; u < u1; ++u) gemm_unit<__half, false>(p, nullptr, ring, sm, rp, redbuf, opnd, p.h * 2, u * ROWS, p.lm_rows, EP_LM, 0, warp, lane);
    }
    if ((p.dbg & 64) && lane == 0 && blockIdx.x < TS_CTAS && (warp == 0 || warp == 7)) {
        g_mega_cyc[warp == 0 ? 3 : 5][blockIdx.x] = sm.c_full[warp];
        g_mega_cyc[4][blockIdx.x] = clock64() - t_start_c;
    }
}

}  // namespace mg

// ------------------------------------------------------------------------------------------------ host side
__global__ void __launch_bounds__(256) retile_kernel(const uint4* __restrict__ in, uint4* __restrict__ out, int n, int kbytes)
{
    const int cpr = kbytes >> 4;   // 16-byte chunks per row
    const size_t total = (size_t)n * cpr;
    for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
        const int row = (int)(i / cpr), c_abs = (int)(i % cpr);
        const int rt = row / mg::ROWS, r = row % mg::ROWS;
        const int kc = c_abs / (mg::STAGE_K / 16), cw = c_abs % (mg::STAGE_K / 16);
        const int kw = min(mg::STAGE_K, kbytes - kc * mg::STAGE_K);
        const size_t off = (size_t)rt * mg::ROWS * kbytes + (size_t)kc * mg::STAGE_BYTES + (size_t)r * kw + ((size_t)(cw ^ (r & 7)) << 4);
        out[off >> 4] = in[i];
    }
}

size_t mega_tiled_bytes(int n, int kbytes) { return (size_t)((n + mg::ROWS - 1) / mg::ROWS) * mg::ROWS * kbytes; }

int mega_retile(const void* w_nk, void* out, int n, int kbytes, cudaStream_t st)
{
    FTCF_REQUIRE(w_nk && out && n > 0 && kbytes > 0 && kbytes % 128 == 0, FTCF_ERR_INVALID, "retile: n=%d kbytes=%d", n, kbytes);
    FTCF_CUDA_CHECK(cudaMemsetAsync(out, 0, mega_tiled_bytes(n, kbytes), st));
    retile_kernel<<<1184, 256, 0, st>>>(static_cast<const uint4*>(w_nk), static_cast<uint4*>(out), n, kbytes);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

std::atomic<int> g_mega_dbg{0}, g_mega_ns{0}, g_mega_inflight{4};

bool mega_supported(int B, int h, int hl, int inter, int dh, int rot, bool w8, int tp, bool parallel_residual)
{
    const int wsz = w8 ? 1 : 2;
    if (B < 1 || B > mg::MAX_B || tp != 1 || !parallel_residual) return false;
    if (dh != 64 && dh != 128) return false;
    if (rot % 2 != 0 || rot > dh) return false;
    // every k extent must be a whole number of 128-byte k-steps and the operand rows 16-byte vectors
    if ((h * wsz) % 128 != 0 || (hl * wsz) % 128 != 0 || (inter * wsz) % 128 != 0 || (h * 2) % 128 != 0) return false;
    if (h % 8 != 0 || hl % 8 != 0 || inter % 8 != 0) return false;
    return true;
}

int mega_plan(mg::Params& p, bool w8)
{
    const int wsz = w8 ? 1 : 2;
    p.ks = (p.inter % p.h == 0 && p.inter > p.h && (p.h * wsz) % mg::STAGE_K == 0) ? p.inter / p.h : 1;   // split k on block boundaries only
    const int kmax_elems = std::max(std::max(p.h, p.hl), p.ks > 1 ? p.h : p.inter);
    p.opnd_pitch = kmax_elems * 2 + 16;
    p.att_max_units = (p.max_len + mg::ATT_UNIT - 1) / mg::ATT_UNIT + 1;
    int dev = 0, max_smem = 0;
    FTCF_CUDA_CHECK(cudaGetDevice(&dev));
    FTCF_CUDA_CHECK(cudaDeviceGetAttribute(&max_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev));
    const long long fixed = (long long)p.B * p.opnd_pitch + (long long)sizeof(mg::Smem) + 256 + 1024 /* static */;
    long long ns = (max_smem - fixed) / mg::STAGE_BYTES;
    if (ns > 16) ns = 16;
    if (g_mega_ns.load() > 0 && ns > g_mega_ns.load()) ns = g_mega_ns.load();
    p.inflight = g_mega_inflight.load();
    if (p.inflight < 1 || p.inflight > ns) p.inflight = (int)ns;
    FTCF_REQUIRE(ns >= 4, FTCF_ERR_UNSUPPORTED, "decode megakernel: batch %d x k %d leaves room for only %lld ring stages", p.B, kmax_elems, ns);
    p.ns = (int)ns;
    return FTCF_OK;
}

// diagnostics (not part of include/ftcf.h): copies the barrier timestamps of the last dbg & 64 launch to the host
extern "C" int ftcf_debug_mega_timestamps(unsigned long long* out, size_t bytes)
{
    const size_t want = sizeof(unsigned long long) * 3 * mg::TS_BARS * mg::TS_CTAS;
    if (bytes < want) return (int)want;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, mg::g_mega_ts, want) == cudaSuccess ? 0 : -1;
}
extern "C" int ftcf_debug_mega_cycles(long long* out, size_t bytes)
{
    const size_t want = sizeof(long long) * 8 * mg::TS_CTAS;
    if (bytes < want) return (int)want;
    cudaDeviceSynchronize();
    return cudaMemcpyFromSymbol(out, mg::g_mega_cyc, want) == cudaSuccess ? 0 : -1;
}

size_t mega_smem_bytes(const mg::Params& p)
{
    return (size_t)p.ns * mg::STAGE_BYTES + (size_t)p.B * p.opnd_pitch + sizeof(mg::Smem) + 256;
}

int mega_launch(const mg::Params& p_in, bool w8, cudaStream_t st)
{
    mg::Params p = p_in;
    p.dbg = g_mega_dbg.load(std::memory_order_relaxed);
    int dev = 0, sms = 0;
    FTCF_CUDA_CHECK(cudaGetDevice(&dev));
    FTCF_CUDA_CHECK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev));
    const size_t smem = mega_smem_bytes(p);
    FTCF_CUDA_CHECK(cudaMemsetAsync(p.gbar, 0, sizeof(unsigned), st));
    cudaError_t err = cudaSuccess;
#define FTCF_MEGA(W8_, DH_)                                                                                              \
    do {                                                                                                                 \
        static size_t configured = 0;                                                                                    \
        auto kern = mg::decode_mega_kernel<W8_, DH_>;                                                                    \
        if (configured < smem) {                                                                                         \
            err = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);                    \
            if (err == cudaSuccess) {                                                                                    \
                int occ = 0;                                                                                             \
                err = cudaOccupancyMaxActiveBlocksPerMultiprocessor(&occ, kern, mg::THREADS, smem);                      \
                if (err == cudaSuccess && occ < 1) {                                                                     \
                    set_error("decode megakernel: a CTA with %zu bytes of shared memory does not fit an SM", smem);      \
                    return FTCF_ERR_UNSUPPORTED;                                                                         \
                }                                                                                                        \
            }                                                                                                            \
            configured = smem;                                                                                           \
        }                                                                                                                \
        if (err == cudaSuccess) kern<<<sms, mg::THREADS, smem, st>>>(p);                                                 \
    } while (0)
    if (w8) {
        if (p.dh == 128) FTCF_MEGA(true, 128);
        else FTCF_MEGA(true, 64);
    } else {
        if (p.dh == 128) FTCF_MEGA(false, 128);
        else FTCF_MEGA(false, 64);
    }
#undef FTCF_MEGA
    FTCF_REQUIRE(err == cudaSuccess, FTCF_ERR_CUDA, "decode megakernel: %s", cudaGetErrorString(err));
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

}  // namespace ftcf
