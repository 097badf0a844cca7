// Attention kernels of the GPT-NeoX path: the per-token decode attention over the KV cache (HBM-bound) and the
// prefill-side bias + rotary + cache scatter and causal attention.
//
// Reference behaviour restated (not ported):
//   * decode: masked_multihead_attention_kernel,
//     kernels/decoder_masked_multihead_attention/decoder_masked_multihead_attention_template.hpp:1099-1919 --
//     q,k,v += bias (fp16), NeoX rotary on the first `rot` dims pairing (i, i + rot/2) at position
//     timestep - pad_count (:1303-1365, utils.h:1325-1337), append k,v at slot seq_len[b], scores = q.k / sqrt(dh) in
//     fp32, the pad gap [input_len, max_input_len) is masked, softmax with 1/(sum + 1e-6), out = P.V;
//     finished rows return immediately (:1176-1178).
//   * prefill: add_fusedQKV_bias_transpose_kernel + transpose_4d_batch_major_{k,v}_cache
//     (kernels/unfused_attention_kernels.cu:1326-1484,1673-1757) and the unfused QK^T / masked softmax / PV chain
//     (layers/attention_layers/GptContextAttentionLayer.cc:194-300, unfused_attention_kernels.cu:255-333).
//
// B200 design of the decode kernel (roofline: HBM, 2 * c * dh * 2 bytes per (sequence, head) per step):
//   * cache layout K,V = [B, heads, max_len, dh] fp16: one (b, h) pair is a single contiguous stream;
//   * a cache row (dh halves) is read by dh/8 adjacent lanes with one 128-bit no-allocate load each, four rows in
//     flight per lane; dot products reduced with warp shuffles inside the row group;
//   * split-KV: grid = (heads, B, splits) so that B*heads*splits covers the 148 SMs several times even at B = 1;
//     every split writes (max, sum, unnormalised out) and the last CTA to arrive (atomic ticket, self-resetting)
//     merges them -- no second launch;
//   * masked pad-gap rows are never loaded.
#include "tma_utils.cuh"

namespace ftcf {

constexpr int MMHA_THREADS = 128;
constexpr int MMHA_BULK_KEYS = 64;       // keys per CTA of the bulk-staged decode attention
std::atomic<int> g_mmha_bulk{1};         // tunable "mmha_bulk": bulk-staged decode attention at small batch (0: off)
constexpr int MMHA_MAX_CHUNK = 4096;
std::atomic<int> g_mmha_splits{0};       // tunable "mmha_splits": force the split count (0: automatic)
std::atomic<int> g_mmha_onepass{1};      // tunable "mmha_onepass": one-pass (online softmax) decode attention; 0: the two-pass kernel
std::atomic<int> g_prefill_mma{1};       // tunable "prefill_mma": tensor-core prefill attention (0: the CUDA-core kernel)
std::atomic<int> g_mmha_prefetch{0};     // tunable "mmha_prefetch": L2 prefetch of the split's cache rows before the dependency wait
std::atomic<int> g_mmha_lite{1};         // tunable "mmha_lite": 64-thread bulk kernel with only the K tile staged (one wave of CTAs beside the GEMMs)
std::atomic<int> g_mmha_pdl{0};          // tunable "mmha_pdl": launch the decode attention with programmatic dependent launch   // keys per split (fp32 scores kept in shared memory)

__device__ __forceinline__ float rotary_angle(int pos, int i, int rot)
{
    // decoder_masked_multihead_attention_utils.h:1325-1329: t_step / 10000^(2i/rot)
    return (float)pos / powf(10000.f, (2.f * (float)i) / (float)rot);
}

// x[d] (already biased, fp16) and its NeoX partner -> rotated value, rounded to fp16
__device__ __forceinline__ __half rotary_neox(__half xd, __half xpartner, int d, int rot, int pos)
{
    const int half_rot = rot >> 1;
    const int i = d < half_rot ? d : d - half_rot;
    float sn, cs;
    sincosf(rotary_angle(pos, i, rot), &sn, &cs);
    const float a = __half2float(xd), b = __half2float(xpartner);
    return __float2half_rn(d < half_rot ? cs * a - sn * b : cs * a + sn * b);
}

struct MmhaP {
    ftcf_mmha_params p;
    int prefetch;
};

template <int DH, bool BEAMS = false>
__global__ void __launch_bounds__(MMHA_THREADS) mmha_decode_kernel(const MmhaP params)
{
    const ftcf_mmha_params& p = params.p;
    constexpr int LPR = DH / 8;               // lanes per cache row (16 bytes each)
    constexpr int NG = MMHA_THREADS / LPR;    // row groups per CTA
    constexpr int UNR = 4;                    // rows per group per block, all in flight at once
    constexpr int BLK = NG * UNR;             // keys per block (32 at dh = 128)

    __shared__ float s_scores[MMHA_MAX_CHUNK];
    __shared__ float s_out[NG][DH];
    __shared__ __align__(16) __half s_q[DH];
    __shared__ __align__(16) __half s_k[DH];
    __shared__ __align__(16) __half s_v[DH];
    __shared__ float s_red[32];
    __shared__ int s_flag;

    const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
    const int H = p.heads, tid = threadIdx.x;
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    pdl_launch_dependents();                  // the O-projection GEMM may start streaming its weights now
    // Everything read before pdl_wait() is constant for the whole decode step: the request state (lengths, finished flags: last
    // written by the previous step's sampling kernels, which are ordinary launches and so completed before this step began)
    // and cache rows of earlier positions.
    if (p.finished != nullptr && p.finished[b]) return;

    const int tlen = p.seq_len[b];
    const int total = tlen + 1;
    const int chunk = ceil_div(total, p.splits);
    const int start = split * chunk;
    const int end = min(start + chunk, total);
    const int owner = tlen / chunk;           // the split that holds the new token
    const int in_len = p.input_len[b], max_in = p.max_input_len;

    const __half* qkv = static_cast<const __half*>(p.qkv) + (size_t)b * 3 * H * DH;
    const __half* bias = static_cast<const __half*>(p.qkv_bias);
    __half* kc = static_cast<__half*>(p.k_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    __half* vc = static_cast<__half*>(p.v_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    const int li = tid % LPR, gi = tid / LPR;

    // ---- L2 prefetch of this split's cached K and V rows, issued BEFORE the dependency wait: the cache does not depend on the
    // QKV GEMM that is still running, so the DRAM latency of the rows hides behind that kernel's tail and the loads after the
    // wait are L2 hits.  (Keeping the rows in registers instead was measured slower: at 128 registers x 256 threads only two
    // CTAs fit per SM next to the GEMM CTAs and the grid ran in waves.)
    for (int pos = start + (tid >> 1); pos < end && params.prefetch && !BEAMS; pos += MMHA_THREADS / 2) {
        if (pos == tlen || (pos >= in_len && pos < max_in)) continue;
        const __half* src = ((tid & 1) ? vc : kc) + (size_t)pos * DH;
#pragma unroll
        for (int c = 0; c < DH * 2; c += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(reinterpret_cast<const char*>(src) + c));
    }

    pdl_wait();                               // qkv of this layer is complete and visible
    const unsigned long long trc_t1 = trc_now(threadIdx.x == 0);
    // Beam search: slot `pos` of this row lives in the cache row of the beam the indirection names (the buffer the previous
    // beam-search step wrote); the new token goes to the row's own slot (template.hpp:1494-1522,1709-1761).
    const int32_t* indir = nullptr;
    const __half *kc0 = kc, *vc0 = vc;        // cache of beam 0 of this row's batch, head h
    if constexpr (BEAMS) {
        indir = p.cache_indir + ((size_t)((*p.step - p.max_input_len) & 1) * p.batch + b) * p.max_len;
        const int beam0 = (b / p.beam_width) * p.beam_width;
        kc0 = static_cast<const __half*>(p.k_cache) + ((size_t)beam0 * H + h) * (size_t)p.max_len * DH;
        vc0 = static_cast<const __half*>(p.v_cache) + ((size_t)beam0 * H + h) * (size_t)p.max_len * DH;
    }
    const size_t beam_stride = (size_t)H * p.max_len * DH;

    // ---- q (all splits), k / v (owner split): bias, rotary, append to the cache
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        const int rot = p.rotary_dim;
        const int pos = (*p.step - 1) - p.pad_count[b];
        const int qi = h * DH + d;
        __half q = qkv[qi];
        if (bias) q = __hadd(q, bias[qi]);
        const bool do_rot = d < rot;
        const int dp = d < (rot >> 1) ? d + (rot >> 1) : d - (rot >> 1);
        if (do_rot) {
            __half qp = qkv[h * DH + dp];
            if (bias) qp = __hadd(qp, bias[h * DH + dp]);
            q = rotary_neox(q, qp, d, rot, pos);
        }
        s_q[d] = q;
        if (split == owner) {
            const int ki = H * DH + qi, vi = 2 * H * DH + qi;
            __half k = qkv[ki], v = qkv[vi];
            if (bias) {
                k = __hadd(k, bias[ki]);
                v = __hadd(v, bias[vi]);
            }
            if (do_rot) {
                __half kp = qkv[H * DH + h * DH + dp];
                if (bias) kp = __hadd(kp, bias[H * DH + h * DH + dp]);
                k = rotary_neox(k, kp, d, rot, pos);
            }
            s_k[d] = k;
            s_v[d] = v;
            kc[(size_t)tlen * DH + d] = k;
            vc[(size_t)tlen * DH + d] = v;
        }
    }
    __syncthreads();

    float q[8];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(&s_q[li * 8]);
        const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(qh[i]);
            q[2 * i] = f.x;
            q[2 * i + 1] = f.y;
        }
    }

    // ---- scores
    for (int base = start; base < end; base += BLK) {
        uint4 kv[UNR];
        bool valid[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int pos = base + gi + u * NG;
            valid[u] = pos < end && !(pos >= in_len && pos < max_in);
            kv[u] = make_uint4(0, 0, 0, 0);
            if (valid[u]) {
                if (pos == tlen) kv[u] = *reinterpret_cast<const uint4*>(&s_k[li * 8]);
                else if (BEAMS) kv[u] = ld_stream_16(kc0 + (size_t)indir[pos] * beam_stride + (size_t)pos * DH + li * 8);
                else kv[u] = ld_stream_16(kc + (size_t)pos * DH + li * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const __half2* kh = reinterpret_cast<const __half2*>(&kv[u]);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(kh[i]);
                dot = fmaf(q[2 * i], f.x, dot);
                dot = fmaf(q[2 * i + 1], f.y, dot);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            const int pos = base + gi + u * NG;
            if (li == 0 && pos < end) s_scores[pos - start] = valid[u] ? dot * p.inv_sqrt_dh : -INFINITY;
        }
    }
    __syncthreads();

    // ---- softmax statistics of this split
    const int cnt = max(end - start, 0);
    float mx = -INFINITY;
    for (int i = tid; i < cnt; i += MMHA_THREADS) mx = fmaxf(mx, s_scores[i]);
    mx = block_max(mx, s_red);
    float sum = 0.f;
    for (int i = tid; i < cnt; i += MMHA_THREADS) {
        const float sc = s_scores[i];
        const float e = (sc == -INFINITY) ? 0.f : __expf(sc - mx);
        s_scores[i] = e;
        sum += e;
    }
    sum = block_sum(sum, s_red);
    __syncthreads();

    // ---- P.V
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int base = start; base < end; base += BLK) {
        uint4 vv[UNR];
        float pr[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int pos = base + gi + u * NG;
            pr[u] = pos < end ? s_scores[pos - start] : 0.f;
            vv[u] = make_uint4(0, 0, 0, 0);
            if (pr[u] != 0.f) {
                if (pos == tlen) vv[u] = *reinterpret_cast<const uint4*>(&s_v[li * 8]);
                else if (BEAMS) vv[u] = ld_stream_16(vc0 + (size_t)indir[pos] * beam_stride + (size_t)pos * DH + li * 8);
                else vv[u] = ld_stream_16(vc + (size_t)pos * DH + li * 8);
            }
        }
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const __half2* vh = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(vh[i]);
                acc[2 * i] = fmaf(pr[u], f.x, acc[2 * i]);
                acc[2 * i + 1] = fmaf(pr[u], f.y, acc[2 * i + 1]);
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_out[gi][li * 8 + i] = acc[i];
    __syncthreads();

    __half* ctx = static_cast<__half*>(p.ctx) + (size_t)b * H * DH + h * DH;
    if (p.splits == 1) {
        for (int d = tid; d < DH; d += MMHA_THREADS) {
            float o = 0.f;
#pragma unroll
            for (int g = 0; g < NG; ++g) o += s_out[g][d];
            ctx[d] = __float2half_rn(o * (1.f / (sum + 1e-6f)));
        }
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t1, end - start, 0);
        return;
    }

    // ---- split-KV: publish the partial, the last arriver merges
    float* part = p.partial + ((size_t)(b * H + h) * p.splits + split) * (DH + 2);
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        float o = 0.f;
#pragma unroll
        for (int g = 0; g < NG; ++g) o += s_out[g][d];
        part[d] = o;
    }
    if (tid == 0) {
        part[DH] = mx;
        part[DH + 1] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&p.counters[b * H + h], 1);
        s_flag = (old == p.splits - 1);
    }
    __syncthreads();
    if (!s_flag) {
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t1, end - start, 0);
        return;
    }
    const unsigned long long trc_t2 = trc_now(threadIdx.x == 0);
    __threadfence();
    const float* all = p.partial + (size_t)(b * H + h) * p.splits * (DH + 2);
    float M = -INFINITY;
    for (int s2 = 0; s2 < p.splits; ++s2) M = fmaxf(M, __ldcg(&all[s2 * (DH + 2) + DH]));
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        float L = 0.f, O = 0.f;
        for (int s2 = 0; s2 < p.splits; ++s2) {
            const float mi = __ldcg(&all[s2 * (DH + 2) + DH]);
            const float wgt = (mi == -INFINITY) ? 0.f : __expf(mi - M);
            L = fmaf(__ldcg(&all[s2 * (DH + 2) + DH + 1]), wgt, L);
            O = fmaf(__ldcg(&all[s2 * (DH + 2) + d]), wgt, O);
        }
        ctx[d] = __float2half_rn(O * (1.f / (L + 1e-6f)));
    }
    if (tid == 0) {
        p.counters[b * H + h] = 0;
        trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t2, end - start, 1);
    }
}

// ---------------------------------------------------------------- decode attention, one pass (online softmax)
// Same semantics and split-KV protocol as mmha_decode_kernel below, but K and V rows of a block are requested TOGETHER and
// consumed in one loop with a running (max, sum, out) per row group: half as many dependent DRAM round trips per CTA (the
// kernel is latency-bound: 66 -> 33 rounds at context 1030 without splits), no score buffer in shared memory (22 KB -> 5 KB,
// more CTAs per SM).  The 8 row groups of the CTA are merged once at the end, then the usual per-split partial + last-arriver
// merge follows.
template <int DH>
__global__ void __launch_bounds__(MMHA_THREADS, 6) mmha_decode_onepass_kernel(const MmhaP params)
{
    const ftcf_mmha_params& p = params.p;
    constexpr int LPR = DH / 8;               // lanes per cache row (16 bytes each)
    constexpr int NG = MMHA_THREADS / LPR;    // row groups per CTA
    constexpr int UNR = 4;                    // rows per group per block: 4 K + 4 V loads in flight per lane
    constexpr int BLK = NG * UNR;

    __shared__ float s_out[NG][DH];
    __shared__ float s_ml[NG][2];
    __shared__ __align__(16) __half s_q[DH];
    __shared__ __align__(16) __half s_k[DH];
    __shared__ __align__(16) __half s_v[DH];
    __shared__ int s_flag;

    const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
    const int H = p.heads, tid = threadIdx.x;
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    if (p.finished != nullptr && p.finished[b]) return;

    const int tlen = p.seq_len[b];
    const int total = tlen + 1;
    const int chunk = ceil_div(total, p.splits);
    const int start = split * chunk;
    const int end = min(start + chunk, total);
    const int owner = tlen / chunk;           // the split that holds the new token
    const int in_len = p.input_len[b], max_in = p.max_input_len;

    const __half* qkv = static_cast<const __half*>(p.qkv) + (size_t)b * 3 * H * DH;
    const __half* bias = static_cast<const __half*>(p.qkv_bias);
    __half* kc = static_cast<__half*>(p.k_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    __half* vc = static_cast<__half*>(p.v_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    const int li = tid % LPR, gi = tid / LPR;

    // ---- q (all splits), k / v (owner split): bias, rotary, append to the cache
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        const int rot = p.rotary_dim;
        const int pos = (*p.step - 1) - p.pad_count[b];
        const int qi = h * DH + d;
        __half q = qkv[qi];
        if (bias) q = __hadd(q, bias[qi]);
        const bool do_rot = d < rot;
        const int dp = d < (rot >> 1) ? d + (rot >> 1) : d - (rot >> 1);
        if (do_rot) {
            __half qp = qkv[h * DH + dp];
            if (bias) qp = __hadd(qp, bias[h * DH + dp]);
            q = rotary_neox(q, qp, d, rot, pos);
        }
        s_q[d] = q;
        if (split == owner) {
            const int ki = H * DH + qi, vi = 2 * H * DH + qi;
            __half k = qkv[ki], v = qkv[vi];
            if (bias) {
                k = __hadd(k, bias[ki]);
                v = __hadd(v, bias[vi]);
            }
            if (do_rot) {
                __half kp = qkv[H * DH + h * DH + dp];
                if (bias) kp = __hadd(kp, bias[H * DH + h * DH + dp]);
                k = rotary_neox(k, kp, d, rot, pos);
            }
            s_k[d] = k;
            s_v[d] = v;
            kc[(size_t)tlen * DH + d] = k;
            vc[(size_t)tlen * DH + d] = v;
        }
    }
    __syncthreads();

    float q[8];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(&s_q[li * 8]);
        const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(qh[i]);
            q[2 * i] = f.x;
            q[2 * i + 1] = f.y;
        }
    }

    float mrun = -INFINITY, lrun = 0.f, acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    for (int base = start; base < end; base += BLK) {
        uint4 kv[UNR], vv[UNR];
        bool valid[UNR];
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const int pos = base + gi + u * NG;
            valid[u] = pos < end && !(pos >= in_len && pos < max_in);
            kv[u] = make_uint4(0, 0, 0, 0);
            vv[u] = make_uint4(0, 0, 0, 0);
            if (valid[u]) {
                if (pos == tlen) {
                    kv[u] = *reinterpret_cast<const uint4*>(&s_k[li * 8]);
                    vv[u] = *reinterpret_cast<const uint4*>(&s_v[li * 8]);
                } else {
                    kv[u] = ld_stream_16(kc + (size_t)pos * DH + li * 8);
                    vv[u] = ld_stream_16(vc + (size_t)pos * DH + li * 8);
                }
            }
        }
        float sc[UNR];
        float mx = -INFINITY;
#pragma unroll
        for (int u = 0; u < UNR; ++u) {
            const __half2* kh = reinterpret_cast<const __half2*>(&kv[u]);
            float dot = 0.f;
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(kh[i]);
                dot = fmaf(q[2 * i], f.x, dot);
                dot = fmaf(q[2 * i + 1], f.y, dot);
            }
#pragma unroll
            for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
            sc[u] = valid[u] ? dot * p.inv_sqrt_dh : -INFINITY;
            mx = fmaxf(mx, sc[u]);
        }
        const float mnew = fmaxf(mrun, mx);
        if (mnew > -INFINITY) {
            const float corr = __expf(mrun - mnew);            // 0 when mrun is still -inf
            lrun *= corr;
#pragma unroll
            for (int i = 0; i < 8; ++i) acc[i] *= corr;
            mrun = mnew;
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const float pr = (sc[u] == -INFINITY) ? 0.f : __expf(sc[u] - mnew);
                lrun += pr;
                const __half2* vh = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(vh[i]);
                    acc[2 * i] = fmaf(pr, f.x, acc[2 * i]);
                    acc[2 * i + 1] = fmaf(pr, f.y, acc[2 * i + 1]);
                }
            }
        }
    }
#pragma unroll
    for (int i = 0; i < 8; ++i) s_out[gi][li * 8 + i] = acc[i];
    if (li == 0) {
        s_ml[gi][0] = mrun;
        s_ml[gi][1] = lrun;
    }
    __syncthreads();

    // ---- merge the row groups (fixed order)
    float mx = -INFINITY;
#pragma unroll
    for (int g = 0; g < NG; ++g) mx = fmaxf(mx, s_ml[g][0]);
    float sum = 0.f;
#pragma unroll
    for (int g = 0; g < NG; ++g) {
        const float mg = s_ml[g][0];
        sum += (mg == -INFINITY) ? 0.f : s_ml[g][1] * __expf(mg - mx);
    }
    __half* ctx = static_cast<__half*>(p.ctx) + (size_t)b * H * DH + h * DH;
    float* part = p.partial + ((size_t)(b * H + h) * p.splits + split) * (DH + 2);
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        float o = 0.f;
#pragma unroll
        for (int g = 0; g < NG; ++g) {
            const float mg = s_ml[g][0];
            o += (mg == -INFINITY) ? 0.f : s_out[g][d] * __expf(mg - mx);
        }
        if (p.splits == 1) ctx[d] = __float2half_rn(o * (1.f / (sum + 1e-6f)));
        else part[d] = o;
    }
    if (p.splits == 1) {
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t0, trc_t0, end - start, 0);
        return;
    }
    if (tid == 0) {
        part[DH] = mx;
        part[DH + 1] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&p.counters[b * H + h], 1);
        s_flag = (old == p.splits - 1);
    }
    __syncthreads();
    if (!s_flag) {
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t0, trc_t0, end - start, 0);
        return;
    }
    const unsigned long long trc_t2 = trc_now(threadIdx.x == 0);
    __threadfence();
    const float* all = p.partial + (size_t)(b * H + h) * p.splits * (DH + 2);
    float M = -INFINITY;
    for (int s2 = 0; s2 < p.splits; ++s2) M = fmaxf(M, __ldcg(&all[s2 * (DH + 2) + DH]));
    for (int d = tid; d < DH; d += MMHA_THREADS) {
        float L = 0.f, O = 0.f;
        for (int s2 = 0; s2 < p.splits; ++s2) {
            const float mi = __ldcg(&all[s2 * (DH + 2) + DH]);
            const float wgt = (mi == -INFINITY) ? 0.f : __expf(mi - M);
            L = fmaf(__ldcg(&all[s2 * (DH + 2) + DH + 1]), wgt, L);
            O = fmaf(__ldcg(&all[s2 * (DH + 2) + d]), wgt, O);
        }
        ctx[d] = __float2half_rn(O * (1.f / (L + 1e-6f)));
    }
    if (tid == 0) {
        p.counters[b * H + h] = 0;
        trc_emit(TRC_MMHA, trc_t0, trc_t0, trc_t2, end - start, 1);
    }
}

// ---------------------------------------------------------------- decode attention, bulk-staged (small batch)
// At batch <= 8 the kernels above are latency-bound: a CTA walks its keys in dependent rounds of 16 KB, and when the FFN2 GEMM
// streams at the same time every round queues behind its deep TMA ring (measured in the 13B decode step: 17 us alone, 26-29 us
// beside the GEMM, on the critical path QKV -> attention -> O of every layer).  Here a CTA owns at most 64 keys of one
// (sequence, head): ONE thread asks for its whole K tile and V tile with two bulk copies (cp.async.bulk, rows of a (b, h) pair
// are contiguous in the [B, H, max_len, dh] cache) BEFORE the dependency wait -- earlier positions do not depend on this
// layer's QKV GEMM -- so every byte the launch needs is in flight at once and has usually landed when q arrives.  Scores,
// softmax and P.V then run out of shared memory; the split partials are merged by the last arriver as in the kernels above.
//
// TH = 128, STAGE_V: K and V tiles in shared memory (33 KB + 5 KB static, 72 registers x 128 threads per CTA).
// TH = 64, !STAGE_V ("lite"): only the K tile is staged (16 KB), the V rows are prefetched into L2 before the wait and read from
// there after the softmax, the cross-group reduction buffer aliases the K tile: 17.5 KB and 64 registers x 64 threads per CTA.  Beside
// a resident FFN2 CTA (31 K registers, 82 KB) an SM then holds 7-8 attention CTAs instead of 3, so the 680-960 CTAs of a 13B layer at
// context 1024-1536 run as ONE wave (measured: one wave 15 us, two waves 25 us after the dependency resolves, profiles/r2_timeline_*).
template <int DH, int TH, bool STAGE_V>
__global__ void __launch_bounds__(TH, STAGE_V ? 6 : 16) mmha_decode_bulk_kernel(const MmhaP params)
{
    const ftcf_mmha_params& p = params.p;
    constexpr int LPR = DH / 8;               // lanes per cache row (16 bytes each)
    constexpr int NG = TH / LPR;              // keys handled per pass
    constexpr int CH = MMHA_BULK_KEYS;
    static_assert(CH <= TH, "one softmax element per thread");

    extern __shared__ __align__(128) uint8_t bulk_smem[];        // [K tile: CH x DH fp16][V tile: CH x DH fp16 when STAGE_V]
    __shared__ __align__(8) uint64_t bar;
    __shared__ float s_out_static[STAGE_V ? NG : 1][STAGE_V ? DH : 1];
    // lite: the reduction buffer reuses the K tile (dead after the scores)
    float(*s_out)[DH] = STAGE_V ? reinterpret_cast<float(*)[DH]>(&s_out_static[0][0]) : reinterpret_cast<float(*)[DH]>(bulk_smem);
    static_assert(STAGE_V || NG * DH * 4 <= CH * DH * 2, "reduction buffer must fit the K tile");
    __shared__ float s_sc[CH];
    __shared__ float s_red[32];
    __shared__ __align__(16) __half s_q[DH];
    __shared__ __align__(16) __half s_k[DH];
    __shared__ __align__(16) __half s_v[DH];
    __shared__ int s_flag;

    const int h = blockIdx.x, b = blockIdx.y, split = blockIdx.z;
    const int H = p.heads, tid = threadIdx.x;
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    pdl_launch_dependents();
    // request state (lengths, finished flags) is constant within a decode step: safe to read before the dependency wait
    if (p.finished != nullptr && p.finished[b]) return;
    const int tlen = p.seq_len[b];
    const int total = tlen + 1;
    const int nact = ceil_div(total, CH);                 // splits that have keys at this step
    if (split >= nact) return;
    const int chunk = ceil_div(total, nact);              // <= CH, balanced
    const int start = split * chunk;
    const int end = min(start + chunk, total);
    const int cnt = end - start;
    const int owner = tlen / chunk;                       // the split that holds the new token
    const int in_len = p.input_len[b], max_in = p.max_input_len;
    const int nload = min(end, tlen) - start;             // cached rows of this split (the new token's row comes from qkv)

    __half* kc = static_cast<__half*>(p.k_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    __half* vc = static_cast<__half*>(p.v_cache) + ((size_t)b * H + h) * (size_t)p.max_len * DH;
    const __half* sK = reinterpret_cast<const __half*>(bulk_smem);
    const __half* sV = sK + CH * DH;
    if (tid == 0) {
        tma::mbar_init(&bar, 1);
        asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
        const uint32_t bytes = nload > 0 ? (uint32_t)nload * DH * 2 : 0u;
        tma::mbar_arrive_expect_tx(&bar, (STAGE_V ? 2 : 1) * bytes);
        if (nload > 0) {
            asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma::smem_u32(sK)),
                         "l"(kc + (size_t)start * DH), "r"(bytes), "r"(tma::smem_u32(&bar))
                         : "memory");
            if (STAGE_V)
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(tma::smem_u32(sV)),
                             "l"(vc + (size_t)start * DH), "r"(bytes), "r"(tma::smem_u32(&bar))
                             : "memory");
        }
    }
    if (!STAGE_V) {      // the V rows start travelling to L2 now (they do not depend on this layer's QKV GEMM either)
        const char* vsrc = reinterpret_cast<const char*>(vc + (size_t)start * DH);
        for (int off = tid * 128; off < nload * DH * 2; off += TH * 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(vsrc + off));
    }
    pdl_wait();                               // qkv of this layer is complete and visible
    const unsigned long long trc_t1 = trc_now(threadIdx.x == 0);

    const __half* qkv = static_cast<const __half*>(p.qkv) + (size_t)b * 3 * H * DH;
    const __half* bias = static_cast<const __half*>(p.qkv_bias);
    const int li = tid % LPR, gi = tid / LPR;
    // ---- q (all splits), k / v (owner split): bias, rotary, append to the cache
    for (int d = tid; d < DH; d += TH) {
        const int rot = p.rotary_dim;
        const int pos = (*p.step - 1) - p.pad_count[b];
        const int qi = h * DH + d;
        __half q = qkv[qi];
        if (bias) q = __hadd(q, bias[qi]);
        const bool do_rot = d < rot;
        const int dp = d < (rot >> 1) ? d + (rot >> 1) : d - (rot >> 1);
        if (do_rot) {
            __half qp = qkv[h * DH + dp];
            if (bias) qp = __hadd(qp, bias[h * DH + dp]);
            q = rotary_neox(q, qp, d, rot, pos);
        }
        s_q[d] = q;
        if (split == owner) {
            const int ki = H * DH + qi, vi = 2 * H * DH + qi;
            __half k = qkv[ki], v = qkv[vi];
            if (bias) {
                k = __hadd(k, bias[ki]);
                v = __hadd(v, bias[vi]);
            }
            if (do_rot) {
                __half kp = qkv[H * DH + h * DH + dp];
                if (bias) kp = __hadd(kp, bias[H * DH + h * DH + dp]);
                k = rotary_neox(k, kp, d, rot, pos);
            }
            s_k[d] = k;
            s_v[d] = v;
            kc[(size_t)tlen * DH + d] = k;
            vc[(size_t)tlen * DH + d] = v;
        }
    }
    __syncthreads();                          // s_q / s_k / s_v and the barrier initialisation are visible to all threads
    float q[8];
    {
        const uint4 qv = *reinterpret_cast<const uint4*>(&s_q[li * 8]);
        const __half2* qh = reinterpret_cast<const __half2*>(&qv);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(qh[i]);
            q[2 * i] = f.x;
            q[2 * i + 1] = f.y;
        }
    }
    tma::mbar_wait(&bar, 0);                  // both tiles have landed
    const unsigned long long trc_t2 = trc_now(threadIdx.x == 0);

    // ---- scores
    for (int j0 = 0; j0 < cnt; j0 += NG) {
        const int j = j0 + gi, pos = start + j;
        const bool valid = j < cnt && !(pos >= in_len && pos < max_in);
        float dot = 0.f;
        if (valid) {
            const uint4 kv = *reinterpret_cast<const uint4*>(pos == tlen ? &s_k[li * 8] : &sK[(size_t)j * DH + li * 8]);
            const __half2* kh = reinterpret_cast<const __half2*>(&kv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(kh[i]);
                dot = fmaf(q[2 * i], f.x, dot);
                dot = fmaf(q[2 * i + 1], f.y, dot);
            }
        }
#pragma unroll
        for (int o = LPR / 2; o > 0; o >>= 1) dot += __shfl_xor_sync(0xffffffffu, dot, o);
        if (li == 0 && j < cnt) s_sc[j] = valid ? dot * p.inv_sqrt_dh : -INFINITY;
    }
    __syncthreads();
    // ---- softmax statistics of this split (cnt <= 64 <= threads)
    const float scv = tid < cnt ? s_sc[tid] : -INFINITY;
    const float mx = block_max(scv, s_red);
    const float ev = (scv == -INFINITY) ? 0.f : __expf(scv - mx);
    const float sum = block_sum(ev, s_red);
    if (tid < cnt) s_sc[tid] = ev;
    __syncthreads();
    // ---- P.V
    float acc[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = 0.f;
    if (STAGE_V) {
        for (int j = gi; j < cnt; j += NG) {
            const float pr = s_sc[j];
            if (pr == 0.f) continue;
            const uint4 vv = *reinterpret_cast<const uint4*>(start + j == tlen ? &s_v[li * 8] : &sV[(size_t)j * DH + li * 8]);
            const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
            for (int i = 0; i < 4; ++i) {
                const float2 f = __half22float2(vh[i]);
                acc[2 * i] = fmaf(pr, f.x, acc[2 * i]);
                acc[2 * i + 1] = fmaf(pr, f.y, acc[2 * i + 1]);
            }
        }
    } else {
        constexpr int UNR = 8;                // V rows of one thread requested together (L2 hits after the prefetch)
        for (int j0 = gi; j0 < cnt; j0 += NG * UNR) {
            uint4 vv[UNR];
            float pr[UNR];
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const int j = j0 + u * NG;
                pr[u] = j < cnt ? s_sc[j] : 0.f;
                vv[u] = make_uint4(0, 0, 0, 0);
                if (pr[u] != 0.f) {
                    if (start + j == tlen) vv[u] = *reinterpret_cast<const uint4*>(&s_v[li * 8]);
                    else vv[u] = ld_stream_16(vc + (size_t)(start + j) * DH + li * 8);
                }
            }
#pragma unroll
            for (int u = 0; u < UNR; ++u) {
                const __half2* vh = reinterpret_cast<const __half2*>(&vv[u]);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(vh[i]);
                    acc[2 * i] = fmaf(pr[u], f.x, acc[2 * i]);
                    acc[2 * i + 1] = fmaf(pr[u], f.y, acc[2 * i + 1]);
                }
            }
        }
    }
    // (lite: s_out aliases the K tile -- every thread passed the barrier after the scores, nobody reads K any more)
#pragma unroll
    for (int i = 0; i < 8; ++i) s_out[gi][li * 8 + i] = acc[i];
    __syncthreads();

    __half* ctx = static_cast<__half*>(p.ctx) + (size_t)b * H * DH + h * DH;
    float* part = p.partial + ((size_t)(b * H + h) * p.splits + split) * (DH + 2);
    for (int d = tid; d < DH; d += TH) {
        float o = 0.f;
#pragma unroll
        for (int g = 0; g < NG; ++g) o += s_out[g][d];
        if (nact == 1) ctx[d] = __float2half_rn(o * (1.f / (sum + 1e-6f)));
        else part[d] = o;
    }
    if (nact == 1) {
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t2, cnt, 0);
        return;
    }
    if (tid == 0) {
        part[DH] = mx;
        part[DH + 1] = sum;
    }
    __threadfence();
    __syncthreads();
    if (tid == 0) {
        const int old = atomicAdd(&p.counters[b * H + h], 1);
        s_flag = (old == nact - 1);
    }
    __syncthreads();
    if (!s_flag) {
        if (tid == 0) trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t2, cnt, 0);
        return;
    }
    __threadfence();
    // Merge of the nact partials by the last arriver, in two load rounds instead of a chain of them (it is the tail of the layer's
    // critical path: 6 us before, under the FFN2 weight stream every dependent round costs 2-3 us): round 1 fetches every
    // split's (max, sum) at once -> weights exp(m_s - M) in shared memory and the normaliser (fixed reduction tree: the result
    // does not depend on arrival order); round 2 requests up to 24 partial rows per dimension before the first is used.
    const float* all = p.partial + (size_t)(b * H + h) * p.splits * (DH + 2);
    float* s_w = reinterpret_cast<float*>(&s_out[0][0]);      // nact weights (nact <= max_len / 64 <= NG * DH)
    float M = -INFINITY;
    for (int s2 = tid; s2 < nact; s2 += TH) M = fmaxf(M, __ldcg(&all[s2 * (DH + 2) + DH]));
    M = block_max(M, s_red);
    float Lp = 0.f;
    for (int s2 = tid; s2 < nact; s2 += TH) {
        const float mi = __ldcg(&all[s2 * (DH + 2) + DH]);
        const float wgt = (mi == -INFINITY) ? 0.f : __expf(mi - M);
        s_w[s2] = wgt;
        Lp = fmaf(__ldcg(&all[s2 * (DH + 2) + DH + 1]), wgt, Lp);
    }
    const float L = block_sum(Lp, s_red);
    __syncthreads();
    constexpr int MB = 24;
    for (int d = tid; d < DH; d += TH) {
        float O0 = 0.f, O1 = 0.f, O2 = 0.f, O3 = 0.f;
        for (int s0 = 0; s0 < nact; s0 += MB) {
            float a[MB];
#pragma unroll
            for (int u = 0; u < MB; ++u) a[u] = (s0 + u < nact) ? __ldcg(&all[(s0 + u) * (DH + 2) + d]) : 0.f;
#pragma unroll
            for (int u = 0; u < MB; u += 4) {
                O0 = fmaf(a[u], (s0 + u < nact) ? s_w[s0 + u] : 0.f, O0);
                O1 = fmaf(a[u + 1], (s0 + u + 1 < nact) ? s_w[s0 + u + 1] : 0.f, O1);
                O2 = fmaf(a[u + 2], (s0 + u + 2 < nact) ? s_w[s0 + u + 2] : 0.f, O2);
                O3 = fmaf(a[u + 3], (s0 + u + 3 < nact) ? s_w[s0 + u + 3] : 0.f, O3);
            }
        }
        ctx[d] = __float2half_rn(((O0 + O1) + (O2 + O3)) * (1.f / (L + 1e-6f)));
    }
    if (tid == 0) {
        p.counters[b * H + h] = 0;
        trc_emit(TRC_MMHA, trc_t0, trc_t1, trc_t2, cnt, 1);
    }
}

// L2 prefetch of the cache rows one decode-attention launch will read (every valid slot < seq_len of every sequence and head),
// launched on a side stream at the START of the layer: the rows travel while the QKV GEMM streams its weights, and the
// attention kernel -- a chain of dependent load rounds -- then runs on L2 hits.  evict_last so the weight stream (evict_first)
// does not push them out again.  Reads request state only (constant within a step); touches no data.
__global__ void __launch_bounds__(256) kv_prefetch_kernel(const __half* __restrict__ k_cache, const __half* __restrict__ v_cache,
                                                          const int32_t* __restrict__ seq_len, const int32_t* __restrict__ input_len,
                                                          const uint8_t* __restrict__ finished, int heads, int dh, int max_len, int max_in)
{
    const int h = blockIdx.x, b = blockIdx.y;
    if (finished != nullptr && finished[b]) return;
    const int tlen = seq_len[b], in_len = input_len[b];
    const size_t base = ((size_t)b * heads + h) * (size_t)max_len * dh;
    const int lines = (dh * 2) / 128;                         // 128-byte lines per row
    const int per_row = 2 * lines;                            // K and V
    for (int i = threadIdx.x; i < tlen * per_row; i += blockDim.x) {
        const int pos = i / per_row, r = i % per_row;
        if (pos >= in_len && pos < max_in) continue;
        const __half* src = (r < lines ? k_cache : v_cache) + base + (size_t)pos * dh;
        asm volatile("prefetch.global.L2::evict_last [%0];" ::"l"(reinterpret_cast<const char*>(src) + (r % lines) * 128));
    }
}

FTCF_TRACE_INSTALLER(trace_install_attention)

// ---------------------------------------------------------------- prefill: bias + rotary + scatter
template <int DH>
__global__ void __launch_bounds__(DH)
prefill_qkv_rotary_scatter_kernel(const __half* __restrict__ qkv, const __half* __restrict__ bias, __half* __restrict__ q_out,
                                  __half* __restrict__ k_cache, __half* __restrict__ v_cache,
                                  const int32_t* __restrict__ tok_batch, const int32_t* __restrict__ tok_pos, int H, int rot,
                                  int max_len)
{
    const int tkn = blockIdx.x, h = blockIdx.y, d = threadIdx.x;
    const int b = tok_batch[tkn], pos = tok_pos[tkn];
    const __half* row = qkv + (size_t)tkn * 3 * H * DH;
    const int qi = h * DH + d, ki = H * DH + qi, vi = 2 * H * DH + qi;
    __half q = row[qi], k = row[ki], v = row[vi];
    if (bias) {
        q = __hadd(q, bias[qi]);
        k = __hadd(k, bias[ki]);
        v = __hadd(v, bias[vi]);
    }
    if (d < rot) {
        const int dp = d < (rot >> 1) ? d + (rot >> 1) : d - (rot >> 1);
        __half qp = row[h * DH + dp], kp = row[H * DH + h * DH + dp];
        if (bias) {
            qp = __hadd(qp, bias[h * DH + dp]);
            kp = __hadd(kp, bias[H * DH + h * DH + dp]);
        }
        q = rotary_neox(q, qp, d, rot, pos);
        k = rotary_neox(k, kp, d, rot, pos);
    }
    q_out[((size_t)tkn * H + h) * DH + d] = q;
    const size_t ci = (((size_t)b * H + h) * max_len + pos) * DH + d;
    k_cache[ci] = k;
    v_cache[ci] = v;
}


// ---------------------------------------------------------------- prefill: causal attention on tensor cores
// FlashAttention-2 form with mma.sync.m16n8k16 (fp16 in, fp32 accumulate): a CTA owns 64 query rows of one (sequence, head),
// each of its 4 warps 16 rows with the Q fragments in registers; K / V tiles of 64 keys go through shared memory (row pitch
// DH + 8 halves: ldmatrix conflict-free), S = Q.K^T and O += P.V run on the tensor cores, the online softmax in registers
// (quad shuffles); P is rounded to fp16 before P.V as the reference's qk_buf is (GptContextAttentionLayer.cc:231-249).
// Key tiles beyond the causal diagonal are never loaded.  Replaces the CUDA-core kernel below for dh = 64 / 128 (it measured
// 0.82 ms per layer at S = 1024: 13 TFLOP/s).
__device__ __forceinline__ void ldmatrix_x4(uint32_t (&r)[4], const void* p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void ldmatrix_x4_trans(uint32_t (&r)[4], const void* p)
{
    const uint32_t a = (uint32_t)__cvta_generic_to_shared(p);
    asm volatile("ldmatrix.sync.aligned.m8n8.x4.trans.shared.b16 {%0,%1,%2,%3}, [%4];" : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]) : "r"(a));
}
__device__ __forceinline__ void mma_f16_16816(float (&d)[4], const uint32_t (&a)[4], uint32_t b0, uint32_t b1)
{
    mma_16816(d, a[0], a[1], a[2], a[3], b0, b1);
}

template <int DH>
__global__ void __launch_bounds__(128)
prefill_attention_mma_kernel(const __half* __restrict__ q, const __half* __restrict__ k_cache, const __half* __restrict__ v_cache,
                             __half* __restrict__ ctx, const int32_t* __restrict__ seq_offsets, int H, int max_len, float scale)
{
    constexpr int QT = 64, KT = 64, PITCH = DH + 8, KS = DH / 16, DT = DH / 8;
    __shared__ __align__(16) __half s_k[KT][PITCH];
    __shared__ __align__(16) __half s_v[KT][PITCH];

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const int off = seq_offsets[b], len = seq_offsets[b + 1] - off;
    if (q0 >= len) return;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31, g = lane >> 2, t = lane & 3;
    const int r0 = q0 + warp * 16 + g, r1 = r0 + 8;          // the two query rows this thread holds
    const int r0c = min(r0, len - 1), r1c = min(r1, len - 1);

    // ---- Q fragments (A operand, row-major): a0 (row g, k 2t..), a1 (row g+8), a2 (row g, k 8+2t..), a3 (row g+8)
    uint32_t qa[KS][4];
    {
        const __half* q0p = q + ((size_t)(off + r0c) * H + h) * DH;
        const __half* q1p = q + ((size_t)(off + r1c) * H + h) * DH;
#pragma unroll
        for (int kk = 0; kk < KS; ++kk) {
            qa[kk][0] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16 + 2 * t);
            qa[kk][1] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16 + 2 * t);
            qa[kk][2] = *reinterpret_cast<const uint32_t*>(q0p + kk * 16 + 8 + 2 * t);
            qa[kk][3] = *reinterpret_cast<const uint32_t*>(q1p + kk * 16 + 8 + 2 * t);
        }
    }
    float o[DT][4];
#pragma unroll
    for (int i = 0; i < DT; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) o[i][j] = 0.f;
    float m0 = -INFINITY, m1 = -INFINITY, l0 = 0.f, l1 = 0.f;

    const __half* kb = k_cache + ((size_t)b * H + h) * (size_t)max_len * DH;
    const __half* vb = v_cache + ((size_t)b * H + h) * (size_t)max_len * DH;
    const int kmax = min(q0 + QT, len);                        // keys this CTA can need: [0, kmax)
    for (int k0 = 0; k0 < kmax; k0 += KT) {
        __syncthreads();
        for (int v = tid; v < KT * DH / 8; v += 128) {
            const int r = v / (DH / 8), c = v % (DH / 8);
            uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
            if (k0 + r < kmax) {
                kk = *reinterpret_cast<const uint4*>(kb + (size_t)(k0 + r) * DH + c * 8);
                vv = *reinterpret_cast<const uint4*>(vb + (size_t)(k0 + r) * DH + c * 8);
            }
            *reinterpret_cast<uint4*>(&s_k[r][c * 8]) = kk;
            *reinterpret_cast<uint4*>(&s_v[r][c * 8]) = vv;
        }
        __syncthreads();
        if (k0 > q0 + warp * 16 + 15) continue;                // this warp's rows are all above the tile (causal)

        // ---- S = Q.K^T for 16 rows x 64 keys
        float sacc[KT / 8][4];
#pragma unroll
        for (int j = 0; j < KT / 8; ++j)
#pragma unroll
            for (int e = 0; e < 4; ++e) sacc[j][e] = 0.f;
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
#pragma unroll
            for (int kk = 0; kk < KS; kk += 2) {
                uint32_t bf[4];   // (keys 8j.., d 16kk..+7), (.., +8..15), (.., +16..23), (.., +24..31)
                ldmatrix_x4(bf, &s_k[j * 8 + (lane & 7)][kk * 16 + (lane >> 3) * 8]);
                mma_f16_16816(sacc[j], qa[kk], bf[0], bf[1]);
                mma_f16_16816(sacc[j], qa[kk + 1], bf[2], bf[3]);
            }
        }
        // ---- scale, causal mask, online softmax (rows r0 and r1)
        float mx0 = -INFINITY, mx1 = -INFINITY;
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
            const int c0 = k0 + j * 8 + 2 * t;
            sacc[j][0] = (c0 <= r0 && c0 < kmax) ? sacc[j][0] * scale : -INFINITY;
            sacc[j][1] = (c0 + 1 <= r0 && c0 + 1 < kmax) ? sacc[j][1] * scale : -INFINITY;
            sacc[j][2] = (c0 <= r1 && c0 < kmax) ? sacc[j][2] * scale : -INFINITY;
            sacc[j][3] = (c0 + 1 <= r1 && c0 + 1 < kmax) ? sacc[j][3] * scale : -INFINITY;
            mx0 = fmaxf(mx0, fmaxf(sacc[j][0], sacc[j][1]));
            mx1 = fmaxf(mx1, fmaxf(sacc[j][2], sacc[j][3]));
        }
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 1));
        mx0 = fmaxf(mx0, __shfl_xor_sync(0xffffffffu, mx0, 2));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 1));
        mx1 = fmaxf(mx1, __shfl_xor_sync(0xffffffffu, mx1, 2));
        const float mn0 = fmaxf(m0, mx0), mn1 = fmaxf(m1, mx1);
        const float c0f = (mn0 == -INFINITY) ? 1.f : __expf(m0 - mn0);     // exp(-inf - x) = 0 on the first tile
        const float c1f = (mn1 == -INFINITY) ? 1.f : __expf(m1 - mn1);
        m0 = mn0;
        m1 = mn1;
        float ps0 = 0.f, ps1 = 0.f;
        uint32_t pa[KT / 16][4];                                             // P as A fragments (fp16)
#pragma unroll
        for (int j = 0; j < KT / 8; ++j) {
            const float p00 = (sacc[j][0] == -INFINITY) ? 0.f : __expf(sacc[j][0] - mn0);
            const float p01 = (sacc[j][1] == -INFINITY) ? 0.f : __expf(sacc[j][1] - mn0);
            const float p10 = (sacc[j][2] == -INFINITY) ? 0.f : __expf(sacc[j][2] - mn1);
            const float p11 = (sacc[j][3] == -INFINITY) ? 0.f : __expf(sacc[j][3] - mn1);
            ps0 += p00 + p01;
            ps1 += p10 + p11;
            pa[j >> 1][(j & 1) * 2 + 0] = f2_to_h2(p00, p01);
            pa[j >> 1][(j & 1) * 2 + 1] = f2_to_h2(p10, p11);
        }
        l0 = l0 * c0f + ps0;
        l1 = l1 * c1f + ps1;
#pragma unroll
        for (int i = 0; i < DT; ++i) {
            o[i][0] *= c0f;
            o[i][1] *= c0f;
            o[i][2] *= c1f;
            o[i][3] *= c1f;
        }
        // ---- O += P.V
#pragma unroll
        for (int ks = 0; ks < KT / 16; ++ks) {
#pragma unroll
            for (int i = 0; i < DT; i += 2) {
                uint32_t bf[4];   // (keys 16ks..+7, d 8i..), (keys +8..15, d 8i..), (keys ..+7, d 8i+8..), (keys +8..15, d 8i+8..)
                ldmatrix_x4_trans(bf, &s_v[ks * 16 + (lane & 7) + ((lane >> 3) & 1) * 8][i * 8 + (lane >> 4) * 8]);
                mma_f16_16816(o[i], pa[ks], bf[0], bf[1]);
                mma_f16_16816(o[i + 1], pa[ks], bf[2], bf[3]);
            }
        }
    }
    // ---- row sums across the quad, normalise, store
    l0 += __shfl_xor_sync(0xffffffffu, l0, 1);
    l0 += __shfl_xor_sync(0xffffffffu, l0, 2);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 1);
    l1 += __shfl_xor_sync(0xffffffffu, l1, 2);
    const float inv0 = 1.f / l0, inv1 = 1.f / l1;
    __half* o0p = ctx + (size_t)(off + r0c) * H * DH + h * DH;
    __half* o1p = ctx + (size_t)(off + r1c) * H * DH + h * DH;
#pragma unroll
    for (int i = 0; i < DT; ++i) {
        if (r0 < len) *reinterpret_cast<uint32_t*>(o0p + i * 8 + 2 * t) = f2_to_h2(o[i][0] * inv0, o[i][1] * inv0);
        if (r1 < len) *reinterpret_cast<uint32_t*>(o1p + i * 8 + 2 * t) = f2_to_h2(o[i][2] * inv1, o[i][3] * inv1);
    }
}

// ---------------------------------------------------------------- prefill: causal attention (CUDA-core flash form)
// 4 lanes per query row, 32 query rows per CTA, K/V tiles of 32 keys staged in shared memory, online softmax.
template <int DH>
__global__ void __launch_bounds__(128)
prefill_attention_kernel(const __half* __restrict__ q, const __half* __restrict__ k_cache, const __half* __restrict__ v_cache,
                         __half* __restrict__ ctx, const int32_t* __restrict__ seq_offsets, int H, int max_len, float scale)
{
    constexpr int QT = 32, KT = 32, DPL = DH / 4, NV = DPL / 8;   // dims per lane, 16-byte vectors per lane
    __shared__ __align__(16) __half s_k[KT][DH];
    __shared__ __align__(16) __half s_v[KT][DH];

    const int b = blockIdx.z, h = blockIdx.y, q0 = blockIdx.x * QT;
    const int off = seq_offsets[b], len = seq_offsets[b + 1] - off;
    if (q0 >= len) return;
    const int tid = threadIdx.x, li = tid & 3, qi = q0 + (tid >> 2);
    const bool q_ok = qi < len;

    float qr[DPL], o[DPL];
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        uint4 v = make_uint4(0, 0, 0, 0);
        if (q_ok) v = *reinterpret_cast<const uint4*>(q + ((size_t)(off + qi) * H + h) * DH + c * 32 + li * 8);
        const __half2* vh = reinterpret_cast<const __half2*>(&v);
#pragma unroll
        for (int i = 0; i < 4; ++i) {
            const float2 f = __half22float2(vh[i]);
            qr[c * 8 + 2 * i] = f.x;
            qr[c * 8 + 2 * i + 1] = f.y;
        }
    }
#pragma unroll
    for (int i = 0; i < DPL; ++i) o[i] = 0.f;
    float mrun = -INFINITY, lrun = 0.f;

    const __half* kb = k_cache + ((size_t)b * H + h) * (size_t)max_len * DH;
    const __half* vb = v_cache + ((size_t)b * H + h) * (size_t)max_len * DH;
    const int kmax = min(q0 + QT, len);   // keys needed by this CTA: [0, kmax)
    for (int k0 = 0; k0 < kmax; k0 += KT) {
        __syncthreads();
        for (int v = tid; v < KT * DH / 8; v += 128) {
            const int r = v / (DH / 8), c = v % (DH / 8);
            uint4 kk = make_uint4(0, 0, 0, 0), vv = kk;
            if (k0 + r < kmax) {
                kk = *reinterpret_cast<const uint4*>(kb + (size_t)(k0 + r) * DH + c * 8);
                vv = *reinterpret_cast<const uint4*>(vb + (size_t)(k0 + r) * DH + c * 8);
            }
            *reinterpret_cast<uint4*>(&s_k[r][c * 8]) = kk;
            *reinterpret_cast<uint4*>(&s_v[r][c * 8]) = vv;
        }
        __syncthreads();
        const int jend = min(KT, kmax - k0);
        for (int j = 0; j < jend; ++j) {
            float dot = 0.f;
#pragma unroll
            for (int c = 0; c < NV; ++c) {
                const uint4 kk = *reinterpret_cast<const uint4*>(&s_k[j][c * 32 + li * 8]);
                const __half2* kh = reinterpret_cast<const __half2*>(&kk);
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const float2 f = __half22float2(kh[i]);
                    dot = fmaf(qr[c * 8 + 2 * i], f.x, dot);
                    dot = fmaf(qr[c * 8 + 2 * i + 1], f.y, dot);
                }
            }
            dot += __shfl_xor_sync(0xffffffffu, dot, 1);
            dot += __shfl_xor_sync(0xffffffffu, dot, 2);
            if (k0 + j <= qi) {                       // causal (uniform inside the 4-lane group)
                const float sc = dot * scale;
                if (sc > mrun) {
                    const float corr = __expf(mrun - sc);   // exp(-inf) = 0 on the first key
                    lrun *= corr;
#pragma unroll
                    for (int i = 0; i < DPL; ++i) o[i] *= corr;
                    mrun = sc;
                }
                const float pr = __expf(sc - mrun);
                lrun += pr;
#pragma unroll
                for (int c = 0; c < NV; ++c) {
                    const uint4 vv = *reinterpret_cast<const uint4*>(&s_v[j][c * 32 + li * 8]);
                    const __half2* vh = reinterpret_cast<const __half2*>(&vv);
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float2 f = __half22float2(vh[i]);
                        o[c * 8 + 2 * i] = fmaf(pr, f.x, o[c * 8 + 2 * i]);
                        o[c * 8 + 2 * i + 1] = fmaf(pr, f.y, o[c * 8 + 2 * i + 1]);
                    }
                }
            }
        }
    }
    if (!q_ok) return;
    const float inv = 1.f / lrun;
#pragma unroll
    for (int c = 0; c < NV; ++c) {
        uint4 out;
        uint32_t* ow = reinterpret_cast<uint32_t*>(&out);
#pragma unroll
        for (int i = 0; i < 4; ++i) ow[i] = f2_to_h2(o[c * 8 + 2 * i] * inv, o[c * 8 + 2 * i + 1] * inv);
        *reinterpret_cast<uint4*>(ctx + (size_t)(off + qi) * H * DH + h * DH + c * 32 + li * 8) = out;
    }
}

}  // namespace ftcf

using namespace ftcf;

static bool mmha_bulk_applies(int batch, int heads, int dh)
{
    // small BATCH, not just few (row, head) pairs: with tensor parallelism a batch of 32 has only 160-320 pairs per rank, but it is
    // a bandwidth problem (hundreds of MB of cache per layer) that belongs to the streaming kernels below
    return g_mmha_bulk.load() != 0 && batch <= 8 && batch * heads <= 320 && (dh == 64 || dh == 128);
}

extern "C" int ftcf_mmha_choose_splits(int batch, int heads, int max_len)
{
    const int ctas = batch * heads;
    if (g_mmha_splits.load() == 0 && mmha_bulk_applies(batch, heads, 128)) return ceil_div(max_len, MMHA_BULK_KEYS);
    int splits = ceil_div(148 * 4, ctas > 0 ? ctas : 1);
    if (g_mmha_splits.load() > 0) splits = g_mmha_splits.load();   // experiment hook
    const int by_len = ceil_div(max_len, 128);          // at least 128 keys per split
    if (splits > by_len) splits = by_len;
    const int need = ceil_div(max_len, MMHA_MAX_CHUNK); // at most MMHA_MAX_CHUNK keys per split
    if (splits < need) splits = need;
    if (splits < 1) splits = 1;
    return splits;
}

extern "C" int ftcf_mmha_prefetch_cache(const ftcf_mmha_params* p, void* stream)
{
    FTCF_REQUIRE(p != nullptr && p->batch > 0 && p->heads > 0, FTCF_ERR_INVALID, "mmha prefetch: bad params");
    kv_prefetch_kernel<<<dim3(p->heads, p->batch), 256, 0, as_stream(stream)>>>(
        static_cast<const __half*>(p->k_cache), static_cast<const __half*>(p->v_cache), p->seq_len, p->input_len, p->finished, p->heads,
        p->dh, p->max_len, p->max_input_len);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_mmha_decode(const ftcf_mmha_params* p, void* stream)
{
    FTCF_REQUIRE(p != nullptr, FTCF_ERR_INVALID, "mmha: null params");
    FTCF_REQUIRE(p->batch > 0 && p->heads > 0 && p->splits >= 1, FTCF_ERR_INVALID, "mmha: bad sizes");
    FTCF_REQUIRE(ceil_div(p->max_len, p->splits) <= MMHA_MAX_CHUNK, FTCF_ERR_INVALID,
                 "mmha: max_len %d needs at least %d splits", p->max_len, ceil_div(p->max_len, MMHA_MAX_CHUNK));
    FTCF_REQUIRE(p->splits == 1 || (p->partial != nullptr && p->counters != nullptr), FTCF_ERR_INVALID,
                 "mmha: split-KV needs scratch");
    FTCF_REQUIRE(p->rotary_dim % 2 == 0 && p->rotary_dim <= p->dh, FTCF_ERR_INVALID, "mmha: rotary_dim %d", p->rotary_dim);
    MmhaP mp{*p, g_mmha_prefetch.load()};
    const dim3 grid(p->heads, p->batch, p->splits);
    cudaError_t lerr = cudaSuccess;
    // one pass where the launch is a single wave of CTAs (measured, 13B decode step: batch 8 -5 %, batch 1 equal, batch 32 +3 %)
    const bool onepass = g_mmha_onepass.load() == 2 || (g_mmha_onepass.load() != 0 && p->batch * p->heads <= 148 * 4);
    const bool pdl = g_mmha_pdl.load() != 0;
    if (p->cache_indir != nullptr) {
        // beam search: rows gather their keys through the cache indirection -- the two-pass kernel is the one that does that
        FTCF_REQUIRE(p->beam_width > 1 && p->batch % p->beam_width == 0, FTCF_ERR_INVALID, "mmha: cache_indir with beam_width %d, %d rows",
                     p->beam_width, p->batch);
        switch (p->dh) {
            case 64: lerr = launch_pdl_if(pdl, mmha_decode_kernel<64, true>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
            case 128: lerr = launch_pdl_if(pdl, mmha_decode_kernel<128, true>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
            case 256: lerr = launch_pdl_if(pdl, mmha_decode_kernel<256, true>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
            default: FTCF_REQUIRE(false, FTCF_ERR_UNSUPPORTED, "mmha: size_per_head %d (supported: 64, 128, 256)", p->dh);
        }
        FTCF_REQUIRE(lerr == cudaSuccess, FTCF_ERR_CUDA, "mmha (beams) launch failed: %s", cudaGetErrorString(lerr));
        FTCF_LAUNCH_CHECK();
        return FTCF_OK;
    }
    // small batch, enough splits that no CTA gets more than 64 keys: the bulk-staged kernel
    if (mmha_bulk_applies(p->batch, p->heads, p->dh) && ceil_div(p->max_len, p->splits) <= MMHA_BULK_KEYS) {
        // The small CTAs pay off only where the 128-thread kernel would need more than one wave beside the GEMMs (~440 CTAs).
        // Measured in the 13B step (profiles/r2_decode_experiments.txt): one GPU, batch 1 (960 CTAs) -2.3 %; batch 2 +0.6 %;
        // two GPUs, batch 1 (480 CTAs) +5 %.
        const bool lite = g_mmha_lite.load() != 0 && p->batch * p->heads <= 48 && (long long)p->batch * p->heads * p->splits > 600;
        const size_t smem = (size_t)(lite ? 1 : 2) * MMHA_BULK_KEYS * p->dh * sizeof(__half);
        if (p->dh == 128) {
            static std::atomic<int> cfg128{0};
            if (!cfg128.exchange(1))
                FTCF_CUDA_CHECK(cudaFuncSetAttribute(mmha_decode_bulk_kernel<128, MMHA_THREADS, true>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                                     (int)(2 * MMHA_BULK_KEYS * 128 * sizeof(__half))));
            lerr = lite ? launch_pdl(mmha_decode_bulk_kernel<128, 64, false>, grid, dim3(64), smem, as_stream(stream), mp)
                        : launch_pdl(mmha_decode_bulk_kernel<128, MMHA_THREADS, true>, grid, dim3(MMHA_THREADS), smem, as_stream(stream), mp);
        } else {
            lerr = lite ? launch_pdl(mmha_decode_bulk_kernel<64, 64, false>, grid, dim3(64), smem, as_stream(stream), mp)
                        : launch_pdl(mmha_decode_bulk_kernel<64, MMHA_THREADS, true>, grid, dim3(MMHA_THREADS), smem, as_stream(stream), mp);
        }
        FTCF_REQUIRE(lerr == cudaSuccess, FTCF_ERR_CUDA, "mmha (bulk) launch failed: %s", cudaGetErrorString(lerr));
        FTCF_LAUNCH_CHECK();
        return FTCF_OK;
    }
    switch (p->dh) {
        case 64: lerr = onepass ? launch_pdl_if(pdl, mmha_decode_onepass_kernel<64>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp)
                                : launch_pdl_if(pdl, mmha_decode_kernel<64>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
        case 128: lerr = onepass ? launch_pdl_if(pdl, mmha_decode_onepass_kernel<128>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp)
                                 : launch_pdl_if(pdl, mmha_decode_kernel<128>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
        case 256: lerr = launch_pdl_if(pdl, mmha_decode_kernel<256>, grid, dim3(MMHA_THREADS), 0, as_stream(stream), mp); break;
        default: FTCF_REQUIRE(false, FTCF_ERR_UNSUPPORTED, "mmha: size_per_head %d (supported: 64, 128, 256)", p->dh);
    }
    FTCF_REQUIRE(lerr == cudaSuccess, FTCF_ERR_CUDA, "mmha launch failed: %s", cudaGetErrorString(lerr));
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_prefill_qkv_rotary_scatter(const void* qkv, const void* qkv_bias, void* q_out, void* k_cache, void* v_cache,
                                               const int32_t* tok_batch, const int32_t* tok_pos, int tokens, int heads, int dh,
                                               int rotary_dim, int max_len, void* stream)
{
    FTCF_REQUIRE(tokens > 0 && heads > 0, FTCF_ERR_INVALID, "prefill scatter: empty");
    FTCF_REQUIRE(rotary_dim % 2 == 0 && rotary_dim <= dh, FTCF_ERR_INVALID, "prefill scatter: rotary_dim %d", rotary_dim);
    const dim3 grid(tokens, heads);
#define FTCF_PS(DH_)                                                                                                   \
    prefill_qkv_rotary_scatter_kernel<DH_><<<grid, DH_, 0, as_stream(stream)>>>(                                       \
        static_cast<const __half*>(qkv), static_cast<const __half*>(qkv_bias), static_cast<__half*>(q_out),            \
        static_cast<__half*>(k_cache), static_cast<__half*>(v_cache), tok_batch, tok_pos, heads, rotary_dim, max_len)
    switch (dh) {
        case 64: FTCF_PS(64); break;
        case 128: FTCF_PS(128); break;
        case 256: FTCF_PS(256); break;
        default: FTCF_REQUIRE(false, FTCF_ERR_UNSUPPORTED, "prefill scatter: size_per_head %d", dh);
    }
#undef FTCF_PS
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_prefill_attention(const void* q, const void* k_cache, const void* v_cache, void* ctx,
                                      const int32_t* seq_offsets, int batch, int max_seq, int heads, int dh, int max_len,
                                      float scale, void* stream)
{
    FTCF_REQUIRE(batch > 0 && max_seq > 0 && heads > 0, FTCF_ERR_INVALID, "prefill attention: empty");
    const bool mma = g_prefill_mma.load() != 0;
    const dim3 grid(ceil_div(max_seq, mma ? 64 : 32), heads, batch);
#define FTCF_PA(DH_)                                                                                                   \
    if (mma)                                                                                                           \
        prefill_attention_mma_kernel<DH_><<<grid, 128, 0, as_stream(stream)>>>(                                        \
            static_cast<const __half*>(q), static_cast<const __half*>(k_cache), static_cast<const __half*>(v_cache),   \
            static_cast<__half*>(ctx), seq_offsets, heads, max_len, scale);                                            \
    else                                                                                                               \
        prefill_attention_kernel<DH_><<<grid, 128, 0, as_stream(stream)>>>(                                            \
            static_cast<const __half*>(q), static_cast<const __half*>(k_cache), static_cast<const __half*>(v_cache),   \
            static_cast<__half*>(ctx), seq_offsets, heads, max_len, scale)
    switch (dh) {
        case 64: FTCF_PA(64); break;
        case 128: FTCF_PA(128); break;
        default: FTCF_REQUIRE(false, FTCF_ERR_UNSUPPORTED, "prefill attention: size_per_head %d (supported: 64, 128)", dh);
    }
#undef FTCF_PA
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
