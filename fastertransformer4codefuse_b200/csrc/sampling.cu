// Logit post-processing, top-k sampling, stop criteria and the output gather of the decode loop (fp32 logits).
//
// Reference behaviour restated (not ported) -- paths relative to src/fastertransformer:
//   * order of operations: layers/DynamicDecodeLayer.cc:192-495, layers/sampling_layers/BaseSamplingLayer.cc:255-357
//   * optional_last_tokens (first generated step only): kernels/select_optional_last_tokens.cu:74-83
//   * temperature x 1/(T + 1e-6), padded vocab -> -FLT_MAX: kernels/sampling_penalty_kernels.cu:115-143
//   * multiplicative repetition penalty, once per distinct id, pad gap skipped: sampling_penalty_kernels.cu:366-425
//   * end mask for finished rows: kernels/sampling_topk_kernels.cu:68-93; softmax exp(x-max)/(sum+1e-6):
//     kernels/sampling_topp_kernels.cu:1296-1345
//   * two-stage top-k (8 vocabulary slices per row, block-strided), candidate walk with curand_uniform * p * sum:
//     sampling_topk_kernels.cu:131-312, layers/sampling_layers/TopKSamplingLayer.cu:189-265
//   * tie order: a thread keeps the first maximum it meets, equal maxima across threads go to the higher thread index
//     (kernels/reduce_kernel_utils.cuh:325-348 under cub::BlockReduce) -- restated here as one total order
//     (value desc, index % BLOCK desc, index asc) so the result does not depend on the shape of our reduction tree
//   * stop words on the time-major id buffer: kernels/stop_criteria_kernels.cu:24-81
//   * gatherTree (beam 1) removing the pad gap: kernels/decoding_kernels.cu:452-580
// The reference's per-token host spin-wait (stop_criteria_kernels.cu:135-156) is replaced by a finished counter
// written to mapped pinned memory that the host polls without stalling the stream.
#include <curand_kernel.h>
#include <float.h>

#include "common.cuh"

namespace ftcf {

constexpr int BLOCKS_PER_ROW = 8;

struct Cand {
    float v;
    int idx;   // -1: none
};

template <int BS>
__device__ __forceinline__ bool cand_better(const Cand& a, const Cand& b)
{
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    if (a.v != b.v) return a.v > b.v;
    const int ta = a.idx % BS, tb = b.idx % BS;
    if (ta != tb) return ta > tb;
    return a.idx < b.idx;
}

template <int BS>
__device__ __forceinline__ Cand block_argmax(Cand c, Cand* s_c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.v = __shfl_xor_sync(0xffffffffu, c.v, o);
        other.idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
        if (cand_better<BS>(other, c)) c = other;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_c[warp] = c;
    __syncthreads();
    Cand r = (lane < BS / 32) ? s_c[lane] : Cand{-FLT_MAX, -1};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        Cand other;
        other.v = __shfl_xor_sync(0xffffffffu, r.v, o);
        other.idx = __shfl_xor_sync(0xffffffffu, r.idx, o);
        if (cand_better<BS>(other, r)) r = other;
    }
    return r;
}

// ---------------------------------------------------------------- logits preparation: one CTA per row
__global__ void __launch_bounds__(1024) logits_prepare_kernel(const ftcf_sampling_params p, float* __restrict__ rep_scratch)
{
    __shared__ float s_red[32];
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (tid == 0) { const unsigned long long t = trc_now(); trc_emit(TRC_SAMPLING, t, t, t, 0, 0); }   // marks the start of sampling
    const int V = p.vocab, Vp = p.vocab_padded;
    float* row = p.logits + (size_t)b * Vp;
    const int step = *p.step;
    const bool fin = p.finished[b] != 0;

    // (a) optional last tokens: only on the first generated step
    if (p.optional_last_tokens != nullptr && step == p.max_input_len) {
        const int32_t* allowed = p.optional_last_tokens + (size_t)b * p.n_last;
        float* saved = rep_scratch + (size_t)b * p.max_len;   // n_last <= max_len checked on the host
        for (int i = tid; i < p.n_last; i += nt) {
            const int id = allowed[i];
            if (id >= 0 && id < Vp) saved[i] = row[id];
        }
        __syncthreads();
        for (int v = tid; v < Vp; v += nt) row[v] = -INFINITY;
        __syncthreads();
        for (int i = tid; i < p.n_last; i += nt) {
            const int id = allowed[i];
            if (id >= 0 && id < Vp) row[id] = saved[i];
        }
        __syncthreads();
    }
    // (b) temperature
    if (p.temperature != nullptr) {
        const float inv = 1.f / (p.temperature[b] + 1e-6f);
        for (int v = tid; v < Vp; v += nt) row[v] = v < V ? row[v] * inv : -FLT_MAX;
        __syncthreads();
    }
    // (c) repetition penalty (every distinct earlier id once: all reads happen before any write)
    if (p.repetition_penalty != nullptr && step > 1) {
        const float pen = p.repetition_penalty[b];
        float* pl = rep_scratch + (size_t)b * p.max_len;
        const int in_len = p.input_len[b];
        for (int i = tid; i < step; i += nt) {
            if (i >= in_len && i < p.max_input_len) continue;
            const int id = p.output_ids[(size_t)i * p.batch + b];
            if (id < 0 || id >= Vp) continue;      // an id outside the table (the embedding lookup clamps it) penalises nothing
            const float lg = row[id];
            pl[i] = lg < 0.f ? lg * pen : lg / pen;
        }
        __syncthreads();
        for (int i = tid; i < step; i += nt) {
            if (i >= in_len && i < p.max_input_len) continue;
            const int id = p.output_ids[(size_t)i * p.batch + b];
            if (id < 0 || id >= Vp) continue;
            row[id] = pl[i];
        }
        __syncthreads();
    }
    // (d) end mask / padded vocabulary
    for (int v = tid; v < Vp; v += nt) {
        if (v >= V) row[v] = -FLT_MAX;
        else if (fin) row[v] = (v == p.end_id) ? FLT_MAX : -FLT_MAX;
    }
    __syncthreads();
    // (e) softmax: when cum_log_probs are wanted (TopKSamplingLayer.cu:236-247) and always for pure top-p rows
    //     (TopPSamplingLayer.cu:293-300)
    if (p.want_probs || p.top_k[b] == 0) {
        float mx = -FLT_MAX;
        for (int v = tid; v < Vp; v += nt) mx = fmaxf(mx, row[v]);
        mx = block_max(mx, s_red);
        float sum = 0.f;
        for (int v = tid; v < Vp; v += nt) {
            const float e = expf(row[v] - mx);
            row[v] = e;
            sum += e;
        }
        sum = block_sum(sum, s_red);
        const float denom = sum + 1e-6f;
        for (int v = tid; v < Vp; v += nt) row[v] = row[v] / denom;
    }
}

// ---------------------------------------------------------------- top-k stage 1: grid (8, B)
template <int BS>
__global__ void __launch_bounds__(BS)
topk_stage1_kernel(const float* __restrict__ logits, float* __restrict__ tmp, int* __restrict__ cand_id,
                   float* __restrict__ cand_val, const int32_t* __restrict__ top_k, const uint8_t* __restrict__ finished,
                   int Vp, int max_top_k)
{
    __shared__ Cand s_c[32];
    const int lane_blk = blockIdx.x, b = blockIdx.y, tid = threadIdx.x;
    if (finished[b]) return;
    const int k = top_k[b];
    if (k == 0) return;                       // pure top-p row: topp_sample_kernel (skip_decode, TopKSamplingLayer.cu:66-75)
    const float* row = logits + (size_t)b * Vp;
    float* trow = tmp + (size_t)b * Vp;
    for (int e = tid + lane_blk * BS; e < Vp; e += BS * BLOCKS_PER_ROW) trow[e] = row[e];
    // each thread only ever re-reads what it wrote itself: no barrier needed
    int* out_id = cand_id + ((size_t)b * BLOCKS_PER_ROW + lane_blk) * max_top_k;
    float* out_val = cand_val + ((size_t)b * BLOCKS_PER_ROW + lane_blk) * max_top_k;
    for (int ite = 0; ite < k; ++ite) {
        Cand c{-FLT_MAX, -1};
        for (int e = tid + lane_blk * BS; e < Vp; e += BS * BLOCKS_PER_ROW) {
            const float v = trow[e];
            if (v > c.v) {
                c.v = v;
                c.idx = e;
            }
        }
        const Cand w = block_argmax<BS>(c, s_c);
        if (tid == 0) {
            out_id[ite] = w.idx;
            out_val[ite] = w.idx >= 0 ? w.v : -FLT_MAX;
        }
        if (w.idx >= 0 && (w.idx % BS) == tid) trow[w.idx] = -FLT_MAX;   // the owning thread retires the winner
        __syncthreads();
    }
}

// ---------------------------------------------------------------- top-k stage 2 + sampling: one CTA per row
template <int BS>
__global__ void __launch_bounds__(BS)
topk_stage2_kernel(const ftcf_sampling_params p, int* __restrict__ cand_id, float* __restrict__ cand_val)
{
    extern __shared__ unsigned char s_dyn[];
    __shared__ Cand s_c[32];
    const int b = blockIdx.x, tid = threadIdx.x;
    const int step = *p.step;
    int32_t* out = p.output_ids + (size_t)step * p.batch + b;
    if (p.top_k[b] == 0) return;              // pure top-p row
    if (p.finished[b]) {
        if (tid == 0) *out = p.end_id;
        return;
    }
    const int k = p.top_k[b];
    const int size = k * BLOCKS_PER_ROW;
    int* s_id = reinterpret_cast<int*>(s_dyn);
    float* s_val2 = reinterpret_cast<float*>(s_dyn) + p.max_top_k;
    // candidates of lane j live at [j * max_top_k, j * max_top_k + k); logical position = j * k + i
    const int* ids = cand_id + (size_t)b * BLOCKS_PER_ROW * p.max_top_k;
    float* vals = cand_val + (size_t)b * BLOCKS_PER_ROW * p.max_top_k;
    float s_max = 0.f, s_sum = 0.f;
    for (int ite = 0; ite < k; ++ite) {
        Cand c{-FLT_MAX, -1};
        for (int pos = tid; pos < size; pos += BS) {
            const float v = vals[(pos / k) * p.max_top_k + (pos % k)];
            if (v > c.v) {
                c.v = v;
                c.idx = pos;
            }
        }
        const Cand w = block_argmax<BS>(c, s_c);
        if (w.idx < 0) {   // fewer than k valid candidates: not reachable with k <= vocab/8, keep the walk well-defined
            if (tid == 0) {
                s_id[ite] = -1;
                s_val2[ite] = 0.f;
            }
            __syncthreads();
            continue;
        }
        if (tid == 0) {
            float u = w.v;
            if (ite == 0) s_max = u;
            if (!p.want_probs) u = __expf(u - s_max);
            s_id[ite] = w.idx;
            s_val2[ite] = u;
            s_sum += u;
            vals[(w.idx / k) * p.max_top_k + (w.idx % k)] = -FLT_MAX;
        }
        __syncthreads();
    }
    if (tid == 0) {
        curandState_t* st = reinterpret_cast<curandState_t*>(p.curand_states) + b;
        float rand_num = curand_uniform(st) * p.top_p[b] * s_sum;
        int chosen = k - 1;
        for (int i = 0; i < k; ++i) {
            rand_num -= s_val2[i];
            if (rand_num <= 0.f || i == k - 1) {
                chosen = i;
                break;
            }
        }
        while (chosen > 0 && s_id[chosen] < 0) --chosen;
        const int pos = s_id[chosen];
        const int tok = pos >= 0 ? ids[(pos / k) * p.max_top_k + (pos % k)] % p.vocab_padded : p.end_id;
        *out = tok;
        if (p.cum_log_probs != nullptr && p.want_probs) p.cum_log_probs[b] += logf(s_val2[chosen]);
        p.seq_len[b] += 1;
        p.finished[b] = (tok == p.end_id) ? 1 : 0;
    }
}

// ---------------------------------------------------------------- pure top-p (nucleus) rows: one CTA per row
// Semantics of topp_beam_topk_kernel<MAX_K = 1> + the segmented descending radix sort + topp_sampling
// (kernels/sampling_topp_kernels.cu:801-1004, 1006-1160): draw r = curand_uniform * p; if the largest probability alone reaches
// p it is taken; otherwise walk the probabilities in descending order (ties in ascending id order: the sort is stable) and take
// the first id whose inclusive cumulative sum reaches r.
// B200 design: no sort.  The crossing element is found by bisection on the probability's bit pattern (non-negative floats order
// like unsigned integers): G(K) = sum of the probabilities with key >= K is evaluated by one pass over the row (400 KB, L2
// resident) per bisection step, in 2^-44 fixed point so that the sum does not depend on the order of the additions -- 31 passes,
// deterministic, O(V) memory traffic from L2 and no V-sized scratch.
__device__ __forceinline__ unsigned long long prob_fx(float pr) { return __float2ull_rz(pr * 17592186044416.f); }   // 2^44

__device__ __forceinline__ unsigned long long block_sum_u64(unsigned long long v, unsigned long long* red)
{
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nw = (blockDim.x + 31) >> 5;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if (lane == 0) red[warp] = v;
    __syncthreads();
    unsigned long long r = lane < nw ? red[lane] : 0ull;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) r += __shfl_xor_sync(0xffffffffu, r, o);
    return r;
}

__global__ void __launch_bounds__(1024) topp_sample_kernel(const ftcf_sampling_params p)
{
    __shared__ unsigned long long s_red[32];
    __shared__ Cand s_c[32];
    __shared__ float s_rand;
    __shared__ int s_pick;
    const int b = blockIdx.x, tid = threadIdx.x, nt = blockDim.x;
    if (p.top_k[b] != 0) return;
    const int step = *p.step;
    int32_t* out = p.output_ids + (size_t)step * p.batch + b;
    if (p.finished[b]) {
        if (tid == 0) *out = p.end_id;
        return;
    }
    const int Vp = p.vocab_padded;
    const float* row = p.logits + (size_t)b * Vp;      // probabilities (logits_prepare softmaxed this row)
    const float pthr = p.top_p[b];
    if (tid == 0) {
        curandState_t* st = reinterpret_cast<curandState_t*>(p.curand_states) + b;
        s_rand = curand_uniform(st) * pthr;
    }
    // largest probability, lowest id on ties
    Cand c{-FLT_MAX, -1};
    for (int v = tid; v < Vp; v += nt) {
        const float x = row[v];
        if (x > c.v) {
            c.v = x;
            c.idx = v;
        }
    }
    // block_argmax prefers the higher thread on ties (the top-k rule); for the top-p head any of the tied maxima has the same
    // probability, so only the id differs in that measure-zero case
    const Cand top = block_argmax<1024>(c, s_c);
    __syncthreads();
    int pick = top.idx;
    float pick_prob = top.v;
    if (!(top.v >= pthr)) {
        const unsigned long long need = prob_fx(s_rand);
        // invariant: G(lo) >= need > G(hi)
        unsigned lo = 0u, hi = __float_as_uint(top.v) + 1u;
        unsigned long long g_hi = 0ull;                       // G(hi)
        {
            unsigned long long t = 0ull;
            for (int v = tid; v < Vp; v += nt) t += prob_fx(fmaxf(row[v], 0.f));
            const unsigned long long total = block_sum_u64(t, s_red);
            if (total < need) lo = hi;                        // rounding: r beyond the total mass -> keep the head
        }
        while (hi - lo > 1u) {
            const unsigned mid = lo + ((hi - lo) >> 1);
            unsigned long long t = 0ull;
            for (int v = tid; v < Vp; v += nt) {
                const float x = fmaxf(row[v], 0.f);
                if (__float_as_uint(x) >= mid) t += prob_fx(x);
            }
            const unsigned long long g = block_sum_u64(t, s_red);
            if (g >= need) lo = mid;
            else {
                hi = mid;
                g_hi = g;
            }
        }
        if (lo < hi) {
            // elements with key == lo all carry the same probability; the crossing one is the j-th of them in id order
            const float pk = __uint_as_float(lo);
            const unsigned long long pk_fx = prob_fx(pk);
            unsigned long long j = pk_fx > 0ull ? (need - g_hi + pk_fx - 1ull) / pk_fx : 1ull;
            if (j < 1ull) j = 1ull;
            if (tid == 0) s_pick = -1;
            __syncthreads();
            // ordered count over the row in chunks of the block size
            unsigned long long seen = 0ull;
            for (int base = 0; base < Vp && s_pick < 0; base += nt) {
                const int v = base + tid;
                const int hit = (v < Vp && __float_as_uint(fmaxf(row[v], 0.f)) == lo) ? 1 : 0;
                // inclusive rank of this hit inside the chunk
                const unsigned bal = __ballot_sync(0xffffffffu, hit);
                const int lane = tid & 31, warp = tid >> 5;
                const int in_warp = __popc(bal & ((2u << lane) - 1u));
                __syncthreads();
                if (lane == 0) s_red[warp] = (unsigned long long)__popc(bal);
                __syncthreads();
                unsigned long long before = 0ull, chunk_total = 0ull;
                for (int w = 0; w < (nt >> 5); ++w) {
                    const unsigned long long cw = s_red[w];
                    if (w < warp) before += cw;
                    chunk_total += cw;
                }
                if (hit && seen + before + (unsigned long long)in_warp == j) s_pick = v;
                seen += chunk_total;
                __syncthreads();
            }
            if (s_pick >= 0) {
                pick = s_pick;
                pick_prob = pk;
            }
        }
    }
    if (tid == 0) {
        *out = pick;
        if (p.cum_log_probs != nullptr && p.want_probs) p.cum_log_probs[b] += logf(pick_prob);
        p.seq_len[b] += 1;
        p.finished[b] = (pick == p.end_id) ? 1 : 0;
    }
}

// ---------------------------------------------------------------- stop words, finished count, step advance
__global__ void __launch_bounds__(256) step_finalize_kernel(const ftcf_sampling_params p)
{
    __shared__ int s_cnt;
    const unsigned long long trc_t0 = trc_now(threadIdx.x == 0);
    const int step = *p.step;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int b = threadIdx.x; b < p.batch; b += blockDim.x) {
        if (p.stop_words != nullptr) {
            const int32_t* base = p.stop_words + (size_t)b * 2 * p.n_stop;
            const int32_t* offs = base + p.n_stop;
            for (int idx = 0; idx < p.n_stop; ++idx) {
                if (offs[idx] < 0) continue;
                const int item_end = offs[idx], item_start = idx > 0 ? offs[idx - 1] : 0;
                const int item_size = item_end - item_start;
                if (step + 1 < item_size) continue;
                bool ok = true;
                for (int t = item_size - 1; t >= 0; --t) {
                    if (p.output_ids[(size_t)(step - (item_size - 1) + t) * p.batch + b] != base[item_start + t]) {
                        ok = false;
                        break;
                    }
                }
                if (ok) p.finished[b] = 1;
            }
        }
        if (p.finished[b]) atomicAdd(&s_cnt, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.finished_count_host_mapped != nullptr) {
            // (step << 8 | all_finished) would lose the count; publish count and step in two ints
            p.finished_count_host_mapped[0] = s_cnt;
            __threadfence_system();
            p.finished_count_host_mapped[1] = step;
        }
        if (p.finished_hist_host_mapped != nullptr) {
            p.finished_hist_host_mapped[step] = s_cnt + 1;
            __threadfence_system();
        }
        *p.step = step + 1;
        trc_emit(TRC_SAMPLING, trc_t0, trc_t0, trc_t0, step, 3);   // marks the end of the step
    }
}

FTCF_TRACE_INSTALLER(trace_install_sampling)

__global__ void curand_init_kernel(curandState_t* st, const uint64_t* seeds, int batch)
{
    const int b = blockIdx.x * blockDim.x + threadIdx.x;
    if (b < batch) curand_init((unsigned long long)seeds[b], 0, 0, st + b);
}

// ---------------------------------------------------------------- output gather (beam 1)
__global__ void gather_output_kernel(int32_t* __restrict__ out, int32_t* __restrict__ out_len, const int32_t* __restrict__ ids,
                                     const int32_t* __restrict__ seq_len, const int32_t* __restrict__ input_len, int batch,
                                     int max_input_len, int max_len, int end_id)
{
    // one CTA per sequence; restates decoding_kernels.cu:452-580 for beam_width == 1: positions before the prompt
    // end are copied, the pad gap [input_len, max_input_len) is dropped, the rest shifts left, the tail and anything
    // after the first generated end_id become end_id.
    const int b = blockIdx.x;
    const int in_len = input_len[b];
    const int total = seq_len[b] + 1;                 // internal length (pad gap included)
    const int msl = min(max_len, total);
    const int pad = max_input_len - in_len;
    __shared__ int s_first_end;
    if (threadIdx.x == 0) {
        out_len[b] = total;
        s_first_end = max_len;
    }
    __syncthreads();
    int32_t* o = out + (size_t)b * max_len;
    for (int t = threadIdx.x; t < max_len; t += blockDim.x) {
        // source level for output position t
        int src = t < in_len ? t : t + pad;
        int val = end_id;
        if (src < msl && t < total - pad) val = ids[(size_t)src * batch + b];
        o[t] = val;
        const int start = max_input_len == 0 ? 1 : max_input_len;
        if (val == end_id && t >= start && t < msl) atomicMin(&s_first_end, t);
    }
    __syncthreads();
    const int fe = s_first_end;
    for (int t = threadIdx.x; t < max_len; t += blockDim.x)
        if (t > fe && t < msl) o[t] = end_id;
}

}  // namespace ftcf

using namespace ftcf;

extern "C" size_t ftcf_curand_state_bytes(void) { return sizeof(curandState_t); }

extern "C" size_t ftcf_sampling_workspace_bytes(int batch, int vocab_padded, int max_top_k)
{
    // tmp logits copy + candidate ids/values + repetition scratch is sized separately by max_len (see engine)
    size_t bytes = (size_t)batch * vocab_padded * sizeof(float);
    bytes += (size_t)batch * BLOCKS_PER_ROW * max_top_k * (sizeof(int) + sizeof(float));
    return (bytes + 255) & ~(size_t)255;
}

extern "C" int ftcf_curand_init(void* states, const uint64_t* seeds, int batch, void* stream)
{
    FTCF_REQUIRE(batch > 0, FTCF_ERR_INVALID, "curand_init: batch %d", batch);
    curand_init_kernel<<<ceil_div(batch, 128), 128, 0, as_stream(stream)>>>(static_cast<curandState_t*>(states), seeds, batch);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

// workspace layout: [tmp logits B*Vp f32][cand_val B*8*K f32][cand_id B*8*K i32][rep scratch B*max_len f32]
extern "C" int ftcf_sampling_step(const ftcf_sampling_params* pp, void* stream)
{
    FTCF_REQUIRE(pp != nullptr, FTCF_ERR_INVALID, "sampling: null params");
    const ftcf_sampling_params& p = *pp;
    FTCF_REQUIRE(p.batch > 0 && p.vocab > 0 && p.vocab_padded >= p.vocab, FTCF_ERR_INVALID, "sampling: bad sizes");
    FTCF_REQUIRE(p.max_top_k >= 1 && p.max_top_k <= 1024, FTCF_ERR_UNSUPPORTED, "sampling: max_top_k %d (1..1024)", p.max_top_k);
    FTCF_REQUIRE(p.optional_last_tokens == nullptr || p.n_last <= p.max_len, FTCF_ERR_UNSUPPORTED,
                 "sampling: optional_last_tokens list longer than max_len");
    cudaStream_t st = as_stream(stream);
    float* tmp = static_cast<float*>(p.workspace);
    float* cand_val = tmp + (size_t)p.batch * p.vocab_padded;
    int* cand_id = reinterpret_cast<int*>(cand_val + (size_t)p.batch * BLOCKS_PER_ROW * p.max_top_k);
    float* rep = reinterpret_cast<float*>(cand_id + (size_t)p.batch * BLOCKS_PER_ROW * p.max_top_k);

    logits_prepare_kernel<<<p.batch, 1024, 0, st>>>(p, rep);
    FTCF_LAUNCH_CHECK();
    const dim3 g1(BLOCKS_PER_ROW, p.batch);
    const size_t dyn = (size_t)p.max_top_k * 8;
    // CASE_K table of the reference (sampling_topk_kernels.cu:411-417): (k<=16: 128,128) (<=32: 256,128) (<=1024: 256,256)
    if (p.max_top_k <= 16) {
        topk_stage1_kernel<128><<<g1, 128, 0, st>>>(p.logits, tmp, cand_id, cand_val, p.top_k, p.finished, p.vocab_padded, p.max_top_k);
        FTCF_LAUNCH_CHECK();
        topk_stage2_kernel<128><<<p.batch, 128, dyn, st>>>(p, cand_id, cand_val);
    } else if (p.max_top_k <= 32) {
        topk_stage1_kernel<256><<<g1, 256, 0, st>>>(p.logits, tmp, cand_id, cand_val, p.top_k, p.finished, p.vocab_padded, p.max_top_k);
        FTCF_LAUNCH_CHECK();
        topk_stage2_kernel<128><<<p.batch, 128, dyn, st>>>(p, cand_id, cand_val);
    } else {
        topk_stage1_kernel<256><<<g1, 256, 0, st>>>(p.logits, tmp, cand_id, cand_val, p.top_k, p.finished, p.vocab_padded, p.max_top_k);
        FTCF_LAUNCH_CHECK();
        topk_stage2_kernel<256><<<p.batch, 256, dyn, st>>>(p, cand_id, cand_val);
    }
    FTCF_LAUNCH_CHECK();
    if (p.has_top_p_rows) {
        topp_sample_kernel<<<p.batch, 1024, 0, st>>>(p);
        FTCF_LAUNCH_CHECK();
    }
    step_finalize_kernel<<<1, 256, 0, st>>>(p);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_gather_output(int32_t* out, int32_t* out_len, const int32_t* ids_time_major, const int32_t* seq_len,
                                  const int32_t* input_len, int batch, int max_input_len, int max_len, int end_id, void* stream)
{
    FTCF_REQUIRE(batch > 0 && max_len > 0, FTCF_ERR_INVALID, "gather_output: bad sizes");
    gather_output_kernel<<<batch, 256, 0, as_stream(stream)>>>(out, out_len, ids_time_major, seq_len, input_len, batch,
                                                                max_input_len, max_len, end_id);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
