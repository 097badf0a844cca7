// Online beam search of the decode loop (beam_width > 1) on fp32 logits, rows = batch x beam.
//
// Reference behaviour restated (not ported) -- paths relative to src/fastertransformer:
//   * wiring: layers/DynamicDecodeLayer.cc:308-408, layers/beam_search_layers/BaseBeamSearchLayer.cu:170-285,
//     layers/beam_search_layers/OnlineBeamSearchLayer.cu:24-170
//   * penalties: kernels/beam_search_penalty_kernels.cu:24-50 (temperature x 1/(T + 1e-6), padded vocabulary -> -FLT_MAX),
//     :84-150 (repetition penalty along the beam's own history through parent_ids, pad gap skipped, every value taken from
//     the unpenalised logit)
//   * per row: online softmax + the 2K best logits by (value desc, id asc) per vocabulary part, parts merged, candidate value =
//     logit - max - log(sum) + cum_log_prob (kernels/online_softmax_beamsearch_kernels.cu:366-520, tie rule
//     kernels/reduce_kernel_utils.cuh:275-322); a finished row offers only (end_id, its cum_log_prob) (:390-398)
//   * per batch: K winners among K x 2K candidates by (score desc, candidate index asc), score = value [/ len^len_penalty]
//     + diversity * (candidate index % K); the winners keep the unpenalised value as cum_log_prob (:101-262)
//   * update of lengths / finished / parents / ids: OnlineBeamSearchLayer.cu:24-60; cache indirection: BaseBeamSearchLayer.cu:24-52
//   * stop words through the parents: kernels/stop_criteria_kernels.cu:24-84; gatherTree: kernels/decoding_kernels.cu:452-580
// The reference spends 6 launches (+ a host sync for the stop decision) per step; here: [penalties] -> candidates -> winners ->
// apply + indirection -> stop words / finished count, no host sync (same mapped finished history as the sampling path).
#include <float.h>

#include "common.cuh"

namespace ftcf {

constexpr int BEAM_THREADS = 256;
constexpr int BEAM_EPT = 8;                       // logits a thread keeps in registers per vocabulary part
constexpr int BEAM_MAX_PART = BEAM_THREADS * BEAM_EPT;
constexpr int BEAM_MAX_K = 32;
constexpr int BEAM_MAX_PARTS = 128;
constexpr int BEAM_PARENT_SMEM = 40 * 1024;     // staging of the generated levels' parents in the penalty kernel

struct BCand {
    float v;
    int idx;   // -1: none
};

__device__ __forceinline__ bool bcand_better(const BCand& a, const BCand& b)
{
    if (a.idx < 0) return false;
    if (b.idx < 0) return true;
    if (a.v != b.v) return a.v > b.v;
    return a.idx < b.idx;
}

// (value desc, id asc) winner of the CTA; s_c holds one entry per warp
__device__ __forceinline__ BCand beam_block_argmax(BCand c, BCand* s_c)
{
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        BCand other;
        other.v = __shfl_xor_sync(0xffffffffu, c.v, o);
        other.idx = __shfl_xor_sync(0xffffffffu, c.idx, o);
        if (bcand_better(other, c)) c = other;
    }
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    __syncthreads();
    if (lane == 0) s_c[warp] = c;
    __syncthreads();
    BCand r = (lane < BEAM_THREADS / 32) ? s_c[lane] : BCand{-FLT_MAX, -1};
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        BCand other;
        other.v = __shfl_xor_sync(0xffffffffu, r.v, o);
        other.idx = __shfl_xor_sync(0xffffffffu, r.idx, o);
        if (bcand_better(other, r)) r = other;
    }
    return r;
}

struct BeamWs {
    float* md;          // [BB, parts, 2]   (max, sum of exp) per vocabulary part
    float* cand_val;    // [BB, parts, C]
    int32_t* cand_id;   // [BB, parts, C]   id inside the row, -1: none
    int32_t* win_word;  // [BB] row * Vp + id of the winner that takes this slot
    float* win_val;     // [BB]
    int32_t* new_seq;   // [BB]
    int32_t* rep_idx;   // [BB, max_len]
    float* rep_val;     // [BB, max_len]
    int parts, part_len;
};

__host__ __device__ inline int beam_parts(int rows, int vocab_padded, int* part_len)
{
    int parts = (296 + rows - 1) / rows;
    const int need = (vocab_padded + BEAM_MAX_PART - 1) / BEAM_MAX_PART;
    if (parts < need) parts = need;
    if (parts > BEAM_MAX_PARTS) parts = BEAM_MAX_PARTS;
    int len = (vocab_padded + parts - 1) / parts;
    if (len < BEAM_THREADS) len = BEAM_THREADS < vocab_padded ? BEAM_THREADS : vocab_padded;
    parts = (vocab_padded + len - 1) / len;
    *part_len = len;
    return parts;
}

__host__ inline size_t beam_ws_layout(BeamWs* w, char* base, int rows, int beam, int vocab_padded, int max_len)
{
    int part_len = 0;
    const int parts = beam_parts(rows, vocab_padded, &part_len);
    const int C = 2 * beam;
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    const size_t o_md = carve((size_t)rows * parts * 2 * 4), o_cv = carve((size_t)rows * parts * C * 4), o_ci = carve((size_t)rows * parts * C * 4),
                 o_ww = carve((size_t)rows * 4), o_wv = carve((size_t)rows * 4), o_ns = carve((size_t)rows * 4),
                 o_ri = carve((size_t)rows * max_len * 4), o_rv = carve((size_t)rows * max_len * 4);
    if (w != nullptr) {
        w->md = (float*)(base + o_md); w->cand_val = (float*)(base + o_cv); w->cand_id = (int32_t*)(base + o_ci);
        w->win_word = (int32_t*)(base + o_ww); w->win_val = (float*)(base + o_wv); w->new_seq = (int32_t*)(base + o_ns);
        w->rep_idx = (int32_t*)(base + o_ri); w->rep_val = (float*)(base + o_rv);
        w->parts = parts; w->part_len = part_len;
    }
    return off;
}

// ---------------------------------------------------------------- penalties: one CTA per row
__global__ void __launch_bounds__(BEAM_THREADS) beam_penalty_kernel(const ftcf_beam_params p, const BeamWs w, int scale_on, int rep_on)
{
    extern __shared__ int32_t s_parents[];            // [levels, K] parents of this batch's generated levels (when they fit)
    __shared__ int s_n;
    const int bb = blockIdx.x, tid = threadIdx.x, K = p.beam_width, BB = p.batch * K;
    const int batch = bb / K, step = *p.step;
    float* row = p.logits + (size_t)bb * p.vocab_padded;
    if (scale_on) {
        const float inv_temp = 1.0f / (p.temperature + 1e-6f);
        for (int v = tid; v < p.vocab_padded; v += BEAM_THREADS) row[v] = v < p.vocab ? row[v] * inv_temp : -FLT_MAX;
    }
    if (!rep_on || step <= 0) return;
    __syncthreads();
    int32_t* idx = w.rep_idx + (size_t)bb * p.max_len;
    float* val = w.rep_val + (size_t)bb * p.max_len;
    const int in_len = p.input_len[bb], max_in = p.max_input_len;
    // History levels, newest first: entry 0 is the slot's own last token, entry 1 + (step - 2 - i) belongs to level i.  Levels
    // inside the pad gap are marked -1.  For level i the reference first replaces `parent` by parent_ids[i][parent] and then
    // reads the token of THAT slot at level i; below max_in every parent is 0, so only the generated levels are a serial walk.
    const int first_gen = max(max_in, 0);             // levels >= first_gen carry real parents
    const int gen_levels = max(step - 1 - first_gen, 0);     // levels first_gen .. step - 2
    const bool staged = (size_t)gen_levels * K * sizeof(int32_t) <= BEAM_PARENT_SMEM;
    if (staged) {
        for (int i = tid; i < gen_levels * K; i += BEAM_THREADS)
            s_parents[i] = p.parent_ids[(size_t)(first_gen + i / K) * BB + batch * K + i % K];
    }
    __syncthreads();
    if (tid == 0) {
        // the slot's own last token is filed under level step - 1, and the write-back skips pad-gap LEVELS: on the first step of a
        // row shorter than max_input_len it is therefore not penalised (beam_search_penalty_kernels.cu:141-147)
        const bool own_in_gap = step - 1 >= in_len && step - 1 < max_in;
        idx[0] = own_in_gap ? -1 : p.output_ids[(size_t)(step - 1) * BB + bb];
        int parent = bb % K;
        for (int i = step - 2; i >= first_gen; --i) {
            parent = staged ? s_parents[(i - first_gen) * K + parent] : p.parent_ids[(size_t)i * BB + batch * K + parent];
            idx[1 + (step - 2 - i)] = p.output_ids[(size_t)i * BB + batch * K + parent];
        }
        s_n = step;                                   // entries 0 .. step - 1
    }
    // context levels (i < max_in): the walk has reached parent 0 (or starts there when nothing was generated yet: then the
    // first lookup parent_ids[i][own slot] is 0 as well)
    for (int i = min(step - 2, first_gen - 1) - tid; i >= 0; i -= BEAM_THREADS) {
        const bool gap = i >= in_len && i < max_in;
        idx[1 + (step - 2 - i)] = gap ? -1 : p.output_ids[(size_t)i * BB + batch * K];
    }
    __syncthreads();
    const int n = s_n;
    const float pen = p.repetition_penalty;
    for (int i = tid; i < n; i += BEAM_THREADS) {
        const int t = idx[i];
        if (t < 0) continue;
        const float x = row[t];
        val[i] = x > 0.f ? x / pen : x * pen;
    }
    __syncthreads();
    for (int i = tid; i < n; i += BEAM_THREADS)
        if (idx[i] >= 0) row[idx[i]] = val[i];
}

// ---------------------------------------------------------------- candidates: grid (parts, rows)
__global__ void __launch_bounds__(BEAM_THREADS) beam_candidates_kernel(const ftcf_beam_params p, const BeamWs w)
{
    __shared__ BCand s_c[BEAM_THREADS / 32];
    __shared__ float s_red[32];
    const int part = blockIdx.x, bb = blockIdx.y, tid = threadIdx.x, C = 2 * p.beam_width;
    if (p.finished[bb]) return;                         // a finished row offers only end_id (see beam_winners_kernel)
    const float* row = p.logits + (size_t)bb * p.vocab_padded;
    const int start = part * w.part_len, end = min(start + w.part_len, p.vocab_padded);
    float x[BEAM_EPT];
    float mx = -FLT_MAX;
#pragma unroll
    for (int e = 0; e < BEAM_EPT; ++e) {
        const int v = start + tid + e * BEAM_THREADS;
        x[e] = v < end ? row[v] : -FLT_MAX;
        mx = fmaxf(mx, x[e]);
    }
    mx = block_max(mx, s_red);
    float sum = 0.f;
#pragma unroll
    for (int e = 0; e < BEAM_EPT; ++e) {
        const int v = start + tid + e * BEAM_THREADS;
        if (v < end) sum += __expf(x[e] - mx);
    }
    sum = block_sum(sum, s_red);
    float* md = w.md + ((size_t)bb * w.parts + part) * 2;
    if (tid == 0) {
        md[0] = mx;
        md[1] = sum;
    }
    float* cv = w.cand_val + ((size_t)bb * w.parts + part) * C;
    int32_t* ci = w.cand_id + ((size_t)bb * w.parts + part) * C;
    unsigned used = 0u;
    for (int r = 0; r < C; ++r) {
        BCand best{-FLT_MAX, -1};
#pragma unroll
        for (int e = 0; e < BEAM_EPT; ++e) {
            const int v = start + tid + e * BEAM_THREADS;
            if (v < end && !(used & (1u << e))) {
                const BCand c{x[e], v};
                if (bcand_better(c, best)) best = c;
            }
        }
        const BCand win = beam_block_argmax(best, s_c);
        if (win.idx >= 0 && (win.idx - start) % BEAM_THREADS == tid) used |= 1u << ((win.idx - start) / BEAM_THREADS);
        if (tid == 0) {
            cv[r] = win.v;
            ci[r] = win.idx;
        }
    }
}

// ---------------------------------------------------------------- winners: one CTA per batch
__global__ void __launch_bounds__(BEAM_THREADS) beam_winners_kernel(const ftcf_beam_params p, const BeamWs w)
{
    __shared__ BCand s_c[BEAM_THREADS / 32];
    __shared__ float s_red[32];
    __shared__ float s_val[BEAM_MAX_K * 2 * BEAM_MAX_K];
    __shared__ int32_t s_word[BEAM_MAX_K * 2 * BEAM_MAX_K];
    __shared__ float s_score[BEAM_MAX_K * 2 * BEAM_MAX_K];
    __shared__ uint8_t s_taken[BEAM_MAX_K * 2 * BEAM_MAX_K];
    const int b = blockIdx.x, tid = threadIdx.x, K = p.beam_width, C = 2 * K, Vp = p.vocab_padded;
    const int n = w.parts * C;                          // merged entries per row
    for (int j = 0; j < K; ++j) {
        const int bb = b * K + j;
        const float cum = p.cum_log_probs[bb];
        if (p.finished[bb]) {
            // +FLT_MAX at end_id, -FLT_MAX elsewhere: max = FLT_MAX, sum = 1, so end_id keeps cum and the rest is -inf
            for (int r = tid; r < C; r += BEAM_THREADS) {
                const int id = r == 0 ? p.end_id : (r - 1) + ((r - 1) >= p.end_id ? 1 : 0);
                s_val[j * C + r] = r == 0 ? cum : -INFINITY;
                s_word[j * C + r] = bb * Vp + id;
            }
            __syncthreads();
            continue;
        }
        const float* md = w.md + (size_t)bb * w.parts * 2;
        float M = -FLT_MAX;
        for (int q = tid; q < w.parts; q += BEAM_THREADS) M = fmaxf(M, md[2 * q]);
        M = block_max(M, s_red);
        float D = 0.f;
        for (int q = tid; q < w.parts; q += BEAM_THREADS) D += md[2 * q + 1] * __expf(md[2 * q] - M);
        D = block_sum(D, s_red);
        const float logD = logf(D);
        const float* cv = w.cand_val + (size_t)bb * n;
        const int32_t* ci = w.cand_id + (size_t)bb * n;
        unsigned used = 0u;                             // n <= 128 parts x 64 = 8192 entries = 32 per thread
        for (int r = 0; r < C; ++r) {
            BCand best{-FLT_MAX, -1};
            for (int i = tid, e = 0; i < n; i += BEAM_THREADS, ++e) {
                if (used & (1u << e)) continue;
                const BCand c{cv[i], ci[i]};
                if (bcand_better(c, best)) best = c;
            }
            const BCand win = beam_block_argmax(best, s_c);
            // mark the winner: ids are unique inside a row, so the entry that carries this id is the one
            for (int i = tid, e = 0; i < n; i += BEAM_THREADS, ++e)
                if (win.idx >= 0 && ci[i] == win.idx) used |= 1u << e;
            if (tid == 0) {
                s_val[j * C + r] = win.idx >= 0 ? (win.v - M - logD) + cum : -INFINITY;
                s_word[j * C + r] = bb * Vp + max(win.idx, 0);
            }
        }
        __syncthreads();
    }
    // ---- K winners among the K * C candidates
    const int total = K * C;
    float len_div = 1.f;
    if (p.length_penalty != 0.f) {
        // the reference indexes finished / sequence_lengths with the batch index of the call: the row `b` of the whole
        // batch when all batches go through one call, row b * K when differing runtime arguments force one call per batch
        const int li = p.args_differ ? b * K : b;
        const int length = p.finished[li] ? p.seq_len[li] : p.seq_len[li] + 1;
        if (length != 1) len_div = powf((float)length, p.length_penalty);
    }
    for (int i = tid; i < total; i += BEAM_THREADS) {
        float sc = s_val[i];
        if (len_div != 1.f) sc = sc / len_div;
        s_score[i] = sc + p.diversity_rate * (float)(i % K);
        s_taken[i] = 0;
    }
    __syncthreads();
    for (int r = 0; r < K; ++r) {
        BCand best{-FLT_MAX, -1};
        for (int i = tid; i < total; i += BEAM_THREADS) {
            if (s_taken[i]) continue;
            const BCand c{s_score[i], i};
            if (bcand_better(c, best)) best = c;
        }
        const BCand win = beam_block_argmax(best, s_c);
        if (tid == 0) {
            const int bb = b * K + r;
            const int word = s_word[win.idx];
            const int parent = (word / Vp) % K;
            const int pr = b * K + parent;
            w.win_word[bb] = word;
            w.win_val[bb] = s_val[win.idx];              // the value WITHOUT length penalty / diversity
            w.new_seq[bb] = p.seq_len[pr] + (p.finished[pr] ? 0 : 1);
            s_taken[win.idx] = 1;
        }
        __syncthreads();
    }
}

// ---------------------------------------------------------------- apply + cache indirection: one CTA per row
__global__ void __launch_bounds__(BEAM_THREADS) beam_apply_kernel(const ftcf_beam_params p, const BeamWs w)
{
    const int bb = blockIdx.x, tid = threadIdx.x, K = p.beam_width, BB = p.batch * K, Vp = p.vocab_padded;
    const int b = bb / K, j = bb % K, step = *p.step;
    const int word = w.win_word[bb];
    const int parent = (word / Vp) % K, tok = word % Vp;
    const bool fin = tok == p.end_id;
    if (tid == 0) {
        p.output_ids[(size_t)step * BB + bb] = tok;
        p.parent_ids[(size_t)step * BB + bb] = parent;
        p.seq_len[bb] = w.new_seq[bb];
        p.finished[bb] = fin ? 1 : 0;
        p.cum_log_probs[bb] = w.win_val[bb];
    }
    if (fin) return;                                   // rows of finished beams are left as they are
    const int parity = (step - p.max_input_len) & 1;
    const int32_t* src = p.cache_indir + ((size_t)parity * BB + b * K + parent) * p.max_len;
    int32_t* tgt = p.cache_indir + ((size_t)(1 - parity) * BB + bb) * p.max_len;
    const int nsteps = min(step + 1, p.max_len);
    for (int t = tid; t < nsteps; t += BEAM_THREADS) tgt[t] = t == step ? j : src[t];
}

// ---------------------------------------------------------------- stop words through the parents, finished count, step advance
__global__ void __launch_bounds__(256) beam_finalize_kernel(const ftcf_beam_params p)
{
    __shared__ int s_cnt;
    const int K = p.beam_width, BB = p.batch * K;
    const int step = *p.step;
    if (threadIdx.x == 0) s_cnt = 0;
    __syncthreads();
    for (int bb = threadIdx.x; bb < BB; bb += blockDim.x) {
        const int b = bb / K, j = bb % K;
        if (p.stop_words != nullptr) {
            const int32_t* base = p.stop_words + (size_t)b * 2 * p.n_stop;
            const int32_t* offs = base + p.n_stop;
            for (int idx = 0; idx < p.n_stop; ++idx) {
                if (offs[idx] < 0) continue;
                const int item_end = offs[idx], item_start = idx > 0 ? offs[idx - 1] : 0;
                const int item_size = item_end - item_start;
                if (step + 1 < item_size) continue;
                bool ok = true;
                int parent = j;
                for (int t = item_size - 1; t >= 0; --t) {
                    const size_t at = (size_t)(step - (item_size - 1) + t) * BB + b * K + parent;
                    if (p.output_ids[at] != base[item_start + t]) {
                        ok = false;
                        break;
                    }
                    parent = p.parent_ids[at];
                    if (parent < 0 || parent >= K) {
                        ok = false;
                        break;
                    }
                }
                if (ok) p.finished[bb] = 1;
            }
        }
        if (p.finished[bb]) atomicAdd(&s_cnt, 1);
    }
    __syncthreads();
    if (threadIdx.x == 0) {
        if (p.finished_count_host_mapped != nullptr) {
            p.finished_count_host_mapped[0] = s_cnt;
            __threadfence_system();
            p.finished_count_host_mapped[1] = step;
        }
        if (p.finished_hist_host_mapped != nullptr) {
            p.finished_hist_host_mapped[step] = s_cnt + 1;
            __threadfence_system();
        }
        *p.step = step + 1;
    }
}

// ---------------------------------------------------------------- output gather with parents: one CTA per batch, one thread per beam
__global__ void gather_output_beams_kernel(int32_t* __restrict__ out, int32_t* __restrict__ out_len, const int32_t* __restrict__ ids,
                                           const int32_t* __restrict__ parents, const int32_t* __restrict__ seq_len,
                                           const int32_t* __restrict__ input_len, int batch, int K, int max_input_len, int max_len, int end_id)
{
    const int b = blockIdx.x, j = threadIdx.x;
    if (j >= K) return;
    const int BB = batch * K, i = b * K + j;
    int longest = -1;
    for (int q = 0; q < K; ++q) longest = max(longest, seq_len[b * K + q] + 1);
    out_len[i] = seq_len[i] + 1;
    int32_t* o = out + (size_t)i * max_len;
    const int msl = min(max_len, longest);
    if (msl <= 0) return;
    const int in_len = input_len[i], pad = max_input_len - in_len;
    for (int t = 0; t < max_len; ++t) o[t] = 0;
    o[msl - 1 - pad] = ids[(size_t)(msl - 1) * BB + i];
    int parent = parents[(size_t)(msl - 1) * BB + i] % K;
    bool found_bad = false;
    for (int level = msl - 2; level >= 0; --level) {
        if (level >= in_len && level < max_input_len) continue;
        const int tgt = level >= max_input_len ? level - pad : level;
        if (parent < 0 || parent > K) {
            o[tgt] = end_id;
            parent = -1;
            found_bad = true;
        } else {
            o[tgt] = ids[(size_t)level * BB + b * K + parent];
            parent = parents[(size_t)level * BB + b * K + parent] % K;
        }
    }
    for (int t = longest - pad; t < max_len; ++t)
        if (t >= 0) o[t] = end_id;
    if (!found_bad) {
        bool fin = false;
        for (int t = max_input_len == 0 ? 1 : max_input_len; t < msl; ++t) {
            if (fin) o[t] = end_id;
            else if (o[t] == end_id) fin = true;
        }
    }
}

}  // namespace ftcf

using namespace ftcf;

extern "C" size_t ftcf_beam_workspace_bytes(int batch, int beam_width, int vocab_padded, int max_len)
{
    if (batch <= 0 || beam_width <= 0 || vocab_padded <= 0 || max_len <= 0) return 0;
    return beam_ws_layout(nullptr, nullptr, batch * beam_width, beam_width, vocab_padded, max_len);
}

extern "C" int ftcf_beam_search_step(const ftcf_beam_params* pp, void* stream)
{
    FTCF_REQUIRE(pp != nullptr, FTCF_ERR_INVALID, "beam search: null params");
    const ftcf_beam_params& p = *pp;
    FTCF_REQUIRE(p.batch > 0 && p.beam_width > 1 && p.vocab > 0 && p.vocab_padded >= p.vocab && p.max_len > 0, FTCF_ERR_INVALID,
                 "beam search: bad sizes (batch %d, beam_width %d, vocab %d/%d)", p.batch, p.beam_width, p.vocab, p.vocab_padded);
    FTCF_REQUIRE(p.beam_width <= BEAM_MAX_K, FTCF_ERR_UNSUPPORTED, "beam search: beam_width %d (supported: 2..%d)", p.beam_width, BEAM_MAX_K);
    FTCF_REQUIRE(p.logits && p.output_ids && p.parent_ids && p.seq_len && p.finished && p.cum_log_probs && p.input_len && p.cache_indir &&
                     p.step && p.workspace,
                 FTCF_ERR_INVALID, "beam search: null tensor");
    FTCF_REQUIRE((long long)p.batch * p.beam_width * p.vocab_padded < (1ll << 31), FTCF_ERR_UNSUPPORTED, "beam search: batch x beam x vocabulary overflows int32");
    const int rows = p.batch * p.beam_width;
    BeamWs w{};
    beam_ws_layout(&w, static_cast<char*>(p.workspace), rows, p.beam_width, p.vocab_padded, p.max_len);
    FTCF_REQUIRE(w.part_len <= BEAM_MAX_PART, FTCF_ERR_UNSUPPORTED, "beam search: vocabulary %d needs more than %d parts", p.vocab_padded, BEAM_MAX_PARTS);
    cudaStream_t st = as_stream(stream);
    const int scale_on = (p.temperature != 1.0f || p.vocab != p.vocab_padded) ? 1 : 0;
    const int rep_on = p.repetition_penalty != 1.0f ? 1 : 0;
    if (scale_on || rep_on) {
        beam_penalty_kernel<<<rows, BEAM_THREADS, rep_on ? BEAM_PARENT_SMEM : 0, st>>>(p, w, scale_on, rep_on);
        FTCF_LAUNCH_CHECK();
    }
    beam_candidates_kernel<<<dim3(w.parts, rows), BEAM_THREADS, 0, st>>>(p, w);
    FTCF_LAUNCH_CHECK();
    beam_winners_kernel<<<p.batch, BEAM_THREADS, 0, st>>>(p, w);
    FTCF_LAUNCH_CHECK();
    beam_apply_kernel<<<rows, BEAM_THREADS, 0, st>>>(p, w);
    FTCF_LAUNCH_CHECK();
    beam_finalize_kernel<<<1, 256, 0, st>>>(p);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}

extern "C" int ftcf_gather_output_beams(int32_t* out, int32_t* out_len, const int32_t* ids_time_major, const int32_t* parent_ids,
                                        const int32_t* seq_len, const int32_t* input_len, int batch, int beam_width, int max_input_len,
                                        int max_len, int end_id, void* stream)
{
    FTCF_REQUIRE(batch > 0 && beam_width > 1 && beam_width <= BEAM_MAX_K && max_len > 0, FTCF_ERR_INVALID, "gather_output_beams: bad sizes");
    FTCF_REQUIRE(out && out_len && ids_time_major && parent_ids && seq_len && input_len, FTCF_ERR_INVALID, "gather_output_beams: null tensor");
    gather_output_beams_kernel<<<batch, 32, 0, as_stream(stream)>>>(out, out_len, ids_time_major, parent_ids, seq_len, input_len, batch,
                                                                    beam_width, max_input_len, max_len, end_id);
    FTCF_LAUNCH_CHECK();
    return FTCF_OK;
}
