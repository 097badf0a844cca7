// Host engine: the request loop of the GPT-NeoX / CodeFuse path behind the C ABI (ftcf_gptneox_*).
//
// Mirrors, in behaviour, ft::GptNeoX<T>::forward (models/gptneox/GptNeoX.cc:385-1052) with GptNeoXContextDecoder
// (models/gptneox/GptNeoXContextDecoder.cc:223-512) for the prompt and GptNeoXDecoder
// (models/gptneox/GptNeoXDecoder.cc:197-389) for each generated token; weight order as th_op/gptneox/GptNeoXOp.h:121-174.
// What is different by design (B200-first):
//   * buffers are plain cudaMalloc slabs that only grow; the KV cache is [L][B, heads/t, max_len, dh] and is NOT
//     zero-filled per request (masked / unwritten slots are never read);
//   * every decode-step kernel takes the loop counter from device memory, so one captured CUDA graph is replayed for
//     all tokens of a request (and across requests of the same shape) instead of ~335 host launches per token;
//   * the per-token host spin-wait of the reference (kernels/stop_criteria_kernels.cu:135-156) is replaced by a
//     finished counter in mapped pinned memory polled two steps behind the stream;
//   * NCCL is loaded at run time (the library the process already has) and used only where the reference all-reduces.
#include <dlfcn.h>

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <string>
#include <vector>

#include "common.cuh"

namespace ftcf {
int splitk_reserve_for_stream(cudaStream_t st);   // gemm_decode.cu: split-K scratch must exist before the decode step is captured
void splitk_release_for_stream(cudaStream_t st);

// ------------------------------------------------------------------------------------------------ NCCL (dlopen)
typedef struct ncclComm* ncclComm_t;
struct NcclUid {
    char b[128];   // ncclUniqueId: 128 opaque bytes, passed by value
};
struct NcclApi {
    void* lib = nullptr;
    int (*GetUniqueId)(void*) = nullptr;
    int (*CommInitRank)(ncclComm_t*, int, NcclUid, int) = nullptr;
    int (*CommDestroy)(ncclComm_t) = nullptr;
    int (*AllReduce)(const void*, void*, size_t, int, int, ncclComm_t, cudaStream_t) = nullptr;
    int (*AllGather)(const void*, void*, size_t, int, ncclComm_t, cudaStream_t) = nullptr;
    const char* (*GetErrorString)(int) = nullptr;
    bool ok = false;
};
static NcclApi g_nccl;

static int nccl_load()
{
    if (g_nccl.ok) return FTCF_OK;
    const char* names[] = {"libnccl.so.2", "libnccl.so"};
    for (const char* nm : names) {
        g_nccl.lib = dlopen(nm, RTLD_NOW | RTLD_GLOBAL);
        if (g_nccl.lib) break;
    }
    FTCF_REQUIRE(g_nccl.lib != nullptr, FTCF_ERR_NCCL, "cannot dlopen libnccl.so.2: %s", dlerror());
#define FTCF_SYM(field, name)                                                                    \
    *reinterpret_cast<void**>(&g_nccl.field) = dlsym(g_nccl.lib, name);                         \
    FTCF_REQUIRE(g_nccl.field != nullptr, FTCF_ERR_NCCL, "libnccl: missing symbol %s", name)
    FTCF_SYM(GetUniqueId, "ncclGetUniqueId");
    FTCF_SYM(CommInitRank, "ncclCommInitRank");
    FTCF_SYM(CommDestroy, "ncclCommDestroy");
    FTCF_SYM(AllReduce, "ncclAllReduce");
    FTCF_SYM(AllGather, "ncclAllGather");
    FTCF_SYM(GetErrorString, "ncclGetErrorString");
#undef FTCF_SYM
    g_nccl.ok = true;
    return FTCF_OK;
}
#define FTCF_NCCL_CHECK(expr)                                                                    \
    do {                                                                                         \
        int _r = (expr);                                                                         \
        if (_r != 0) {                                                                           \
            set_error("%s failed: %s", #expr, g_nccl.GetErrorString ? g_nccl.GetErrorString(_r) : "?"); \
            return FTCF_ERR_NCCL;                                                                \
        }                                                                                        \
    } while (0)
constexpr int NCCL_FLOAT16 = 6, NCCL_FLOAT32 = 7, NCCL_SUM = 0;

// ------------------------------------------------------------------------------------------------ small kernels
// tok_b names the ROW (request row x beam + beam 0) a prompt token belongs to; the ids come from the request's [batch, S] tensor
__global__ void gather_prompt_ids_kernel(int32_t* out, const int32_t* ids, const int32_t* tok_b, const int32_t* tok_p, int T, int S, int beam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < T) out[i] = ids[(size_t)(tok_b[i] / beam) * S + tok_p[i]];
}
// rows = request batch x beam: every beam of a request starts from the same prompt (invokeTileGptInputs, kernels/gpt_kernels.cu)
__global__ void ids_to_time_major_kernel(int32_t* out, const int32_t* ids, int rows, int S, int beam)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < rows * S) {
        const int r = i / S, t = i % S;
        out[(size_t)t * rows + r] = ids[(size_t)(r / beam) * S + t];
    }
}
// x[b] = table[output_ids[*step - 1][b]]  (invokeEmbeddingLookupPosEncodingPadCount, kernels/decoding_kernels.cu:260)
__global__ void __launch_bounds__(256)
embedding_prev_token_kernel(__half* __restrict__ out, const __half* __restrict__ table, const int32_t* __restrict__ out_ids,
                            const int32_t* __restrict__ step, int B, int n, int vocab)
{
    const int nvec = n >> 3;
    const size_t vi = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    if (vi >= (size_t)B * nvec) return;
    const int row = (int)(vi / nvec), c = (int)(vi % nvec);
    int id = out_ids[(size_t)(*step - 1) * B + row];
    id = min(max(id, 0), vocab - 1);
    *reinterpret_cast<uint4*>(out + (size_t)row * n + c * 8) = ld_ro_16(table + (size_t)id * n + c * 8);
}
// [t, B, Vl] -> [B, t*Vl]  (invokeTransposeAxis01, kernels/gpt_kernels.cu:297-327)
__global__ void transpose_logits_kernel(float* __restrict__ out, const float* __restrict__ in, int t, int B, int Vl)
{
    const size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x;
    const size_t total = (size_t)t * B * Vl;
    if (i >= total) return;
    const int v = (int)(i % Vl);
    const int b = (int)((i / Vl) % B);
    const int r = (int)(i / ((size_t)Vl * B));
    out[(size_t)b * t * Vl + (size_t)r * Vl + v] = in[i];
}
__global__ void fill_i32_kernel(int32_t* p, int v, int n)
{
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    if (i < n) p[i] = v;
}

// ------------------------------------------------------------------------------------------------ engine
struct DevBuf {
    void* p = nullptr;
    size_t cap = 0;
    int ensure(size_t bytes)
    {
        if (bytes <= cap) return FTCF_OK;
        g_capture_generation.fetch_add(1, std::memory_order_relaxed);
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
        const size_t want = (bytes + 255) & ~(size_t)255;
        FTCF_CUDA_CHECK(cudaMalloc(&p, want));
        cap = want;
        return FTCF_OK;
    }
    void release()
    {
        if (p) cudaFree(p);
        p = nullptr;
        cap = 0;
    }
    template <typename T>
    T* as() const { return static_cast<T*>(p); }
};

struct LayerW {
    const __half *ln1_b, *ln1_g, *qkv_b, *o_b, *ffn1_b, *ffn2_b, *ln2_b, *ln2_g;
    const void* w[4];        // K-major weights: u8 [n,k] (int8) or fp16 [n,k]
    const __half* scale[4];  // int8 only
};

}  // namespace ftcf

using namespace ftcf;

struct ftcf_gptneox {
    ftcf_gptneox_config cfg{};
    cudaStream_t stream = nullptr;          // the engine's own non-blocking stream: all work (and the captured graph) runs here
    cudaStream_t caller_stream = nullptr;   // the stream handed in at construction (torch's current stream, GptNeoXOp.h:180)
    cudaEvent_t caller_ev = nullptr;
    cudaStream_t side = nullptr;            // second branch of the decode layer (FFN) so that it overlaps the attention branch
    cudaEvent_t ev_fork = nullptr, ev_join = nullptr;
    cudaStream_t side2 = nullptr;           // KV-cache L2 prefetch of the decode layer
    cudaEvent_t ev_join2 = nullptr;
    int h = 0, Hl = 0, hl = 0, inter_l = 0, Vp = 0, Vl = 0, t = 1, rank = 0;
    std::vector<LayerW> layers;
    const __half *wte = nullptr, *lnf_g = nullptr, *lnf_b = nullptr, *lm_head = nullptr;
    std::vector<DevBuf> owned;   // re-laid-out weights (fp16 transposes, plain->B200 int8)
    ncclComm_t comm = nullptr;
    // tensor-parallel exchange area (ftcf_tp_exchange): decode-size all-reduces run as remote stores from the O / FFN2 GEMM
    // epilogues into every rank's area (CUDA IPC over NVLink) instead of residual kernel + ncclAllReduce
    void* tp_area = nullptr;            // this rank's [data][counters], cudaMalloc'ed (IPC-exportable)
    size_t tp_data_bytes = 0;
    std::vector<void*> tp_opened;       // peers' areas mapped into this process
    ftcf_tp_exchange tpx{};             // peer pointers + geometry; step / step_base are filled per request
    bool tp_fused = false;              // the exchange area is up on EVERY rank
    static constexpr int kTpMaxRows = 32;

    // options
    int opt_cuda_graph = 1, opt_gemm_impl = 0, opt_step_timing = 0, opt_two_branch = 1, opt_fused_ln = 1, opt_kv_prefetch = 0, opt_pro_ctas = 148, opt_tp_fused = 1, opt_layer_hints = 1;
    // CTA targets of the four decode GEMMs of a layer in the fused path (0: pro_ctas for QKV / FFN1, the kernel's default for O / FFN2)
    int opt_qkv_ctas = 0, opt_ffn1_ctas = 0, opt_o_ctas = 0, opt_ffn2_ctas = 160,   // FFN2 at one CTA per SM leaves the attention kernel its slots (profiles/r2_decode_experiments.txt)
         opt_ffn2_no_pdl = 0, opt_ffn2_stages = 0, opt_o_stages = 0;

    // request-sized buffers (grow only)
    DevBuf kv, x, x2, n1, n2, qkv, qbuf, ctx, attn, inter, ffn, logits, logits_local, logits_gather, samp_ws, small, mmha_part,
        prompt_meta, lm_pad, attn_b, ffn_b;
    bool fused_on = false;              // decode layers run with the residual + LayerNorm prologue fused into the QKV / FFN1 / LM-head GEMMs
    int32_t* host_flag = nullptr;       // mapped pinned: [0] finished count, [1] step it belongs to
    int32_t* host_flag_dev = nullptr;
    int32_t* host_stage = nullptr;      // pinned staging for small uploads / callback reads
    size_t host_stage_cap = 0;
    int32_t* host_hist = nullptr;       // mapped pinned: [step] = 1 + finished count after that step (0: not written yet)
    int32_t* host_hist_dev = nullptr;
    size_t host_hist_cap = 0;           // entries

    // cached decode graphs, most recently used first (alternating request shapes -- the normal serving case -- replay instead
    // of re-capturing a 200-node graph per request)
    struct CachedGraph {
        std::string key;
        cudaGraphExec_t exec = nullptr;
        long long nodes = 0;
    };
    std::vector<CachedGraph> graphs;
    static constexpr size_t kMaxGraphs = 8;
    void drop_graphs()
    {
        for (auto& g : graphs)
            if (g.exec) cudaGraphExecDestroy(g.exec);
        graphs.clear();
    }

    std::vector<float> last_step_ms;
};

namespace {

struct Small {   // carved out of one small device slab; all int32 / float / u8 arrays of size B (or max_len * B)
    int32_t *out_ids, *seq_len, *input_len, *pad_count, *top_k, *step, *counters, *tok_b, *tok_p, *seq_off, *last_idx, *prompt_ids,
        *gathered, *gathered_len, *parent_ids, *cache_indir;
    int beam;                   // 1: sampling; > 1: beam search (rows = request batch x beam)
    float *top_p, *temperature, *rep_pen, *cum_log;
    uint8_t* finished;
    uint64_t* seeds;
    void* curand;
};

int engine_gemm(ftcf_gptneox* e, cudaStream_t st, const void* x, int layer, int kind, const __half* bias, void* y, int m, int n, int k, int act,
                int target_ctas = 0)
{
    const LayerW& L = e->layers[layer];
    if (e->cfg.int8_mode == 1) {
        const ftcf_launch_hint hint{target_ctas, 0, 0};
        return ftcf_gemm_w8a16_ex(x, static_cast<const uint8_t*>(L.w[kind]), L.scale[kind], bias, y, m, n, k, act, e->opt_gemm_impl,
                                  target_ctas > 0 ? &hint : nullptr, st);
    }
    return ftcf_gemm_f16(x, L.w[kind], bias, y, m, n, k, n, act, 0, e->opt_gemm_impl, st);
}

int engine_allreduce(ftcf_gptneox* e, void* buf, size_t count)
{
    if (e->t == 1) return FTCF_OK;
    FTCF_NCCL_CHECK(g_nccl.AllReduce(buf, buf, count, NCCL_FLOAT16, NCCL_SUM, e->comm, e->stream));
    g_launch_count.fetch_add(1, std::memory_order_relaxed);
    return FTCF_OK;
}

// Sets up the tensor-parallel exchange area: every rank exports its area as a CUDA IPC handle, the handles travel through
// ncclAllGather, every rank maps its peers.  All ranks then agree (ncclAllReduce MIN) on whether everybody succeeded; if not,
// the engine keeps the residual kernel + ncclAllReduce path (never a CPU path) -- e.g. when peer access is not available.
int tp_exchange_setup(ftcf_gptneox* e)
{
    const int t = e->t;
    if (t <= 1 || t > 8 || e->cfg.int8_mode != 1 || e->h % 128 != 0) return FTCF_OK;
    cudaStream_t st = e->stream;
    e->tp_data_bytes = (size_t)2 * 2 * t * ftcf_gptneox::kTpMaxRows * (e->h / 2) * 8;   // 8-byte flagged words, two columns each
    const size_t area_bytes = e->tp_data_bytes;
    int ok = 1;
    cudaIpcMemHandle_t mine{};
    if (cudaMalloc(&e->tp_area, area_bytes) != cudaSuccess || cudaMemset(e->tp_area, 0, area_bytes) != cudaSuccess ||
        cudaIpcGetMemHandle(&mine, e->tp_area) != cudaSuccess) {
        cudaGetLastError();
        ok = 0;
    }
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    char* dev_handles = nullptr;
    FTCF_CUDA_CHECK(cudaMalloc(&dev_handles, (size_t)t * 64 + 16));
    FTCF_CUDA_CHECK(cudaMemcpyAsync(dev_handles + (size_t)e->rank * 64, &mine, 64, cudaMemcpyHostToDevice, st));
    FTCF_NCCL_CHECK(g_nccl.AllGather(dev_handles + (size_t)e->rank * 64, dev_handles, 64, /*ncclChar*/ 0, e->comm, st));
    std::vector<cudaIpcMemHandle_t> all(t);
    FTCF_CUDA_CHECK(cudaMemcpyAsync(all.data(), dev_handles, (size_t)t * 64, cudaMemcpyDeviceToHost, st));
    FTCF_CUDA_CHECK(cudaStreamSynchronize(st));
    std::vector<void*> base(t, nullptr);
    for (int r = 0; r < t && ok; ++r) {
        if (r == e->rank) { base[r] = e->tp_area; continue; }
        void* p = nullptr;
        if (cudaIpcOpenMemHandle(&p, all[r], cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) {
            cudaGetLastError();
            ok = 0;
            break;
        }
        e->tp_opened.push_back(p);
        base[r] = p;
    }
    // unanimous?
    int32_t* flag = reinterpret_cast<int32_t*>(dev_handles + (size_t)t * 64);
    int32_t okv = ok;
    FTCF_CUDA_CHECK(cudaMemcpyAsync(flag, &okv, 4, cudaMemcpyHostToDevice, st));
    FTCF_NCCL_CHECK(g_nccl.AllReduce(flag, flag, 1, /*ncclInt32*/ 2, /*ncclMin*/ 3, e->comm, st));
    FTCF_CUDA_CHECK(cudaMemcpyAsync(&okv, flag, 4, cudaMemcpyDeviceToHost, st));
    FTCF_CUDA_CHECK(cudaStreamSynchronize(st));
    cudaFree(dev_handles);
    e->tp_fused = okv == 1;
    if (e->tp_fused) {
        ftcf_tp_exchange& x = e->tpx;
        x = ftcf_tp_exchange{};
        for (int r = 0; r < t; ++r) x.peer_data[r] = base[r];
        x.tp = t; x.rank = e->rank; x.m_max = ftcf_gptneox::kTpMaxRows; x.h = e->h; x.layer_num = e->cfg.layer_num;
    }
    return FTCF_OK;
}

#define FTCF_TRY(expr)                \
    do {                              \
        int _s = (expr);              \
        if (_s != FTCF_OK) return _s; \
    } while (0)

// One transformer layer on m rows of x (in place: x <- layer(x)).  attn_fn fills e->ctx from e->qkv.
template <typename AttnFn>
int run_layer(ftcf_gptneox* e, int l, int m, AttnFn&& attn_fn, bool tp_decode = false)
{
    const LayerW& L = e->layers[l];
    const ftcf_gptneox_config& c = e->cfg;
    cudaStream_t st = e->stream;
    __half* x = e->x.as<__half>();
    if (c.use_gptj_residual) {
        // Parallel residual: the attention branch (LN1, QKV, attention, O) and the FFN branch (LN2, FFN1, FFN2) only meet in
        // the residual add (GptNeoXDecoder.cc:301-360).  At decode sizes every kernel is a short HBM-bound stream, so the two
        // branches run on two streams (two parallel branches of the captured graph): one branch's kernel boundaries, LayerNorm
        // and the latency-bound attention kernel are covered by the other branch's weight streaming.
        const bool fork = e->opt_two_branch != 0 && m <= 32;
        cudaStream_t sb = fork ? e->side : st;
        if (fork) {
            FTCF_CUDA_CHECK(cudaEventRecord(e->ev_fork, st));
            FTCF_CUDA_CHECK(cudaStreamWaitEvent(sb, e->ev_fork, 0));
        }
        // decode with m <= 4 rows (tensor parallel, or fused_ln = 2): the two LayerNorms run as the prologue of the GEMMs that
        // consume them (the residual add cannot: with t > 1 its sum goes through the all-reduce first)
        // the two GEMMs that start the branches share the CTA slots (2 per SM) instead of queueing behind each other: at batch 32 the
        // QKV GEMM -- head of the critical path QKV -> attention -> O -- otherwise ends 30 us after FFN1 (profiles/r2_timeline_b32.txt)
        const int share = (fork && e->opt_layer_hints != 0) ? e->opt_pro_ctas : 0;
        const bool ln_pro = e->opt_fused_ln != 0 && m <= 4 && e->h % 128 == 0 && e->h <= 16384 && (e->t > 1 || e->opt_fused_ln == 2);
        auto ln_gemm = [&](cudaStream_t s2, const __half* g, const __half* b, int kind, const __half* bias, void* y, int n, int act) -> int {
            ftcf_ln_prologue pro{};
            pro.x = x; pro.gamma = g; pro.beta = b; pro.eps = c.layernorm_eps;
            pro.cta_hint = fork ? e->opt_pro_ctas : 0;
            if (c.int8_mode == 1)
                return ftcf_gemm_w8a16_ln(&pro, static_cast<const uint8_t*>(L.w[kind]), L.scale[kind], bias, y, m, n, e->h, act, s2);
            return ftcf_gemm_f16_ln(&pro, L.w[kind], bias, y, m, n, e->h, n, act, 0, s2);
        };
        if (ln_pro) {
            FTCF_TRY(ln_gemm(sb, L.ln2_g, L.ln2_b, 2, L.ffn1_b, e->inter.p, e->inter_l, 1));
        } else {
            FTCF_TRY(ftcf_layernorm(x, L.ln2_g, L.ln2_b, e->n2.p, m, e->h, c.layernorm_eps, sb));
            FTCF_TRY(engine_gemm(e, sb, e->n2.p, l, 2, L.ffn1_b, e->inter.p, m, e->inter_l, e->h, 1, share));
        }
        // decode rows with the exchange area up: the O / FFN2 epilogues store their tiles into every rank's area and
        // ftcf_tp_gather_residual rebuilds the all-reduced residual (no residual kernel, no ncclAllReduce)
        const bool tp_push = tp_decode && e->tp_fused && e->opt_tp_fused != 0 && c.int8_mode == 1 && m <= ftcf_gptneox::kTpMaxRows;
        if (tp_push) FTCF_TRY(ftcf_gemm_w8a16_tp_push(e->inter.p, static_cast<const uint8_t*>(L.w[3]), L.scale[3], &e->tpx, 1, l, m, e->h, e->inter_l, nullptr, sb));
        else FTCF_TRY(engine_gemm(e, sb, e->inter.p, l, 3, nullptr, e->ffn.p, m, e->h, e->inter_l, 0, share > 0 ? e->opt_ffn2_ctas : 0));
        if (fork) FTCF_CUDA_CHECK(cudaEventRecord(e->ev_join, sb));
        if (ln_pro) {
            FTCF_TRY(ln_gemm(st, L.ln1_g, L.ln1_b, 0, nullptr, e->qkv.p, 3 * e->hl, 0));
        } else {
            FTCF_TRY(ftcf_layernorm(x, L.ln1_g, L.ln1_b, e->n1.p, m, e->h, c.layernorm_eps, st));
            FTCF_TRY(engine_gemm(e, st, e->n1.p, l, 0, nullptr, e->qkv.p, m, 3 * e->hl, e->h, 0, share));
        }
        FTCF_TRY(attn_fn(l));
        if (tp_push) FTCF_TRY(ftcf_gemm_w8a16_tp_push(e->ctx.p, static_cast<const uint8_t*>(L.w[1]), L.scale[1], &e->tpx, 0, l, m, e->h, e->hl, nullptr, st));
        else FTCF_TRY(engine_gemm(e, st, e->ctx.p, l, 1, nullptr, e->attn.p, m, e->h, e->hl, 0));
        if (fork) FTCF_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_join, 0));
        // ffn2_b slot holds the summed (o + ffn2) bias, already divided by t (huggingface_convert.py:35-41,192-206)
        if (tp_push) {
            FTCF_TRY(ftcf_tp_gather_residual(&e->tpx, l, x, L.ffn2_b, x, m, st));
        } else {
            FTCF_TRY(ftcf_add_bias_attn_ffn_residual(x, e->ffn.p, e->attn.p, x, L.ffn2_b, m, e->h, e->t, st));
            FTCF_TRY(engine_allreduce(e, x, (size_t)m * e->h));
        }
    } else {
        FTCF_TRY(ftcf_layernorm(x, L.ln1_g, L.ln1_b, e->n1.p, m, e->h, c.layernorm_eps, st));
        FTCF_TRY(engine_gemm(e, st, e->n1.p, l, 0, nullptr, e->qkv.p, m, 3 * e->hl, e->h, 0));
        FTCF_TRY(attn_fn(l));
        FTCF_TRY(engine_gemm(e, st, e->ctx.p, l, 1, nullptr, e->attn.p, m, e->h, e->hl, 0));
        FTCF_TRY(engine_allreduce(e, e->attn.p, (size_t)m * e->h));
        // x2 = attn + bias_o + x ; n1 = LN2(x2)
        FTCF_TRY(ftcf_add_bias_residual_layernorm(x, e->attn.p, L.o_b, e->x2.p, L.ln2_g, L.ln2_b, e->n1.p, m, e->h, c.layernorm_eps, st));
        FTCF_TRY(engine_gemm(e, st, e->n1.p, l, 2, L.ffn1_b, e->inter.p, m, e->inter_l, e->h, 1));
        FTCF_TRY(engine_gemm(e, st, e->inter.p, l, 3, nullptr, e->ffn.p, m, e->h, e->inter_l, 0));
        FTCF_TRY(engine_allreduce(e, e->ffn.p, (size_t)m * e->h));
        FTCF_TRY(ftcf_add_bias_residual(x, e->ffn.p, e->x2.p, L.ffn2_b, m, e->h, st));
    }
    return FTCF_OK;
}

size_t kv_layer_elems(const ftcf_gptneox* e, int B, int max_len) { return (size_t)B * e->Hl * max_len * e->cfg.size_per_head; }

}  // namespace

extern "C" int ftcf_nccl_unique_id(void* out128)
{
    FTCF_REQUIRE(out128 != nullptr, FTCF_ERR_INVALID, "nccl_unique_id: null");
    FTCF_TRY(nccl_load());
    FTCF_NCCL_CHECK(g_nccl.GetUniqueId(out128));
    return FTCF_OK;
}

extern "C" int ftcf_gptneox_create(ftcf_gptneox** out, const ftcf_gptneox_config* cfg, const void* const* weights, size_t n_weights,
                                   const void* const* int8_weights, const void* const* scales, size_t n_int8,
                                   const void* nccl_unique_id, void* stream)
{
    FTCF_REQUIRE(out && cfg && weights, FTCF_ERR_INVALID, "create: null argument");
    FTCF_TRY(ftcf_device_check());
    const ftcf_gptneox_config& c = *cfg;
    const int L = c.layer_num, t = c.tensor_para_size;
    FTCF_REQUIRE(L > 0 && c.head_num > 0 && c.size_per_head > 0 && c.inter_size > 0 && c.vocab_size > 0, FTCF_ERR_INVALID,
                 "create: non-positive model dimension");
    FTCF_REQUIRE(t >= 1 && c.tensor_para_rank >= 0 && c.tensor_para_rank < t, FTCF_ERR_INVALID, "create: tensor_para rank %d of %d",
                 c.tensor_para_rank, t);
    FTCF_REQUIRE(c.head_num % t == 0 && c.inter_size % t == 0, FTCF_ERR_INVALID, "create: head_num %d / inter_size %d not divisible by tensor_para_size %d",
                 c.head_num, c.inter_size, t);
    FTCF_REQUIRE(c.int8_mode == 0 || c.int8_mode == 1, FTCF_ERR_UNSUPPORTED, "create: int8_mode %d (0 or 1; the driver asserts the same, codefuse_example.py:199)",
                 c.int8_mode);
    FTCF_REQUIRE(c.size_per_head == 64 || c.size_per_head == 128, FTCF_ERR_UNSUPPORTED, "create: size_per_head %d (64 or 128)", c.size_per_head);
    FTCF_REQUIRE(n_weights == (size_t)12 * L + 4, FTCF_ERR_INVALID, "create: expected %d weight tensors, got %zu", 12 * L + 4, n_weights);
    FTCF_REQUIRE(c.int8_mode == 0 || (int8_weights && scales && n_int8 == (size_t)4 * L), FTCF_ERR_INVALID,
                 "create: int8_mode=1 needs %d int8 weights and scales", 4 * L);

    auto* e = new ftcf_gptneox();
    e->cfg = c;
    e->caller_stream = as_stream(stream);
    // Stream capture is not allowed on the legacy default stream, which is what torch hands over by default, so the
    // engine owns a stream and orders it after the caller's with an event at the start of every request.
    // The attention branch of a decode layer (QKV -> attention -> O) is a chain of three dependent, partly latency-bound
    // kernels; the FFN branch is two big weight streams.  The main stream (attention branch) gets the highest priority so that
    // its CTAs are placed first and the latency-bound part overlaps the FFN streaming instead of trailing it.
    int prio_least = 0, prio_greatest = 0;
    cudaDeviceGetStreamPriorityRange(&prio_least, &prio_greatest);
    if (const char* pe = std::getenv("FTCF_STREAM_PRIO")) {   // experiment hook: 0 = no priorities, -1 = reversed
        const int v = std::atoi(pe);
        if (v == 0) prio_greatest = prio_least = 0;
        if (v < 0) std::swap(prio_least, prio_greatest);
    }
    if (cudaStreamCreateWithPriority(&e->stream, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->caller_ev, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->side, cudaStreamNonBlocking, prio_least) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_fork, cudaEventDisableTiming) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_join, cudaEventDisableTiming) != cudaSuccess ||
        cudaStreamCreateWithPriority(&e->side2, cudaStreamNonBlocking, prio_greatest) != cudaSuccess ||
        cudaEventCreateWithFlags(&e->ev_join2, cudaEventDisableTiming) != cudaSuccess) {
        set_error("create: cannot create the engine stream");
        delete e;
        return FTCF_ERR_CUDA;
    }
    cudaEventRecord(e->caller_ev, e->caller_stream);
    cudaStreamWaitEvent(e->stream, e->caller_ev, 0);
    e->t = t;
    e->rank = c.tensor_para_rank;
    e->h = c.head_num * c.size_per_head;
    e->Hl = c.head_num / t;
    e->hl = e->Hl * c.size_per_head;
    e->inter_l = c.inter_size / t;
    // vocab padding for fp16, models/gptneox/GptNeoX.cc:319-323
    e->Vp = (int)(std::ceil(std::ceil((double)c.vocab_size / t) / 8.0) * 8 * t);
    e->Vl = e->Vp / t;
    auto W = [&](int field, int l) { return static_cast<const __half*>(weights[(size_t)field * L + l]); };
    e->layers.resize(L);
    int status = FTCF_OK;
    const int gk[4] = {e->h, e->hl, e->h, e->inter_l};              // k of {qkv, o, ffn1, ffn2}
    const int gn[4] = {3 * e->hl, e->h, e->inter_l, e->h};          // n
    const int wfield[4] = {2, 4, 6, 8};
    for (int l = 0; l < L && status == FTCF_OK; ++l) {
        LayerW& lw = e->layers[l];
        lw.ln1_b = W(0, l); lw.ln1_g = W(1, l); lw.qkv_b = W(3, l); lw.o_b = W(5, l);
        lw.ffn1_b = W(7, l); lw.ffn2_b = W(9, l); lw.ln2_b = W(10, l); lw.ln2_g = W(11, l);
        for (int kind = 0; kind < 4 && status == FTCF_OK; ++kind) {
            const size_t elems = (size_t)gk[kind] * gn[kind];
            if (c.int8_mode == 1) {
                const void* q = int8_weights[(size_t)kind * L + l];
                lw.scale[kind] = static_cast<const __half*>(scales[(size_t)kind * L + l]);
                if (!q || !lw.scale[kind]) { set_error("create: missing int8 weight/scale (kind %d layer %d)", kind, l); status = FTCF_ERR_INVALID; break; }
                if (c.int8_layout == 0) {
                    lw.w[kind] = q;
                } else {
                    // plain int8 [k, n] -> K-major biased uint8 on the host (load-time only)
                    std::vector<int8_t> hq(elems);
                    std::vector<uint8_t> ho(elems);
                    if (cudaMemcpy(hq.data(), q, elems, cudaMemcpyDeviceToHost) != cudaSuccess) { set_error("create: D2H of int8 weight failed"); status = FTCF_ERR_CUDA; break; }
                    if (c.int8_layout == 2) ftcf_int8_ampere_to_b200_host(hq.data(), gk[kind], gn[kind], ho.data());   // reference-made *.q.bin
                    else ftcf_int8_plain_to_b200_host(hq.data(), gk[kind], gn[kind], ho.data());
                    e->owned.emplace_back();
                    status = e->owned.back().ensure(elems);
                    if (status == FTCF_OK && cudaMemcpy(e->owned.back().p, ho.data(), elems, cudaMemcpyHostToDevice) != cudaSuccess) { set_error("create: H2D failed"); status = FTCF_ERR_CUDA; }
                    lw.w[kind] = e->owned.back().p;
                }
            } else {
                const __half* wkn = W(wfield[kind], l);
                if (!wkn) { set_error("create: missing fp16 weight (kind %d layer %d)", kind, l); status = FTCF_ERR_INVALID; break; }
                e->owned.emplace_back();
                status = e->owned.back().ensure(elems * sizeof(__half));
                if (status == FTCF_OK) status = ftcf_transpose_f16(wkn, e->owned.back().p, gk[kind], gn[kind], e->stream);
                lw.w[kind] = e->owned.back().p;
                lw.scale[kind] = nullptr;
            }
        }
    }
    e->wte = static_cast<const __half*>(weights[(size_t)12 * L]);
    e->lnf_g = static_cast<const __half*>(weights[(size_t)12 * L + 1]);   // weight then bias, GptNeoXOp.h:172-173
    e->lnf_b = static_cast<const __half*>(weights[(size_t)12 * L + 2]);
    e->lm_head = static_cast<const __half*>(weights[(size_t)12 * L + 3]);
    if (status == FTCF_OK && (!e->wte || !e->lnf_g || !e->lnf_b || !e->lm_head)) { set_error("create: missing embedding / final layernorm / lm_head"); status = FTCF_ERR_INVALID; }
    if (status == FTCF_OK && e->Vp != c.vocab_size) {
        // zero-padded copy of the LM head so that every rank's slice is whole (GptNeoX.cc:749-764)
        status = e->lm_pad.ensure((size_t)e->Vp * e->h * sizeof(__half));
        if (status == FTCF_OK) {
            cudaMemsetAsync(e->lm_pad.p, 0, (size_t)e->Vp * e->h * sizeof(__half), e->stream);
            cudaMemcpyAsync(e->lm_pad.p, e->lm_head, (size_t)c.vocab_size * e->h * sizeof(__half), cudaMemcpyDeviceToDevice, e->stream);
            e->lm_head = e->lm_pad.as<__half>();
        }
    }
    if (status == FTCF_OK) status = splitk_reserve_for_stream(e->stream);
    if (status == FTCF_OK) status = splitk_reserve_for_stream(e->side);
    if (status == FTCF_OK && t > 1) {
        if (!nccl_unique_id) { set_error("create: tensor_para_size %d needs an NCCL unique id", t); status = FTCF_ERR_INVALID; }
        if (status == FTCF_OK) status = nccl_load();
        if (status == FTCF_OK) {
            NcclUid uid;
            memcpy(uid.b, nccl_unique_id, 128);
            int r = g_nccl.CommInitRank(&e->comm, t, uid, e->rank);
            if (r != 0) { set_error("ncclCommInitRank failed: %s", g_nccl.GetErrorString(r)); status = FTCF_ERR_NCCL; }
        }
    }
    if (status == FTCF_OK && t > 1) status = tp_exchange_setup(e);
    if (status == FTCF_OK) {
        if (cudaHostAlloc(&e->host_flag, 64, cudaHostAllocMapped) != cudaSuccess ||
            cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->host_flag_dev), e->host_flag, 0) != cudaSuccess) {
            set_error("create: mapped pinned allocation failed");
            status = FTCF_ERR_CUDA;
        }
    }
    if (status == FTCF_OK && cudaStreamSynchronize(e->stream) != cudaSuccess) { set_error("create: stream sync failed: %s", cudaGetErrorString(cudaGetLastError())); status = FTCF_ERR_CUDA; }
    if (status != FTCF_OK) {
        ftcf_gptneox_destroy(e);
        return status;
    }
    *out = e;
    return FTCF_OK;
}

extern "C" void ftcf_gptneox_destroy(ftcf_gptneox* e)
{
    if (!e) return;
    e->drop_graphs();
    for (auto& b : e->owned) b.release();
    DevBuf* bufs[] = {&e->kv, &e->x, &e->x2, &e->n1, &e->n2, &e->qkv, &e->qbuf, &e->ctx, &e->attn, &e->inter, &e->ffn, &e->logits,
                      &e->logits_local, &e->logits_gather, &e->samp_ws, &e->small, &e->mmha_part, &e->prompt_meta, &e->lm_pad, &e->attn_b,
                      &e->ffn_b};
    for (DevBuf* b : bufs) b->release();
    if (e->host_flag) cudaFreeHost(e->host_flag);
    if (e->host_stage) cudaFreeHost(e->host_stage);
    if (e->host_hist) cudaFreeHost(e->host_hist);
    for (void* p : e->tp_opened) cudaIpcCloseMemHandle(p);
    if (e->tp_area) cudaFree(e->tp_area);
    if (e->comm && g_nccl.ok) g_nccl.CommDestroy(e->comm);
    if (e->caller_ev) cudaEventDestroy(e->caller_ev);
    if (e->ev_fork) cudaEventDestroy(e->ev_fork);
    if (e->ev_join) cudaEventDestroy(e->ev_join);
    if (e->ev_join2) cudaEventDestroy(e->ev_join2);
    if (e->side2) cudaStreamDestroy(e->side2);
    if (e->side) { splitk_release_for_stream(e->side); cudaStreamDestroy(e->side); }
    if (e->stream) { splitk_release_for_stream(e->stream); cudaStreamDestroy(e->stream); }
    delete e;
}

extern "C" int ftcf_gptneox_set_option(ftcf_gptneox* e, const char* name, int value)
{
    FTCF_REQUIRE(e && name, FTCF_ERR_INVALID, "set_option: null");
    const std::string n(name);
    if (n == "cuda_graph") e->opt_cuda_graph = value;
    else if (n == "gemm_impl") e->opt_gemm_impl = value;
    else if (n == "step_timing") e->opt_step_timing = value;
    else if (n == "two_branch") e->opt_two_branch = value;
    else if (n == "fused_ln") e->opt_fused_ln = value;
    else if (n == "kv_prefetch") e->opt_kv_prefetch = value;
    else if (n == "pro_ctas") e->opt_pro_ctas = value;
    else if (n == "tp_fused") e->opt_tp_fused = value;
    else if (n == "layer_hints") e->opt_layer_hints = value;
    else if (n == "qkv_ctas") e->opt_qkv_ctas = value;
    else if (n == "ffn1_ctas") e->opt_ffn1_ctas = value;
    else if (n == "o_ctas") e->opt_o_ctas = value;
    else if (n == "ffn2_ctas") e->opt_ffn2_ctas = value;
    else if (n == "ffn2_no_pdl") e->opt_ffn2_no_pdl = value;
    else if (n == "ffn2_stages") e->opt_ffn2_stages = value;
    else if (n == "o_stages") e->opt_o_stages = value;
    else FTCF_REQUIRE(false, FTCF_ERR_INVALID, "set_option: unknown option %s", name);
    e->drop_graphs();   // anything captured may be stale
    return FTCF_OK;
}

extern "C" int ftcf_gptneox_last_step_ms(ftcf_gptneox* e, float* out, int n)
{
    if (!e || !out) return 0;
    const int c = std::min<int>(n, (int)e->last_step_ms.size());
    for (int i = 0; i < c; ++i) out[i] = e->last_step_ms[i];
    return c;
}

namespace {

template <typename T>
const T* pick(const T* arr, int n, int b)
{
    return n <= 1 ? arr : arr + b;
}

// The decode step: embedding of the previous token, L layers, final LN, LM head, sampling.  Everything reads the loop
// counter from device memory, so the same sequence can be captured once and replayed.
int decode_step(ftcf_gptneox* e, const Small& s, const ftcf_sampling_params& sp, int B, int max_len, int S, int splits, bool run_layers)
{
    const ftcf_gptneox_config& c = e->cfg;
    cudaStream_t st = e->stream;
    const int dh = c.size_per_head;
    if (e->fused_on) {
        // B <= 4, one GPU, parallel residual: the residual add of layer l-1 and the LayerNorms of layer l are the prologue of
        // layer l's QKV and FFN1 GEMMs (ftcf_gemm_*_ln); the residual stream ping-pongs between x and x2 and the branch
        // outputs between two (attn, ffn) pairs, so no kernel overwrites what a concurrently running one still reads.
        __half* xb[2] = {e->x.as<__half>(), e->x2.as<__half>()};
        __half* attn[2] = {e->attn.as<__half>(), e->attn_b.as<__half>()};
        __half* ffn[2] = {e->ffn.as<__half>(), e->ffn_b.as<__half>()};
        const int L = run_layers ? c.layer_num : 0;
        const bool w8 = c.int8_mode == 1;
        const bool tp_fused = e->t > 1;          // fused_on with t > 1 implies the exchange area is up (see ftcf_gptneox_forward)
        // Tensor parallel: the all-reduced residual of layer l - 1 is rebuilt from the exchanged partials by ftcf_tp_gather_residual, a
        // 10-CTA kernel that heads EACH branch of layer l (the FFN branch keeps its own copy of the residual stream in xs), so both
        // chains hang off it by programmatic launch.  Measured at tp 2 / 4 (profiles/r2_tp_experiments.txt): gather in every QKV /
        // FFN1 CTA's prologue 424 / 390 tokens/s, one gather before the fork 369 / 490, one per branch 460 / --.
        __half* xs[2] = {e->n1.as<__half>(), e->n2.as<__half>()};
        if (run_layers) {
            const size_t total = (size_t)B * e->h / 8;
            embedding_prev_token_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(xb[0], e->wte, s.out_ids, s.step, B, e->h, c.vocab_size);
            FTCF_LAUNCH_CHECK();
        }
        const size_t per_layer = kv_layer_elems(e, B, max_len);
        auto prologue = [&](int l, const __half* g, const __half* b, bool store) {
            ftcf_ln_prologue pro{};
            pro.gamma = g; pro.beta = b; pro.eps = c.layernorm_eps;
            // QKV and FFN1 start together: half of the SM slots each (measured: QKV otherwise queues behind FFN1's CTAs for ~25 us)
            pro.cta_hint = (l < c.layer_num && e->opt_two_branch) ? e->opt_pro_ctas : 0;
            if (l < c.layer_num && store && e->opt_qkv_ctas > 0) pro.cta_hint = e->opt_qkv_ctas;       // store == the QKV prologue
            if (l < c.layer_num && !store && e->opt_ffn1_ctas > 0) pro.cta_hint = e->opt_ffn1_ctas;
            if (l == 0) {
                pro.x = xb[0];
            } else {
                pro.x = xb[(l - 1) & 1];
                pro.add_bias = e->layers[l - 1].ffn2_b;   // (b_o + b_ffn2) / t, huggingface_convert.py:35-41,192-206
                if (tp_fused) {
                    // the partial sums of EVERY rank's O / FFN2 of layer l - 1 were pushed into this rank's exchange area by their
                    // epilogues and summed by the branch's gather kernel -- the all-reduce of GptNeoXDecoder.cc:348-359
                    pro.x = store ? xb[l & 1] : xs[l & 1];
                    pro.add_bias = nullptr;
                    return pro;
                }
                pro.add_ffn = ffn[(l - 1) & 1];
                pro.add_attn = attn[(l - 1) & 1];
                if (store) pro.x_out = xb[l & 1];
            }
            return pro;
        };
        for (int l = 0; l < L; ++l) {
            const LayerW& lw = e->layers[l];
            cudaStream_t sb = e->opt_two_branch ? e->side : st;
            // The FFN branch is issued first: measured best (its two big weight streams take the SMs, the attention branch's
            // shorter kernels fill in).  Issuing QKV first, or forking only after QKV, was 8-25 % slower per token.
            if (e->opt_two_branch) {
                FTCF_CUDA_CHECK(cudaEventRecord(e->ev_fork, st));
                FTCF_CUDA_CHECK(cudaStreamWaitEvent(sb, e->ev_fork, 0));
            }
            ftcf_mmha_params mp{};
            mp.qkv = e->qkv.p;
            mp.qkv_bias = lw.qkv_b;
            mp.k_cache = e->kv.as<__half>() + (size_t)(2 * l) * per_layer;
            mp.v_cache = e->kv.as<__half>() + (size_t)(2 * l + 1) * per_layer;
            mp.ctx = e->ctx.p;
            mp.seq_len = s.seq_len; mp.input_len = s.input_len; mp.pad_count = s.pad_count; mp.finished = s.finished; mp.step = s.step;
            mp.partial = e->mmha_part.as<float>(); mp.counters = s.counters;
            mp.batch = B; mp.heads = e->Hl; mp.dh = dh; mp.rotary_dim = c.rotary_embedding_dim;
            mp.max_len = max_len; mp.max_input_len = S; mp.splits = splits;
            mp.inv_sqrt_dh = 1.f / std::sqrt((float)dh);
            if (s.beam > 1) { mp.cache_indir = s.cache_indir; mp.beam_width = s.beam; }
            const bool kvpf = e->opt_two_branch && e->opt_kv_prefetch;
            if (kvpf) {   // this layer's cache rows start travelling to L2 now, while the GEMMs stream their weights
                FTCF_CUDA_CHECK(cudaStreamWaitEvent(e->side2, e->ev_fork, 0));
                FTCF_TRY(ftcf_mmha_prefetch_cache(&mp, e->side2));
                FTCF_CUDA_CHECK(cudaEventRecord(e->ev_join2, e->side2));
            }
            if (tp_fused && l > 0) {
                const void* bias = e->layers[l - 1].ffn2_b;   // (b_o + b_ffn2) / t, huggingface_convert.py:35-41,192-206
                FTCF_TRY(ftcf_tp_gather_residual(&e->tpx, l - 1, xb[(l - 1) & 1], bias, xb[l & 1], B, st));
                FTCF_TRY(ftcf_tp_gather_residual(&e->tpx, l - 1, l == 1 ? xb[0] : xs[(l - 1) & 1], bias, xs[l & 1], B, sb));
            }
            const ftcf_ln_prologue p2 = prologue(l, lw.ln2_g, lw.ln2_b, false);
            const ftcf_ln_prologue p1 = prologue(l, lw.ln1_g, lw.ln1_b, true);
            const ftcf_launch_hint h2{e->opt_ffn2_ctas, e->opt_ffn2_no_pdl, e->opt_ffn2_stages};
            const ftcf_launch_hint ho{e->opt_o_ctas, 0, e->opt_o_stages};
            auto ffn1 = [&]() -> int {
                if (w8) return ftcf_gemm_w8a16_ln(&p2, static_cast<const uint8_t*>(lw.w[2]), lw.scale[2], lw.ffn1_b, e->inter.p, B, e->inter_l, e->h, 1, sb);
                return ftcf_gemm_f16_ln(&p2, lw.w[2], lw.ffn1_b, e->inter.p, B, e->inter_l, e->h, e->inter_l, 1, 0, sb);
            };
            auto ffn2 = [&]() -> int {
                if (tp_fused) return ftcf_gemm_w8a16_tp_push(e->inter.p, static_cast<const uint8_t*>(lw.w[3]), lw.scale[3], &e->tpx, 1, l, B, e->h, e->inter_l, &h2, sb);
                if (w8) return ftcf_gemm_w8a16_ex(e->inter.p, static_cast<const uint8_t*>(lw.w[3]), lw.scale[3], nullptr, ffn[l & 1], B, e->h, e->inter_l, 0, e->opt_gemm_impl, &h2, sb);
                return ftcf_gemm_f16(e->inter.p, lw.w[3], nullptr, ffn[l & 1], B, e->h, e->inter_l, e->h, 0, 0, 1, sb);
            };
            auto qkv = [&]() -> int {
                if (w8) return ftcf_gemm_w8a16_ln(&p1, static_cast<const uint8_t*>(lw.w[0]), lw.scale[0], nullptr, e->qkv.p, B, 3 * e->hl, e->h, 0, st);
                return ftcf_gemm_f16_ln(&p1, lw.w[0], nullptr, e->qkv.p, B, 3 * e->hl, e->h, 3 * e->hl, 0, 0, st);
            };
            auto oproj = [&]() -> int {
                if (tp_fused) return ftcf_gemm_w8a16_tp_push(e->ctx.p, static_cast<const uint8_t*>(lw.w[1]), lw.scale[1], &e->tpx, 0, l, B, e->h, e->hl, &ho, st);
                if (w8) return ftcf_gemm_w8a16_ex(e->ctx.p, static_cast<const uint8_t*>(lw.w[1]), lw.scale[1], nullptr, attn[l & 1], B, e->h, e->hl, 0, e->opt_gemm_impl, &ho, st);
                return ftcf_gemm_f16(e->ctx.p, lw.w[1], nullptr, attn[l & 1], B, e->h, e->hl, e->h, 0, 0, 1, st);
            };
            // (measured and removed: QKV alone first, then FFN beside the attention: +25 % per token; FFN2 held back until the
            // attention finished: +4 %; profiles/r2_decode_experiments.txt)
            FTCF_TRY(ffn1());
            FTCF_TRY(ffn2());
            if (e->opt_two_branch) FTCF_CUDA_CHECK(cudaEventRecord(e->ev_join, sb));
            FTCF_TRY(qkv());
            FTCF_TRY(ftcf_mmha_decode(&mp, st));
            FTCF_TRY(oproj());
            if (e->opt_two_branch) FTCF_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_join, 0));
            if (kvpf) FTCF_CUDA_CHECK(cudaStreamWaitEvent(st, e->ev_join2, 0));
        }
        if (!tp_fused) {
            // final LayerNorm (on the last layer's residual sum) as the prologue of the LM head
            const ftcf_ln_prologue pf = prologue(L, e->lnf_g, e->lnf_b, false);
            return ftcf_gemm_f16_ln(&pf, e->lm_head, nullptr, e->logits.p, B, e->Vp, e->h, e->Vp, 0, 1, st);
        }
        // tensor parallel: the last layer's exchange is gathered into e->x, then the sharded LM head below
        if (L > 0) FTCF_TRY(ftcf_tp_gather_residual(&e->tpx, L - 1, xb[(L - 1) & 1], e->layers[L - 1].ffn2_b, e->x.p, B, st));
    } else if (run_layers) {
        const size_t total = (size_t)B * e->h / 8;
        embedding_prev_token_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->x.as<__half>(), e->wte, s.out_ids, s.step, B, e->h,
                                                                                      c.vocab_size);
        FTCF_LAUNCH_CHECK();
        const size_t per_layer = kv_layer_elems(e, B, max_len);
        for (int l = 0; l < c.layer_num; ++l) {
            auto attn = [&](int layer) -> int {
                ftcf_mmha_params mp{};
                mp.qkv = e->qkv.p;
                mp.qkv_bias = e->layers[layer].qkv_b;
                mp.k_cache = e->kv.as<__half>() + (size_t)(2 * layer) * per_layer;
                mp.v_cache = e->kv.as<__half>() + (size_t)(2 * layer + 1) * per_layer;
                mp.ctx = e->ctx.p;
                mp.seq_len = s.seq_len;
                mp.input_len = s.input_len;
                mp.pad_count = s.pad_count;
                mp.finished = s.finished;
                mp.step = s.step;
                mp.partial = e->mmha_part.as<float>();
                mp.counters = s.counters;
                mp.batch = B; mp.heads = e->Hl; mp.dh = dh; mp.rotary_dim = c.rotary_embedding_dim;
                mp.max_len = max_len; mp.max_input_len = S; mp.splits = splits;
                mp.inv_sqrt_dh = 1.f / std::sqrt((float)dh);
                if (s.beam > 1) { mp.cache_indir = s.cache_indir; mp.beam_width = s.beam; }
                return ftcf_mmha_decode(&mp, st);
            };
            FTCF_TRY(run_layer(e, l, B, attn, e->t > 1));
        }
    }
    // the LM head stays on the streaming kernel up to 32 rows (1 GB of fp16 weights, 6.9 TB/s there)
    const int lm_impl = e->opt_gemm_impl == 2 ? (B <= 32 ? 1 : 0) : (e->opt_gemm_impl == 0 && B <= 32 ? 1 : e->opt_gemm_impl);
    // final LN on x (decode) -- for the first generated token the caller has put the prefill's last-token rows in x
    FTCF_TRY(ftcf_layernorm(e->x.p, e->lnf_g, e->lnf_b, e->n1.p, B, e->h, c.layernorm_eps, st));
    if (e->t == 1) {
        FTCF_TRY(ftcf_gemm_f16(e->n1.p, e->lm_head, nullptr, e->logits.p, B, e->Vp, e->h, e->Vp, 0, 1, lm_impl, st));
    } else {
        const __half* slice = e->lm_head + (size_t)e->rank * e->Vl * e->h;
        FTCF_TRY(ftcf_gemm_f16(e->n1.p, slice, nullptr, e->logits_local.p, B, e->Vl, e->h, e->Vl, 0, 1, lm_impl, st));
        FTCF_NCCL_CHECK(g_nccl.AllGather(e->logits_local.p, e->logits_gather.p, (size_t)B * e->Vl, NCCL_FLOAT32, e->comm, st));
        g_launch_count.fetch_add(1, std::memory_order_relaxed);
        const size_t total = (size_t)e->t * B * e->Vl;
        transpose_logits_kernel<<<(unsigned)((total + 255) / 256), 256, 0, st>>>(e->logits.as<float>(), e->logits_gather.as<float>(), e->t, B, e->Vl);
        FTCF_LAUNCH_CHECK();
    }
    return FTCF_OK;
}

}  // namespace

extern "C" int ftcf_gptneox_forward(ftcf_gptneox* e, const ftcf_gptneox_request* rq, ftcf_gptneox_stats* stats)
{
    FTCF_REQUIRE(e && rq, FTCF_ERR_INVALID, "forward: null argument");
    const ftcf_gptneox_request& r = *rq;
    const ftcf_gptneox_config& c = e->cfg;
    // Beam search (beam_width K > 1) runs the decode loop on B = batch x K rows, like the reference (GptNeoX.cc:88-156); the prompt
    // is prefilled ONCE per request row into the cache row of beam 0 -- the cache indirection starts at 0 and only ever names
    // other beams for generated positions (BaseBeamSearchLayer.cu:24-52), so the K - 1 redundant prefills of GptNeoX.cc:589-682
    // are never read.
    const int K = r.beam_width > 1 ? r.beam_width : 1;
    const int Bq = r.batch, B = Bq * K, S = r.max_input_len, out_len = r.output_len;
    FTCF_REQUIRE(Bq > 0 && S >= 1 && out_len >= 1, FTCF_ERR_INVALID, "forward: batch %d, input length %d, output_len %d", Bq, S, out_len);
    FTCF_REQUIRE(r.input_ids && r.input_lengths && r.output_ids && r.sequence_lengths, FTCF_ERR_INVALID, "forward: null tensor");
    FTCF_REQUIRE(K <= 32, FTCF_ERR_UNSUPPORTED, "forward: beam_width %d (supported: 1..32)", K);
    FTCF_REQUIRE(K == 1 || S > 1, FTCF_ERR_UNSUPPORTED, "forward: beam search needs a prompt (max_input_len > 1)");
    FTCF_REQUIRE(K == 1 || r.optional_last_tokens == nullptr, FTCF_ERR_UNSUPPORTED, "forward: optional_last_tokens with beam search");
    const int max_len = S + out_len;
    cudaStream_t st = e->stream;
    FTCF_CUDA_CHECK(cudaEventRecord(e->caller_ev, e->caller_stream));   // inputs were produced on the caller's stream
    FTCF_CUDA_CHECK(cudaStreamWaitEvent(st, e->caller_ev, 0));
    const long long launches0 = g_launch_count.load();
    const int dh = c.size_per_head, L = c.layer_num;

    // ---- host copies of the lengths, token bookkeeping for the padding-removed prefill (GptNeoXContextDecoder.cc:285-308)
    std::vector<int32_t> lens_q(Bq), lens(B);
    FTCF_CUDA_CHECK(cudaMemcpyAsync(lens_q.data(), r.input_lengths, Bq * sizeof(int32_t), cudaMemcpyDeviceToHost, st));
    FTCF_CUDA_CHECK(cudaStreamSynchronize(st));
    int T = 0;
    for (int b = 0; b < Bq; ++b) {
        FTCF_REQUIRE(lens_q[b] >= 1 && lens_q[b] <= S, FTCF_ERR_INVALID, "forward: input_lengths[%d] = %d outside [1, %d]", b, lens_q[b], S);
        T += lens_q[b];
    }
    for (int b = 0; b < B; ++b) lens[b] = lens_q[b / K];
    const bool has_prefill = S > 1;
    const int m_max = has_prefill ? std::max(T, B) : B;

    // ---- sampling arguments (TopKSamplingLayer.cu:28-78 setup rules)
    std::vector<int32_t> ks(B);
    std::vector<float> ps(B), temps(B, 1.f), reps(B, 1.f);
    std::vector<uint64_t> seeds(B, 0);
    int max_top_k = 1;
    bool any_temp = false, any_rep = false, any_topp = false;
    for (int b = 0; b < B; ++b) {
        const int bq = b / K;
        int k = r.top_k_host ? *pick(r.top_k_host, r.n_top_k, bq) : 0;
        float p = r.top_p_host ? *pick(r.top_p_host, r.n_top_p, bq) : 0.f;
        if (k < 0) k = 0;
        if (k == 0 && p == 0.f) k = 1;
        if (k > 0 && p == 0.f) p = 1.f;
        any_topp |= k == 0;                     // pure top-p row (TopPSamplingLayer.cu:60-78)
        if (k > 1024) k = 1024;
        p = std::min(std::max(p, 0.f), 1.f);
        ks[b] = k;
        ps[b] = p;
        max_top_k = std::max(max_top_k, k);
        if (r.temperature_host) temps[b] = *pick(r.temperature_host, r.n_temperature, bq);
        if (r.repetition_penalty_host) reps[b] = *pick(r.repetition_penalty_host, r.n_repetition_penalty, bq);
        if (r.random_seed_host) seeds[b] = (uint64_t)*pick(r.random_seed_host, r.n_random_seed, bq);
        any_temp |= temps[b] != 1.f;
        any_rep |= reps[b] != 1.f;
    }

    // ---- buffers
    const size_t per_layer = kv_layer_elems(e, B, max_len);
    FTCF_TRY(e->kv.ensure(per_layer * 2 * L * sizeof(__half)));
    FTCF_TRY(e->x.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->x2.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->n1.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->n2.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->qkv.ensure((size_t)m_max * 3 * e->hl * 2));
    FTCF_TRY(e->qbuf.ensure((size_t)m_max * e->hl * 2));
    FTCF_TRY(e->ctx.ensure((size_t)m_max * e->hl * 2));
    FTCF_TRY(e->attn.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->inter.ensure((size_t)m_max * e->inter_l * 2));
    FTCF_TRY(e->ffn.ensure((size_t)m_max * e->h * 2));
    FTCF_TRY(e->logits.ensure((size_t)B * e->Vp * 4));
    if (e->t > 1) {
        FTCF_TRY(e->logits_local.ensure((size_t)B * e->Vl * 4));
        FTCF_TRY(e->logits_gather.ensure((size_t)B * e->Vp * 4));
    }
    const size_t ws_bytes = K > 1 ? ftcf_beam_workspace_bytes(Bq, K, e->Vp, max_len)
                                  : ftcf_sampling_workspace_bytes(B, e->Vp, max_top_k) + (size_t)B * max_len * 4 + 256;
    FTCF_TRY(e->samp_ws.ensure(ws_bytes));
    const int splits = ftcf_mmha_choose_splits(B, e->Hl, max_len);
    FTCF_TRY(e->mmha_part.ensure((size_t)B * e->Hl * splits * (dh + 2) * 4 + 256));
    e->fused_on = e->opt_fused_ln == 1 && B <= 4 && c.use_gptj_residual != 0 && e->h % 128 == 0 && e->h <= 16384 &&
                  (e->t == 1 || (e->tp_fused && e->opt_tp_fused != 0 && c.int8_mode == 1));
    if (e->fused_on) {
        FTCF_TRY(e->attn_b.ensure((size_t)B * e->h * 2));
        FTCF_TRY(e->ffn_b.ensure((size_t)B * e->h * 2));
    }
    // small slab layout
    Small s{};
    size_t off = 0;
    auto carve = [&](size_t bytes) { size_t o = off; off += (bytes + 255) & ~(size_t)255; return o; };
    // Everything a captured decode graph points into comes first and is sized by (B, max_len) only; the arrays sized by the
    // token count T (prefill bookkeeping, never touched by the graph) come last, so two requests of the same shape with
    // different ragged lengths see the same offsets.
    const size_t o_out_ids = carve((size_t)max_len * B * 4), o_seq = carve(B * 4), o_inlen = carve(B * 4), o_pad = carve(B * 4),
                 o_topk = carve(B * 4), o_step = carve(4), o_cnt = carve((size_t)B * e->Hl * 4), o_seqoff = carve((B + 1) * 4),
                 o_last = carve(B * 4), o_gath = carve((size_t)B * max_len * 4), o_glen = carve(B * 4), o_topp = carve(B * 4),
                 o_temp = carve(B * 4), o_rep = carve(B * 4), o_cum = carve(B * 4), o_fin = carve(B), o_seeds = carve(B * 8),
                 o_curand = carve((size_t)B * ftcf_curand_state_bytes()),
                 o_parent = carve(K > 1 ? (size_t)max_len * B * 4 : 0), o_indir = carve(K > 1 ? (size_t)2 * B * max_len * 4 : 0),
                 o_tokb = carve((size_t)std::max(T, 1) * 4),
                 o_tokp = carve((size_t)std::max(T, 1) * 4), o_pids = carve((size_t)std::max(T, 1) * 4);
    FTCF_TRY(e->small.ensure(off));
    char* sb = e->small.as<char>();
    s.out_ids = (int32_t*)(sb + o_out_ids); s.seq_len = (int32_t*)(sb + o_seq); s.input_len = (int32_t*)(sb + o_inlen);
    s.pad_count = (int32_t*)(sb + o_pad); s.top_k = (int32_t*)(sb + o_topk); s.step = (int32_t*)(sb + o_step);
    s.counters = (int32_t*)(sb + o_cnt); s.tok_b = (int32_t*)(sb + o_tokb); s.tok_p = (int32_t*)(sb + o_tokp);
    s.seq_off = (int32_t*)(sb + o_seqoff); s.last_idx = (int32_t*)(sb + o_last); s.prompt_ids = (int32_t*)(sb + o_pids);
    s.gathered = (int32_t*)(sb + o_gath); s.gathered_len = (int32_t*)(sb + o_glen); s.top_p = (float*)(sb + o_topp);
    s.temperature = (float*)(sb + o_temp); s.rep_pen = (float*)(sb + o_rep); s.cum_log = (float*)(sb + o_cum);
    s.finished = (uint8_t*)(sb + o_fin); s.seeds = (uint64_t*)(sb + o_seeds); s.curand = sb + o_curand;
    s.parent_ids = (int32_t*)(sb + o_parent); s.cache_indir = (int32_t*)(sb + o_indir); s.beam = K;

    // ---- pinned staging: [lens B][pad B][seq_len B][ks B][ps B][temps B][reps B][seeds 2B][tok_b T][tok_p T][seq_off B+1][last B][step 1]
    const size_t stage_ints = (size_t)10 * B + 2 * (size_t)std::max(T, 1) + (B + 1) + B + 1 + 2 * (size_t)B + 16;
    if (stage_ints * 4 > e->host_stage_cap) {
        if (e->host_stage) cudaFreeHost(e->host_stage);
        e->host_stage = nullptr;
        FTCF_CUDA_CHECK(cudaHostAlloc(&e->host_stage, stage_ints * 4, cudaHostAllocDefault));
        e->host_stage_cap = stage_ints * 4;
    }
    int32_t* hs = e->host_stage;
    size_t hp = 0;
    auto up = [&](void* dst, const void* src, size_t bytes) -> int {
        memcpy(hs + hp, src, bytes);
        FTCF_CUDA_CHECK(cudaMemcpyAsync(dst, hs + hp, bytes, cudaMemcpyHostToDevice, st));
        hp += (bytes + 3) / 4;
        return FTCF_OK;
    };
    std::vector<int32_t> pad(B), seq0(B), tok_b(std::max(T, 1)), tok_p(std::max(T, 1)), seq_off(B + 1), last(B);
    {
        int tix = 0;
        seq_off[0] = 0;
        for (int b = 0; b < B; ++b) {
            pad[b] = S - lens[b];
            seq0[b] = has_prefill ? S - 1 : 0;   // invokeDecodingInitialize(max_input_length - 1), GptNeoX.cc:687-695
            if (b % K == 0)                      // rows of beams > 0 hold no prompt tokens (empty sequences for the prefill kernels)
                for (int p = 0; p < lens[b]; ++p) { tok_b[tix] = b; tok_p[tix] = p; ++tix; }
            seq_off[b + 1] = tix;
            last[b] = tix - 1;                   // every beam starts from the last prompt token of its request row
        }
    }
    const int32_t step0 = S;
    FTCF_TRY(up(s.input_len, lens.data(), B * 4));
    FTCF_TRY(up(s.pad_count, pad.data(), B * 4));
    FTCF_TRY(up(s.seq_len, seq0.data(), B * 4));
    FTCF_TRY(up(s.top_k, ks.data(), B * 4));
    FTCF_TRY(up(s.top_p, ps.data(), B * 4));
    FTCF_TRY(up(s.temperature, temps.data(), B * 4));
    FTCF_TRY(up(s.rep_pen, reps.data(), B * 4));
    FTCF_TRY(up(s.seeds, seeds.data(), B * 8));
    FTCF_TRY(up(s.tok_b, tok_b.data(), (size_t)std::max(T, 1) * 4));
    FTCF_TRY(up(s.tok_p, tok_p.data(), (size_t)std::max(T, 1) * 4));
    FTCF_TRY(up(s.seq_off, seq_off.data(), (B + 1) * 4));
    FTCF_TRY(up(s.last_idx, last.data(), B * 4));
    FTCF_TRY(up(s.step, &step0, 4));
    FTCF_CUDA_CHECK(cudaMemsetAsync(s.finished, 0, B, st));
    FTCF_CUDA_CHECK(cudaMemsetAsync(s.cum_log, 0, B * 4, st));
    if (K > 1) {
        // cum_log_probs 0 for beam 0 and -1e20 for the others: the first step only expands beam 0 (decoding_kernels.cu:24-60)
        std::vector<float> cum0(B, -1e20f);
        for (int b = 0; b < B; b += K) cum0[b] = 0.f;
        FTCF_TRY(up(s.cum_log, cum0.data(), (size_t)B * 4));
        FTCF_CUDA_CHECK(cudaMemsetAsync(s.parent_ids, 0, (size_t)max_len * B * 4, st));
        FTCF_CUDA_CHECK(cudaMemsetAsync(s.cache_indir, 0, (size_t)2 * B * max_len * 4, st));      // GptNeoX.cc:568-570
    }
    FTCF_CUDA_CHECK(cudaMemsetAsync(s.counters, 0, (size_t)B * e->Hl * 4, st));
    FTCF_CUDA_CHECK(cudaMemsetAsync(s.out_ids, 0, (size_t)max_len * B * 4, st));
    ids_to_time_major_kernel<<<ceil_div(B * S, 256), 256, 0, st>>>(s.out_ids, r.input_ids, B, S, K);
    FTCF_LAUNCH_CHECK();
    FTCF_TRY(ftcf_curand_init(s.curand, s.seeds, B, st));
    e->host_flag[0] = 0;
    e->host_flag[1] = -1;
    if ((size_t)max_len > e->host_hist_cap) {
        if (e->host_hist) cudaFreeHost(e->host_hist);
        e->host_hist = nullptr;
        e->host_hist_cap = 0;
        const size_t want = ((size_t)max_len + 1023) & ~(size_t)1023;
        FTCF_CUDA_CHECK(cudaHostAlloc(&e->host_hist, want * sizeof(int32_t), cudaHostAllocMapped));
        FTCF_CUDA_CHECK(cudaHostGetDevicePointer(reinterpret_cast<void**>(&e->host_hist_dev), e->host_hist, 0));
        e->host_hist_cap = want;
        g_capture_generation.fetch_add(1, std::memory_order_relaxed);
    }
    memset(e->host_hist, 0, (size_t)max_len * sizeof(int32_t));   // the previous request has drained (forward ends with a sync)

    if (e->tp_fused) {
        // the exchange epochs restart with the request, so the area is zeroed (every rank's earlier pushes were consumed before its
        // previous request returned).  A peer's first push of THIS request follows its prefill, whose all-reduces need this
        // rank's, which follow this memset in stream order; without a prefill a one-int all-reduce provides the same ordering.
        const size_t used = (size_t)2 * 2 * e->t * e->tpx.m_max * (e->h / 2) * 8;
        FTCF_CUDA_CHECK(cudaMemsetAsync(e->tp_area, 0, used, st));
        if (!has_prefill) FTCF_NCCL_CHECK(g_nccl.AllReduce(s.counters, s.counters, 1, /*ncclInt32*/ 2, NCCL_SUM, e->comm, st));
        e->tpx.step = s.step;
        e->tpx.step_base = has_prefill ? S + 1 : S;      // the first loop iteration that runs the layers
    }

    ftcf_sampling_params sp{};
    sp.logits = e->logits.as<float>(); sp.output_ids = s.out_ids; sp.seq_len = s.seq_len; sp.finished = s.finished;
    sp.cum_log_probs = s.cum_log; sp.input_len = s.input_len; sp.top_k = s.top_k; sp.top_p = s.top_p;
    sp.temperature = any_temp ? s.temperature : nullptr;
    sp.repetition_penalty = any_rep ? s.rep_pen : nullptr;
    sp.optional_last_tokens = r.optional_last_tokens; sp.n_last = r.optional_last_tokens ? r.n_last : 0;
    sp.stop_words = r.stop_words; sp.n_stop = r.stop_words ? r.n_stop : 0;
    sp.curand_states = s.curand; sp.step = s.step; sp.finished_count_host_mapped = e->host_flag_dev;
    sp.workspace = e->samp_ws.p;
    sp.batch = B; sp.vocab = c.vocab_size; sp.vocab_padded = e->Vp; sp.max_top_k = max_top_k;
    sp.max_input_len = S; sp.max_len = max_len; sp.end_id = c.end_id; sp.want_probs = r.return_cum_log_probs ? 1 : 0;
    sp.has_top_p_rows = any_topp ? 1 : 0;
    sp.finished_hist_host_mapped = e->host_hist_dev;

    ftcf_beam_params bp{};
    if (K > 1) {
        auto first = [](const float* a, int n, float dflt) { return a != nullptr && n > 0 ? a[0] : dflt; };
        auto differ = [](const float* a, int n) {
            for (int i = 1; a != nullptr && i < n; ++i)
                if (a[i] != a[0]) return true;
            return false;
        };
        bp.logits = e->logits.as<float>(); bp.output_ids = s.out_ids; bp.parent_ids = s.parent_ids; bp.seq_len = s.seq_len;
        bp.finished = s.finished; bp.cum_log_probs = s.cum_log; bp.input_len = s.input_len; bp.cache_indir = s.cache_indir;
        bp.stop_words = r.stop_words; bp.n_stop = r.stop_words ? r.n_stop : 0; bp.step = s.step;
        bp.finished_count_host_mapped = e->host_flag_dev; bp.finished_hist_host_mapped = e->host_hist_dev; bp.workspace = e->samp_ws.p;
        bp.batch = Bq; bp.beam_width = K; bp.vocab = c.vocab_size; bp.vocab_padded = e->Vp; bp.max_input_len = S; bp.max_len = max_len;
        bp.end_id = c.end_id;
        // element 0 of every runtime argument serves the whole batch (DynamicDecodeLayer.cc:308-408 hands the tensors down unsliced)
        bp.temperature = first(r.temperature_host, r.n_temperature, 1.f);
        bp.repetition_penalty = first(r.repetition_penalty_host, r.n_repetition_penalty, 1.f);
        bp.diversity_rate = first(r.beam_search_diversity_rate_host, r.n_beam_search_diversity_rate, 0.f);
        bp.length_penalty = first(r.len_penalty_host, r.n_len_penalty, 0.f);
        bp.args_differ = (differ(r.temperature_host, r.n_temperature) || differ(r.repetition_penalty_host, r.n_repetition_penalty) ||
                          differ(r.beam_search_diversity_rate_host, r.n_beam_search_diversity_rate) || differ(r.len_penalty_host, r.n_len_penalty))
                             ? 1 : 0;
    }
    auto token_step = [&]() -> int { return K > 1 ? ftcf_beam_search_step(&bp, st) : ftcf_sampling_step(&sp, st); };

    cudaEvent_t ev0, ev1, ev2;
    FTCF_CUDA_CHECK(cudaEventCreate(&ev0));
    FTCF_CUDA_CHECK(cudaEventCreate(&ev1));
    FTCF_CUDA_CHECK(cudaEventCreate(&ev2));
    struct EvGuard { cudaEvent_t a, b, c; ~EvGuard() { cudaEventDestroy(a); cudaEventDestroy(b); cudaEventDestroy(c); } } evg{ev0, ev1, ev2};
    FTCF_CUDA_CHECK(cudaEventRecord(ev0, st));

    // ---- prefill
    if (has_prefill) {
        gather_prompt_ids_kernel<<<ceil_div(T, 256), 256, 0, st>>>(s.prompt_ids, r.input_ids, s.tok_b, s.tok_p, T, S, K);
        FTCF_LAUNCH_CHECK();
        FTCF_TRY(ftcf_embedding_lookup(e->x.p, e->wte, s.prompt_ids, T, e->h, c.vocab_size, st));
        int max_seq = 0;
        for (int b = 0; b < B; ++b) max_seq = std::max(max_seq, lens[b]);
        const float scale = __half2float(__float2half_rn(1.f / std::sqrt((float)dh)));   // fp16 constant, GptContextAttentionLayer.cc:205
        for (int l = 0; l < L; ++l) {
            auto attn = [&](int layer) -> int {
                __half* kc = e->kv.as<__half>() + (size_t)(2 * layer) * per_layer;
                __half* vc = e->kv.as<__half>() + (size_t)(2 * layer + 1) * per_layer;
                FTCF_TRY(ftcf_prefill_qkv_rotary_scatter(e->qkv.p, e->layers[layer].qkv_b, e->qbuf.p, kc, vc, s.tok_b, s.tok_p, T, e->Hl, dh,
                                                         c.rotary_embedding_dim, max_len, st));
                return ftcf_prefill_attention(e->qbuf.p, kc, vc, e->ctx.p, s.seq_off, B, max_seq, e->Hl, dh, max_len, scale, st);
            };
            FTCF_TRY(run_layer(e, l, T, attn));
        }
        // last-token rows -> x2, then x <- x2 (invokeLookupHiddenStateOfLastToken, kernels/gpt_kernels.cu:438-470)
        FTCF_TRY(ftcf_embedding_lookup(e->x2.p, e->x.p, s.last_idx, B, e->h, T, st));
        FTCF_CUDA_CHECK(cudaMemcpyAsync(e->x.p, e->x2.p, (size_t)B * e->h * 2, cudaMemcpyDeviceToDevice, st));
    }
    FTCF_CUDA_CHECK(cudaEventRecord(ev1, st));

    // ---- decode loop (GptNeoX.cc:776-1048)
    e->last_step_ms.clear();
    std::vector<cudaEvent_t> step_ev;
    const bool timing = e->opt_step_timing != 0;
    auto rec_step = [&]() {
        if (!timing) return;
        cudaEvent_t ev;
        cudaEventCreate(&ev);
        cudaEventRecord(ev, st);
        step_ev.push_back(ev);
    };
    rec_step();

    const bool want_trace = r.logits_trace != nullptr && r.logits_trace_steps > 0;
    const bool use_graph = e->opt_cuda_graph != 0 && out_len > 2 && !want_trace;
    char keybuf[400];
    snprintf(keybuf, sizeof(keybuf), "B%d K%d:%a:%a:%a:%a:%d S%d M%d k%d t%d r%d p%d l%p s%p n%d/%d kv%p x%p sm%p lg%p f%d g%lld", B, K,
             bp.temperature, bp.repetition_penalty, bp.diversity_rate, bp.length_penalty, bp.args_differ, S, max_len, max_top_k,
             (int)any_temp, (int)any_rep + 2 * (int)any_topp, sp.want_probs, (const void*)sp.optional_last_tokens, (const void*)sp.stop_words,
             sp.n_last, sp.n_stop, e->kv.p, e->x.p, e->small.p, e->logits.p, (int)e->fused_on + 2 * (int)(e->tp_fused && e->opt_tp_fused),
             g_capture_generation.load(std::memory_order_relaxed));
    const std::string key(keybuf);

    constexpr int kExitLag = 2;
    int steps_done = 0;
    std::vector<int32_t> last_seq(B, -1);
    std::vector<int32_t> cb_tok(B), cb_idx(B), cb_seq(B);
    for (int step = S; step < max_len; ++step) {
        const bool run_layers = !(has_prefill && step == S);
        if (use_graph && run_layers) {
            size_t gi = 0;
            while (gi < e->graphs.size() && e->graphs[gi].key != key) ++gi;
            if (gi == e->graphs.size()) {
                cudaGraph_t g = nullptr;
                const long long n0 = g_launch_count.load();
                FTCF_CUDA_CHECK(cudaStreamBeginCapture(st, cudaStreamCaptureModeThreadLocal));
                int rc = decode_step(e, s, sp, B, max_len, S, splits, true);
                if (rc == FTCF_OK) rc = token_step();
                cudaError_t ce = cudaStreamEndCapture(st, &g);
                if (rc != FTCF_OK) { if (g) cudaGraphDestroy(g); return rc; }
                FTCF_CUDA_CHECK(ce);
                ftcf_gptneox::CachedGraph cg;
                cg.key = key;
                cg.nodes = g_launch_count.load() - n0;
                g_launch_count.fetch_sub(cg.nodes);   // captured, not launched yet
                cudaError_t ie = cudaGraphInstantiate(&cg.exec, g, 0);
                cudaGraphDestroy(g);
                FTCF_CUDA_CHECK(ie);
                if (e->graphs.size() >= ftcf_gptneox::kMaxGraphs) {
                    cudaGraphExecDestroy(e->graphs.back().exec);
                    e->graphs.pop_back();
                }
                e->graphs.insert(e->graphs.begin(), cg);
            } else if (gi != 0) {
                std::rotate(e->graphs.begin(), e->graphs.begin() + gi, e->graphs.begin() + gi + 1);
            }
            FTCF_CUDA_CHECK(cudaGraphLaunch(e->graphs[0].exec, st));
            g_launch_count.fetch_add(e->graphs[0].nodes);
        } else {
            FTCF_TRY(decode_step(e, s, sp, B, max_len, S, splits, run_layers));
            if (want_trace && steps_done < r.logits_trace_steps) {   // raw logits, before the sampler edits them in place
                for (int b = 0; b < B; ++b)
                    FTCF_CUDA_CHECK(cudaMemcpyAsync(r.logits_trace + ((size_t)steps_done * B + b) * c.vocab_size,
                                                    e->logits.as<float>() + (size_t)b * e->Vp, (size_t)c.vocab_size * 4,
                                                    cudaMemcpyDeviceToDevice, st));
            }
            FTCF_TRY(token_step());
        }
        ++steps_done;
        rec_step();

        const bool last_iter = step + 1 >= max_len;
        if (r.callback && !last_iter) {
            // streaming: the reference synchronises every token here too (pybind_callback_utils.cc:36-76)
            FTCF_CUDA_CHECK(cudaMemcpyAsync(cb_tok.data(), s.out_ids + (size_t)step * B, B * 4, cudaMemcpyDeviceToHost, st));
            FTCF_CUDA_CHECK(cudaMemcpyAsync(cb_seq.data(), s.seq_len, B * 4, cudaMemcpyDeviceToHost, st));
            FTCF_CUDA_CHECK(cudaStreamSynchronize(st));
            for (int b = 0; b < B; ++b) {
                if (cb_seq[b] == last_seq[b]) cb_tok[b] = c.end_id;
                else last_seq[b] = cb_seq[b];
                cb_idx[b] = cb_seq[b] - S;
            }
            if (e->rank == 0) r.callback(r.callback_user, step, cb_tok.data(), cb_idx.data(), B);
            if (e->host_flag[0] >= B) break;
        } else if (!last_iter && step - kExitLag >= S) {
            // Early exit without a per-token stream sync, and identical on every tensor-parallel rank: the decision for loop
            // iteration `step` is taken on the finished count the sampler published FOR step - kExitLag (a step-indexed value
            // in mapped pinned memory), which the host waits for -- it is kExitLag steps behind the launches, so the wait is
            // normally over before it starts, and it bounds the run-ahead of the host.  Every rank therefore launches exactly
            // (first all-finished step + kExitLag) steps and the same number of NCCL collectives.  Running kExitLag extra
            // steps after every row finished only appends end_id to rows that no longer advance.
            const volatile int32_t* hist = e->host_hist;
            const int chk = step - kExitLag;
            for (long long spin = 0; hist[chk] == 0; ++spin) {
                if ((spin & 0xfff) == 0xfff) {
                    const cudaError_t q = cudaStreamQuery(st);
                    if (q == cudaSuccess) break;                  // drained: the value is there now (or the step never ran)
                    FTCF_REQUIRE(q == cudaErrorNotReady, FTCF_ERR_CUDA, "forward: CUDA error during decode: %s", cudaGetErrorString(q));
                }
            }
            if (hist[chk] - 1 >= B) break;
        }
    }
    FTCF_CUDA_CHECK(cudaEventRecord(ev2, st));

    // ---- outputs (setOutputTensors, GptNeoX.cc:1090-1181)
    if (K > 1)
        FTCF_TRY(ftcf_gather_output_beams(r.output_ids, r.sequence_lengths, s.out_ids, s.parent_ids, s.seq_len, s.input_len, Bq, K, S, max_len,
                                          c.end_id, st));
    else
        FTCF_TRY(ftcf_gather_output(r.output_ids, r.sequence_lengths, s.out_ids, s.seq_len, s.input_len, B, S, max_len, c.end_id, st));
    if (r.cum_log_probs) FTCF_CUDA_CHECK(cudaMemcpyAsync(r.cum_log_probs, s.cum_log, B * 4, cudaMemcpyDeviceToDevice, st));
    FTCF_CUDA_CHECK(cudaStreamSynchronize(st));
    {
        cudaError_t le = cudaGetLastError();
        FTCF_REQUIRE(le == cudaSuccess, FTCF_ERR_CUDA, "forward: CUDA error after the request: %s", cudaGetErrorString(le));
    }
    if (timing) {
        for (size_t i = 1; i < step_ev.size(); ++i) {
            float ms = 0.f;
            cudaEventElapsedTime(&ms, step_ev[i - 1], step_ev[i]);
            e->last_step_ms.push_back(ms);
        }
    }
    for (auto ev : step_ev) cudaEventDestroy(ev);
    if (stats) {
        stats->steps = steps_done;
        cudaEventElapsedTime(&stats->prefill_ms, ev0, ev1);
        cudaEventElapsedTime(&stats->decode_ms, ev1, ev2);
        stats->kernel_launches = g_launch_count.load() - launches0;
    }
    return FTCF_OK;
}
