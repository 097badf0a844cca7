/*
 * ftcf.h -- C ABI of libftcf.so, the B200 (sm_100a) engine behind the CodeFuse / GPT-NeoX path.
 *
 * This is the drop-in boundary below the reference's Python surface: the two pybind11 modules
 * `libth_gptneox` (GptNeoXOp, reference src/fastertransformer/th_op/gptneox/GptNeoXOp.cc:190-212) and
 * `libth_common` (symmetric_quantize_last_axis_of_batched_matrix_int8, reference
 * src/fastertransformer/th_op/common/WeightOnlyQuantOps.cc:344-349) are thin torch-facing shims over the
 * entry points declared here.  Plain pointers and sizes only; no torch / C++ types cross this line.
 *
 * Conventions
 *   - every function returns 0 (FTCF_OK) or a non-zero ftcf_status; ftcf_last_error() gives the message
 *     (thread-local).  Nothing here falls back to a CPU path: without a usable sm_100 device the calls fail.
 *   - `stream` is a cudaStream_t passed as void*.
 *   - fp16 buffers are passed as void* (IEEE binary16, row-major); ids / lengths are int32.
 *   - device pointers unless the name ends in _host.
 *
 * Each entry cites the reference interface it stands in for (paths relative to
 * /root/reference/src/fastertransformer).
 */
#ifndef FTCF_H_
#define FTCF_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef enum {
    FTCF_OK = 0,
    FTCF_ERR_INVALID = 1,     /* bad argument (the binding turns it into a Python RuntimeError)       */
    FTCF_ERR_CUDA = 2,        /* CUDA runtime / driver error                                          */
    FTCF_ERR_NCCL = 3,        /* NCCL error                                                           */
    FTCF_ERR_UNSUPPORTED = 4  /* shape or feature outside what the sm_100a kernels implement          */
} ftcf_status;

const char* ftcf_last_error(void);
/* Library/ABI version, bumped when a signature changes. */
int ftcf_abi_version(void);
/* 0 when the current device is compute capability 10.x and the kernels in this build can run on it. */
int ftcf_device_check(void);
/* Kernels (and NCCL collectives) launched by this library in this process so far. */
long long ftcf_launch_count(void);
/* Process-wide tuning knobs (defaults are the tuned values): "pdl" 0/1 programmatic dependent launch,
 * "skinny_target_ctas" / "decode_target_ctas" CTAs per streaming / tcgen05 decode-GEMM launch, "decode_impl" 3 (tcgen05) or 1
 * (streaming mma.sync kernel) for the INT8 decode GEMMs, "mmha_bulk" 0/1, ... (the full list is in csrc/capi.cu). */
int ftcf_set_tunable(const char* name, int value);

/* Debug: per-CTA timeline of the decode kernels.  Between start and stop every instrumented kernel appends one 56-byte record
 * per CTA {u64 t_start, t_after_dependency_wait, t_first_data, t_end (globaltimer ns); i32 kind, cta, n_cta, a, b, pad}
 * (kind: 1 int8 GEMM, 2 fp16 GEMM [a = n, b = k], 10 decode attention, 20 LayerNorm, 21 residual, 30 sampling).
 * tools/trace_step.py turns it into a per-launch timeline of one decode step. */
int ftcf_debug_trace_start(unsigned capacity);
int ftcf_debug_trace_stop(void* out_host, unsigned max_records, unsigned* n);
/* Debug: device buffer of 64 x 8 int64 (or NULL to remove) that CTA (0,0,0) of every decode-GEMM launch fills with per-K-step
 * clock stamps: converter warp [0] before / [1] after the wait for the weight stage, [2] after the conversion, [3] after the
 * wait for a free TMEM stage, [4] after tcgen05.st; MMA warp [5] before / [6] after its wait for the TMEM stage, [7] after
 * issuing the MMAs and commits (tools/decode_gemm_probe.py). */
int ftcf_debug_decode_probe(void* dev_buf);

/* ------------------------------------------------------------------------------------------------
 * Weight-only INT8 quantiser (CPU).  Replaces ft::symmetric_quantize<half,half|float> +
 * preprocess_weights_for_mixed_gemm, kernels/cutlass_kernels/cutlass_preprocessors.cc:577-673,500-539,
 * reached through th_op/common/WeightOnlyQuantOps.cc:140-233.
 *   weight      [e, k, n] (e = 1 for a 2-D matrix), dtype: 0 = fp32, 1 = fp16, 2 = bf16, host memory
 *   processed   e*k*n bytes: B200 layout = W^T, i.e. [e][n][k] with k contiguous, value q + 128 as uint8
 *   unprocessed optional (may be NULL) plain int8 [e, k, n]
 *   scales      [e, n] in the weight's dtype (absmax / 128, rounded to that dtype)
 * The same rounding as the reference: q = clip(round_half_away(w / scale_fp32), -128, 127).
 * ---------------------------------------------------------------------------------------------- */
int ftcf_symmetric_quantize_int8_host(const void* weight_host, int dtype, size_t e, size_t k, size_t n,
                                      uint8_t* processed_host, int8_t* unprocessed_host, void* scales_host);
/* Layout helpers (CPU): plain int8 [k,n] <-> B200 layout, and the reference's sm80 ("Ampere") *.q.bin layout
 * (cutlass_preprocessors.cc:133-201,207-348,350-370,437-498) -> B200 layout, for checkpoints made by the
 * reference's quant_and_save.py. */
int ftcf_int8_plain_to_b200_host(const int8_t* q_kn, size_t k, size_t n, uint8_t* out_nk);
int ftcf_int8_ampere_to_b200_host(const int8_t* processed_ampere, size_t k, size_t n, uint8_t* out_nk);

/* ------------------------------------------------------------------------------------------------
 * Kernels.  One launcher per kernel family; all asynchronous on `stream`.
 * ---------------------------------------------------------------------------------------------- */

/* y[m,n] = act(x[m,k] . dequant(W)[k,n] + bias), W given as W^T uint8 [n,k] (value q+128), per-column fp16 scale.
 * fp32 accumulate, scale applied in the epilogue, fp16 out.  act: 0 none, 1 tanh-GELU.  bias may be NULL.
 * Replaces CutlassFpAIntBGemmRunner<half,uint8_t>::gemm / gemm_bias_act,
 * kernels/cutlass_kernels/fpA_intB_gemm/fpA_intB_gemm_template.h:461-570.
 * impl: 0 = auto, 1 = force the skinny (m <= 32 streaming) kernel, 2 = force the tcgen05 kernel. */
int ftcf_gemm_w8a16(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n,
                    int k, int act, int impl, void* stream);
/* Same, with a launch hint (optional, NULL = defaults; never changes results):
 *   target_ctas  CTAs the launch should aim for (0: automatic).  The decode layer runs its two branches side by side and gives
 *                each GEMM a share of the GPU's CTA slots, so that neither branch queues behind the other;
 *   no_pdl       1: launch without the programmatic-dependent-launch attribute (the kernel then starts after its predecessor
 *                has finished instead of parking its CTAs on the SMs while it waits);
 *   stages       see below. */
typedef struct {
    int32_t target_ctas, no_pdl;
    int32_t stages;   /* tcgen05 decode GEMM: depth of the 16 KB weight ring (0: automatic = as deep as lets two CTAs share an SM;
                         deeper rings take the SM for one CTA -- for a GEMM that runs beside small kernels, not beside another GEMM) */
} ftcf_launch_hint;
int ftcf_gemm_w8a16_ex(const void* x, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m, int n,
                       int k, int act, int impl, const ftcf_launch_hint* hint, void* stream);

/* y[m,n] = x[m,k] . W, W given K-major as W^T fp16 [n,k]; fp32 accumulate.  out_f32 = 1 writes fp32 (LM head
 * logits, models/gptneox/GptNeoX.cc:869-912), else fp16 with optional bias + tanh-GELU applied with the
 * reference's fp16 rounding points (layers/FfnLayer.cc:294-309, kernels/activation_kernels.cu:50-72).
 * ldy = row pitch of y in elements.  Replaces cublasMMWrapper::Gemm, utils/cublasMMWrapper.cc:154-328. */
int ftcf_gemm_f16(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                  int out_f32, int impl, void* stream);
int ftcf_gemm_f16_ex(const void* x, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy, int act,
                     int out_f32, int impl, const ftcf_launch_hint* hint, void* stream);

/* Decode-step variants (m <= 4 rows) with the layer's glue fused in as a prologue: every CTA first builds its input
 *   r = ((add_ffn + add_attn) + add_bias) + x     -- the previous layer's parallel-residual add (skipped when add_ffn is NULL)
 *   a = LayerNorm(r; gamma, beta, eps)            -- fp32 statistics, fp16 normalisation
 * in shared memory and the GEMM runs on `a`; x_out (optional) receives r.  Same arithmetic as
 * ftcf_add_bias_attn_ffn_residual (tp = 1) followed by ftcf_layernorm (kernels/add_residual_kernels.cu:116-176,
 * kernels/layernorm_kernels.cu:158-286), without their two launches on the layer's critical path. */
/* Tensor-parallel exchange fused into the decode GEMMs: replaces the residual kernel + ncclAllReduce of
 * models/gptneox/GptNeoXDecoder.cc:348-359 (and the one-shot kernel of kernels/custom_ar_kernels.cu:139,202) at decode sizes.
 * Every rank owns an exchange area that all ranks of the node map (CUDA IPC over NVLink), made of 8-byte "flagged words":
 *   area [2 slots][2 kinds: 0 attn, 1 ffn][tp source ranks][m_max][h / 2] x { fp16 pair, uint32 epoch }
 * Push side (O / FFN2 GEMM of rank r): the epilogue stores each output tile into (slot, kind, source r) of EVERY rank's area --
 * remote 8-byte stores from the kernel that computed the tile.  Data and flag travel in ONE store, so there is no fence and no
 * separate flag write on the critical path (a system-scope release after remote stores costs 6-9 us beside a streaming GEMM).
 * Gather side (ftcf_tp_gather_residual, one small kernel heading each branch of the next layer): polls the words it needs until
 * their epoch is the expected one, rebuilds every rank's partial  o_r = ((ffn_r + attn_r) + bias) + half(x / tp)  with the reference's
 * fp16 adds (kernels/add_residual_kernels.cu:116-176), sums the tp partials in rank order in fp32 and rounds once: the all-reduce.
 * The exchange index g = (*step - step_base) * layer_num + layer is evaluated on the device (one captured graph serves every
 * token): slot = g & 1, epoch = g / 2 + 1.  The area is zeroed at the start of a request (epoch 0 never matches). */
typedef struct {
    void* peer_data[8];            /* exchange area of every rank, own rank included (device pointers valid in this process) */
    int32_t tp, rank, m_max, h;
    const int32_t* step;           /* device scalar: the decode loop's step */
    int32_t step_base, layer_num;
} ftcf_tp_exchange;

typedef struct {
    const void *x, *add_ffn, *add_attn, *add_bias, *gamma, *beta;   /* fp16: [m,k] [m,k] [m,k] [k] [k] [k] */
    void* x_out;                                                    /* [m,k] fp16 or NULL; must not alias x */
    float eps;
    int32_t cta_hint;   /* 0: automatic; > 0: CTAs this launch should aim for (the engine gives the two GEMMs that start a layer
                           together half of the SM slots each, so that neither queues behind the other) */
} ftcf_ln_prologue;
int ftcf_gemm_w8a16_ln(const ftcf_ln_prologue* pro, const uint8_t* w_nk, const void* scale, const void* bias, void* y, int m,
                       int n, int k, int act, void* stream);
int ftcf_gemm_f16_ln(const ftcf_ln_prologue* pro, const void* w_nk, const void* bias, void* y, int m, int n, int k, int ldy,
                     int act, int out_f32, void* stream);
/* INT8 GEMM whose epilogue pushes the [m, n = h] output into every rank's exchange area (kind 0: O GEMM, 1: FFN2 GEMM) instead
 * of writing y.  m <= ex->m_max, n == ex->h, k a multiple of 128. */
int ftcf_gemm_w8a16_tp_push(const void* x, const uint8_t* w_nk, const void* scale, const ftcf_tp_exchange* ex, int kind, int layer,
                            int m, int n, int k, const ftcf_launch_hint* hint, void* stream);
/* Stand-alone gather side: x_out[m,h] = all-reduced residual of exchange `layer` (see ftcf_tp_exchange); x is that layer's input. */
int ftcf_tp_gather_residual(const ftcf_tp_exchange* ex, int layer, const void* x, const void* bias, void* x_out, int m, void* stream);

/* out[k,n] -> out_t[n,k] fp16 transpose (load-time re-layout of fp16 weights to K-major). */
int ftcf_transpose_f16(const void* in_kn, void* out_nk, int k, int n, void* stream);

/* LayerNorm, fp32 statistics and fp16 normalisation as kernels/layernorm_kernels.cu:158-286 (invokeGeneralLayerNorm
 * :1653-1735).  Optional fused pre-add: if `residual_out` != NULL first computes
 * residual_out = x + add1 (+ add_bias) in fp16 (invokeGeneralAddBiasResidualPreLayerNorm), then normalises it. */
int ftcf_layernorm(const void* x, const void* gamma, const void* beta, void* y, int m, int n, float eps, void* stream);
int ftcf_add_bias_residual_layernorm(const void* x, const void* add1, const void* add_bias, void* residual_out,
                                     const void* gamma, const void* beta, void* y, int m, int n, float eps,
                                     void* stream);

/* out = (half)(x / tp) + ffn + attn + bias   (fp16 adds, kernels/add_residual_kernels.cu:116-176)
 * and  out = x + y + bias                     (invokeAddBiasResidual, sequential-residual path). */
int ftcf_add_bias_attn_ffn_residual(void* out, const void* ffn, const void* attn, const void* x, const void* bias,
                                    int m, int n, int tp, void* stream);
int ftcf_add_bias_residual(void* out, const void* y, const void* x, const void* bias, int m, int n, void* stream);

/* Row gather from the embedding table (kernels/gpt_kernels.cu:32-105, decoding_kernels.cu:260).
 * ids[i] for i < m; ids_stride lets the caller point into a time-major id buffer. */
int ftcf_embedding_lookup(void* out, const void* table, const int32_t* ids, int m, int n, int vocab, void* stream);

/* Decode attention for one new token per sequence (masked_multihead_attention_kernel,
 * kernels/decoder_masked_multihead_attention/decoder_masked_multihead_attention_template.hpp:1099-1919;
 * parameter block kernels/decoder_masked_multihead_attention.h:51-158). */
typedef struct {
    const void* qkv;            /* [B, 3*heads*dh] fp16: q | k | v, each [heads, dh]                          */
    const void* qkv_bias;       /* [3*heads*dh] fp16 or NULL                                                  */
    void* k_cache;              /* [B, heads, max_len, dh] fp16 (this layer, this rank)                       */
    void* v_cache;              /* [B, heads, max_len, dh] fp16                                               */
    void* ctx;                  /* [B, heads*dh] fp16 out                                                     */
    const int32_t* seq_len;     /* [B] cache slot of the new token == number of earlier slots                 */
    const int32_t* input_len;   /* [B] prompt lengths: slots [input_len, max_input_len) are the masked pad gap */
    const int32_t* pad_count;   /* [B] total_padding_tokens (rotary position = timestep - pad_count)          */
    const uint8_t* finished;    /* [B] or NULL; finished rows are skipped (template.hpp:1176-1178)            */
    const int32_t* step;        /* device scalar: the loop's `step`; timestep = step - 1                      */
    float* partial;             /* split-KV scratch [B*heads*splits*(dh+2)] fp32 (unused when splits == 1)    */
    int32_t* counters;          /* [B*heads] zero-initialised, self-resetting                                 */
    int32_t batch, heads, dh, rotary_dim, max_len, max_input_len, splits;
    float inv_sqrt_dh;
    /* beam search (NULL / 0 otherwise): batch counts batch x beam rows; slot t < seq_len of row r is read from cache row
     * (r / beam_width) * beam_width + cache_indir[parity][r][t], parity = (*step - max_input_len) & 1 -- the buffer the previous
     * beam-search step wrote (template.hpp:1494-1522,1709-1761; double buffer GptNeoX.cc:118-122,777-778) */
    const int32_t* cache_indir; /* [2, B, max_len] */
    int32_t beam_width;
} ftcf_mmha_params;
int ftcf_mmha_decode(const ftcf_mmha_params* p, void* stream);
/* Optional companion of ftcf_mmha_decode: L2 prefetch of the cache rows that launch will read (same params; only the cache
 * pointers, lengths and flags are used).  Meant for a side stream at the start of the layer; never changes results. */
int ftcf_mmha_prefetch_cache(const ftcf_mmha_params* p, void* stream);
/* Scratch sizing / split choice for ftcf_mmha_decode. */
int ftcf_mmha_choose_splits(int batch, int heads, int max_len);

/* Prefill: qkv + bias, NeoX rotary at the token's position, q kept, k/v scattered into the cache
 * (add_fusedQKV_bias_transpose_kernel + transpose_4d_batch_major_{k,v}_cache,
 * kernels/unfused_attention_kernels.cu:1326-1484,1673-1757), on padding-removed tokens. */
int ftcf_prefill_qkv_rotary_scatter(const void* qkv, const void* qkv_bias, void* q_out, void* k_cache, void* v_cache,
                                    const int32_t* tok_batch, const int32_t* tok_pos, int tokens, int heads, int dh,
                                    int rotary_dim, int max_len, void* stream);
/* Causal attention over each sequence's own prompt (replaces the unfused QK^T / softmax / PV chain,
 * layers/attention_layers/GptContextAttentionLayer.cc:194-300).  q [T, heads, dh]; ctx [T, heads*dh]. */
int ftcf_prefill_attention(const void* q, const void* k_cache, const void* v_cache, void* ctx,
                           const int32_t* seq_offsets /* [B+1] */, int batch, int max_seq, int heads, int dh,
                           int max_len, float scale, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Sampling stack on fp32 logits (layers/DynamicDecodeLayer.cc:192-495, sampling_layers/TopKSamplingLayer.cu).
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float* logits;                 /* [B, vocab_padded] fp32, modified in place                              */
    int32_t* output_ids;           /* [max_len, B] time-major                                                */
    int32_t* seq_len;              /* [B]                                                                    */
    uint8_t* finished;             /* [B]                                                                    */
    float* cum_log_probs;          /* [B] or NULL                                                            */
    const int32_t* input_len;      /* [B]                                                                    */
    const int32_t* top_k;          /* [B] runtime k (already through the setup rules); 0 = pure top-p row    */
    const float* top_p;            /* [B]                                                                    */
    const float* temperature;      /* [B] or NULL (NULL == all 1)                                            */
    const float* repetition_penalty; /* [B] or NULL                                                          */
    const int32_t* optional_last_tokens; /* [B, n_last] (-1 padded) or NULL; applied when step == max_input_len */
    const int32_t* stop_words;     /* [B, 2, n_stop] or NULL                                                 */
    void* curand_states;           /* [B] curandState_t                                                      */
    int32_t* step;                 /* device scalar, read; incremented by ftcf_sampling_advance              */
    int32_t* finished_count_host_mapped; /* device-visible pinned int (or NULL): #finished after this step   */
    void* workspace;               /* ftcf_sampling_workspace_bytes()                                        */
    int32_t batch, vocab, vocab_padded, max_top_k, n_last, n_stop, max_input_len, max_len, end_id;
    int32_t want_probs;            /* 1: softmax before top-k and accumulate cum_log_probs                   */
    int32_t has_top_p_rows;        /* 1: some row has top_k == 0 (pure top-p, layers/sampling_layers/TopPSamplingLayer.cu) */
    int32_t* finished_hist_host_mapped;  /* device-visible pinned int[max_len] (or NULL): entry [step] = 1 + #finished after
                                            that step (0 = not written yet).  Lets the host take the early-exit decision on a
                                            step-indexed value, i.e. identically on every tensor-parallel rank, without the
                                            per-token stream sync of kernels/stop_criteria_kernels.cu:135-156 */
} ftcf_sampling_params;
size_t ftcf_sampling_workspace_bytes(int batch, int vocab_padded, int max_top_k);
size_t ftcf_curand_state_bytes(void);
/* curand_init(seed[b], 0, 0) per row (kernels/sampling_topk_kernels.cu:32-55). */
int ftcf_curand_init(void* states, const uint64_t* seeds, int batch, void* stream);
/* One decoding step of logit post-processing + top-k sampling + stop criteria; advances *step by one at the end. */
int ftcf_sampling_step(const ftcf_sampling_params* p, void* stream);
/* gatherTree for beam 1: time-major ids -> [B, max_len] with the pad gap removed (kernels/decoding_kernels.cu:452-580). */
int ftcf_gather_output(int32_t* out /* [B, max_len] */, int32_t* out_len /* [B] */, const int32_t* ids_time_major,
                       const int32_t* seq_len, const int32_t* input_len, int batch, int max_input_len, int max_len,
                       int end_id, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Online beam search (beam_width > 1): layers/DynamicDecodeLayer.cc:308-408 -> layers/beam_search_layers/BaseBeamSearchLayer.cu:170-285
 * -> OnlineBeamSearchLayer.cu:83-170.  Rows are batch x beam ("B*K").
 * ---------------------------------------------------------------------------------------------- */
typedef struct {
    float* logits;                 /* [B*K, vocab_padded] fp32, modified in place by the penalties                   */
    int32_t* output_ids;           /* [max_len, B*K] time-major                                                      */
    int32_t* parent_ids;           /* [max_len, B*K] time-major: the beam slot each token's history continues in     */
    int32_t* seq_len;              /* [B*K]                                                                          */
    uint8_t* finished;             /* [B*K]                                                                          */
    float* cum_log_probs;          /* [B*K]; initial state 0 for beam 0, -1e20 for the others (decoding_kernels.cu:24-60) */
    const int32_t* input_len;      /* [B*K] (tiled)                                                                  */
    int32_t* cache_indir;          /* [2, B*K, max_len], zero-initialised; see ftcf_mmha_params                      */
    const int32_t* stop_words;     /* [B, 2, n_stop] or NULL                                                         */
    int32_t* step;                 /* device scalar, read; advanced by one at the end                                */
    int32_t* finished_count_host_mapped;  /* as in ftcf_sampling_params (counts rows, i.e. up to B*K)                */
    int32_t* finished_hist_host_mapped;
    void* workspace;               /* ftcf_beam_workspace_bytes()                                                    */
    int32_t batch, beam_width, vocab, vocab_padded, n_stop, max_input_len, max_len, end_id;
    float temperature, repetition_penalty, diversity_rate, length_penalty;
    int32_t args_differ;           /* 1: the runtime arguments differ between rows; the reference then runs the batches one by
                                      one, which only changes which row's length the length penalty looks at (b * K instead of b) */
} ftcf_beam_params;
size_t ftcf_beam_workspace_bytes(int batch, int beam_width, int vocab_padded, int max_len);
/* One decoding step: penalties (kernels/beam_search_penalty_kernels.cu), log-softmax + 2K candidates per row, K winners per
 * batch (kernels/online_softmax_beamsearch_kernels.cu), update of ids / parents / lengths / finished flags, cache indirection,
 * stop words (kernels/stop_criteria_kernels.cu:24-84), finished count; *step += 1. */
int ftcf_beam_search_step(const ftcf_beam_params* p, void* stream);
/* gatherTree with parents (kernels/decoding_kernels.cu:452-580): out [B, K, max_len], out_len [B, K]. */
int ftcf_gather_output_beams(int32_t* out, int32_t* out_len, const int32_t* ids_time_major, const int32_t* parent_ids,
                             const int32_t* seq_len, const int32_t* input_len, int batch, int beam_width, int max_input_len,
                             int max_len, int end_id, void* stream);

/* ------------------------------------------------------------------------------------------------
 * Model-level engine: the request loop of ft::GptNeoX<T>::forward (models/gptneox/GptNeoX.cc:385-1052) with
 * GptNeoXContextDecoder (…ContextDecoder.cc:223-512) and GptNeoXDecoder (…Decoder.cc:197-389) underneath.
 * ---------------------------------------------------------------------------------------------- */
typedef struct ftcf_gptneox ftcf_gptneox;

typedef struct {
    int32_t head_num, size_per_head, inter_size, layer_num, vocab_size, rotary_embedding_dim;
    int32_t start_id, end_id;
    int32_t tensor_para_size, tensor_para_rank;
    int32_t int8_mode;            /* 0: fp16 weights, 1: weight-only int8                                    */
    int32_t use_gptj_residual;    /* 1: parallel residual                                                    */
    float layernorm_eps;          /* 1e-5 (models/gptneox/GptNeoX.h:42-43)                                   */
    int32_t int8_layout;          /* 0: B200 layout (ours), 1: plain int8 [k,n], 2: the reference's sm80 *.q.bin layout (1, 2: re-laid out at load) */
} ftcf_gptneox_config;

/* weights: 12*L+4 fp16 device pointers in GptNeoXOp order (th_op/gptneox/GptNeoXOp.h:121-174); int8_weights and
 * scales: 4*L each ({qkv,o,ffn1,ffn2}*L + layer), NULL entries allowed when int8_mode == 0.
 * nccl_unique_id: 128 bytes from ftcf_nccl_unique_id (same on all ranks) or NULL when tensor_para_size == 1. */
int ftcf_gptneox_create(ftcf_gptneox** out, const ftcf_gptneox_config* cfg, const void* const* weights,
                        size_t n_weights, const void* const* int8_weights, const void* const* scales, size_t n_int8,
                        const void* nccl_unique_id, void* stream);
void ftcf_gptneox_destroy(ftcf_gptneox* h);
int ftcf_nccl_unique_id(void* out128);

typedef void (*ftcf_token_callback)(void* user, int32_t step, const int32_t* last_tokens_host /* [B * beam] */,
                                    const int32_t* idxs_host /* [B * beam] */, int32_t rows /* B * beam */);
typedef struct {
    const int32_t* input_ids;       /* device [B, S]                                                         */
    const int32_t* input_lengths;   /* device [B]                                                            */
    int32_t batch, max_input_len, output_len;
    int32_t beam_width;             /* 0 or 1: sampling; > 1: beam search (see ftcf_gptneox_forward)          */
    /* sampling arguments on the HOST, each either NULL, 1 element or B elements (n_* gives the count)       */
    const int32_t* top_k_host;  int32_t n_top_k;
    const float* top_p_host;    int32_t n_top_p;
    const float* temperature_host; int32_t n_temperature;
    const float* repetition_penalty_host; int32_t n_repetition_penalty;
    const int64_t* random_seed_host; int32_t n_random_seed;
    /* beam search only (beam_width > 1): element 0 applies to the whole batch, as in the reference (DynamicDecodeLayer.cc:308-408) */
    const float* beam_search_diversity_rate_host; int32_t n_beam_search_diversity_rate;
    const float* len_penalty_host; int32_t n_len_penalty;
    const int32_t* stop_words;      /* device [B, 2, n_stop] or NULL */ int32_t n_stop;
    const int32_t* optional_last_tokens; /* device [B, n_last] or NULL */ int32_t n_last;
    int32_t return_cum_log_probs;
    ftcf_token_callback callback;   /* may be NULL */ void* callback_user;
    /* outputs (device) */
    int32_t* output_ids;            /* [B, beam, S + output_len]                                             */
    int32_t* sequence_lengths;      /* [B, beam]                                                             */
    float* cum_log_probs;           /* [B, beam] or NULL                                                     */
    /* optional debug taps (device, may be NULL): fp32 logits of every step [steps, B, vocab]                */
    float* logits_trace;  int32_t logits_trace_steps;
} ftcf_gptneox_request;

typedef struct {
    int32_t steps;                  /* decode-loop iterations executed                                       */
    float prefill_ms, decode_ms;    /* CUDA-event times of the two phases                                    */
    int64_t kernel_launches;        /* kernels launched by this request                                      */
} ftcf_gptneox_stats;

int ftcf_gptneox_forward(ftcf_gptneox* h, const ftcf_gptneox_request* req, ftcf_gptneox_stats* stats);
/* Per-step decode times (ms) of the last request, up to n entries; returns the count written. */
int ftcf_gptneox_last_step_ms(ftcf_gptneox* h, float* out, int n);
/* Engine options: "cuda_graph" (0/1), "gemm_impl" (0 auto / 1 skinny / 2 tcgen05), "step_timing" (0/1). */
int ftcf_gptneox_set_option(ftcf_gptneox* h, const char* name, int value);

#ifdef __cplusplus
}
#endif
#endif /* FTCF_H_ */
