// TEST INFRASTRUCTURE.  C entry points over the REFERENCE's own CUDA kernels, compiled (oracle/Makefile, target
// _ref/libref_kernels.so) from the sources where they lie under /root/reference for sm_100a:
//   kernels/layernorm_kernels.cu                invokeGeneralLayerNorm<half>                (K4)
//   kernels/add_residual_kernels.cu             invokeAddBiasAttentionFfnResidual<half>     (K5)
//   kernels/decoder_masked_multihead_attention/ mmha_launch_kernel<uint16_t, 128 | 64>      (K1), parameters filled exactly as
//                                               fusedQKV_masked_attention_dispatch does (layers/attention_layers/
//                                               DecoderSelfAttentionLayer.cc:36-146, neox_rotary_style = true, q_scaling = 1)
//   kernels/sampling_topk_kernels.cu            invokeCurandBatchInitialize, invokeAddBiasEndMask, invokeBatchTopKSampling (K12)
//   kernels/sampling_topp_kernels.cu            invokeAddBiasSoftMax
//   kernels/sampling_penalty_kernels.cu         invokeBatchApplyTemperaturePenalty, invokeBatchApplyRepetitionPenalty
//   kernels/sampling_topp_kernels.cu            invokeTopPInitialize + invokeBatchTopPSampling (pure top-p rows)
//   kernels/stop_criteria_kernels.cu            invokeStopWordsCriterion
//   kernels/unfused_attention_kernels.cu        invokeAddFusedQKVBiasTranspose (prefill bias + NeoX rotary + split), invokeMaskedSoftmax
//   kernels/decoding_kernels.cu                 invokeGatherTree (output gather with the pad gap removed; with parents for beams)
//   kernels/beam_search_penalty_kernels.cu      invokeAddBiasApplyPenalties (temperature / repetition penalty along the beam history)
//   kernels/online_softmax_beamsearch_kernels.cu invokeTopkSoftMax (log-softmax + 2K candidates per row + K winners per batch)
// so that `-m gpu` tests can compare our kernels with the reference's on the same inputs on the B200 (tests/test_ref_kernels_gpu.py).
// Nothing here is part of the product; no reference source is copied.
#include <cuda_fp16.h>
#include <cuda_runtime.h>

#include <cstring>
#include <string>

#include "src/fastertransformer/kernels/add_residual_kernels.h"
#include "src/fastertransformer/kernels/decoder_masked_multihead_attention.h"
#include "src/fastertransformer/kernels/layernorm_kernels.h"
#include "src/fastertransformer/kernels/sampling_penalty_kernels.h"
#include "src/fastertransformer/kernels/sampling_topk_kernels.h"
#include "src/fastertransformer/kernels/sampling_topp_kernels.h"
#include "src/fastertransformer/kernels/stop_criteria_kernels.h"
#include "src/fastertransformer/kernels/decoding_kernels.h"
#include "src/fastertransformer/kernels/beam_search_penalty_kernels.h"
#include "src/fastertransformer/kernels/online_softmax_beamsearch_kernels.h"
#include "src/fastertransformer/kernels/unfused_attention_kernels.h"
#include "src/fastertransformer/utils/Tensor.h"
#include "src/fastertransformer/utils/logger.h"

namespace ft = fastertransformer;

// the two host symbols the kernel objects pull in (utils/logger.cc, utils/Tensor.cc are not built)
namespace fastertransformer {
Logger::Logger() {}
std::string Tensor::getNumpyTypeDesc(DataType) const { return "x"; }
}  // namespace fastertransformer

// explicit instantiations live in decoder_masked_multihead_attention_{128,64}.cu
template <typename T, int Dh, int Dh_MAX, typename KERNEL_PARAMS_TYPE>
void mmha_launch_kernel(const KERNEL_PARAMS_TYPE& params, const cudaStream_t& stream);

static cudaStream_t S(void* s) { return reinterpret_cast<cudaStream_t>(s); }
static int done() { return cudaGetLastError() == cudaSuccess ? 0 : 1; }

extern "C" int ref_layernorm_half(void* out, const void* in, const void* gamma, const void* beta, float eps, int m, int n, void* stream)
{
    ft::invokeGeneralLayerNorm<half>(static_cast<half*>(out), static_cast<const half*>(in), static_cast<const half*>(gamma),
                                     static_cast<const half*>(beta), eps, m, n, (float*)nullptr, 0, S(stream));
    return done();
}

extern "C" int ref_add_bias_attn_ffn_residual_half(void* out, const void* ffn, const void* attn, const void* x, const void* bias, int m,
                                                   int n, int tp, void* stream)
{
    ft::invokeAddBiasAttentionFfnResidual<half>(static_cast<half*>(out), static_cast<const half*>(ffn), static_cast<const half*>(attn),
                                                static_cast<const half*>(x), static_cast<const half*>(bias), m, n, tp, S(stream));
    return done();
}

// k_cache: [B, H, Dh/8, max_len, 8] fp16, v_cache: [B, H, max_len, Dh] (the reference's layouts)
extern "C" int ref_mmha_half(const void* qkv, const void* qkv_bias, void* k_cache, void* v_cache, void* ctx, const void* finished,
                             const int* sequence_lengths, int batch, int heads, int dh, int rotary_dim, int memory_max_len,
                             int max_input_len, const int* total_padding_tokens, int step, const void* masked_tokens, void* stream)
{
    using DataType = uint16_t;
    Masked_multihead_attention_params<DataType> params;
    memset(&params, 0, sizeof(params));
    const int hidden = heads * dh;
    if (qkv_bias != nullptr) {
        params.q_bias = static_cast<const DataType*>(qkv_bias);
        params.k_bias = static_cast<const DataType*>(qkv_bias) + hidden;
        params.v_bias = static_cast<const DataType*>(qkv_bias) + 2 * hidden;
    }
    params.out = static_cast<DataType*>(ctx);
    params.q = static_cast<const DataType*>(qkv);
    params.k = static_cast<const DataType*>(qkv) + hidden;
    params.v = static_cast<const DataType*>(qkv) + 2 * hidden;
    params.stride = 3 * hidden;
    params.finished = const_cast<bool*>(static_cast<const bool*>(finished));
    params.k_cache = static_cast<DataType*>(k_cache);
    params.v_cache = static_cast<DataType*>(v_cache);
    params.cache_indir = nullptr;
    params.batch_size = batch;
    params.beam_width = 1;
    params.memory_max_len = memory_max_len;
    params.length_per_sample = sequence_lengths;
    params.timestep = step - 1;
    params.num_heads = heads;
    params.hidden_size_per_head = dh;
    params.rotary_embedding_dim = rotary_dim;
    params.neox_rotary_style = true;
    params.inv_sqrt_dh = 1.F / sqrtf((float)dh);
    params.total_padding_tokens = total_padding_tokens;
    params.masked_tokens = static_cast<const bool*>(masked_tokens);
    params.max_input_length = max_input_len;
    const cudaStream_t st = S(stream);
    if (dh == 128) mmha_launch_kernel<DataType, 128, 128, Masked_multihead_attention_params<DataType>>(params, st);
    else if (dh == 64) mmha_launch_kernel<DataType, 64, 64, Masked_multihead_attention_params<DataType>>(params, st);
    else return 2;
    return done();
}

extern "C" size_t ref_curand_state_bytes() { return sizeof(curandState_t); }
extern "C" int ref_curand_batch_init(void* states, int batch, const unsigned long long* seeds_dev, void* stream)
{
    ft::invokeCurandBatchInitialize(static_cast<curandState_t*>(states), batch, seeds_dev, S(stream));
    return done();
}
extern "C" int ref_temperature_penalty(float* logits, const float* temperatures, int batch, int vocab, int vocab_padded, void* stream)
{
    ft::invokeBatchApplyTemperaturePenalty<float>(logits, (const float*)nullptr, temperatures, batch, vocab, vocab_padded, S(stream));
    return done();
}
extern "C" int ref_repetition_penalty(float* logits, const float* penalties, const int* output_ids, int batch, int vocab_padded,
                                      const int* input_lengths, int max_input_len, int step, void* stream)
{
    ft::invokeBatchApplyRepetitionPenalty<float>(logits, penalties, output_ids, batch, batch, vocab_padded, input_lengths, max_input_len, step,
                                                 ft::RepetitionPenaltyType::Multiplicative, S(stream));
    return done();
}
extern "C" int ref_add_bias_end_mask(float* logits, const int* end_ids, const void* finished, int batch, int vocab, int vocab_padded, void* stream)
{
    ft::invokeAddBiasEndMask<float>(logits, (const float*)nullptr, end_ids, static_cast<const bool*>(finished), batch, vocab, vocab_padded, S(stream));
    return done();
}
extern "C" int ref_add_bias_softmax(float* logits, const int* end_ids, const void* finished, int batch, int vocab_padded, int vocab, void* stream)
{
    ft::invokeAddBiasSoftMax<float>(logits, (const float*)nullptr, end_ids, static_cast<const bool*>(finished), batch, vocab_padded, vocab, S(stream));
    return done();
}
// workspace == NULL: returns the size needed through *workspace_size
extern "C" int ref_batch_topk_sampling(void* workspace, size_t* workspace_size, const float* log_probs, int* ids, int* sequence_length,
                                       void* finished, float* cum_log_probs, void* curand_states, int max_top_k, const int* top_ks,
                                       const float* top_ps, int vocab_padded, const int* end_ids, int batch, void* stream)
{
    size_t ws = *workspace_size;
    ft::invokeBatchTopKSampling<float>(workspace, ws, log_probs, ids, sequence_length, static_cast<bool*>(finished), cum_log_probs,
                                       (float*)nullptr, static_cast<curandState_t*>(curand_states), max_top_k, top_ks, 1.0f, top_ps,
                                       vocab_padded, end_ids, S(stream), batch, (const bool*)nullptr);
    *workspace_size = ws;
    return done();
}

// Pure top-p rows, as TopPSamplingLayer::runSampling drives them (layers/sampling_layers/TopPSamplingLayer.cu:256-331): `probs`
// holds the softmax of every row (ref_add_bias_softmax).  Scratch is allocated per call (test infrastructure).
extern "C" int ref_batch_topp_sampling(const float* probs, int* ids, int* sequence_length, void* finished, float* cum_log_probs,
                                       void* curand_states, int batch, int vocab_padded, const int* end_ids, float max_top_p,
                                       const float* top_ps, void* stream)
{
    cudaStream_t st = S(stream);
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceProp prop;
    if (cudaGetDeviceProperties(&prop, dev) != cudaSuccess) return 3;
    size_t ws = 0, cub = 0;
    ft::invokeBatchTopPSampling<float>(nullptr, ws, cub, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr,
                                       batch, (size_t)vocab_padded, nullptr, max_top_p, top_ps, st, &prop, nullptr);
    void* wsp = nullptr;
    int *id_vals = nullptr, *offs = nullptr;
    if (cudaMalloc(&wsp, ws + 256) != cudaSuccess || cudaMalloc(&id_vals, (size_t)batch * vocab_padded * sizeof(int)) != cudaSuccess ||
        cudaMalloc(&offs, 2 * (size_t)(batch + 1) * sizeof(int)) != cudaSuccess)
        return 4;
    int* begin_offs = offs + (batch + 1);
    ft::invokeTopPInitialize(id_vals, offs, begin_offs, batch, vocab_padded, st);
    ft::invokeBatchTopPSampling<float>(wsp, ws, cub, ids, sequence_length, static_cast<bool*>(finished), cum_log_probs, nullptr, probs, id_vals,
                                       offs, begin_offs, static_cast<curandState_t*>(curand_states), batch, (size_t)vocab_padded, end_ids,
                                       max_top_p, top_ps, st, &prop, nullptr);
    cudaStreamSynchronize(st);
    cudaFree(wsp);
    cudaFree(id_vals);
    cudaFree(offs);
    return done();
}

// stop_words [B, 2, n]; output_ids time-major [max_len, B]; beam_width 1 (kernels/stop_criteria_kernels.cu:24-81)
extern "C" int ref_stop_words_criterion(const int* output_ids, const int* stop_words, void* finished, int stop_words_len, int batch, int step,
                                        void* stream)
{
    ft::invokeStopWordsCriterion(output_ids, nullptr, stop_words, static_cast<bool*>(finished), 0, (size_t)stop_words_len, batch, 1, step, S(stream));
    return done();
}

// Prefill: qkv [token_num, 3 * H * Dh] (padding removed) + bias -> q / k / v [B, H, S, Dh] with NeoX rotary at the token's index in
// its own sequence, exactly as GptContextAttentionLayer drives it (layers/attention_layers/GptContextAttentionLayer.cc:140-170).
extern "C" int ref_prefill_qkv_bias_rotary_transpose(void* q_buf, void* k_buf, void* v_buf, void* qkv, const void* qkv_bias,
                                                     const int* padding_offset, int batch, int seq_len, int token_num, int heads, int dh,
                                                     int rotary_dim, void* stream)
{
    ft::invokeAddFusedQKVBiasTranspose<half>(static_cast<half*>(q_buf), static_cast<half*>(k_buf), static_cast<half*>(v_buf),
                                             ft::PrefixPromptBatchWeightsParam<half>{}, static_cast<half*>(qkv),
                                             static_cast<const half*>(qkv_bias), padding_offset, batch, seq_len, token_num, heads, dh,
                                             rotary_dim, 1, (const float*)nullptr, 0, S(stream));
    return done();
}

// attention_score (fp16) = softmax(qk (fp32) * scale + (1 - mask) * -10000), kernels/unfused_attention_kernels.cu:255-333
extern "C" int ref_masked_softmax_half(void* attention_score, const float* qk, const void* attention_mask, int batch, int heads, int q_len,
                                       int k_len, float scale, void* stream)
{
    ft::MaskedSoftmaxParam<half, float> param;
    param.attention_score = static_cast<half*>(attention_score);
    param.qk = qk;
    param.attention_mask = static_cast<const half*>(attention_mask);
    param.batch_size = batch;
    param.q_length = q_len;
    param.k_length = k_len;
    param.num_heads = heads;
    param.qk_scale = __float2half(scale);
    ft::invokeMaskedSoftmax(param, S(stream));
    return done();
}

// output_ids [B, 1, max_time] <- time-major step_ids, pad gap [input_len, max_input_length) removed, as GptNeoX<T>::setOutputTensors
// fills gatherTreeParam for sampling (models/gptneox/GptNeoX.cc:1141-1164; kernels/decoding_kernels.cu:452-580).
// `sequence_lengths` is read and updated in place (+1, :1147-1148); `scratch` is the [B, max_time] transposed buffer.
extern "C" int ref_gather_tree_sampling(int* output_ids, int* sequence_lengths, int* scratch, int max_time, int batch, const int* step_ids,
                                        const int* end_tokens, const int* input_lengths, int max_input_length, void* stream)
{
    ft::gatherTreeParam param;
    param.beams = scratch;
    param.max_sequence_lengths = sequence_lengths;
    param.max_sequence_length_final_step = 1;
    param.max_time = max_time;
    param.batch_size = batch;
    param.beam_width = 1;
    param.step_ids = step_ids;
    param.parent_ids = nullptr;
    param.end_tokens = end_tokens;
    param.max_input_length = max_input_length;
    param.prefix_soft_prompt_lengths = nullptr;
    param.input_lengths = input_lengths;
    param.max_prefix_soft_prompt_length = 0;
    param.max_input_without_prompt_length = max_input_length;
    param.stream = S(stream);
    param.output_ids = output_ids;
    ft::invokeGatherTree(param);
    return done();
}

// ---- beam search (beam_width > 1), as layers/beam_search_layers/BaseBeamSearchLayer.cu:228-262 and OnlineBeamSearchLayer.cu:124-142
// call them (no BeamHypotheses: models/gptneox/GptNeoX.cc passes none).
extern "C" int ref_beam_penalties(float* logits, int step, const int* output_ids, const int* parent_ids, const int* input_lengths,
                                  const int* sequence_lengths, int max_input_length, int batch, int beam_width, int vocab, int vocab_padded,
                                  const int* end_ids, float temperature, float repetition_penalty, void* stream)
{
    const ft::RepetitionPenaltyType type = repetition_penalty != 1.0f ? ft::RepetitionPenaltyType::Multiplicative : ft::RepetitionPenaltyType::None;
    ft::invokeAddBiasApplyPenalties<float>(step, logits, output_ids + (size_t)(step - 1) * batch * beam_width, output_ids, parent_ids,
                                           input_lengths, sequence_lengths, (const float*)nullptr, 0, max_input_length, batch, batch, beam_width,
                                           vocab, vocab_padded, end_ids, temperature, repetition_penalty, type, 0, S(stream));
    return done();
}

// ids [batch * beam_width] receives row * vocab_padded + token of the K winners per batch; cum_log_probs is updated in place.
// workspace: floats, sized as OnlineBeamSearchLayer.cu:175-182.
extern "C" size_t ref_beam_topk_workspace_floats(int batch)
{
    return (size_t)(ceil(batch * 64 * (64 * 2) / 4.) * 4 * 2 + ceil(batch * (64 * 2) * 128 * (2 * (4 * 2) + 2) / 4.) * 4);
}
extern "C" int ref_beam_topk_softmax(const float* logits, const void* finished, const int* sequence_lengths, float* cum_log_probs, int* ids,
                                     void* workspace, size_t workspace_floats, int batch, int beam_width, int vocab_padded, const int* end_ids,
                                     float diversity_rate, float length_penalty, void* stream)
{
    ft::BeamHypotheses hyps;
    ft::invokeTopkSoftMax<float>(logits, (const float*)nullptr, static_cast<const bool*>(finished), sequence_lengths, cum_log_probs,
                                 (float*)nullptr, ids, workspace, (int)workspace_floats, &hyps, batch, beam_width, vocab_padded, end_ids,
                                 diversity_rate, length_penalty, S(stream));
    return done();
}

// output_ids [B, beam, max_time] <- time-major step_ids / parent_ids, as GptNeoX<T>::setOutputTensors (models/gptneox/GptNeoX.cc:1141-1164)
extern "C" int ref_gather_tree_beams(int* output_ids, int* sequence_lengths, int* scratch, int max_time, int batch, int beam_width,
                                     const int* step_ids, const int* parent_ids, const int* end_tokens, const int* input_lengths,
                                     int max_input_length, void* stream)
{
    ft::gatherTreeParam param;
    param.beams = scratch;
    param.max_sequence_lengths = sequence_lengths;
    param.max_sequence_length_final_step = 1;
    param.max_time = max_time;
    param.batch_size = batch;
    param.beam_width = beam_width;
    param.step_ids = step_ids;
    param.parent_ids = parent_ids;
    param.end_tokens = end_tokens;
    param.max_input_length = max_input_length;
    param.prefix_soft_prompt_lengths = nullptr;
    param.input_lengths = input_lengths;
    param.max_prefix_soft_prompt_length = 0;
    param.max_input_without_prompt_length = max_input_length;
    param.stream = S(stream);
    param.output_ids = output_ids;
    ft::invokeGatherTree(param);
    return done();
}
