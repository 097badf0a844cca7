// pybind11 module `libth_gptneox`: the reference's Python class GptNeoXOp over the C ABI of libftcf.so.
//
// Same constructor and forward() arguments, return values and argument-error behaviour as
// src/fastertransformer/th_op/gptneox/GptNeoXOp.cc:190-212 (constructor :25-106, forward :113-185), so that
// examples/pytorch/codefuse/codefuse_example.py:469-470,533-536,575-589 runs unchanged with --lib_path pointing at
// fastertransformer4codefuse_b200/lib.  Everything below this file is plain pointers and sizes (include/ftcf.h).
#include <ATen/cuda/CUDAContext.h>
#include <torch/extension.h>

#include <cstdlib>
#include <cstring>
#include <memory>
#include <string>
#include <vector>

#include "ftcf.h"

namespace py = pybind11;

namespace {

void check_status(int status)
{
    // argument errors and run-time errors both surface as a Python RuntimeError (the reference exits the process on
    // run-time CUDA errors, GptNeoXOp.h:370-380; raising is the allowed replacement, never partial tensors)
    TORCH_CHECK(status == FTCF_OK, "libftcf error ", status, ": ", ftcf_last_error());
}

const void* data_or_null(const at::Tensor& t) { return t.defined() && t.numel() > 0 ? t.data_ptr() : nullptr; }

struct CallbackCtx {
    py::object fn;
    int end_id;
    int beam;
    std::string error;
    bool failed = false;
};

// {"last_tokens": [[id] * beam] * B, "idxs": [[n] * beam] * B}, th_op/gptneox/utils/pybind_callback_utils.cc:59-103
void token_callback(void* user, int32_t /*step*/, const int32_t* toks, const int32_t* idxs, int32_t rows)
{
    auto* ctx = static_cast<CallbackCtx*>(user);
    if (ctx->failed) return;
    try {
        py::list last_tokens, last_idxs;
        for (int b = 0; b < rows / ctx->beam; ++b) {
            py::list t, i;
            for (int j = 0; j < ctx->beam; ++j) {
                t.append(py::int_(toks[b * ctx->beam + j]));
                i.append(py::int_(idxs[b * ctx->beam + j]));
            }
            last_tokens.append(t);
            last_idxs.append(i);
        }
        py::dict msg;
        msg["last_tokens"] = last_tokens;
        msg["idxs"] = last_idxs;
        ctx->fn(msg);
    } catch (py::error_already_set& e) {
        ctx->failed = true;
        ctx->error = e.what();
    }
}

class GptNeoXOp {
public:
    GptNeoXOp(py::object comm, int64_t rank, int64_t head_num, int64_t size_per_head, int64_t inter_size, int64_t layer_num,
              int64_t vocab_size, int64_t rotary_embedding_dim, int64_t start_id, int64_t end_id, int64_t tensor_para_size,
              int64_t pipeline_para_size, int64_t int8_mode, int64_t /*max_seq_len*/, bool use_gptj_residual,
              std::vector<at::Tensor> weights, std::vector<at::Tensor> int8_weights, std::vector<at::Tensor> scale)
        : weights_(std::move(weights)), int8_weights_(std::move(int8_weights)), scale_(std::move(scale)), end_id_((int)end_id)
    {
        TORCH_CHECK(pipeline_para_size == 1, "pipeline_para_size must be 1 (codefuse_example.py:647 fixes it)");
        TORCH_CHECK((int64_t)weights_.size() == 12 * layer_num + 4, "expected ", 12 * layer_num + 4, " weight tensors, got ", weights_.size());
        const auto st = weights_[0].scalar_type();
        // the reference dispatches on weights[0]: Float -> FTGptNeoX<float>, Half -> FTGptNeoX<half> (GptNeoXOp.cc:46,56-105)
        TORCH_CHECK(st == at::kHalf || st == at::kFloat, "Wrong tensor type: weights must be fp16 or fp32");
        for (const auto& t : weights_) {                    // CHECK_INPUT, GptNeoXOp.cc:52-54
            if (t.numel() == 0) continue;
            TORCH_CHECK(t.is_cuda(), "weights must be CUDA tensors");
            TORCH_CHECK(t.is_contiguous(), "weights must be contiguous");
            TORCH_CHECK(t.scalar_type() == st, "weights must share one dtype");
        }
        if (st == at::kFloat) {
            // fp32 checkpoints are accepted and rounded to fp16 once, here: the B200 engine computes in fp16 with fp32 accumulation
            // only (a deliberate difference from FTGptNeoX<float>); the fp16 copies are what we keep alive
            for (auto& t : weights_)
                if (t.numel() > 0) t = t.to(at::kHalf).contiguous();
            for (auto& t : scale_)
                if (t.defined() && t.numel() > 0 && t.scalar_type() == at::kFloat) t = t.to(at::kHalf).contiguous();
        }
        ftcf_gptneox_config cfg{};
        cfg.head_num = (int)head_num; cfg.size_per_head = (int)size_per_head; cfg.inter_size = (int)inter_size;
        cfg.layer_num = (int)layer_num; cfg.vocab_size = (int)vocab_size; cfg.rotary_embedding_dim = (int)rotary_embedding_dim;
        cfg.start_id = (int)start_id; cfg.end_id = (int)end_id; cfg.tensor_para_size = (int)tensor_para_size;
        cfg.tensor_para_rank = (int)(rank % tensor_para_size); cfg.int8_mode = (int)int8_mode;
        cfg.use_gptj_residual = use_gptj_residual ? 1 : 0; cfg.layernorm_eps = 1e-5f;
        // bytes of the int8 tensors: 0 = made by OUR libth_common (B200 layout), 1 = plain int8 [k,n], 2 = reference-made
        // sm80 `*.q.bin` files
        const char* lay = std::getenv("FTCF_INT8_LAYOUT");
        cfg.int8_layout = lay ? std::atoi(lay) : 0;

        std::vector<const void*> w(weights_.size());
        for (size_t i = 0; i < weights_.size(); ++i) w[i] = data_or_null(weights_[i]);
        const size_t n8 = int8_mode == 1 ? int8_weights_.size() : 0;
        TORCH_CHECK(int8_mode != 1 || (n8 == (size_t)(4 * layer_num) && scale_.size() == n8), "int8_mode=1 needs ", 4 * layer_num,
                    " int8 weights and scales");
        std::vector<const void*> q(n8 ? n8 : 1, nullptr), s(n8 ? n8 : 1, nullptr);
        for (size_t i = 0; i < n8; ++i) {
            TORCH_CHECK(int8_weights_[i].is_cuda() && scale_[i].is_cuda(), "int8 weights / scales must be CUDA tensors");
            q[i] = data_or_null(int8_weights_[i]);
            s[i] = data_or_null(scale_[i]);
        }
        char uid[128];
        const void* uid_ptr = nullptr;
        if (tensor_para_size > 1) {
            // rank 0 of the group makes the ncclUniqueId, everyone receives it through the torch process group
            // (replaces nccl_inherit::ftNcclInitialize, th_op/gptneox/utils/nccl_inherit_utils.cc:8-68)
            py::module_ dist = py::module_::import("torch.distributed");
            const int group_rank = dist.attr("get_rank")(comm).cast<int>();
            py::list box;
            if (group_rank == 0) {
                check_status(ftcf_nccl_unique_id(uid));
                box.append(py::bytes(uid, 128));
            } else {
                box.append(py::none());
            }
            const int src = dist.attr("get_global_rank")(comm, 0).cast<int>();
            dist.attr("broadcast_object_list")(box, py::arg("src") = src, py::arg("group") = comm);
            const std::string got = box[0].cast<std::string>();
            TORCH_CHECK(got.size() == 128, "NCCL id exchange failed");
            std::memcpy(uid, got.data(), 128);
            uid_ptr = uid;
        }
        void* stream = at::cuda::getCurrentCUDAStream().stream();   // captured at construction, GptNeoXOp.h:180
        check_status(ftcf_gptneox_create(&handle_, &cfg, w.data(), w.size(), n8 ? q.data() : nullptr, n8 ? s.data() : nullptr, n8, uid_ptr,
                                         stream));
        if (const char* opts = std::getenv("FTCF_OPTIONS")) {        // experiment hook: "cuda_graph=0,two_branch=0"
            std::string o(opts);
            size_t p = 0;
            while (p < o.size()) {
                size_t c = o.find(',', p);
                if (c == std::string::npos) c = o.size();
                const std::string item = o.substr(p, c - p);
                const size_t eq = item.find('=');
                if (eq != std::string::npos) check_status(ftcf_gptneox_set_option(handle_, item.substr(0, eq).c_str(), std::atoi(item.c_str() + eq + 1)));
                p = c + 1;
            }
        }
    }
    ~GptNeoXOp() { ftcf_gptneox_destroy(handle_); }
    GptNeoXOp(const GptNeoXOp&) = delete;
    GptNeoXOp& operator=(const GptNeoXOp&) = delete;

    std::vector<at::Tensor> forward(at::Tensor input_ids, at::Tensor input_lengths, int64_t output_len, c10::optional<int64_t> beam_width_opt,
                                    c10::optional<at::Tensor> top_k_opt, c10::optional<at::Tensor> top_p_opt,
                                    c10::optional<at::Tensor> beam_search_diversity_rate_opt, c10::optional<at::Tensor> temperature_opt,
                                    c10::optional<at::Tensor> len_penalty_opt, c10::optional<at::Tensor> repetition_penalty_opt,
                                    c10::optional<at::Tensor> random_seed_opt, c10::optional<at::Tensor> stop_words_list_opt,
                                    c10::optional<at::Tensor> optional_last_tokens_opt, c10::optional<int64_t> return_cum_log_probs_opt,
                                    py::object callback_opt)
    {
        // GptNeoXOp.cc:133-145
        TORCH_CHECK(input_ids.is_cuda(), "input_ids must be a CUDA tensor");
        TORCH_CHECK(input_ids.is_contiguous(), "input_ids must be contiguous");
        TORCH_CHECK(input_ids.scalar_type() == at::kInt, "input_ids dtype should be int32");
        TORCH_CHECK(input_lengths.is_cuda(), "input_lengths must be a CUDA tensor");
        TORCH_CHECK(input_lengths.is_contiguous(), "input_lengths must be contiguous");
        TORCH_CHECK(input_lengths.scalar_type() == at::kInt, "input_lengths dtype should be int32");
        TORCH_CHECK(input_ids.dim() == 2, "input_ids must be [batch, max_input_length]");
        const int64_t rcl = return_cum_log_probs_opt.has_value() ? *return_cum_log_probs_opt : 0;
        TORCH_CHECK(rcl == 0 || rcl == 1, "return_cum_log_probs should be 0 (no return cum_log_probs),  1 (the cumulative log probs of generated sequences)");
        const int beam = beam_width_opt.has_value() && *beam_width_opt > 1 ? (int)*beam_width_opt : 1;
        const int64_t B = input_ids.size(0), S = input_ids.size(1);
        auto opt_i32 = at::TensorOptions().dtype(at::kInt).device(input_ids.device());
        at::Tensor output_ids = at::empty({B, beam, S + output_len}, opt_i32);
        at::Tensor sequence_lengths = at::empty({B, beam}, opt_i32);
        at::Tensor cum_log_probs = at::empty({B, beam}, opt_i32.dtype(at::kFloat));

        std::vector<at::Tensor> keep;
        auto host = [&](const c10::optional<at::Tensor>& t, at::ScalarType dt, const void*& ptr, int32_t& n) {
            ptr = nullptr;
            n = 0;
            if (!t.has_value() || !t->defined()) return;
            at::Tensor h = t->detach().to(at::kCPU, dt).contiguous().reshape({-1});
            keep.push_back(h);
            ptr = h.data_ptr();
            n = (int32_t)h.numel();
        };
        ftcf_gptneox_request rq{};
        rq.input_ids = input_ids.data_ptr<int32_t>();
        rq.input_lengths = input_lengths.data_ptr<int32_t>();
        rq.batch = (int32_t)B; rq.max_input_len = (int32_t)S; rq.output_len = (int32_t)output_len;
        rq.beam_width = beam;
        const void* p = nullptr;
        host(top_k_opt, at::kInt, p, rq.n_top_k); rq.top_k_host = static_cast<const int32_t*>(p);
        host(top_p_opt, at::kFloat, p, rq.n_top_p); rq.top_p_host = static_cast<const float*>(p);
        host(temperature_opt, at::kFloat, p, rq.n_temperature); rq.temperature_host = static_cast<const float*>(p);
        host(repetition_penalty_opt, at::kFloat, p, rq.n_repetition_penalty); rq.repetition_penalty_host = static_cast<const float*>(p);
        host(random_seed_opt, at::kLong, p, rq.n_random_seed); rq.random_seed_host = static_cast<const int64_t*>(p);
        host(beam_search_diversity_rate_opt, at::kFloat, p, rq.n_beam_search_diversity_rate);
        rq.beam_search_diversity_rate_host = static_cast<const float*>(p);
        host(len_penalty_opt, at::kFloat, p, rq.n_len_penalty); rq.len_penalty_host = static_cast<const float*>(p);
        if (stop_words_list_opt.has_value() && stop_words_list_opt->defined()) {
            at::Tensor sw = stop_words_list_opt->contiguous();
            TORCH_CHECK(sw.is_cuda() && sw.scalar_type() == at::kInt && sw.dim() == 3, "stop_words_list must be a CUDA int32 tensor [batch, 2, n]");
            keep.push_back(sw);
            rq.stop_words = sw.data_ptr<int32_t>();
            rq.n_stop = (int32_t)sw.size(2);
        }
        if (optional_last_tokens_opt.has_value() && optional_last_tokens_opt->defined()) {
            at::Tensor ol = optional_last_tokens_opt->contiguous();
            TORCH_CHECK(ol.is_cuda() && ol.scalar_type() == at::kInt && ol.dim() == 2, "optional_last_tokens must be a CUDA int32 tensor [batch, n]");
            keep.push_back(ol);
            rq.optional_last_tokens = ol.data_ptr<int32_t>();
            rq.n_last = (int32_t)ol.size(1);
        }
        rq.return_cum_log_probs = (int32_t)rcl;
        CallbackCtx cb{callback_opt, end_id_, beam, {}, false};
        if (!callback_opt.is_none()) {
            rq.callback = token_callback;
            rq.callback_user = &cb;
        }
        rq.output_ids = output_ids.data_ptr<int32_t>();
        rq.sequence_lengths = sequence_lengths.data_ptr<int32_t>();
        rq.cum_log_probs = rcl > 0 ? cum_log_probs.data_ptr<float>() : nullptr;
        check_status(ftcf_gptneox_forward(handle_, &rq, &stats_));   // the GIL stays held, as in the reference
        TORCH_CHECK(!cb.failed, "streaming callback raised: ", cb.error);
        if (rcl > 0) return {output_ids, sequence_lengths, cum_log_probs};
        return {output_ids, sequence_lengths};
    }

    void set_option(const std::string& name, int64_t value) { check_status(ftcf_gptneox_set_option(handle_, name.c_str(), (int)value)); }
    std::vector<float> last_step_ms() const      // per-step decode times of the last request (option "step_timing" = 1)
    {
        std::vector<float> v(8192);
        v.resize((size_t)ftcf_gptneox_last_step_ms(handle_, v.data(), (int)v.size()));
        return v;
    }
    py::dict last_stats() const
    {
        py::dict d;
        d["steps"] = stats_.steps;
        d["prefill_ms"] = stats_.prefill_ms;
        d["decode_ms"] = stats_.decode_ms;
        d["kernel_launches"] = stats_.kernel_launches;
        return d;
    }

private:
    std::vector<at::Tensor> weights_, int8_weights_, scale_;   // kept alive: the engine stores raw pointers (GptNeoXOp.h:402-404)
    ftcf_gptneox* handle_ = nullptr;
    ftcf_gptneox_stats stats_{};
    int end_id_;
};

}  // namespace

PYBIND11_MODULE(libth_gptneox, module)
{
    py::class_<GptNeoXOp>(module, "GptNeoXOp")
        .def(py::init<py::object, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t, int64_t,
                      int64_t, bool, std::vector<at::Tensor>, std::vector<at::Tensor>, std::vector<at::Tensor>>())
        .def("forward", &GptNeoXOp::forward)
        .def("set_option", &GptNeoXOp::set_option)
        .def("last_stats", &GptNeoXOp::last_stats)
        .def("last_step_ms", &GptNeoXOp::last_step_ms);
}
