"""ctypes view of include/ftcf.h (libftcf.so).

This is the only way Python code in this package reaches the engine: raw device pointers and sizes through the
C ABI, exactly what the pybind11 shims (csrc/binding/) pass.  There is no CPU or PyTorch fallback: if the shared
library is missing or the device is not sm_100 the calls raise.
"""
from __future__ import annotations

import ctypes as C
import os
from typing import Optional

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_DIR = os.path.join(_HERE, "lib")
LIB_PATH = os.path.join(LIB_DIR, "libftcf.so")


class FtcfError(RuntimeError):
    pass


class MmhaParams(C.Structure):
    _fields_ = [
        ("qkv", C.c_void_p), ("qkv_bias", C.c_void_p), ("k_cache", C.c_void_p), ("v_cache", C.c_void_p),
        ("ctx", C.c_void_p), ("seq_len", C.c_void_p), ("input_len", C.c_void_p), ("pad_count", C.c_void_p),
        ("finished", C.c_void_p), ("step", C.c_void_p), ("partial", C.c_void_p), ("counters", C.c_void_p),
        ("batch", C.c_int32), ("heads", C.c_int32), ("dh", C.c_int32), ("rotary_dim", C.c_int32),
        ("max_len", C.c_int32), ("max_input_len", C.c_int32), ("splits", C.c_int32), ("inv_sqrt_dh", C.c_float),
        ("cache_indir", C.c_void_p), ("beam_width", C.c_int32),
    ]


class BeamParams(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("output_ids", C.c_void_p), ("parent_ids", C.c_void_p), ("seq_len", C.c_void_p),
        ("finished", C.c_void_p), ("cum_log_probs", C.c_void_p), ("input_len", C.c_void_p), ("cache_indir", C.c_void_p),
        ("stop_words", C.c_void_p), ("step", C.c_void_p), ("finished_count_host_mapped", C.c_void_p),
        ("finished_hist_host_mapped", C.c_void_p), ("workspace", C.c_void_p),
        ("batch", C.c_int32), ("beam_width", C.c_int32), ("vocab", C.c_int32), ("vocab_padded", C.c_int32), ("n_stop", C.c_int32),
        ("max_input_len", C.c_int32), ("max_len", C.c_int32), ("end_id", C.c_int32),
        ("temperature", C.c_float), ("repetition_penalty", C.c_float), ("diversity_rate", C.c_float), ("length_penalty", C.c_float),
        ("args_differ", C.c_int32),
    ]


class SamplingParams(C.Structure):
    _fields_ = [
        ("logits", C.c_void_p), ("output_ids", C.c_void_p), ("seq_len", C.c_void_p), ("finished", C.c_void_p),
        ("cum_log_probs", C.c_void_p), ("input_len", C.c_void_p), ("top_k", C.c_void_p), ("top_p", C.c_void_p),
        ("temperature", C.c_void_p), ("repetition_penalty", C.c_void_p), ("optional_last_tokens", C.c_void_p),
        ("stop_words", C.c_void_p), ("curand_states", C.c_void_p), ("step", C.c_void_p),
        ("finished_count_host_mapped", C.c_void_p), ("workspace", C.c_void_p),
        ("batch", C.c_int32), ("vocab", C.c_int32), ("vocab_padded", C.c_int32), ("max_top_k", C.c_int32),
        ("n_last", C.c_int32), ("n_stop", C.c_int32), ("max_input_len", C.c_int32), ("max_len", C.c_int32),
        ("end_id", C.c_int32), ("want_probs", C.c_int32), ("has_top_p_rows", C.c_int32),
        ("finished_hist_host_mapped", C.c_void_p),
    ]


class LaunchHint(C.Structure):
    _fields_ = [("target_ctas", C.c_int32), ("no_pdl", C.c_int32), ("stages", C.c_int32)]


class TpExchange(C.Structure):
    _fields_ = [("peer_data", C.c_void_p * 8), ("tp", C.c_int32), ("rank", C.c_int32),
                ("m_max", C.c_int32), ("h", C.c_int32), ("step", C.c_void_p), ("step_base", C.c_int32), ("layer_num", C.c_int32)]


class LnPrologue(C.Structure):
    _fields_ = [("x", C.c_void_p), ("add_ffn", C.c_void_p), ("add_attn", C.c_void_p), ("add_bias", C.c_void_p),
                ("gamma", C.c_void_p), ("beta", C.c_void_p), ("x_out", C.c_void_p), ("eps", C.c_float), ("cta_hint", C.c_int32)]


class GptNeoXConfig(C.Structure):
    _fields_ = [
        ("head_num", C.c_int32), ("size_per_head", C.c_int32), ("inter_size", C.c_int32), ("layer_num", C.c_int32),
        ("vocab_size", C.c_int32), ("rotary_embedding_dim", C.c_int32), ("start_id", C.c_int32), ("end_id", C.c_int32),
        ("tensor_para_size", C.c_int32), ("tensor_para_rank", C.c_int32), ("int8_mode", C.c_int32),
        ("use_gptj_residual", C.c_int32), ("layernorm_eps", C.c_float), ("int8_layout", C.c_int32),
    ]


TOKEN_CALLBACK = C.CFUNCTYPE(None, C.c_void_p, C.c_int32, C.POINTER(C.c_int32), C.POINTER(C.c_int32), C.c_int32)


class GptNeoXRequest(C.Structure):
    _fields_ = [
        ("input_ids", C.c_void_p), ("input_lengths", C.c_void_p),
        ("batch", C.c_int32), ("max_input_len", C.c_int32), ("output_len", C.c_int32), ("beam_width", C.c_int32),
        ("top_k_host", C.c_void_p), ("n_top_k", C.c_int32),
        ("top_p_host", C.c_void_p), ("n_top_p", C.c_int32),
        ("temperature_host", C.c_void_p), ("n_temperature", C.c_int32),
        ("repetition_penalty_host", C.c_void_p), ("n_repetition_penalty", C.c_int32),
        ("random_seed_host", C.c_void_p), ("n_random_seed", C.c_int32),
        ("beam_search_diversity_rate_host", C.c_void_p), ("n_beam_search_diversity_rate", C.c_int32),
        ("len_penalty_host", C.c_void_p), ("n_len_penalty", C.c_int32),
        ("stop_words", C.c_void_p), ("n_stop", C.c_int32),
        ("optional_last_tokens", C.c_void_p), ("n_last", C.c_int32),
        ("return_cum_log_probs", C.c_int32),
        ("callback", TOKEN_CALLBACK), ("callback_user", C.c_void_p),
        ("output_ids", C.c_void_p), ("sequence_lengths", C.c_void_p), ("cum_log_probs", C.c_void_p),
        ("logits_trace", C.c_void_p), ("logits_trace_steps", C.c_int32),
    ]


class GptNeoXStats(C.Structure):
    _fields_ = [("steps", C.c_int32), ("prefill_ms", C.c_float), ("decode_ms", C.c_float), ("kernel_launches", C.c_int64)]


# name -> (restype, argtypes); every symbol include/ftcf.h declares
SIGNATURES = {
    "ftcf_last_error": (C.c_char_p, []),
    "ftcf_abi_version": (C.c_int, []),
    "ftcf_device_check": (C.c_int, []),
    "ftcf_launch_count": (C.c_longlong, []),
    "ftcf_set_tunable": (C.c_int, [C.c_char_p, C.c_int]),
    "ftcf_debug_trace_start": (C.c_int, [C.c_uint]),
    "ftcf_debug_trace_stop": (C.c_int, [C.c_void_p, C.c_uint, C.POINTER(C.c_uint)]),
    "ftcf_debug_decode_probe": (C.c_int, [C.c_void_p]),
    "ftcf_symmetric_quantize_int8_host": (C.c_int, [C.c_void_p, C.c_int, C.c_size_t, C.c_size_t, C.c_size_t, C.c_void_p,
                                                    C.c_void_p, C.c_void_p]),
    "ftcf_int8_plain_to_b200_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "ftcf_int8_ampere_to_b200_host": (C.c_int, [C.c_void_p, C.c_size_t, C.c_size_t, C.c_void_p]),
    "ftcf_gemm_w8a16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                  C.c_int, C.c_int, C.c_void_p]),
    "ftcf_gemm_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                C.c_int, C.c_int, C.c_void_p]),
    "ftcf_gemm_w8a16_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_int, C.POINTER(LaunchHint), C.c_void_p]),
    "ftcf_gemm_f16_ex": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.POINTER(LaunchHint), C.c_void_p]),
    "ftcf_gemm_w8a16_ln": (C.c_int, [C.POINTER(LnPrologue), C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int,
                                     C.c_int, C.c_void_p]),
    "ftcf_gemm_f16_ln": (C.c_int, [C.POINTER(LnPrologue), C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int,
                                   C.c_int, C.c_int, C.c_void_p]),
    "ftcf_gemm_w8a16_tp_push": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.POINTER(TpExchange), C.c_int, C.c_int, C.c_int, C.c_int,
                                          C.c_int, C.POINTER(LaunchHint), C.c_void_p]),
    "ftcf_tp_gather_residual": (C.c_int, [C.POINTER(TpExchange), C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ftcf_transpose_f16": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_void_p]),
    "ftcf_layernorm": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "ftcf_add_bias_residual_layernorm": (C.c_int, [C.c_void_p] * 7 + [C.c_int, C.c_int, C.c_float, C.c_void_p]),
    "ftcf_add_bias_attn_ffn_residual": (C.c_int, [C.c_void_p] * 5 + [C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ftcf_add_bias_residual": (C.c_int, [C.c_void_p] * 4 + [C.c_int, C.c_int, C.c_void_p]),
    "ftcf_embedding_lookup": (C.c_int, [C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_void_p]),
    "ftcf_mmha_decode": (C.c_int, [C.POINTER(MmhaParams), C.c_void_p]),
    "ftcf_mmha_prefetch_cache": (C.c_int, [C.POINTER(MmhaParams), C.c_void_p]),
    "ftcf_mmha_choose_splits": (C.c_int, [C.c_int, C.c_int, C.c_int]),
    "ftcf_prefill_qkv_rotary_scatter": (C.c_int, [C.c_void_p] * 7 + [C.c_int] * 5 + [C.c_void_p]),
    "ftcf_prefill_attention": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_float, C.c_void_p]),
    "ftcf_sampling_workspace_bytes": (C.c_size_t, [C.c_int, C.c_int, C.c_int]),
    "ftcf_curand_state_bytes": (C.c_size_t, []),
    "ftcf_curand_init": (C.c_int, [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]),
    "ftcf_sampling_step": (C.c_int, [C.POINTER(SamplingParams), C.c_void_p]),
    "ftcf_gather_output": (C.c_int, [C.c_void_p] * 5 + [C.c_int] * 4 + [C.c_void_p]),
    "ftcf_beam_workspace_bytes": (C.c_size_t, [C.c_int] * 4),
    "ftcf_beam_search_step": (C.c_int, [C.POINTER(BeamParams), C.c_void_p]),
    "ftcf_gather_output_beams": (C.c_int, [C.c_void_p] * 6 + [C.c_int] * 5 + [C.c_void_p]),
    "ftcf_gptneox_create": (C.c_int, [C.POINTER(C.c_void_p), C.POINTER(GptNeoXConfig), C.POINTER(C.c_void_p), C.c_size_t,
                                      C.POINTER(C.c_void_p), C.POINTER(C.c_void_p), C.c_size_t, C.c_void_p, C.c_void_p]),
    "ftcf_gptneox_destroy": (None, [C.c_void_p]),
    "ftcf_nccl_unique_id": (C.c_int, [C.c_void_p]),
    "ftcf_gptneox_forward": (C.c_int, [C.c_void_p, C.POINTER(GptNeoXRequest), C.POINTER(GptNeoXStats)]),
    "ftcf_gptneox_last_step_ms": (C.c_int, [C.c_void_p, C.POINTER(C.c_float), C.c_int]),
    "ftcf_gptneox_set_option": (C.c_int, [C.c_void_p, C.c_char_p, C.c_int]),
}

_lib: Optional[C.CDLL] = None


def load() -> C.CDLL:
    """Load libftcf.so (built by `__graft_entry__.build()` / `make -C csrc`).  Raises if it is missing."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise FtcfError(f"{LIB_PATH} is missing: build it with `python -c 'import __graft_entry__ as g; g.build()'`; "
                        "there is no fallback path")
    lib = C.CDLL(LIB_PATH, mode=C.RTLD_GLOBAL)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)          # AttributeError here == the library does not export a declared symbol
        fn.restype = res
        fn.argtypes = args
    # experiment hook: FTCF_TUNABLES="skinny_target_ctas=148,pdl=0"
    for item in filter(None, os.environ.get("FTCF_TUNABLES", "").split(",")):
        key, _, val = item.partition("=")
        if lib.ftcf_set_tunable(key.strip().encode(), int(val)) != 0:
            raise FtcfError(f"FTCF_TUNABLES: {lib.ftcf_last_error().decode()}")
    _lib = lib
    return lib


def check(status: int) -> None:
    if status != 0:
        msg = load().ftcf_last_error().decode("utf-8", "replace")
        raise FtcfError(f"libftcf error {status}: {msg}")
