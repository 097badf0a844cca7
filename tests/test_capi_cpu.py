"""CPU tests of the drop-in boundary: libftcf.so loads, exports every symbol include/ftcf.h declares, the ctypes view
covers them, the CPU-side quantiser entry points agree with the oracle / the reference's goldens, and argument errors
surface as errors (no compute kernels are launched here -- there is no GPU)."""
import glob
import os
import re

import numpy as np
import pytest
import torch

from fastertransformer4codefuse_b200 import capi, quant
from oracle import quant_ref as Q

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
GOLD = os.path.join(ROOT, "tests", "golden")


def _declared_symbols():
    text = open(os.path.join(ROOT, "include", "ftcf.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(ftcf_[a-z0-9_]+)\s*\(", text)))


def test_header_symbols_are_exported_and_bound():
    lib = capi.load()
    names = _declared_symbols()
    assert len(names) >= 25
    for n in names:
        assert hasattr(lib, n), f"libftcf.so does not export {n} (declared in include/ftcf.h)"
        assert n in capi.SIGNATURES, f"capi.SIGNATURES has no prototype for {n}"
    assert lib.ftcf_abi_version() >= 1


def test_no_silent_cpu_fallback():
    # without an sm_100 device the device check must FAIL (loudly), never pretend
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    lib = capi.load()
    assert lib.ftcf_device_check() != 0
    assert lib.ftcf_last_error()
    with pytest.raises(capi.FtcfError):
        capi.check(lib.ftcf_device_check())


def test_bad_arguments_are_reported():
    lib = capi.load()
    assert lib.ftcf_set_tunable(b"no_such_knob", 1) == 1
    assert b"unknown tunable" in lib.ftcf_last_error()
    assert lib.ftcf_gemm_w8a16(None, None, None, None, None, 1, 1, 1, 0, 0, None) == 1       # FTCF_ERR_INVALID
    assert lib.ftcf_symmetric_quantize_int8_host(None, 1, 1, 1, 1, None, None, None) == 1
    assert lib.ftcf_mmha_decode(None, None) == 1


def test_host_side_sizing_and_kernel_choice():
    """Pure host logic of the C ABI (no launch): decode-attention split choice and the kernel family it implies, beam-search
    workspace sizing, beam-search argument checks."""
    lib = capi.load()
    # batch 1, 13B heads: the bulk-staged kernel, one CTA per 64 keys; it publishes nothing at larger batches
    assert lib.ftcf_mmha_choose_splits(1, 40, 1536) == 24
    assert lib.ftcf_mmha_choose_splits(4, 40, 1536) == 24
    # batch 32 (and tensor-parallel ranks at batch 32, 5-20 heads each): the streaming kernels, at least 128 keys per split
    for heads in (40, 20, 10, 5):
        s = lib.ftcf_mmha_choose_splits(32, heads, 2560)
        assert 1 <= s <= 2560 // 128, (heads, s)
    assert lib.ftcf_mmha_choose_splits(32, 40, 2560) == 1
    # a split never gets more keys than the score buffer holds (4096)
    assert lib.ftcf_mmha_choose_splits(64, 40, 32768) >= 8
    # beam-search workspace: grows with rows, beams and history length; zero for nonsense
    a = lib.ftcf_beam_workspace_bytes(1, 3, 100864, 1536)
    assert a > 0 and lib.ftcf_beam_workspace_bytes(2, 3, 100864, 1536) > a and lib.ftcf_beam_workspace_bytes(1, 8, 100864, 1536) > a
    assert lib.ftcf_beam_workspace_bytes(1, 3, 100864, 4096) > a
    assert lib.ftcf_beam_workspace_bytes(0, 3, 100864, 1536) == 0
    assert lib.ftcf_beam_search_step(None, None) == 1                                     # FTCF_ERR_INVALID
    bp = capi.BeamParams()
    bp.batch, bp.beam_width, bp.vocab, bp.vocab_padded, bp.max_len = 1, 33, 100, 100, 10
    assert lib.ftcf_beam_search_step(bp, None) != 0 and b"beam_width" in lib.ftcf_last_error()
    bp.beam_width = 1
    assert lib.ftcf_beam_search_step(bp, None) == 1                                       # beam search needs beam_width > 1
    assert lib.ftcf_gather_output_beams(None, None, None, None, None, None, 1, 3, 4, 8, 0, None) == 1


@pytest.mark.parametrize("dtype", [torch.float16, torch.float32, torch.bfloat16])
@pytest.mark.parametrize("shape", [(64, 48), (2, 32, 16), (128, 200)])
def test_libth_common_quantiser_vs_oracle(dtype, shape):
    torch.manual_seed(sum(shape))
    w = (torch.randn(*shape) * 0.03).to(dtype)
    w.reshape(-1)[3] = 0
    proc, scales, unproc = quant.symmetric_quantize_last_axis_of_batched_matrix_int8(w, return_unprocessed=True)
    assert proc.shape == w.shape and proc.dtype == torch.int8
    assert scales.dtype == dtype and tuple(scales.shape) == ((shape[-1],) if len(shape) == 2 else (shape[0], shape[-1]))
    w3 = w.reshape((-1,) + tuple(shape[-2:]))
    for e in range(w3.shape[0]):
        q_ref, s_ref = Q.symmetric_quantize_unprocessed(w3[e].float().numpy())
        assert np.array_equal(unproc.reshape(w3.shape)[e].numpy(), q_ref)
        assert torch.equal(scales.reshape(w3.shape[0], -1)[e], torch.from_numpy(s_ref).to(dtype))
        k, n = shape[-2:]
        got = proc.reshape(w3.shape[0], -1)[e].numpy().view(np.uint8)
        assert np.array_equal(got, Q.to_b200_layout(q_ref).reshape(-1).view(np.uint8))


@pytest.mark.parametrize("fixture", sorted(glob.glob(os.path.join(GOLD, "quant_ref_*.npz"))))
def test_libth_common_quantiser_vs_reference_goldens(fixture):
    """Same matrices the reference's object code quantised: identical int8 values and fp16 scales; and the loader for
    reference-made sm80 `*.q.bin` bytes lands on exactly our layout."""
    z = np.load(fixture)
    w = torch.from_numpy(z["w"])
    proc, scales, unproc = quant.symmetric_quantize_last_axis_of_batched_matrix_int8(w, return_unprocessed=True)
    assert np.array_equal(unproc.numpy(), z["unprocessed"])
    assert np.array_equal(scales.to(torch.float16).numpy().view(np.uint16), z["scales"].view(np.uint16))
    k, n = z["w"].shape
    conv = quant.ampere_layout_to_b200(torch.from_numpy(z["processed"]), k, n)
    assert torch.equal(conv.reshape(-1), proc.reshape(-1))


def test_quantiser_rejects_bad_input():
    with pytest.raises(RuntimeError):
        quant.symmetric_quantize_last_axis_of_batched_matrix_int8(torch.zeros(4, dtype=torch.float16))
    with pytest.raises(RuntimeError):
        quant.symmetric_quantize_last_axis_of_batched_matrix_int8(torch.zeros(4, 4, dtype=torch.int32))


def test_plain_to_b200_layout():
    lib = capi.load()
    g = np.random.default_rng(3)
    q = g.integers(-128, 128, size=(48, 40)).astype(np.int8)
    out = np.empty(48 * 40, np.uint8)
    assert lib.ftcf_int8_plain_to_b200_host(q.ctypes.data, 48, 40, out.ctypes.data) == 0
    assert np.array_equal(out, Q.to_b200_layout(q).reshape(-1).view(np.uint8))


def test_mmha_split_choice_bounds():
    lib = capi.load()
    for B, H, L in [(1, 40, 1536), (32, 5, 2560), (8, 20, 1536), (1, 40, 16), (1, 1, 20000)]:
        s = lib.ftcf_mmha_choose_splits(B, H, L)
        assert s >= 1 and (L + s - 1) // s <= 4096
