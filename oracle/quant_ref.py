"""Oracle (numpy restatement) of the reference's weight-only INT8 quantiser.

TEST INFRASTRUCTURE -- never imported by the product path.

Follows, function by function (all paths relative to /root/reference):
  * symmetric_quantize<half,half/float>
        src/fastertransformer/kernels/cutlass_kernels/cutlass_preprocessors.cc:577-673
  * permute_B_rows_for_mixed_gemm                      ... :133-201
  * subbyte_transpose (int8 case)                      ... :207-348
  * interleave_column_major_tensor                     ... :437-498
  * add_bias_and_interleave_int8s_inplace              ... :350-370
  * preprocess_weights_for_mixed_gemm (sm80 details)   ... :500-539
    layout details for uint8 on sm>=75: ColumnMajorTileInterleave<64, 2>
        src/fastertransformer/cutlass_extensions/include/cutlass_extensions/gemm/kernel/mixed_gemm_B_layout.h:59-72
  * binding: symmetric_quantize_last_axis_of_batched_matrix_int8
        src/fastertransformer/th_op/common/WeightOnlyQuantOps.cc:140-233

Pinned against: the golden vectors in
tests/weight_only_quant_ops/th_weight_quant_ops_unit_tests.py:36-39,110-116,133
(tests/test_oracle_quant.py) and, when oracle/_ref/libref_quant.so is built,
against the reference's own object code compiled from cutlass_preprocessors.cc.

Besides the reference ("Ampere") layout this module restates the B200-native
layout our library emits -- W^T, i.e. [n, k] with k contiguous, biased to
uint8 (+128) -- so tests can go between the two.
"""
from __future__ import annotations

import numpy as np

# tests/weight_only_quant_ops/th_weight_quant_ops_unit_tests.py:36-37
ROW_PERMUTATION_INT8 = np.array([0, 1, 8, 9, 2, 3, 10, 11, 4, 5, 12, 13, 6, 7, 14, 15])

ROWS_PER_COLUMN_TILE = 64   # ThreadblockK, mixed_gemm_B_layout.h:62
COLUMNS_INTERLEAVED = 2     # 128 B cache line / 64


def _round_half_away(x: np.ndarray) -> np.ndarray:
    """C `round()` (cutlass_preprocessors.cc:631): halves away from zero."""
    return np.sign(x) * np.floor(np.abs(x) + np.float32(0.5))


def symmetric_quantize_unprocessed(w: np.ndarray):
    """Per-column symmetric int8 quantisation of w[k, n] (or [e, k, n]).

    cutlass_preprocessors.cc:603-640.  Scale = absmax/128 in fp32; the value
    divided by is the fp32 scale (NOT the fp16-rounded one that is stored).
    Returns (q int8 same shape, scale fp32 [n] / [e, n]).
    """
    w32 = np.asarray(w).astype(np.float32)
    col_max = np.abs(w32).max(axis=-2)                       # :613-618
    scale = (col_max * np.float32(1.0 / 128.0)).astype(np.float32)   # :622-625
    with np.errstate(divide="ignore", invalid="ignore"):
        scaled = w32 / scale[..., None, :]
    r = _round_half_away(scaled.astype(np.float32))
    # int8_t(std::max(-128.f, std::min(127.f, x))) -- std::min(127, NaN) -> 127
    r = np.where(np.isnan(r), np.float32(127.0), r)
    q = np.clip(r, -128.0, 127.0).astype(np.int8)
    return q, scale


def permute_b_rows(q: np.ndarray) -> np.ndarray:
    """cutlass_preprocessors.cc:133-201 for int8: inside each group of 16 rows,
    out[row] = in[perm[row]]."""
    k, n = q.shape[-2:]
    assert k % 16 == 0 and n % 8 == 0, "rows % 16 and cols % 8 (cutlass_preprocessors.cc:170-177)"
    t = q.reshape(q.shape[:-2] + (k // 16, 16, n))
    return t[..., ROW_PERMUTATION_INT8, :].reshape(q.shape)


def subbyte_transpose_int8(q: np.ndarray) -> np.ndarray:
    """cutlass_preprocessors.cc:207-348 (int8): data moves to column-major while
    the logical shape is kept (test: th_weight_quant_ops_unit_tests.py:133)."""
    return np.ascontiguousarray(np.swapaxes(q, -1, -2)).reshape(q.shape)


def interleave_column_major(q_colmajor: np.ndarray, k: int, n: int) -> np.ndarray:
    """cutlass_preprocessors.cc:437-498 with rows_per_tile=64, interleave=2.

    Input: bytes of a column-major [k, n] matrix (flat view [n, k]).  Works on
    32-bit words (4 rows each).  Output has the same number of bytes.
    """
    assert k % 4 == 0 and n % ROWS_PER_COLUMN_TILE == 0
    lead = q_colmajor.shape[:-2]
    src = q_colmajor.reshape(lead + (n, k // 4, 4))
    vec_rows = k // 4
    vrt = ROWS_PER_COLUMN_TILE // 4           # vec_rows_per_tile = 16
    il = COLUMNS_INTERLEAVED
    dst = np.empty(lead + (n // il, vec_rows * il, 4), dtype=q_colmajor.dtype)
    read_col = np.arange(n)[:, None]
    vec_read_row = np.arange(vec_rows)[None, :]
    base_vec_row = (vec_read_row // vrt) * vrt
    vec_write_row = il * base_vec_row + vrt * (read_col % il) + vec_read_row % vrt
    write_col = np.broadcast_to(read_col // il, vec_write_row.shape)
    dst[..., write_col, vec_write_row, :] = src
    return dst.reshape(q_colmajor.shape)


def add_bias_and_interleave_int8s(q: np.ndarray) -> np.ndarray:
    """cutlass_preprocessors.cc:350-370: +128 (wrapping) then swap bytes 1<->2 of
    every 4.  Golden: th_weight_quant_ops_unit_tests.py:110-116."""
    flat = (q.astype(np.int16) + 128).astype(np.uint8).reshape(-1, 4)
    out = flat[:, [0, 2, 1, 3]]
    return np.ascontiguousarray(out).reshape(q.shape).view(np.int8)


def preprocess_weights_ampere(q: np.ndarray) -> np.ndarray:
    """cutlass_preprocessors.cc:500-539 for sm80 + int8.  q: int8 [k, n]."""
    k, n = q.shape[-2:]
    x = permute_b_rows(q)
    x = subbyte_transpose_int8(x)
    x = interleave_column_major(x, k, n)
    return add_bias_and_interleave_int8s(x)


def unprocess_weights_ampere(p: np.ndarray, k: int, n: int) -> np.ndarray:
    """Inverse of preprocess_weights_ampere: processed bytes -> plain int8 [k, n]."""
    lead = p.shape[:-2] if p.ndim > 2 else ()
    flat = p.reshape(-1, 4).view(np.uint8)[:, [0, 2, 1, 3]]
    x = (flat.astype(np.int16) - 128).astype(np.int8).reshape(lead + (n // COLUMNS_INTERLEAVED, (k // 4) * COLUMNS_INTERLEAVED, 4))
    vec_rows = k // 4
    vrt = ROWS_PER_COLUMN_TILE // 4
    il = COLUMNS_INTERLEAVED
    read_col = np.arange(n)[:, None]
    vec_read_row = np.arange(vec_rows)[None, :]
    base_vec_row = (vec_read_row // vrt) * vrt
    vec_write_row = il * base_vec_row + vrt * (read_col % il) + vec_read_row % vrt
    write_col = np.broadcast_to(read_col // il, vec_write_row.shape)
    colmajor = x[..., write_col, vec_write_row, :]            # [n, k/4, 4]
    permuted = np.swapaxes(colmajor.reshape(lead + (n, k)), -1, -2)   # [k, n] row-permuted
    inv = np.argsort(ROW_PERMUTATION_INT8)
    t = permuted.reshape(lead + (k // 16, 16, n))
    return np.ascontiguousarray(t[..., inv, :].reshape(lead + (k, n)))


def symmetric_quantize_last_axis_of_batched_matrix_int8(w: np.ndarray):
    """The reference binding (WeightOnlyQuantOps.cc:140-233): returns
    (processed int8 with the input's shape, scale [n] in the weight's dtype)."""
    q, scale = symmetric_quantize_unprocessed(w)
    return preprocess_weights_ampere(q), scale.astype(np.asarray(w).dtype)


# ---------------------------------------------------------------- B200 layout
def to_b200_layout(q: np.ndarray) -> np.ndarray:
    """Plain int8 [k, n] -> the layout our kernels stream: W^T [n, k], k
    contiguous, stored as uint8 = q + 128 (bytes returned in an int8 array of
    the *original* [k, n] shape, as the binding does)."""
    t = np.ascontiguousarray(np.swapaxes(q, -1, -2))
    u = (t.astype(np.int16) + 128).astype(np.uint8)
    return u.view(np.int8).reshape(q.shape)


def from_b200_layout(p: np.ndarray, k: int, n: int) -> np.ndarray:
    lead = p.shape[:-2] if p.ndim > 2 else ()
    u = p.reshape(lead + (n, k)).view(np.uint8)
    return np.ascontiguousarray(np.swapaxes((u.astype(np.int16) - 128).astype(np.int8), -1, -2))


def dequantize(q: np.ndarray, scale: np.ndarray) -> np.ndarray:
    """fp32 view of what the GEMM multiplies by: q * scale (scale as stored,
    i.e. already rounded to the weight dtype)."""
    return q.astype(np.float32) * np.asarray(scale).astype(np.float32)[..., None, :]
