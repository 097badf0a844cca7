"""B200-native (sm_100a) decoder for the CodeFuse / GPT-NeoX path.

Layout
  csrc/        CUDA kernels, the C++ engine and the C ABI (include/ftcf.h) -> lib/libftcf.so
  csrc/binding pybind11 shims `libth_gptneox` / `libth_common` (the reference's Python surface) -> lib/
  capi.py      ctypes view of the C ABI
  gptneox_op.py, quant.py   Python mirrors of the reference operator interface over the C ABI
  weights.py   FT-layout weight containers (synthetic init, tensor-parallel split)
  checkpoint.py  HF -> FT checkpoint files, *.q.bin / *.s.bin, per-rank loader (the reference's converter / quant_and_save / loader)
Nothing in this package imports `oracle/` and nothing falls back to a CPU implementation.
"""
from .capi import FtcfError, load  # noqa: F401
