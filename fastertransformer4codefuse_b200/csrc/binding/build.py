#!/usr/bin/env python
"""Builds the two pybind11 shims next to libftcf.so:  lib/libth_gptneox.so, lib/libth_common.so.

They are ordinary CPython extension modules named exactly as the reference's (`import libth_gptneox` after
`sys.path.append(lib_path)`, codefuse_example.py:468-470) and link libftcf.so through $ORIGIN.  g++ is called directly
(no JIT cache: the .so files must travel in-tree)."""
import os
import subprocess
import sys
import sysconfig

import torch
from torch.utils import cpp_extension as ce

HERE = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.dirname(os.path.dirname(HERE))
LIB = os.path.join(PKG, "lib")
INC = os.path.join(os.path.dirname(PKG), "include")


def build_one(name: str, src: str, cuda: bool) -> None:
    out = os.path.join(LIB, f"{name}.so")
    srcp = os.path.join(HERE, src)
    deps = [srcp, os.path.join(INC, "ftcf.h"), os.path.abspath(__file__)]
    if os.path.exists(out) and all(os.path.getmtime(out) >= os.path.getmtime(d) for d in deps):
        return
    incs = ce.include_paths("cuda" if cuda else "cpu") + [sysconfig.get_paths()["include"], INC]
    libdirs = ce.library_paths("cuda" if cuda else "cpu") + [LIB]
    cmd = ["g++", "-O2", "-std=c++17", "-fPIC", "-shared", "-fvisibility=hidden", "-Wno-attributes",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", f"-DTORCH_EXTENSION_NAME={name}",
           "-DTORCH_API_INCLUDE_EXTENSION_H"]
    cmd += [f"-I{p}" for p in incs] + [srcp, "-o", out] + [f"-L{p}" for p in libdirs]
    cmd += ["-lftcf", "-ltorch_python", "-ltorch", "-ltorch_cpu", "-lc10"] + (["-lc10_cuda", "-ltorch_cuda"] if cuda else [])
    cmd += ["-Wl,-rpath,$ORIGIN"] + [f"-Wl,-rpath,{p}" for p in ce.library_paths("cpu")]
    subprocess.run(cmd, check=True)


def main() -> None:
    os.makedirs(LIB, exist_ok=True)
    if not os.path.exists(os.path.join(LIB, "libftcf.so")):
        sys.exit("build libftcf.so first (make -C fastertransformer4codefuse_b200/csrc)")
    build_one("libth_common", "th_common.cc", cuda=False)
    build_one("libth_gptneox", "th_gptneox.cc", cuda=True)


if __name__ == "__main__":
    main()
