"""Build `libgmf_b200.so` in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgmf_b200.so")
SOURCES = ["gmf_api.cu"]
HEADERS = ["common.cuh", "linear_tc.cuh", "attn_args.cuh", "sc_common.cuh", "sc_attn_v9.cuh", "fus_attn_v2.cuh", "ffn_fused.cuh", "pcn_qkv.cuh", "tail.cuh", "dgr_head.cuh", "dgr_head_api.inl", "matcher.cuh", "matcher_api.inl", "compat_api.inl", os.path.join("..", "..", "include", "gmf_b200.h")]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False) -> str:
    if not force and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17", "-DGMF_SC_DBG=" + os.environ.get("GMF_SC_DBG", "0"),] + (["-DGMF_FFN_TRACE"] if os.environ.get("GMF_FFN_TRACE") else []) + (["-DGMF_PCN_TRACE"] if os.environ.get("GMF_PCN_TRACE") else []) + [
           "-Xcompiler", "-fPIC", "-shared", "-o", LIB] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building libgmf_b200.so")
    if verbose:
        print(r.stdout + r.stderr)
    return LIB


if __name__ == "__main__":
    print(build(force=True, verbose="-v" in sys.argv))
