"""Build `libgmf_b200.so` in-tree with nvcc for sm_100a (no torch dependency in the library)."""
from __future__ import annotations

import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
LIB = os.path.join(HERE, "libgmf_b200.so")
SOURCES = ["gmf_api.cu"]
HEADERS = ["common.cuh", "linear_tc.cuh", "attn_args.cuh", "sc_common.cuh", "sc_attn_v9.cuh", "fus_attn_v2.cuh", "ffn_fused.cuh", "kv_proj_all.cuh", "pcn_qkv.cuh", "tail.cuh", "dgr_head.cuh", "dgr_head_api.inl", "dgr_train.cuh", "dgr_train_api.inl", "pdsc_train.cuh", "pdsc_train_api.inl", "matcher.cuh", "matcher_api.inl", "compat_api.inl", "sm_baseline.cuh", "sm_api.inl", "se3_refine.cuh", os.path.join("..", "..", "include", "gmf_b200.h")]


def _stale() -> bool:
    if not os.path.isfile(LIB):
        return True
    t = os.path.getmtime(LIB)
    return any(os.path.getmtime(os.path.join(CSRC, f)) > t for f in SOURCES + HEADERS)


def build(force: bool = False, verbose: bool = False, out: str | None = None, defines=()) -> str:
    """`out` / `defines` build a development variant (e.g. -DGMF_SC_DBG=2) next to the product library; the product is always
    gmf_b200/libgmf_b200.so built with no extra defines."""
    target = out or LIB
    if not force and out is None and not _stale():
        return LIB
    nvcc = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
    cmd = [nvcc, "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-lineinfo", "-std=c++17"] + [f"-D{d}" for d in defines] + [
           "-Xcompiler", "-fPIC", "-shared", "-o", target] + [os.path.join(CSRC, s) for s in SOURCES]
    if verbose:
        cmd.insert(1, "-Xptxas")
        cmd.insert(2, "-v")
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
        raise RuntimeError("nvcc failed building " + target)
    if verbose:
        print(r.stdout + r.stderr)
    return target


if __name__ == "__main__":
    # python -m gmf_b200.build [-v] [--out build/libvariant.so] [-DNAME=VALUE ...]
    argv = sys.argv[1:]
    out = argv[argv.index("--out") + 1] if "--out" in argv else None
    if out:
        os.makedirs(os.path.dirname(os.path.abspath(out)), exist_ok=True)
    print(build(force=True, verbose="-v" in argv, out=out, defines=[x[2:] for x in argv if x.startswith("-D")]))
