"""TEST INFRASTRUCTURE ONLY — golden vectors of the correspondence construction from the reference's own code.

The block lives inside `ThreeDMatchTest.__getitem__` (datasets/ThreeDMatch.py), which needs the dataset on disk, so the script
reads the reference's source lines 384-391 (matcher) and 401-402, 412-414 (gather, corr_pos) from /root/reference at run time and
executes exactly those statements on seeded inputs.  Nothing is copied into the repository.   python oracle/gen_golden_matcher.py
"""
from __future__ import annotations

import os
import sys
import textwrap
import types

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.matcher_oracle import synth_descriptors    # noqa: E402

REF = "/root/reference/GMF_PointDSC/datasets/ThreeDMatch.py"
CASES = {"matcher_n700_m650_d32": dict(ns=700, nt=650, d=32, seed=3, mutual=False),
         "matcher_n900_m1000_d33_mutual": dict(ns=900, nt=1000, d=33, seed=4, mutual=True)}


def reference_block():
    lines = open(REF).read().split("\n")
    groups = [lines[383:391], lines[400:402], lines[411:414]]        # 1-based :384-391, :401-402, :412-414 (ThreeDMatchTest.__getitem__)
    src = "\n".join(textwrap.dedent("\n".join(g)) for g in groups)
    assert "np.argmin(distance, axis=1)" in src and "corr_pos.mean(0)" in src, "reference lines moved"
    return compile(src, REF, "exec")


def main():
    code = reference_block()
    for name, c in CASES.items():
        s, t, sk, tk = synth_descriptors(c["ns"], c["nt"], c["d"], c["seed"])
        ns = dict(np=np, self=types.SimpleNamespace(use_mutual=c["mutual"], in_dim=6), src_desc=s, tgt_desc=t, src_keypts=sk, tgt_keypts=tk)
        exec(code, ns)
        path = os.path.join(ROOT, "tests", "golden", name + ".npz")
        np.savez_compressed(path, ns=c["ns"], nt=c["nt"], d=c["d"], seed=c["seed"], mutual=int(c["mutual"]),
                            source_idx=ns["source_idx"].astype(np.int32), corr=ns["corr"].astype(np.int32),
                            corr_pos=ns["corr_pos"].astype(np.float32))
        print(name, ns["corr"].shape, os.path.getsize(path))


if __name__ == "__main__":
    main()
