"""TEST INFRASTRUCTURE ONLY — golden vectors of DGR's weighted_procrustes from the reference's own source.
core/registration.py imports core.knn / core.loss (MinkowskiEngine, open3d: absent), so the function's source lines (91-113) are read
from /root/reference at run time and executed; nothing is copied into the repository.   python oracle/gen_golden_procrustes.py"""
from __future__ import annotations

import os
import sys
import textwrap

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from oracle.procrustes_oracle import synth_problem     # noqa: E402

REF = "/root/reference/GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/registration.py"


def main():
    lines = open(REF).read().split("\n")
    src = textwrap.dedent("\n".join(lines[90:113]))
    assert src.startswith("def weighted_procrustes(X, Y, w, eps):"), "reference lines moved"
    ns = {"torch": torch, "np": np}
    exec(compile(src, REF, "exec"), ns)
    out = {}
    for name, (n, seed) in {"a": (2000, 1), "b": (317, 2)}.items():
        X, Y, w = synth_problem(n, seed)
        R, t = ns["weighted_procrustes"](X, Y, w[:, None], 1e-6)
        out[f"R_{name}"], out[f"t_{name}"] = R.numpy(), t.numpy()
        out[f"n_{name}"], out[f"seed_{name}"] = n, seed
    path = os.path.join(ROOT, "tests", "golden", "procrustes_dgr.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path))


if __name__ == "__main__":
    main()
