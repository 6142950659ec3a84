"""TEST INFRASTRUCTURE ONLY — recipe that makes the UNMODIFIED reference travel to the GPU box.

The reference path is pure Python/PyTorch (no C sources to compile), so "building" it means copying the handful of modules the
path imports, byte for byte, from the read-only tree `/root/reference` into `oracle/_ref/` (git-ignored, NOT gpurun-ignored: it
ships with the snapshot like the built `.so`).  Nothing is edited; `oracle/ref_shim.py` applies the same two import shims it
applies to `/root/reference`.  Run by `__graft_entry__.build()` in the build container; on the GPU box the
prebuilt copy is used as is.  Only `tests/`, `smoke()` and bench.py's CPU legs ever import from `oracle/_ref`.

    python oracle/build_ref.py        # -> oracle/_ref/GMF_PointDSC/{models,utils}/..., oracle/_ref/dgr_fcgf/..., MANIFEST.json
"""
from __future__ import annotations

import hashlib
import json
import os
import shutil

HERE = os.path.dirname(os.path.abspath(__file__))
REF_ROOT = os.environ.get("GMF_REFERENCE_ROOT", "/root/reference")
OUT = os.path.join(HERE, "_ref")

# (source relative to the reference root, destination relative to oracle/_ref)
FILES = [
    ("GMF_PointDSC/models/PointDSC.py", "GMF_PointDSC/models/PointDSC.py"),
    ("GMF_PointDSC/models/fusion_layer.py", "GMF_PointDSC/models/fusion_layer.py"),
    ("GMF_PointDSC/models/common.py", "GMF_PointDSC/models/common.py"),
    ("GMF_PointDSC/models/Img_Encoder.py", "GMF_PointDSC/models/Img_Encoder.py"),
    ("GMF_PointDSC/models/resnet.py", "GMF_PointDSC/models/resnet.py"),
    ("GMF_PointDSC/utils/SE3.py", "GMF_PointDSC/utils/SE3.py"),
    ("GMF_PointDSC/utils/__init__.py", "GMF_PointDSC/utils/__init__.py"),
    # training losses (§8f N2: training step of the path)
    ("GMF_PointDSC/libs/loss.py", "GMF_PointDSC/libs/loss.py"),
    # DGR bottleneck fusion head (cfg#5) and the pose solvers of §8f N3
    ("GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py", "dgr_fcgf/model/perceiver_io.py"),
    ("GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/registration.py", "dgr_fcgf/core/registration.py"),
    ("GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/loss.py", "dgr_fcgf/core/loss.py"),
    # classical spectral-matching baseline (§8f N4)
    ("GMF_PointDSC/baseline_scripts/baseline_3DMatch.py", "GMF_PointDSC/baseline_scripts/baseline_3DMatch.py"),
]


def available() -> bool:
    return os.path.isfile(os.path.join(REF_ROOT, FILES[0][0]))


def build(verbose: bool = False) -> str | None:
    """Copy the reference modules (unchanged) into oracle/_ref.  Returns the output directory, or None when the reference tree is
    not present (GPU box: the directory shipped with the snapshot is used)."""
    if not available():
        return OUT if os.path.isdir(OUT) else None
    manifest = {}
    for src_rel, dst_rel in FILES:
        src = os.path.join(REF_ROOT, src_rel)
        if not os.path.isfile(src):
            continue
        dst = os.path.join(OUT, dst_rel)
        os.makedirs(os.path.dirname(dst), exist_ok=True)
        shutil.copyfile(src, dst)
        with open(dst, "rb") as f:
            manifest[dst_rel] = {"source": src_rel, "sha256": hashlib.sha256(f.read()).hexdigest()}
        if verbose:
            print("copied", src_rel)
    with open(os.path.join(OUT, "MANIFEST.json"), "w") as f:
        json.dump(manifest, f, indent=1, sort_keys=True)
    return OUT


if __name__ == "__main__":
    print(build(verbose=True))
