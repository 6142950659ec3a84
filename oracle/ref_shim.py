"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference model in the build container.

Works only where `/root/reference` exists (never on the GPU box).  Two shims, no source edits
(SURVEY.md Appendix B): a stub `torchvision.models.utils` module (removed upstream, imported by
models/resnet.py:3) and a replacement for the ImageNet download triggered by
`ImageEncoder()` -> `resnet34(pretrained=True)` (models/Img_Encoder.py:13, resnet.py:219-224).
"""
from __future__ import annotations

import os
import sys
import types

REF_ROOT = os.environ.get("GMF_REFERENCE_ROOT", "/root/reference")
REF_PKG = os.path.join(REF_ROOT, "GMF_PointDSC")


def available() -> bool:
    return os.path.isfile(os.path.join(REF_PKG, "models", "PointDSC.py"))


def load_reference():
    """Returns the reference `PointDSC` class (and patches its backbone download)."""
    if not available():
        raise RuntimeError("reference tree not present")
    sys.dont_write_bytecode = True
    if REF_PKG not in sys.path:
        sys.path.insert(0, REF_PKG)
    if "torchvision.models.utils" not in sys.modules:
        stub = types.ModuleType("torchvision.models.utils")
        stub.load_state_dict_from_url = lambda *a, **k: None
        sys.modules["torchvision.models.utils"] = stub
    import models.resnet as resnet  # type: ignore

    resnet.load_state_dict_from_url = (
        lambda url, progress=True: resnet.ResNet(3, resnet.BasicBlock, [3, 4, 6, 3]).state_dict())
    from models.PointDSC import PointDSC  # type: ignore

    return PointDSC


def build_reference(state_dict, cfg):
    """Construct the reference module with `cfg` and load `state_dict` strictly."""
    import torch

    PointDSC = load_reference()
    m = PointDSC(in_dim=6, num_layers=cfg["num_layers"], num_channels=128, num_iterations=cfg["num_iterations"],
                 ratio=cfg["ratio"], inlier_threshold=cfg["inlier_threshold"], sigma_d=cfg["sigma_d"], k=cfg["k"],
                 nms_radius=cfg["nms_radius"]).eval()
    missing = m.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        m.sigma_spat.fill_(cfg["sigma_d"])
    return m
