"""TEST INFRASTRUCTURE ONLY — import the UNMODIFIED reference model.

Source tree, in order: `$GMF_REFERENCE_ROOT`, `/root/reference` (build container), `oracle/_ref` (the byte-for-byte copy made by
`oracle/build_ref.py`; git-ignored, travels to the GPU box with the snapshot).  Two shims, no source edits:
a stub `torchvision.models.utils` module (removed upstream, imported by models/resnet.py:3) and a replacement for the ImageNet
download triggered by `ImageEncoder()` -> `resnet34(pretrained=True)` (models/Img_Encoder.py:13, resnet.py:219-224).
"""
from __future__ import annotations

import os
import sys
import types

HERE = os.path.dirname(os.path.abspath(__file__))


def _roots():
    env = os.environ.get("GMF_REFERENCE_ROOT")
    if env:
        yield env, "reference tree ($GMF_REFERENCE_ROOT)"
    yield "/root/reference", "reference tree (/root/reference)"
    yield os.path.join(HERE, "_ref"), "oracle/_ref (unmodified copy made by oracle/build_ref.py)"


def locate():
    """(package dir holding models/ and utils/, description) or (None, None)."""
    for root, what in _roots():
        pkg = os.path.join(root, "GMF_PointDSC")
        if os.path.isfile(os.path.join(pkg, "models", "PointDSC.py")):
            return pkg, what
    return None, None


def available() -> bool:
    return locate()[0] is not None


def source() -> str:
    return locate()[1] or "absent"


def dgr_file(rel: str):
    """Path of a file of GMF_DeepGlobalRegistration_fcgf (e.g. 'model/perceiver_io.py'), or None."""
    for root, _ in _roots():
        for sub in ("GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf", "dgr_fcgf"):
            p = os.path.join(root, sub, rel)
            if os.path.isfile(p):
                return p
    return None


def load_reference():
    """Returns the reference `PointDSC` class (and patches its backbone download)."""
    pkg, _ = locate()
    if pkg is None:
        raise RuntimeError("reference tree not present (neither /root/reference nor oracle/_ref)")
    sys.dont_write_bytecode = True
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    if "torchvision.models.utils" not in sys.modules:
        stub = types.ModuleType("torchvision.models.utils")
        stub.load_state_dict_from_url = lambda *a, **k: None
        sys.modules["torchvision.models.utils"] = stub
    import models.resnet as resnet  # type: ignore

    resnet.load_state_dict_from_url = (
        lambda url, progress=True: resnet.ResNet(3, resnet.BasicBlock, [3, 4, 6, 3]).state_dict())
    from models.PointDSC import PointDSC  # type: ignore

    return PointDSC


def _construct(cfg):
    PointDSC = load_reference()
    return PointDSC(in_dim=6, num_layers=cfg["num_layers"], num_channels=128, num_iterations=cfg["num_iterations"],
                    ratio=cfg["ratio"], inlier_threshold=cfg["inlier_threshold"], sigma_d=cfg["sigma_d"], k=cfg["k"],
                    nms_radius=cfg["nms_radius"]).eval()


def build_reference(state_dict, cfg):
    """Construct the reference module with `cfg` and load `state_dict` strictly."""
    import torch

    m = _construct(cfg)
    missing = m.load_state_dict(state_dict, strict=True)
    assert not missing.missing_keys and not missing.unexpected_keys
    with torch.no_grad():
        m.sigma_spat.fill_(cfg["sigma_d"])
    return m


def build_reference_hot_path(state_dict, cfg):
    """The unmodified reference module fed with image TOKENS instead of images: loads a hot-path `state_dict` (no backbone keys;
    everything else must match exactly) and bypasses `encoder.image_encoder`.  Use with `forward_tokens`."""
    import torch

    m = _construct(cfg)
    res = m.load_state_dict(state_dict, strict=False)
    assert not res.unexpected_keys, res.unexpected_keys
    assert all(k.startswith("encoder.image_encoder.") or k.endswith("num_batches_tracked") for k in res.missing_keys), res.missing_keys
    with torch.no_grad():
        m.sigma_spat.fill_(cfg["sigma_d"])
    # stand-in for the backbone (outside the timed path): the caller passes the image tokens [B, T, 128] where the reference expects
    # images, reshaped to the [B, 128, 1, T] feature map that NonLocalNet.forward flattens again (PointDSC.py:129-135)
    m.encoder.image_encoder = torch.nn.Identity()
    return m


def forward_tokens(model, corr_pos, src, tgt, p_tok, q_tok, testing=True):
    """reference forward, one pair at a time (the reference asserts bs == 1 in testing mode, PointDSC.py:279,504), stacked."""
    import torch

    outs = []
    with torch.no_grad():
        for b in range(corr_pos.shape[0]):
            data = {"corr_pos": corr_pos[b:b + 1], "src_keypts": src[b:b + 1], "tgt_keypts": tgt[b:b + 1],
                    "p_image": p_tok[b:b + 1].permute(0, 2, 1).unsqueeze(2), "q_image": q_tok[b:b + 1].permute(0, 2, 1).unsqueeze(2)}
            if testing:
                data["testing"] = True
            outs.append(model(data))
    return {"final_trans": torch.cat([o["final_trans"] for o in outs]), "final_labels": torch.cat([o["final_labels"] for o in outs])}
