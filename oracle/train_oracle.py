"""TEST INFRASTRUCTURE ONLY — oracle of the GMF-PointDSC training step.

The checker for the CUDA training step is torch autograd of the UNMODIFIED reference: the reference `PointDSC` module in training mode
(models/PointDSC.py:191-266, image backbone bypassed with the image tokens as in oracle/ref_shim.py) followed by the reference's own
`ClassificationLoss` / `SpectralMatchingLoss` (libs/loss.py:66-139), both loaded from `/root/reference` or `oracle/_ref`.
`loss_head_closed_form` restates the loss head and its analytic gradient the way `sm_loss_fused_kernel` / `bce_kernel`
(gmf_b200/csrc/pdsc_train.cuh) compute them; tests/test_pdsc_train.py pins it to autograd of the reference losses on the CPU.
Only `tests/` may import this file.
"""
from __future__ import annotations

import importlib.util
import os
import sys

import torch

from . import ref_shim


def load_reference_losses():
    """(ClassificationLoss, SpectralMatchingLoss) classes of the reference's libs/loss.py, or None when no reference tree is present."""
    pkg, _ = ref_shim.locate()
    if pkg is None:
        return None
    path = os.path.join(pkg, "libs", "loss.py")
    if not os.path.isfile(path):
        return None
    sys.dont_write_bytecode = True
    if pkg not in sys.path:
        sys.path.insert(0, pkg)
    spec = importlib.util.spec_from_file_location("ref_pointdsc_loss", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    return mod.ClassificationLoss, mod.SpectralMatchingLoss


def reference_training_step(sd, cfg, data, balanced=False, w_class=1.0, w_sm=1.0, dtype=torch.float64):
    """One forward + backward of the reference (trainer.py:134-160 with weight_transformation = 0, config_3DMatch.py:52).
    Returns losses, logits, gradients by state_dict key, token gradients and the updated running statistics."""
    losses = load_reference_losses()
    assert losses is not None
    Cls, Sm = losses
    m = ref_shim.build_reference_hot_path(sd, cfg).to(dtype).train()
    p_tok = data["p_tokens"].to(dtype).clone().requires_grad_(True)
    q_tok = data["q_tokens"].to(dtype).clone().requires_grad_(True)
    gt = data["gt_labels"].to(dtype)
    inp = {"corr_pos": data["corr_pos"].to(dtype), "src_keypts": data["src_keypts"].to(dtype), "tgt_keypts": data["tgt_keypts"].to(dtype),
           "p_image": p_tok.permute(0, 2, 1).unsqueeze(2), "q_image": q_tok.permute(0, 2, 1).unsqueeze(2)}
    # cal_seed_trans builds its identity matrices with the default dtype (models/common.py:40-45): run the float64 oracle under it
    old = torch.get_default_dtype()
    torch.set_default_dtype(dtype)
    try:
        res = m(inp)
    finally:
        torch.set_default_dtype(old)
    # the reference losses cast the labels with .float() (libs/loss.py:91-95,138), so they only accept fp32 predictions: the model runs in
    # `dtype`, the two loss classes on fp32 casts of its outputs (the cast is differentiable)
    cl = Cls(balanced=balanced)(res["final_labels"].float(), gt.float())["loss"]
    sl = Sm(balanced=balanced)(res["M"].float(), gt.float())
    loss = w_class * cl + w_sm * sl
    loss.backward()
    grads = {k: (v.grad.detach().clone() if v.grad is not None else None) for k, v in m.named_parameters()}
    state = {k: v.detach().clone() for k, v in m.state_dict().items()}
    return {"class_loss": float(cl), "sm_loss": float(sl), "loss": float(loss), "logits": res["final_labels"].detach(), "grads": grads,
            "final_trans": res["final_trans"].detach(),
            "d_p_tokens": p_tok.grad.detach(), "d_q_tokens": q_tok.grad.detach(), "state": state}


def loss_head_closed_form(feat, logits, gt, sigma, balanced, w_class=1.0, w_sm=1.0):
    """Loss head and analytic gradient as the CUDA kernels compute them.  feat [B,N,C] (un-normalised), logits / gt [B,N], sigma scalar.
    Returns class_loss, sm_loss, d loss / d logits, d loss / d feat (through the normalisation and M only) and d loss / d sigma."""
    B, N, _ = feat.shape
    nrm = feat.norm(dim=-1, keepdim=True).clamp_min(1e-12)
    fh = feat / nrm
    # ClassificationLoss (libs/loss.py:86-93)
    n = B * N
    pos = gt.sum()
    pw = ((n - pos - 1).clamp_min(0) + 1) / ((pos - 1).clamp_min(0) + 1) if balanced else torch.tensor(1.0, dtype=feat.dtype)
    wy = 1 + (pw - 1) * gt
    sp = torch.nn.functional.softplus(-logits)
    class_loss = ((1 - gt) * logits + wy * sp).sum() / n
    dlogit = w_class * ((1 - gt) - wy * (1 - torch.sigmoid(logits))) / n
    # SpectralMatchingLoss (libs/loss.py:118-139) on M = clamp(1 - (1 - fh fh^T) / sigma^2, 0, 1), zero diagonal (PointDSC.py:231-234)
    s = fh @ fh.transpose(1, 2)
    pre = 1 - (1 - s) / sigma ** 2
    m = pre.clamp(0, 1)
    eye = torch.eye(N, dtype=torch.bool)[None]
    posm = ((gt[:, :, None] + gt[:, None, :]) == 2) & ~eye
    k = gt.sum(dim=1)
    kk = k * (k - 1)
    if balanced:
        cp = 1.0 / (((kk - 1).clamp_min(0) + 1) * B)
        cn = 1.0 / (((N * N - kk - 1).clamp_min(0) + 1) * B)
    else:
        cp = cn = torch.full((B,), 2.0 / (B * N * N), dtype=feat.dtype)
    e = torch.where(posm, m - 1, m)
    coef = torch.where(posm, cp[:, None, None], cn[:, None, None])
    dm = (coef * e).masked_fill(eye, 0)
    sm_loss = (0.5 * dm * e).sum()
    inside = (pre >= 0) & (pre <= 1)
    g = torch.where(inside, dm / sigma ** 2, torch.zeros_like(dm))
    dsigma = w_sm * torch.where(inside, dm * 2 * (1 - s) / sigma ** 3, torch.zeros_like(dm)).sum()
    dfh = w_sm * 2 * (g @ fh)
    dfeat = (dfh - fh * (fh * dfh).sum(-1, keepdim=True)) / nrm
    return class_loss, sm_loss, dlogit, dfeat, dsigma
