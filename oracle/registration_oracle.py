"""TEST INFRASTRUCTURE ONLY — DGR's pose solver `GlobalRegistration` (GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/registration.py:
16-66 ortho2rotation, :91-113 weighted_procrustes, :116-132 Transformation, :135-194 GlobalRegistration; core/loss.py:42-61 HighDimSmoothL1Loss)
restated with torch autograd on CPU, plus a loader that executes the reference's OWN source of those functions (the modules import packages
that are absent here - core.knn needs MinkowskiEngine - so the import lines are dropped) to pin the restatement."""
from __future__ import annotations

import os
import re

import numpy as np
import torch

EPS = float(np.finfo(np.float32).eps)


def ortho2rotation(p):                                          # :16-66
    x_raw, y_raw = p[:, 0:3], p[:, 3:6]
    x = x_raw / torch.clamp(torch.sqrt((x_raw ** 2).sum(1, keepdim=True)), min=1e-8)
    proj = (x * y_raw).sum(1, keepdim=True) / torch.clamp((x ** 2).sum(1, keepdim=True), min=1e-8) * x
    y = y_raw - proj
    y = y / torch.clamp(torch.sqrt((y ** 2).sum(1, keepdim=True)), min=1e-8)
    z = torch.cross(x, y, dim=1)
    return torch.stack((x, y, z), dim=2)


def smooth_l1(X, Y, w, q, eps=EPS):                             # loss.py:51-61, w [N,1]
    sq = torch.sum(((X - Y) / q) ** 2, dim=1, keepdim=True)
    half = 0.5 * (sq < 1).to(X.dtype)
    loss = (0.5 - half) * (torch.sqrt(sq + eps) - 0.5) + half * sq
    return (loss * w).sum() / w.sum()


def weighted_procrustes(X, Y, w, eps):                          # :91-113, w [N,1]
    w_norm = w / (torch.abs(w).sum() + eps)
    mux = (w_norm * X).sum(0, keepdim=True)
    muy = (w_norm * Y).sum(0, keepdim=True)
    Sxy = (Y - muy).t().mm(w_norm * (X - mux)).double()
    U, _, Vh = torch.linalg.svd(Sxy)
    V = Vh.t()
    S = torch.eye(3, dtype=torch.float64)
    if torch.det(U) * torch.det(V) < 0:
        S[-1, -1] = -1
    R = U.mm(S.mm(V.t())).to(X.dtype)
    t = (muy.squeeze() - R.mm(mux.t()).squeeze()).to(X.dtype)
    return R, t


def global_registration(X, Y, w, quantization_size=1.0, max_iter=1000, max_break_count=20, break_threshold_ratio=1e-5):
    """X, Y [N,3], w [N,1] -> (R [3,3], t [3], info dict) following :135-194 line by line."""
    R, t = weighted_procrustes(X, Y, w, EPS)
    rot6d = torch.cat([R[:, 0], R[:, 1]])[None].clone().requires_grad_(True)      # Transformation.__init__ :121-125
    trans = t[None].clone().requires_grad_(True)
    fwd = lambda pts: pts @ ortho2rotation(rot6d)[0].t() + trans                # noqa: E731  (:130-132)
    opt = torch.optim.Adam([rot6d, trans], lr=1e-1)
    sched = torch.optim.lr_scheduler.ExponentialLR(opt, gamma=0.999)
    loss_prev = smooth_l1(fwd(X), Y, w, quantization_size).item()
    brk, i, loss = 0, 0, None
    for i in range(max_iter):
        loss = smooth_l1(fwd(X), Y, w, quantization_size)
        if loss.item() < 1e-7:
            break
        opt.zero_grad()
        loss.backward()
        opt.step()
        sched.step()
        if abs(loss_prev - loss.item()) < loss_prev * break_threshold_ratio:
            brk += 1
            if brk >= max_break_count:
                break
        loss_prev = loss.item()
    return ortho2rotation(rot6d.detach())[0], trans.detach()[0], {"iterations": i, "loss": float(loss.item()) if loss is not None else loss_prev, "break_count": brk}


def reference_global_registration():
    """`GlobalRegistration` compiled from the reference's own source text (registration.py + loss.py with their import lines dropped), or None."""
    from oracle import ref_shim
    reg, los = ref_shim.dgr_file("core/registration.py"), ref_shim.dgr_file("core/loss.py")
    if not reg or not los:
        return None
    ns = {"np": np, "torch": torch, "optim": torch.optim, "nn": torch.nn}
    for path in (los, reg):
        text = open(path).read()
        text = re.sub(r"^(from|import) .*$", "", text, flags=re.M)          # numpy / torch / optim are provided above; core.knn is not needed
        exec(compile(text, path, "exec"), ns)
    return ns["GlobalRegistration"]


def synth_problem(n, seed, inlier=0.4, noise=0.01, extent=3.0):
    """Correspondences with predicted-inlier-like weights: inliers get w in [0.5, 1], outliers w in [0, 0.05] (a few confident outliers)."""
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, 3, generator=g) * extent
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 2] *= -1
    t = torch.rand(3, generator=g)
    Y = X @ q.T + t + noise * torch.randn(n, 3, generator=g)
    out = torch.rand(n, generator=g) > inlier
    Y[out] = torch.rand(int(out.sum()), 3, generator=g) * extent
    w = torch.where(out, 0.05 * torch.rand(n, generator=g), 0.5 + 0.5 * torch.rand(n, generator=g))
    conf = out & (torch.rand(n, generator=g) < 0.02)
    w[conf] = 0.9
    return X, Y, w[:, None].contiguous()
