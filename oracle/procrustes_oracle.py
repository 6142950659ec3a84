"""TEST INFRASTRUCTURE ONLY — DGR weighted Procrustes (GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/core/registration.py:91-113)
restated in torch fp64 on CPU.  Pinned against the reference function itself: oracle/gen_golden_procrustes.py executes the reference's own
source lines (the module imports packages that are absent here) on seeded inputs and commits the result (tests/golden/procrustes_*.npz)."""
from __future__ import annotations

import torch


def weighted_procrustes(X, Y, w, eps):
    X, Y, w = X.double(), Y.double(), w.double()
    W1 = torch.abs(w).sum()                                   # :99
    w_norm = (w / (W1 + eps))[:, None]                        # :100 (w arrives as [N, 1] in the reference's caller)
    mux = (w_norm * X).sum(0, keepdim=True)
    muy = (w_norm * Y).sum(0, keepdim=True)
    Sxy = (Y - muy).t().mm(w_norm * (X - mux))                # :105
    U, D, V = Sxy.svd()
    S = torch.eye(3, dtype=torch.float64)
    if U.det() * V.det() < 0:                                 # :108-109
        S[-1, -1] = -1
    R = U.mm(S.mm(V.t()))
    t = muy.squeeze() - R.mm(mux.t()).squeeze()
    return R, t


def synth_problem(n, seed, inlier=0.4, noise=0.01):
    g = torch.Generator().manual_seed(seed)
    X = torch.rand(n, 3, generator=g) * 3.0
    q, _ = torch.linalg.qr(torch.randn(3, 3, generator=g))
    if torch.det(q) < 0:
        q[:, 2] *= -1
    t = torch.rand(3, generator=g)
    Y = X @ q.T + t + noise * torch.randn(n, 3, generator=g)
    out = torch.rand(n, generator=g) > inlier
    Y[out] = torch.rand(int(out.sum()), 3, generator=g) * 3.0
    w = torch.where(out, 0.05 * torch.rand(n, generator=g), 0.5 + 0.5 * torch.rand(n, generator=g))
    return X, Y, w
