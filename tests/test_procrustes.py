"""DGR weighted Procrustes (first half): oracle pinned to the reference function's own source (golden vectors), CUDA
path (gmf_weighted_procrustes) against the oracle.  Tolerance: rotation 0.01 deg, translation 1 mm (BASELINE.json north_star)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle.procrustes_oracle import synth_problem, weighted_procrustes


def _rot_err_deg(R1, R2):
    return float(torch.rad2deg(2 * torch.asin(torch.clamp((R1.double() - R2.double()).norm() / (2 * 2 ** 0.5), max=1.0))))


def test_oracle_matches_reference_lines():
    g = np.load(os.path.join(ROOT, "tests", "golden", "procrustes_dgr.npz"))
    for name in ("a", "b"):
        X, Y, w = synth_problem(int(g[f"n_{name}"]), int(g[f"seed_{name}"]))
        R, t = weighted_procrustes(X, Y, w, 1e-6)
        assert _rot_err_deg(R, torch.from_numpy(g[f"R_{name}"])) < 1e-3
        assert (t.float() - torch.from_numpy(g[f"t_{name}"])).abs().max() < 1e-5


@pytest.mark.gpu
def test_cuda_weighted_procrustes_matches_oracle_batched():
    from gmf_b200.engine import Engine
    eng = Engine(num_layers=1)
    probs = [synth_problem(1500, s) for s in (3, 4, 5)]
    X, Y, w = (torch.stack([p[i] for p in probs]).cuda() for i in range(3))
    R, t = eng.weighted_procrustes(X, Y, w, eps=1e-6)
    for b, (x, y, ww) in enumerate(probs):
        Rr, tr = weighted_procrustes(x, y, ww, 1e-6)
        assert _rot_err_deg(R[b].cpu(), Rr) < 0.01
        assert (t[b].cpu().double() - tr).abs().max() < 1e-3
        assert abs(float(torch.det(R[b].cpu().double())) - 1.0) < 1e-5
    # degenerate: all weights zero -> finite output (identity-like), no NaN
    R0, t0 = eng.weighted_procrustes(X[:1], Y[:1], torch.zeros_like(w[:1]))
    assert torch.isfinite(R0).all() and torch.isfinite(t0).all()
