"""TEST INFRASTRUCTURE ONLY — the classical spectral-matching baseline `SM` (GMF_PointDSC/baseline_scripts/baseline_3DMatch.py:19-53)
restated in torch on CPU, and a loader that executes the reference's OWN source lines of `SM` (the script imports open3d / datasets at
module level, which are absent here, so only the function body is compiled) to pin the restatement."""
from __future__ import annotations

import re
import types

import torch

from oracle import pointdsc_oracle as O


def sm_baseline(corr, src_keypts, tgt_keypts, inlier_threshold=0.10, top_ratio=0.1, iters=10, dtype=torch.float32):
    """corr [N,6] (src | tgt), src_keypts / tgt_keypts [1,N,3] -> (trans [1,4,4], labels [1,N], leading_eig [1,N])."""
    corr, src_keypts, tgt_keypts = corr.to(dtype), src_keypts.to(dtype), tgt_keypts.to(dtype)
    diff = corr[:, None, :] - corr[None, :, :]                                                       # :20
    M = diff[:, :, 0:3].pow(2).sum(-1).sqrt() - diff[:, :, 3:6].pow(2).sum(-1).sqrt()                # :21
    M = M[None]
    sigma = inlier_threshold / 3                                                                     # :33
    M = torch.clamp(4.5 - M ** 2 / 2 / sigma ** 2, min=0)                                            # :34
    n = M.shape[1]
    M[:, torch.arange(n), torch.arange(n)] = 0                                                       # :35
    v = torch.ones_like(M[:, :, 0:1])
    for _ in range(iters):                                                                           # :38-41
        v = torch.bmm(M, v)
        v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
    v = v.squeeze(-1)
    top = torch.argsort(v, dim=1, descending=True)[:, 0:int(n * top_ratio)]                          # :45
    labels = torch.zeros_like(v)
    labels[0, top[0]] = 1
    trans = O.rigid_transform_3d(src_keypts, tgt_keypts, v * labels)                                 # :51
    return trans, labels, v


def reference_sm():
    """The reference's `SM` function compiled from its own source text (baseline_3DMatch.py:19-53), or None when the tree is absent."""
    import os

    from oracle import ref_shim
    pkg, _ = ref_shim.locate()
    if pkg is None:
        return None
    path = os.path.join(pkg, "baseline_scripts", "baseline_3DMatch.py")
    if not os.path.isfile(path):
        return None
    text = open(path).read()
    m = re.search(r"^def SM\(.*?(?=^def )", text, re.S | re.M)
    if not m:
        return None
    ref_shim.load_reference()                                  # puts the reference package on sys.path (models.common)
    from models.common import rigid_transform_3d               # type: ignore
    ns = {"torch": torch, "rigid_transform_3d": rigid_transform_3d}
    exec(compile(m.group(0), path, "exec"), ns)
    fn = ns["SM"]
    return lambda corr, s, t, thr, ratio=0.1: fn(corr, s, t, types.SimpleNamespace(inlier_threshold=thr), top_ratio=ratio)
