"""Deterministic synthetic weights and 3DMatch/KITTI-shaped fragment pairs.

Nothing here depends on the reference or on `oracle/`: the same generator feeds the CUDA path,
the oracle and the golden-fixture script, so all three see bit-identical inputs on any machine
with the same torch build (CPU generators only).
"""
from __future__ import annotations

import zlib
from typing import Dict, Mapping, Tuple

import torch


def _gen(seed: int, name: str) -> torch.Generator:
    g = torch.Generator(device="cpu")
    g.manual_seed((seed * 1000003 + zlib.crc32(name.encode())) % (2 ** 63))
    return g


def synth_state_dict(shapes: Mapping[str, Tuple[int, ...]], seed: int = 0, plain_init: bool = False) -> Dict[str, torch.Tensor]:
    """Name-seeded random weights.  Conv/Linear weights ~ xavier-normal (PointDSC.py:183-188);
    norm scales/shifts, biases and BN running statistics are mildly perturbed (unless
    `plain_init`) so BN folding, LayerNorm affine and bias paths are actually exercised."""
    out: Dict[str, torch.Tensor] = {}
    for name, shape in shapes.items():
        g = _gen(seed, name)
        shape = tuple(shape)
        if name == "sigma":
            t = torch.tensor([1.0 if plain_init else 0.9])
        elif name == "sigma_spat":
            t = torch.tensor([0.10])            # overwritten by the module constructor's sigma_d
        elif name.endswith("num_batches_tracked"):
            t = torch.zeros((), dtype=torch.int64)
        elif name.endswith("running_mean"):
            t = torch.zeros(shape) if plain_init else 0.05 * torch.randn(shape, generator=g)
        elif name.endswith("running_var"):
            t = torch.ones(shape) if plain_init else 1.0 + 0.2 * torch.rand(shape, generator=g)
        elif len(shape) >= 2:                    # conv / linear weight
            rf = 1
            for d in shape[2:]:
                rf *= d
            fan_in, fan_out = shape[1] * rf, shape[0] * rf
            if "cpe.proj" in name:               # depthwise taps: keep O(0.3)
                t = 0.3 * torch.randn(shape, generator=g)
            elif len(shape) == 4:                # backbone conv2d: kaiming fan_out
                t = torch.randn(shape, generator=g) * (2.0 / fan_out) ** 0.5
            else:
                t = torch.randn(shape, generator=g) * (2.0 / (fan_in + fan_out)) ** 0.5
        elif name.endswith("weight"):            # LayerNorm / BatchNorm scale
            t = torch.ones(shape) if plain_init else 1.0 + 0.1 * torch.randn(shape, generator=g)
        else:                                    # biases
            t = (0.02 if plain_init else 0.05) * torch.randn(shape, generator=g)
        out[name] = t.to(torch.float32) if t.dtype != torch.int64 else t
    return out


def synth_tokens(batch: int, tokens: int, seed: int, channels: int = 128) -> torch.Tensor:
    """Stand-in for backbone output: non-negative (post-ReLU) token features, mean|x| ~ 2 like the
    random-init ResNet trunk."""
    g = _gen(seed, "tokens")
    return torch.relu(2.5 * torch.randn(batch, tokens, channels, generator=g) + 1.0).contiguous()


def synth_pairs(batch: int, n: int, seed: int = 0, extent: float = 3.0, inlier_ratio: float = 0.30,
                noise: float = 0.0) -> Dict[str, torch.Tensor]:
    """`batch` synthetic fragment pairs with `n` putative correspondences each.

    src ~ U[0,L)^3, rigid motion (R,t), tgt = R src + t (+ N(0,noise^2)) for inliers and U[0,L)^3
    for outliers; corr_pos = [src|tgt] - mean_N (datasets/ThreeDMatch.py:411-414)."""
    src_l, tgt_l, gt_l, lab_l = [], [], [], []
    for b in range(batch):
        g = _gen(seed, f"pair{b}")
        src = torch.rand(n, 3, generator=g) * extent
        q, r = torch.linalg.qr(torch.randn(3, 3, generator=g))
        q = q * torch.sign(torch.diagonal(r))[None, :]
        if torch.det(q) < 0:
            q[:, 2] = -q[:, 2]
        t = torch.rand(3, generator=g) * (extent / 3.0)
        tgt = src @ q.T + t
        if noise > 0:
            tgt = tgt + noise * torch.randn(n, 3, generator=g)
        inl = torch.rand(n, generator=g) < inlier_ratio
        rnd = torch.rand(n, 3, generator=g) * extent
        tgt = torch.where(inl[:, None], tgt, rnd)
        gt = torch.eye(4)
        gt[:3, :3], gt[:3, 3] = q, t
        src_l.append(src), tgt_l.append(tgt), gt_l.append(gt), lab_l.append(inl.float())
    src, tgt = torch.stack(src_l), torch.stack(tgt_l)
    corr = torch.cat([src, tgt], dim=-1)
    corr = corr - corr.mean(dim=1, keepdim=True)
    return {"corr_pos": corr.contiguous(), "src_keypts": src.contiguous(), "tgt_keypts": tgt.contiguous(),
            "gt_trans": torch.stack(gt_l), "gt_labels": torch.stack(lab_l)}
