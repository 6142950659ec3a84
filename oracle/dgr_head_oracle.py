"""TEST INFRASTRUCTURE ONLY — CPU restatement of the DGR bottleneck fusion head.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference legs may import this module; the
product path (gmf_b200/dgr_head.py -> libgmf_b200.so) never does.

Restates `PerceiverIO.forward` with depth=0 from
GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py (line numbers below) in plain torch
ops on CPU, in fp32 or fp64.  Pinned against the UNMODIFIED reference module run in the build container:
tests/golden/dgr_head_*.npz (oracle/gen_golden_dgr.py); tests/test_oracle_golden.py checks the restatement against them.
"""
from __future__ import annotations

from typing import Mapping

import torch
import torch.nn.functional as F


def _ln(x, w, b):                                   # nn.LayerNorm, eps 1e-5 (PreNorm :31-51)
    return F.layer_norm(x, (x.shape[-1],), w, b, 1e-5)


def _cpe(x, w, b):                                  # ConvPosEnc.forward :126-136: depthwise 3-tap conv along tokens + identity
    xt = x.transpose(0, 1).unsqueeze(0)             # [1, C, L]
    y = F.conv1d(xt, w, b, stride=1, padding=1, groups=w.shape[0]) + xt
    return y[0].transpose(0, 1)


def dgr_head_forward(sd: Mapping[str, torch.Tensor], latents: torch.Tensor, image_feat: torch.Tensor, pe: bool = True,
                     dtype=torch.float32, capture: bool = False):
    """latents [M, latent_dim], image_feat [T, dim] -> [M, latent_dim]   (perceiver_io.py:187-221, batch of one)."""
    W = {k: v.to(dtype) for k, v in sd.items()}
    x = latents.to(dtype)
    data = image_feat.to(dtype)
    cap = {}
    if pe:                                          # :197-202
        x = _cpe(x, W["cpe.proj_q.weight"], W["cpe.proj_q.bias"])
        data = _cpe(data, W["cpe.proj_content.weight"], W["cpe.proj_content.bias"])
    a, f = "cross_attend_blocks.0.", "cross_attend_blocks.1."
    # cross_attn(x, context=data) + x   (:208; PreNorm :43-51; Attention.forward :83-102)
    xn = _ln(x, W[a + "norm.weight"], W[a + "norm.bias"])
    cn = _ln(data, W[a + "norm_context.weight"], W[a + "norm_context.bias"])
    q = xn @ W[a + "fn.to_q.weight"].T
    kv = cn @ W[a + "fn.to_kv.weight"].T
    k, v = kv.chunk(2, dim=-1)                      # :91
    d = q.shape[-1]
    sim = (q @ k.T) * d ** -0.5                     # :95, scale = dim_head ** -0.5 (:76), one head
    attn = sim.softmax(dim=-1)                      # :98
    o = attn @ v                                    # :100
    x1 = o @ W[a + "fn.to_out.weight"].T + W[a + "fn.to_out.bias"] + x
    # cross_ff(x) + x   (:211; FeedForward :58-66, GEGLU :53-56: x * gelu(gates), exact erf gelu)
    h = _ln(x1, W[f + "norm.weight"], W[f + "norm.bias"]) @ W[f + "fn.net.0.weight"].T + W[f + "fn.net.0.bias"]
    val, gates = h.chunk(2, dim=-1)
    g = val * F.gelu(gates)
    out = g @ W[f + "fn.net.2.weight"].T + W[f + "fn.net.2.bias"] + x1
    if capture:
        cap.update(x0=x, q=q, k=k, v=v, o=o, x1=x1, g=g)
        return out, cap
    return out


def synth_latents(m: int, seed: int, channels: int = 256) -> torch.Tensor:
    """Stand-in for the sparse U-Net bottleneck features (post BN + ReLU block outputs, resunet_new.py:654-658)."""
    g = torch.Generator().manual_seed(0x0D6A + seed)
    return torch.relu(torch.randn(m, channels, generator=g)) * 1.2
