"""state_dict contract of the GMF-PointDSC hot path (everything except the image backbone).

The key names and shapes are the drop-in contract with the reference module
(reference: GMF_PointDSC/models/PointDSC.py:147-190 constructor, fusion_layer.py:131-170,
SURVEY.md Appendix A).  ``hot_path_spec`` enumerates, in one canonical order, every tensor the
CUDA path consumes; the C-ABI (`gmf_weight_count` / `gmf_weight_spec`, include/gmf_b200.h)
exposes the same table from the native side and `pack_state_dict` cross-checks the two.
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import torch

C = 128          # num_channels (kernels are specialised on it)
DH = 64          # fusion cross-attention head dim (cross_dim_head = C // 2)
FF = 1024        # GEGLU feed-forward width (dim * mult * 2)


def _fusion_spec(prefix: str, pe: bool) -> List[Tuple[str, Tuple[int, ...]]]:
    """FusionLayer(dim=128, latent_dim=128, cross_heads=1, cross_dim_head=64, depth=0, pe=pe)."""
    s: List[Tuple[str, Tuple[int, ...]]] = []
    if pe:  # ConvPosEnc, fusion_layer.py:97-128 (depthwise k=3)
        s += [(prefix + "cpe.proj_q.weight", (C, 1, 3)), (prefix + "cpe.proj_q.bias", (C,)),
              (prefix + "cpe.proj_content.weight", (C, 1, 3)), (prefix + "cpe.proj_content.bias", (C,))]
    a = prefix + "cross_attend_blocks.0."
    f = prefix + "cross_attend_blocks.1."
    s += [(a + "norm.weight", (C,)), (a + "norm.bias", (C,)),
          (a + "norm_context.weight", (C,)), (a + "norm_context.bias", (C,)),
          (a + "fn.to_q.weight", (DH, C)), (a + "fn.to_kv.weight", (2 * DH, C)),
          (a + "fn.to_out.weight", (C, DH)), (a + "fn.to_out.bias", (C,)),
          (f + "norm.weight", (C,)), (f + "norm.bias", (C,)),
          (f + "fn.net.0.weight", (FF, C)), (f + "fn.net.0.bias", (FF,)),
          (f + "fn.net.2.weight", (C, FF // 2)), (f + "fn.net.2.bias", (C,))]
    return s


def _bn_spec(prefix: str, ch: int) -> List[Tuple[str, Tuple[int, ...]]]:
    return [(prefix + "weight", (ch,)), (prefix + "bias", (ch,)),
            (prefix + "running_mean", (ch,)), (prefix + "running_var", (ch,))]


def hot_path_spec(num_layers: int, in_dim: int = 6) -> "OrderedDict[str, Tuple[int, ...]]":
    """Ordered name -> shape table of every non-backbone tensor the path reads."""
    s: List[Tuple[str, Tuple[int, ...]]] = [("sigma", (1,)), ("sigma_spat", (1,)),
                                            ("encoder.layer0.weight", (C, in_dim, 1)),
                                            ("encoder.layer0.bias", (C,))]
    s += _fusion_spec("encoder.fusion_layer_1.", pe=False)
    for i in range(num_layers):
        p = f"encoder.blocks.PointCN_layer_{i}."
        s += [(p + "0.weight", (C, C, 1)), (p + "0.bias", (C,))] + _bn_spec(p + "1.", C)
        n = f"encoder.blocks.NonLocal_layer_{i}."
        s += [(n + "fc_message.0.weight", (C // 2, C, 1)), (n + "fc_message.0.bias", (C // 2,))]
        s += _bn_spec(n + "fc_message.1.", C // 2)
        s += [(n + "fc_message.3.weight", (C // 2, C // 2, 1)), (n + "fc_message.3.bias", (C // 2,))]
        s += _bn_spec(n + "fc_message.4.", C // 2)
        s += [(n + "fc_message.6.weight", (C, C // 2, 1)), (n + "fc_message.6.bias", (C,))]
        for q in "qkv":
            s += [(n + f"projection_{q}.weight", (C, C, 1)), (n + f"projection_{q}.bias", (C,))]
        s += _fusion_spec(n + "fusion_layer_2.", pe=True)
    s += [("classification.0.weight", (32, C, 1)), ("classification.0.bias", (32,)),
          ("classification.2.weight", (32, 32, 1)), ("classification.2.bias", (32,)),
          ("classification.4.weight", (1, 32, 1)), ("classification.4.bias", (1,))]
    return OrderedDict(s)


def pack_state_dict(state_dict: Dict[str, torch.Tensor], num_layers: int, in_dim: int = 6) -> torch.Tensor:
    """Flatten the hot-path tensors of a (reference-compatible) state_dict into one fp32 host buffer
    in `hot_path_spec` order.  Missing keys / wrong shapes raise, like load_state_dict(strict=True)."""
    chunks = []
    for name, shape in hot_path_spec(num_layers, in_dim).items():
        if name not in state_dict:
            raise KeyError(f"state_dict is missing hot-path tensor {name!r}")
        t = state_dict[name].detach().to("cpu", torch.float32)
        if tuple(t.shape) != tuple(shape):
            raise ValueError(f"{name}: expected shape {shape}, got {tuple(t.shape)}")
        chunks.append(t.reshape(-1))
    return torch.cat(chunks).contiguous()
