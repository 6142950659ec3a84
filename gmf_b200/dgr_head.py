"""Drop-in mirror of the DGR inlier network's bottleneck fusion head.

Reference: `PerceiverIO` in GMF_DeepGlobalRegistration/GMF_DeepGlobalRegistration_fcgf/model/perceiver_io.py:140-221,
instantiated at model/resunet_new.py:516-525 as

    PerceiverIO(dim=128, depth=0, latent_dim=256, cross_heads=1, latent_heads=8, cross_dim_head=128,
                latent_dim_head=128, pe=True)

and called from `ResUNet2.transformer` (:694-705) as `self.perceiver_io(image, queries_encoder=P_att)` with
P_att = [1, M, 256] (all active stride-8 voxels of the batch as ONE sequence) and image = [1, T, 128].

Same constructor signature, same parameter tree (a reference `state_dict` loads with strict=True), same
`forward(data, mask=None, queries_encoder=None)`.  The sub-modules are parameter containers only: the arithmetic runs
in the sm_100a CUDA library behind the C ABI (`gmf_dgr_head_*` in include/gmf_b200.h).  There is no PyTorch fallback.
"""
from __future__ import annotations

import ctypes as C
from typing import List, Optional, Tuple

import numpy as np
import torch
import torch.nn as nn

from . import _lib


def dgr_head_shapes(pe: bool = True, latent_dim: int = 256, dim: int = 128, cross_dim_head: int = 128):
    """state_dict key -> shape of the reference module (static mirror of the library's table; usable without the library)."""
    s = {}
    if pe:
        s["cpe.proj_q.weight"] = (latent_dim, 1, 3); s["cpe.proj_q.bias"] = (latent_dim,)
        s["cpe.proj_content.weight"] = (dim, 1, 3); s["cpe.proj_content.bias"] = (dim,)
    a, f = "cross_attend_blocks.0.", "cross_attend_blocks.1."
    s[a + "norm.weight"] = (latent_dim,); s[a + "norm.bias"] = (latent_dim,)
    s[a + "norm_context.weight"] = (dim,); s[a + "norm_context.bias"] = (dim,)
    s[a + "fn.to_q.weight"] = (cross_dim_head, latent_dim); s[a + "fn.to_kv.weight"] = (2 * cross_dim_head, dim)
    s[a + "fn.to_out.weight"] = (latent_dim, cross_dim_head); s[a + "fn.to_out.bias"] = (latent_dim,)
    s[f + "norm.weight"] = (latent_dim,); s[f + "norm.bias"] = (latent_dim,)
    s[f + "fn.net.0.weight"] = (latent_dim * 8, latent_dim); s[f + "fn.net.0.bias"] = (latent_dim * 8,)
    s[f + "fn.net.2.weight"] = (latent_dim, latent_dim * 4); s[f + "fn.net.2.bias"] = (latent_dim,)
    return s


def dgr_head_spec(pe: bool = True) -> List[Tuple[str, int]]:
    """(state_dict key, numel) in the order gmf_dgr_head_load_weights expects (queried from the library)."""
    lib = _lib.load()
    out = []
    for i in range(lib.gmf_dgr_head_weight_count(int(pe))):
        name = C.create_string_buffer(256)
        numel = C.c_int64()
        _lib.check(lib.gmf_dgr_head_weight_spec(int(pe), i, name, 256, C.byref(numel)))
        out.append((name.value.decode(), int(numel.value)))
    return out


def pack_dgr_state_dict(sd, pe: bool = True) -> np.ndarray:
    parts = []
    for key, numel in dgr_head_spec(pe):
        if key not in sd:
            raise KeyError(f"state_dict is missing {key}")
        t = sd[key]
        a = t.detach().cpu().numpy() if isinstance(t, torch.Tensor) else np.asarray(t)
        a = np.ascontiguousarray(a, dtype=np.float32).reshape(-1)
        if a.size != numel:
            raise ValueError(f"{key}: expected {numel} elements, got {a.size}")
        parts.append(a)
    return np.concatenate(parts)


class DgrHeadEngine:
    """Owns one `gmf_dgr_head` handle (packed weights + workspace) on one device."""

    def __init__(self, device: int = 0, pe: bool = True, latent_dim: int = 256, dim: int = 128, cross_dim_head: int = 128):
        if not torch.cuda.is_available():
            raise _lib.GmfError("gmf_b200 needs a CUDA device (sm_100a); there is no CPU fallback")
        self.lib = _lib.load()
        self.device = device
        self.pe = pe
        h = C.c_void_p()
        _lib.check(self.lib.gmf_dgr_head_create(C.byref(h), device, latent_dim, dim, cross_dim_head, int(pe)))
        self.h = h

    def load_state_dict(self, sd) -> None:
        flat = pack_dgr_state_dict(sd, self.pe)
        _lib.check(self.lib.gmf_dgr_head_load_weights(self.h, flat.ctypes.data_as(C.c_void_p), flat.size))

    def forward(self, latents: torch.Tensor, image_feat: torch.Tensor) -> torch.Tensor:
        """latents [M,256], image_feat [T,128] (cuda fp32) -> [M,256]"""
        assert latents.is_cuda and image_feat.is_cuda and latents.dtype == torch.float32 and image_feat.dtype == torch.float32
        x = latents.contiguous()
        ctx = image_feat.contiguous()
        out = torch.empty_like(x)
        st = torch.cuda.current_stream(x.device).cuda_stream
        _lib.check(self.lib.gmf_dgr_head_forward(self.h, x.data_ptr(), ctx.data_ptr(), x.shape[0], ctx.shape[0], out.data_ptr(),
                                                 C.c_void_p(st)))
        return out

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.gmf_dgr_head_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass


class DgrHeadTrainer:
    """One data-parallel training step of the head (BASELINE.json configs[4]; reference loop: core/trainer.py:226-300 with
    optim.SGD(lr, momentum, weight_decay), core/trainer.py:75-79).  Parameters, gradients and the momentum buffer are flat fp32 CUDA tensors in
    state_dict order; forward / backward / SGD run in the CUDA library, the gradient all-reduce is ONE torch.distributed call over the flat
    4.5 MB buffer (NCCL over NVLink when the process group is nccl; gloo in the CPU-side tests of the host logic).

        tr = DgrHeadTrainer(device, pe=True); tr.load_state_dict(sd)
        out = tr.forward(latents, image_feat)            # saves activations
        tr.backward(d_out)                               # fills tr.grads (and d_latents / d_image_feat)
        tr.step(lr=0.1, momentum=0.8, weight_decay=1e-4) # all-reduce (mean) + SGD
    """

    def __init__(self, device: int = 0, pe: bool = True):
        self.engine = DgrHeadEngine(device, pe)
        self.lib, self.pe, self.device = self.engine.lib, pe, torch.device("cuda", device)
        self.spec = dgr_head_spec(pe)
        n = int(self.lib.gmf_dgr_head_param_count(int(pe)))
        assert n == sum(k for _, k in self.spec)
        self.params = torch.zeros(n, device=self.device)
        self.grads = torch.zeros(n, device=self.device)
        self.momentum_buf = torch.zeros(n, device=self.device)
        self.steps = 0
        self._ws = None
        self._saved = None

    def load_state_dict(self, sd) -> None:
        self.params.copy_(torch.from_numpy(pack_dgr_state_dict(sd, self.pe)))
        self.steps = 0

    def state_dict(self, shapes=None):
        shapes = shapes or dgr_head_shapes(self.pe)
        out, o = {}, 0
        for key, numel in self.spec:
            out[key] = self.params[o:o + numel].detach().cpu().reshape(shapes[key]).clone()
            o += numel
        return out

    def grad_dict(self, shapes=None):
        shapes = shapes or dgr_head_shapes(self.pe)
        out, o = {}, 0
        for key, numel in self.spec:
            out[key] = self.grads[o:o + numel].detach().cpu().reshape(shapes[key]).clone()
            o += numel
        return out

    def _workspace(self, M, T):
        need = int(self.lib.gmf_dgr_head_train_workspace_bytes(M, T))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws

    def forward(self, latents: torch.Tensor, image_feat: torch.Tensor) -> torch.Tensor:
        x, ctx = latents.contiguous().float(), image_feat.contiguous().float()
        assert x.is_cuda and ctx.is_cuda and x.shape[1] == 256 and ctx.shape[1] == 128
        ws = self._workspace(x.shape[0], ctx.shape[0])
        out = torch.empty_like(x)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.gmf_dgr_head_train_forward(self.engine.h, self.params.data_ptr(), x.data_ptr(), ctx.data_ptr(), x.shape[0], ctx.shape[0],
                                                       out.data_ptr(), ws.data_ptr(), ws.numel(), st))
        self._saved = (x, ctx)
        return out

    def backward(self, d_out: torch.Tensor, want_input_grads: bool = True):
        if self._saved is None:
            raise _lib.GmfError("DgrHeadTrainer.backward needs a preceding forward (activations live in the workspace)")
        x, ctx = self._saved
        d_out = d_out.contiguous().float()
        d_x = torch.empty_like(x) if want_input_grads else None
        d_ctx = torch.empty_like(ctx) if want_input_grads else None
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.gmf_dgr_head_train_backward(self.engine.h, self.params.data_ptr(), x.data_ptr(), ctx.data_ptr(), d_out.data_ptr(),
                                                        x.shape[0], ctx.shape[0], d_x.data_ptr() if want_input_grads else None,
                                                        d_ctx.data_ptr() if want_input_grads else None, self.grads.data_ptr(),
                                                        self._ws.data_ptr(), self._ws.numel(), st))
        return d_x, d_ctx

    def step(self, lr: float = 0.1, momentum: float = 0.8, weight_decay: float = 1e-4, group=None) -> None:
        """all-reduce (sum) of the flat gradient over the process group, then SGD with the mean gradient"""
        from .shard import exchange_gradients
        world, _ = exchange_gradients(self.grads, group)
        st = C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)
        _lib.check(self.lib.gmf_sgd_step(self.params.data_ptr(), self.grads.data_ptr(), self.momentum_buf.data_ptr(), self.params.numel(), lr, momentum,
                                         weight_decay, 1.0 / world, 1 if self.steps == 0 else 0, st))
        self.steps += 1


# ---- parameter containers with the reference's key layout (perceiver_io.py) ----
class _Attention(nn.Module):                       # :68-81
    def __init__(self, query_dim, context_dim, heads, dim_head):
        super().__init__()
        inner = dim_head * heads
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, query_dim)


class _GEGLU(nn.Module):                           # :53-56
    pass


class _FeedForward(nn.Module):                     # :58-66
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), _GEGLU(), nn.Linear(dim * mult, dim))


class _PreNorm(nn.Module):                         # :31-37
    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if context_dim is not None else None


class _ConvPosEnc(nn.Module):                      # :105-124
    def __init__(self, dim_q, dim_content, k=3):
        super().__init__()
        self.proj_q = nn.Conv1d(dim_q, dim_q, k, 1, k // 2, groups=dim_q)
        self.proj_content = nn.Conv1d(dim_content, dim_content, k, 1, k // 2, groups=dim_content)


class PerceiverIO(nn.Module):
    """perceiver_io.py:140-221 with depth=0 (the only depth the reference instantiates)."""

    def __init__(self, depth, dim, latent_dim=512, cross_heads=1, latent_heads=8, cross_dim_head=64, latent_dim_head=64,
                 weight_tie_layers=False, pe=False):
        super().__init__()
        if depth != 0:
            raise NotImplementedError("the reference builds this head with depth=0 (resunet_new.py:518); latent self-attention is not built")
        if cross_heads != 1:
            raise NotImplementedError("cross_heads=1 only (resunet_new.py:520)")
        self.pe = pe
        self.dim, self.latent_dim, self.cross_dim_head = dim, latent_dim, cross_dim_head
        if pe:
            self.cpe = _ConvPosEnc(dim_q=latent_dim, dim_content=dim)
        self.cross_attend_blocks = nn.ModuleList([
            _PreNorm(latent_dim, _Attention(latent_dim, dim, heads=cross_heads, dim_head=cross_dim_head), context_dim=dim),
            _PreNorm(latent_dim, _FeedForward(latent_dim)),
        ])
        self.layers = nn.ModuleList([])
        self._engine: Optional[DgrHeadEngine] = None
        self._engine_version = -1

    def _sync_engine(self, device: torch.device) -> DgrHeadEngine:
        version = sum(p._version for p in self.parameters())
        if self._engine is None or self._engine.device != (device.index or 0):
            self._engine = DgrHeadEngine(device.index or 0, self.pe, self.latent_dim, self.dim, self.cross_dim_head)
            self._engine_version = -1
        if self._engine_version != version:
            self._engine.load_state_dict(self.state_dict())
            self._engine_version = version
        return self._engine

    @torch.no_grad()
    def forward(self, data, mask=None, queries_encoder=None):
        if mask is not None:
            raise NotImplementedError("mask is never passed by the reference (resunet_new.py:701)")
        x = queries_encoder
        if x.dim() != 3 or x.shape[0] != 1 or data.shape[0] != 1:
            raise ValueError("expected queries_encoder [1, M, latent_dim] and data [1, T, dim] (resunet_new.py:696-701)")
        eng = self._sync_engine(x.device)
        return eng.forward(x[0].float(), data[0].float()).unsqueeze(0)
