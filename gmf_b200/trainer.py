"""Training step of GMF-PointDSC on the CUDA path.

Mirrors one iteration of the reference's `Trainer.train_epoch` (GMF_PointDSC/libs/trainer.py:123-168): training-mode forward of the path
after the image backbone, ClassificationLoss + SpectralMatchingLoss (libs/loss.py:66-139), analytic backward, the finite-gradient guard
(:161-166) and `torch.optim.Adam` (train_3DMatch.py:52-58).  Parameters, gradients and the Adam moments are flat fp32 CUDA tensors in
`hot_path_spec` order, so the data-parallel gradient exchange is ONE `torch.distributed.all_reduce` (NCCL) of `self.grads`.  The image
backbone stays in PyTorch: `forward_backward` returns the gradients with respect to the image tokens for autograd to carry on.
No CPU fallback: everything here calls the C ABI (`gmf_pointdsc_train_*`, `gmf_adam_step`).
"""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional

import torch

from . import _lib
from .weights import hot_path_spec, pack_state_dict


def trainable_mask(num_layers: int) -> torch.Tensor:
    """uint8 mask over the flat parameter buffer: 0 for BatchNorm running statistics (buffers, not parameters) and `sigma_spat`
    (requires_grad=False, PointDSC.py:165)."""
    parts = []
    for name, shape in hot_path_spec(num_layers).items():
        n = 1
        for d in shape:
            n *= d
        frozen = name.endswith("running_mean") or name.endswith("running_var") or name == "sigma_spat"
        parts.append(torch.full((n,), 0 if frozen else 1, dtype=torch.uint8))
    return torch.cat(parts)


class PointDSCTrainer:
    def __init__(self, num_layers: int = 12, device: int = 0, balanced: bool = False, weight_classification: float = 1.0,
                 weight_spectralmatching: float = 1.0, precision: str = "tf32x3"):
        """precision: "tf32x3" = error-compensated tensor-pipe products (fp32-level gradients), "tf32" = plain TF32 (3x less tensor work)."""
        if precision not in ("tf32", "tf32x3"):
            raise ValueError("precision must be 'tf32' or 'tf32x3'")
        self.x3 = 1 if precision == "tf32x3" else 0
        self.lib = _lib.load()
        self.num_layers, self.dev_index = int(num_layers), int(device)
        self.device = torch.device("cuda", self.dev_index)
        self.balanced, self.w_class, self.w_sm = bool(balanced), float(weight_classification), float(weight_spectralmatching)
        self.spec = hot_path_spec(self.num_layers)
        n = int(self.lib.gmf_pointdsc_param_count(self.num_layers))
        assert n == sum(int(torch.Size(s).numel()) for s in self.spec.values())
        self.params = torch.zeros(n, dtype=torch.float32, device=self.device)
        self.grads = torch.zeros_like(self.params)
        self.exp_avg = torch.zeros_like(self.params)
        self.exp_avg_sq = torch.zeros_like(self.params)
        self.mask = trainable_mask(self.num_layers).to(self.device)
        self.steps = 0
        self._ws: Optional[torch.Tensor] = None
        self._saved = None

    # ---- parameters ----
    def load_state_dict(self, sd: Dict[str, torch.Tensor]) -> None:
        self.params.copy_(pack_state_dict(sd, self.num_layers))

    def _unflatten(self, flat: torch.Tensor) -> Dict[str, torch.Tensor]:
        out, o = {}, 0
        host = flat.detach().cpu()
        for name, shape in self.spec.items():
            n = int(torch.Size(shape).numel())
            out[name] = host[o:o + n].reshape(shape).clone()
            o += n
        return out

    def state_dict(self) -> Dict[str, torch.Tensor]:
        return self._unflatten(self.params)

    def grad_dict(self) -> Dict[str, torch.Tensor]:
        return self._unflatten(self.grads)

    # ---- one step ----
    def _stream(self):
        return C.c_void_p(torch.cuda.current_stream(self.device).cuda_stream)

    def forward_backward(self, corr_pos, src_keypts, tgt_keypts, p_tokens, q_tokens, gt_labels, want_token_grads: bool = True) -> dict:
        """corr_pos [B,N,6], src/tgt_keypts [B,N,3], p/q_tokens [B,T,128] (backbone output), gt_labels [B,N] -> losses (device tensor of
        {class_loss, sm_loss, loss}), logits [B,N] (`final_labels` of the training-mode forward), features [B,N,128], d_p_tokens / d_q_tokens."""
        t = [x.to(self.device, torch.float32).contiguous() for x in (corr_pos, src_keypts, tgt_keypts, p_tokens, q_tokens, gt_labels)]
        cp, sk, tk, pt, qt, gt = t
        B, N, T = cp.shape[0], cp.shape[1], pt.shape[1]
        assert cp.shape == (B, N, 6) and sk.shape == (B, N, 3) and tk.shape == (B, N, 3) and pt.shape == (B, T, 128) and qt.shape == (B, T, 128) and gt.shape == (B, N)
        need = int(self.lib.gmf_pointdsc_train_workspace_bytes(self.num_layers, B, N, T, self.x3))
        if need == 0:
            raise _lib.GmfError("gmf_pointdsc_train_workspace_bytes: unsupported shape")
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        losses = torch.empty(3, dtype=torch.float32, device=self.device)
        logits = torch.empty(B, N, dtype=torch.float32, device=self.device)
        feats = torch.empty(B, N, 128, dtype=torch.float32, device=self.device)
        st = self._stream()
        _lib.check(self.lib.gmf_pointdsc_train_forward(self.dev_index, self.num_layers, self.params.data_ptr(), cp.data_ptr(), sk.data_ptr(), tk.data_ptr(),
                                                       pt.data_ptr(), qt.data_ptr(), gt.data_ptr(), B, N, T, int(self.balanced), self.w_class, self.w_sm,
                                                       self.x3, losses.data_ptr(), logits.data_ptr(), feats.data_ptr(), None, self._ws.data_ptr(), self._ws.numel(),
                                                       st))
        d_p = torch.empty_like(pt) if want_token_grads else None
        d_q = torch.empty_like(qt) if want_token_grads else None
        _lib.check(self.lib.gmf_pointdsc_train_backward(self.dev_index, self.num_layers, self.params.data_ptr(), cp.data_ptr(), pt.data_ptr(), qt.data_ptr(),
                                                        B, N, T, self.w_class, self.w_sm, self.x3, None, None, None, self.grads.data_ptr(),
                                                        d_p.data_ptr() if want_token_grads else None, d_q.data_ptr() if want_token_grads else None,
                                                        self._ws.data_ptr(), self._ws.numel(), st))
        return {"losses": losses, "class_loss": losses[0], "sm_loss": losses[1], "loss": losses[2], "final_labels": logits, "features": feats,
                "d_p_tokens": d_p, "d_q_tokens": d_q}

    def final_trans(self, out: dict, src_keypts, tgt_keypts, num_iterations: int = 10, ratio: float = 0.1, inlier_threshold: float = 0.10, k: int = 40):
        """`final_trans` of the training-mode forward (PointDSC.py:246-253: top-`ratio` seeds by confidence without NMS, per-seed spectral matching
        + weighted Kabsch, best hypothesis by inlier count, no post-refinement) from the `features` / `final_labels` of `forward_backward`, on the
        seed kernels of the inference engine - what the reference's TransformationLoss (libs/loss.py:12-63; logging only at weight 0) consumes."""
        from .engine import Engine
        eng = getattr(self, "_engine", None)
        cfg = (num_iterations, ratio, inlier_threshold, k)
        if eng is None or self._engine_cfg != cfg:
            with torch.cuda.device(self.device):
                eng = self._engine = Engine(self.num_layers, num_iterations, k, ratio, inlier_threshold, inlier_threshold, self.device)
            self._engine_cfg = cfg
        eng.load_state_dict(self.state_dict())                       # sigma / sigma_spat of the current step
        src, tgt = src_keypts.to(self.device, torch.float32).contiguous(), tgt_keypts.to(self.device, torch.float32).contiguous()
        with torch.no_grad(), torch.cuda.device(self.device):
            normed = torch.nn.functional.normalize(out["features"], p=2, dim=-1)
            seeds = eng.pick_seeds(src, out["final_labels"], use_nms=False)
            seed_trans = eng.seed_hypotheses(normed, src, tgt, seeds)[0]
            return eng.score_hypotheses(seed_trans, src, tgt, refine=False)[0]

    def step(self, lr: float = 1e-4, weight_decay: float = 1e-6, betas=(0.9, 0.999), eps: float = 1e-8, group=None) -> bool:
        """all-reduce (sum) of the flat gradient over the process group, the reference's finite-gradient guard, then Adam on the mean
        gradient.  Returns False when the step was skipped because of a non-finite gradient (trainer.py:161-168)."""
        from .shard import exchange_gradients
        world, ok = exchange_gradients(self.grads, group, check_finite=True)
        if not ok:
            return False
        self.steps += 1
        _lib.check(self.lib.gmf_adam_step(self.params.data_ptr(), self.grads.data_ptr(), self.exp_avg.data_ptr(), self.exp_avg_sq.data_ptr(),
                                          self.mask.data_ptr(), self.params.numel(), lr, betas[0], betas[1], eps, weight_decay, 1.0 / world, self.steps,
                                          self._stream()))
        return True


class _TrainForward(torch.autograd.Function):
    """Autograd node of the training-mode path: (image tokens, hot-path parameters) -> (logits, M, features).  Used by
    `gmf_b200.PointDSC.forward` in training mode, so that the reference's own trainer (losses in Python on `final_labels` / `M`,
    `loss.backward()`, any torch optimiser; libs/trainer.py:134-168) runs unchanged on the CUDA path."""

    @staticmethod
    def forward(ctx, st: "TrainState", corr_pos, src, tgt, p_tok, q_tok, *params):
        lib, dev = st.lib, corr_pos.device
        B, N, T = corr_pos.shape[0], corr_pos.shape[1], p_tok.shape[1]
        flat = torch.cat([p.detach().reshape(-1).float() for p in params])
        need = int(lib.gmf_pointdsc_train_workspace_bytes(st.num_layers, B, N, T, st.x3))
        if need == 0:
            raise _lib.GmfError("gmf_pointdsc_train_workspace_bytes: unsupported shape")
        if st.ws is None or st.ws.numel() < need or st.ws.device != dev:
            st.ws = None
            st.ws = torch.empty(need, dtype=torch.uint8, device=dev)
        cp, sk, tk, pt, qt = (x.detach().float().contiguous() for x in (corr_pos, src, tgt, p_tok, q_tok))
        logits = torch.empty(B, N, device=dev)
        feats = torch.empty(B, N, 128, device=dev)
        M = torch.empty(B, N, N, device=dev)
        s_ = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(lib.gmf_pointdsc_train_forward(dev.index or 0, st.num_layers, flat.data_ptr(), cp.data_ptr(), sk.data_ptr(), tk.data_ptr(), pt.data_ptr(),
                                                  qt.data_ptr(), None, B, N, T, 0, 1.0, 1.0, st.x3, None, logits.data_ptr(), feats.data_ptr(), M.data_ptr(),
                                                  st.ws.data_ptr(), st.ws.numel(), s_))
        # BatchNorm running statistics were updated inside `flat`: write them back into the module's buffers
        o = 0
        with torch.no_grad():
            for p in params:
                n = p.numel()
                if getattr(p, "_gmf_running_stat", False):
                    p.copy_(flat[o:o + n].reshape(p.shape))
                o += n
        st.generation += 1                                    # the saved activations live in st.ws: a later forward overwrites them
        ctx.generation = st.generation
        ctx.st, ctx.flat, ctx.inputs, ctx.shapes = st, flat, (cp, pt, qt), [p.shape for p in params]
        ctx.dims = (B, N, T)
        return logits, M, feats

    @staticmethod
    def backward(ctx, d_logits, d_M, d_feats):
        st, (cp, pt, qt), (B, N, T) = ctx.st, ctx.inputs, ctx.dims
        if ctx.generation != st.generation:
            raise _lib.GmfError("gmf_b200 training-mode backward: the activations of this forward were overwritten by a later training-mode forward of "
                                "the same module (one workspace per module) - call backward() before the next forward")
        dev = cp.device
        grads = torch.empty_like(ctx.flat)
        d_p, d_q = torch.empty_like(pt), torch.empty_like(qt)
        ptr = lambda t: None if t is None else t.contiguous().float().data_ptr()
        keep = [None if t is None else t.contiguous().float() for t in (d_logits, d_M, d_feats)]
        if all(t is None for t in keep):
            keep[0] = torch.zeros(B, N, device=dev)
        s_ = C.c_void_p(torch.cuda.current_stream(dev).cuda_stream)
        _lib.check(st.lib.gmf_pointdsc_train_backward(dev.index or 0, st.num_layers, ctx.flat.data_ptr(), cp.data_ptr(), pt.data_ptr(), qt.data_ptr(), B, N, T,
                                                      1.0, 1.0, st.x3, *(None if t is None else t.data_ptr() for t in keep), grads.data_ptr(), d_p.data_ptr(),
                                                      d_q.data_ptr(), st.ws.data_ptr(), st.ws.numel(), s_))
        out, o = [], 0
        for shp in ctx.shapes:
            n = int(torch.Size(shp).numel())
            out.append(grads[o:o + n].reshape(shp))
            o += n
        return (None, None, None, None, d_p, d_q, *out)


class TrainState:
    """per-module state of the training-mode path (library handle, workspace, precision)"""

    def __init__(self, num_layers: int, precision: str = "tf32x3"):
        if precision not in ("tf32", "tf32x3"):
            raise ValueError("precision must be 'tf32' or 'tf32x3'")
        self.lib = _lib.load()
        self.num_layers, self.x3, self.ws, self.generation = int(num_layers), 1 if precision == "tf32x3" else 0, None, 0


def train_forward(st: TrainState, corr_pos, src, tgt, p_tok, q_tok, named_tensors: Dict[str, torch.Tensor]):
    """named_tensors: the module's state (parameters and buffers, keep_vars=True).  Returns logits, M, features with autograd history."""
    params = []
    for name in hot_path_spec(st.num_layers):
        t = named_tensors[name]
        t._gmf_running_stat = name.endswith("running_mean") or name.endswith("running_var")
        params.append(t)
    return _TrainForward.apply(st, corr_pos, src, tgt, p_tok, q_tok, *params)
