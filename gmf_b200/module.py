"""Drop-in mirror of the reference `PointDSC(nn.Module)` (GMF_PointDSC/models/PointDSC.py:146-266).

Same constructor signature, same parameter / buffer tree (so a reference `state_dict` loads with
strict=True), same `forward(data) -> {"final_trans", "final_labels", "M"}` contract.  The image backbone runs
in PyTorch (gmf_b200/backbone.py); everything after the image tokens runs in the sm_100a CUDA library behind
one custom op (`gmf_b200::pointdsc_forward`) over the C ABI in include/gmf_b200.h.  The modules below are
parameter containers only — there is no PyTorch implementation of the hot path to fall back to.
"""
from __future__ import annotations

from typing import Dict, List, Tuple

import torch
import torch.nn as nn

from .backbone import ImageEncoder
from .engine import Engine

_ENGINES: Dict[int, Engine] = {}


@torch.library.custom_op("gmf_b200::pointdsc_forward", mutates_args=())
def _pointdsc_forward(corr_pos: torch.Tensor, src: torch.Tensor, tgt: torch.Tensor, p_tok: torch.Tensor, q_tok: torch.Tensor,
                      engine_id: int, testing: bool) -> List[torch.Tensor]:
    out = _ENGINES[engine_id].forward(corr_pos, src, tgt, p_tok, q_tok, testing=testing, want_feat=not testing)
    feat = out["feat"] if out["feat"] is not None else torch.empty(0, device=corr_pos.device)
    return [out["final_trans"], out["final_labels"], out["confidence"], out["seeds"], feat]


@_pointdsc_forward.register_fake
def _(corr_pos, src, tgt, p_tok, q_tok, engine_id, testing):
    B, N = corr_pos.shape[0], corr_pos.shape[1]
    S = _ENGINES[engine_id].num_seeds(N)
    f = corr_pos.new_empty(B, N, 128) if not testing else corr_pos.new_empty(0)
    return [corr_pos.new_empty(B, 4, 4), corr_pos.new_empty(B, N), corr_pos.new_empty(B, N),
            corr_pos.new_empty(B, S, dtype=torch.int32), f]


# ------------------------------------------------------------------------------------------------
# parameter containers with the reference's key layout (models/fusion_layer.py, models/PointDSC.py)
# ------------------------------------------------------------------------------------------------
class _Attention(nn.Module):                       # fusion_layer.py:71-80
    def __init__(self, query_dim, context_dim, inner):
        super().__init__()
        self.to_q = nn.Linear(query_dim, inner, bias=False)
        self.to_kv = nn.Linear(context_dim, inner * 2, bias=False)
        self.to_out = nn.Linear(inner, context_dim)


class _GEGLU(nn.Module):
    pass


class _FeedForward(nn.Module):                     # fusion_layer.py:59-66
    def __init__(self, dim, mult=4):
        super().__init__()
        self.net = nn.Sequential(nn.Linear(dim, dim * mult * 2), _GEGLU(), nn.Linear(dim * mult, dim))


class _PreNorm(nn.Module):                         # fusion_layer.py:32-38
    def __init__(self, dim, fn, context_dim=None):
        super().__init__()
        self.fn = fn
        self.norm = nn.LayerNorm(dim)
        self.norm_context = nn.LayerNorm(context_dim) if context_dim is not None else None


class _ConvPosEnc(nn.Module):                      # fusion_layer.py:97-116
    def __init__(self, dim_q, dim_content):
        super().__init__()
        self.proj_q = nn.Conv1d(dim_q, dim_q, 3, 1, 1, groups=dim_q)
        self.proj_content = nn.Conv1d(dim_content, dim_content, 3, 1, 1, groups=dim_content)


class _FusionLayer(nn.Module):                     # fusion_layer.py:131-170 with depth=0
    def __init__(self, dim, latent_dim, dim_head, pe):
        super().__init__()
        self.pe = pe
        if pe:
            self.cpe = _ConvPosEnc(latent_dim, dim)
        self.cross_attend_blocks = nn.ModuleList([
            _PreNorm(latent_dim, _Attention(latent_dim, dim, dim_head), context_dim=dim),
            _PreNorm(latent_dim, _FeedForward(latent_dim)),
        ])
        self.layers = nn.ModuleList([])


class _NonLocalBlock(nn.Module):                   # PointDSC.py:10-38
    def __init__(self, c):
        super().__init__()
        self.fc_message = nn.Sequential(nn.Conv1d(c, c // 2, 1), nn.BatchNorm1d(c // 2), nn.ReLU(inplace=True),
                                        nn.Conv1d(c // 2, c // 2, 1), nn.BatchNorm1d(c // 2), nn.ReLU(inplace=True),
                                        nn.Conv1d(c // 2, c, 1))
        self.projection_q = nn.Conv1d(c, c, 1)
        self.projection_k = nn.Conv1d(c, c, 1)
        self.projection_v = nn.Conv1d(c, c, 1)
        self.fusion_layer_2 = _FusionLayer(c, c, c // 2, pe=True)


class _NonLocalNet(nn.Module):                     # PointDSC.py:77-112
    def __init__(self, in_dim, num_layers, c):
        super().__init__()
        self.num_layers = num_layers
        self.blocks = nn.ModuleDict()
        self.layer0 = nn.Conv1d(in_dim, c, 1, bias=True)
        self.image_encoder = ImageEncoder()
        self.fusion_layer_1 = _FusionLayer(c, c, c // 2, pe=False)
        for i in range(num_layers):
            self.blocks[f"PointCN_layer_{i}"] = nn.Sequential(nn.Conv1d(c, c, 1, bias=True), nn.BatchNorm1d(c), nn.ReLU(inplace=True))
            self.blocks[f"NonLocal_layer_{i}"] = _NonLocalBlock(c)


class PointDSC(nn.Module):
    """B200-native GMF-PointDSC.  `.eval()`: the inference path (eval-mode BatchNorm; `data` without the 'testing' key returns logits as
    final_labels and the feature-similarity matrix M, no gradients).  `.train()`: the training-mode forward of the reference (batch-statistics
    BatchNorm, running statistics updated, logits and M with autograd history), so that the reference's trainer - losses in Python,
    `loss.backward()`, any torch optimiser (libs/trainer.py:134-168) - runs unchanged; forward and backward are the CUDA training step of
    gmf_b200/trainer.py (`train_precision`: "tf32x3" or "tf32")."""

    train_precision = "tf32x3"

    def __init__(self, in_dim=6, num_layers=6, num_channels=128, num_iterations=10, ratio=0.1, inlier_threshold=0.10,
                 sigma_d=0.10, k=40, nms_radius=0.10):
        super().__init__()
        if in_dim != 6 or num_channels != 128:
            raise ValueError("the sm_100a kernels are specialised on in_dim=6, num_channels=128")
        self.num_iterations, self.ratio, self.num_channels = num_iterations, ratio, num_channels
        self.inlier_threshold, self.k, self.nms_radius, self.num_layers = inlier_threshold, k, nms_radius, num_layers
        self.sigma = nn.Parameter(torch.tensor([1.0]), requires_grad=True)
        self.sigma_spat = nn.Parameter(torch.tensor([float(sigma_d)]), requires_grad=False)
        self.encoder = _NonLocalNet(in_dim, num_layers, num_channels)
        self.classification = nn.Sequential(nn.Conv1d(num_channels, 32, 1), nn.ReLU(inplace=True), nn.Conv1d(32, 32, 1),
                                            nn.ReLU(inplace=True), nn.Conv1d(32, 1, 1))
        for m in self.modules():                   # PointDSC.py:183-188
            if isinstance(m, (nn.Conv1d, nn.Linear)):
                nn.init.xavier_normal_(m.weight, gain=1)
            elif isinstance(m, nn.BatchNorm1d):
                nn.init.constant_(m.weight, 1)
                nn.init.constant_(m.bias, 0)
        self._engine_id = -1
        self._packed_sig: Tuple = ()

    def __del__(self):
        _ENGINES.pop(getattr(self, "_engine_id", -1), None)

    # ---- engine management ------------------------------------------------------------------
    def _weights_signature(self) -> Tuple:
        return tuple((t.data_ptr(), t._version) for n, t in self.state_dict(keep_vars=True).items()
                     if not n.startswith("encoder.image_encoder."))

    def engine(self) -> Engine:
        dev = self.sigma.device
        if dev.type != "cuda":
            raise RuntimeError("gmf_b200.PointDSC runs on CUDA (sm_100a) only: call .cuda() first; there is no CPU path")
        eng = _ENGINES.get(self._engine_id)
        if eng is None or eng.device != dev:
            _ENGINES.pop(self._engine_id, None)        # the module moved: release the old device's workspace / weights / streams
            with torch.cuda.device(dev):
                eng = Engine(self.num_layers, self.num_iterations, self.k, self.ratio, self.inlier_threshold, self.nms_radius, dev)
            self._engine_id = max(_ENGINES.keys(), default=-1) + 1
            _ENGINES[self._engine_id] = eng
            self._packed_sig = ()
        sig = self._weights_signature()
        if sig != self._packed_sig:                # (re)pack after load_state_dict / parameter updates
            eng.load_state_dict(self.state_dict())
            self._packed_sig = sig
        return eng

    # "fp32": the reference-equivalent eager trunk (default).  "bf16" / "bf16_graph": opt-in channels_last + bf16 autocast trunk, the latter
    # replayed from a CUDA graph (backbone.py tokens_fast; 28 -> 20 ms per 2 x 64 images at 480 x 640).  The tokens then carry bf16 rounding
    # noise (~2e-2 abs on tokens of magnitude ~5), which costs ~3e-3 on the logits (tests/test_gpu_parity.py).
    backbone_mode = "fp32"

    @torch.no_grad()
    def image_tokens(self, p_image, q_image):
        enc = self.encoder.image_encoder
        if self.backbone_mode == "fp32":
            return enc.tokens(p_image), enc.tokens(q_image)
        if self.backbone_mode not in ("bf16", "bf16_graph"):
            raise ValueError(f"unknown backbone_mode {self.backbone_mode!r}")
        sig = tuple((t.data_ptr(), t._version) for t in enc.state_dict(keep_vars=True).values())
        if getattr(self, "_bb_sig", None) != sig:              # weights changed: drop captured graphs
            enc.invalidate_fast()
            self._bb_sig = sig
        g = self.backbone_mode == "bf16_graph"
        return enc.tokens_fast(p_image, g), enc.tokens_fast(q_image, g)

    def _forward_train(self, data):
        """PointDSC.forward in training mode (PointDSC.py:207-266 with self.training): `final_labels` = logits and `M` carry autograd history
        back to the hot-path parameters and, through the image tokens, to the PyTorch backbone; `final_trans` (top-`ratio` seeds without NMS,
        no post-refinement, :246,253) is computed without gradient, as in the reference where the hypothesis selection is not differentiable
        with respect to anything the default losses use (weight_transformation = 0, config_3DMatch.py:52)."""
        from .trainer import TrainState, train_forward
        if "testing" in data.keys():
            raise RuntimeError("'testing' data with a module in training mode: call .eval() first (the reference asserts bs == 1 there)")
        corr_pos, src, tgt = data["corr_pos"], data["src_keypts"], data["tgt_keypts"]
        eng = self.engine()
        st = getattr(self, "_train_state", None)
        if st is None or st.x3 != (1 if self.train_precision == "tf32x3" else 0):
            st = self._train_state = TrainState(self.num_layers, self.train_precision)
        with torch.cuda.device(corr_pos.device):
            enc = self.encoder.image_encoder
            p_tok, q_tok = enc.tokens(data["p_image"]), enc.tokens(data["q_image"])     # with autograd: the backbone trains too
            logits, M, feats = train_forward(st, corr_pos, src, tgt, p_tok, q_tok, self.state_dict(keep_vars=True))
            for m in self.modules():               # nn.BatchNorm1d bookkeeping of a training-mode forward (the hot-path BatchNorms ran in CUDA)
                if isinstance(m, nn.BatchNorm1d) and m.num_batches_tracked is not None:
                    m.num_batches_tracked += 1
            with torch.no_grad():
                normed = torch.nn.functional.normalize(feats, p=2, dim=-1)
                seeds = eng.pick_seeds(src, logits.detach(), use_nms=False)
                seed_trans = eng.seed_hypotheses(normed, src, tgt, seeds)[0]
                final_trans = eng.score_hypotheses(seed_trans, src, tgt, refine=False)[0]
        return {"final_trans": final_trans, "final_labels": logits, "M": M}

    def forward(self, data):
        if self.training:
            return self._forward_train(data)
        with torch.no_grad():
            return self._forward_eval(data)

    def _forward_eval(self, data):
        testing = "testing" in data.keys()
        corr_pos, src, tgt = data["corr_pos"], data["src_keypts"], data["tgt_keypts"]
        eng = self.engine()
        for name in ("corr_pos", "src_keypts", "tgt_keypts", "p_image", "q_image"):
            if data[name].device != eng.device:      # the C side dereferences raw pointers on the engine's device
                raise RuntimeError(f"data['{name}'] is on {data[name].device}, the model on {eng.device}")
        with torch.cuda.device(corr_pos.device):
            p_tok, q_tok = self.image_tokens(data["p_image"], data["q_image"])
            trans, labels, conf, _seeds, feat = _pointdsc_forward(corr_pos, src, tgt, p_tok, q_tok, self._engine_id, testing)
            if testing:
                return {"final_trans": trans, "final_labels": labels, "M": None}
            # training-mode outputs (PointDSC.py:231-234, 260): M from normalised features, logits as labels
            M = _ENGINES[self._engine_id].feature_compat(feat)
            return {"final_trans": trans, "final_labels": conf, "M": M}
