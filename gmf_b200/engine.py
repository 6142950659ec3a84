"""Host-side handle on the C-ABI context: weight upload, workspace management and per-stage calls.

All tensors handed to the engine must be CUDA fp32 contiguous (int32 for indices); torch is used only for
device memory and the current stream."""
from __future__ import annotations

import ctypes as C
from typing import Dict, Optional, Tuple

import torch

from . import _lib
from .weights import hot_path_spec, pack_state_dict


def _ptr(t: Optional[torch.Tensor]):
    return None if t is None else C.c_void_p(t.data_ptr())


def _chk(t: torch.Tensor, dtype=torch.float32) -> torch.Tensor:
    if not t.is_cuda:
        raise _lib.GmfError("gmf_b200 ops need CUDA tensors (no CPU fallback)")
    return t.detach().to(dtype).contiguous()


class Engine:
    def __init__(self, num_layers=12, num_iterations=10, k=40, ratio=0.1, inlier_threshold=0.10, nms_radius=0.10,
                 device: Optional[torch.device] = None):
        if not torch.cuda.is_available():
            raise _lib.GmfError("gmf_b200 needs a CUDA (sm_100a) device; there is no CPU fallback")
        self.lib = _lib.load()
        self.device = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
        self.cfg = _lib.GmfConfig(num_layers, num_iterations, k, ratio, inlier_threshold, nms_radius)
        self.num_layers, self.ratio, self.k = num_layers, ratio, k
        h = C.c_void_p()
        _lib.check(self.lib.gmf_create(C.byref(h), self.device.index or 0, C.byref(self.cfg)))
        self.h = h
        self._ws: Optional[torch.Tensor] = None
        self._verify_spec()

    def _verify_spec(self):
        spec = hot_path_spec(self.num_layers)
        n = self.lib.gmf_weight_count(self.num_layers)
        assert n == len(spec), (n, len(spec))
        buf = C.create_string_buffer(256)
        numel = C.c_int64()
        for i, (name, shape) in enumerate(spec.items()):
            _lib.check(self.lib.gmf_weight_spec(self.num_layers, i, buf, 256, C.byref(numel)))
            want = 1
            for d in shape:
                want *= d
            assert buf.value.decode() == name and numel.value == want, (i, buf.value, name)

    def __del__(self):
        try:
            if getattr(self, "h", None):
                self.lib.gmf_destroy(self.h)
                self.h = None
        except Exception:
            pass

    # ---- weights -------------------------------------------------------------------------------
    def load_state_dict(self, state_dict: Dict[str, torch.Tensor]):
        flat = pack_state_dict(state_dict, self.num_layers)
        _lib.check(self.lib.gmf_load_weights(self.h, C.c_void_p(flat.data_ptr()), flat.numel()))

    # ---- helpers --------------------------------------------------------------------------------
    def num_seeds(self, n: int) -> int:
        return int(n * float(self.ratio))               # PointDSC.py:246,270: int(num_corr * self.ratio), Python double

    def workspace(self, B: int, N: int, T: int) -> Tuple[torch.Tensor, int]:
        need = int(self.lib.gmf_workspace_bytes(self.h, B, N, max(T, 1)))
        if self._ws is None or self._ws.numel() < need:
            self._ws = None
            self._ws = torch.empty(need, dtype=torch.uint8, device=self.device)
        return self._ws, self._ws.numel()

    @staticmethod
    def _stream():
        return C.c_void_p(torch.cuda.current_stream().cuda_stream)

    # ---- whole path -----------------------------------------------------------------------------
    def forward(self, corr_pos, src, tgt, p_tok, q_tok, testing=True, want_feat=False):
        corr_pos, src, tgt, p_tok, q_tok = map(_chk, (corr_pos, src, tgt, p_tok, q_tok))
        B, N, _ = corr_pos.shape
        T = p_tok.shape[1]
        S = self.num_seeds(N)
        dev = corr_pos.device
        trans = torch.empty(B, 4, 4, device=dev)
        labels = torch.empty(B, N, device=dev)
        conf = torch.empty(B, N, device=dev)
        seeds = torch.empty(B, S, dtype=torch.int32, device=dev)
        feat = torch.empty(B, N, 128, device=dev) if want_feat else None
        ws, nbytes = self.workspace(B, N, T)
        _lib.check(self.lib.gmf_pointdsc_forward(self.h, _ptr(corr_pos), _ptr(src), _ptr(tgt), _ptr(p_tok), _ptr(q_tok), B, N, T,
                                                 1 if testing else 0, _ptr(trans), _ptr(labels), _ptr(conf), _ptr(seeds),
                                                 _ptr(feat), _ptr(ws), nbytes, self._stream()))
        return {"final_trans": trans, "final_labels": labels, "confidence": conf, "seeds": seeds, "feat": feat}

    def forward_host(self, corr_pos, src, tgt, p_tok, q_tok, out_trans, out_labels, out_conf=None, testing=True):
        """HOST tensors in, HOST tensors out (H2D + forward + D2H + stream sync inside the C call)."""
        B, N, _ = corr_pos.shape
        T = p_tok.shape[1]
        _lib.check(self.lib.gmf_pointdsc_forward_host(self.h, _ptr(corr_pos), _ptr(src), _ptr(tgt), _ptr(p_tok), _ptr(q_tok), B, N, T,
                                                      1 if testing else 0, _ptr(out_trans), _ptr(out_labels), _ptr(out_conf),
                                                      self._stream()))

    def forward_host_async(self, corr_pos, src, tgt, p_tok, q_tok, out_trans, out_labels, out_conf=None, testing=True):
        """Like forward_host but returns once the work is enqueued; call `synchronize()` before reading the outputs.  Consecutive
        calls overlap: the uploads of call k+1 run while call k computes (double-buffered staging inside the library)."""
        B, N, _ = corr_pos.shape
        T = p_tok.shape[1]
        _lib.check(self.lib.gmf_pointdsc_forward_host_async(self.h, _ptr(corr_pos), _ptr(src), _ptr(tgt), _ptr(p_tok), _ptr(q_tok), B, N, T,
                                                            1 if testing else 0, _ptr(out_trans), _ptr(out_labels), _ptr(out_conf),
                                                            self._stream()))

    def feature_compat(self, feat):
        """Training-mode `M` (PointDSC.py:231-234): feat [B,N,128] un-normalised encoder features -> [B,N,N]."""
        feat = _chk(feat)
        B, N, _ = feat.shape
        M = torch.empty(B, N, N, device=feat.device)
        nb = int(self.lib.gmf_feature_compat_workspace_bytes(B, N))
        ws = torch.empty(nb, dtype=torch.uint8, device=feat.device)
        _lib.check(self.lib.gmf_feature_compat(self.h, _ptr(feat), B, N, _ptr(M), _ptr(ws), nb, self._stream()))
        return M

    def weighted_procrustes(self, X, Y, w, eps=1e-6):
        """DGR weighted Procrustes (core/registration.py:91-113): X, Y [B,N,3], w [B,N] -> R [B,3,3], t [B,3]."""
        X, Y, w = _chk(X), _chk(Y), _chk(w)
        B, N, _ = X.shape
        R, t = torch.empty(B, 3, 3, device=X.device), torch.empty(B, 3, device=X.device)
        _lib.check(self.lib.gmf_weighted_procrustes(self.h, _ptr(X), _ptr(Y), _ptr(w), B, N, float(eps), _ptr(R), _ptr(t), self._stream()))
        return R, t

    def global_registration(self, X, Y, w, quantization_size=1.0, max_iter=1000, max_break_count=20, break_threshold_ratio=1e-5):
        """DGR `GlobalRegistration` (core/registration.py:135-194): weighted Procrustes + robust SE(3) Adam refinement, all on the device.
        X, Y [B,N,3], w [B,N] -> (R [B,3,3], t [B,3], info [B,3] = exit iteration, loss, break count)."""
        X, Y, w = _chk(X), _chk(Y), _chk(w)
        B, N, _ = X.shape
        R, t, info = torch.empty(B, 3, 3, device=X.device), torch.empty(B, 3, device=X.device), torch.empty(B, 3, device=X.device)
        _lib.check(self.lib.gmf_global_registration(self.h, _ptr(X), _ptr(Y), _ptr(w), B, N, float(quantization_size), int(max_iter),
                                                    int(max_break_count), float(break_threshold_ratio), _ptr(R), _ptr(t), _ptr(info), self._stream()))
        return R, t, info

    def sm_baseline(self, src, tgt, inlier_threshold=0.10, top_ratio=0.1, iters=10):
        """Classical spectral matching `SM` (baseline_scripts/baseline_3DMatch.py:19-53): src, tgt [B,N,3] ->
        (trans [B,4,4], labels [B,N], leading_eig [B,N])."""
        src, tgt = _chk(src), _chk(tgt)
        B, N, _ = src.shape
        trans, labels, eig = torch.empty(B, 4, 4, device=src.device), torch.empty(B, N, device=src.device), torch.empty(B, N, device=src.device)
        nb = int(self.lib.gmf_sm_workspace_bytes(B, N, float(top_ratio)))
        ws = torch.empty(max(nb, 1), dtype=torch.uint8, device=src.device)
        _lib.check(self.lib.gmf_sm_baseline(self.h, _ptr(src), _ptr(tgt), B, N, float(inlier_threshold), float(top_ratio), int(iters),
                                            _ptr(trans), _ptr(labels), _ptr(eig), _ptr(ws), nb, self._stream()))
        return trans, labels, eig

    def synchronize(self):
        _lib.check(self.lib.gmf_stream_synchronize(self.h, self._stream()))

    # ---- stages ---------------------------------------------------------------------------------
    def fusion_layer(self, layer: int, queries, context):
        queries, context = _chk(queries), _chk(context)
        B, Lq, _ = queries.shape
        Lk = context.shape[1]
        out = torch.empty_like(queries)
        ws, nb = self.workspace(B, max(Lq, 2), Lk)
        _lib.check(self.lib.gmf_fusion_layer(self.h, layer, _ptr(queries), _ptr(context), B, Lq, Lk, _ptr(out), _ptr(ws), nb, self._stream()))
        return out

    def sc_attention(self, layer: int, feat, src, tgt):
        feat, src, tgt = _chk(feat), _chk(src), _chk(tgt)
        B, N, _ = feat.shape
        out = torch.empty_like(feat)
        ws, nb = self.workspace(B, N, 1)
        _lib.check(self.lib.gmf_sc_attention(self.h, layer, _ptr(feat), _ptr(src), _ptr(tgt), B, N, _ptr(out), _ptr(ws), nb, self._stream()))
        return out

    def encoder_layer(self, layer: int, feat, src, tgt, image_feat):
        feat, src, tgt, image_feat = _chk(feat), _chk(src), _chk(tgt), _chk(image_feat)
        B, N, _ = feat.shape
        T = image_feat.shape[1]
        out = torch.empty_like(feat)
        ws, nb = self.workspace(B, N, T)
        _lib.check(self.lib.gmf_encoder_layer(self.h, layer, _ptr(feat), _ptr(src), _ptr(tgt), _ptr(image_feat), B, N, T, _ptr(out),
                                              _ptr(ws), nb, self._stream()))
        return out

    def classify(self, feat):
        feat = _chk(feat)
        B, N, _ = feat.shape
        normed, conf = torch.empty_like(feat), torch.empty(B, N, device=feat.device)
        _lib.check(self.lib.gmf_classify(self.h, _ptr(feat), B, N, _ptr(normed), _ptr(conf), self._stream()))
        return normed, conf

    def pick_seeds(self, src, confidence, use_nms=True):
        src, confidence = _chk(src), _chk(confidence)
        B, N, _ = src.shape
        seeds = torch.empty(B, self.num_seeds(N), dtype=torch.int32, device=src.device)
        ws, nb = self.workspace(B, N, 1)
        _lib.check(self.lib.gmf_pick_seeds(self.h, _ptr(src), _ptr(confidence), B, N, 1 if use_nms else 0, _ptr(seeds), _ptr(ws), nb,
                                           self._stream()))
        return seeds

    def seed_hypotheses(self, normed, src, tgt, seeds):
        normed, src, tgt = _chk(normed), _chk(src), _chk(tgt)
        seeds = _chk(seeds, torch.int32)
        B, N, _ = normed.shape
        S = seeds.shape[1]
        k = min(self.k, N - 1)
        trans = torch.empty(B, S, 4, 4, device=normed.device)
        knn = torch.empty(B, S, k, dtype=torch.int32, device=normed.device)
        w = torch.empty(B, S, k, device=normed.device)
        ws, nb = self.workspace(B, N, 1)
        _lib.check(self.lib.gmf_seed_hypotheses(self.h, _ptr(normed), _ptr(src), _ptr(tgt), _ptr(seeds), B, N, S, _ptr(trans), _ptr(knn),
                                                _ptr(w), _ptr(ws), nb, self._stream()))
        return trans, knn, w

    def score_hypotheses(self, seed_trans, src, tgt, refine=True):
        seed_trans, src, tgt = _chk(seed_trans), _chk(src), _chk(tgt)
        B, S = seed_trans.shape[:2]
        N = src.shape[1]
        dev = src.device
        final, labels = torch.empty(B, 4, 4, device=dev), torch.empty(B, N, device=dev)
        counts = torch.empty(B, S, dtype=torch.int32, device=dev)
        best = torch.empty(B, dtype=torch.int32, device=dev)
        pre = torch.empty(B, 4, 4, device=dev)
        ws, nb = self.workspace(B, N, 1)
        _lib.check(self.lib.gmf_score_hypotheses(self.h, _ptr(seed_trans), _ptr(src), _ptr(tgt), B, N, S, 1 if refine else 0, _ptr(final),
                                                 _ptr(labels), _ptr(counts), _ptr(best), _ptr(pre), _ptr(ws), nb, self._stream()))
        return final, labels, counts, best, pre

    def rigid_transform_3d(self, A, B, weights=None):
        A, B = _chk(A), _chk(B)
        weights = None if weights is None else _chk(weights)
        M, k, _ = A.shape
        out = torch.empty(M, 4, 4, device=A.device)
        _lib.check(self.lib.gmf_rigid_transform_3d(self.h, _ptr(A), _ptr(B), _ptr(weights), M, k, _ptr(out), self._stream()))
        return out

    # ---- debug ----------------------------------------------------------------------------------
    def debug_linear(self, x, w, bias, residual=None, relu=True):
        x = _chk(x)
        w, bias = w.detach().float().cpu().contiguous(), bias.detach().float().cpu().contiguous()
        residual = None if residual is None else _chk(residual)
        rows, k = x.shape
        nout = w.shape[0]
        out = torch.empty(rows, nout, device=x.device)
        _lib.check(self.lib.gmf_debug_linear(self.h, _ptr(x), _ptr(w), _ptr(bias), _ptr(residual), rows, k, nout, 1 if relu else 0,
                                             _ptr(out), self._stream()))
        return out

    def debug_attention(self, q, k, v, scale, src=None, tgt=None, sigma_d=0.1):
        q, k, v = _chk(q), _chk(k), _chk(v)
        src = None if src is None else _chk(src)
        tgt = None if tgt is None else _chk(tgt)
        B, Lq, D = q.shape
        Lk = k.shape[1]
        out = torch.empty_like(q)
        ws, nb = self.workspace(B, max(Lq, Lk), max(Lq, Lk))
        _lib.check(self.lib.gmf_debug_attention(self.h, _ptr(q), _ptr(k), _ptr(v), _ptr(src), _ptr(tgt), B, Lq, Lk, D, scale, sigma_d,
                                                _ptr(out), _ptr(ws), nb, self._stream()))
        return out

    PROFILE_CATEGORIES = ["pcn_qkv", "fusion_q_proj", "fusion_kv_proj", "ffn_geglu", "attn_fusion", "attn_sc", "prep_layer0", "classify",
                          "pick_seeds", "seed_knn", "spectral_kabsch", "score_refine", "other"]

    def profile(self, enable: bool):
        _lib.check(self.lib.gmf_profile_enable(self.h, 1 if enable else 0))

    def profile_read(self):
        """-> {category: (total_ms, launches)} since profile(True)."""
        out = {}
        for i, name in enumerate(self.PROFILE_CATEGORIES):
            ms, n = C.c_double(), C.c_int64()
            _lib.check(self.lib.gmf_profile_read(self.h, i, C.byref(ms), C.byref(n)))
            out[name] = (ms.value, n.value)
        return out

    def launch_count(self, reset=False) -> int:
        return int(self.lib.gmf_launch_count(1 if reset else 0))
