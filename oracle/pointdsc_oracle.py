"""TEST INFRASTRUCTURE ONLY — CPU restatement ("oracle") of the GMF-PointDSC forward hot path.

Only `tests/`, `__graft_entry__.smoke()` and `bench.py`'s cpu_baseline / `--impl reference` legs may
import this file; the product package `gmf_b200` never does (its ops fail loudly without the CUDA
library).  This is a *functional* fp32 torch-CPU restatement of the reference algorithm, written
against a flat `state_dict`, so that it can travel to the GPU box where `/root/reference` does not
exist.  Every function cites the reference lines it follows (paths relative to
`/root/reference/GMF_PointDSC/`).

Parity pinning: the reference ships no golden vectors or unit tests for this path, so
the oracle is pinned against *outputs of the reference itself executed in the build container*:
`oracle/gen_golden.py` imports the unmodified reference (oracle/ref_shim.py), runs it on seeded
synthetic inputs and commits the outputs under `tests/golden/`; `tests/test_oracle_golden.py`
checks this restatement against those fixtures (and against the live reference when present).
"""
from __future__ import annotations

from typing import Dict, Optional

import torch
import torch.nn.functional as F

Tensor = torch.Tensor


# --------------------------------------------------------------------------------------------
# fusion layer (models/fusion_layer.py)
# --------------------------------------------------------------------------------------------
def conv_pos_enc(x: Tensor, w: Tensor, b: Tensor) -> Tensor:
    """ConvPosEnc.forward, fusion_layer.py:118-128: depthwise Conv1d(k=3,pad=1) along the token
    axis plus identity.  x: [B, L, C]; w: [C,1,3]; b: [C]."""
    xt = x.transpose(1, 2)
    return (F.conv1d(xt, w, b, padding=1, groups=w.shape[0]) + xt).transpose(1, 2)


def cross_attention(sd, p: str, x: Tensor, ctx: Tensor) -> Tensor:
    """PreNorm(Attention) with one head, fusion_layer.py:32-52 and :71-94."""
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    cn = F.layer_norm(ctx, (ctx.shape[-1],), sd[p + "norm_context.weight"], sd[p + "norm_context.bias"], 1e-5)
    q = xn @ sd[p + "fn.to_q.weight"].T                       # :84 (no bias)
    kv = cn @ sd[p + "fn.to_kv.weight"].T                     # :86
    d = q.shape[-1]
    k, v = kv[..., :d], kv[..., d:]                           # chunk(2, -1) :87
    sim = torch.einsum("bid,bjd->bij", q, k) * d ** -0.5      # :90
    out = torch.einsum("bij,bjd->bid", sim.softmax(dim=-1), v)  # :91-92
    return out @ sd[p + "fn.to_out.weight"].T + sd[p + "fn.to_out.bias"]  # :94


def geglu_ff(sd, p: str, x: Tensor) -> Tensor:
    """PreNorm(FeedForward): Linear -> GEGLU (erf GELU) -> Linear, fusion_layer.py:54-69."""
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    h = xn @ sd[p + "fn.net.0.weight"].T + sd[p + "fn.net.0.bias"]
    half = h.shape[-1] // 2
    h = h[..., :half] * F.gelu(h[..., half:])                 # value * gelu(gate) :55-57
    return h @ sd[p + "fn.net.2.weight"].T + sd[p + "fn.net.2.bias"]


def fusion_layer(sd, p: str, ctx: Tensor, queries: Tensor, pe: bool) -> Tensor:
    """FusionLayer.forward with depth=0, fusion_layer.py:172-201.  ctx = `data`, queries =
    `queries_encoder`; both [B, L, 128]."""
    x = queries
    if pe:                                                    # :182-186
        x = conv_pos_enc(x, sd[p + "cpe.proj_q.weight"], sd[p + "cpe.proj_q.bias"])
        ctx = conv_pos_enc(ctx, sd[p + "cpe.proj_content.weight"], sd[p + "cpe.proj_content.bias"])
    x = cross_attention(sd, p + "cross_attend_blocks.0.", x, ctx) + x   # :190
    x = geglu_ff(sd, p + "cross_attend_blocks.1.", x) + x               # :191
    return x


# --------------------------------------------------------------------------------------------
# PointDSC encoder (models/PointDSC.py:10-143)
# --------------------------------------------------------------------------------------------
def compat_matrix(src: Tensor, tgt: Tensor, sigma_spat: Tensor) -> Dict[str, Tensor]:
    """Length-consistency matrix, PointDSC.py:216-221."""
    sd_ = torch.norm(src[:, :, None, :] - src[:, None, :, :], dim=-1)
    td_ = torch.norm(tgt[:, :, None, :] - tgt[:, None, :, :], dim=-1)
    c = torch.clamp(1.0 - (sd_ - td_) ** 2 / sigma_spat ** 2, min=0)
    return {"src_dist": sd_, "compat": c}


def _bn_eval(x: Tensor, sd, p: str) -> Tensor:
    """Eval-mode BatchNorm1d on [B, C, N] (eps 1e-5)."""
    return F.batch_norm(x, sd[p + "running_mean"], sd[p + "running_var"], sd[p + "weight"], sd[p + "bias"],
                        False, 0.0, 1e-5)


def sc_nonlocal_attention(sd, p: str, feat: Tensor, compat: Tensor) -> Tensor:
    """Q/K/V 1x1 convs + softmax(compat * QK^T/sqrt(C)) V, PointDSC.py:54-64.  feat: [B,C,N]."""
    q = F.conv1d(feat, sd[p + "projection_q.weight"], sd[p + "projection_q.bias"])
    k = F.conv1d(feat, sd[p + "projection_k.weight"], sd[p + "projection_k.bias"])
    v = F.conv1d(feat, sd[p + "projection_v.weight"], sd[p + "projection_v.bias"])
    logits = torch.einsum("bco,bci->boi", q, k) / feat.shape[1] ** 0.5        # :60
    w = torch.softmax(compat * logits, dim=-1)                                # :62 (multiplicative!)
    return torch.einsum("boi,bci->bco", w, v)                                  # :64


def fc_message(sd, p: str, msg: Tensor) -> Tensor:
    """conv-BN-ReLU, conv-BN-ReLU, conv; PointDSC.py:13-21,65."""
    x = torch.relu(_bn_eval(F.conv1d(msg, sd[p + "0.weight"], sd[p + "0.bias"]), sd, p + "1."))
    x = torch.relu(_bn_eval(F.conv1d(x, sd[p + "3.weight"], sd[p + "3.bias"]), sd, p + "4."))
    return F.conv1d(x, sd[p + "6.weight"], sd[p + "6.bias"])


def nonlocal_block(sd, p: str, feat: Tensor, compat: Tensor, image_feat: Tensor, cap=None) -> Tensor:
    """NonLocalBlock.forward, PointDSC.py:40-74."""
    msg = sc_nonlocal_attention(sd, p, feat, compat)
    message = fc_message(sd, p + "fc_message.", msg)
    fused = fusion_layer(sd, p + "fusion_layer_2.", image_feat, feat.permute(0, 2, 1), pe=True).permute(0, 2, 1)
    if cap is not None:
        cap["sc_msg"] = msg.permute(0, 2, 1).contiguous()
        cap["fusion2"] = fused.permute(0, 2, 1).contiguous()
    return message + fused                                                     # :73


def encoder(sd, num_layers: int, corr_pos: Tensor, compat: Tensor, p_tok: Tensor, q_tok: Tensor,
            cap: Optional[dict] = None) -> Tensor:
    """NonLocalNet.forward after the image backbone, PointDSC.py:137-143.  Tokens are [B,T,128].
    Fusion-1: queries = q-image tokens, context = p-image tokens (:137)."""
    image_feat = fusion_layer(sd, "encoder.fusion_layer_1.", p_tok, q_tok, pe=False)
    feat = F.conv1d(corr_pos.permute(0, 2, 1), sd["encoder.layer0.weight"], sd["encoder.layer0.bias"])
    if cap is not None:
        cap["image_feat"] = image_feat
        cap["feat_in"] = []
        cap["feat_out"] = []
    for i in range(num_layers):
        p = f"encoder.blocks.PointCN_layer_{i}."
        feat = torch.relu(_bn_eval(F.conv1d(feat, sd[p + "0.weight"], sd[p + "0.bias"]), sd, p + "1."))
        lc = {} if cap is not None else None
        out = nonlocal_block(sd, f"encoder.blocks.NonLocal_layer_{i}.", feat, compat, image_feat, lc)
        if cap is not None:
            cap["feat_in"].append(feat.permute(0, 2, 1).contiguous())
            cap["feat_out"].append(out.permute(0, 2, 1).contiguous())
            cap.setdefault("layers", []).append(lc)
        feat = out
    return feat.permute(0, 2, 1)                                               # [B, N, C]


def classify(sd, feat: Tensor) -> Tensor:
    """classification MLP 128-32-32-1 on un-normalised features, PointDSC.py:175-181,241."""
    x = feat.permute(0, 2, 1)
    x = torch.relu(F.conv1d(x, sd["classification.0.weight"], sd["classification.0.bias"]))
    x = torch.relu(F.conv1d(x, sd["classification.2.weight"], sd["classification.2.bias"]))
    return F.conv1d(x, sd["classification.4.weight"], sd["classification.4.bias"]).squeeze(1)


# --------------------------------------------------------------------------------------------
# seed phase (PointDSC.py:268-528, models/common.py:10-75)
# --------------------------------------------------------------------------------------------
def pick_seeds(src_dist: Tensor, scores: Tensor, radius: float, max_num: int) -> Tensor:
    """NMS seed selection (bs == 1), PointDSC.py:268-286.  Literal: scores * is_local_max sorted
    descending — with negative logits the suppressed points (score * 0) rank first."""
    assert scores.shape[0] == 1
    rel = (scores.T >= scores) | (src_dist[0] >= radius)
    is_max = rel.min(-1)[0].float()
    return torch.argsort(scores * is_max, dim=1, descending=True)[:, :max_num]


def knn_indices(x: Tensor, k: int) -> Tensor:
    """common.py:53-75 with normalized=True, ignore_self=True: top-(k+1) smallest of 2 - 2 x x^T,
    rank 0 dropped."""
    d = 2 - 2 * torch.matmul(x, x.transpose(2, 1))
    return d.topk(k=k + 1, dim=-1, largest=False)[1][:, :, 1:]


def leading_eigenvector(m: Tensor, iters: int) -> Tensor:
    """Power iteration with the *global* allclose early exit, PointDSC.py:429-448."""
    v = torch.ones_like(m[:, :, 0:1])
    last = v
    for _ in range(iters):
        v = torch.bmm(m, v)
        v = v / (torch.norm(v, dim=1, keepdim=True) + 1e-6)
        if torch.allclose(v, last):
            break
        last = v
    return v.squeeze(-1)


def rigid_transform_3d(a: Tensor, b: Tensor, w: Optional[Tensor] = None) -> Tensor:
    """Weighted Kabsch, common.py:10-50.  a,b: [M,k,3], w: [M,k] -> [M,4,4]."""
    if w is None:
        w = torch.ones_like(a[:, :, 0])
    ws = w.sum(dim=1, keepdim=True)[:, :, None] + 1e-6
    ca = (a * w[:, :, None]).sum(dim=1, keepdim=True) / ws
    cb = (b * w[:, :, None]).sum(dim=1, keepdim=True) / ws
    h = (a - ca).transpose(1, 2) @ (w[:, :, None] * (b - cb))      # A^T diag(w) B  (:34-35)
    u, _, v = torch.svd(h)                                         # :38 (V, not V^T, despite the name)
    e = torch.eye(3, dtype=a.dtype)[None].repeat(a.shape[0], 1, 1)
    e[:, 2, 2] = torch.det(v @ u.transpose(1, 2))
    r = v @ e @ u.transpose(1, 2)
    t = cb.transpose(1, 2) - r @ ca.transpose(1, 2)
    out = torch.eye(4, dtype=a.dtype)[None].repeat(a.shape[0], 1, 1)   # utils/SE3.py:73-96
    out[:, :3, :3] = r
    out[:, :3, 3:4] = t
    return out


def seed_hypotheses(feat_n: Tensor, src: Tensor, tgt: Tensor, seeds: Tensor, k: int, sigma: Tensor,
                    sigma_spat: Tensor, iters: int, cap: Optional[dict] = None) -> Tensor:
    """kNN -> 40x40 compat -> power iteration -> per-seed Kabsch; PointDSC.py:323-407."""
    bs, n, c = feat_n.shape
    k = min(k, n - 1)
    idx = knn_indices(feat_n, k).gather(1, seeds[:, :, None].expand(-1, -1, k))      # :327-329
    flat = idx.reshape(bs, -1)
    kf = feat_n.gather(1, flat[:, :, None].expand(-1, -1, c)).view(bs, -1, k, c)     # :335
    mf = torch.clamp(1 - (1 - kf @ kf.transpose(2, 3)) / sigma ** 2, min=0).view(-1, k, k)   # :337-339
    ks = src.gather(1, flat[:, :, None].expand(-1, -1, 3)).view(bs, -1, k, 3)
    kt = tgt.gather(1, flat[:, :, None].expand(-1, -1, 3)).view(bs, -1, k, 3)
    dd = ((ks[:, :, :, None] - ks[:, :, None]) ** 2).sum(-1) ** 0.5 - ((kt[:, :, :, None] - kt[:, :, None]) ** 2).sum(-1) ** 0.5
    ms = torch.clamp(1 - dd ** 2 / sigma_spat ** 2, min=0).view(-1, k, k)            # :349-352
    m = mf * ms
    ar = torch.arange(k)
    m[:, ar, ar] = 0                                                                 # :361
    w = leading_eigenvector(m, iters).view(bs, -1, k)
    w = w / (w.sum(dim=-1, keepdim=True) + 1e-6)                                     # :365
    trans = rigid_transform_3d(ks.view(-1, k, 3), kt.view(-1, k, 3), w.view(-1, k)).view(bs, -1, 4, 4)
    if cap is not None:
        cap.update(knn_idx=idx, seed_M=m, seed_weight=w, seed_trans=trans)
    return trans


def score_hypotheses(trans: Tensor, src: Tensor, tgt: Tensor, tau: float):
    """Inlier counting over all seeds x points, PointDSC.py:413-427."""
    pred = torch.einsum("bsnm,bmk->bsnk", trans[:, :, :3, :3], src.permute(0, 2, 1)) + trans[:, :, :3, 3:4]
    l2 = torch.norm(pred.permute(0, 1, 3, 2) - tgt[:, None], dim=-1)
    fit = (l2 < tau).float().mean(dim=-1)
    best = fit.argmax(dim=1)
    final = trans.gather(1, best[:, None, None, None].expand(-1, -1, 4, 4)).squeeze(1)
    lab = (l2.gather(1, best[:, None, None].expand(-1, -1, l2.shape[2])).squeeze(1) < tau).float()
    return fit, best, final, lab


def post_refinement(trans: Tensor, src: Tensor, tgt: Tensor, inlier_threshold: float) -> Tensor:
    """<=20 reweighted Kabsch rounds, stop when the inlier count repeats; PointDSC.py:493-528
    (bs == 1).  Threshold quirk kept: 0.10 if inlier_threshold == 0.10 else 1.2 (:505-508)."""
    assert trans.shape[0] == 1
    tau = 0.10 if inlier_threshold == 0.10 else 1.2
    prev = 0
    for _ in range(20):
        warped = (trans[:, :3, :3] @ src.permute(0, 2, 1) + trans[:, :3, 3:4]).permute(0, 2, 1)   # SE3.py:43-57
        l2 = torch.norm(warped - tgt, dim=-1)
        inl = (l2 < tau)[0]
        num = int(inl.sum())
        if abs(num - prev) < 1:
            break
        prev = num
        trans = rigid_transform_3d(src[:, inl, :], tgt[:, inl, :], (1 / (1 + (l2 / tau) ** 2))[:, inl])
    return trans


# --------------------------------------------------------------------------------------------
# whole forward (PointDSC.forward, PointDSC.py:191-266)
# --------------------------------------------------------------------------------------------
def tail_single(sd, cfg: dict, feat: Tensor, src: Tensor, tgt: Tensor, src_dist: Tensor, cap: Optional[dict] = None,
                confidence_override: Optional[Tensor] = None):
    """Everything after the encoder for ONE pair in testing mode (bs == 1 semantics)."""
    feat_n = F.normalize(feat, p=2, dim=-1)                                    # :229
    conf = classify(sd, feat) if confidence_override is None else confidence_override   # :241
    n = feat.shape[1]
    seeds = pick_seeds(src_dist, conf, cfg["nms_radius"], int(n * cfg["ratio"]))        # :244
    sc = {} if cap is not None else None
    trans = seed_hypotheses(feat_n, src, tgt, seeds, cfg["k"], sd["sigma"], sd["sigma_spat"],
                            cfg["num_iterations"], sc)
    fit, best, final, lab = score_hypotheses(trans, src, tgt, cfg["inlier_threshold"])
    refined = post_refinement(final, src, tgt, cfg["inlier_threshold"])                 # :257
    if cap is not None:
        cap.update(sc)
        cap.update(normed=feat_n, confidence=conf, seeds=seeds, fitness=fit, best=best, pre_refine=final)
    return refined, lab, conf


DEFAULT_CFG = dict(num_layers=12, num_iterations=10, ratio=0.1, inlier_threshold=0.10, sigma_d=0.10, k=40,
                   nms_radius=0.10)


@torch.no_grad()
def forward_testing(sd, cfg: dict, corr_pos: Tensor, src: Tensor, tgt: Tensor, p_tok: Tensor, q_tok: Tensor,
                    capture: bool = False) -> Dict[str, Tensor]:
    """Testing-mode forward on image *tokens* (the timed path starts after the backbone).
    B > 1 == the reference looped over pairs and stacked (reference asserts bs == 1, :279,:504)."""
    outs = {"final_trans": [], "final_labels": [], "confidence": [], "seeds": []}
    caps = []
    for b in range(corr_pos.shape[0]):
        cap = {} if capture else None
        sl = slice(b, b + 1)
        cm = compat_matrix(src[sl], tgt[sl], sd["sigma_spat"])
        feat = encoder(sd, cfg["num_layers"], corr_pos[sl], cm["compat"], p_tok[sl], q_tok[sl], cap)
        tcap = {} if capture else None
        tr, lab, conf = tail_single(sd, cfg, feat, src[sl], tgt[sl], cm["src_dist"], tcap)
        outs["final_trans"].append(tr), outs["final_labels"].append(lab), outs["confidence"].append(conf)
        if capture:
            cap.update(tcap)
            cap["feat"] = feat
            caps.append(cap)
    res = {k: torch.cat(v) for k, v in outs.items() if v}
    res["M"] = None
    if capture:
        res["capture"] = caps
    return res


def rotation_error_deg(r1: Tensor, r2: Tensor) -> Tensor:
    """fp64 geodesic angle via 2*asin(||R1-R2||_F / (2*sqrt 2)) (fp32 acos is too noisy)."""
    d = (r1.double() - r2.double()).flatten(-2).norm(dim=-1)
    return torch.rad2deg(2 * torch.asin(torch.clamp(d / (2 * 2 ** 0.5), max=1.0)))
