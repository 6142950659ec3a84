"""Precision attribution probe (CPU, test infrastructure): emulate the CUDA path's operand roundings inside the oracle, one op class at a
time, and report max|dlogit| against the fp32 oracle.  Usage: python tools/probe_precision.py [kitti|3dmatch] [N] [T] [layers]"""
import os
import sys

import torch
import torch.nn.functional as F

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmf_b200.synth import synth_pairs, synth_state_dict, synth_tokens   # noqa: E402
from gmf_b200.weights import hot_path_spec                               # noqa: E402
from oracle import pointdsc_oracle as O                                  # noqa: E402


def bf16(x):
    return x.to(torch.bfloat16).to(torch.float32)


def fp16(x):
    return x.clamp(-65504, 65504).to(torch.float16).to(torch.float32)


def tf32(x):
    i = x.contiguous().view(torch.int32)
    i = (i + 0xFFF + ((i >> 13) & 1)) & ~0x1FFF
    return i.view(torch.float32)


MODE = set()
STATS = {}


def sc_attn(sd, p, feat, compat):
    q = F.conv1d(feat, sd[p + "projection_q.weight"], sd[p + "projection_q.bias"])
    k = F.conv1d(feat, sd[p + "projection_k.weight"], sd[p + "projection_k.bias"])
    v = F.conv1d(feat, sd[p + "projection_v.weight"], sd[p + "projection_v.bias"])
    if "sc_proj_tf32" in MODE:
        f_ = tf32(feat)
        q = F.conv1d(f_, tf32(sd[p + "projection_q.weight"]), sd[p + "projection_q.bias"])
        k = F.conv1d(f_, tf32(sd[p + "projection_k.weight"]), sd[p + "projection_k.bias"])
        v = F.conv1d(f_, tf32(sd[p + "projection_v.weight"]), sd[p + "projection_v.bias"])
    if "sc_qk_bf16" in MODE:
        q, k = bf16(q), bf16(k)
    if "sc_qk_fp16" in MODE:
        q, k = fp16(q * (1.4426950408889634 / feat.shape[1] ** 0.5)) / (1.4426950408889634 / feat.shape[1] ** 0.5), fp16(k)
    if "sc_qk_split" in MODE:                                 # 2-term split: qh kh + qh kl + ql kh
        qh, kh = bf16(q), bf16(k)
        ql, kl = bf16(q - qh), bf16(k - kh)
        logits = (torch.einsum("bco,bci->boi", qh, kh) + torch.einsum("bco,bci->boi", qh, kl) + torch.einsum("bco,bci->boi", ql, kh)) / feat.shape[1] ** 0.5
    else:
        logits = torch.einsum("bco,bci->boi", q, k) / feat.shape[1] ** 0.5
    STATS.setdefault("sc_logit_absmax", []).append(float((compat * logits).abs().max()))
    w = torch.softmax(compat * logits, dim=-1)
    if "sc_pv_bf16" in MODE:
        # the kernel rounds the UNnormalised p = exp2(t - ref) to bf16 and V to bf16, and normalises with the fp32 row sum
        t = compat * logits
        pu = torch.exp(t - t.max(-1, keepdim=True)[0])
        w = bf16(pu) / pu.sum(-1, keepdim=True)
        v = bf16(v)
    if "sc_pv_bf16_rsum" in MODE:                             # normalise with the sum of the ROUNDED probabilities
        t = compat * logits
        pu = bf16(torch.exp(t - t.max(-1, keepdim=True)[0]))
        w = pu / pu.sum(-1, keepdim=True)
        v = bf16(v)
    if "sc_p_bf16_v_fp16" in MODE:
        t = compat * logits
        pu = torch.exp(t - t.max(-1, keepdim=True)[0])
        w = bf16(pu) / pu.sum(-1, keepdim=True)
        v = fp16(v)
    return torch.einsum("boi,bci->bco", w, v)


def cross_attention(sd, p, x, ctx):
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    cn = F.layer_norm(ctx, (ctx.shape[-1],), sd[p + "norm_context.weight"], sd[p + "norm_context.bias"], 1e-5)
    wq, wkv, wo = sd[p + "fn.to_q.weight"], sd[p + "fn.to_kv.weight"], sd[p + "fn.to_out.weight"]
    if "fus_proj_tf32" in MODE:
        xn, cn, wq, wkv = tf32(xn), tf32(cn), tf32(wq), tf32(wkv)
    q = xn @ wq.T
    kv = cn @ wkv.T
    d = q.shape[-1]
    k, v = kv[..., :d], kv[..., d:]
    if "fus_qk_bf16" in MODE:
        q, k = bf16(q), bf16(k)
    if "fus_qk_fp16" in MODE:
        q, k = fp16(q), fp16(k)
    sim = torch.einsum("bid,bjd->bij", q, k) * d ** -0.5
    STATS.setdefault("fus_logit_absmax", []).append(float(sim.abs().max()))
    w = sim.softmax(dim=-1)
    if "fus_pv_bf16" in MODE:
        pu = torch.exp(sim - sim.max(-1, keepdim=True)[0])
        w = bf16(pu) / pu.sum(-1, keepdim=True)
        v = bf16(v)
    if "fus_pv_bf16_rsum" in MODE:
        pu = bf16(torch.exp(sim - sim.max(-1, keepdim=True)[0]))
        w = pu / pu.sum(-1, keepdim=True)
        v = bf16(v)
    if "fus_p_bf16_v_fp16" in MODE:
        pu = torch.exp(sim - sim.max(-1, keepdim=True)[0])
        w = bf16(pu) / pu.sum(-1, keepdim=True)
        v = fp16(v)
    out = torch.einsum("bij,bjd->bid", w, v)
    if "fus_proj_tf32" in MODE:
        out, wo = tf32(out), tf32(wo)
    return out @ wo.T + sd[p + "fn.to_out.bias"]


_orig_conv1d = F.conv1d


def conv1d_tf32(x, w, b=None, *a, **k):
    """PointCN / fc_message 1x1 convolutions run as TF32 tensor-core GEMMs (classifier [32|1 out channels] and layer0 [6 in] stay fp32)."""
    if w.dim() == 3 and w.shape[2] == 1 and w.shape[0] in (64, 128) and w.shape[1] in (64, 128) and not a and not k:
        if "mlp_tf32" in MODE:
            return _orig_conv1d(tf32(x), tf32(w), b)
        if "mlp_tf32_act" in MODE:
            return _orig_conv1d(tf32(x), w, b)
        if "mlp_tf32_w" in MODE:
            return _orig_conv1d(x, tf32(w), b)
        if "mlp_f16x2" in MODE:      # 2-term fp16 split of both operands, 3 products
            xh, wh = fp16(x), fp16(w)
            xl, wl = fp16(x - xh), fp16(w - wh)
            return _orig_conv1d(xh, wh, b) + _orig_conv1d(xl, wh) + _orig_conv1d(xh, wl)
        if "mlp_bf16x3" in MODE:     # 3-term bf16 split of the activations, 2-term of the weights: xh wh + xm wh + xh wm (+ xl wh + xm wm + xh wl)
            xh = bf16(x); xm = bf16(x - xh); xl = bf16(x - xh - xm)
            wh = bf16(w); wm = bf16(w - wh); wl = bf16(w - wh - wm)
            return (_orig_conv1d(xh, wh, b) + _orig_conv1d(xm, wh) + _orig_conv1d(xh, wm) + _orig_conv1d(xl, wh) + _orig_conv1d(xm, wm) + _orig_conv1d(xh, wl))
    return _orig_conv1d(x, w, b, *a, **k)


def geglu_ff(sd, p, x):
    xn = F.layer_norm(x, (x.shape[-1],), sd[p + "norm.weight"], sd[p + "norm.bias"], 1e-5)
    w1, w2 = sd[p + "fn.net.0.weight"], sd[p + "fn.net.2.weight"]
    if "ffn_16" in MODE:
        xn, w1 = fp16(xn), fp16(w1)
    h = xn @ w1.T + sd[p + "fn.net.0.bias"]
    half = h.shape[-1] // 2
    h = h[..., :half] * F.gelu(h[..., half:])
    if "ffn_16" in MODE:
        h, w2 = tf32(h), tf32(w2)
    return h @ w2.T + sd[p + "fn.net.2.bias"]


def run(sd, cfg, args, modes):
    MODE.clear(); MODE.update(modes); STATS.clear()
    O.sc_nonlocal_attention, O.cross_attention, O.geglu_ff = sc_attn, cross_attention, geglu_ff
    O.F.conv1d = conv1d_tf32
    return O.forward_testing(sd, cfg, *args)


def main():
    shape = sys.argv[1] if len(sys.argv) > 1 else "kitti"
    n = int(sys.argv[2]) if len(sys.argv) > 2 else 1500
    t = int(sys.argv[3]) if len(sys.argv) > 3 else 600
    layers = int(sys.argv[4]) if len(sys.argv) > 4 else 12
    extent, thr, noise = (60.0, 1.2, 0.04) if shape == "kitti" else (3.0, 0.10, 0.002)
    torch.set_num_threads(os.cpu_count() or 8)
    cfg = dict(O.DEFAULT_CFG, num_layers=layers, inlier_threshold=thr, nms_radius=thr, sigma_d=thr)
    wseed, dseed, plain = int(os.environ.get("WSEED", 0)), int(os.environ.get("DSEED", 301)), os.environ.get("PLAIN", "1") == "1"
    sd = synth_state_dict(hot_path_spec(layers), seed=wseed, plain_init=plain)
    sd["sigma_spat"] = torch.tensor([thr])
    pr = synth_pairs(1, n, seed=dseed, extent=extent, inlier_ratio=float(os.environ.get("INLIERS", 0.30)), noise=noise)
    args = [pr["corr_pos"], pr["src_keypts"], pr["tgt_keypts"], synth_tokens(1, t, 1), synth_tokens(1, t, 2)]
    ref = run(sd, cfg, args, [])
    print(shape, "N", n, "T", t, "layers", layers, "logit range", float(ref["confidence"].min()), float(ref["confidence"].max()))
    print("  SC logit |c*qk/sqrt(C)| max per layer:", [round(x, 1) for x in STATS["sc_logit_absmax"]])
    print("  fusion logit max per call:", [round(x, 1) for x in STATS["fus_logit_absmax"]])
    allm = ["sc_proj_tf32", "fus_proj_tf32", "mlp_tf32", "ffn_16"]
    if len(sys.argv) > 5:                                    # explicit list of '+'-joined mode sets
        for ms in sys.argv[5:]:
            modes = [m for m in ms.replace("ALL", "+".join(allm)).split("+") if m]
            out = run(sd, cfg, args, modes)
            print(f"  {'+'.join(modes):90s} max|dlogit| {float((out['confidence'] - ref['confidence']).abs().max()):.5f}")
        return
    for modes in (["mlp_tf32"], ["ffn_16"],
                  allm + ["sc_qk_bf16", "sc_pv_bf16", "fus_qk_bf16", "fus_pv_bf16"],
                  allm + ["sc_qk_fp16", "sc_pv_bf16", "fus_qk_fp16", "fus_pv_bf16"],
                  allm + ["sc_qk_fp16", "sc_pv_bf16_rsum", "fus_qk_fp16", "fus_pv_bf16_rsum"],
                  ["sc_pv_bf16_rsum"], ["fus_pv_bf16_rsum"],
                  ["sc_qk_fp16", "sc_pv_bf16_rsum", "sc_proj_tf32", "fus_qk_fp16", "fus_pv_bf16_rsum", "fus_proj_tf32"],
                  ["sc_qk_fp16"], ["sc_p_bf16_v_fp16"], ["fus_qk_fp16"], ["fus_p_bf16_v_fp16"],
                  ["sc_qk_fp16", "sc_pv_bf16", "sc_proj_tf32", "fus_qk_fp16", "fus_pv_bf16", "fus_proj_tf32"],
                  ["sc_qk_fp16", "sc_p_bf16_v_fp16", "sc_proj_tf32", "fus_qk_fp16", "fus_p_bf16_v_fp16", "fus_proj_tf32"],
                  ["sc_qk_bf16"], ["sc_pv_bf16"], ["sc_proj_tf32"], ["fus_qk_bf16"], ["fus_pv_bf16"], ["fus_proj_tf32"],
                  ["sc_qk_split", "sc_pv_bf16", "sc_proj_tf32", "fus_qk_bf16", "fus_pv_bf16", "fus_proj_tf32"],
                  ["sc_qk_bf16", "sc_pv_bf16", "sc_proj_tf32", "fus_qk_bf16", "fus_pv_bf16", "fus_proj_tf32"]):
        out = run(sd, cfg, args, modes)
        print(f"  {'+'.join(modes):90s} max|dlogit| {float((out['confidence'] - ref['confidence']).abs().max()):.5f}")


if __name__ == "__main__":
    main()
