"""GPU bring-up probe: runs one named check in this process and prints error statistics.
Usage: python tools/gpu_probe.py <step>   (tools/run_probe.sh runs every step in its own process with a timeout)."""
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from gmf_b200.engine import Engine  # noqa: E402


def stats(name, got, ref):
    got, ref = got.float().cpu(), ref.float().cpu()
    d = (got - ref).abs()
    print(f"  {name}: max|d|={d.max().item():.3e} mean|d|={d.mean().item():.3e} ref_absmax={ref.abs().max().item():.3e} "
          f"nan={int(torch.isnan(got).sum())}", flush=True)
    return d.max().item()


def tf32(x):
    return (x.view(torch.int32) & ~0x1FFF).view(torch.float32) if False else x


def step_linear():
    eng = Engine(num_layers=1)
    g = torch.Generator().manual_seed(0)
    for (k, nout, relu, res, rows) in [(128, 128, True, False, 128), (128, 128, True, False, 1000), (128, 64, True, False, 300),
                                       (64, 64, True, False, 257), (64, 128, False, True, 515)]:
        x = torch.randn(rows, k, generator=g)
        w = torch.randn(nout, k, generator=g) / k ** 0.5
        b = torch.randn(nout, generator=g)
        r = torch.randn(rows, nout, generator=g) if res else None
        ref = x.double() @ w.double().T + b.double()
        if relu:
            ref = ref.clamp(min=0)
        if res:
            ref = ref + r.double()
        out = eng.debug_linear(x.cuda(), w, b, None if r is None else r.cuda(), relu=relu)
        torch.cuda.synchronize()
        stats(f"linear k={k} nout={nout} rows={rows}", out, ref.float())
        if rows == 128 and k == 128:
            d = (out.cpu() - ref.float()).abs()
            print("   row err profile (first 16 rows):", [f"{v:.1e}" for v in d.max(dim=1)[0][:16].tolist()])
            print("   col err profile (first 16 cols):", [f"{v:.1e}" for v in d.max(dim=0)[0][:16].tolist()])


def attn_ref(q, k, v, scale, src=None, tgt=None, sigma=0.1):
    s = torch.einsum("bid,bjd->bij", q.double(), k.double()) * scale
    if src is not None:
        ds = torch.cdist(src.double(), src.double())
        dt = torch.cdist(tgt.double(), tgt.double())
        c = torch.clamp(1 - (ds - dt) ** 2 / sigma ** 2, min=0)
        s = s * c
    return torch.einsum("bij,bjd->bid", s.softmax(-1), v.double()).float()


def step_attn():
    eng = Engine(num_layers=1)
    g = torch.Generator().manual_seed(1)
    for (B, Lq, Lk) in [(1, 128, 128), (1, 100, 300), (2, 515, 1000)]:
        q, k, v = (torch.randn(B, L, 64, generator=g) for L in (Lq, Lk, Lk))
        out = eng.debug_attention(q.cuda(), k.cuda(), v.cuda(), 0.125)
        torch.cuda.synchronize()
        stats(f"attn64 B={B} Lq={Lq} Lk={Lk}", out, attn_ref(q, k, v, 0.125))
    # large-magnitude logits exercise the lazy rescale path
    q, k, v = (torch.randn(1, L, 64, generator=g) for L in (256, 1024, 1024))
    k[:, 700:] *= 6.0
    out = eng.debug_attention(q.cuda(), k.cuda(), v.cuda(), 0.5)
    torch.cuda.synchronize()
    stats("attn64 rescale", out, attn_ref(q, k, v, 0.5))


def step_attn_sc():
    from gmf_b200.synth import synth_pairs
    eng = Engine(num_layers=1)
    g = torch.Generator().manual_seed(2)
    for (B, N) in [(1, 128), (1, 300), (2, 1000)]:
        pr = synth_pairs(B, N, seed=N)
        q, k, v = (torch.randn(B, N, 128, generator=g) for _ in range(3))
        out = eng.debug_attention(q.cuda(), k.cuda(), v.cuda(), 128 ** -0.5, pr["src_keypts"].cuda(), pr["tgt_keypts"].cuda(), 0.1)
        torch.cuda.synchronize()
        stats(f"attnSC B={B} N={N}", out, attn_ref(q, k, v, 128 ** -0.5, pr["src_keypts"], pr["tgt_keypts"], 0.1))


def _golden(name="l2_n384_3dmatch"):
    from conftest import golden_cfg, golden_state_dict, load_golden
    meta, fx = load_golden(name)
    cfg, sd = golden_cfg(meta), golden_state_dict(meta)
    eng = Engine(num_layers=cfg["num_layers"], num_iterations=cfg["num_iterations"], k=cfg["k"], ratio=cfg["ratio"],
                 inlier_threshold=cfg["inlier_threshold"], nms_radius=cfg["nms_radius"])
    eng.load_state_dict(sd)
    return meta, fx, cfg, sd, eng


def step_fusion():
    from oracle import pointdsc_oracle as O
    meta, fx, cfg, sd, eng = _golden()
    out = eng.fusion_layer(-1, fx["q_tok"].cuda(), fx["p_tok"].cuda())
    torch.cuda.synchronize()
    stats("fusion_1 vs golden image_feat", out, fx["image_feat"])
    feat_in = fx["feat_out_0"]
    ref = O.fusion_layer(sd, "encoder.blocks.NonLocal_layer_1.fusion_layer_2.", fx["image_feat"], feat_in, pe=True)
    out = eng.fusion_layer(1, feat_in.cuda(), fx["image_feat"].cuda())
    torch.cuda.synchronize()
    stats("fusion_2 (layer 1) vs oracle", out, ref)


def step_layer():
    from oracle import pointdsc_oracle as O
    meta, fx, cfg, sd, eng = _golden()
    src, tgt = fx["src"], fx["tgt"]
    cm = O.compat_matrix(src, tgt, sd["sigma_spat"])
    feat_in = fx["feat_out_0"]                                   # input of layer 1 (token-major)
    import torch.nn.functional as F
    p = "encoder.blocks.PointCN_layer_1."
    f1 = torch.relu(O._bn_eval(F.conv1d(feat_in.permute(0, 2, 1), sd[p + "0.weight"], sd[p + "0.bias"]), sd, p + "1."))
    msg_ref = O.sc_nonlocal_attention(sd, "encoder.blocks.NonLocal_layer_1.", f1, cm["compat"]).permute(0, 2, 1)
    msg = eng.sc_attention(1, f1.permute(0, 2, 1).contiguous().cuda(), src.cuda(), tgt.cuda())
    torch.cuda.synchronize()
    stats("sc_attention (layer 1) vs oracle", msg, msg_ref)
    out = eng.encoder_layer(1, feat_in.cuda(), src.cuda(), tgt.cuda(), fx["image_feat"].cuda())
    torch.cuda.synchronize()
    stats("encoder_layer 1 vs golden feat_out_1", out, fx["feat_out_1"])


def step_tail():
    from oracle import pointdsc_oracle as O
    meta, fx, cfg, sd, eng = _golden()
    src, tgt, feat = fx["src"], fx["tgt"], fx["feat"]
    normed, conf = eng.classify(feat.cuda())
    torch.cuda.synchronize()
    stats("confidence vs golden", conf, fx["confidence"])
    stats("normed", normed, torch.nn.functional.normalize(feat, dim=-1))
    seeds = eng.pick_seeds(src.cuda(), fx["confidence"].cuda())
    torch.cuda.synchronize()
    print("  seeds identical:", bool((seeds.cpu().long() == fx["seeds"]).all()), "overlap",
          len(set(seeds[0].tolist()) & set(fx["seeds"][0].tolist())), "/", fx["seeds"].shape[1])
    nf = torch.nn.functional.normalize(feat, dim=-1)
    trans, knn, w = eng.seed_hypotheses(nf.cuda(), src.cuda(), tgt.cuda(), fx["seeds"].int().cuda())
    torch.cuda.synchronize()
    cap = {}
    O.seed_hypotheses(nf, src, tgt, fx["seeds"], cfg["k"], sd["sigma"], sd["sigma_spat"], cfg["num_iterations"], cap)
    print("  knn identical rows:", int((knn.cpu().long() == cap["knn_idx"]).all(-1).sum()), "/", knn.shape[1])
    stats("seed weights", w, cap["seed_weight"])
    stats("seed_trans vs golden", trans, fx["seed_trans"])
    final, labels, counts, best, pre = eng.score_hypotheses(fx["seed_trans"].cuda(), src.cuda(), tgt.cuda(), refine=True)
    torch.cuda.synchronize()
    stats("fitness", counts.float() / src.shape[1], fx["fitness"])
    stats("pre_refine", pre, fx["pre_refine"])
    stats("final_trans", final, fx["final_trans"])
    print("  labels equal:", bool((labels.cpu() == fx["final_labels"]).all()))


def step_e2e():
    from oracle import pointdsc_oracle as O
    for name in ["l2_n384_3dmatch", "l12_n512_3dmatch", "l2_n300_kitti"]:
        meta, fx, cfg, sd, eng = _golden(name)
        t0 = time.time()
        out = eng.forward(fx["corr_pos"].cuda(), fx["src"].cuda(), fx["tgt"].cuda(), fx["p_tok"].cuda(), fx["q_tok"].cuda(),
                          testing=True, want_feat=True)
        torch.cuda.synchronize()
        print(name, f"({time.time() - t0:.3f}s, launches={eng.launch_count(True)})")
        stats("feat", out["feat"], fx["feat"])
        stats("confidence", out["confidence"], fx["confidence"])
        print("  seeds overlap", len(set(out["seeds"][0].tolist()) & set(fx["seeds"][0].tolist())), "/", fx["seeds"].shape[1])
        re = O.rotation_error_deg(out["final_trans"].cpu()[:, :3, :3], fx["final_trans"][:, :3, :3]).max().item()
        te = (out["final_trans"].cpu()[:, :3, 3] - fx["final_trans"][:, :3, 3]).norm(dim=-1).max().item()
        print(f"  RE={re:.5f} deg  TE={te * 1000:.4f} mm  labels_equal={bool((out['final_labels'].cpu() == fx['final_labels']).all())}")


def step_big():
    """cfg#2-shaped single chunk: N=5000, T=4800, a few pairs — smoke + rough timing."""
    from gmf_b200.synth import synth_pairs, synth_state_dict, synth_tokens
    from gmf_b200.weights import hot_path_spec
    B = int(os.environ.get("PROBE_B", "4"))
    eng = Engine(num_layers=12)
    eng.load_state_dict(synth_state_dict(hot_path_spec(12), seed=0, plain_init=True))
    pr = synth_pairs(B, 5000, seed=3, noise=0.002)
    p_tok, q_tok = synth_tokens(B, 4800, 1).cuda(), synth_tokens(B, 4800, 2).cuda()
    args = [pr["corr_pos"].cuda(), pr["src_keypts"].cuda(), pr["tgt_keypts"].cuda(), p_tok, q_tok]
    for it in range(3):
        torch.cuda.synchronize()
        t0 = time.time()
        out = eng.forward(*args, testing=True)
        torch.cuda.synchronize()
        print(f"  iter {it}: {time.time() - t0:.4f}s for B={B} -> {B / (time.time() - t0):.1f} pairs/s")
    from oracle import pointdsc_oracle as O
    re = O.rotation_error_deg(out["final_trans"].cpu()[:, :3, :3], pr["gt_trans"][:, :3, :3])
    te = (out["final_trans"].cpu()[:, :3, 3] - pr["gt_trans"][:, :3, 3]).norm(dim=-1)
    print("  RE vs gt (deg):", [f"{v:.4f}" for v in re.tolist()], " TE (mm):", [f"{v * 1000:.3f}" for v in te.tolist()])
    print("  nan in conf:", int(torch.isnan(out["confidence"]).sum()), "inliers:", out["final_labels"].sum(dim=1).tolist())


if __name__ == "__main__":
    name = sys.argv[1]
    print(f"== {name}", flush=True)
    globals()["step_" + name]()
    print(f"== {name} done", flush=True)
