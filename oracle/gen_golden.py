"""TEST INFRASTRUCTURE ONLY — generate `tests/golden/*.npz` from the UNMODIFIED reference.

Run in the build container (needs /root/reference):   python oracle/gen_golden.py
Weights and inputs are name-seeded synthetic tensors (gmf_b200.synth), so the fixtures only store
inputs that are cheap to keep (image tokens, points) and the reference's outputs; the state_dict is
regenerated bit-identically from `weight_seed` on any machine.
"""
from __future__ import annotations

import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmf_b200.backbone import ImageEncoder          # noqa: E402
from gmf_b200.synth import synth_pairs, synth_state_dict   # noqa: E402
from gmf_b200.weights import hot_path_spec          # noqa: E402
from oracle import ref_shim                         # noqa: E402

CASES = {
    # name: (num_layers, N, H, W, extent, sigma_d/inlier_thr/nms, inlier_ratio, noise, weight_seed, data_seed, plain_init)
    "l2_n384_3dmatch": dict(num_layers=2, n=384, h=64, w=96, extent=3.0, thr=0.10, ratio_in=0.30, noise=0.002,
                            wseed=1, dseed=11, plain=False, stages=True),
    "l12_n512_3dmatch": dict(num_layers=12, n=512, h=64, w=96, extent=3.0, thr=0.10, ratio_in=0.30, noise=0.0,
                             wseed=2, dseed=12, plain=True, stages=False),
    "l2_n300_kitti": dict(num_layers=2, n=300, h=48, w=160, extent=60.0, thr=1.2, ratio_in=0.40, noise=0.04,
                          wseed=3, dseed=13, plain=False, stages=False),
}


def full_shapes(num_layers):
    shapes = dict(hot_path_spec(num_layers))
    for k, v in ImageEncoder().state_dict().items():
        shapes["encoder.image_encoder." + k] = tuple(v.shape)
    for i in range(num_layers):
        shapes[f"encoder.blocks.PointCN_layer_{i}.1.num_batches_tracked"] = ()
        for j in (1, 4):
            shapes[f"encoder.blocks.NonLocal_layer_{i}.fc_message.{j}.num_batches_tracked"] = ()
    return shapes


def case_cfg(c):
    return dict(num_layers=c["num_layers"], num_iterations=10, ratio=0.1, inlier_threshold=c["thr"], sigma_d=c["thr"],
                k=40, nms_radius=c["thr"])


def case_state_dict(c):
    sd = synth_state_dict(full_shapes(c["num_layers"]), seed=c["wseed"], plain_init=c["plain"])
    sd["sigma_spat"] = torch.tensor([c["thr"]], dtype=torch.float32)
    return sd


def run_case(name, c):
    cfg = case_cfg(c)
    sd = case_state_dict(c)
    ref = ref_shim.build_reference(sd, cfg)
    pairs = synth_pairs(1, c["n"], seed=c["dseed"], extent=c["extent"], inlier_ratio=c["ratio_in"], noise=c["noise"])
    g = torch.Generator().manual_seed(c["dseed"])
    p_img = torch.rand(1, 3, c["h"], c["w"], generator=g)
    q_img = torch.rand(1, 3, c["h"], c["w"], generator=g)
    cap = {}
    hooks = []
    enc = ref.encoder

    def grab(key, idx=None, post=lambda t: t):
        def fn(_m, _i, o):
            cap.setdefault(key, {})[idx] = post(o.detach().clone())
        return fn

    hooks.append(enc.fusion_layer_1.register_forward_hook(grab("image_feat", 0)))
    for i in range(c["num_layers"]):
        hooks.append(enc.blocks[f"NonLocal_layer_{i}"].register_forward_hook(grab("feat_out", i, lambda t: t.permute(0, 2, 1).contiguous())))
    hooks.append(ref.classification.register_forward_hook(grab("confidence", 0, lambda t: t.squeeze(1))))
    # image tokens exactly as the reference builds them (PointDSC.py:129-135)
    with torch.no_grad():
        pt = enc.image_encoder(p_img)
        qt = enc.image_encoder(q_img)
        p_tok = pt.view(1, 128, -1).permute(0, 2, 1).contiguous()
        q_tok = qt.view(1, 128, -1).permute(0, 2, 1).contiguous()
        data = {"corr_pos": pairs["corr_pos"], "src_keypts": pairs["src_keypts"], "tgt_keypts": pairs["tgt_keypts"],
                "p_image": p_img, "q_image": q_img, "testing": True}
        # seeds / seed_trans: wrap the bound methods (no source edits)
        orig_pick, orig_cst = ref.pick_seeds, ref.cal_seed_trans

        def pick(*a, **k):
            s = orig_pick(*a, **k)
            cap["seeds"] = s.clone()
            return s

        def cst(*a, **k):
            r = orig_cst(*a, **k)
            cap["seed_trans"], cap["fitness"], cap["pre_refine"] = r[0].clone(), r[1].clone(), r[2].clone()
            return r

        ref.pick_seeds, ref.cal_seed_trans = pick, cst
        out = ref(data)
    for h in hooks:
        h.remove()
    fx = dict(
        corr_pos=pairs["corr_pos"], src=pairs["src_keypts"], tgt=pairs["tgt_keypts"], gt_trans=pairs["gt_trans"],
        p_tok=p_tok, q_tok=q_tok, p_image=p_img, q_image=q_img,
        final_trans=out["final_trans"], final_labels=out["final_labels"], confidence=cap["confidence"][0],
        seeds=cap["seeds"], fitness=cap["fitness"], pre_refine=cap["pre_refine"], image_feat=cap["image_feat"][0],
        feat=cap["feat_out"][c["num_layers"] - 1],
    )
    if c["stages"]:
        fx["seed_trans"] = cap["seed_trans"]
        for i in range(c["num_layers"]):
            fx[f"feat_out_{i}"] = cap["feat_out"][i]
    meta = {k: v for k, v in c.items()}
    np.savez_compressed(os.path.join(ROOT, "tests", "golden", name + ".npz"),
                        meta=np.array(repr(meta)), **{k: v.numpy() for k, v in fx.items()})
    print(name, "final_trans\n", out["final_trans"][0].numpy(), "\n inliers", int(out["final_labels"].sum()),
          "conf range", float(cap["confidence"][0].min()), float(cap["confidence"][0].max()))


if __name__ == "__main__":
    if not ref_shim.available():
        raise SystemExit("reference not present; fixtures are generated in the build container only")
    torch.set_num_threads(8)
    for n, c in CASES.items():
        run_case(n, c)
