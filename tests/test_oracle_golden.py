"""Pin the oracle restatement against outputs of the UNMODIFIED reference (tests/golden/*.npz,
made by oracle/gen_golden.py) and, when /root/reference is present, against the live reference."""
import pytest
import torch

from conftest import GOLDEN_CASES, golden_cfg, golden_state_dict, load_golden
from oracle import pointdsc_oracle as O
from oracle import ref_shim


@pytest.mark.parametrize("name", GOLDEN_CASES)
def test_oracle_matches_reference_golden(name):
    meta, fx = load_golden(name)
    cfg, sd = golden_cfg(meta), golden_state_dict(meta)
    torch.set_num_threads(8)
    out = O.forward_testing(sd, cfg, fx["corr_pos"], fx["src"], fx["tgt"], fx["p_tok"], fx["q_tok"], capture=True)
    cap = out["capture"][0]
    assert torch.allclose(cap["image_feat"], fx["image_feat"], atol=2e-5, rtol=1e-5)
    scale = float(fx["feat"].abs().max())
    assert (cap["feat"] - fx["feat"]).abs().max() <= 2e-5 * max(1.0, scale)
    assert (out["confidence"] - fx["confidence"]).abs().max() <= 2e-5 * max(1.0, float(fx["confidence"].abs().max()))
    assert torch.equal(cap["seeds"], fx["seeds"])
    assert torch.allclose(cap["fitness"], fx["fitness"], atol=1e-6)
    assert torch.allclose(cap["pre_refine"], fx["pre_refine"], atol=1e-5)
    assert torch.equal(out["final_labels"], fx["final_labels"])
    assert float(O.rotation_error_deg(out["final_trans"][:, :3, :3], fx["final_trans"][:, :3, :3]).max()) < 1e-3
    assert (out["final_trans"][:, :3, 3] - fx["final_trans"][:, :3, 3]).abs().max() < 1e-4
    if "seed_trans" in fx:
        assert torch.allclose(cap["seed_trans"], fx["seed_trans"], atol=1e-4)
        for i in range(meta["num_layers"]):
            assert torch.allclose(cap["feat_out"][i], fx[f"feat_out_{i}"], atol=2e-5, rtol=1e-5)


@pytest.mark.skipif(not ref_shim.available(), reason="reference tree only exists in the build container")
def test_oracle_matches_live_reference_with_backbone():
    """Fresh case not in the fixtures, through the real backbone of both implementations."""
    from gmf_b200.backbone import ImageEncoder
    from gmf_b200.synth import synth_pairs, synth_state_dict
    from oracle.gen_golden import full_shapes

    cfg = dict(O.DEFAULT_CFG, num_layers=3)
    sd = synth_state_dict(full_shapes(3), seed=7)
    ref = ref_shim.build_reference(sd, cfg)
    pairs = synth_pairs(1, 256, seed=5, noise=0.001)
    g = torch.Generator().manual_seed(3)
    p_img, q_img = torch.rand(1, 3, 48, 64, generator=g), torch.rand(1, 3, 48, 64, generator=g)
    with torch.no_grad():
        out_ref = ref({"corr_pos": pairs["corr_pos"], "src_keypts": pairs["src_keypts"], "tgt_keypts": pairs["tgt_keypts"],
                       "p_image": p_img, "q_image": q_img, "testing": True})
    enc = ImageEncoder().eval()
    enc.load_state_dict({k[len("encoder.image_encoder."):]: v for k, v in sd.items() if k.startswith("encoder.image_encoder.")})
    out = O.forward_testing(sd, cfg, pairs["corr_pos"], pairs["src_keypts"], pairs["tgt_keypts"],
                            enc.tokens(p_img), enc.tokens(q_img))
    assert torch.equal(out["final_labels"], out_ref["final_labels"])
    assert torch.allclose(out["final_trans"], out_ref["final_trans"], atol=1e-5)


def test_rigid_transform_recovers_known_pose():
    g = torch.Generator().manual_seed(0)
    a = torch.randn(4, 40, 3, generator=g)
    q, _ = torch.linalg.qr(torch.randn(4, 3, 3, generator=g))
    q[:, :, 2] *= torch.sign(torch.det(q))[:, None]
    t = torch.randn(4, 3, generator=g)
    b = a @ q.transpose(1, 2) + t[:, None]
    tr = O.rigid_transform_3d(a, b, torch.rand(4, 40, generator=g) + 0.1)
    assert torch.allclose(tr[:, :3, :3], q, atol=1e-5) and torch.allclose(tr[:, :3, 3], t, atol=1e-5)
