"""Classical SM baseline (baseline_scripts/baseline_3DMatch.py:19-53): the oracle restatement is pinned to the reference's own
source lines (when the reference tree or oracle/_ref is present); the CUDA path (gmf_sm_baseline, N x N matrix never stored) is compared
with the oracle.  Tolerances: leading eigenvector 1e-5 abs (unit-norm vector), labels identical except ties at the top-k boundary, pose 0.01 deg /
1 mm."""
import pytest
import torch

from conftest import record
from gmf_b200.synth import synth_pairs
from oracle import pointdsc_oracle as O
from oracle import sm_oracle


def _case(n, seed, extent=3.0, thr=0.10, inliers=0.3):
    pr = synth_pairs(1, n, seed=seed, extent=extent, inlier_ratio=inliers, noise=0.002 * extent / 3.0)
    return pr, torch.cat([pr["src_keypts"][0], pr["tgt_keypts"][0]], dim=-1), thr


def test_oracle_matches_reference_source_lines():
    ref = sm_oracle.reference_sm()
    if ref is None:
        pytest.skip("reference tree / oracle/_ref not present")
    for n, seed, extent, thr in [(400, 1, 3.0, 0.10), (257, 2, 60.0, 0.6)]:
        pr, corr, _ = _case(n, seed, extent, thr)
        t_ref, l_ref = ref(corr[None], pr["src_keypts"], pr["tgt_keypts"], thr)   # the caller passes corr_pos [1, N, 6]
        t, l, _v = sm_oracle.sm_baseline(corr, pr["src_keypts"], pr["tgt_keypts"], thr)
        assert torch.equal(l, l_ref)
        assert torch.allclose(t, t_ref, atol=1e-5)


@pytest.mark.gpu
@pytest.mark.parametrize("n,extent,thr,inliers", [(1000, 3.0, 0.10, 0.3), (5000, 3.0, 0.10, 0.3), (1301, 60.0, 0.6, 0.4)])
def test_cuda_sm_baseline_matches_oracle(n, extent, thr, inliers):
    from gmf_b200.engine import Engine
    eng = Engine(num_layers=1)
    pr, corr, _ = _case(n, 40 + n, extent, thr, inliers)
    torch.set_num_threads(8)
    t_ref, l_ref, v_ref = sm_oracle.sm_baseline(corr, pr["src_keypts"], pr["tgt_keypts"], thr, dtype=torch.float64)
    trans, labels, v = eng.sm_baseline(pr["src_keypts"].cuda(), pr["tgt_keypts"].cuda(), inlier_threshold=thr)
    dv = float((v.cpu().double() - v_ref).abs().max())
    re = float(O.rotation_error_deg(trans.cpu()[:, :3, :3], t_ref[:, :3, :3].float()).max())
    te = float((trans.cpu()[:, :3, 3].double() - t_ref[:, :3, 3]).norm(dim=-1).max())
    # labels: identical except entries whose eigenvector value ties (to fp32 noise) with the k-th largest
    diff = (labels.cpu().double() != l_ref)[0]
    kth = torch.sort(v_ref[0], descending=True)[0][int(n * 0.1) - 1]
    near = (v_ref[0][diff] - kth).abs().max() if diff.any() else torch.tensor(0.0)
    record(f"sm_baseline_n{n}", max_abs_dv=dv, rot_err_deg=re, trans_err_mm=te * 1e3, labels_differing=int(diff.sum()), boundary_gap=float(near))
    assert dv < 1e-5 and float(near) < 1e-6 and int(diff.sum()) <= 2
    assert re < 0.01 and te < 1e-3
    # batched == per pair
    pr2 = synth_pairs(3, 700, seed=5, noise=0.002)
    tb, lb, vb = eng.sm_baseline(pr2["src_keypts"].cuda(), pr2["tgt_keypts"].cuda())
    for b in range(3):
        t1, l1, v1 = eng.sm_baseline(pr2["src_keypts"][b:b + 1].cuda(), pr2["tgt_keypts"][b:b + 1].cuda())
        assert torch.equal(t1[0], tb[b]) and torch.equal(l1[0], lb[b]) and torch.equal(v1[0], vb[b])


@pytest.mark.gpu
def test_cuda_sm_baseline_error_paths():
    from gmf_b200 import _lib
    from gmf_b200.engine import Engine
    eng = Engine(num_layers=1)
    x = torch.zeros(1, 8, 3, device="cuda")
    with pytest.raises(_lib.GmfError):
        eng.sm_baseline(x, x, top_ratio=0.01)                  # int(8 * 0.01) == 0 seeds
    with pytest.raises(_lib.GmfError):
        eng.sm_baseline(x, x, inlier_threshold=0.0)
