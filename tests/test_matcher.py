"""Correspondence construction: oracle pinned against vectors produced by the reference's own source lines
(oracle/gen_golden_matcher.py); CUDA path (gmf_build_correspondences) against the oracle.  Index parity: identical, except rows whose
two best distances are closer than the fp32 summation-order noise of the dot products (BLAS vs sequential FMA), which are checked to be
such near-ties.  corr_pos: 1e-5 abs (the mean over rows is accumulated in a different order)."""
import os

import numpy as np
import pytest
import torch

from conftest import ROOT
from oracle.matcher_oracle import build_correspondences as oracle_build, synth_descriptors

GOLDEN = ["matcher_n700_m650_d32", "matcher_n900_m1000_d33_mutual"]


def load(name):
    g = np.load(os.path.join(ROOT, "tests", "golden", name + ".npz"))
    s, t, sk, tk = synth_descriptors(int(g["ns"]), int(g["nt"]), int(g["d"]), int(g["seed"]))
    return g, s, t, sk, tk, bool(g["mutual"])


@pytest.mark.parametrize("name", GOLDEN)
def test_oracle_matches_reference_lines(name):
    g, s, t, sk, tk, mutual = load(name)
    o = oracle_build(s, t, sk, tk, mutual)
    assert np.array_equal(o["source_idx"], g["source_idx"]) and np.array_equal(o["corr"], g["corr"])
    assert np.abs(o["corr_pos"] - g["corr_pos"]).max() == 0.0


def _engine():
    from gmf_b200.engine import Engine
    return Engine(num_layers=1)


def _run(eng, s, t, sk, tk, mutual):
    from gmf_b200.matcher import build_correspondences
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
    out = build_correspondences(eng, to(s), to(t), to(sk), to(tk), use_mutual=mutual)
    torch.cuda.synchronize()
    return {k: v.cpu().numpy() for k, v in out.items()}


def _check(out, ref, sk, tk, mutual):
    dist = ref["distance"]
    sidx = out["source_idx"]
    diff = np.nonzero(sidx != ref["source_idx"])[0]
    for i in diff:                                           # only fp32 near-ties may differ
        assert abs(dist[i, sidx[i]] - dist[i, ref["source_idx"][i]]) <= 2e-6, (i, dist[i, sidx[i]], dist[i].min())
    assert len(diff) <= max(1, len(sidx) // 500)
    if len(diff) == 0:
        n = int(out["n_corr"])
        assert n == ref["corr"].shape[0]
        assert np.array_equal(out["corr"][:n], ref["corr"])
        assert np.array_equal(out["src_keypts"][:n], ref["src_keypts"]) and np.array_equal(out["tgt_keypts"][:n], ref["tgt_keypts"])
        assert np.abs(out["corr_pos"][:n] - ref["corr_pos"]).max() <= 1e-5
        assert (out["corr"][n:] == -1).all() and (out["corr_pos"][n:] == 0).all()
        if not mutual:
            assert n == len(sidx)


@pytest.mark.gpu
@pytest.mark.parametrize("name", GOLDEN)
def test_cuda_matcher_matches_reference_golden(name):
    g, s, t, sk, tk, mutual = load(name)
    out = _run(_engine(), s, t, sk, tk, mutual)
    ref = oracle_build(s, t, sk, tk, mutual)
    assert np.array_equal(ref["source_idx"], g["source_idx"])
    _check(out, ref, sk, tk, mutual)


@pytest.mark.gpu
@pytest.mark.parametrize("ns,nt,d,mutual", [(5000, 5000, 32, False), (5000, 4321, 33, True), (1, 1, 32, True), (129, 127, 8, True), (300, 2000, 64, False)])
def test_cuda_matcher_matches_oracle_sizes(ns, nt, d, mutual):
    s, t, sk, tk = synth_descriptors(ns, nt, d, seed=ns + nt)
    out = _run(_engine(), s, t, sk, tk, mutual)
    ref = oracle_build(s, t, sk, tk, mutual)
    _check(out, ref, sk, tk, mutual)


@pytest.mark.gpu
def test_cuda_matcher_tensor_pipe_variant_matches_oracle():
    """GMF_MATCH_IMPL=1: the same matcher as an error-compensated tf32 GEMM with the argmin in its epilogue (kept as an alternative path)."""
    os.environ["GMF_MATCH_IMPL"] = "1"
    try:
        eng = _engine()
    finally:
        del os.environ["GMF_MATCH_IMPL"]
    for ns, nt, d, mutual in [(700, 650, 32, False), (900, 1000, 33, True), (129, 300, 64, True)]:
        s, t, sk, tk = synth_descriptors(ns, nt, d, seed=ns)
        _check(_run(eng, s, t, sk, tk, mutual), oracle_build(s, t, sk, tk, mutual), sk, tk, mutual)


@pytest.mark.gpu
def test_cuda_matcher_exact_ties_take_first_index_and_batches_are_independent():
    """Duplicate target rows give exactly equal distances: np.argmin keeps the first index; batched call == per-pair calls."""
    s, t, sk, tk = synth_descriptors(400, 300, 32, seed=7)
    t2 = np.concatenate([t, t], axis=0)                      # every target appears twice -> ties between j and j + 300
    tk2 = np.concatenate([tk, tk + 1.0], axis=0)
    eng = _engine()
    out = _run(eng, s, t2, sk, tk2, False)
    assert (out["source_idx"] < 300).all()
    ref = oracle_build(s, t2, sk, tk2, False)
    _check(out, ref, sk, tk2, False)
    from gmf_b200.matcher import build_correspondences
    s_b, t_b, sk_b, tk_b = synth_descriptors(400, 600, 32, seed=8)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
    both = build_correspondences(eng, torch.stack([to(s), to(s_b)]), torch.stack([to(t2), to(t_b)]), torch.stack([to(sk), to(sk_b)]),
                                 torch.stack([to(tk2), to(tk_b)]), use_mutual=True)
    one = build_correspondences(eng, to(s_b), to(t_b), to(sk_b), to(tk_b), use_mutual=True)
    for k in one:
        assert torch.equal(both[k][1], one[k]), k


@pytest.mark.gpu
def test_matched_pairs_feed_the_forward_path():
    """descriptors -> correspondences -> PointDSC forward without leaving the device (shape / dtype contract of corr_pos, keypoints)."""
    from gmf_b200.matcher import build_correspondences
    from gmf_b200.synth import synth_state_dict, synth_tokens
    from gmf_b200.weights import hot_path_spec
    from gmf_b200.engine import Engine
    eng = Engine(num_layers=1)
    eng.load_state_dict(synth_state_dict(hot_path_spec(1), seed=0, plain_init=True))
    s, t, sk, tk = synth_descriptors(600, 600, 32, seed=11)
    to = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()      # noqa: E731
    m = build_correspondences(eng, to(s), to(t), to(sk), to(tk), use_mutual=False)
    out = eng.forward(m["corr_pos"][None], m["src_keypts"][None], m["tgt_keypts"][None], synth_tokens(1, 96, 1).cuda(), synth_tokens(1, 96, 2).cuda(),
                      testing=True)
    assert out["final_trans"].shape == (1, 4, 4) and torch.isfinite(out["final_trans"]).all()
