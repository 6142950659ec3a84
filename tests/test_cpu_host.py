"""CPU-only checks: the C-ABI library builds, loads and exports every declared symbol; the weight table agrees between
Python and C; the nn.Module mirror keeps the reference's state_dict layout; host-side sharding logic (gloo, world 2)."""
import ctypes as C
import os
import re

import pytest
import torch

from conftest import ROOT


@pytest.fixture(scope="module")
def lib():
    import __graft_entry__ as g
    g.build()
    from gmf_b200 import _lib
    return _lib.load()


def test_library_exports_every_declared_symbol(lib):
    hdr = open(os.path.join(ROOT, "include", "gmf_b200.h")).read()
    declared = set(re.findall(r"\b(gmf_[a-z0-9_]+)\s*\(", hdr))
    assert len(declared) >= 20
    for name in declared:
        assert hasattr(lib, name), name
    from gmf_b200 import _lib
    assert declared == set(_lib.EXPORTS)


def test_integration_doc_accounts_for_every_export():
    """INTEGRATION.md section 3 maps every declared entry point to the reference symbol it replaces (or says it has none)."""
    hdr = open(os.path.join(ROOT, "include", "gmf_b200.h")).read()
    doc = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    for name in set(re.findall(r"\b(gmf_[a-z0-9_]+)\s*\(", hdr)):
        short = name.replace("gmf_dgr_head", "")          # the table abbreviates `gmf_dgr_head_create / _weight_spec / ...`
        assert name in doc or (name.startswith("gmf_dgr_head_") and short in doc), name


def test_weight_table_matches_python_spec(lib):
    from gmf_b200.weights import hot_path_spec
    for layers in (1, 2, 12):
        spec = hot_path_spec(layers)
        assert lib.gmf_weight_count(layers) == len(spec)
        buf, numel = C.create_string_buffer(256), C.c_int64()
        for i, (name, shape) in enumerate(spec.items()):
            assert lib.gmf_weight_spec(layers, i, buf, 256, C.byref(numel)) == 0
            n = 1
            for d in shape:
                n *= d
            assert buf.value.decode() == name and numel.value == n
    assert lib.gmf_weight_spec(2, 10 ** 6, None, 0, None) != 0
    assert b"out of range" in lib.gmf_last_error()


def test_pack_state_dict_validates():
    from gmf_b200.synth import synth_state_dict
    from gmf_b200.weights import hot_path_spec, pack_state_dict
    sd = synth_state_dict(hot_path_spec(2), seed=3)
    flat = pack_state_dict(sd, 2)
    assert flat.numel() == sum(v.numel() for v in sd.values())
    bad = dict(sd)
    bad.pop("classification.4.bias")
    with pytest.raises(KeyError):
        pack_state_dict(bad, 2)
    bad = dict(sd)
    bad["encoder.layer0.weight"] = torch.zeros(128, 5, 1)
    with pytest.raises(ValueError):
        pack_state_dict(bad, 2)


def test_module_mirror_state_dict_layout():
    from gmf_b200 import PointDSC
    from gmf_b200.weights import hot_path_spec
    m = PointDSC(num_layers=3, sigma_d=1.2, inlier_threshold=1.2, nms_radius=1.2)
    sd = m.state_dict()
    spec = hot_path_spec(3)
    for k, shape in spec.items():
        assert tuple(sd[k].shape) == tuple(shape), k
    rest = [k for k in sd if k not in spec]
    assert all(k.startswith("encoder.image_encoder.backbone.") or k.endswith("num_batches_tracked") for k in rest)
    assert float(sd["sigma_spat"]) == pytest.approx(1.2) and float(sd["sigma"]) == 1.0
    assert not m.sigma_spat.requires_grad and m.sigma.requires_grad
    with pytest.raises(ValueError):
        PointDSC(num_channels=64)


@pytest.mark.skipif(torch.cuda.is_available(), reason="checks the loud failure without a GPU")
def test_no_cpu_fallback():
    from gmf_b200 import PointDSC
    from gmf_b200._lib import GmfError
    from gmf_b200.engine import Engine
    with pytest.raises(GmfError):
        Engine(num_layers=1)
    m = PointDSC(num_layers=1).eval()
    data = {"corr_pos": torch.zeros(1, 16, 6), "src_keypts": torch.zeros(1, 16, 3), "tgt_keypts": torch.zeros(1, 16, 3),
            "p_image": torch.zeros(1, 3, 32, 32), "q_image": torch.zeros(1, 3, 32, 32), "testing": True}
    with pytest.raises(RuntimeError):
        m(data)


def test_synthetic_generators_are_deterministic():
    from gmf_b200.synth import synth_pairs, synth_state_dict, synth_tokens
    from gmf_b200.weights import hot_path_spec
    a, b = synth_pairs(2, 50, seed=4, noise=0.01), synth_pairs(2, 50, seed=4, noise=0.01)
    assert all(torch.equal(a[k], b[k]) for k in a)
    assert torch.allclose(a["corr_pos"].mean(dim=1), torch.zeros(2, 6), atol=1e-5)
    r = a["gt_trans"][:, :3, :3]
    assert torch.allclose(r @ r.transpose(1, 2), torch.eye(3).expand(2, 3, 3), atol=1e-5) and (torch.det(r) > 0).all()
    s1, s2 = synth_state_dict(hot_path_spec(1), 5), synth_state_dict(hot_path_spec(1), 5)
    assert all(torch.equal(s1[k], s2[k]) for k in s1)
    assert (synth_tokens(1, 10, 0) >= 0).all()


def test_shard_range_partitions():
    from gmf_b200.shard import shard_range
    for n in (1, 7, 64, 256):
        for w in (1, 2, 3, 8):
            parts = [shard_range(n, r, w) for r in range(w)]
            assert parts[0][0] == 0 and parts[-1][1] == n
            assert all(parts[i][1] == parts[i + 1][0] for i in range(w - 1))
            assert max(h - l for l, h in parts) - min(h - l for l, h in parts) <= 1


def _gloo_worker(rank, world, port, q):
    import torch.distributed as dist
    from gmf_b200.shard import gather_poses, shard_range
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    lo, hi = shard_range(7, rank, world)
    local = torch.arange(lo, hi, dtype=torch.float32)[:, None, None] * torch.ones(1, 4, 4)
    allp = gather_poses(local, 7, rank, world)
    q.put((rank, allp[:, 0, 0].tolist()))
    dist.destroy_process_group()


def test_host_gather_world2_gloo():
    import torch.multiprocessing as mp
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 29500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = dict(q.get(timeout=120) for _ in range(2))
    [p.join(timeout=60) for p in ps]
    assert res[0] == res[1] == [float(i) for i in range(7)]


def _gloo_grad_worker(rank, world, port, q):
    import torch.distributed as dist
    from gmf_b200.shard import exchange_gradients
    dist.init_process_group("gloo", init_method=f"tcp://127.0.0.1:{port}", rank=rank, world_size=world)
    g = torch.full((1000,), float(rank + 1))
    w, ok = exchange_gradients(g, check_finite=True)
    bad = torch.ones(8)
    if rank == 1:
        bad[3] = float("nan")                              # a non-finite gradient on ONE rank ...
    _, ok_bad = exchange_gradients(bad, check_finite=True)
    q.put((rank, w, ok, g.unique().tolist(), ok_bad))      # ... must make EVERY rank skip the step
    dist.destroy_process_group()


def test_gradient_exchange_world2_gloo():
    """The collective of the data-parallel training steps (trainer.py / dgr_head.py `step`): sum over ranks, world size for the mean,
    and a finite guard that all ranks agree on."""
    import torch.multiprocessing as mp
    from gmf_b200.shard import exchange_gradients
    g = torch.arange(5.0)
    assert exchange_gradients(g, check_finite=True) == (1, True) and torch.equal(g, torch.arange(5.0))       # no process group: untouched
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    port = 31500 + os.getpid() % 2000
    ps = [ctx.Process(target=_gloo_grad_worker, args=(r, 2, port, q)) for r in range(2)]
    [p.start() for p in ps]
    res = [q.get(timeout=120) for _ in range(2)]
    [p.join(timeout=60) for p in ps]
    for rank, w, ok, vals, ok_bad in res:
        assert w == 2 and ok and vals == [3.0] and not ok_bad


def test_host_chunk_plan_properties(lib):
    """plan_host_chunks (host entry point): chunks cover the batch exactly, respect the cap, start small and never shrink into a short tail."""
    buf = (C.c_int * 256)()
    for B, N, cap in [(64, 5000, 48), (32, 10000, 48), (256, 5000, 48), (9, 5000, 48), (5, 5000, 2), (1, 300, 48), (17, 1000, 8), (3000, 300, 48)]:
        n = lib.gmf_debug_plan_host_chunks(B, N, cap, 148, buf, 256)
        sizes = [buf[i] for i in range(min(n, 256))]
        assert n >= 1 and sum(sizes) == B and all(1 <= s <= cap for s in sizes), (B, N, cap, sizes)
        if B > 8 and n > 1:
            assert sizes[0] <= max(1, B // 9) * 1.25 + 1                # small first chunk: its upload is the only exposed one
            assert all(sizes[i + 1] <= 2.6 * sizes[i] + 1 for i in range(n - 2)), sizes   # uploads stay ahead of the kernels
    assert lib.gmf_debug_plan_host_chunks(64, 5000, 48, 148, buf, 256) == 3 and [buf[i] for i in range(3)] == [7, 14, 43]
    assert lib.gmf_debug_plan_host_chunks(0, 5000, 48, 148, buf, 256) < 0


def test_gelu_polynomial_constants_match_erf_gelu():
    """gelu_erf / geglu2 (gmf_b200/csrc/linear_tc.cuh) evaluate F.gelu (erf form, fusion_layer.py:57) as max(x, 0) - |x| 2^(q(|x|) - 1) with a
    degree-5 polynomial q ~ log2 erfc(|x| / sqrt2) (tools/fit_gelu.py).  The compiled-in coefficients are checked here in fp32 arithmetic against
    the exact function over the whole range the kernels can see: |error| <= 2e-6 (the TF32 / fp16 rounding of the product is ~5e-4 relative)."""
    import math
    import os
    import re

    import numpy as np

    src = open(os.path.join(os.path.dirname(__file__), "..", "gmf_b200", "csrc", "linear_tc.cuh")).read()
    c = [np.float32(re.search(r"#define GMF_GELU_C%d \(([-+0-9.e]+)f\)" % i, src).group(1)) for i in range(1, 6)]
    x = np.concatenate([np.linspace(-60, 60, 600001), [1e4, -1e4, 0.0]]).astype(np.float32)
    n = -np.abs(x)
    q = (-c[4]) * n + c[3]
    q = q * n - c[2]
    q = q * n + c[1]
    q = q * n - c[0]
    with np.errstate(over="ignore", under="ignore"):
        g = n * np.exp2(q * n - np.float32(1.0)) + np.maximum(x, np.float32(0.0))
    ref = np.array([0.5 * v * (1.0 + math.erf(v / math.sqrt(2.0))) for v in x.astype(np.float64)])
    assert not np.isnan(g).any()
    assert np.abs(g - ref).max() <= 2e-6
