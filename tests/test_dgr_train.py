"""DGR head training step (BASELINE.json configs[4]; reference core/trainer.py:226-300 around model/perceiver_io.py:187-221).
Oracle = torch autograd of oracle/dgr_head_oracle.py, itself pinned to autograd of the UNMODIFIED reference `PerceiverIO` (oracle/_ref or
/root/reference).  CUDA path (gmf_dgr_head_train_forward / _backward / gmf_sgd_step): forward within the inference tolerance, every weight
gradient within 1e-2 of its own max-norm (TF32 products, fp32 accumulate), SGD step equal to torch.optim.SGD on the same gradients."""
import importlib.util
import sys

import pytest
import torch

from conftest import record
from gmf_b200.dgr_head import dgr_head_shapes
from gmf_b200.synth import synth_state_dict, synth_tokens
from oracle import ref_shim
from oracle.dgr_head_oracle import dgr_head_forward, synth_latents


def _oracle_grads(sd, x, ctx, d_out, pe=True, dtype=torch.float64):
    W = {k: v.detach().clone().to(dtype).requires_grad_(True) for k, v in sd.items()}
    xx, cc = x.detach().clone().to(dtype).requires_grad_(True), ctx.detach().clone().to(dtype).requires_grad_(True)
    out = dgr_head_forward(W, xx, cc, pe=pe, dtype=dtype)
    (out * d_out.to(dtype)).sum().backward()
    return out.detach(), {k: v.grad for k, v in W.items()}, xx.grad, cc.grad


def test_oracle_autograd_matches_reference_module_autograd():
    path = ref_shim.dgr_file("model/perceiver_io.py")
    if path is None:
        pytest.skip("reference tree / oracle/_ref not present")
    sys.dont_write_bytecode = True
    spec = importlib.util.spec_from_file_location("ref_perceiver_io_train", path)
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    for pe in (True, False):
        sd = synth_state_dict(dgr_head_shapes(pe), seed=3)
        m = mod.PerceiverIO(dim=128, depth=0, latent_dim=256, cross_heads=1, latent_heads=8, cross_dim_head=128, latent_dim_head=128, pe=pe).double()
        m.load_state_dict({k: v.double() for k, v in sd.items()}, strict=True)
        x, ctx = synth_latents(90, 7).double().requires_grad_(True), synth_tokens(1, 70, 8)[0].double().requires_grad_(True)
        g = torch.Generator().manual_seed(1)
        d_out = torch.randn(90, 256, generator=g, dtype=torch.float64)
        out = m(ctx.unsqueeze(0), queries_encoder=x.unsqueeze(0))[0]
        (out * d_out).sum().backward()
        o_out, o_g, o_dx, o_dc = _oracle_grads(sd, x.detach(), ctx.detach(), d_out, pe)
        assert torch.allclose(out.detach(), o_out, atol=1e-10)
        assert torch.allclose(x.grad, o_dx, atol=1e-10) and torch.allclose(ctx.grad, o_dc, atol=1e-10)
        for k, p in m.named_parameters():
            assert torch.allclose(p.grad, o_g[k], atol=1e-9), k


def test_gradient_allreduce_host_logic_gloo(tmp_path):
    """The trainer's flat-buffer all-reduce + 1 / world scaling, exercised with world_size 2 on gloo CPU tensors (the CUDA kernels are not needed for
    this logic: sum over ranks, then SGD on the mean)."""
    import torch.multiprocessing as mp
    mp.spawn(_gloo_worker, args=(2, str(tmp_path / "rdzv")), nprocs=2, join=True)


def _gloo_worker(rank, world, rdzv):
    import torch.distributed as dist
    dist.init_process_group("gloo", init_method=f"file://{rdzv}", rank=rank, world_size=world)
    g = torch.full((1000,), float(rank + 1))
    dist.all_reduce(g, op=dist.ReduceOp.SUM)
    assert torch.equal(g, torch.full((1000,), 3.0))
    p, buf = torch.ones(1000), torch.zeros(1000)
    gm = g * (1.0 / world) + 1e-4 * p                          # grad_scale = 1 / world, weight decay
    buf = gm.clone()
    p = p - 0.1 * buf
    assert abs(float(p[0]) - (1 - 0.1 * (1.5 + 1e-4))) < 1e-6
    dist.destroy_process_group()


@pytest.mark.gpu
@pytest.mark.parametrize("m_rows,t_ctx,pe", [(512, 300, True), (2048, 4800, True), (333, 257, False), (1, 1, True)])
def test_cuda_training_step_matches_autograd(m_rows, t_ctx, pe):
    from gmf_b200.dgr_head import DgrHeadTrainer
    sd = synth_state_dict(dgr_head_shapes(pe), seed=11)
    x, ctx = synth_latents(m_rows, 51), synth_tokens(1, t_ctx, 52)[0]
    g = torch.Generator().manual_seed(2)
    d_out = torch.randn(m_rows, 256, generator=g) / m_rows      # d(mean-like loss)/d out
    torch.set_num_threads(8)
    o_out, o_g, o_dx, o_dc = _oracle_grads(sd, x, ctx, d_out, pe)
    tr = DgrHeadTrainer(0, pe)
    tr.load_state_dict(sd)
    out = tr.forward(x.cuda(), ctx.cuda())
    d_x, d_c = tr.backward(d_out.cuda())
    torch.cuda.synchronize()
    fe = float((out.cpu().double() - o_out).abs().max())
    assert fe < 6e-3, fe
    grads = tr.grad_dict()
    worst = ("", 0.0)
    for k, gref in o_g.items():
        rel = float((grads[k].double() - gref).abs().max() / gref.abs().max().clamp_min(1e-12))
        if rel > worst[1]:
            worst = (k, rel)
        assert rel < 1e-2, (k, rel)
    rx = float((d_x.cpu().double() - o_dx).abs().max() / o_dx.abs().max())
    rc = float((d_c.cpu().double() - o_dc).abs().max() / o_dc.abs().max().clamp_min(1e-12))
    record(f"dgr_train_m{m_rows}_t{t_ctx}_pe{int(pe)}", fwd_max_abs_err=fe, worst_weight_grad=worst[0], worst_weight_grad_rel=worst[1], d_latents_rel=rx, d_ctx_rel=rc)
    assert rx < 1e-2 and rc < 1e-2
    # SGD (lr 0.1, momentum 0.8, weight decay 1e-4: core/trainer.py:75-79 with the reference's config defaults) == torch.optim.SGD on the same gradients
    params = [torch.nn.Parameter(v.clone().float()) for v in sd.values()]
    opt = torch.optim.SGD(params, lr=0.1, momentum=0.8, weight_decay=1e-4)
    for _ in range(2):                                          # two steps: momentum buffer initialisation, then its update
        for p_, k in zip(params, sd.keys()):
            p_.grad = grads[k].clone().float()
        opt.step()
        tr.step(lr=0.1, momentum=0.8, weight_decay=1e-4)
    new = tr.state_dict()
    for p_, k in zip(params, sd.keys()):
        assert torch.allclose(new[k], p_.detach(), atol=1e-6), k


@pytest.mark.gpu
def test_cuda_training_error_paths():
    from gmf_b200 import _lib
    from gmf_b200.dgr_head import DgrHeadTrainer
    tr = DgrHeadTrainer(0, True)
    with pytest.raises(_lib.GmfError):
        tr.backward(torch.zeros(4, 256, device="cuda"))        # no forward yet
    lib = _lib.load()
    assert lib.gmf_dgr_head_train_workspace_bytes(0, 5) == 0
    assert lib.gmf_sgd_step(None, None, None, 0, 0.1, 0.8, 0.0, 1.0, 1, None) == -1
