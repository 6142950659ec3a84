"""ONE data-parallel training step of GMF-PointDSC per rank and iteration at the reference's training shape
(config_3DMatch.py: batch 16, num_node 1000; 4800 image tokens per fragment, 12 layers): training-mode forward + losses + analytic backward +
flat gradient all-reduce over NCCL / NVLink + Adam (lr 1e-4, weight decay 1e-6).  Every rank holds its own batch (weak scaling, one process per
GPU like the reference's one-process trainer would be replicated); the only collective is the all-reduce of the flat fp32 gradient.

    python tools/bench_pdsc_train.py [--iters 5] [--precision tf32x3|tf32] [--batch 16] [--corr 1000] [--tokens 4800] [--layers 12]
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 tools/bench_pdsc_train.py

Rank 0 prints one JSON line: steps/s and pairs/s (max over ranks of the CUDA-event time), the split forward+backward / all-reduce+Adam, and the
unmodified reference (torch autograd on the host cores, oracle/_ref) timed on ONE pair-batch of 2 beside it (bounded sample)."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmf_b200.synth import synth_pairs, synth_state_dict, synth_tokens   # noqa: E402
from gmf_b200.trainer import PointDSCTrainer                             # noqa: E402
from gmf_b200.weights import hot_path_spec                               # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=5)
    ap.add_argument("--batch", type=int, default=16)
    ap.add_argument("--corr", type=int, default=1000)
    ap.add_argument("--tokens", type=int, default=4800)
    ap.add_argument("--layers", type=int, default=12)
    ap.add_argument("--precision", default="tf32x3")
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = synth_state_dict(hot_path_spec(a.layers), seed=9)
    sd["sigma_spat"] = torch.tensor([0.1])
    tr = PointDSCTrainer(a.layers, local, precision=a.precision)
    tr.load_state_dict(sd)
    data = synth_pairs(a.batch, a.corr, seed=100 + rank, noise=0.01)                 # every rank its own batch
    pt, qt = synth_tokens(a.batch, a.tokens, 200 + rank), synth_tokens(a.batch, a.tokens, 300 + rank)
    args = [x.to(dev) for x in (data["corr_pos"], data["src_keypts"], data["tgt_keypts"], pt, qt, data["gt_labels"])]
    from gmf_b200 import _lib
    lib = _lib.load()

    def step(timers=None):
        ev = [torch.cuda.Event(enable_timing=True) for _ in range(3)] if timers is not None else None
        if ev: ev[0].record()
        out = tr.forward_backward(*args)
        if ev: ev[1].record()
        tr.step(lr=1e-4, weight_decay=1e-6)                      # NCCL all-reduce (sum) of the flat gradient + guard + Adam on the mean
        if ev:
            ev[2].record()
            timers.append(ev)
        return out
    for _ in range(2):
        out = step()
    torch.cuda.synchronize()
    if world > 1:
        dist.barrier()
    lib.gmf_launch_count(1)
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    timers = []
    e0.record()
    for _ in range(a.iters):
        out = step(timers)
    e1.record()
    torch.cuda.synchronize()
    launches = int(lib.gmf_launch_count(0)) // a.iters
    ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
    if world > 1:
        dist.all_reduce(ms, op=dist.ReduceOp.MAX)
    parts = [sum(ev[i].elapsed_time(ev[i + 1]) for ev in timers) / len(timers) for i in range(2)]
    cpu = None
    if rank == 0 and not a.no_cpu:
        from oracle import ref_shim, train_oracle
        if ref_shim.available() and train_oracle.load_reference_losses() is not None:
            torch.set_num_threads(os.cpu_count() or 1)
            small = {k: (v[:2] if torch.is_tensor(v) and v.shape[0] == a.batch else v) for k, v in data.items()}
            small["p_tokens"], small["q_tokens"] = pt[:2], qt[:2]
            cfg = dict(num_layers=a.layers, num_iterations=10, ratio=0.1, inlier_threshold=0.1, sigma_d=0.1, k=40, nms_radius=0.1)
            t1 = time.perf_counter()
            train_oracle.reference_training_step(sd, cfg, small, dtype=torch.float32)
            dt = time.perf_counter() - t1
            cpu = {"pairs_per_s": 2.0 / dt, "ms_per_pair": dt * 500.0, "cores": os.cpu_count(), "kind": "reference",
                   "sample": "one forward + backward of the unmodified reference module and losses on 2 pairs of the same shape (no optimiser step)"}
    if rank == 0:
        msv = float(ms.item())
        print(json.dumps({"metric": "GMF-PointDSC training pairs/sec (forward + losses + backward + NCCL gradient all-reduce + Adam)",
                          "workload": f"{a.batch} pairs x {a.corr} correspondences, {a.tokens} image tokens per fragment, {a.layers} layers per rank and step; "
                                      f"{tr.params.numel() / 1e6:.2f} M parameters; image backbone outside (token gradients returned)",
                          "precision": a.precision, "n_gpus": world, "scaling": "weak", "value": world * a.batch * 1000.0 / msv, "unit": "pairs/s",
                          "steps_per_s_per_gpu": 1000.0 / msv, "ms_per_step": msv, "split_ms": {"forward_backward": parts[0], "allreduce_adam": parts[1]},
                          "allreduce": {"backend": dist.get_backend() if world > 1 else None, "ranks": world, "bytes": int(tr.grads.numel() * 4)},
                          "gpu_launches_per_step": launches, "workspace_gb": tr._ws.numel() / 1e9, "loss": float(out["loss"]), "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
