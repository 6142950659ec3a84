"""cfg#5 (BASELINE.json configs[4]): ONE data-parallel training step of the DGR bottleneck fusion head per rank and iteration:
forward (activations saved) + backward + flat gradient all-reduce over NCCL / NVLink + SGD (lr 0.1, momentum 0.8, weight decay 1e-4).
Each rank holds its own batch of M bottleneck latents and T image tokens (weak scaling, like the reference's DataLoader-per-process DDP-style
setup would); the only collective is the 4.5 MB fp32 gradient all-reduce.

    python tools/bench_dgr_train.py [--iters 30]                                           # 1 GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/bench_dgr_train.py

Rank 0 prints one JSON line per (M, T): steps/s (max over ranks of the CUDA-event time), the split forward / backward / all-reduce + SGD, and
the oracle (torch autograd of the reference restatement on the host cores) beside it."""
import argparse
import json
import os
import sys
import time

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmf_b200.dgr_head import DgrHeadTrainer, dgr_head_shapes    # noqa: E402
from gmf_b200.synth import synth_state_dict, synth_tokens        # noqa: E402


def flops(m, t):
    fwd = 2.0 * m * (256 * 128 * 2) + 2.0 * t * 128 * 256 + 4.0 * m * t * 128 + 2.0 * m * (256 * 2048 + 1024 * 256)
    return 3.0 * fwd                                            # backward = 2 x forward


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=30)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from oracle.dgr_head_oracle import dgr_head_forward, synth_latents
    rank, world, local = int(os.environ.get("RANK", 0)), int(os.environ.get("WORLD_SIZE", 1)), int(os.environ.get("LOCAL_RANK", 0))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)
    sd = synth_state_dict(dgr_head_shapes(True), seed=9)
    tr = DgrHeadTrainer(local, pe=True)
    for m, t in [(512, 300), (2048, 4800)]:
        tr.load_state_dict(sd)
        x = synth_latents(m, 31 + rank).to(dev)                 # every rank its own batch
        c = synth_tokens(1, t, 32 + rank)[0].contiguous().to(dev)
        tgt = torch.zeros(m, 256, device=dev)

        def step(timers=None):
            ev = [torch.cuda.Event(enable_timing=True) for _ in range(4)] if timers is not None else None
            if ev: ev[0].record()
            out = tr.forward(x, c)
            d_out = (out - tgt) * (2.0 / out.numel())           # d/d out of mean((out - target)^2): a stand-in for the network's BCE loss upstream
            if ev: ev[1].record()
            tr.backward(d_out, want_input_grads=True)
            if ev: ev[2].record()
            tr.step(lr=0.1, momentum=0.8, weight_decay=1e-4)    # NCCL all-reduce (sum) of the flat gradient + SGD on the mean
            if ev:
                ev[3].record()
                timers.append(ev)
        for _ in range(5):
            step()
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        timers = []
        e0.record()
        for _ in range(a.iters):
            step(timers)
        e1.record()
        torch.cuda.synchronize()
        ms = torch.tensor([e0.elapsed_time(e1) / a.iters], device=dev)
        if world > 1:
            dist.all_reduce(ms, op=dist.ReduceOp.MAX)
        parts = [sum(ev[i].elapsed_time(ev[i + 1]) for ev in timers) / len(timers) for i in range(3)]
        cpu = None
        if rank == 0 and not a.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            W = {k: v.clone().requires_grad_(True) for k, v in sd.items()}
            xc, cc = x.cpu(), c.cpu()

            def cpu_step():
                out = dgr_head_forward(W, xc, cc, pe=True)
                ((out - tgt.cpu()) ** 2).mean().backward()
            cpu_step()
            t1 = time.perf_counter()
            for _ in range(3):
                cpu_step()
            cpu = {"ms_per_step": (time.perf_counter() - t1) * 1000 / 3, "cores": os.cpu_count(), "kind": "port (torch autograd of the oracle, no optimiser)"}
        if rank == 0:
            f = flops(m, t)
            msv = float(ms.item())
            print(json.dumps({"metric": "DGR bottleneck head training steps/sec per GPU (cfg#5: forward + backward + NCCL gradient all-reduce + SGD)",
                              "workload": f"M={m} latents x 256, T={t} tokens x 128 per rank, pe=True, 0.89 M parameters",
                              "n_gpus": world, "scaling": "weak", "value": 1000.0 / msv, "aggregate_steps_per_s": world * 1000.0 / msv, "ms_per_step": msv,
                              "split_ms": {"forward": parts[0], "backward": parts[1], "allreduce_sgd": parts[2]},
                              "allreduce": {"backend": dist.get_backend() if world > 1 else None, "ranks": world, "bytes": int(tr.grads.numel() * 4)},
                              "achieved_tflops_per_gpu": f / (msv * 1e-3) / 1e12, "cpu_baseline": cpu}), flush=True)
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
