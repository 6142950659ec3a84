"""cfg#5 (BASELINE.json configs[4]): DGR bottleneck fusion head forward, M latent rows x 256 over T image tokens x 128.
Prints one JSON line per (M, T): device-resident forwards/s (CUDA events around each forward, L2 flushed between iterations),
end-to-end through the PerceiverIO mirror with host tensors (H2D + forward + D2H inside the timed region), the oracle on the
host cores, and the achieved fraction of the tensor roofline.   python tools/bench_dgr_head.py [--iters 50]"""
import argparse
import json
import os
import sys
import time

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from gmf_b200.dgr_head import DgrHeadEngine, dgr_head_shapes     # noqa: E402
from gmf_b200.synth import synth_state_dict, synth_tokens        # noqa: E402
from gmf_b200 import _lib                                        # noqa: E402


def flops(m, t):
    # to_q + to_out, to_kv, QK^T + PV (d = 128), FFN 256 -> 2048 -> (GEGLU) 1024 -> 256
    return 2.0 * m * (256 * 128 * 2) + 2.0 * t * 128 * 256 + 4.0 * m * t * 128 + 2.0 * m * (256 * 2048 + 1024 * 256)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--iters", type=int, default=50)
    ap.add_argument("--no-cpu", action="store_true")
    a = ap.parse_args()
    from oracle.dgr_head_oracle import dgr_head_forward, synth_latents
    peaks = os.path.join(ROOT, "MEASURED_PEAKS.json")
    peak_tf = json.load(open(peaks)).get("bf16_tflops_sustained", 1391.3) if os.path.isfile(peaks) else 1400.0
    dev = torch.device("cuda", 0)
    sd = synth_state_dict(dgr_head_shapes(True), seed=9)
    eng = DgrHeadEngine(0, pe=True)
    eng.load_state_dict(sd)
    lib = _lib.load()
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    for m, t in [(512, 300), (512, 4800), (2048, 300), (2048, 4800)]:
        xh = synth_latents(m, 31).pin_memory()
        ch = synth_tokens(1, t, 32)[0].contiguous().pin_memory()
        oh = torch.empty(m, 256).pin_memory()
        x, c = xh.to(dev), ch.to(dev)
        for _ in range(5):
            out = eng.forward(x, c)
        torch.cuda.synchronize()
        lib.gmf_launch_count(1)
        eng.forward(x, c)
        launches = int(lib.gmf_launch_count(1))
        ms = []
        for _ in range(a.iters):
            flush.fill_(1)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = eng.forward(x, c)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        # end to end: pinned host tensors in, pinned host tensor out
        t0 = None
        for it in range(a.iters + 5):
            if it == 5:
                torch.cuda.synchronize()
                t0 = time.perf_counter()
            xd, cd = xh.to(dev, non_blocking=True), ch.to(dev, non_blocking=True)
            oh.copy_(eng.forward(xd, cd), non_blocking=True)
            torch.cuda.synchronize()
        e2e_ms = (time.perf_counter() - t0) * 1000 / a.iters
        cpu = None
        if not a.no_cpu:
            torch.set_num_threads(os.cpu_count() or 1)
            dgr_head_forward(sd, xh, ch, pe=True)
            t1 = time.perf_counter()
            n = 5
            for _ in range(n):
                ref = dgr_head_forward(sd, xh, ch, pe=True)
            cpu_ms = (time.perf_counter() - t1) * 1000 / n
            err = float((out.cpu() - ref).abs().max())
            cpu = {"ms_per_forward": cpu_ms, "cores": os.cpu_count(), "kind": "port", "max_abs_diff_vs_cuda": err}
        f = flops(m, t)
        print(json.dumps({"metric": "DGR bottleneck head forwards/sec (cfg#5)", "workload": f"M={m} latents x 256, T={t} tokens x 128, pe=True",
                          "value": 1000.0 / med, "ms_per_forward": med, "ms_min": ms[0], "gpu_launches": launches,
                          "e2e": {"ms_per_forward": e2e_ms, "h2d_bytes": (xh.numel() + ch.numel()) * 4, "d2h_bytes": oh.numel() * 4},
                          "roofline": {"bound": "tensor", "flops": f, "achieved": f / (med * 1e-3) / 1e12, "peak": peak_tf, "unit": "TFLOP/s",
                                       "frac": f / (med * 1e-3) / 1e12 / peak_tf, "note": "launch/latency bound at this size"},
                          "cpu_baseline": cpu}), flush=True)


if __name__ == "__main__":
    main()
