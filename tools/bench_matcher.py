"""Correspondence construction timing: Ns = Nt = 5000 descriptors of dimension 32 (FCGF) / 33 (FPFH) per pair.
Device-resident time per call (CUDA events, L2 flushed between calls) for B = 1 and B = 64 pairs, the NumPy oracle on the host cores,
and the fraction of the FP32 FMA-pipe roofline (148 SMs x 128 FMA/clk x 2 x SM clock).   python tools/bench_matcher.py"""
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from gmf_b200.engine import Engine                       # noqa: E402
from gmf_b200.matcher import build_correspondences       # noqa: E402
from oracle.matcher_oracle import build_correspondences as oracle_build, synth_descriptors   # noqa: E402


def main():
    eng = Engine(num_layers=1)
    dev = torch.device("cuda", 0)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    peak = 148 * 128 * 2 * 1.965e9 / 1e12
    for (n, d, mutual, B) in [(5000, 32, False, 1), (5000, 32, True, 1), (5000, 32, False, 64), (5000, 33, True, 64)]:
        s, t, sk, tk = synth_descriptors(n, n, d, seed=5)
        to = lambda a: torch.from_numpy(a).to(dev)[None].repeat(B, 1, 1).contiguous()      # noqa: E731
        S, T, SK, TK = to(s), to(t), to(sk), to(tk)
        for _ in range(3):
            out = build_correspondences(eng, S, T, SK, TK, use_mutual=mutual)
        torch.cuda.synchronize()
        ms = []
        for _ in range(20):
            flush.fill_(0)
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            out = build_correspondences(eng, S, T, SK, TK, use_mutual=mutual)
            e1.record()
            torch.cuda.synchronize()
            ms.append(e0.elapsed_time(e1))
        ms.sort()
        med = ms[len(ms) // 2]
        t0 = time.perf_counter()
        ref = oracle_build(s, t, sk, tk, mutual)
        cpu_ms = (time.perf_counter() - t0) * 1000
        same = float((out["source_idx"][0].cpu().numpy() == ref["source_idx"]).mean())
        flops = 2.0 * n * n * d * (2 if mutual else 1) * B
        print(json.dumps({"workload": f"Ns=Nt={n}, D={d}, mutual={mutual}, B={B}", "ms_per_call": med, "pairs_per_s": B / (med * 1e-3),
                          "roofline": {"bound": "fp32 FMA pipe", "achieved": flops / (med * 1e-3) / 1e12, "peak": peak, "unit": "TFLOP/s",
                                       "frac": flops / (med * 1e-3) / 1e12 / peak},
                          "cpu_baseline": {"ms_per_pair": cpu_ms, "cores": os.cpu_count(), "kind": "port (NumPy, BLAS threads)"},
                          "source_idx_agreement": same}), flush=True)


if __name__ == "__main__":
    main()
