"""Per-kernel time of one PointDSC training step (torch.profiler / CUPTI sees the library's launches too).  GPU box only.
    python tools/profile_pdsc_train.py [tf32|tf32x3]"""
import os
import sys

import torch
from torch.profiler import ProfilerActivity, profile

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from gmf_b200.synth import synth_pairs, synth_state_dict, synth_tokens  # noqa: E402
from gmf_b200.trainer import PointDSCTrainer  # noqa: E402
from gmf_b200.weights import hot_path_spec  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "tf32x3"
B, N, T, L = 16, 1000, 4800, 12
sd = synth_state_dict(hot_path_spec(L), seed=9)
tr = PointDSCTrainer(L, 0, precision=prec)
tr.load_state_dict(sd)
d = synth_pairs(B, N, seed=100, noise=0.01)
args = [x.cuda() for x in (d["corr_pos"], d["src_keypts"], d["tgt_keypts"], synth_tokens(B, T, 200), synth_tokens(B, T, 300), d["gt_labels"])]
for _ in range(2):
    tr.forward_backward(*args)
torch.cuda.synchronize()
with profile(activities=[ProfilerActivity.CUDA]) as p:
    tr.forward_backward(*args)
    torch.cuda.synchronize()
print(p.key_averages().table(sort_by="cuda_time_total", row_limit=25, max_name_column_width=70))
