"""Diagnostic (GPU box): per-tensor gradient errors of the CUDA PointDSC training step against autograd of the unmodified reference.
    python tools/check_pdsc_train.py LAYERS B N T [balanced 0/1] [tf32|tf32x3]"""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
from test_pdsc_train import _case, _cfg  # noqa: E402
from oracle import train_oracle  # noqa: E402
from gmf_b200.trainer import PointDSCTrainer  # noqa: E402

L, B, N, T = (int(x) for x in sys.argv[1:5])
balanced = len(sys.argv) > 5 and sys.argv[5] == "1"
sd, data = _case(L, B, N, T, 21)
torch.set_num_threads(8)
t0 = time.time()
ref = train_oracle.reference_training_step(sd, _cfg(L), data, balanced=balanced)
print(f"reference step: {time.time() - t0:.1f} s; losses", ref["class_loss"], ref["sm_loss"])
tr = PointDSCTrainer(L, 0, balanced=balanced, precision=sys.argv[6] if len(sys.argv) > 6 else "tf32x3")
tr.load_state_dict(sd)
out = tr.forward_backward(data["corr_pos"], data["src_keypts"], data["tgt_keypts"], data["p_tokens"], data["q_tokens"], data["gt_labels"])
torch.cuda.synchronize()
print("cuda losses", out["losses"].tolist())
print("logit max abs err", float((out["final_labels"].cpu().double() - ref["logits"]).abs().max()), "max |logit|", float(ref["logits"].abs().max()))
grads = tr.grad_dict()
rows = []
for k, g in ref["grads"].items():
    if k.startswith("encoder.image_encoder.") or g is None:
        continue
    mine = grads[k].double().reshape(g.shape)
    rows.append((float((mine - g).abs().max() / g.abs().max().clamp_min(1e-30)), k, float(g.abs().max()), float(mine.abs().max())))
for r in rows:
    print(f"{r[0]:10.3e}  {r[1]:70s} ref max {r[2]:.3e}  cuda max {r[3]:.3e}")
for n_ in ("d_p_tokens", "d_q_tokens"):
    print(n_, float((out[n_].cpu().double() - ref[n_]).abs().max() / ref[n_].abs().max()))
new = tr.state_dict()
print("running stats max abs err", max(float((new[k].double() - ref["state"][k]).abs().max()) for k in new if "running_" in k))
