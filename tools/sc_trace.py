"""Print the per-tile role timeline recorded by GMF_SC_TRACE (CTA (0,0) of one SC-attention launch)."""
import sys
import numpy as np
t = np.fromfile(sys.argv[1], dtype=np.int64).reshape(4, 256, 4)
t0 = t[t > 0].min()
lo, hi = int(sys.argv[2]) if len(sys.argv) > 2 else 30, int(sys.argv[3]) if len(sys.argv) > 3 else 50
print("tile | softmax: wait_start got_S loaded done (grp) | mma: step_start got_P got_V issued")
for j in range(lo, hi):
    g = j & 1
    s = t[g, j] - t0
    m = t[2, j] - t0
    print(f"{j:4d} | g{g} {s[0]:8d} {s[1]-s[0]:6d} {s[2]-s[1]:5d} {s[3]-s[2]:6d} | {m[0]:8d} {m[1]-m[0]:6d} {m[2]-m[1]:5d} {m[3]-m[2]:6d}")
d = np.diff(t[2, 20:140, 1])
print("mean clk per tile (mma got_P deltas):", d.mean())
for g in (0, 1):
    js = np.arange(20 + g, 140, 2)
    print(f"group {g}: mean wait {np.mean(t[g, js, 1]-t[g, js, 0]):.0f}  load {np.mean(t[g, js, 2]-t[g, js, 1]):.0f}  compute+store {np.mean(t[g, js, 3]-t[g, js, 2]):.0f}")
js = np.arange(20, 140)
print(f"mma: wait P {np.mean(t[2, js, 1]-t[2, js, 0]):.0f}  wait V {np.mean(t[2, js, 2]-t[2, js, 1]):.0f}  issue {np.mean(t[2, js, 3]-t[2, js, 2]):.0f}")
print("issue_sd(j): wait_k  mma_issue  commits   (tile j issued during step j-3)")
for j in range(lo, hi):
    q = t[3, j]
    print(f"{j:4d} | start {q[0]-t0:8d} wait_k {q[1]-q[0]:6d} mma {q[2]-q[1]:5d} commit {q[3]-q[2]:5d}")
