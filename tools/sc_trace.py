"""Decode gpurun_out/sc_trace.bin (GMF_SC_TRACE build of the SC attention kernel): clock64 timeline of one CTA.
roles 0..3 = softmax warps 0, 4, 8, 12 (group 0 half 0/1, group 1 half 0/1): stamps [loop top, s_full got, tile in registers + s_free arrive,
compute done, pv_done got, P stored + p_ready arrive]; role 4 = score issuer [wait s_free start, k_full got, issued]; role 5 = PV issuer
[wait start, p_ready got, issued]; role 6 = producer [K stage issue, V stage issue]."""
import sys

import numpy as np

t = np.fromfile(sys.argv[1] if len(sys.argv) > 1 else "gpurun_out/sc_trace.bin", dtype=np.int64).reshape(7, 64, 8)
t0 = t[t > 0].min()
r = lambda x: int(x - t0) if x > 0 else -1      # noqa: E731
print("tile | softmax g0h0: top sfull ld done pvdone pready | g1h0 ... | score issuer: wait kfull issued | pv issuer: wait got issued")
for j in range(6, 40):
    g = j & 1
    a = t[2 * g, j, :6]
    b = t[2 * g + 1, j, :6]
    print(f"{j:3d} g{g} | h0 {[r(x) for x in a]} | h1 {[r(x) for x in b]} | S {[r(x) for x in t[4, j, :3]]} | PV {[r(x) for x in t[5, j, :3]]}")
# phase statistics over tiles 8..55
def stat(name, v):
    v = np.array([x for x in v if x > -10**8])
    print(f"{name:52s} mean {v.mean():7.1f}  p50 {np.median(v):7.1f}  max {v.max():7.0f}")
for role in range(4):
    js = [j for j in range(8, 56) if (j & 1) == (role >> 1)]
    s = t[role]
    stat(f"softmax role {role}: wait s_full", [s[j, 1] - s[j, 0] for j in js])
    stat(f"softmax role {role}: tmem ld + s_free arrive", [s[j, 2] - s[j, 1] for j in js])
    stat(f"softmax role {role}: compute", [s[j, 3] - s[j, 2] for j in js])
    stat(f"softmax role {role}: wait pv_done", [s[j, 4] - s[j, 3] for j in js])
    stat(f"softmax role {role}: P store + arrive", [s[j, 5] - s[j, 4] for j in js])
    stat(f"softmax role {role}: period (2 tiles)", [s[j + 2, 0] - s[j, 0] for j in js[:-1]])
js = list(range(8, 56))
stat("score issuer: wait s_free", [t[4, j, 1] - t[4, j, 0] for j in js])
stat("score issuer: issue", [t[4, j, 2] - t[4, j, 1] for j in js])
stat("score issuer: period per tile", [t[4, j + 1, 2] - t[4, j, 2] for j in js[:-1]])
stat("pv issuer: wait p_ready", [t[5, j, 1] - t[5, j, 0] for j in js])
stat("pv issuer: issue", [t[5, j, 2] - t[5, j, 1] for j in js])
stat("score issued -> softmax s_full got (MMA exec + wake)", [t[(j & 1) * 2, j, 1] - t[4, j, 2] for j in js])
stat("softmax tile in regs -> score issuer got s_free (j+3)", [t[4, j + 3, 1] - max(t[(j & 1) * 2, j, 2], t[(j & 1) * 2 + 1, j, 2]) for j in js[:-3]])
stat("softmax p_ready -> pv issuer got", [t[5, j, 1] - max(t[(j & 1) * 2, j, 5], t[(j & 1) * 2 + 1, j, 5]) for j in js])
