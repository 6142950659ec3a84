"""Timeline of the second tile of CTA 3 of the all-layer context K / V projection kernel (GMF_FFN_TRACE build; tools/run_ffn_trace_only.sh)."""
import numpy as np
t = np.fromfile('gpurun_out/kv_trace.bin', dtype=np.int64).reshape(8, 64)
t0 = t[0, 0]
r = lambda x: int(x - t0) if x else None
print('layer | LN warp: start, a_free got, a_ready | MMA: start, acc_free got, operands got | epilogue: start, acc got, staging free, stores issued')
for l in range(12):
    print(l, '|', [r(x) for x in t[0, 3 * l:3 * l + 3]], '|', [r(x) for x in t[1, 3 * l:3 * l + 3]], '|', [r(x) for x in t[2, 4 * l:4 * l + 4]])
