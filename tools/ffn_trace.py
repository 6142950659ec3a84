"""Timeline of the SECOND tile of CTA 3 of the persistent fused FFN kernel (GMF_FFN_TRACE build: python -m gmf_b200.build --out
build/libgmf_ffntrace.so -DGMF_FFN_TRACE; tools/run_ffn_iter.sh runs it).  clock64 stamps relative to the worker's start of the tile."""
import numpy as np
t = np.fromfile('gpurun_out/ffn_trace.bin', dtype=np.int64).reshape(8, 64)
t0 = t[0, 0]
r = lambda x: int(x - t0) if x else None
print('worker: tile start 0 | geglu loop + tail end', r(t[0, 2]), '| out_full got', r(t[0, 3]), '| epilogue end', r(t[0, 4]))
print('epilogue (warp 0): residual landed', r(t[0, 7]), '| OUT read', r(t[0, 8]), '| sums done', r(t[0, 9]), '| pair barrier 1', r(t[0, 10]), '| image written + barrier 2', r(t[0, 11]))
print('layernorm warp (tile after this one... stamps of ITS second tile): start', r(t[4, 0]), 'a_free got', r(t[4, 1]), 'a_ready', r(t[4, 2]))
print('pass | worker: wait_acc_start got_acc geglu_done hfree_got | mma1: start got_free issued | mma2: start got')
for p in range(8):
    w = t[3, 4 * p:4 * p + 4]; m = t[1, 4 * p:4 * p + 3]; m2 = t[2, 2 * p:2 * p + 2]
    print(p, '|', [r(x) for x in w], '|', [r(x) for x in m], '|', [r(x) for x in m2])
print('tail | mma2:', [r(x) for x in t[2, 16:18]])
