import numpy as np
t=np.fromfile('gpurun_out/ffn_trace.bin',dtype=np.int64).reshape(4,64)
t0=t[0,0]
r=lambda x:int(x-t0)
print('worker: prologue done',r(t[0,1]),' geglu loop end',r(t[0,2]),' out_full got',r(t[0,3]),' epilogue end',r(t[0,4]))
print('pass | worker: wait_acc_start got_acc geglu_done hfree_got | mma1: start got_free kc0 kc1 | mma2: start got')
for p in range(8):
    w=t[3,4*p:4*p+4]; m=t[1,4*p:4*p+4]; m2=t[2,2*p:2*p+2]
    print(p,'|',[r(x) for x in w],'|',[r(x) for x in m],'|',[r(x) for x in m2])
