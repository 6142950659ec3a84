import sys, torch
sys.path.insert(0, '.')
from oracle import registration_oracle as RO
from gmf_b200.engine import Engine
eng = Engine(num_layers=1)
x, y, ww = RO.synth_problem(1500, 11)
X, Y, w = x[None].cuda(), y[None].cuda(), ww[None, :, 0].contiguous().cuda()
for iters in (1, 5, 10, 20, 35, 50, 75, 100, 150):
    R, t, info = eng.global_registration(X, Y, w, 0.05, iters, 20, 0.0)
    Ro, to, io = RO.global_registration(x, y, ww, 0.05, iters, 20, 0.0)
    print(iters, 'dR', float((R[0].cpu() - Ro).abs().max()), 'dt', float((t[0].cpu() - to).abs().max()), 'loss', float(info[0, 1]), io['loss'])
