"""Weighted minimax fit of q(a) ~ log2 erfc(a / sqrt2) (polynomial without constant term) behind gelu_erf / geglu2 in
gmf_b200/csrc/linear_tc.cuh:  gelu(x) = x/2 + |x|/2 (1 - 2^q(|x|)).  Prints, per degree, the erfc and gelu errors and the fp32 coefficients
(degree 5 is the one compiled in; the last line is the Abramowitz-Stegun 7.1.25 form it replaced).   python tools/fit_gelu.py"""
import numpy as np
from scipy.special import erfc, erf
a=np.linspace(0,8,40001)
f=np.log2(erfc(a/np.sqrt(2)))
def gelu(x): return 0.5*x*(1+erf(x/np.sqrt(2)))
for deg in (4,5,6,7):
    # weighted LSQ, iterate reweighting towards minimax of E error
    w=np.ones_like(a)
    E=erfc(a/np.sqrt(2))
    for it in range(60):
        W=w*E
        V=np.vander(a,deg+1,increasing=True)[:,1:]  # no constant term
        c,_res,_r,_s=np.linalg.lstsq(V*W[:,None], f*W, rcond=None)
        q=V@c
        err=np.abs(np.exp2(q)-E)
        w=w*(1+5*err/err.max())
        w/=w.max()
    c32=c.astype(np.float32)
    x=np.linspace(-10,10,400001).astype(np.float32)
    ax=np.abs(x)
    qq=np.zeros_like(ax)
    for k in c32[::-1]: qq=(qq+k)*ax
    Eh=np.exp2(qq.astype(np.float32))
    hx=0.5*x
    g=hx+np.abs(hx)*(1-Eh)
    ref=gelu(x.astype(np.float64))
    print(deg, 'max erfc err', err.max(), 'gelu abs err', np.abs(g-ref).max(), 'rel-to-|x|', (np.abs(g-ref)/np.maximum(np.abs(x),1e-3)).max(), c32)
# current A&S
x=np.linspace(-10,10,400001)
ax=np.abs(x); t=1/(1+0.33267263*ax); p=((0.7478556*t-0.0958798)*t+0.3480242)*t*np.exp2(-0.72134752*x*x)
g=0.5*x+np.abs(0.5*x)*(1-p); print('A&S gelu abs err', np.abs(g-gelu(x)).max())
