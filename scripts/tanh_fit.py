"""Minimax fit (Lawson iteration) of the odd polynomial tanh(x) = x + x s Q(s), s = x^2, on |x| <= T for the fused kernel's
hybrid tanh (pinn_fused.cu: fused_tanh); prints the float32 coefficients and the relative error of a float32 evaluation."""
import numpy as np
from numpy.polynomial import chebyshev as Ch
F=np.float32
def fit(T, n):
    # tanh(x) = x + x*s*Q(s), s=x^2 in [0,T^2]; minimise relative error |x s Q - (tanh - x)|/tanh  -> weight
    N=4000
    k=np.arange(N); s=(0.5-0.5*np.cos(np.pi*(k+0.5)/N))*T*T
    x=np.sqrt(s); x[x==0]=1e-12
    target=(np.tanh(x)-x)/(x*s)       # Q(s)
    w = (x*s)/np.tanh(x)              # relative-error weight
    # iteratively reweighted LS -> approx minimax (Lawson)
    lw=np.ones(N)
    V=np.vander(s, n, increasing=True)
    for it in range(200):
        A=V*(w*lw)[:,None]; b=target*w*lw
        c,*_=np.linalg.lstsq(A,b,rcond=None)
        err=np.abs((V@c-target)*w)
        lw=lw*(err/err.max()+1e-3)**0.5; lw/=lw.max()
    return c, err.max()
def eval32(c, x):
    x=x.astype(F); s=(x*x).astype(F)
    p=F(c[-1])*np.ones_like(s)
    for k in range(len(c)-2,-1,-1): p=(p*s+F(c[k])).astype(F)
    # fma semantic approximated: compute in f64 then round each step
    r=(p.astype(np.float64)*s).astype(F)
    return (r.astype(np.float64)*x+x).astype(F)
for T in (0.25,0.35,0.45,0.55):
    for n in (3,4,5):
        c,e=fit(T,n)
        x=np.linspace(1e-4,T,200001)
        a=eval32(c,x); ref=np.tanh(x.astype(F).astype(np.float64))
        rel=np.abs(a-ref)/ref
        print("T=%.2f n=%d minimax rel %.2e  fp32 eval max rel %.2e (%.2f ulp) mean signed %.2e"%(T,n,e,rel.max(),rel.max()/2**-24, ((a-ref)/ref).mean()), [float(F(v)).hex() for v in c])
