import numpy as np
from scipy.special import erfc, erf
from scipy.optimize import least_squares
def gelu(x): return 0.5*x*(1+erf(x/np.sqrt(2)))
f=np.float32
def approx(x, c):
    hx=(f(0.5)*x.astype(f)).astype(f); t=np.abs(hx)
    q=f(c[-1])*np.ones_like(t)
    for k in range(len(c)-2,-1,-1):
        q=(q*t+f(c[k])).astype(f)
    e=np.exp2((q*t).astype(f)).astype(f)
    w=(hx+t).astype(f)
    return (w - t*e).astype(f)
xs=np.concatenate([np.linspace(-12,12,48001), np.linspace(-0.5,0.5,4001)])
deg=4
t=np.linspace(1e-4,4.5,4000)   # t' = |x|/2
qstar=np.log2(np.maximum(erfc(np.sqrt(2)*t),1e-300))/t
c=np.polyfit(t,qstar,deg)[::-1]
def res(c):
    tt=np.abs(xs)/2
    q=np.polyval(c[::-1],tt); e=np.exp2(q*tt)
    a=(0.5*xs+tt)-tt*e
    return (a-gelu(xs))
for p in (2,4,8,16,32):
    r=least_squares(lambda c: np.sign(res(c))*np.abs(res(c)*1e4)**(p/2), c, method='lm'); c=r.x
a=approx(xs,c); g=gelu(xs)
print('coeffs', [repr(float(f(v))) for v in c])
print('max abs err (fp32 eval)', np.max(np.abs(a-g)))
m=np.abs(xs)<6
print('max rel err |x|<6', np.max(np.abs(a-g)[m]/np.maximum(np.abs(g)[m],1e-30)))
print('tails', approx(np.array([20.,100.,1e4,-20.,-100.,-1e4, 0.0, -0.0]),c))
