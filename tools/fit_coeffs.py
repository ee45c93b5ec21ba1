"""Fit the odd/even polynomial kernels used by csrc/llampc_math.cuh (near-minimax: Chebyshev-node least squares
followed by a few Remez-style reweighting passes) and report fp32 Horner evaluation error."""
import numpy as np
from numpy.polynomial import polynomial as P

def fit(fun_over_arg, lo, hi, deg, iters=30, n=4001, rel=False):
    # fit g(s) ~ poly(s) on s in [lo,hi]; weights iterate towards equi-oscillation (Lawson)
    k = np.arange(n)
    s = 0.5*(lo+hi) + 0.5*(hi-lo)*np.cos(np.pi*(k+0.5)/n)
    g = fun_over_arg(s)
    w = np.ones(n)
    for _ in range(iters):
        V = np.vander(s, deg+1, increasing=True)
        c, *_ = np.linalg.lstsq(V*w[:,None], g*w, rcond=None)
        e = np.abs(V@c - g)
        w = w*(e/e.max()+1e-3)**0.5
        w /= w.max()
    return c

def horner32(c, s):
    s = s.astype(np.float32); acc = np.full_like(s, np.float32(c[-1]))
    for ck in c[-2::-1]:
        acc = (acc*s + np.float32(ck)).astype(np.float32)   # not fused, pessimistic
    return acc

def report(name, c, f_exact, lo, hi, even=False):
    z = np.linspace(lo, hi, 2000001)
    s32 = (z.astype(np.float32)*z.astype(np.float32)).astype(np.float32)
    p = horner32(c, s32)
    if even:
        val = p
    else:
        # z + z*s*p
        val = (z.astype(np.float32) + (z.astype(np.float32)*s32).astype(np.float32)*p).astype(np.float32)
    ex = f_exact(z.astype(np.float32).astype(np.float64))
    err = np.abs(val.astype(np.float64)-ex)
    pe = np.abs((np.polyval(c[::-1], z*z)*(1 if even else z**3) + (0 if even else z)) - f_exact(z))
    print(f"{name}: deg {len(c)-1} in s; poly abs err {pe.max():.2e}; fp32 eval abs err max {err.max():.2e} mean {err.mean():.2e}")
    print("   coeffs:", ", ".join(f"{v:.10e}f" for v in c))

# atan(z) = z + z*s*A(s), s = z^2 in [0,1]
for deg in (6,7,8,9):
    c = fit(lambda s: np.where(s>0,(np.arctan(np.sqrt(s))/np.sqrt(np.maximum(s,1e-300))-1)/np.maximum(s,1e-300), -1/3), 1e-12, 1.0, deg)
    report("atan[-1,1]", c, np.arctan, -1, 1)
# sin(t) = t + t*s*S(s) on |t|<=pi/2
for deg in (3,4,5):
    c = fit(lambda s: (np.sin(np.sqrt(s))/np.sqrt(s)-1)/s, 1e-12, (np.pi/2)**2, deg)
    report("sin[-pi/2,pi/2]", c, np.sin, -np.pi/2, np.pi/2)
# small angle |a|<=0.5: sin a = a + a*s*S(s); cos a = 1 + s*C(s)
for deg in (1,2,3):
    c = fit(lambda s: (np.sin(np.sqrt(s))/np.sqrt(s)-1)/s, 1e-12, 0.25, deg)
    report("sin[-.5,.5]", c, np.sin, -.5, .5)
for deg in (2,3,4):
    c = fit(lambda s: (np.cos(np.sqrt(s))-1)/s, 1e-12, 0.25, deg)
    z = np.linspace(-.5,.5,200001); s32=(z.astype(np.float32)**2).astype(np.float32)
    val = (np.float32(1)+s32*horner32(c,s32)).astype(np.float32)
    print(f"cos[-.5,.5] deg {deg}: fp32 abs err {np.abs(val-np.cos(z.astype(np.float32).astype(np.float64))).max():.2e}")
    print("   coeffs:", ", ".join(f"{v:.10e}f" for v in c))
