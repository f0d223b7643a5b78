"""Coefficients of gelu_fast (csrc/tc_epilogue.cuh): weighted minimax-style fit of log2(0.5 * erfc(a / sqrt 2)) on [0, 6.2] (degree 5).

gelu(x) = relu(x) - |x| * q(|x|) with q = 0.5 erfc(|x| / sqrt 2) = 2 ** L(|x|); the weight is the sensitivity of the GELU
value to an error in L (|x| q ln 2).  Prints the monomial coefficients and the fp32 end-to-end error over [-8, 8]."""
import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as C, polynomial as P

X, DEG = 6.2, 5
xs = np.linspace(0, X, 400001)
L = np.log2(0.5 * erfc(xs / np.sqrt(2)))
w = xs * 0.5 * erfc(xs / np.sqrt(2)) * np.log(2) + 1e-12
t = 2 * xs / X - 1
ww = w.copy()
for _ in range(40):                      # iteratively re-weighted least squares towards the minimax solution
    c = C.chebfit(t, L, DEG, w=ww)
    err = (C.chebval(t, c) - L) * w
    ww = ww * (1 + 4 * np.abs(err) / np.abs(err).max()); ww = ww / ww.max() * w.max()
c = C.chebfit(t, L, DEG, w=ww)
pt = C.cheb2poly(c)
coef, acc = np.zeros(DEG + 1), np.array([1.0])
for ck in pt:
    coef[:len(acc)] += ck * acc
    acc = P.polymul(acc, np.array([-1.0, 2.0 / X]))
xf = np.linspace(-60, 60, 4000001).astype(np.float32)
ax = np.abs(xf)          # no clamp: the degree-5 fit has a negative leading coefficient
cf = coef.astype(np.float32)
Lh = np.full_like(ax, cf[-1])
for k in range(DEG - 1, -1, -1):
    Lh = Lh * ax + cf[k]
g = np.maximum(xf, 0) - np.abs(xf) * np.exp2(Lh).astype(np.float32)
ref = 0.5 * xf.astype(np.float64) * (1 + erf(xf.astype(np.float64) / np.sqrt(2)))
print(f"coefficients c0..c{DEG}:", [float(np.float32(v)) for v in coef])
print("fp32 max |gelu error| on [-60, 60]:", float(np.abs(g - ref).max()))
