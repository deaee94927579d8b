"""Fit Q in  GELU(x) = 0.5 x (1 + tanh(x Q(x^2)))  so that tanh(x Q(x^2)) == erf(x / sqrt 2)
(the exact-erf GELU of nn.GELU(), not the 'tanh GELU' constants).  Output: the three coefficients
used by gelu_tanh_fit() in csrc/gemm_tc.cu and the max abs deviation from the exact function."""
import numpy as np
from scipy.special import erf
from scipy.optimize import least_squares

x = np.linspace(1e-3, 6.0, 6000)
exact = lambda z: 0.5 * z * (1 + erf(z / np.sqrt(2)))


def model(c, z):
    s = np.minimum(z * z, 25.0)
    return z * (c[0] + s * (c[1] + s * c[2]))


resid = lambda c: 0.5 * x * (1 + np.tanh(model(c, x))) - exact(x)
c = least_squares(resid, np.array([0.7978845608, 0.0356774, 0.0]), xtol=1e-15, ftol=1e-15, gtol=1e-15).x
for _ in range(30):  # iteratively re-weighted towards minimax
    e = np.abs(resid(c))
    w = e / e.max() + 0.05
    c = least_squares(lambda cc: resid(cc) * w, c, xtol=1e-15, ftol=1e-15, gtol=1e-15).x
xx = np.linspace(-12, 12, 240001)
err = np.abs(0.5 * xx * (1 + np.tanh(model(c, xx))) - exact(xx))
print("coefficients", [float("%.9e" % v) for v in c])
print("max abs error %.3e at x = %.3f" % (err.max(), xx[err.argmax()]))
