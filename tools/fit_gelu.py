import numpy as np
from scipy.special import erfc, erf
from numpy.polynomial import chebyshev as C, polynomial as P

def fit(deg, tmax):
    # R(t) = log2(0.5*erfc(t/sqrt2)) ; fit weighted so that abs error of Q = exp2(R) is minimised
    n = 4000
    k = np.arange(n)
    t = 0.5 * tmax * (1 - np.cos(np.pi * (k + 0.5) / n))
    Q = 0.5 * erfc(t / np.sqrt(2))
    R = np.log2(Q)
    w = Q * np.maximum(t, 0.3)  # abs error of x*Q
    coef = None
    # iteratively reweighted least squares towards minimax of w*(R - p)
    ww = w.copy()
    for it in range(60):
        V = np.vander(t, deg + 1, increasing=True)
        A = V * ww[:, None]
        coef, *_ = np.linalg.lstsq(A, R * ww, rcond=None)
        err = np.abs((V @ coef - R) * w)
        ww = ww * (1 + 2.0 * err / err.max()) ** 0.5
        ww /= ww.max() / w.max()
    return coef

def eval32(coef, x):
    x = x.astype(np.float32)
    t = np.minimum(np.abs(x), np.float32(TMAX))
    r = np.float32(coef[-1]) * np.ones_like(t)
    for c in coef[-2::-1]:
        r = (r.astype(np.float64) * t.astype(np.float64) + np.float64(np.float32(c))).astype(np.float32)  # fma
    q = np.exp2(r.astype(np.float64)).astype(np.float32)
    cdf = np.where(x >= 0, np.float32(1) - q, q)
    return (x * cdf).astype(np.float32)

for deg in (6, 7, 8, 9):
    for TMAX in (6.0, 7.5, 9.0):
        coef = fit(deg, TMAX)
        x = np.linspace(-12, 12, 2000001)
        ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
        got = eval32(coef, x)
        err = np.abs(got - ref)
        rel = err / np.maximum(np.abs(ref), 1e-30)
        print(deg, TMAX, "max abs err %.3e at x=%.3f ; max abs err/max(1,|x|) %.3e" % (err.max(), x[err.argmax()], (err / np.maximum(1, np.abs(x))).max()))
print("----")
TMAX = 9.0
coef = fit(6, TMAX)
x = np.linspace(-12, 12, 2000001)
ref = 0.5 * x * (1 + erf(x / np.sqrt(2)))
got = eval32(coef, x)
err = np.abs(got - ref)
for lo, hi in ((-12, -6), (-6, -3), (-3, -1), (-1, 0), (0, 1), (1, 3), (3, 12)):
    m = (x >= lo) & (x < hi)
    print(lo, hi, "max abs %.3e" % err[m].max())
print(", ".join("%.9ef" % np.float32(c) for c in coef))
