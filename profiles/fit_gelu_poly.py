"""Coefficients of the forward GELU's polynomial (csrc/b200at_gelu.cuh): Phi(-x) = 2^P7(x) on [0, 6.5].

    python profiles/fit_gelu_poly.py

Weighted least squares (weight = Phi(-x), floor 1e-7: the error that matters is the absolute error of Phi, not of its
logarithm) in the Chebyshev basis, converted to monomials in t = -x; then the error of the fp32 Horner evaluation with
fused multiply-adds against double precision."""
import numpy as np
from scipy.special import erf, log_ndtr

XMAX = 6.5


def horner_fma32(d, t):
    acc = np.full(t.shape, np.float32(d[-1]), dtype=np.float32)
    for k in range(len(d) - 2, -1, -1):
        acc = (acc.astype(np.float64) * t.astype(np.float64) + np.float64(np.float32(d[k]))).astype(np.float32)
    return acc


def main():
    x = np.linspace(0, XMAX, 400001)
    y = log_ndtr(-x) / np.log(2.0)
    h = np.exp2(y)
    cheb = np.polynomial.chebyshev.chebfit(x / (XMAX / 2) - 1, y, 7, w=np.maximum(h, 1e-7))
    mono_u = np.polynomial.Polynomial(np.polynomial.chebyshev.cheb2poly(cheb))
    c = mono_u(np.polynomial.Polynomial([-1, 2 / XMAX])).coef          # in x = |v|
    d = [ck * (-1) ** k for k, ck in enumerate(c)]                      # in t = -|v|
    for k, v in enumerate(d):
        print(f'#define B200AT_GELU_P{k} ({np.float32(v):.9e}f)')
    p = horner_fma32(d, (-x).astype(np.float32)).astype(np.float64)
    err = np.abs(np.exp2(p) - h)
    print(f'Phi(-x): max abs error {err.max():.3e}; x * error {np.max(x * err):.3e}')
    v = np.linspace(-12, 12, 960001)
    nax = -np.abs(v)
    hh = np.exp2(horner_fma32(d, np.maximum(nax, -XMAX).astype(np.float32)).astype(np.float64))
    g = nax * hh + np.maximum(v, 0)
    print(f'GELU over [-12, 12]: max abs error {np.abs(g - 0.5 * v * (1 + erf(v / np.sqrt(2)))).max():.3e}')
    t = 1 / (1 + 0.23164189 * x)
    q = (((0.5307027 * t - 0.72657603) * t + 0.7107069) * t - 0.14224836) * t + 0.1274148
    print(f'A&S 7.1.26 form it replaces: x * error {np.max(x * np.abs(q * t * np.exp(-x * x / 2) - h)):.3e}')


if __name__ == '__main__':
    main()
