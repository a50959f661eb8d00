#!/usr/bin/env python
"""Fit the polynomial behind ``gelu_fast`` (open_provence_b200/csrc/common.cuh).

    Phi(-t) = 0.5 * erfc(t / sqrt 2) = 2^-(1 + t * g(t)),  t in [0, 6]

g is fitted by iteratively re-weighted least squares (weight = Phi(-t) * t, the factor that turns an error in
g into an absolute error in Phi) and the fp32 Horner evaluation is checked against the exact erf form.
"""
import numpy as np
from numpy.polynomial import chebyshev as Ch
from scipy.special import erfc

TMAX, DEG = 6.0, 5
t = np.linspace(1e-6, TMAX, 200001)
phi_neg = 0.5 * erfc(t / np.sqrt(2))
g = (-np.log2(phi_neg) - 1) / t
w = phi_neg * t
x = 2 * t / TMAX - 1
ww = w.copy()
for _ in range(30):
    c = Ch.chebfit(x, g, DEG, w=ww)
    err = (Ch.chebval(x, c) - g) * w * np.log(2)
    ww = ww * (1 + 2 * np.abs(err) / np.abs(err).max())
    ww /= ww.max()
poly = np.poly1d([0.0])
for k, ck in enumerate(Ch.cheb2poly(c)):
    poly = poly + ck * np.poly1d([2 / TMAX, -1]) ** k
coef = poly.coeffs[::-1]
tf = t.astype(np.float32)
acc = np.full_like(tf, np.float32(coef[-1]))
for ck in coef[-2::-1]:
    acc = (acc * tf + np.float32(ck)).astype(np.float32)
e = np.exp2((-(tf * acc) - np.float32(1)).astype(np.float32).astype(np.float64))
print("coefficients (ascending):", [float(np.float32(ck)) for ck in coef])
print("max abs error Phi  %.3e" % np.abs(e - phi_neg).max())
print("max abs error gelu %.3e" % max(np.abs(t * (1 - e) - t * (1 - phi_neg)).max(), np.abs(t * e - t * phi_neg).max()))
