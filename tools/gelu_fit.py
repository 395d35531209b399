#!/usr/bin/env python
"""Fit and check of the MUFU-free GEGLU gate (gemm_tc5.cu gelu_poly2): erf(z) = z P(w), w = 2 z^2 / Z^2 - 1 on |z| <= Z,
weighted least squares on Chebyshev nodes (weight z^2: the error of gelu is ~ x z dP), evaluated exactly as the device does
it (fp32, saturating-FMA clamp, Horner in w) against the exact gelu.   python tools/gelu_fit.py [--Z 3.4 --deg 10]"""
import argparse
from math import erf

import numpy as np
from numpy.polynomial import chebyshev as C

ap = argparse.ArgumentParser()
ap.add_argument("--Z", type=float, default=3.4)
ap.add_argument("--deg", type=int, default=10)
a = ap.parse_args()
verf = np.vectorize(erf)
Z, deg, n = a.Z, a.deg, 6000
t = np.cos(np.pi * (np.arange(n) + 0.5) / n)
z = np.sqrt((t + 1) * 0.5 * Z * Z)
f = np.where(z > 1e-9, verf(z) / np.maximum(z, 1e-9), 2 / np.sqrt(np.pi))
wgt = z * z + 0.05
c = np.linalg.lstsq(C.chebvander(t, deg) * wgt[:, None], f * wgt, rcond=None)[0]
mono = C.cheb2poly(c)
print("coefficients of w^0 .. w^%d:" % deg, [float(np.float32(v)) for v in mono])


def device(xs):
    x = xs.astype(np.float32)
    tt = np.clip(x * np.float32(0.70710678118654752 * 0.5 / Z) + np.float32(0.5), 0, 1).astype(np.float32)
    zc = (tt * np.float32(2 * Z) - np.float32(Z)).astype(np.float32)
    w = ((zc * np.float32(2 / (Z * Z))) * zc - np.float32(1)).astype(np.float32)
    p = np.float32(mono[-1]) * np.ones_like(w)
    for k in range(len(mono) - 2, -1, -1):
        p = (p * w + np.float32(mono[k])).astype(np.float32)
    hx = (np.float32(0.5) * x).astype(np.float32)
    return (hx * (zc * p).astype(np.float32) + hx).astype(np.float64)


xs = np.linspace(-14, 14, 1400001)
ref = 0.5 * xs * (1 + verf(xs / np.sqrt(2)))
err = np.abs(device(xs) - ref)
print(f"max |gelu_poly2 - gelu| = {err.max():.2e} at x = {xs[err.argmax()]:.3f}")
