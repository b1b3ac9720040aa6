"""Fit of the polynomial in csrc/fm.cu:atan2_finite: atan(a) = a * P(a^2) on a in [0, 1], degree 8 in a^2, iteratively
re-weighted least squares towards the minimax solution; prints the F32 coefficients and the max |error| of the F32 Horner
evaluation over 2e6 points."""
import numpy as np


def f(u):
    a = np.sqrt(u)
    return np.where(a > 0, np.arctan(a) / np.where(a > 0, a, 1), 1.0)


deg = 8
u = np.cos(np.pi * (np.arange(4000) + 0.5) / 4000) * 0.5 + 0.5
A = np.vander(u, deg + 1, increasing=True)
w = np.ones_like(u)
for _ in range(60):
    c, *_ = np.linalg.lstsq(A * w[:, None], f(u) * w, rcond=None)
    err = A @ c - f(u)
    w = w * (1 + 3 * np.abs(err) / np.abs(err).max())
    w /= w.mean()
a = np.linspace(0, 1, 2000001).astype(np.float32)
uu = (a * a).astype(np.float32)
c32 = c.astype(np.float32)
p = np.full_like(uu, c32[-1])
for k in range(deg - 1, -1, -1):
    p = (p * uu + c32[k]).astype(np.float32)
r = (p * a).astype(np.float32)
print("max |error| (F32 Horner):", np.abs(r.astype(np.float64) - np.arctan(a.astype(np.float64))).max())
print("coefficients (constant first):", [repr(float(x)) for x in c32])
