"""Fit of order_assoc_tc_shift (feos_torch_b200/csrc/order.cuh): ln(T_c / T_c0) of 2B-associating PC-SAFT fluids as a
cubic in a = 1/m, b = ln kappa_AB, c = eps_AB / (10 eps).  The critical temperatures are located by bisection on the
converged / failed mask of the VLE solver itself (CPU build of the device functions, tests/hostsim).  Only the ORDER in
which the kernels visit the states depends on the result.

    python oracle/fits/fit_order_tc_shift.py
"""
import itertools
import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(ROOT, "tests", "hostsim"))
import sim  # noqa: E402

rng = np.random.default_rng(0)
n = 6000
m = rng.uniform(1.0, 6.0, n)
kap = np.exp(rng.uniform(np.log(0.001), np.log(0.3), n))
r = rng.uniform(3.0, 20.0, n)  # eps_ab / eps
eps, sig = 250.0, 3.5
P = np.zeros((n, 8))
P[:, 0], P[:, 1], P[:, 2], P[:, 4], P[:, 5], P[:, 6], P[:, 7] = m, sig, eps, kap, r * eps, 1, 1
mc = np.clip(m, 1, 5.5)
tc0 = eps * (0.6380 + mc * (0.7070 - 0.06418 * mc))  # non-associating T_c
lo, hi = 0.9 * tc0, 10 * tc0
for _ in range(24):
    mid = 0.5 * (lo + hi)
    failed = sim.pure_vle(P, mid)[3].astype(bool)
    lo = np.where(failed, lo, mid)
    hi = np.where(failed, mid, hi)
y = np.log(lo / tc0)
a, b, c = 1.0 / m, np.log(kap), r / 10.0
idx = [(i, j, k) for i, j, k in itertools.product(range(4), repeat=3) if i + j + k <= 3]
X = np.stack([a**i * b**j * c**k for i, j, k in idx], 1)
w, *_ = np.linalg.lstsq(X, y, rcond=None)
res = y - X @ w
print("T_c / T_c0 range %.2f .. %.2f, fit rms %.3f, max %.3f" % (np.exp(y.min()), np.exp(y.max()), res.std(), np.abs(res).max()))
print("// monomials a^i b^j c^k, i + j + k <= 3, k fastest")
print("{" + ", ".join("%.6e" % v for v in w) + "}")
