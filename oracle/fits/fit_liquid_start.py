"""Fit of the liquid-density start correlation used by the forward kernels (csrc/liquid_start.cuh):
saturated liquid density of non-associating, non-polar PC-SAFT fluids, rho_L / rho_max = F(tau, 1/m) with
tau = (T / eps) / t_c(m), from saturated densities computed by the oracle's compiled CPU solver."""
import sys
import numpy as np, torch
sys.path.insert(0, "/root/repo")
from oracle import cpu_solver

ms = np.arange(1.0, 8.01, 0.125)
ts = np.arange(0.40, 4.2, 0.01)
M, TS = np.meshgrid(ms, ts, indexing="ij")
P = np.zeros((M.size, 8)); P[:, 0] = M.ravel(); P[:, 1] = 3.5; P[:, 2] = 200.0
T = TS.ravel() * 200.0
rv, rl, failed = cpu_solver.pure_vle(torch.from_numpy(P), torch.from_numpy(T))
ok = (~failed).numpy().reshape(M.shape)
rl = rl.numpy().reshape(M.shape); rv = rv.numpy().reshape(M.shape)
rho_max = 0.5 / (np.pi / 6 * M * 3.5**3)
# critical T* per m: largest converged T* (grid 0.01) -- refine by requiring rl/rv > 1.02
def tc_of(m):
    """critical T* by bisection on 'the solver finds two phases with rho_L / rho_V > 1.01'"""
    lo, hi = 0.9, 4.0
    for _ in range(30):
        mid = 0.5 * (lo + hi)
        p1 = torch.tensor([[m, 3.5, 200.0, 0, 0, 0, 0, 0]], dtype=torch.float64)
        a, b, f = cpu_solver.pure_vle(p1, torch.tensor([mid * 200.0], dtype=torch.float64))
        if (not f.item()) and (b / a).item() > 1.01: lo = mid
        else: hi = mid
    return lo
tc = np.array([tc_of(m) for m in ms])
print("tc(m) samples", list(zip(ms[::8], tc[::8])))
# fit tc(m) polynomial in m (deg 4)
ctc = np.polyfit(1.0 / ms, tc, 6)
print("tc fit max rel err", np.abs(np.polyval(ctc, 1.0 / ms) / tc - 1).max(), list(ctc))
tau = TS / np.polyval(ctc, 1.0 / M)
sel = ok & (tau > 0.28) & (tau < 0.985)
x, im, y = tau[sel], 1.0 / M[sel], (rl / rho_max)[sel]
terms = [(a, b) for a in range(0, 5) for b in range(0, 3)]
# sqrt(1 - tau) captures the critical scaling of the liquid branch
def feats(x, im):
    w = np.sqrt(np.maximum(1.0 - x, 0.0))
    return np.stack([w**a * im**b for a, b in terms], 1)
X = feats(x, im)
coef, *_ = np.linalg.lstsq(X, y, rcond=None)
err = X @ coef / y - 1
print("n", len(y), "F fit rel err min %.4f max %.4f rms %.4f" % (err.min(), err.max(), np.sqrt((err**2).mean())))
print("terms", terms)
print("coef", ", ".join("%.10e" % c for c in coef))
print("ctc", ", ".join("%.10e" % c for c in ctc))
for lo, hi in [(0.28,0.4),(0.4,0.6),(0.6,0.8),(0.8,0.9),(0.9,0.95),(0.95,0.97),(0.97,0.985)]:
    s = (x>=lo)&(x<hi)
    print("tau [%.2f,%.2f): n %d err min %.4f max %.4f" % (lo,hi,s.sum(), err[s].min(), err[s].max()))
# refit on tau <= 0.97
s = x <= 0.97
coef2, *_ = np.linalg.lstsq(X[s], y[s], rcond=None)
e2 = X[s] @ coef2 / y[s] - 1
print("refit tau<=0.97: min %.4f max %.4f rms %.4f" % (e2.min(), e2.max(), np.sqrt((e2**2).mean())))
print("coef2", ", ".join("%.10e" % c for c in coef2))
# check on the C2 workload
import bench
params, Tw, _ = bench.make_workload(100000, 1000, "cpu")
m = params[:,0].numpy(); tstar = (Tw/params[:,2]).numpy()
tauw = tstar/np.polyval(ctc, 1.0/m)
rvw, rlw, fw = cpu_solver.pure_vle(params, Tw)
okw = (~fw).numpy() & (tauw<=0.97) & (tauw>=0.28)
rm = 0.5/(np.pi/6*m*params[:,1].numpy()**3)
pred = feats(tauw[okw], 1/m[okw]) @ coef2
ew = pred/(rlw.numpy()[okw]/rm[okw]) - 1
print("C2 workload: covered %.3f of converged, err min %.4f max %.4f" % (okw.sum()/(~fw).numpy().sum(), ew.min(), ew.max()))
