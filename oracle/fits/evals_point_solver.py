import sys, numpy as np
sys.path.insert(0, "/root/repo/tests/hostsim"); sys.path.insert(0, "/root/repo/tests")
import sim, json
rng = np.random.default_rng(5)
n = 4000
P = np.tile(np.array([1, 3.5, 150, 0, 0, 0, 0, 0, 1, 3.5, 200, 0, 0, 0, 0, 0], dtype=float), (n, 1))
K = np.tile(np.array([-0.1, 0.0]), (n, 1))
T = (0.6 + 0.25 * rng.random(n)) * 150 * 1.2757
z = 0.05 + 0.9 * rng.random(n)
for bubble in (True, False):
    rho, val, st, ev = sim.mix_point(P, K, T, z, bubble, robust=False)
    print("C4", "bubble" if bubble else "dew", "mean evals %.3f" % ev.mean(), "retry", int((st == 2).sum()), "fail", int((st == 1).sum()), "hist", np.bincount(ev)[:30].tolist())
# a second plain pair with chains and a wider range
P2 = np.tile(np.array([2.5, 3.6, 230, 0, 0, 0, 0, 0, 4.0, 3.9, 260, 0, 0, 0, 0, 0], dtype=float), (n, 1))
K2 = np.tile(np.array([0.02, 0.0]), (n, 1))
T2 = 250 + 250 * rng.random(n)
for bubble in (True, False):
    rho, val, st, ev = sim.mix_point(P2, K2, T2, z, bubble, robust=False)
    print("chains", "bubble" if bubble else "dew", "mean evals %.3f" % ev.mean(), "retry", int((st == 2).sum()), "fail", int((st == 1).sum()))
g = json.load(open("/root/repo/tests/golden/mix.json"))
for grid in ("category_grid", "random_grid"):
    e = g[grid]
    Pg = np.array(e["params"]).reshape(-1, 16)
    for bubble in (True, False):
        rho, val, st, ev = sim.mix_point(Pg, np.array(e["kij"]), np.array(e["T"]), np.array(e["z"]), bubble, robust=False)
        print(grid, "bubble" if bubble else "dew", "mean evals %.3f" % ev.mean(), "retry", int((st == 2).sum()), "fail", int((st == 1).sum()))
