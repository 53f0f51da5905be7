"""One-off extended parity check on a GPU box (not part of the test suite: ~1 minute of oracle time): N seeded random
(category, T, z, k_ij) states per run through the kernels and through the oracle (oracle/models.py + oracle/solver.py),
with seeds that differ from the golden grids'.  Prints, per property, the number of states compared, mask
disagreements (all / outside the near-critical exemption of tests/test_gpu_mix.py), and the largest value and
gradient deviations.

    python profiles/extended_parity.py [seed] [states per category]
"""
import os
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))
import cases  # noqa: E402
import feos_torch_b200 as ft  # noqa: E402
from oracle import eos as oeos, models as om, solver as osolver  # noqa: E402

F64 = torch.float64
seed = int(sys.argv[1]) if len(sys.argv) > 1 else 77
per = int(sys.argv[2]) if len(sys.argv) > 2 else 40
g = torch.Generator().manual_seed(seed)
P0 = torch.tensor(cases.MIX_PARAMS, dtype=F64)
ncat = P0.shape[0]
n = per * ncat
cat = torch.arange(n) % ncat
P = P0[cat]
K = torch.zeros((n, 2), dtype=F64)
K[:, 0] = -0.1 + 0.15 * torch.rand(n, generator=g, dtype=F64)
T = 150.0 + 140.0 * torch.rand(n, generator=g, dtype=F64)
z = 0.02 + 0.96 * torch.rand(n, generator=g, dtype=F64)
p0 = torch.full_like(T, 1e5)
print(f"seed {seed}: {n} states over {ncat} categories, T 150-290 K, z 0.02-0.98, k_ij -0.1..0.05")
for bubble in (True, False):
    ko = K.clone().requires_grad_(True)
    po = P.clone().requires_grad_(True)
    mo = om.PcSaftMix(po, ko)
    vo, no = (mo.bubble_point if bubble else mo.dew_point)(T, z, p0)
    vo.sum().backward()
    rho, _ = osolver.binary_bubble_dew(lambda T_, r_: oeos.mix_helmholtz(P, K, T_, r_), osolver.mix_max_density(P), T, z, p0, bubble)
    ratio_o = rho[:, 2:4].sum(dim=1) / rho[:, 0:2].sum(dim=1)
    kg = K.cuda().requires_grad_(True)
    pg = P.cuda().requires_grad_(True)
    mg = ft.PcSaftMix(pg, kg)
    vg, ng = (mg.bubble_point if bubble else mg.dew_point)(T.cuda(), z.cuda(), p0.cuda())
    vg.sum().backward()
    ng = ng.cpu()
    both = ~ng & ~no
    full_g = torch.full((n,), float("nan"), dtype=F64)
    full_g[~ng] = vg.detach().cpu()
    full_o = torch.full((n,), float("nan"), dtype=F64)
    full_o[~no] = vo.detach()
    dv = ((full_g[both] - full_o[both]).abs() / full_o[both].abs()).max().item()
    gk = ((kg.grad.cpu()[both, 0] - ko.grad[both, 0]).abs() / ko.grad[both, 0].abs()).max().item()
    gp_o = po.grad.reshape(n, 16)[both]
    gp_g = pg.grad.cpu().reshape(n, 16)[both]
    gp = ((gp_g - gp_o).abs() / gp_o.abs().max(dim=1, keepdim=True).values).max().item()
    diff = ng != no
    near = diff & ((~no & (ratio_o < 1.5)) | ng)  # oracle's solution near-critical, or only the kernels fail
    only_kernel_fails = int((diff & ng).sum())
    print(f"{'bubble' if bubble else 'dew   '}: compared {int(both.sum())}, oracle failed {int(no.sum())}, kernels failed {int(ng.sum())}, "
          f"mask differs on {int(diff.sum())} (oracle near-critical: {int((diff & ~no & (ratio_o < 1.5)).sum())}, "
          f"kernels fail where the oracle converges: {only_kernel_fails}); "
          f"max rel value {dv:.2e}, k_ij gradient {gk:.2e}, parameter gradients {gp:.2e}")
