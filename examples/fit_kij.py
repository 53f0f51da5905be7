"""Fit a binary interaction parameter k_ij to bubble-point pressures, data-parallel over the GPUs of
one node (the use case of the publication the reference cites, README.md:44-55).

    python examples/fit_kij.py                                   # one GPU
    python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 examples/fit_kij.py

Every rank owns a round-robin shard of the (T, x) states; the only communication is one all-reduce of
the k_ij gradient (2 doubles) per step (feos_torch_b200.distributed.fit_step)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from feos_torch_b200 import PcSaftMix  # noqa: E402
from feos_torch_b200 import distributed as fd  # noqa: E402


def main(n=200_000, steps=30):
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    if world > 1:
        dist.init_process_group("nccl", device_id=dev)

    g = torch.Generator().manual_seed(0)
    pair = torch.tensor([[1, 3.5, 150, 0, 0, 0, 0, 0], [1, 3.5, 200, 0, 0, 0, 0, 0]], dtype=torch.float64, device=dev)
    T_all = (0.6 + 0.25 * torch.rand(n, generator=g, dtype=torch.float64)) * 150 * 1.2757
    x_all = 0.05 + 0.9 * torch.rand(n, generator=g, dtype=torch.float64)
    T, x = fd.shard(T_all, rank, world).to(dev), fd.shard(x_all, rank, world).to(dev)
    k = T.shape[0]
    p0 = torch.full((k,), 1e5, dtype=torch.float64, device=dev)

    def pressures(kij):
        return PcSaftMix(pair.expand(k, 2, 8), kij.expand(k, 2)).bubble_point(T, x, p0)[0]

    with torch.no_grad():  # synthetic "experimental" data from k_ij = -0.12
        target = pressures(torch.tensor([-0.12, 0.0], dtype=torch.float64, device=dev))

    kij = torch.tensor([-0.02, 0.0], dtype=torch.float64, device=dev, requires_grad=True)
    opt = torch.optim.LBFGS([kij], lr=1.0, max_iter=1, history_size=5)
    for step in range(steps):
        def closure():
            return torch.tensor(fd.fit_step(lambda: ((pressures(kij) / target - 1.0) ** 2).sum(), [kij]), device=dev)

        loss = opt.step(closure)
        if rank == 0 and (step % 5 == 0 or step == steps - 1):
            print(f"step {step:3d}  loss {float(loss):.6e}  k_ij {kij[0].item():+.8f}")
    if rank == 0:
        assert abs(kij[0].item() + 0.12) < 1e-6, kij
        print("recovered k_ij = -0.12")
    if world > 1:
        dist.destroy_process_group()


if __name__ == "__main__":
    main()
