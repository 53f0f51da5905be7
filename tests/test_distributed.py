"""world_size-2 gloo tests (CPU) of the multi-GPU host logic: round-robin sharding of the batch and the
single all-reduce of the shared-leaf gradients of a fit loss.  The kernels are per-state independent, so
a cheap differentiable stand-in for the property evaluation exercises exactly the same plumbing."""
import os

import torch
import torch.distributed as dist
import torch.multiprocessing as mp

from feos_torch_b200 import distributed as fd


def _worker(rank, world, port, n, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        g = torch.Generator().manual_seed(0)
        T = torch.rand(n, generator=g, dtype=torch.float64) + 1.0
        target = torch.rand(n, generator=g, dtype=torch.float64)
        k = torch.tensor([0.3, -0.2], dtype=torch.float64, requires_grad=True)  # shared leaf (like k_ij)
        idx = fd.shard_indices(n)
        assert torch.equal(fd.shard(T), T[idx])

        def loss_fn():
            Tl, tl = fd.shard(T), fd.shard(target)
            pred = k[0] * Tl + k[1] * Tl * Tl  # stand-in for a property with a shared parameter
            return ((pred - tl) ** 2).sum()

        total = fd.fit_step(loss_fn, [k])
        mask = (fd.shard(T) > 1.5)
        gathered = [torch.empty_like(fd.shard(T, r, world) > 1.5) for r in range(world)]
        dist.all_gather(gathered, mask)
        full_mask = fd.unshard_mask(gathered, n)
        if rank == 0:
            out.put((total, k.grad.tolist(), full_mask.tolist()))
    finally:
        dist.destroy_process_group()


def test_sharding_and_gradient_allreduce_world2():
    n, world, port = 1000, 2, 29517  # equal shards: all_gather needs equal sizes
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    procs = [ctx.Process(target=_worker, args=(r, world, port, n, q)) for r in range(world)]
    for p in procs:
        p.start()
    total, grad, mask = q.get(timeout=120)
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    # single-process reference
    g = torch.Generator().manual_seed(0)
    T = torch.rand(n, generator=g, dtype=torch.float64) + 1.0
    target = torch.rand(n, generator=g, dtype=torch.float64)
    k = torch.tensor([0.3, -0.2], dtype=torch.float64, requires_grad=True)
    loss = ((k[0] * T + k[1] * T * T - target) ** 2).sum()
    loss.backward()
    assert abs(total - loss.item()) < 1e-9 * abs(loss.item())
    assert torch.allclose(torch.tensor(grad, dtype=torch.float64), k.grad, rtol=1e-12)
    assert mask == (T > 1.5).tolist()


def test_shard_indices_cover_batch_once():
    n = 37
    for world in (1, 2, 4, 8):
        seen = torch.cat([fd.shard_indices(n, r, world) for r in range(world)])
        assert sorted(seen.tolist()) == list(range(n))
