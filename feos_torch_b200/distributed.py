"""Multi-GPU use of the hot path: one process per GPU (``torch.distributed``, NCCL over NVLink).

Every state point is an independent solve (the reference maps over states with rayon,
`/root/reference/src/pcsaft.rs:86-92`), so the batch axis is sharded with NO collective on the data
path.  States are dealt out round-robin (state i -> rank i mod world) rather than in contiguous blocks
because the cost of a state correlates with T/T_c and inputs are often sorted by temperature.

The only exchange is the sum of the gradients of SHARED leaves (a k_ij, segment parameters, one
parameter row expanded over the batch) of a fitting loss: a handful of doubles, one all-reduce per
step -- latency bound, independent of the batch size.  Per-state parameters need no collective.
"""
import torch
import torch.distributed as dist


def shard_indices(n, rank=None, world=None):
    """Indices of the states owned by ``rank``: rank, rank + world, rank + 2 world, ..."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return torch.arange(rank, n, world)


def shard(tensor, rank=None, world=None):
    """Round-robin shard of a per-state tensor along dim 0 (a view; call .contiguous() if needed)."""
    rank = dist.get_rank() if rank is None else rank
    world = dist.get_world_size() if world is None else world
    return tensor[rank::world]


def unshard_mask(local_masks, n):
    """Inverse of ``shard`` for the per-rank failed masks gathered in rank order -> mask of length n."""
    world = len(local_masks)
    out = torch.empty(n, dtype=local_masks[0].dtype, device=local_masks[0].device)
    for r, m in enumerate(local_masks):
        out[r::world] = m
    return out


def all_reduce_gradients(leaves, group=None):
    """Sum the ``.grad`` of shared leaves over all ranks with ONE all-reduce (flattened into a single
    buffer).  Leaves without a gradient contribute zeros, so every rank issues the same collective."""
    leaves = list(leaves)
    if not leaves or not dist.is_initialized() or dist.get_world_size(group) == 1:
        return
    flat = torch.cat([(p.grad if p.grad is not None else torch.zeros_like(p)).reshape(-1) for p in leaves])
    dist.all_reduce(flat, op=dist.ReduceOp.SUM, group=group)
    off = 0
    for p in leaves:
        k = p.numel()
        p.grad = flat[off:off + k].view_as(p).clone()
        off += k


def fit_step(loss_fn, shared_leaves, optimizer=None, group=None):
    """One data-parallel step of a fit: ``loss_fn()`` evaluates the loss over this rank's shard (through
    PcSaftPure / PcSaftMix / GcPcSaftMix), gradients of the shared leaves are summed over the ranks and
    the optimizer (if given) steps.  Returns the global loss as a float."""
    for p in shared_leaves:
        p.grad = None
    loss = loss_fn()
    loss.backward()
    all_reduce_gradients(shared_leaves, group)
    total = loss.detach().clone()
    if dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(total, op=dist.ReduceOp.SUM, group=group)
    if optimizer is not None:
        optimizer.step()
    return float(total)
