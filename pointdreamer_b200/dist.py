"""Multi-GPU sharding of the path: one process per GPU (torchrun), independent units, ONE
all-gather at the end.

  * batch of shapes: shape i runs on rank i % world; finished atlases are all-gathered;
  * one shape, G GPUs: view v's diffusion chain runs on rank v % G with the noise-stream slot of
    chain v (chain0 = v), so the gathered views are bit-identical to a single-GPU run.
There is no per-step traffic; NCCL over NVLink on GPUs, gloo in the CPU tests.
"""
import torch
import torch.distributed as dist


def world():
    if dist.is_available() and dist.is_initialized():
        return dist.get_rank(), dist.get_world_size()
    return 0, 1


def shard_indices(n_items, rank=None, world_size=None):
    """Indices of the items (shapes or views) this rank owns: i % world == rank."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    return list(range(rank, n_items, world_size))


def gather_stacked(local, n_items):
    """All-gather per-rank stacks [n_local, ...] of round-robin-sharded items back into item
    order [n_items, ...] on every rank.  n_items must be divisible by the world size."""
    r, w = world()
    if w == 1:
        return local
    if n_items % w != 0:
        raise ValueError(f"{n_items} items do not shard evenly over {w} ranks")
    local = local.contiguous()
    out = torch.empty((w,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1, *local.shape[1:]), local)
    # out[rank][j] is item j*w + rank  ->  [n_local, w, ...] -> [n_items, ...]
    return out.transpose(0, 1).reshape((n_items,) + tuple(local.shape[1:])).contiguous()


def inpaint_views_sharded(inpainter, sparse_imgs, masks):
    """DDNM for one shape split by view across the ranks, then one all-gather of the views."""
    r, w = world()
    V = sparse_imgs.shape[0]
    if w == 1:
        return inpainter.inpaint_batch(sparse_imgs, masks, chain0=0)
    idx = shard_indices(V)
    outs = []
    for v in idx:  # chain0 = v keeps every chain on the noise slot the serial reference uses
        outs.append(inpainter.inpaint_batch(sparse_imgs[v:v + 1], masks[v:v + 1], chain0=v))
    return gather_stacked(torch.cat(outs, 0), V)
