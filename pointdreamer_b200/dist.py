"""Multi-GPU sharding of the path: one process per GPU (torchrun), independent units, ONE
all-gather at the end.

  * batch of shapes: shape i runs on rank i % world; finished atlases are all-gathered;
  * one shape, G GPUs: rank r runs views r*V/G .. (r+1)*V/G-1 as one U-Net batch, each chain on
    the noise-stream slot of its view (chain0 = first view), so the gathered views are
    bit-identical to a single-GPU run.
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


def gather_blocks(local):
    """All-gather per-rank stacks [n_local, ...] of BLOCK-sharded items (rank r owns items
    r*n_local .. (r+1)*n_local-1) into item order on every rank."""
    r, w = world()
    if w == 1:
        return local
    local = local.contiguous()
    out = torch.empty((w * local.shape[0],) + tuple(local.shape[1:]), dtype=local.dtype,
                      device=local.device)
    dist.all_gather_into_tensor(out, local)
    return out


def view_block(n_views, rank=None, world_size=None):
    """(first view, count) of the contiguous block of views this rank owns."""
    r, w = world()
    rank = r if rank is None else rank
    world_size = w if world_size is None else world_size
    if n_views % world_size != 0:
        raise ValueError(f"{n_views} views do not shard evenly over {world_size} ranks")
    n = n_views // world_size
    return rank * n, n


def inpaint_views_sharded(inpainter, sparse_imgs, masks):
    """DDNM for ONE shape split by view across the ranks: rank r runs the contiguous
    block of V/G chains it owns as ONE U-Net batch, with chain0 = its first view so every chain
    keeps the slot of the reference's noise stream (diffusion.py:493-499, 552); then one
    all-gather of the views.  The result equals the single-GPU run bit for bit."""
    r, w = world()
    V = sparse_imgs.shape[0]
    if w == 1:
        return inpainter.inpaint_batch(sparse_imgs, masks, chain0=0)
    v0, n = view_block(V)
    local = inpainter.inpaint_batch(sparse_imgs[v0:v0 + n].contiguous(),
                                    masks[v0:v0 + n].contiguous(), chain0=v0)
    return gather_blocks(local)
