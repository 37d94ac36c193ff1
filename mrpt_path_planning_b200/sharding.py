"""Multi-GPU partitioning of the hot path. One process per GPU, torch.distributed for the plumbing.

Three axes, only the third needs a collective:
  * node batches (config C3)      -> contiguous slice of the batch per rank, cloud replicated, no collective
  * independent queries (C4)      -> queries q with q % world == rank, cloud replicated, no collective
  * very large clouds (C5)        -> points i with i % world == rank; every rank evaluates ALL edges against its
                                     shard and the partial per-edge minima are merged with all_reduce(MIN). The
                                     per-edge value is an order-independent min over points (+inf when no point
                                     constrains the path), so the merged tensor equals the single-GPU result bit
                                     for bit.
The functions below are backend-agnostic (NCCL on the GPUs, gloo in the CPU tests).
"""
import numpy as np


def slice_for_rank(n, rank, world):
    """Contiguous, balanced slice [lo, hi) of n items for this rank (node batches)."""
    base, rem = divmod(n, world)
    lo = rank * base + min(rank, rem)
    return lo, lo + base + (1 if rank < rem else 0)


def queries_for_rank(n, rank, world):
    """Round-robin query ids of this rank (independent planning queries)."""
    return list(range(rank, n, world))


def cloud_shard(xs, ys, rank, world):
    """Points of this rank's shard of the obstacle cloud (strided by index: every shard covers the whole map,
    so the per-rank work stays balanced for any node placement)."""
    return np.ascontiguousarray(xs[rank::world]), np.ascontiguousarray(ys[rank::world])


def allreduce_min_(t, group=None):
    """In-place element-wise minimum over ranks of a torch tensor of partial per-edge minima."""
    import torch.distributed as dist

    if dist.is_available() and dist.is_initialized() and dist.get_world_size(group) > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MIN, group=group)
    return t


def attach_ranks(ctx, rank, world, max_nodes=0, group=None):
    """The hand-shake of the library's multi-GPU layer (include/mppb.h, "several GPUs") over torch.distributed: rank 0
    draws the NCCL unique id, every rank builds its communicator on the context's device; with max_nodes > 0 every
    rank also creates its receive slots for the fused merge and attaches to all the others' (CUDA IPC handles).
    After this, ctx.eval_nodes_sharded_device() evaluates against the rank's shard and merges on the context's stream;
    torch.distributed is not involved in the data path."""
    import torch.distributed as dist

    from . import capi

    box = [capi.comm_unique_id() if rank == 0 else None]
    dist.broadcast_object_list(box, src=0, group=group)
    ctx.comm_init(world, rank, box[0])
    if max_nodes > 0 and world > 1:
        handles = [None] * world
        dist.all_gather_object(handles, ctx.peer_rows_create(max_nodes), group=group)
        ctx.peer_rows_attach(world, rank, handles)
        dist.barrier(group=group)  # nobody evaluates before everybody is attached
    return ctx.sharded_mode


def eval_nodes_cloud_sharded(ctx, n_nodes, d_nodes_xycs, clip, d_out, group=None):
    """Stage (a)+(b) against this rank's cloud shard (already uploaded into ctx), then the min-merge, all inside the
    library (mppb_eval_nodes_sharded_device: fused peer-memory merge or NCCL, see attach_ranks).
    d_nodes_xycs / d_out are device arrays on the context's stream; d_out is [n_nodes][total_paths] float32."""
    ctx.eval_nodes_sharded_device(n_nodes, d_nodes_xycs, clip, d_out)
    return d_out


def gather_rows(local_rows, n_total, rank, world, group=None):
    """Rank-0 gather of node-sharded result rows (host side; used by tests and tools, not by the bench)."""
    import torch
    import torch.distributed as dist

    if world == 1:
        return local_rows
    cols = local_rows.shape[1]
    sizes = [slice_for_rank(n_total, r, world) for r in range(world)]
    bufs = [torch.empty((hi - lo, cols), dtype=local_rows.dtype) for lo, hi in sizes]
    dist.all_gather(bufs, local_rows.contiguous(), group=group) if len({hi - lo for lo, hi in sizes}) == 1 else \
        dist.all_gather_object(bufs, local_rows)
    return torch.cat([b if isinstance(b, torch.Tensor) else torch.as_tensor(b) for b in bufs], 0)
