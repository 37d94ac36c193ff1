"""Two GPUs, one process each: the library's multi-GPU layer (include/mppb.h, "several GPUs"). The cloud is sharded by
point index, every rank evaluates all nodes against its shard, and the partial rows are merged by ncclAllReduce(min)
and by the fused peer-memory epilogue of tpobs_grid_kernel. Both must equal the single-GPU evaluation of the whole
cloud bit for bit (the per-path value is an order-independent minimum over points) and the CPU checker.
Skipped when fewer than two CUDA devices are visible."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

pytestmark = pytest.mark.gpu


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _inputs():
    from mrpt_path_planning_b200 import workloads as wl

    xs, ys = wl.uniform_cloud(200000, 60.0, seed=71)
    poses = wl.node_poses(300, 60.0, margin=3.0, seed=72)
    return xs, ys, poses


def _worker(rank, world, port, tmp):
    import torch
    import torch.distributed as dist

    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)  # carries the hand-shake bytes only
    from mrpt_path_planning_b200 import capi, sharding

    torch.cuda.set_device(rank)
    ctx = capi.Context(rank)
    for p in capi.ackermann_ptgs():
        ctx.add_ptg(p)
    xs, ys, poses = _inputs()
    ctx.set_obstacles_shard(xs, ys, rank, world)
    n, cols = len(poses), ctx.total_paths
    d_nodes = torch.from_numpy(capi.nodes_xycs_from_poses(poses)).cuda(rank)
    d_out = torch.empty((n, cols), dtype=torch.float32, device="cuda:%d" % rank)
    torch.cuda.synchronize()
    out = {}
    assert sharding.attach_ranks(ctx, rank, world) == "nccl"
    ctx.eval_nodes_sharded_device(n, d_nodes, 5.0, d_out)
    ctx.sync()
    out["nccl"] = d_out.cpu().numpy().copy()
    assert sharding.attach_ranks(ctx, rank, world, max_nodes=n) == "fused"
    for step in range(5):  # several epochs: both parity halves of the receive area, flags, counters
        d_out.zero_()
        torch.cuda.synchronize()
        ctx.eval_nodes_sharded_device(n - 7 * step, d_nodes, 5.0, d_out)
        ctx.sync()
        out["fused%d" % step] = d_out.cpu().numpy()[:n - 7 * step].copy()
    np.savez(os.path.join(tmp, "rank%d.npz" % rank), **out)
    dist.barrier()
    ctx.close()
    dist.destroy_process_group()


def test_cloud_sharded_over_two_gpus_equals_one_gpu(tmp_path, capi, orc, oracle_ackermann):
    import torch
    import torch.multiprocessing as mp

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    xs, ys, poses = _inputs()
    ctx = capi.Context(0)
    for p in capi.ackermann_ptgs():
        ctx.add_ptg(p)
    ctx.set_obstacles(xs, ys)
    whole = ctx.eval_nodes(poses, 5.0)
    ctx.close()
    for rank in range(world):
        r = np.load(tmp_path / ("rank%d.npz" % rank))
        assert np.array_equal(r["nccl"], whole)
        for step in range(5):
            assert np.array_equal(r["fused%d" % step], whole[:len(poses) - 7 * step]), (rank, step)
    ref, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses[:32], 5.0, mode=1, nthreads=0)
    init = np.concatenate([[p.init_dist(k) for k in range(p.num_paths)] for p in capi.ackermann_ptgs()])
    assert np.array_equal(np.minimum(init[None, :], whole[:32].astype(np.float64)), ref)
