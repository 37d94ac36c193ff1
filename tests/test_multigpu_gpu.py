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
        ctx.set_grid_near((1, 2, 0, 2, 2)[step])  # near-first form of the step (two launches share the epilogue's count) and plain
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


def test_queries_over_two_devices_in_one_process(capi, orc, oracle_ackermann):
    """mppb_planner_plan_multi: independent queries dealt to planners on two devices of this process, run at once;
    every result equals the single-planner batch and the oracle planner's tree."""
    import torch

    from mrpt_path_planning_b200 import workloads as wl
    from test_oracle_known_answers import C1_PARAMS, C1_TIMESTAMPS
    from test_planner_gpu import _random_queries

    if torch.cuda.device_count() < 2:
        pytest.skip("needs two CUDA devices")
    extent = 40.0
    xs, ys = wl.walls_cloud(6000, extent, wall_fraction=0.8, seed=33)
    qs = _random_queries(np.random.default_rng(35), 9, extent, 3.0, 8.0)
    params = list(C1_PARAMS)
    params[8] = 40
    ptgs = capi.ackermann_ptgs()
    ctxs = [capi.Context(d) for d in range(2)]
    planners = []
    for c in ctxs:
        pl = capi.Planner(c, ptgs, params, C1_TIMESTAMPS)
        pl.add_obstacles(xs, ys)
        pl.add_waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True)
        planners.append(pl)
    multi = capi.Planner.plan_multi(planners, qs, n_threads_per_planner=2)
    single = planners[0].plan_many(qs, n_threads=2)
    trees_single = [planners[0].tree(i) for i in range(len(qs))]
    multi = capi.Planner.plan_multi(planners, qs, n_threads_per_planner=2)  # (again: planners[0] now holds its share)
    for i, q in enumerate(qs):
        for key in ("success", "tree_nodes", "tree_parents", "expanded", "path_cost", "best_node", "path_len"):
            assert multi[i][key] == single[i][key], (i, key)
        assert np.array_equal(planners[i % 2].tree(i // 2), trees_single[i])
    ora = orc.Planner(oracle_ackermann, params, C1_TIMESTAMPS)
    ora.add_obstacles(xs, ys)
    ora.add_waypoints(orc.Waypoints([20.0, 10.0], [20.0, 30.0], 3.0, 0.5, True))
    ora.plan(*qs[3][:4])
    assert np.array_equal(planners[1].tree(1), ora.tree())
    for pl in planners:
        pl.close()
    for c in ctxs:
        c.close()
