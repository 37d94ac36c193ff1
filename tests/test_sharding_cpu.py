"""Multi-process (gloo, world_size 2, CPU) tests of the N>1 host logic: shard assignment and the min-merge of a
cloud-sharded evaluation. The per-shard evaluator here is the oracle (a checker), standing in for the GPU kernel,
because this test runs without a GPU; the merge code under test is the product's sharding.py."""
import os
import socket
import sys

import numpy as np
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

from mrpt_path_planning_b200 import sharding, workloads as wl  # noqa: E402


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    p = s.getsockname()[1]
    s.close()
    return p


def _worker(rank, world, port, tmp):
    os.environ.update(MASTER_ADDR="127.0.0.1", MASTER_PORT=str(port))
    dist.init_process_group("gloo", rank=rank, world_size=world)
    from oracle import orc

    ptgs = orc.ackermann_ptgs()
    xs, ys = wl.uniform_cloud(4000, 30.0, seed=41)
    poses = wl.node_poses(24, 30.0, margin=2.0, seed=42)
    # (iii) cloud sharding: partial minima per rank, merged with all_reduce(MIN)
    sx, sy = sharding.cloud_shard(xs, ys, rank, world)
    part, _ = orc.eval_nodes(ptgs, sx, sy, poses, 5.0, mode=1, nthreads=1)
    t = torch.from_numpy(part.astype(np.float32))
    sharding.allreduce_min_(t)
    # (i) node sharding: each rank evaluates its slice against the whole cloud
    lo, hi = sharding.slice_for_rank(len(poses), rank, world)
    rows, _ = orc.eval_nodes(ptgs, xs, ys, poses[lo:hi], 5.0, mode=1, nthreads=1)
    full = sharding.gather_rows(torch.from_numpy(rows.astype(np.float32)), len(poses), rank, world)
    if rank == 0:
        np.save(os.path.join(tmp, "merged.npy"), t.numpy())
        np.save(os.path.join(tmp, "gathered.npy"), full.numpy())
    dist.barrier()
    dist.destroy_process_group()


def test_cloud_sharded_min_merge_and_node_sharding_world2(tmp_path, orc, oracle_ackermann):
    world = 2
    mp.spawn(_worker, args=(world, _free_port(), str(tmp_path)), nprocs=world, join=True)
    xs, ys = wl.uniform_cloud(4000, 30.0, seed=41)
    poses = wl.node_poses(24, 30.0, margin=2.0, seed=42)
    ref, _ = orc.eval_nodes(oracle_ackermann, xs, ys, poses, 5.0, mode=1, nthreads=0)
    ref = ref.astype(np.float32)
    assert np.array_equal(np.load(tmp_path / "merged.npy"), ref)
    assert np.array_equal(np.load(tmp_path / "gathered.npy"), ref)


def test_shard_assignment_covers_everything_once():
    for n in (0, 1, 7, 4096, 4099):
        for world in (1, 2, 3, 8):
            seen = []
            for r in range(world):
                lo, hi = sharding.slice_for_rank(n, r, world)
                seen += list(range(lo, hi))
            assert seen == list(range(n))
            q = sorted(sum((sharding.queries_for_rank(n, r, world) for r in range(world)), []))
            assert q == list(range(n))
    xs = np.arange(10, dtype=np.float32)
    parts = [sharding.cloud_shard(xs, xs, r, 3)[0] for r in range(3)]
    assert sorted(np.concatenate(parts).tolist()) == xs.tolist()
