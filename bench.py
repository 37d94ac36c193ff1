#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: candidate edges evaluated / s).

Workload (config C3 of BASELINE.json / SURVEY.md 8d): synthetic 2^20-point obstacle cloud
U[0,200)^2 (PCG64 seed 20240601), a 4096-node expansion batch (seed 20240602 + rank) x both
Ackermann PTGs x all 121 trajectories = 991 232 candidate edges per step and per GPU.
A step = one pass of stage (a)+(b) over one node batch against the resident cloud.

  python bench.py --gpus N --steps K --warmup W            (ours; under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the reference's own code, oracle/_ref, on host cores)

Prints ONE JSON line (rank 0). `value` = device-timed throughput with inputs resident in HBM;
`e2e` = the same metric through the host C-ABI call (pinned host buffers, H2D + kernel + D2H and
the host-side cos/sin inside the timed region).
"""
import argparse
import json
import os
import statistics
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)

from mrpt_path_planning_b200 import workloads as wl  # noqa: E402

EXTENT = 200.0
CLIP = 5.0
N_PTGS, K_PATHS = 2, 121
METRIC = "candidate edges evaluated/sec (node x PTG x k, N obstacle pts)"
UNIT = "edges/s"


def algorithmic_work(b, n):
    """reference-equivalent FLOPs (every node tests every point) and the binned figure."""
    n_in = n * (2 * CLIP) ** 2 / (EXTENT * EXTENT)
    flop_ref = b * n * 6.0 + b * n_in * (8 + 6 * N_PTGS)
    flop_binned = b * n_in * (6 + 8 + 6 * N_PTGS)
    hbm_bytes = 8.0 * n + 48.0 * b + 4.0 * b * N_PTGS * K_PATHS
    return dict(n_in=n_in, flop_ref=flop_ref, flop_binned=flop_binned, hbm_bytes=hbm_bytes)


class ClockSampler:
    """Samples nvidia-smi clocks / throttle reasons during the timed region."""

    Q = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
         "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, index):
        self.index = index
        self.proc = None
        self.lines = []

    def start(self):
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "-i", str(self.index), "--query-gpu=" + self.Q, "--format=csv,noheader,nounits",
                 "-lms", "100"], stdout=subprocess.PIPE, stderr=subprocess.DEVNULL, text=True)
            self.t = threading.Thread(target=self._read, daemon=True)
            self.t.start()
        except Exception:
            self.proc = None

    def _read(self):
        for line in self.proc.stdout:
            self.lines.append(line.strip())

    def stop(self):
        if not self.proc:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvidia-smi unavailable"]}
        time.sleep(0.15)
        self.proc.terminate()
        try:
            self.proc.wait(timeout=2)
        except Exception:
            self.proc.kill()
        sm, mx, reasons = [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for ln in self.lines:
            f = [x.strip() for x in ln.split(",")]
            if len(f) < 9:
                continue
            try:
                sm.append(float(f[1]))
                mx.append(float(f[2]))
            except ValueError:
                continue
            for name, v in zip(names, f[5:9]):
                if v.lower() == "active":
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "samples": len(sm), "reasons": sorted(reasons)}


def bind_to_gpu_numa_node(index, local_world=1):
    """One process per GPU: pin this process to the CPUs next to its GPU (NVML's ideal affinity), so that the pinned
    host buffers the kernels write into over PCIe are allocated on the GPU's own NUMA node; when several ranks share
    those CPUs, each takes a disjoint slice of them (its host threads — the per-call cos/sin of the e2e path, the
    planner's workers — then never compete with another rank's for a core)."""
    try:
        import pynvml

        pynvml.nvmlInit()
        h = pynvml.nvmlDeviceGetHandleByIndex(index)
        pynvml.nvmlDeviceSetCpuAffinity(h)
    except Exception:
        pass
    try:
        cpus = sorted(os.sched_getaffinity(0))
        if local_world > 1 and len(cpus) >= local_world:
            k = len(cpus) // local_world
            mine = cpus[(index % local_world) * k:(index % local_world + 1) * k]
            os.sched_setaffinity(0, mine)
        return len(os.sched_getaffinity(0))
    except Exception:
        return None


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_checker():
    """(module, kind): the reference's own translation units (oracle/_ref/libmpp_ref.so, "reference") when the
    prebuilt library is present, else the restatement (oracle/orc_*.hpp, "port"). Same interface either way."""
    from oracle import orc, ref

    if ref.available():
        return ref.module(), "reference"
    return orc, "port"


def cpu_reference_rate(xs, ys, poses, budget_s, nthreads=0):
    """Times the reference call pattern (TPS_Astar::cached_local_obstacles -> transform_pc_square_clipping, then
    tp_obstacles_single_path per (PTG,k)) on a bounded sample of the node batch. Returns
    (edges_per_s, nodes_done, seconds, threads, outputs, kind)."""
    orc, kind = cpu_checker()

    optgs = orc.ackermann_ptgs()
    threads = orc.lib().orc_num_threads() if nthreads <= 0 else nthreads
    chunk = max(threads * 2, 8)
    done, secs, outs = 0, 0.0, []
    # calibration chunk (also warms the page cache / allocator)
    orc.eval_nodes(optgs, xs, ys, poses[:threads], CLIP, mode=0, nthreads=threads)
    while done < len(poses) and secs < budget_s:
        sl = poses[done:done + chunk]
        out, t = orc.eval_nodes(optgs, xs, ys, sl, CLIP, mode=0, nthreads=threads)
        outs.append(out)
        done += len(sl)
        secs += t
    return done * N_PTGS * K_PATHS / secs, done, secs, threads, np.concatenate(outs, 0), kind


def run_reference(args, rank, world):
    if rank != 0:
        return
    xs, ys = wl.uniform_cloud(args.points, EXTENT)
    poses = wl.node_poses(args.nodes, EXTENT)
    orc, kind = cpu_checker()

    optgs = orc.ackermann_ptgs()
    threads = orc.lib().orc_num_threads()
    # size the per-step sample so that (steps + warmup) samples finish in ~2 minutes
    _, t_probe = orc.eval_nodes(optgs, xs, ys, poses[:threads], CLIP, mode=0, nthreads=threads)
    per_node = max(t_probe / threads, 1e-6)  # wall seconds per node with all threads busy
    total_steps = args.steps + args.warmup
    sample = int(max(threads, min(args.nodes, 120.0 / (per_node * total_steps))))
    times = []
    for s in range(total_steps):
        lo = (s * sample) % max(1, args.nodes - sample + 1)
        _, t = orc.eval_nodes(optgs, xs, ys, poses[lo:lo + sample], CLIP, mode=0, nthreads=threads)
        if s >= args.warmup:
            times.append(t)
    ms = 1e3 * sum(times) / len(times)
    value = sample * N_PTGS * K_PATHS / (ms * 1e-3)
    desc = "%d of %d nodes per step vs the full %d-point cloud, %d host threads" % (
        sample, args.nodes, args.points, threads)
    emit(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": kind, "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


TRAFFIC_FILE = "r2z_traffic.json"  # ncu counters of the step's launches of tpobs_grid_kernel (dram bytes, warp instructions)
C4_COSTMAP = (0.10, 0.5, 4.0)  # resolution, preferredClearanceDistance, maxCost
C4_CPU_SAMPLE_CAP = 100
WAYPOINTS01 = ([3.0, 5.0, 9.0, 20.0, 26.0], [2.0, 5.0, 8.0, 7.0, 3.0])  # share/mvsim-demo-waypoints01.yaml targets


def demo_inputs():
    """Configs C1/C2: the reference's README command line (obstacles_01.txt, start [0.5 0 0], goal [4 2.5 45])."""
    obs = np.loadtxt(os.path.join(ROOT, "tests", "golden", "obstacles_01.txt")).astype(np.float32)
    xs, ys = obs[:, 0].copy(), obs[:, 1].copy()
    start, goal = [0.5, 0, 0, 0, 0, 0], [4, 2.5, np.deg2rad(45.0), 0, 0, 0]
    lo = np.minimum(np.minimum(obs.min(0), np.float32([0.5 - 2, 0 - 2])), np.float32([4 - 2, 2.5 - 2]))
    hi = np.maximum(np.maximum(obs.max(0), np.float32([0.5 + 2, 0 + 2])), np.float32([4 + 2, 2.5 + 2]))
    return xs, ys, start, goal, [float(lo[0]), float(lo[1]), -np.pi], [float(hi[0]), float(hi[1]), np.pi]


def planner_bench(ctx_factory, rank, world, args, with_cpu):
    """Planner-level numbers of BASELINE.json's metric: TPS_Astar plan ms on C1/C2 (single query, latency) and
    plans/s on a C4-shaped batch (independent queries in lock-step, sharded by query id across ranks)."""
    from mrpt_path_planning_b200 import capi

    out = {}
    xs, ys, start, goal, lo, hi = demo_inputs()
    ctx = ctx_factory()
    if with_cpu:
        orc, cpu_kind = cpu_checker()
        out["cpu_kind"] = cpu_kind
    for name, mk_ptgs, wps in (("c1_holonomic", capi.holonomic_ptgs, False), ("c2_ackermann", capi.ackermann_ptgs, True)):
        ptgs = mk_ptgs()
        pl = capi.Planner(ctx, ptgs, wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
        pl.add_obstacles(xs, ys)
        t0 = time.perf_counter()
        pl.add_costmap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
        t_cm = time.perf_counter() - t0
        if wps:
            pl.add_waypoints(WAYPOINTS01[0], WAYPOINTS01[1], 1.5, 0.5, True)
        pl.plan(start, goal, lo, hi)  # warm-up: staging buffers, table upload
        ts = []
        for _ in range(5):
            r = pl.plan(start, goal, lo, hi)
            ts.append(r["time_s"])
        st = pl.gpu_stats()
        d = {"plan_ms": 1e3 * min(ts), "success": bool(r["success"]), "nodes_expanded": int(r["expanded"]),
             "tps_points": int(r["tps_points"]), "edges_costed": int(r["edges_costed"]),
             "gpu_batches": st["collision_batches"] + st["cost_batches"], "gpu_ms": 1e3 * st["gpu_seconds"],
             "phase_ms": {k: 1e3 * v for k, v in st["phase_seconds"].items()}, "costmap_build_ms": 1e3 * t_cm,
             "phase_note": "B_collision_gpu and gpu_ms include the host's second half of phase A, which runs while the "
                           "collision kernels do (split entry points + mppb_sync)"}
        if with_cpu:
            optgs = orc.holonomic_ptgs() if name.startswith("c1") else orc.ackermann_ptgs()
            opl = orc.Planner(optgs, wl.DEMO_ASTAR_PARAMS, wl.DEMO_ASTAR_TIMESTAMPS)
            opl.add_obstacles(xs, ys)
            t0 = time.perf_counter()
            ocm = orc.CostMap(xs, ys, 0.04, 0.5, 4.0, False, 15.0, start[:3])
            d["cpu_costmap_build_ms"] = 1e3 * (time.perf_counter() - t0)
            opl.add_costmap(ocm)
            if wps:
                opl.add_waypoints(orc.Waypoints(WAYPOINTS01[0], WAYPOINTS01[1], 1.5, 0.5, True))
            ro = opl.plan(start, goal, lo, hi)
            d["cpu_plan_ms"] = 1e3 * ro["time_s"]
            d["identical_tree"] = bool(np.array_equal(pl.tree(), opl.tree()))
        pl.close()
        out[name] = d

    # C4: 4096 independent queries in total (strong scaling: 4096 / N per rank, by query id) on a synthetic 200x200 m
    # map of 2^20 points (round pillars: workloads.pillars_map), demo planner parameters, the obstacle costmap as cost
    # evaluator (share/costmap-obstacles.yaml values at 0.10 m cells: one map for the whole world), per-query world box
    # = start/goal +- 4 m, an expansion cap instead of a time limit
    nq_total = args.queries
    mx, my, pillars = wl.pillars_map(args.points, EXTENT)
    allq = wl.planning_queries_free(nq_total, pillars, EXTENT)
    mine = [allq[i] for i in range(rank, nq_total, world)]
    params = list(wl.DEMO_ASTAR_PARAMS)
    params[8] = args.c4_max_expansions
    ptgs = capi.ackermann_ptgs()
    pl = capi.Planner(ctx, ptgs, params, wl.DEMO_ASTAR_TIMESTAMPS)
    pl.add_obstacles(mx, my)
    t0 = time.perf_counter()
    pl.add_costmap(mx, my, C4_COSTMAP[0], C4_COSTMAP[1], C4_COSTMAP[2], False, 0.0, None)
    t_cm = time.perf_counter() - t0
    cpus_total = os.cpu_count() or 1
    cpus = cpus_total
    try:
        import torch

        gpus_on_box = max(world, torch.cuda.device_count())
    except Exception:
        gpus_on_box = world
    runs = {}
    # (a) every host thread this rank may use; (b) FOUR host threads per rank at every N (what a rank has on a 32-core
    # box with 8 GPUs): the figure to compare across N, "equal host threads per rank". The fixed-thread run takes every
    # fourth query of the job (1024 in total at every N), so that the N=1 run stays within seconds.
    avail = len(os.sched_getaffinity(0)) if hasattr(os, "sched_getaffinity") else max(1, cpus // world)
    for label, n_threads, qsel in (("all_threads", max(1, min(avail, cpus // world)), mine),
                                   ("fixed_threads", max(1, min(avail, 4)), [allq[i] for i in range(4 * rank, nq_total, 4 * world)])):
        pl.plan_many(mine[:16], n_threads=n_threads)  # warm-up
        t0 = time.perf_counter()
        res = pl.plan_many(qsel, n_threads=n_threads)
        t_batch = time.perf_counter() - t0
        st = pl.gpu_stats()
        runs[label] = {"host_threads": n_threads, "batch_s": t_batch, "queries": len(qsel),
                       "expansions": int(sum(r["expanded"] for r in res)), "successes": int(sum(r["success"] for r in res)),
                       "gpu_batches": st["collision_batches"] + st["cost_batches"], "gpu_s": st["gpu_seconds"],
                       "phase_seconds": st["phase_seconds"]}
    d = dict(runs["all_threads"])
    d.update({"queries_total": nq_total, "queries_this_rank": len(mine), "max_expansions": args.c4_max_expansions,
              "map": "workloads.pillars_map: %d points on %d round pillars (r 0.5-1.5 m), 200x200 m" % (len(mx), len(pillars)),
              "cost_evaluators": "CostEvaluatorCostMap %.2f m cells, clearance %.1f m, maxCost %.0f (whole map, %d s build)" % (
                  C4_COSTMAP[0], C4_COSTMAP[1], C4_COSTMAP[2], round(t_cm)),
              "costmap_build_s": t_cm, "scaling": "strong (queries_total fixed)", "fixed_threads": runs["fixed_threads"],
              "tree_nodes": int(sum(r["tree_nodes"] for r in res)),
              "max_rss_gb": __import__("resource").getrusage(__import__("resource").RUSAGE_SELF).ru_maxrss / 1048576.0})
    if with_cpu:
        # bounded CPU sample: the first few queries through the reference planner, one per host thread, at a small
        # expansion cap (the reference scans the whole 2^20-point cloud for every expanded node: ~6 ms per expansion)
        # and with the costmap limited to 40 m around the first start (the reference's builder takes a kd-tree query
        # per cell); the product plans the same queries with the same settings, and the trees must be identical
        from concurrent.futures import ThreadPoolExecutor

        n_cpu = min(4, cpus)
        sparams = list(params)
        sparams[8] = C4_CPU_SAMPLE_CAP
        # (costmap of the CPU sample: 12 m around the first start, from the points within reach of its cells — the
        # checker's nearest-point search is a plain scan, so the full map is out of its reach; cells and costs are
        # those the full cloud would give inside that window)
        s0 = mine[0][0]
        near = (np.abs(mx - s0[0]) < 12.0 + 1.5) & (np.abs(my - s0[1]) < 12.0 + 1.5)
        cmx, cmy = np.ascontiguousarray(mx[near]), np.ascontiguousarray(my[near])
        if len(cmx) == 0:
            cmx, cmy = mx[:1].copy(), my[:1].copy()
        cm_args = (C4_COSTMAP[0], C4_COSTMAP[1], C4_COSTMAP[2], False, 12.0, s0[:3])
        t0 = time.perf_counter()
        ocm = orc.CostMap(cmx, cmy, *cm_args)
        d["cpu_costmap_build_s"] = time.perf_counter() - t0
        optgs_pool = [orc.ackermann_ptgs() for _ in range(n_cpu)]

        def one(i):
            opl = orc.Planner(optgs_pool[i], sparams, wl.DEMO_ASTAR_TIMESTAMPS)
            opl.add_obstacles(mx, my)
            opl.add_costmap(ocm)
            return opl.plan(*mine[i][:4]), opl.tree()

        t0 = time.perf_counter()
        with ThreadPoolExecutor(max_workers=n_cpu) as ex:
            cpu_res = list(ex.map(one, range(n_cpu)))
        t_cpu = time.perf_counter() - t0
        pl2 = capi.Planner(ctx, ptgs, sparams, wl.DEMO_ASTAR_TIMESTAMPS)
        pl2.add_obstacles(mx, my)
        pl2.add_costmap(cmx, cmy, *cm_args)
        pl2.plan_many(mine[:n_cpu], n_threads=n_cpu)
        d["cpu_sample"] = {"queries": n_cpu, "threads": n_cpu, "max_expansions": C4_CPU_SAMPLE_CAP,
                           "costmap": "as above, limited to 12 m around the first start",
                           "expansions_per_s": sum(r["expanded"] for r, _ in cpu_res) / t_cpu, "seconds": t_cpu}
        d["identical_trees_on_sample"] = bool(all(np.array_equal(pl2.tree(i), t) for i, (_, t) in enumerate(cpu_res)))
        pl2.close()
    pl.close()
    ctx.close()
    out["c4_batch"] = d
    return out


def config_dict(args, **extra):
    c = {"workload": "C3: synthetic 1M-point cloud U[0,200)^2, 4096-node batch x 2 Ackermann PTGs x 121 k, clip 5 m",
         "points": args.points, "nodes_per_gpu": args.nodes, "ptgs": N_PTGS, "paths_per_ptg": K_PATHS,
         "edges_per_step_per_gpu": args.nodes * N_PTGS * K_PATHS,
         "l2": "flushed between timed steps (256 MiB write)", "parallelism": "nodes sharded per GPU, cloud replicated"}
    c.update(extra)
    return c


def cloud_sharded_bench(ctx, dev, stream, rank, world, args, d_nodes, d_out, flush):
    """Config C5 in weak-scaling form: a cloud of world x 2^23 points, sharded by point index; every rank evaluates the
    SAME 4096-node batch against its shard and the per-edge minima are merged inside the library
    (mppb_eval_nodes_sharded_device): by the fused peer-memory epilogue of tpobs_grid_kernel, and, for comparison, by
    ncclAllReduce(min). Rank 0 checks the merged rows against an evaluation of the whole cloud on one GPU (all nodes,
    bit for bit) and against the reference's own CPU code on the whole cloud (a node sample)."""
    import torch
    import torch.distributed as dist

    from mrpt_path_planning_b200 import capi, sharding

    n_total = args.c5_points_per_gpu * world
    xs, ys = wl.uniform_cloud(n_total, EXTENT, seed=wl.C5_CLOUD_SEED)
    ctx.set_obstacles_shard(xs, ys, rank, world)
    poses = wl.node_poses(args.nodes, EXTENT, seed=wl.C3_NODES_SEED)  # same nodes on every rank
    with torch.cuda.stream(stream):
        d_nodes.copy_(torch.from_numpy(capi.nodes_xycs_from_poses(poses)).to(dev))

    def steps(n, fn):
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(n):
                flush.zero_()
                a, b = (torch.cuda.Event(enable_timing=True) for _ in range(2))
                a.record(stream)
                fn()
                b.record(stream)
                evs.append((a, b))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs]

    def timed(fn):
        steps(3, fn)
        dist.barrier()
        ts = steps(max(5, args.steps // 2), fn)
        t = torch.tensor([sum(ts) / len(ts)], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    res = {"points_total": n_total, "points_per_gpu": int(ctx.num_obstacles), "nodes": args.nodes,
           "merge_bytes": int(d_out.numel() * 4), "scaling": "weak in cloud size (edges fixed)"}
    # (1) the shard evaluation alone, (2) + ncclAllReduce(min) by the library, (3) the fused peer-memory merge
    res["kernel_only_ms"] = timed(lambda: ctx.eval_nodes_device(args.nodes, d_nodes, CLIP, d_out))
    sharding.attach_ranks(ctx, rank, world, max_nodes=0)
    assert ctx.sharded_mode == "nccl"
    res["nccl_ms_per_step"] = timed(lambda: ctx.eval_nodes_sharded_device(args.nodes, d_nodes, CLIP, d_out))
    torch.cuda.synchronize(dev)
    merged_nccl = d_out.cpu().numpy().copy() if rank == 0 else None
    sharding.attach_ranks(ctx, rank, world, max_nodes=args.nodes)
    assert ctx.sharded_mode == "fused"
    res["fused_ms_per_step"] = timed(lambda: ctx.eval_nodes_sharded_device(args.nodes, d_nodes, CLIP, d_out))
    ctx.sync()
    res["ms_per_step"] = min(res["fused_ms_per_step"], res["nccl_ms_per_step"])
    res["merge_share_nccl"] = 1.0 - res["kernel_only_ms"] / res["nccl_ms_per_step"]
    res["merge_share_fused"] = 1.0 - res["kernel_only_ms"] / res["fused_ms_per_step"]
    edges = args.nodes * N_PTGS * K_PATHS
    res["edges_per_s"] = edges / (res["ms_per_step"] * 1e-3)
    if rank == 0:
        merged = d_out.cpu().numpy()
        res["fused_equals_nccl"] = bool(np.array_equal(merged, merged_nccl))
        # the whole cloud on ONE GPU (a second context on this rank's device): must give the merged rows bit for bit
        one = capi.Context(dev.index)
        for p in capi.ackermann_ptgs(ctx=one):
            one.add_ptg(p)
        one.set_obstacles(xs, ys)
        h_one = one.eval_nodes(poses, CLIP)
        res["parity_vs_single_gpu_whole_cloud"] = bool(np.array_equal(h_one, merged))
        one.close()
        # the reference's own CPU evaluation of the whole cloud on a node sample
        orc, kind = cpu_checker()
        optgs = orc.ackermann_ptgs()
        n_ref = min(16, args.nodes)
        ref, t_ref = orc.eval_nodes(optgs, xs, ys, poses[:n_ref], CLIP, mode=0, nthreads=0)
        init = np.concatenate([np.array([p.init_dist(k) for k in range(p.num_paths)]) for p in capi.ackermann_ptgs()])
        res["parity_vs_cpu_checker"] = {"kind": kind, "nodes": n_ref, "seconds": t_ref,
                                        "identical": bool(np.array_equal(np.minimum(init[None, :], merged[:n_ref].astype(np.float64)), ref))}
    dist.barrier()
    return res


def run_ours(args, rank, local_rank, world):
    import torch

    from mrpt_path_planning_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libmpp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    affinity = bind_to_gpu_numa_node(local_rank, world) if world > 1 else None
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = capi.Context(local_rank, stream.cuda_stream)

    # ---- set-up (not timed as steps; reported separately) ----
    capi.ackermann_ptgs(num_paths=5, ctx=ctx)  # warm the builder's buffers / module load
    t0 = time.perf_counter()
    ptgs = capi.ackermann_ptgs(ctx=ctx)  # paths on host threads, collision grids on the GPU (colgrid_build.cu)
    t_tables = time.perf_counter() - t0
    t_grid = ctx.colgrid_build_stats()
    for p in ptgs:
        ctx.add_ptg(p)
    xs, ys = wl.uniform_cloud(args.points, EXTENT)
    poses = wl.node_poses(args.nodes, EXTENT, seed=wl.C3_NODES_SEED + rank)
    t0 = time.perf_counter()
    ctx.set_obstacles(xs, ys)
    ctx.sync()
    t_cloud_first = time.perf_counter() - t0  # includes one-time allocations and kernel loading
    t0 = time.perf_counter()
    ctx.set_obstacles(xs, ys)
    ctx.sync()
    t_cloud = time.perf_counter() - t0
    B, cols = args.nodes, ctx.total_paths
    assert cols == N_PTGS * K_PATHS

    with torch.cuda.stream(stream):
        d_nodes = torch.from_numpy(capi.nodes_xycs_from_poses(poses)).to(dev)
        d_out = torch.empty((B, cols), dtype=torch.float32, device=dev)
        flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    h_poses = capi.pinned_empty((B, 3), np.float64)
    h_poses[:] = poses
    h_out = capi.pinned_empty((B, cols), np.float32)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize(dev)

    def device_steps(n, timed):
        evs = []
        with torch.cuda.stream(stream):
            for _ in range(n):
                flush.zero_()
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(stream)
                ctx.eval_nodes_device(B, d_nodes, CLIP, d_out)
                b.record(stream)
                evs.append((a, b))
        torch.cuda.synchronize(dev)
        return [a.elapsed_time(b) for a, b in evs] if timed else None

    def e2e_steps(n):
        ts = []
        for _ in range(n):
            with torch.cuda.stream(stream):
                flush.zero_()
            torch.cuda.synchronize(dev)
            t = time.perf_counter()
            ctx.eval_nodes(h_poses, CLIP, out=h_out)  # sincos + H2D + kernel + D2H + sync
            ts.append(time.perf_counter() - t)
        return ts

    fp32_peak = ctx.measure_fp32_tflops()
    fp64_peak = ctx.measure_fp64_tflops()

    device_steps(args.warmup, False)
    barrier()
    # one sampler per node (rank 0, all GPUs of the job in one nvidia-smi process): eight pollers contend for the driver
    sampler = ClockSampler(",".join(str(i) for i in range(world)) if world > 1 else local_rank)
    if rank == 0:
        sampler.start()
    l0 = ctx.launch_count
    near0 = ctx.grid_near_stats()
    step_ms = device_steps(args.steps, True)
    ctx.sync()
    launches = ctx.launch_count - l0
    near1 = ctx.grid_near_stats()
    near_r, near_bound = ctx.grid_near_window(CLIP)
    barrier()
    e2e_steps(max(1, args.warmup // 2))
    barrier()
    e2e_s = e2e_steps(args.steps)
    barrier()
    clocks = sampler.stop()

    dev_ms = sum(step_ms) / len(step_ms)
    e2e_ms = 1e3 * sum(e2e_s) / len(e2e_s)

    # the structured variant of C3 (70 % of the points on wall segments at 0.05 m spacing: dense cells,
    # long runs of occupied cells, large empty areas) on the same node batch; reported next to the headline, not as it
    structured = None
    if rank == 0 and not args.no_structured:
        wx, wy = wl.walls_cloud(args.points, EXTENT)
        ctx.set_obstacles(wx, wy)
        device_steps(3, False)
        sn0 = ctx.grid_near_stats()
        s_ms = device_steps(max(5, args.steps // 2), True)
        ctx.sync()
        sn1 = ctx.grid_near_stats()
        structured = {"workload": "C3 nodes, 70 % of the 2^20 points on wall segments (0.05 m spacing)",
                      "ms_per_step": sum(s_ms) / len(s_ms), "edges_per_s": B * cols / (1e-3 * sum(s_ms) / len(s_ms)),
                      "near_first_steps": sn1[0] - sn0[0],
                      "note": "free space around most nodes: the adaptive switch keeps these batches on the plain one-launch form"}
        ctx.set_obstacles(xs, ys)
        ctx.sync()

    # ---- secondary measurements (not the headline): planner level and, for N > 1, the cloud-sharded path ----
    planner = None
    if not args.no_planner:
        planner = planner_bench(lambda: capi.Context(local_rank), rank, world, args,
                                with_cpu=(world == 1 and not args.no_cpu_baseline))
    c5 = None
    if world > 1 and not args.no_c5:
        c5 = cloud_sharded_bench(ctx, dev, stream, rank, world, args, d_nodes, d_out, flush)
    c4_s = planner["c4_batch"]["batch_s"] if planner else 0.0
    c4f_s = planner["c4_batch"]["fixed_threads"]["batch_s"] if planner else 0.0
    c4_exp = float(planner["c4_batch"]["expansions"]) if planner else 0.0
    c4f_exp = float(planner["c4_batch"]["fixed_threads"]["expansions"]) if planner else 0.0
    c4_succ = float(planner["c4_batch"]["successes"]) if planner else 0.0
    c4f_q = float(planner["c4_batch"]["fixed_threads"]["queries"]) if planner else 0.0
    per_rank = None
    if world > 1:
        mine = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        allr = [torch.zeros_like(mine) for _ in range(world)]
        dist.all_gather(allr, mine)
        per_rank = {"kernel_ms": [float(a[0]) for a in allr], "e2e_ms": [float(a[1]) for a in allr]}
        t = torch.tensor([dev_ms, e2e_ms, c4_s, c4f_s], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms, c4_s, c4f_s = t.tolist()
        t = torch.tensor([c4_exp, c4f_exp, c4_succ, c4f_q], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        c4_exp, c4f_exp, c4_succ, c4f_q = t.tolist()
    if planner:
        c4 = planner["c4_batch"]
        c4["plans_per_s"] = c4["queries_total"] / c4_s
        c4["batch_s_max_over_ranks"] = c4_s
        c4["expansions_per_s"] = c4_exp / c4_s
        c4["successes_total"] = int(c4_succ)
        c4["success_rate"] = c4_succ / c4["queries_total"]
        c4["fixed_threads"]["plans_per_s"] = c4f_q / c4f_s
        c4["fixed_threads"]["queries_total"] = int(c4f_q)
        c4["fixed_threads"]["expansions_per_s"] = c4f_exp / c4f_s
        c4["fixed_threads"]["batch_s_max_over_ranks"] = c4f_s
    edges_total = world * B * cols
    value = edges_total / (dev_ms * 1e-3)
    e2e_value = edges_total / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    work = algorithmic_work(B, args.points)
    peaks, peak_src = measured_peaks()
    kernel_s = dev_ms * 1e-3  # the step is one launch of tpobs_grid_kernel, or two in its near-first form (near + rest)
    achieved_tf = work["flop_ref"] / kernel_s / 1e12
    # near-first form: points the two launches actually visit (the bins that intersect the near window of every node,
    # the whole window again for the nodes that needed the second launch)
    near_steps, near_nodes, near_pending = (near1[i] - near0[i] for i in range(3))
    near_first = None
    flop_visited = work["flop_binned"]
    if near_steps:
        bin_size = 0.5
        n_near = args.points * (2 * near_r + bin_size) ** 2 / (EXTENT * EXTENT)
        frac_pending = near_pending / max(near_nodes, 1)
        flop_visited = B * (n_near + frac_pending * work["n_in"]) * (6 + 8 + 6 * N_PTGS)
        near_first = {"steps": int(near_steps), "of_steps": args.steps, "nodes": int(near_nodes),
                      "nodes_needing_second_launch": int(near_pending), "window_half_size_m": near_r, "bound_m": near_bound,
                      "points_visited_per_node": n_near + frac_pending * work["n_in"],
                      "points_in_clip_window_per_node": work["n_in"], "flop_per_launch_visited": flop_visited,
                      "achieved_visited": flop_visited / kernel_s / 1e12,
                      "frac_visited": flop_visited / kernel_s / 1e12 / fp32_peak if fp32_peak else None,
                      "what": "rows whose paths are all blocked within bound_m by points of the near window are final "
                              "(every table entry has d >= distance(origin, cell) - largest robot radius); the other "
                              "nodes go to a second launch over the whole clipping window"}
    # Roofline of the dominant kernel (the step is exactly one launch of tpobs_grid_kernel). Nothing on this path is a
    # dense contraction and the cloud is L2-resident, so neither of the contract's two bounds ("hbm", "tensor") is the
    # limiter; the headline fraction is the PHYSICAL one against the CUDA-core FP32 peak: the FLOPs the binned kernel
    # has to execute (B * N_in * (6 + 8 + 6 P)) over the measured launch time. Next to it: the HBM figure, the
    # issue-slot figure (warp instructions of the ncu capture against the SMs' issue rate: what actually bounds the
    # kernel) and, under its own key, SURVEY 8(d)'s reference-equivalent convention, which exceeds 1 because the
    # kernel skips the 99 % of the reference's per-point tests that the bin index makes unnecessary.
    binned_tf = work["flop_binned"] / kernel_s / 1e12
    roofline = {
        "bound": "fp32 (CUDA cores; the contract's 'hbm' and 'tensor' figures do not bind this kernel: see 'hbm')",
        "kernel": "tpobs_grid_kernel (near pass + rest pass of the near-first form)" if near_first else "tpobs_grid_kernel",
        "achieved": binned_tf, "peak": fp32_peak,
        "unit": "TFLOP/s", "frac": binned_tf / fp32_peak if fp32_peak else None,
        "peak_source": "FMA microbenchmark on this GPU (MEASURED_PEAKS.json has no CUDA-core figure)",
        "convention": "FLOPs of the binned algorithm over the whole clipping window, B * N_in * (6 + 8 + 6 P) per step "
                      "(the figure of the earlier rounds' records); in its near-first form the step executes only "
                      "near_first.flop_per_launch_visited of them: see near_first.frac_visited",
        "flop_per_launch": work["flop_binned"], "near_first": near_first,
        "hbm": {"algorithmic_bytes": work["hbm_bytes"], "achieved": work["hbm_bytes"] / kernel_s / 1e9,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peak_src,
                "frac": work["hbm_bytes"] / kernel_s / 1e9 / peaks.get("hbm_gbs", 1)},
        "survey_convention": {
            "what": "SURVEY 8(d) reference-equivalent FLOPs: every node tests every point (26.2 kFLOP/edge)",
            "flop_per_launch": work["flop_ref"], "achieved": achieved_tf,
            "frac": achieved_tf / fp32_peak if fp32_peak else None},
        "fp64_peak_tflops": fp64_peak, "traffic": None,
    }
    tpath = os.path.join(ROOT, "profiles", TRAFFIC_FILE)
    if os.path.exists(tpath) and B == 4096 and args.points == (1 << 20):
        t = json.load(open(tpath))["tpobs_grid_kernel"]
        roofline["traffic"] = t["dram_bytes_read"] + t["dram_bytes_write"]  # bytes per launch, from the ncu capture
        roofline["traffic_source"] = "profiles/" + TRAFFIC_FILE
        # issue slots: one warp instruction per scheduler and cycle, 4 schedulers per SM
        sm_hz = 1e6 * (clocks.get("sm_mhz") or peaks.get("sm_max_mhz") or 1965.0)
        issue_peak = 148 * 4 * sm_hz
        roofline["issue_slots"] = {"warp_instructions_per_launch": t["warp_instructions"],
                                   "achieved_per_s": t["warp_instructions"] / kernel_s, "peak_per_s": issue_peak,
                                   "frac": t["warp_instructions"] / kernel_s / issue_peak,
                                   "ncu_issue_active_pct": t["issue_active_pct"], "ncu_warps_active_pct": t["warps_active_pct"]}
        roofline["limiter"] = ("per-CTA latency (a chain of dependent look-ups and block barriers per node) and instruction "
                               "issue: %.0f %% issue-slot utilisation at %.0f %% occupancy, %.1f M warp instructions per step "
                               "(ncu); neither HBM nor FMA throughput bounds this kernel" % (
                                   t["issue_active_pct"], t["warps_active_pct"], t["warp_instructions"] / 1e6))
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 transform + u32/f32 table min", "data": "synthetic",
        "config": config_dict(args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h_poses.nbytes),
                "d2h_bytes_per_step": int(h_out.nbytes),
                "note": "mppb_eval_nodes with pinned host buffers: host cos/sin (shared with the library's helper thread when the process may run on >= 8 CPUs) + H2D of the poses + kernel, whose result rows are stored straight into the host buffer over PCIe (counted as D2H bytes) + sync; cloud resident"},
        "gpu_launches": int(launches), "roofline": roofline, "cpu_affinity_cpus": affinity, "per_rank": per_rank,
        "setup": {"ptg_tables_s": t_tables, "ptg_tables_builder": "gpu (mppb_build_collision_grid)",
                  "last_collision_grid": t_grid, "cloud_upload_and_binning_s": t_cloud,
                  "cloud_upload_and_binning_first_call_s": t_cloud_first},
    }
    if structured:
        out["structured_variant"] = structured
    if planner:
        out["planner"] = planner
    if c5:
        out["c5_cloud_sharded"] = c5
    if world == 1 and not args.no_cpu_baseline:
        rate, done, secs, threads, ref_out, kind = cpu_reference_rate(xs, ys, poses, args.cpu_budget)
        t0 = time.perf_counter()
        cpu_checker()[0].ackermann_ptgs()
        out["setup"]["cpu_ptg_tables_s"] = time.perf_counter() - t0  # internal_initialize of both PTGs on the host
        init = np.concatenate([p.init_dists() for p in ptgs])
        ctx.eval_nodes(h_poses, CLIP, out=h_out)
        got = np.minimum(init[None, :], h_out[:done].astype(np.float64))
        out["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": kind,
            "sample": "first %d of %d nodes vs the full %d-point cloud (%.1f s on %d host threads)" % (
                done, B, args.points, secs, threads),
            "parity_on_sample": bool(np.array_equal(got, ref_out))}
    emit(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


_JSON_OUT = None


def emit(line):
    out = _JSON_OUT or sys.stdout
    out.write(line + "\n")
    out.flush()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--nodes", type=int, default=4096)
    ap.add_argument("--points", type=int, default=1 << 20)
    ap.add_argument("--cpu-budget", type=float, default=15.0, help="seconds of CPU baseline work")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-planner", action="store_true", help="skip the planner-level (C1/C2/C4) measurements")
    ap.add_argument("--no-structured", action="store_true", help="skip the structured (wall map) variant of C3")
    ap.add_argument("--no-c5", action="store_true", help="skip the cloud-sharded (C5) measurement at N > 1")
    ap.add_argument("--queries", type=int, default=4096, help="C4: planning queries in total (split over the ranks)")
    ap.add_argument("--c4-max-expansions", type=int, default=5000,
                    help="C4: expansion cap per query (5000: >= 90 %% of the queries reach their goal)")
    ap.add_argument("--c5-points-per-gpu", type=int, default=1 << 23)
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    # stdout carries the ONE JSON line and nothing else: libraries that write to file descriptor 1 (NCCL prints
    # its version banner there) are pointed at stderr, the JSON line goes to a private copy of the real stdout
    global _JSON_OUT
    sys.stdout.flush()
    _JSON_OUT = os.fdopen(os.dup(1), "w")
    os.dup2(2, 1)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
