#!/usr/bin/env python
"""bench.py — headline benchmark of the hot path (BASELINE.json metric: candidate edges evaluated / s).

Workload (config C3 of BASELINE.json / SURVEY.md 8d): synthetic 2^20-point obstacle cloud
U[0,200)^2 (PCG64 seed 20240601), a 4096-node expansion batch (seed 20240602 + rank) x both
Ackermann PTGs x all 121 trajectories = 991 232 candidate edges per step and per GPU.
A step = one pass of stage (a)+(b) over one node batch against the resident cloud.

  python bench.py --gpus N --steps K --warmup W            (ours; under torchrun for N > 1)
  python bench.py --impl reference --gpus N --steps K ...  (CPU arm: the oracle port on host cores)

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


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        return json.load(open(p)), "measured"
    return {"hbm_gbs": 6650.0, "bf16_tflops": 1590.0}, "fallback"


def cpu_reference_rate(xs, ys, poses, budget_s, nthreads=0):
    """Times the oracle's restatement of the reference call pattern (transform_pc_square_clipping +
    tp_obstacles_single_path per (PTG,k)) on a bounded sample of the node batch. Returns
    (edges_per_s, nodes_done, seconds, threads, outputs)."""
    from oracle import orc

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
    return done * N_PTGS * K_PATHS / secs, done, secs, threads, np.concatenate(outs, 0)


def run_reference(args, rank, world):
    if rank != 0:
        return
    xs, ys = wl.uniform_cloud(args.points, EXTENT)
    poses = wl.node_poses(args.nodes, EXTENT)
    from oracle import orc

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
    print(json.dumps({
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": config_dict(args, sample_nodes=sample),
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": threads, "kind": "port", "sample": desc},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }))


def config_dict(args, **extra):
    c = {"workload": "C3: synthetic 1M-point cloud U[0,200)^2, 4096-node batch x 2 Ackermann PTGs x 121 k, clip 5 m",
         "points": args.points, "nodes_per_gpu": args.nodes, "ptgs": N_PTGS, "paths_per_ptg": K_PATHS,
         "edges_per_step_per_gpu": args.nodes * N_PTGS * K_PATHS,
         "l2": "flushed between timed steps (256 MiB write)", "parallelism": "nodes sharded per GPU, cloud replicated"}
    c.update(extra)
    return c


def run_ours(args, rank, local_rank, world):
    import torch

    from mrpt_path_planning_b200 import capi

    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device — libmpp_b200 has no CPU fallback")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        import torch.distributed as dist

        dist.init_process_group("nccl", device_id=dev)
    stream = torch.cuda.Stream(device=dev)
    ctx = capi.Context(local_rank, stream.cuda_stream)

    # ---- set-up (not timed as steps; reported separately) ----
    t0 = time.perf_counter()
    ptgs = capi.ackermann_ptgs()
    t_tables = time.perf_counter() - t0
    for p in ptgs:
        ctx.add_ptg(p)
    xs, ys = wl.uniform_cloud(args.points, EXTENT)
    poses = wl.node_poses(args.nodes, EXTENT, seed=wl.C3_NODES_SEED + rank)
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
    sampler = ClockSampler(local_rank)
    sampler.start()
    l0 = ctx.launch_count
    step_ms = device_steps(args.steps, True)
    launches = ctx.launch_count - l0
    barrier()
    e2e_steps(max(1, args.warmup // 2))
    barrier()
    e2e_s = e2e_steps(args.steps)
    barrier()
    clocks = sampler.stop()

    dev_ms = sum(step_ms) / len(step_ms)
    e2e_ms = 1e3 * sum(e2e_s) / len(e2e_s)
    if world > 1:
        t = torch.tensor([dev_ms, e2e_ms], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        dev_ms, e2e_ms = t.tolist()
    edges_total = world * B * cols
    value = edges_total / (dev_ms * 1e-3)
    e2e_value = edges_total / (e2e_ms * 1e-3)

    if rank != 0:
        if world > 1:
            dist.destroy_process_group()
        return

    work = algorithmic_work(B, args.points)
    peaks, peak_src = measured_peaks()
    kernel_s = dev_ms * 1e-3  # the step is exactly one launch of tpobs_grid_kernel
    achieved_tf = work["flop_ref"] / kernel_s / 1e12
    roofline = {
        "bound": "fp32", "kernel": "tpobs_grid_kernel", "achieved": achieved_tf, "peak": fp32_peak,
        "unit": "TFLOP/s", "frac": achieved_tf / fp32_peak if fp32_peak else None,
        "peak_source": "FMA microbenchmark on this GPU (MEASURED_PEAKS.json has no CUDA-core figure)",
        "convention": "SURVEY 8(d) reference-equivalent FLOPs: every node tests every point (26.2 kFLOP/edge); "
                      "the kernel bins the cloud and skips out-of-window points, so frac can exceed 1",
        "flop_per_launch": work["flop_ref"],
        "binned": {"flop_per_launch": work["flop_binned"], "achieved": work["flop_binned"] / kernel_s / 1e12,
                   "frac": work["flop_binned"] / kernel_s / 1e12 / fp32_peak if fp32_peak else None},
        "hbm": {"algorithmic_bytes": work["hbm_bytes"], "achieved": work["hbm_bytes"] / kernel_s / 1e9,
                "peak": peaks.get("hbm_gbs"), "unit": "GB/s", "peak_source": peak_src,
                "frac": work["hbm_bytes"] / kernel_s / 1e9 / peaks.get("hbm_gbs", 1)},
        "fp64_peak_tflops": fp64_peak, "traffic": None,
    }
    out = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": dev_ms, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 transform + u32/f32 table min", "data": "synthetic",
        "config": config_dict(args), "clocks": clocks,
        "e2e": {"value": e2e_value, "unit": UNIT, "ms_per_step": e2e_ms, "h2d_bytes_per_step": int(h_poses.nbytes),
                "d2h_bytes_per_step": int(h_out.nbytes),
                "note": "mppb_eval_nodes: host cos/sin + H2D(pinned) + kernel + D2H(pinned) + sync; cloud resident"},
        "gpu_launches": int(launches), "roofline": roofline,
        "setup": {"ptg_tables_s": t_tables, "cloud_upload_and_binning_s": t_cloud},
    }
    if world == 1 and not args.no_cpu_baseline:
        rate, done, secs, threads, ref_out = cpu_reference_rate(xs, ys, poses, args.cpu_budget)
        init = np.concatenate([p.init_dists() for p in ptgs])
        ctx.eval_nodes(h_poses, CLIP, out=h_out)
        got = np.minimum(init[None, :], h_out[:done].astype(np.float64))
        out["cpu_baseline"] = {
            "value": rate, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": "first %d of %d nodes vs the full %d-point cloud (%.1f s on %d host threads)" % (
                done, B, args.points, secs, threads),
            "parity_on_sample": bool(np.array_equal(got, ref_out))}
    print(json.dumps(out))
    if world > 1:
        dist.destroy_process_group()


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
    args = ap.parse_args()
    args.warmup = max(args.warmup, 3)
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    if args.impl == "reference":
        run_reference(args, rank, world)
    else:
        run_ours(args, rank, local_rank, world)


if __name__ == "__main__":
    main()
