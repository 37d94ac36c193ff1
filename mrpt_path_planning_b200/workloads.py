"""Synthetic workloads of BASELINE.json / SURVEY.md section 8(d) (inputs only; no compute)."""
import numpy as np

C3_CLOUD_SEED = 20240601
C3_NODES_SEED = 20240602
C3_WALLS_SEED = 20240603
C5_CLOUD_SEED = 20240605


def uniform_cloud(n, extent=200.0, seed=C3_CLOUD_SEED):
    """n float32 points, x,y ~ U[0,extent) (config C3: n=2**20, extent=200)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xs = rng.uniform(0.0, extent, n).astype(np.float32)
    ys = rng.uniform(0.0, extent, n).astype(np.float32)
    return xs, ys


def node_poses(b, extent=200.0, margin=5.0, seed=C3_NODES_SEED):
    """b node poses, x,y ~ U[margin, extent-margin), phi ~ U[-pi,pi), float64 (config C3: b=4096)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    x = rng.uniform(margin, extent - margin, b)
    y = rng.uniform(margin, extent - margin, b)
    phi = rng.uniform(-np.pi, np.pi, b)
    return np.stack([x, y, phi], axis=1)


def walls_cloud(n, extent=200.0, wall_fraction=0.7, spacing=0.05, seed=C3_WALLS_SEED):
    """Structured variant: wall_fraction of the points on random wall segments at `spacing` m."""
    rng = np.random.Generator(np.random.PCG64(seed))
    n_wall = int(n * wall_fraction)
    xs = np.empty(n, dtype=np.float64)
    ys = np.empty(n, dtype=np.float64)
    i = 0
    while i < n_wall:
        length = rng.uniform(2.0, 30.0)
        m = min(n_wall - i, int(length / spacing) + 1)
        x0, y0 = rng.uniform(0, extent, 2)
        th = rng.uniform(0, np.pi)
        t = np.arange(m) * spacing
        xs[i:i + m] = x0 + t * np.cos(th)
        ys[i:i + m] = y0 + t * np.sin(th)
        i += m
    xs[n_wall:] = rng.uniform(0, extent, n - n_wall)
    ys[n_wall:] = rng.uniform(0, extent, n - n_wall)
    np.clip(xs, 0, extent, out=xs)
    np.clip(ys, 0, extent, out=ys)
    return xs.astype(np.float32), ys.astype(np.float32)


C4_QUERIES_SEED = 20240604

# share/mvsim-demo-astar-planner-params.yaml (values typed here as data): grid 0.40 m / 25 deg, metric weights 0.1,
# 13 trajectories, 3 speeds, 5 interpolated segments; [7] = time limit (0 = none), [8] = expansion cap (0 = none)
DEMO_ASTAR_PARAMS = [0.40, 25.0 * np.pi / 180.0, 0.1, 0.1, 13, 3, 5, 0, 0]
DEMO_ASTAR_TIMESTAMPS = [0.5, 1.5, 4.0]


def planning_queries(n, extent=200.0, dmin=10.0, dmax=40.0, margin=4.0, seed=C4_QUERIES_SEED):
    """Config C4: n independent queries, start/goal ~ U[10, extent-10)^2 with distance in [dmin, dmax], random
    headings, per-query world box = start/goal +- margin. Returns tuples (start6, goal6, bbox_min3, bbox_max3, False)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    out = []
    while len(out) < n:
        s = rng.uniform(10.0, extent - 10.0, 2)
        g = rng.uniform(10.0, extent - 10.0, 2)
        d = float(np.hypot(*(g - s)))
        hs, hg = rng.uniform(-np.pi, np.pi, 2)
        if not (dmin <= d <= dmax):
            continue
        lo = np.minimum(s, g) - margin
        hi = np.maximum(s, g) + margin
        out.append(([s[0], s[1], hs, 0.0, 0.0, 0.0], [g[0], g[1], hg, 0.0, 0.0, 0.0],
                    [lo[0], lo[1], -np.pi], [hi[0], hi[1], np.pi], False))
    return out


C4_MAP_SEED = 20240606


def pillars_map(n, extent=200.0, r_min=0.5, r_max=1.5, spacing=0.005, seed=C4_MAP_SEED):
    """Config C4's "synthetic 200x200 m map": n float32 points on the outlines of round pillars (radius U[r_min, r_max],
    points every `spacing` m along the outline), centres U[0, extent)^2, drawn until the n points are used up — about
    800 pillars for n = 2**20, one per ~48 m^2, i.e. a map a car-like robot can cross in any direction (the uniform C3
    cloud, 26 points per m^2, leaves no free space at all). Returns (xs, ys, pillars[m][3] = cx, cy, r)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    xs = np.empty(n, dtype=np.float64)
    ys = np.empty(n, dtype=np.float64)
    pillars = []
    i = 0
    while i < n:
        cx, cy = rng.uniform(0.0, extent, 2)
        r = rng.uniform(r_min, r_max)
        m = min(n - i, max(8, int(2.0 * np.pi * r / spacing)))
        a = np.arange(m) * (2.0 * np.pi / m)
        xs[i:i + m] = cx + r * np.cos(a)
        ys[i:i + m] = cy + r * np.sin(a)
        pillars.append((cx, cy, r))
        i += m
    return xs.astype(np.float32), ys.astype(np.float32), np.array(pillars)


def planning_queries_free(n, pillars, extent=200.0, dmin=10.0, dmax=40.0, margin=4.0, clearance=1.0,
                          seed=C4_QUERIES_SEED):
    """planning_queries() on a pillars_map(): start and goal keep `clearance` m from every pillar outline (a start pose
    in collision cannot be planned from)."""
    rng = np.random.Generator(np.random.PCG64(seed))
    pc, pr = pillars[:, :2], pillars[:, 2]

    def free(p):
        return bool(np.all(np.hypot(pc[:, 0] - p[0], pc[:, 1] - p[1]) > pr + clearance))

    out = []
    while len(out) < n:
        s = rng.uniform(10.0, extent - 10.0, 2)
        g = rng.uniform(10.0, extent - 10.0, 2)
        d = float(np.hypot(*(g - s)))
        hs, hg = rng.uniform(-np.pi, np.pi, 2)
        if not (dmin <= d <= dmax) or not free(s) or not free(g):
            continue
        lo = np.minimum(s, g) - margin
        hi = np.maximum(s, g) + margin
        out.append(([s[0], s[1], hs, 0.0, 0.0, 0.0], [g[0], g[1], hg, 0.0, 0.0, 0.0],
                    [lo[0], lo[1], -np.pi], [hi[0], hi[1], np.pi], False))
    return out
