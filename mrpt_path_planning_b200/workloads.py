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
