"""Deterministic synthetic spinning-LiDAR scans.

A small "DOALS-shaped town": ground plane, an outer 60 x 40 x 8 m hall, 24 axis-aligned boxes and
6 vertical poles.  A sensor with H beams x W azimuth steps is ray-cast against that world; the
nearest hit gets additive range noise and is returned in the *sensor* frame, row-major
(`idx = row * W + col`, the organised layout the reference assumes, odom.cc:128-130).  Everything
is seeded, so the oracle, the CUDA path and the benchmark all see bit-identical float32 inputs.

Only numpy is used; nothing here runs on the GPU and nothing here is part of the timed path.
"""
from __future__ import annotations

import math
from dataclasses import dataclass

import numpy as np

WORLD_SEED = 7
HALL_MIN = np.array([-30.0, -20.0, 0.0])
HALL_MAX = np.array([30.0, 20.0, 8.0])
RANGE_SIGMA = 0.01
RANGE_MIN = 0.5
RANGE_MAX = 100.0


@dataclass(frozen=True)
class World:
    box_min: np.ndarray  # (B,3)
    box_max: np.ndarray  # (B,3)
    cyl_xy: np.ndarray  # (C,2)
    cyl_r: np.ndarray  # (C,)
    cyl_h: np.ndarray  # (C,)


def make_world(seed: int = WORLD_SEED, n_boxes: int = 24, n_cyl: int = 6) -> World:
    rng = np.random.default_rng(seed)
    centre = np.stack([rng.uniform(-26, 26, n_boxes), rng.uniform(-17, 17, n_boxes)], axis=1)
    # keep a 3 m corridor around the trajectory (y ~ 0) free of obstacles
    centre[:, 1] = np.where(np.abs(centre[:, 1]) < 3.5, np.sign(centre[:, 1] + 1e-9) * 3.5 + centre[:, 1], centre[:, 1])
    half = np.stack([rng.uniform(0.5, 3.0, n_boxes), rng.uniform(0.5, 3.0, n_boxes)], axis=1)
    height = rng.uniform(1.0, 6.0, n_boxes)
    box_min = np.concatenate([centre - half, np.zeros((n_boxes, 1))], axis=1)
    box_max = np.concatenate([centre + half, height[:, None]], axis=1)
    cyl_xy = np.stack([rng.uniform(-24, 24, n_cyl), rng.uniform(4.0, 15.0, n_cyl) * rng.choice([-1.0, 1.0], n_cyl)], axis=1)
    cyl_r = np.full(n_cyl, 0.4)
    cyl_h = rng.uniform(3.0, 7.0, n_cyl)
    return World(box_min, box_max, cyl_xy, cyl_r, cyl_h)


def pose(frame: int, rate_hz: float = 10.0) -> np.ndarray:
    """Ground-truth sensor->world pose (4x4 float64) of `frame`.

    1 m/s along +x from x=-8, 0.3 m sinusoidal sway in y, 5 deg/s yaw, sensor 1.5 m above ground.
    """
    t = frame / rate_hz
    yaw = math.radians(5.0) * t
    T = np.eye(4)
    c, s = math.cos(yaw), math.sin(yaw)
    T[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    T[:3, 3] = [-8.0 + 1.0 * t, 0.3 * math.sin(0.5 * t), 1.5]
    return T


def _ray_dirs(beams: int, cols: int) -> np.ndarray:
    elev = np.radians(np.linspace(-22.5, 22.5, beams))
    azim = 2.0 * np.pi * np.arange(cols) / cols
    ce, se = np.cos(elev)[:, None], np.sin(elev)[:, None]
    d = np.stack([ce * np.cos(azim)[None, :], ce * np.sin(azim)[None, :], np.broadcast_to(se, (beams, cols))], axis=-1)
    return d.reshape(-1, 3)


def _cast(world: World, o: np.ndarray, d: np.ndarray) -> np.ndarray:
    """Nearest hit range for rays o + t d (d unit, (N,3)); inf if none."""
    n = d.shape[0]
    with np.errstate(divide="ignore", invalid="ignore"):
        inv = 1.0 / d
        # hall: we are inside, take the exit distance
        t1 = (HALL_MIN[None, :] - o[None, :]) * inv
        t2 = (HALL_MAX[None, :] - o[None, :]) * inv
        best = np.min(np.maximum(t1, t2), axis=1)
        # boxes: slab test, chunked over boxes to bound memory
        for b0, b1 in zip(world.box_min, world.box_max):
            ta = (b0[None, :] - o[None, :]) * inv
            tb = (b1[None, :] - o[None, :]) * inv
            tn = np.max(np.minimum(ta, tb), axis=1)
            tf = np.min(np.maximum(ta, tb), axis=1)
            hit = (tn <= tf) & (tn > 0.0)
            best = np.where(hit & (tn < best), tn, best)
        # vertical cylinders
        a = d[:, 0] ** 2 + d[:, 1] ** 2
        for (cx, cy), r, h in zip(world.cyl_xy, world.cyl_r, world.cyl_h):
            ox, oy = o[0] - cx, o[1] - cy
            bq = ox * d[:, 0] + oy * d[:, 1]
            cq = ox * ox + oy * oy - r * r
            disc = bq * bq - a * cq
            ok = (disc > 0.0) & (a > 1e-12)
            tc = (-bq - np.sqrt(np.where(ok, disc, 0.0))) / np.where(a > 1e-12, a, 1.0)
            z = o[2] + tc * d[:, 2]
            hit = ok & (tc > 0.0) & (z >= 0.0) & (z <= h)
            best = np.where(hit & (tc < best), tc, best)
    assert best.shape == (n,)
    return best


def scan(frame: int, beams: int = 64, cols: int = 1024, world: World | None = None, noise_seed: int | None = None) -> np.ndarray:
    """One scan in the sensor frame as float32 (N,4) with w = 1 (pcl::PointXYZI data[3], PCL convention)."""
    world = world or make_world()
    T = pose(frame)
    dirs_s = _ray_dirs(beams, cols)
    dirs_w = dirs_s @ T[:3, :3].T
    rng_hit = _cast(world, T[:3, 3], dirs_w)
    rng = np.random.default_rng(1000 + frame if noise_seed is None else noise_seed)
    rng_noisy = rng_hit + rng.normal(0.0, RANGE_SIGMA, rng_hit.shape)
    keep = np.isfinite(rng_noisy) & (rng_noisy > RANGE_MIN) & (rng_noisy < RANGE_MAX)
    pts = (dirs_s * rng_noisy[:, None])[keep]
    out = np.ones((pts.shape[0], 4), dtype=np.float32)
    out[:, :3] = pts.astype(np.float32)
    return out


def organized_scan(frame: int, beams: int = 64, cols: int = 1024, world: World | None = None, dropout: float = 0.0,
                   noise_seed: int | None = None) -> np.ndarray:
    """The same scan kept organised, as the segmentation stage wants it (detection.cpp:296-327): float32
    (beams, cols, 4), sensor frame, row 0 = the TOP beam (groundRemoval walks up from row H-1), NaN where the ray
    has no return; `dropout` removes that fraction of the returns at random on top."""
    world = world or make_world()
    T = pose(frame)
    dirs_s = _ray_dirs(beams, cols)
    rng_hit = _cast(world, T[:3, 3], dirs_s @ T[:3, :3].T)
    rng = np.random.default_rng(1000 + frame if noise_seed is None else noise_seed)
    rng_noisy = rng_hit + rng.normal(0.0, RANGE_SIGMA, rng_hit.shape)
    keep = np.isfinite(rng_noisy) & (rng_noisy > RANGE_MIN) & (rng_noisy < RANGE_MAX)
    if dropout > 0.0:
        keep &= rng.random(rng_hit.shape) >= dropout
    out = np.full((beams * cols, 4), np.nan, dtype=np.float32)
    out[keep, :3] = (dirs_s * rng_noisy[:, None])[keep].astype(np.float32)
    out[keep, 3] = 1.0
    return out.reshape(beams, cols, 4)[::-1].copy()


def organized_transform(scan: np.ndarray, T: np.ndarray) -> np.ndarray:
    """pcl::transformPointCloud on an organised scan: NaN points stay NaN."""
    flat = scan.reshape(-1, scan.shape[-1])
    out = np.full_like(flat, np.nan)
    ok = np.isfinite(flat[:, 0])
    out[ok] = transform(flat[ok], T)
    return out.reshape(scan.shape)


def transform(points: np.ndarray, T: np.ndarray) -> np.ndarray:
    """Apply a 4x4 pose to (N,4) float32 points (float64 arithmetic, rounded once)."""
    out = np.ones_like(points, dtype=np.float32)
    out[:, :3] = (points[:, :3].astype(np.float64) @ T[:3, :3].T + T[:3, 3]).astype(np.float32)
    return out


def voxel_filter(points: np.ndarray, leaf: float) -> np.ndarray:
    """Centroid-per-voxel down-sampling (what pcl::VoxelGrid does, odom.cc:469-474)."""
    key = np.floor(points[:, :3].astype(np.float64) / leaf).astype(np.int64)
    key -= key.min(axis=0)
    dims = key.max(axis=0) + 1
    flat = (key[:, 0] * dims[1] + key[:, 1]) * dims[2] + key[:, 2]
    order = np.argsort(flat, kind="stable")
    flat_s = flat[order]
    start = np.flatnonzero(np.concatenate([[True], flat_s[1:] != flat_s[:-1]]))
    sums = np.add.reduceat(points[order, :3].astype(np.float64), start, axis=0)
    cnt = np.diff(np.concatenate([start, [len(flat_s)]]))
    out = np.ones((len(start), 4), dtype=np.float32)
    out[:, :3] = (sums / cnt[:, None]).astype(np.float32)
    return out


def submap(n_points: int, frames, beams: int = 64, cols: int = 1024, world: World | None = None, shuffle_seed: int = 99) -> np.ndarray:
    """World-frame keyframe submap: union of `frames` moved by their ground-truth pose, shuffled, truncated."""
    world = world or make_world()
    parts = [transform(scan(f, beams, cols, world), pose(f)) for f in frames]
    cloud = np.concatenate(parts, axis=0)
    perm = np.random.default_rng(shuffle_seed).permutation(cloud.shape[0])
    cloud = cloud[perm]
    if cloud.shape[0] < n_points:
        raise ValueError(f"submap has only {cloud.shape[0]} points, wanted {n_points}")
    return np.ascontiguousarray(cloud[:n_points])


def perturbed_guess(T: np.ndarray, dxyz=(0.10, 0.10, 0.02), dyaw_deg: float = 1.0) -> np.ndarray:
    """Ground truth moved by a fixed offset: the S2M initial guess of config C2 (float32 4x4)."""
    c, s = math.cos(math.radians(dyaw_deg)), math.sin(math.radians(dyaw_deg))
    D = np.eye(4)
    D[:3, :3] = [[c, -s, 0.0], [s, c, 0.0], [0.0, 0.0, 1.0]]
    D[:3, 3] = dxyz
    return (D @ T).astype(np.float32)


def workload_c1(beams: int = 64, cols: int = 1024):
    """C1: S2S pair, target = frame 0, source = frame 1, guess = I."""
    w = make_world()
    return scan(1, beams, cols, w), scan(0, beams, cols, w), np.eye(4, dtype=np.float32)


def workload_c2(n_target: int = 500_000, beams: int = 64, cols: int = 1024, src_frame: int = 50):
    """C2: S2M, source = frame 50 (sensor frame), target = n_target-point submap, guess = perturbed truth."""
    w = make_world()
    per = beams * cols
    need = int(math.ceil(n_target / (0.95 * per))) + 1
    frames = [5 * i for i in range(need)]
    tgt = submap(n_target, frames, beams, cols, w)
    src = scan(src_frame, beams, cols, w)
    return src, tgt, perturbed_guess(pose(src_frame))


def workload_c4(n_target: int = 2_000_000, beams: int = 128, cols: int = 2048, src_frame: int = 50):
    """C4: as C2 with an OS1-128-shaped 128x2048 scan (~262k points) against a 2M-point submap."""
    return workload_c2(n_target, beams, cols, src_frame)
