"""Synthetic scan generators for the BASELINE.json configs.

All clouds are float32 (n, 4) arrays in pcl::PointXYZ layout (x, y, z, pad=1.0; 16-byte stride) in
MAP coordinates, with a float64 sensor origin, exactly what
VDBMapping::insertPointCloud(cloud, origin, source_id) takes
(/root/reference/include/vdb_mapping/VDBMapping.hpp:399-406).
RNG: numpy Generator(PCG64(1234 + cfg [+ scan index])); range noise = range * (1 + 0.01 N(0,1));
about 1 % of the points are NaN to exercise the filter at VDBMapping.hpp:505-510.
"""
from __future__ import annotations

from dataclasses import dataclass

import numpy as np


@dataclass(frozen=True)
class ScanConfig:
    name: str
    cfg: int
    resolution: float
    max_range: float
    n_points: int
    prob_hit: float = 0.7
    prob_miss: float = 0.4
    prob_thres_min: float = 0.12
    prob_thres_max: float = 0.97
    description: str = ""


CONFIGS = {
    1: ScanConfig("cfg1_lidar64_0.1m_50m", 1, 0.1, 50.0, 64 * 2048,
                  description="synthetic 64-beam LiDAR scan (~130k pts, 0.1 m voxels, 50 m max range)"),
    2: ScanConfig("cfg2_os1-128_0.05m_30m", 2, 0.05, 30.0, 128 * 2048,
                  description="Ouster OS1-128-style synthetic scan (262k pts, 0.05 m voxels, 30 m range)"),
    3: ScanConfig("cfg3_rgbd_640x480_0.02m_5m", 3, 0.02, 5.0, 640 * 480,
                  description="RGB-D depth-image cloud (640x480 = 307k pts, 0.02 m voxels, 5 m range)"),
    4: ScanConfig("cfg4_multilidar_1M_0.02m_100m", 4, 0.02, 100.0, 4 * 128 * 2048,
                  description="1M-point merged multi-LiDAR scan, 0.02 m voxels, 100 m range, bounded scene"),
}


def _finish(points64: np.ndarray, rng: np.random.Generator, nan_frac: float) -> np.ndarray:
    n = points64.shape[0]
    out = np.ones((n, 4), dtype=np.float32)
    out[:, :3] = points64.astype(np.float32)
    if nan_frac > 0:
        idx = rng.random(n) < nan_frac
        out[idx, rng.integers(0, 3)] = np.nan
    return out


def _box_scene_range(o: np.ndarray, d: np.ndarray, half_x: float, half_y: float, ground_z: float,
                     ceil_z: float | None, no_hit: float) -> np.ndarray:
    """Distance along unit directions d from o to an axis-aligned room (world-fixed)."""
    with np.errstate(divide="ignore", invalid="ignore"):
        t = np.full(d.shape[0], no_hit)
        for axis, lim in ((0, half_x), (1, half_y)):
            for s in (+1.0, -1.0):
                tt = (s * lim - o[axis]) / d[:, axis]
                ok = (tt > 0) & np.isfinite(tt)
                t = np.where(ok & (tt < t), tt, t)
        tg = (ground_z - o[2]) / d[:, 2]
        ok = (tg > 0) & np.isfinite(tg)
        t = np.where(ok & (tg < t), tg, t)
        if ceil_z is not None:
            tc = (ceil_z - o[2]) / d[:, 2]
            ok = (tc > 0) & np.isfinite(tc)
            t = np.where(ok & (tc < t), tc, t)
    return t


def _lidar_dirs(n_beams: int, n_az: int, elev_lo_deg: float, elev_hi_deg: float, yaw_deg: float,
                pitch_deg: float = 0.0) -> np.ndarray:
    el = np.deg2rad(np.linspace(elev_lo_deg, elev_hi_deg, n_beams))
    az = np.deg2rad(yaw_deg) + 2.0 * np.pi * np.arange(n_az) / n_az
    # azimuth-major firing order (all beams of one column, then the next column), like a spinning lidar
    A, E = np.meshgrid(az, el, indexing="ij")
    d = np.stack([np.cos(E) * np.cos(A), np.cos(E) * np.sin(A), np.sin(E)], axis=-1).reshape(-1, 3)
    if pitch_deg:
        p = np.deg2rad(pitch_deg)
        R = np.array([[np.cos(p), 0, np.sin(p)], [0, 1, 0], [-np.sin(p), 0, np.cos(p)]])
        d = d @ R.T
    return d


def lidar_scan(cfg: int, k: int = 0, nan_frac: float = 0.01, sensor: int = 0):
    """cfg 1 / 2 / 4 scan number k. Returns (points (n,4) f32, origin (3,) f64).
    sensor > 0 (cfg 2 only): another OS1-128 of a merged multi-LiDAR rig sharing the same origin, yawed by a
    fraction of the azimuth step and pitched by 2.5 deg per sensor (used for the multi-GPU runs: one LiDAR per GPU)."""
    c = CONFIGS[cfg]
    rng = np.random.default_rng(np.random.PCG64(1234 + cfg + 1000 * k + 7919 * sensor))
    R = c.max_range
    if cfg == 1:
        origin = np.array([0.0, 0.0, 0.0])
        d = _lidar_dirs(64, 2048, -24.8, 2.0, 0.0)
        r = _box_scene_range(origin, d, 0.8 * R, 0.8 * R, -1.8, None, 2.0 * R)
    elif cfg == 2:
        origin = np.array([0.137 * k, 0.061 * k, 0.013 * k])
        d = _lidar_dirs(128, 2048, -22.5, 22.5, 0.5 * k + (360.0 / 2048.0) * (sensor / 8.0), pitch_deg=2.5 * sensor)
        r = _box_scene_range(origin, d, 0.8 * R, 0.8 * R, -1.8, None, 2.0 * R)
    elif cfg == 4:
        origin = np.array([0.211 * k, 0.093 * k, 0.0])
        parts = [_lidar_dirs(128, 2048, -22.5, 22.5, 0.5 * k + 0.044 * s, pitch_deg=(-6.0 + 4.0 * s)) for s in range(4)]
        d = np.concatenate(parts, axis=0)
        # bounded hall: walls at 18 m / 34 m, floor 1.8 m below, ceiling 7 m above; two "door" sectors
        # (6 % of azimuths) see nothing and are clipped at max range
        r = _box_scene_range(origin, d, 18.0, 34.0, -1.8, 7.0, 2.0 * R)
        az = np.arctan2(d[:, 1], d[:, 0])
        door = ((np.abs(az - 0.3) < 0.09) | (np.abs(az + 2.1) < 0.09)) & (np.abs(d[:, 2]) < 0.05)
        r = np.where(door, 2.0 * R, r)
    else:
        raise ValueError(cfg)
    r = r * (1.0 + 0.01 * rng.standard_normal(r.shape[0]))
    pts = origin[None, :] + d * r[:, None]
    return _finish(pts, rng, nan_frac), origin


def rgbd_scan(k: int = 0, nan_frac: float = 0.01):
    """cfg 3: 640x480 pinhole (fx=fy=525, cx=319.5, cy=239.5), camera looking +x."""
    rng = np.random.default_rng(np.random.PCG64(1234 + 3 + 1000 * k))
    origin = np.array([0.011 * k, 0.007 * k, 0.003 * k])
    v, u = np.meshgrid(np.arange(480), np.arange(640), indexing="ij")
    u = u.reshape(-1).astype(np.float64)
    v = v.reshape(-1).astype(np.float64)
    depth = 1.0 + 3.5 * (0.5 + 0.5 * np.sin(0.02 * u + 0.1 * k) * np.cos(0.03 * v)) + 0.005 * rng.standard_normal(u.shape[0])
    # a band of far pixels so that some rays exceed the 5 m range and get clipped
    depth = np.where(v < 6, 6.0, depth)
    x = depth
    y = -(u - 319.5) / 525.0 * depth
    z = -(v - 239.5) / 525.0 * depth
    pts = origin[None, :] + np.stack([x, y, z], axis=-1)
    return _finish(pts, rng, nan_frac), origin


def make_scan(cfg: int, k: int = 0, nan_frac: float = 0.01, sensor: int = 0):
    return rgbd_scan(k, nan_frac) if cfg == 3 else lidar_scan(cfg, k, nan_frac, sensor)


def small_scan(seed: int, n: int = 4096, scale: float = 3.0, nan_frac: float = 0.02):
    """Small random cloud for fast parity tests (mix of short, long, axis-aligned and tie-prone rays)."""
    rng = np.random.default_rng(seed)
    p = rng.normal(size=(n, 3)) * scale
    q = n // 8
    p[:q] = np.round(p[:q] * 4) / 4          # lattice points -> exact ties in the DDA
    p[q:2 * q, 1:] = 0.0                      # axis-aligned
    p[2 * q:3 * q] *= 4.0                     # long rays (clipped)
    p[3 * q:3 * q + 8] = 0.0                  # zero-length rays
    origin = rng.normal(size=3) * 0.3
    return _finish(p + origin, rng, nan_frac), origin
