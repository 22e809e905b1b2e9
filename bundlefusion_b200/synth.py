"""Seeded synthetic RGB-D frames: an analytic room (axis-aligned box 6x3x6 m,
three spheres, sinusoid-textured walls) seen from a camera on a Lissajous path.

numpy only (host-side input generation; not part of the hot path).  Frames are what the
reference's sensor contract delivers (RGBDSensor::getDepthFloat / getColorRGBX): depth float32
in metres with invalid = -inf, colour uchar4 RGBA.
"""
from __future__ import annotations

import hashlib

import numpy as np

ROOM_MIN = np.array([-3.0, -1.5, -3.0])
ROOM_MAX = np.array([3.0, 1.5, 3.0])
SPHERES = [(np.array([1.2, -0.9, 1.5]), 0.6), (np.array([-1.5, -1.0, 0.5]), 0.5), (np.array([0.3, -1.1, -1.6]), 0.4)]


def lissajous_pose(i: int, n_total: int = 5000) -> np.ndarray:
    """Camera-to-world 4x4 (float32).  <= ~2 cm / ~1 deg per frame at n_total = 5000."""
    s = 2.0 * np.pi * i / float(n_total)
    pos = np.array([1.2 * np.sin(3.0 * s), 0.25 * np.sin(2.0 * s + 0.3), 1.2 * np.sin(2.0 * s + np.pi / 2)])
    yaw = 4.0 * s + 0.4 * np.sin(5.0 * s)
    pitch = 0.15 * np.sin(3.0 * s + 1.0)
    cy, sy, cp, sp = np.cos(yaw), np.sin(yaw), np.cos(pitch), np.sin(pitch)
    Ry = np.array([[cy, 0, sy], [0, 1, 0], [-sy, 0, cy]])
    Rx = np.array([[1, 0, 0], [0, cp, -sp], [0, sp, cp]])
    T = np.eye(4)
    T[:3, :3] = Ry @ Rx
    T[:3, 3] = pos
    return T.astype(np.float32)


def _raycast(T: np.ndarray, W: int, H: int, fx: float, fy: float, mx: float, my: float):
    u, v = np.meshgrid(np.arange(W, dtype=np.float64), np.arange(H, dtype=np.float64))
    d_cam = np.stack([(u - mx) / fx, (v - my) / fy, np.ones_like(u)], axis=-1)       # z = 1 rays
    R, o = T[:3, :3].astype(np.float64), T[:3, 3].astype(np.float64)
    d = d_cam @ R.T                                                                  # world dirs (z_cam = 1 scale)
    # box (from inside): nearest exit
    with np.errstate(divide="ignore", invalid="ignore"):
        t1 = (ROOM_MIN - o) / d
        t2 = (ROOM_MAX - o) / d
    tfar = np.maximum(t1, t2)
    t_box = np.min(tfar, axis=-1)
    axis = np.argmin(tfar, axis=-1)
    t = t_box.copy()
    obj = np.zeros(t.shape, dtype=np.int32)        # 0 = wall, k = sphere k
    for k, (c, r) in enumerate(SPHERES, start=1):
        oc = o - c
        a = np.sum(d * d, axis=-1)
        b = 2.0 * (d @ oc)
        cc = oc @ oc - r * r
        disc = b * b - 4 * a * cc
        ok = disc > 0
        ts = np.where(ok, (-b - np.sqrt(np.where(ok, disc, 0.0))) / (2 * a), np.inf)
        hit = ok & (ts > 1e-4) & (ts < t)
        t = np.where(hit, ts, t)
        obj = np.where(hit, k, obj)
    p = o + d * t[..., None]
    return t, p, obj, axis                                                           # t == camera-space z


def make_frame(i: int, W: int = 640, H: int = 480, n_total: int = 5000, seed: int = 1234,
               noise: bool = True, dropout: float = 0.02, pose: np.ndarray | None = None):
    """Returns (depth float32 [H,W], color uint8 [H,W,4], pose float32 [4,4])."""
    T = lissajous_pose(i, n_total) if pose is None else np.asarray(pose, dtype=np.float32)
    fx = fy = 525.0 * W / 640.0
    mx, my = (W - 1) / 2.0, (H - 1) / 2.0
    z, p, obj, axis = _raycast(T, W, H, fx, fy, mx, my)
    rng = np.random.Generator(np.random.MT19937(seed + 7919 * i))
    if noise:
        z = z + rng.standard_normal(z.shape) * (0.0012 * z * z)
    depth = z.astype(np.float32)
    if dropout > 0:
        depth[rng.random(z.shape) < dropout] = -np.inf
    depth[~np.isfinite(z)] = -np.inf
    # procedural texture: sinusoids in world coordinates, different phase per surface
    tex = 0.5 + 0.5 * np.sin(7.0 * p[..., 0] + 0.5 * axis) * np.sin(5.0 * p[..., 1] + 1.3) * np.sin(6.0 * p[..., 2] + obj)
    base = np.array([[200, 180, 160], [220, 80, 60], [60, 200, 90], [70, 90, 230]], dtype=np.float64)[obj]
    rgb = np.clip(base * (0.35 + 0.65 * tex[..., None]), 0, 255).astype(np.uint8)
    color = np.concatenate([rgb, np.full((H, W, 1), 255, np.uint8)], axis=-1)
    return depth, np.ascontiguousarray(color), T


def frame_sha256(depth: np.ndarray, color: np.ndarray) -> str:
    h = hashlib.sha256()
    h.update(np.ascontiguousarray(depth).tobytes())
    h.update(np.ascontiguousarray(color).tobytes())
    return h.hexdigest()


def plane_frame(W: int, H: int, z0: float, color=(128, 64, 32)):
    """Fronto-parallel wall at depth z0 (known-answer tests)."""
    depth = np.full((H, W), z0, dtype=np.float32)
    col = np.zeros((H, W, 4), dtype=np.uint8)
    col[..., 0], col[..., 1], col[..., 2], col[..., 3] = color[0], color[1], color[2], 255
    return depth, col
