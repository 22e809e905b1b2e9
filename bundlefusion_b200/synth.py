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


RICH_FREQS = (4.0, 9.0, 20.0, 45.0)         # lattice cells per metre of the four octaves of the "rich" texture


def _lattice_hash(ix, iy, iz, salt):
    """integer lattice point -> pseudo-random value in [0, 1); int64 arithmetic, the same formula as synth_gpu"""
    h = (ix * 73856093) ^ (iy * 19349669) ^ (iz * 83492791) ^ salt
    h = (h * 2654435761) & 0xFFFFFFFF
    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
    h = h ^ (h >> 13)
    return (h & 0xFFFFFF).astype(np.float64) / float(1 << 24)


def rich_texture(p: np.ndarray) -> np.ndarray:
    """World-anchored multi-scale value noise in [0, 1] (trilinear, smooth-stepped lattice noise, four octaves): blobs from ~2 cm to ~25 cm,
    so that a difference-of-Gaussians detector finds features at every pyramid level from any viewpoint -- the sinusoid texture of make_frame
    yields ~14 SIFT key points per 320x240 frame, this one hundreds."""
    acc = np.zeros(p.shape[:-1])
    for o, f in enumerate(RICH_FREQS):
        q = p * f
        i0 = np.floor(q).astype(np.int64)
        t = q - i0
        t = t * t * (3.0 - 2.0 * t)
        v = 0.0
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (t[..., 0] if dx else 1 - t[..., 0]) * (t[..., 1] if dy else 1 - t[..., 1]) * (t[..., 2] if dz else 1 - t[..., 2])
                    v = v + w * _lattice_hash(i0[..., 0] + dx, i0[..., 1] + dy, i0[..., 2] + dz, 1013 * (o + 1))
        acc += v - 0.5
    return np.clip(0.5 + 0.55 * acc, 0.0, 1.0)


def make_frame(i: int, W: int = 640, H: int = 480, n_total: int = 5000, seed: int = 1234,
               noise: bool = True, dropout: float = 0.02, pose: np.ndarray | None = None, texture: str = "sinusoid"):
    """Returns (depth float32 [H,W], color uint8 [H,W,4], pose float32 [4,4]).  texture: "sinusoid"  or "rich"
    (rich_texture: enough SIFT features for the frame loop to track on)."""
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
    if texture == "rich":
        tex = rich_texture(p)
    else:
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


# ------------------------------------------------------------------------------------------------
# synthetic bundle-adjustment problems
# ------------------------------------------------------------------------------------------------
def se3_exp(rot, trans) -> np.ndarray:
    """float64 SE(3) exponential T = [exp(w) | V(w) u] (the parametrisation of FL/Solver/LieDerivUtil.h:160-207)."""
    w = np.asarray(rot, np.float64); u = np.asarray(trans, np.float64)
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        R = np.eye(3) + K; Vm = np.eye(3) + 0.5 * K
    else:
        A, B, Cc = np.sin(th) / th, (1 - np.cos(th)) / th ** 2, (1 - np.sin(th) / th) / th ** 2
        R = np.eye(3) + A * K + B * K @ K
        Vm = np.eye(3) + B * K + Cc * K @ K
    T = np.eye(4); T[:3, :3] = R; T[:3, 3] = Vm @ u
    return T


def se3_log(T):
    """float64 inverse of se3_exp -> (rot, trans)."""
    R = T[:3, :3]; t = T[:3, 3]
    c = np.clip((np.trace(R) - 1) / 2, -1, 1); th = np.arccos(c)
    if th < 1e-9:
        w = np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]]) / 2
    else:
        w = th / (2 * np.sin(th)) * np.array([R[2, 1] - R[1, 2], R[0, 2] - R[2, 0], R[1, 0] - R[0, 1]])
    th = np.linalg.norm(w)
    K = np.array([[0, -w[2], w[1]], [w[2], 0, -w[0]], [-w[1], w[0], 0]])
    if th < 1e-9:
        Vm = np.eye(3) + 0.5 * K
    else:
        Vm = np.eye(3) + (1 - np.cos(th)) / th ** 2 * K + (1 - np.sin(th) / th) / th ** 2 * K @ K
    return w, np.linalg.solve(Vm, t)


def make_ba_problem(n_images: int, degree: int = 6, corr_per_pair: int = 25, noise: float = 0.002, outliers: float = 0.0,
                    perturb_rot: float = 0.02, perturb_trans: float = 0.03, seed: int = 7, stride: int = 10):
    """Keyframes on the Lissajous path (every `stride`-th frame), a co-visibility graph (each image paired with its
    `degree` successors), `corr_per_pair` exact 3-D correspondences per pair observed in both camera frames with
    Gaussian noise (and a fraction of gross outliers).  Returns dict with EntryJ-layout arrays and float32 poses:
    corr (C,8) uint32 view-compatible, gt (N,4,4), init_rot/init_trans (N,3) float32 (image 0 = ground truth)."""
    rng = np.random.Generator(np.random.MT19937(seed))
    gt = np.stack([lissajous_pose(stride * k).astype(np.float64) for k in range(n_images)])
    pairs = [(i, j) for i in range(n_images) for j in range(i + 1, min(n_images, i + 1 + degree))]
    C = len(pairs) * corr_per_pair
    corr = np.zeros(C, dtype=[("i", "<u4"), ("j", "<u4"), ("pi", "<f4", 3), ("pj", "<f4", 3)])
    k = 0
    for (i, j) in pairs:
        # points in front of camera i, 0.5 .. 3 m away
        z = rng.uniform(0.5, 3.0, corr_per_pair)
        xy = rng.uniform(-0.5, 0.5, (corr_per_pair, 2)) * z[:, None]
        pi = np.concatenate([xy, z[:, None]], 1)
        X = pi @ gt[i][:3, :3].T + gt[i][:3, 3]
        Tj_inv = np.linalg.inv(gt[j])
        pj = X @ Tj_inv[:3, :3].T + Tj_inv[:3, 3]
        pi_n = pi + rng.standard_normal(pi.shape) * noise
        pj_n = pj + rng.standard_normal(pj.shape) * noise
        if outliers > 0:
            bad = rng.random(corr_per_pair) < outliers
            pj_n[bad] += rng.standard_normal((int(bad.sum()), 3)) * 0.3
        sl = slice(k, k + corr_per_pair)
        corr["i"][sl], corr["j"][sl], corr["pi"][sl], corr["pj"][sl] = i, j, pi_n, pj_n
        k += corr_per_pair
    rot = np.zeros((n_images, 3), np.float32); trans = np.zeros((n_images, 3), np.float32)
    for n in range(n_images):
        T = gt[n]
        if n > 0:
            T = se3_exp(rng.standard_normal(3) * perturb_rot, rng.standard_normal(3) * perturb_trans) @ T
        w, u = se3_log(T)
        rot[n], trans[n] = w, u
    return {"corr": corr, "gt": gt, "init_rot": rot, "init_trans": trans, "pairs": pairs}


def ba_reference_f64(corr, rot0, trans0, n_gn: int = 10, w_sparse: float = 1.0):
    """Independent float64 Gauss-Newton (dense normal equations, numpy) on the sparse energy of
    FL/Solver/SolverBundlingEquationsLie.h:42-57 with left-multiplicative updates; image 0 fixed.
    Returns (rot, trans) float64 [N,3]."""
    N = len(rot0)
    T = [se3_exp(rot0[k], trans0[k]) for k in range(N)]
    valid = corr["i"] != 0xFFFFFFFF
    ci, cj = corr["i"][valid].astype(int), corr["j"][valid].astype(int)
    pi, pj = corr["pi"][valid].astype(np.float64), corr["pj"][valid].astype(np.float64)
    for _ in range(n_gn):
        Tm = np.stack(T)
        Pi = np.einsum("cab,cb->ca", Tm[ci][:, :3, :3], pi) + Tm[ci][:, :3, 3]
        Pj = np.einsum("cab,cb->ca", Tm[cj][:, :3, :3], pj) + Tm[cj][:, :3, 3]
        r = Pi - Pj

        def J(P):  # d(exp(e) P)/de = [-[P]x | I], columns (rot, trans)
            Jm = np.zeros((len(P), 3, 6))
            Jm[:, 0, 1], Jm[:, 0, 2] = P[:, 2], -P[:, 1]
            Jm[:, 1, 0], Jm[:, 1, 2] = -P[:, 2], P[:, 0]
            Jm[:, 2, 0], Jm[:, 2, 1] = P[:, 1], -P[:, 0]
            Jm[:, 0, 3] = Jm[:, 1, 4] = Jm[:, 2, 5] = 1
            return Jm
        Ji, Jj = J(Pi), -J(Pj)
        H = np.zeros((6 * N, 6 * N)); g = np.zeros(6 * N)
        for c in range(len(ci)):
            a, b = 6 * ci[c], 6 * cj[c]
            H[a:a + 6, a:a + 6] += Ji[c].T @ Ji[c]; H[b:b + 6, b:b + 6] += Jj[c].T @ Jj[c]
            H[a:a + 6, b:b + 6] += Ji[c].T @ Jj[c]; H[b:b + 6, a:a + 6] += Jj[c].T @ Ji[c]
            g[a:a + 6] += Ji[c].T @ r[c]; g[b:b + 6] += Jj[c].T @ r[c]
        H *= w_sparse; g *= w_sparse
        d = np.zeros(6 * N)
        d[6:] = np.linalg.solve(H[6:, 6:] + 1e-12 * np.eye(6 * N - 6), -g[6:])
        for k in range(1, N):
            T[k] = se3_exp(d[6 * k:6 * k + 3], d[6 * k + 3:6 * k + 6]) @ T[k]
        if np.abs(d).max() < 1e-10:
            break
    out = [se3_log(Tk) for Tk in T]
    return np.array([o[0] for o in out]), np.array([o[1] for o in out])


# ------------------------------------------------------------------------------------------------
# dense cache frames (FL/CUDACache.cpp:45-86 without the optional bilateral / Gaussian pre-filters)
# ------------------------------------------------------------------------------------------------
def cache_intrinsics(W_in=640, H_in=480, cw=80, ch=60):
    """FL/CUDACache.cpp:20-24: intrinsics of the down-sampled cache from the input intrinsics."""
    fx = 525.0 * W_in / 640.0; fy = fx
    mx, my = (W_in - 1) / 2.0, (H_in - 1) / 2.0
    return (np.float32(fx * cw / W_in), np.float32(fy * ch / H_in), np.float32(mx * (cw - 1) / (W_in - 1)), np.float32(my * (ch - 1) / (H_in - 1)))


def make_cache_frame(depth: np.ndarray, color: np.ndarray, cw: int = 80, ch: int = 60) -> dict:
    """One CUDACachedFrame (FL/CUDACacheUtil.h:41-53) as host arrays: depth [ch,cw] f32, campos [ch,cw,4] f32, normals [ch,cw,4] f32,
    normalsU [ch,cw,4] u8, intensity [ch,cw] f32, intensityDerivs [ch,cw,2] f32.  Invalid = -inf (0 for the uchar4 normals)."""
    F = np.float32
    H_in, W_in = depth.shape
    fx = fy = F(525.0 * W_in / 640.0); mx, my = F((W_in - 1) / 2.0), F((H_in - 1) / 2.0)
    u, v = np.meshgrid(np.arange(W_in, dtype=F), np.arange(H_in, dtype=F))
    valid = depth != -np.inf
    cam = np.full((H_in, W_in, 4), -np.inf, F)
    with np.errstate(invalid="ignore"):
        cam[..., 0] = np.where(valid, (u - mx) / fx * depth, -np.inf)
        cam[..., 1] = np.where(valid, (v - my) / fy * depth, -np.inf)
        cam[..., 2] = np.where(valid, depth, -np.inf)
        cam[..., 3] = np.where(valid, F(1.0), -np.inf)
    # computeNormals (FL/CUDAImageUtil.cu:404-431): -normalize((P(x,y+1)-P(x,y-1)) x (P(x+1,y)-P(x-1,y))), all five valid
    nrm = np.full((H_in, W_in, 4), -np.inf, F)
    P = cam[..., :3]
    ok = valid.copy(); ok[1:-1, 1:-1] &= valid[2:, 1:-1] & valid[:-2, 1:-1] & valid[1:-1, 2:] & valid[1:-1, :-2]
    ok[0, :] = ok[-1, :] = False; ok[:, 0] = ok[:, -1] = False
    with np.errstate(invalid="ignore", divide="ignore"):
        a = P[2:, 1:-1] - P[:-2, 1:-1]; b = P[1:-1, 2:] - P[1:-1, :-2]
        n = np.cross(a, b).astype(F)
        l = np.sqrt((n * n).sum(-1)).astype(F)
        inner = ok[1:-1, 1:-1] & (l > 0)
        nn = np.where(inner[..., None], n / -l[..., None], -np.inf)
    nrm[1:-1, 1:-1, :3] = nn
    nrm[1:-1, 1:-1, 3] = np.where(inner, F(0.0), -np.inf)
    # nearest-neighbour resample with scale (in-1)/(out-1)  (FL/CUDAImageUtil.cu:93-110)
    xs = (np.arange(cw, dtype=F) * F((W_in - 1) / (cw - 1)) + F(0.5)).astype(np.int64)
    ys = (np.arange(ch, dtype=F) * F((H_in - 1) / (ch - 1)) + F(0.5)).astype(np.int64)
    yy, xx = np.meshgrid(ys, xs, indexing="ij")
    out = {"depth": np.ascontiguousarray(depth[yy, xx]), "campos": np.ascontiguousarray(cam[yy, xx]), "normals": np.ascontiguousarray(nrm[yy, xx])}
    nd = out["normals"]
    nu = np.zeros((ch, cw, 4), np.uint8)
    vn = nd[..., 0] != -np.inf
    with np.errstate(invalid="ignore"):
        q = (np.where(vn[..., None], nd[..., :3], 0) + F(1.0)) / F(2.0)
        r = q * F(255)
        nu[..., :3] = np.where(vn[..., None], np.where(r >= 0, np.floor(r + F(0.5)), np.ceil(r - F(0.5))), 0).astype(np.uint8)   # (uchar)round(p*255)
    out["normalsU"] = nu
    c = color[yy, xx].astype(F)
    inten = ((F(0.299) * c[..., 0] + F(0.587) * c[..., 1] + F(0.114) * c[..., 2]) / F(255.0)).astype(F)
    out["intensity"] = np.ascontiguousarray(inten)
    # computeIntensityDerivatives (FL/CUDAImageUtil.cu:260-300): Sobel / 8 on the interior
    dv = np.full((ch, cw, 2), -np.inf, F)
    I = inten
    ru = (-I[:-2, :-2] + I[:-2, 2:] - F(2) * I[1:-1, :-2] + F(2) * I[1:-1, 2:] - I[2:, :-2] + I[2:, 2:]) / F(8.0)
    rv = (-I[:-2, :-2] - F(2) * I[:-2, 1:-1] - I[:-2, 2:] + I[2:, :-2] + F(2) * I[2:, 1:-1] + I[2:, 2:]) / F(8.0)
    dv[1:-1, 1:-1, 0] = ru; dv[1:-1, 1:-1, 1] = rv
    out["intensityDerivs"] = dv
    return out


def make_dense_ba_problem(n_images: int = 11, stride: int = 3, start: int = 100, corr_per_pair: int = 25, noise: float = 0.002,
                          perturb_rot: float = 0.01, perturb_trans: float = 0.015, seed: int = 17, W: int = 640, H: int = 480):
    """A local-chunk problem (FL/OnlineBundler.cpp:41: 11 frames) with BOTH terms: sparse correspondences between all pairs and
    dense 80x60 cache frames rendered from the synthetic room at the ground-truth poses."""
    # noise-free depth: the reference smooths depth (bilateral, sigma_d 1.0 sigma_r 0.05) before it takes normals; that filter
    # belongs to row a20 (CUDACache::storeFrame) and is not part of this generator yet
    frames = [make_frame(start + stride * k, W, H, noise=False, dropout=0.0) for k in range(n_images)]
    gt = np.stack([f[2].astype(np.float64) for f in frames])
    caches = [make_cache_frame(f[0], f[1]) for f in frames]
    rng = np.random.Generator(np.random.MT19937(seed))
    pairs = [(i, j) for i in range(n_images) for j in range(i + 1, n_images)]
    corr = np.zeros(len(pairs) * corr_per_pair, dtype=[("i", "<u4"), ("j", "<u4"), ("pi", "<f4", 3), ("pj", "<f4", 3)])
    k = 0
    for (i, j) in pairs:
        z = rng.uniform(0.6, 2.5, corr_per_pair); xy = rng.uniform(-0.4, 0.4, (corr_per_pair, 2)) * z[:, None]
        pi = np.concatenate([xy, z[:, None]], 1)
        X = pi @ gt[i][:3, :3].T + gt[i][:3, 3]
        Tj = np.linalg.inv(gt[j]); pj = X @ Tj[:3, :3].T + Tj[:3, 3]
        sl = slice(k, k + corr_per_pair)
        corr["i"][sl], corr["j"][sl] = i, j
        corr["pi"][sl] = pi + rng.standard_normal(pi.shape) * noise; corr["pj"][sl] = pj + rng.standard_normal(pj.shape) * noise
        k += corr_per_pair
    rot = np.zeros((n_images, 3), np.float32); trans = np.zeros((n_images, 3), np.float32)
    for n in range(n_images):
        T = gt[n] if n == 0 else se3_exp(rng.standard_normal(3) * perturb_rot, rng.standard_normal(3) * perturb_trans) @ gt[n]
        rot[n], trans[n] = se3_log(T)
    return {"corr": corr, "gt": gt, "init_rot": rot, "init_trans": trans, "caches": caches, "pairs": pairs, "intrinsics": cache_intrinsics(W, H)}


# ---- synthetic SIFT descriptors (row a18) ------------------------------------------------------------------------------
def quantize_descriptors(v: np.ndarray) -> np.ndarray:
    """SiftGPU's unsigned-char descriptor convention: non-negative, L2-normalised to 512, rounded, clipped to 255."""
    v = np.maximum(np.asarray(v, np.float64), 0.0)
    n = np.linalg.norm(v, axis=1, keepdims=True)
    n[n == 0] = 1.0
    return np.clip(np.floor(v / n * 512.0 + 0.5), 0, 255).astype(np.uint8)


def make_sift_descriptors(n: int, seed: int = 0) -> np.ndarray:
    """n gradient-histogram-like descriptors [n,128] uint8 (sparse exponential bins, the usual 0.2 clamp before renormalising)."""
    rng = np.random.default_rng(seed)
    v = rng.exponential(1.0, (n, 128)) * (rng.random((n, 128)) < 0.6)
    v /= np.maximum(np.linalg.norm(v, axis=1, keepdims=True), 1e-12)
    v = np.minimum(v, 0.2)
    return quantize_descriptors(v)


def make_sift_pair(n1: int, n2: int, n_common: int, noise: float = 0.05, seed: int = 0):
    """Two descriptor sets sharing n_common features (re-observed with noise, in shuffled order).
    Returns (des1 [n1,128] u8, des2 [n2,128] u8, truth [n_common,2] = (index in 1, index in 2))."""
    rng = np.random.default_rng(seed + 7919)
    n_common = min(n_common, n1, n2)
    d1 = make_sift_descriptors(n1, seed)
    base = d1[:n_common].astype(np.float64)
    obs = quantize_descriptors(base + rng.normal(0.0, noise * 512.0 / np.sqrt(128.0), base.shape))
    d2 = np.concatenate([obs, make_sift_descriptors(n2 - n_common, seed + 1)]) if n2 > n_common else obs
    p1, p2 = rng.permutation(n1), rng.permutation(n2)
    inv1, inv2 = np.argsort(p1), np.argsort(p2)
    truth = np.stack([inv1[:n_common], inv2[:n_common]], 1)
    return np.ascontiguousarray(d1[p1]), np.ascontiguousarray(d2[p2]), truth


# ---- synthetic key points + raw matches for the Kabsch filter (row a19) --------------------------------------------------
def make_filter_problem(n_pairs: int = 6, n_inliers: int = 40, n_outliers: int = 12, noise: float = 0.002, seed: int = 0, W: int = 640, H: int = 480):
    """The current frame plus n_pairs earlier frames observing the same 3-D points.  Returns key points [K,4] (x, y, scale, depth), the
    manager-layout raw matches of every pair (earlier frame p -> current frame), already sorted by distance, the inverse intrinsics
    and the ground-truth transforms T_p with T_p * X_p = X_cur."""
    rng = np.random.default_rng(seed)
    fx = 525.0 * W / 640.0; mx, my = (W - 1) / 2.0, (H - 1) / 2.0
    K = np.array([[fx, 0, mx, 0], [0, fx, my, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    Kinv = np.linalg.inv(K)
    n = n_inliers + n_outliers
    P = n_pairs + 1
    cur = n_pairs                                               # the current frame is the last image
    keys = np.zeros((P * n, 4), np.float32)
    num = np.zeros(P, np.int32); dists = np.full((P, 128), 999.0, np.float32); idxs = np.full((P, 128, 2), 0xFFFFFFFF, np.uint32)
    # points in the current frame: spread over the image, well separated (> 5 px), depths 0.8 .. 3 m
    gx, gy = np.meshgrid(np.linspace(30, W - 30, 10), np.linspace(30, H - 30, 8))
    pix = np.stack([gx.ravel(), gy.ravel()], 1)[rng.permutation(80)[:n]] + rng.uniform(-6, 6, (n, 2))
    z = rng.uniform(0.8, 3.0, n)
    Xc = np.stack([(pix[:, 0] - mx) / fx * z, (pix[:, 1] - my) / fx * z, z], 1)
    keys[cur * n:(cur + 1) * n] = np.c_[pix, np.ones(n), z]
    T_gt = np.zeros((P, 4, 4)); T_gt[cur] = np.eye(4)
    for p in range(n_pairs):
        T = se3_exp(rng.standard_normal(3) * 0.08, rng.standard_normal(3) * 0.15)          # X_cur = T X_p
        T_gt[p] = T
        Ti = np.linalg.inv(T)
        Xp = Xc @ Ti[:3, :3].T + Ti[:3, 3] + rng.standard_normal((n, 3)) * noise
        Xp[n_inliers:] += rng.uniform(0.15, 0.5, (n_outliers, 3)) * rng.choice([-1, 1], (n_outliers, 3))    # gross outliers
        keys[p * n:(p + 1) * n] = np.c_[Xp[:, 0] / Xp[:, 2] * fx + mx, Xp[:, 1] / Xp[:, 2] * fx + my, np.ones(n), Xp[:, 2]]
        order = rng.permutation(n)
        num[p] = n
        dists[p, :n] = np.sort(rng.uniform(0.05, 0.6, n)).astype(np.float32)
        idxs[p, :n, 0] = p * n + order; idxs[p, :n, 1] = cur * n + order
    return {"keys": keys, "num": num, "dists": dists, "idxs": idxs, "Kinv": Kinv.astype(np.float32), "T_gt": T_gt, "cur": cur, "n": n,
            "n_inliers": n_inliers, "P": P}


# ---- problems for the surface-area and dense-verify match filters (row a19) ----------------------------------------------------
def make_area_problem(seed: int = 0, W: int = 640, H: int = 480):
    """Filtered matches of 8 earlier frames against the current one (index 8), manager layout, with patches of controlled size:
    pair 0: 20 matches spread over the image in both images (large area)      pair 1: 6 matches inside a 4 x 4 pixel patch in both (tiny)
    pair 2: tiny in image p, spread in the current image (kept: BOTH must be small)   pair 3: no matches
    pair 4: 25 matches spread        pair 5: 3 matches spread       pair 6: 2 matches (degenerate: extent 0 in one axis)
    pair 7: 12 matches on a thin line (area ~ 0)."""
    rng = np.random.default_rng(seed)
    fx = 525.0 * W / 640.0; mx, my = (W - 1) / 2.0, (H - 1) / 2.0
    K = np.array([[fx, 0, mx, 0], [0, fx, my, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float64)
    P, cur = 9, 8
    counts = [20, 6, 8, 0, 25, 3, 2, 12]
    keys, num, fidx = [], np.zeros(P, np.int32), np.full((P, 25, 2), 0xFFFFFFFF, np.uint32)

    def spread(n):
        return np.c_[rng.uniform(40, W - 40, n), rng.uniform(40, H - 40, n), np.ones(n), rng.uniform(0.8, 3.0, n)]

    def tiny(n):
        c = rng.uniform(100, 300, 2)
        return np.c_[c[0] + rng.uniform(-2, 2, n), c[1] + rng.uniform(-2, 2, n), np.ones(n), 1.5 + rng.uniform(-0.002, 0.002, n)]

    def line(n):
        t = np.linspace(0, 1, n)
        return np.c_[100 + 300 * t, 120 + 200 * t, np.ones(n), np.full(n, 2.0)]

    kinds = [(spread, spread), (tiny, tiny), (tiny, spread), None, (spread, spread), (spread, spread), (spread, spread), (line, line)]
    for p, (n, kind) in enumerate(zip(counts, kinds)):
        num[p] = n
        if n == 0:
            continue
        a, b = kind[0](n), kind[1](n)
        base = sum(len(k) for k in keys)
        keys += [a, b]
        fidx[p, :n, 0] = base + np.arange(n); fidx[p, :n, 1] = base + n + np.arange(n)
    return {"keys": np.concatenate(keys).astype(np.float32), "num": num, "fidx": fidx, "Kinv": np.linalg.inv(K).astype(np.float32), "cur": cur, "P": P}


def make_dense_verify_problem(n_prev: int = 5, stride: int = 4, start: int = 200, W: int = 640, H: int = 480):
    """Cached 80x60 frames of the synthetic room at n_prev earlier poses plus the current one (last), and per pair a transform frame p ->
    current frame: exact for even p, grossly wrong (0.35 m / 12 degrees off) for odd p."""
    frames = [make_frame(start + stride * k, W, H, noise=False, dropout=0.0) for k in range(n_prev + 1)]
    gt = np.stack([f[2].astype(np.float64) for f in frames])
    caches = [make_cache_frame(f[0], f[1]) for f in frames]
    cur = n_prev
    T = np.zeros((n_prev + 1, 4, 4), np.float32)
    for p in range(n_prev + 1):
        rel = np.linalg.inv(gt[cur]) @ gt[p]
        if p % 2 == 1:
            rel = se3_exp(np.array([0.0, 0.21, 0.0]), np.array([0.35, 0.0, 0.1])) @ rel
        T[p] = rel
    fx, fy, mx, my = cache_intrinsics(W, H)
    K = np.array([[fx, 0, mx, 0], [0, fy, my, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    return {"caches": caches, "T": T, "K": K, "cur": cur, "P": n_prev + 1, "W": 80, "H": 60, "gt": gt}


def make_fuse_problem(n_images: int = 6, n_points: int = 140, key_stride: int = 256, seed: int = 0, outlier_frac: float = 0.1, invalid_frac: float = 0.05):
    """A solved chunk for SIFTImageManager::fuseToGlobal: n_points 3-D points seen by random subsets of n_images cameras near the origin, one key
    per observation (image-major global index image * key_stride + key), correspondences between the observations of a point in image pairs
    (a random subset of the pairs, so that tracks have to be chained), some of them displaced by > 3 cm (they join tracks without contributing a
    position) and some invalidated.  Returns corr (EntryJ), keyIdx, transforms, keys, descs, numKeys, K."""
    rng = np.random.default_rng(seed)
    fx = 525.0; K = np.array([[fx, 0, 319.5, 0], [0, fx, 239.5, 0], [0, 0, 1, 0], [0, 0, 0, 1]], np.float32)
    T = np.stack([se3_exp(rng.standard_normal(3) * 0.03, rng.standard_normal(3) * 0.05) for _ in range(n_images)]).astype(np.float32)
    T[0] = np.eye(4, dtype=np.float32)
    pts = np.c_[rng.uniform(-1, 1, n_points), rng.uniform(-0.7, 0.7, n_points), rng.uniform(1.0, 3.0, n_points)]
    keys = np.zeros((n_images * key_stride, 4), np.float32); descs = rng.integers(0, 256, (n_images * key_stride, 128)).astype(np.uint8)
    num = np.zeros(n_images, np.int32)
    obs = {}                                                  # (point, image) -> (key index, camera-space position)
    for i in range(n_images):
        seen = np.nonzero(rng.random(n_points) < 0.6)[0]
        rng.shuffle(seen)
        Tinv = np.linalg.inv(T[i].astype(np.float64))
        for p in seen[: key_stride]:
            c = (Tinv @ np.r_[pts[p], 1.0])[:3].astype(np.float32)
            k = int(num[i]); num[i] += 1
            keys[i * key_stride + k] = (fx * c[0] / c[2] + 319.5, fx * c[1] / c[2] + 239.5, float(rng.choice([3.2, 6.4, 12.8])), c[2])
            obs[(int(p), i)] = (i * key_stride + k, c)
    dt = np.dtype([("i", "<u4"), ("j", "<u4"), ("pi", "<f4", (3,)), ("pj", "<f4", (3,))])
    rows, kidx = [], []
    for i in range(n_images):
        for j in range(i + 1, n_images):
            if rng.random() < 0.35:
                continue                                       # this image pair was not matched
            for p in range(n_points):
                if (p, i) in obs and (p, j) in obs and rng.random() < 0.8:
                    (ki, ci), (kj, cj) = obs[(p, i)], obs[(p, j)]
                    cj2 = cj + (np.array([0.06, 0.0, 0.02], np.float32) if rng.random() < outlier_frac else 0)
                    rows.append((i, j, ci, cj2.astype(np.float32))); kidx.append((ki, kj))
    corr = np.array(rows, dtype=dt); kidx = np.array(kidx, np.uint32)
    bad = rng.random(len(corr)) < invalid_frac
    corr["i"][bad] = 0xFFFFFFFF; corr["j"][bad] = 0xFFFFFFFF
    return {"corr": corr, "keyIdx": kidx, "transforms": T, "keys": keys, "descs": descs, "numKeys": num, "K": K, "keyStride": key_stride}
