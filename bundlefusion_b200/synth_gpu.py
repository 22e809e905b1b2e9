"""torch (GPU) version of synth.make_frame for building large frame banks quickly (input generation only -- plumbing, never
inside a timed region).  Same scene, intrinsics and noise model as synth.py; the random stream differs
(torch generator), so use one generator or the other consistently within an experiment."""
from __future__ import annotations

import numpy as np

from . import synth


def _rich_texture(torch, p):
    """synth.rich_texture in torch (same lattice hash, same octaves)"""
    acc = torch.zeros(p.shape[:-1], dtype=torch.float64, device=p.device)
    for o, f in enumerate(synth.RICH_FREQS):
        q = p * f
        i0 = torch.floor(q).to(torch.int64)
        t = q - i0
        t = t * t * (3.0 - 2.0 * t)
        v = torch.zeros_like(acc)
        for dx in (0, 1):
            for dy in (0, 1):
                for dz in (0, 1):
                    w = (t[..., 0] if dx else 1 - t[..., 0]) * (t[..., 1] if dy else 1 - t[..., 1]) * (t[..., 2] if dz else 1 - t[..., 2])
                    h = ((i0[..., 0] + dx) * 73856093) ^ ((i0[..., 1] + dy) * 19349669) ^ ((i0[..., 2] + dz) * 83492791) ^ (1013 * (o + 1))
                    h = (h * 2654435761) & 0xFFFFFFFF
                    h = ((h ^ (h >> 15)) * 2246822519) & 0xFFFFFFFF
                    h = h ^ (h >> 13)
                    v = v + w * ((h & 0xFFFFFF).to(torch.float64) / float(1 << 24))
        acc = acc + (v - 0.5)
    return torch.clamp(0.5 + 0.55 * acc, 0.0, 1.0)


def make_frames(indices, W=640, H=480, n_total=5000, seed=1234, device="cuda:0", noise=True, dropout=0.02, texture="sinusoid"):
    """Returns (depth [B,H,W] float32, color [B,H,W,4] uint8, poses [B,4,4] float32 numpy) on `device`."""
    import torch
    dev = torch.device(device)
    B = len(indices)
    poses = np.stack([synth.lissajous_pose(int(i), n_total) for i in indices]).astype(np.float32)
    T = torch.from_numpy(poses.astype(np.float64)).to(dev)
    fx = fy = 525.0 * W / 640.0
    mx, my = (W - 1) / 2.0, (H - 1) / 2.0
    v, u = torch.meshgrid(torch.arange(H, dtype=torch.float64, device=dev), torch.arange(W, dtype=torch.float64, device=dev), indexing="ij")
    dcam = torch.stack([(u - mx) / fx, (v - my) / fy, torch.ones_like(u)], -1)                  # [H,W,3]
    depth = torch.empty(B, H, W, dtype=torch.float32, device=dev)
    color = torch.empty(B, H, W, 4, dtype=torch.uint8, device=dev)
    gen = torch.Generator(device=dev)
    rmin = torch.tensor(synth.ROOM_MIN, dtype=torch.float64, device=dev)
    rmax = torch.tensor(synth.ROOM_MAX, dtype=torch.float64, device=dev)
    base = torch.tensor([[200, 180, 160], [220, 80, 60], [60, 200, 90], [70, 90, 230]], dtype=torch.float64, device=dev)
    for b in range(B):
        R, o = T[b, :3, :3], T[b, :3, 3]
        d = dcam @ R.T
        t1, t2 = (rmin - o) / d, (rmax - o) / d
        tfar = torch.maximum(t1, t2)
        t, axis = tfar.min(-1)
        obj = torch.zeros_like(axis)
        for k, (c, r) in enumerate(synth.SPHERES, start=1):
            c = torch.tensor(c, dtype=torch.float64, device=dev)
            oc = o - c
            a = (d * d).sum(-1); bq = 2.0 * (d @ oc); cc = (oc @ oc) - r * r
            disc = bq * bq - 4 * a * cc
            ok = disc > 0
            ts = torch.where(ok, (-bq - torch.sqrt(torch.clamp(disc, min=0.0))) / (2 * a), torch.full_like(a, float("inf")))
            hit = ok & (ts > 1e-4) & (ts < t)
            t = torch.where(hit, ts, t); obj = torch.where(hit, torch.full_like(obj, k), obj)
        p = o + d * t[..., None]
        z = t
        gen.manual_seed(seed + 7919 * int(indices[b]))
        if noise:
            z = z + torch.randn(z.shape, generator=gen, device=dev, dtype=torch.float64) * (0.0012 * z * z)
        dep = z.to(torch.float32)
        if dropout > 0:
            dep = torch.where(torch.rand(z.shape, generator=gen, device=dev) < dropout, torch.full_like(dep, float("-inf")), dep)
        depth[b] = dep
        if texture == "rich":
            tex = _rich_texture(torch, p)
        else:
            tex = 0.5 + 0.5 * torch.sin(7.0 * p[..., 0] + 0.5 * axis) * torch.sin(5.0 * p[..., 1] + 1.3) * torch.sin(6.0 * p[..., 2] + obj)
        rgb = torch.clamp(base[obj] * (0.35 + 0.65 * tex[..., None]), 0, 255).to(torch.uint8)
        color[b, ..., :3] = rgb
        color[b, ..., 3] = 255
    return depth, color, poses
