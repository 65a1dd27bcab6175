"""The generators of synth.py restated on torch tensors, so that the large BASELINE configs (C4: 50 M vs 20 M, C5: 200 M vs
200 M points) can be produced directly in HBM in seconds instead of minutes of numpy per rank.

Same counter-based construction: every coordinate is a pure function of (seed, point index, lane).  The integer part
(splitmix64 on int64 with wrap-around, logical shifts emulated by masks) and the 24-bit uniforms are bit-identical to
synth.py; the Gaussian noise goes through fp64 log / cos / sqrt of the device's libm before the final rounding to fp32,
so a coordinate can differ from the numpy generator by one fp32 ulp (tests/test_synth_torch.py bounds it).  Bench input
only — the parity tests keep using synth.py.
"""
import math

import numpy as np
import torch

from . import synth

_MASK = {k: (1 << (64 - k)) - 1 for k in (27, 30, 31, 40)}


def _s64(v):
    """python int (uint64 constant) -> the int64 with the same bit pattern"""
    v &= 0xFFFFFFFFFFFFFFFF
    return v - (1 << 64) if v >= (1 << 63) else v


def _lsr(z, k):
    return (z >> k) & _MASK[k]


def _mix(z):
    z = z + _s64(0x9E3779B97F4A7C15)
    z = (z ^ _lsr(z, 30)) * _s64(0xBF58476D1CE4E5B9)
    z = (z ^ _lsr(z, 27)) * _s64(0x94D049BB133111EB)
    return z ^ _lsr(z, 31)


def _mix_scalar(v):
    v = (v + 0x9E3779B97F4A7C15) & 0xFFFFFFFFFFFFFFFF
    v = ((v ^ (v >> 30)) * 0xBF58476D1CE4E5B9) & 0xFFFFFFFFFFFFFFFF
    v = ((v ^ (v >> 27)) * 0x94D049BB133111EB) & 0xFFFFFFFFFFFFFFFF
    return v ^ (v >> 31)


def uniform24(seed, idx, lane):
    """synth.uniform24 on an int64 index tensor; returns fp32."""
    s = _mix_scalar((int(seed) + int(lane) * 0xD1B54A32D192ED03) & 0xFFFFFFFFFFFFFFFF)
    h = _mix((idx * _s64(0x2545F4914F6CDD1D) + int(lane)) ^ _s64(s))
    return _lsr(h, 40).to(torch.float32) * (1.0 / (1 << 24))


def gaussian(seed, idx, lane):
    u1 = uniform24(seed, idx, lane).to(torch.float64).clamp_min(2.0 ** -25)
    u2 = uniform24(seed, idx, lane + 1).to(torch.float64)
    return torch.sqrt(-2.0 * torch.log(u1)) * torch.cos(2.0 * math.pi * u2)


def _add_noise(cols, seed, idx, sigma):
    if sigma > 0:
        for d in range(3):
            cols[d] = cols[d] + sigma * gaussian(seed, idx, 8 + 2 * d)
    return cols


def _finish(cols, out):
    for d in range(3):
        out[:, d] = cols[d].to(torch.float32).to(torch.float64)


def _chunks(n, start, chunk):
    for b in range(0, n, chunk):
        yield b, min(n, b + chunk), start + b


def uniform_box(n, side, seed, noise_sigma=0.0, start=0, device="cpu", chunk=1 << 24):
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    for b, e, s in _chunks(n, start, chunk):
        idx = torch.arange(s, s + (e - b), dtype=torch.int64, device=device)
        cols = [uniform24(seed, idx, d).to(torch.float64) * float(side) for d in range(3)]
        _finish(_add_noise(cols, seed, idx, noise_sigma), out[b:e])
    return out


def _ground_z(x, y):
    return 0.05 * (torch.sin(0.21 * x) + torch.sin(0.17 * y + 0.5) + torch.sin(0.05 * (x + y)))


def outdoor_scene(n, seed, noise_sigma, start=0, device="cpu", chunk=1 << 24):
    patches = synth._surface_patches_outdoor()
    areas = np.array([p[1] for p in patches])
    cdf = torch.tensor(np.cumsum(areas) / areas.sum(), dtype=torch.float64, device=device)
    kinds = torch.tensor([0 if p[0] == "ground" else (1 if p[0] == "wall" else 2) for p in patches], device=device)
    par_np = np.zeros((len(patches), 5))
    for i, p in enumerate(patches):
        if p[2] is not None:
            par_np[i, :len(p[2])] = p[2]
    par = torch.tensor(par_np, dtype=torch.float64, device=device)
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    for b, e, s in _chunks(n, start, chunk):
        idx = torch.arange(s, s + (e - b), dtype=torch.int64, device=device)
        sel = torch.searchsorted(cdf, uniform24(seed, idx, 3).to(torch.float64), right=True).clamp_max(len(patches) - 1)
        u = uniform24(seed, idx, 0).to(torch.float64)
        v = uniform24(seed, idx, 1).to(torch.float64)
        k = kinds[sel]
        p = par[sel]
        # ground
        gx, gy = u * 200.0, v * 200.0
        gz = _ground_z(gx, gy)
        # walls: (cx, cy, ang, length, height)
        t = (u - 0.5) * p[:, 3]
        wx = p[:, 0] + t * torch.cos(p[:, 2])
        wy = p[:, 1] + t * torch.sin(p[:, 2])
        wz = _ground_z(wx, wy) + v * p[:, 4]
        # trunks: (cx, cy, rad, height)
        ang = u * (2 * math.pi)
        cx = p[:, 0] + p[:, 2] * torch.cos(ang)
        cy = p[:, 1] + p[:, 2] * torch.sin(ang)
        cz = _ground_z(p[:, 0], p[:, 1]) + v * p[:, 3]
        x = torch.where(k == 0, gx, torch.where(k == 1, wx, cx))
        y = torch.where(k == 0, gy, torch.where(k == 1, wy, cy))
        z = torch.where(k == 0, gz, torch.where(k == 1, wz, cz))
        _finish(_add_noise([x, y, z], seed, idx, noise_sigma), out[b:e])
    return out


def indoor_scene(n, seed, noise_sigma, start=0, rooms=10, device="cpu", chunk=1 << 24):
    face_area = np.array([64.0, 64.0, 24.0, 24.0, 24.0, 24.0])
    cdf = torch.tensor(np.cumsum(face_area) / face_area.sum(), dtype=torch.float64, device=device)
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    for b, e, s in _chunks(n, start, chunk):
        idx = torch.arange(s, s + (e - b), dtype=torch.int64, device=device)
        room = (uniform24(seed, idx, 4).to(torch.float64) * (rooms * rooms)).to(torch.int64).clamp_max(rooms * rooms - 1)
        face = torch.searchsorted(cdf, uniform24(seed, idx, 3).to(torch.float64), right=True).clamp_max(5)
        u = uniform24(seed, idx, 0).to(torch.float64)
        v = uniform24(seed, idx, 1).to(torch.float64)
        ox = (room % rooms).to(torch.float64) * 8.0
        oy = (room // rooms).to(torch.float64) * 8.0
        c002 = torch.full_like(u, 0.02)
        c798 = torch.full_like(u, 7.98)
        x = torch.where(face < 2, u * 8.0, torch.where(face == 2, c002, torch.where(face == 3, c798, u * 8.0)))
        y = torch.where(face < 2, v * 8.0, torch.where((face == 2) | (face == 3), u * 8.0, torch.where(face == 4, c002, c798)))
        z = torch.where(face == 0, torch.zeros_like(u), torch.where(face == 1, torch.full_like(u, 3.0), v * 3.0))
        _finish(_add_noise([x + ox, y + oy, z], seed, idx, noise_sigma), out[b:e])
    return out


def site_scene(n, seed, noise_sigma, start=0, device="cpu", chunk=1 << 24):
    b = synth._site_buildings()
    areas = np.array([1000.0 * 1000.0] + [p[3] * p[4] for p in b])
    cdf = torch.tensor(np.cumsum(areas) / areas.sum(), dtype=torch.float64, device=device)
    par_np = np.zeros((len(areas), 5))
    par_np[1:] = np.array(b)
    par = torch.tensor(par_np, dtype=torch.float64, device=device)
    out = torch.empty((n, 3), dtype=torch.float64, device=device)
    for b0, e0, s in _chunks(n, start, chunk):
        idx = torch.arange(s, s + (e0 - b0), dtype=torch.int64, device=device)
        sel = torch.searchsorted(cdf, uniform24(seed, idx, 3).to(torch.float64), right=True).clamp_max(len(areas) - 1)
        u = uniform24(seed, idx, 0).to(torch.float64) + uniform24(seed, idx, 5).to(torch.float64) * 2.0 ** -24
        v = uniform24(seed, idx, 1).to(torch.float64) + uniform24(seed, idx, 6).to(torch.float64) * 2.0 ** -24
        p = par[sel]
        g = sel == 0
        t = (u - 0.5) * p[:, 3]
        x = torch.where(g, u * 1000.0, p[:, 0] + t * torch.cos(p[:, 2]))
        y = torch.where(g, v * 1000.0, p[:, 1] + t * torch.sin(p[:, 2]))
        z = (12.0 * torch.sin(x / 160.0) + 8.0 * torch.sin(y / 115.0 + 0.7) + 5.0 * torch.sin((x + y) / 47.0)
             + 0.4 * torch.sin(x / 3.1) * torch.cos(y / 2.7)) + torch.where(g, torch.zeros_like(v), v * p[:, 4])
        _finish(_add_noise([x, y, z], seed, idx, noise_sigma), out[b0:e0])
    return out


def make_pair(name, scale=1.0, device="cpu"):
    """(est, gt, cfg) as (N, 3) fp64 tensors on `device` — the same clouds as synth.make_pair (see the module note)."""
    cfg = dict(synth.CONFIGS[name])
    n_est = max(1, int(round(cfg["n_est"] * scale)))
    n_gt = max(1, int(round(cfg["n_gt"] * scale)))
    cfg["n_est"], cfg["n_gt"] = n_est, n_gt
    if cfg["kind"] == "box":
        side = synth.box_side_for_density(n_gt)
        cfg["side"] = side
        gt = uniform_box(n_gt, side, synth.GT_SEED, device=device)
        est = uniform_box(n_est, side, synth.EST_SEED, noise_sigma=synth.EST_NOISE_SIGMA, device=device)
    elif cfg["kind"] == "outdoor":
        gt = outdoor_scene(n_gt, synth.GT_SEED, synth.GT_SURFACE_NOISE_SIGMA, device=device)
        est = outdoor_scene(n_est, synth.EST_SEED, synth.EST_NOISE_SIGMA, device=device)
    elif cfg["kind"] == "site":
        gt = site_scene(n_gt, synth.GT_SEED, synth.GT_SURFACE_NOISE_SIGMA, device=device)
        est = site_scene(n_est, synth.EST_SEED, synth.EST_NOISE_SIGMA, device=device)
    else:
        gt = indoor_scene(n_gt, synth.GT_SEED, synth.GT_SURFACE_NOISE_SIGMA, device=device)
        est = indoor_scene(n_est, synth.EST_SEED, synth.EST_NOISE_SIGMA, device=device)
    return est, gt, cfg
