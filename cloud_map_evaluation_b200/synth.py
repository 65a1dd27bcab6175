"""Synthetic cloud generators for the BASELINE.json configs.

Counter-based RNG: every coordinate is a pure function of (seed, point index, lane), so any index range can be
generated independently (shards can self-generate) and the order is reproducible.  All coordinates are produced as
fp32 and widened to fp64 — the reference's storage type (`std::vector<Eigen::Vector3d>`) — so every implementation
sees identical values.
"""
import numpy as np

GT_SEED = 20250001
EST_SEED = 20250002

_M64 = np.uint64(0xFFFFFFFFFFFFFFFF)


def _mix(z):
    """splitmix64 finaliser on uint64 arrays."""
    z = (z + np.uint64(0x9E3779B97F4A7C15)) & _M64
    z = ((z ^ (z >> np.uint64(30))) * np.uint64(0xBF58476D1CE4E5B9)) & _M64
    z = ((z ^ (z >> np.uint64(27))) * np.uint64(0x94D049BB133111EB)) & _M64
    return z ^ (z >> np.uint64(31))


def uniform24(seed, idx, lane):
    """24-bit uniform in [0,1) as fp32; idx is a uint64 array of point indices."""
    with np.errstate(over="ignore"):
        s = _mix(np.uint64(seed) + np.uint64(lane) * np.uint64(0xD1B54A32D192ED03))
        h = _mix(s ^ (idx * np.uint64(0x2545F4914F6CDD1D) + np.uint64(lane)))
    return ((h >> np.uint64(40)).astype(np.float32) * np.float32(1.0 / (1 << 24))).astype(np.float32)


def gaussian(seed, idx, lane):
    """Standard normal (Box-Muller, fp64) from lanes (lane, lane+1)."""
    u1 = uniform24(seed, idx, lane).astype(np.float64)
    u2 = uniform24(seed, idx, lane + 1).astype(np.float64)
    u1 = np.maximum(u1, 2.0 ** -25)
    return np.sqrt(-2.0 * np.log(u1)) * np.cos(2.0 * np.pi * u2)


def _finish(xyz32):
    return np.ascontiguousarray(xyz32.astype(np.float32).astype(np.float64))


def uniform_box(n, side, seed, noise_sigma=0.0, start=0, origin=(0.0, 0.0, 0.0)):
    """n points uniform in a cube of edge `side` (+ optional isotropic Gaussian noise), indices start..start+n."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    cols = []
    for d in range(3):
        c = uniform24(seed, idx, d).astype(np.float64) * float(side) + float(origin[d])
        if noise_sigma > 0:
            c = c + noise_sigma * gaussian(seed, idx, 8 + 2 * d)
        cols.append(c)
    return _finish(np.stack(cols, axis=1))


def box_side_for_density(n, rho=12500.0):
    """Cube edge holding n points at rho pts/m^3 (100k -> 2.0 m, 1M -> 4.31 m, 10M -> 9.28 m)."""
    return float((n / rho) ** (1.0 / 3.0))


def _surface_patches_outdoor():
    """C4 scene description: ground plane + 40 wall rectangles + 500 cylindrical trunks (area-weighted sampling)."""
    rng = np.random.RandomState(4242)
    patches = [("ground", 200.0 * 200.0, None)]
    for _ in range(40):
        cx, cy = rng.uniform(10, 190, 2)
        ang = rng.uniform(0, np.pi)
        length = rng.uniform(8, 30)
        height = rng.uniform(3, 10)
        patches.append(("wall", length * height, (cx, cy, ang, length, height)))
    for _ in range(500):
        cx, cy = rng.uniform(5, 195, 2)
        rad = rng.uniform(0.15, 0.5)
        height = rng.uniform(2, 8)
        patches.append(("trunk", 2 * np.pi * rad * height, (cx, cy, rad, height)))
    return patches


def _ground_z(x, y):
    return 0.05 * (np.sin(0.21 * x) + np.sin(0.17 * y + 0.5) + np.sin(0.05 * (x + y)))


def outdoor_scene(n, seed, noise_sigma, start=0):
    """Newer-College-scale synthetic surfaces (config C4).  noise_sigma > 0 keeps neighbourhood covariances
    non-singular (exact planes make det(cov) a rounding-noise quantity, on which no two implementations agree)."""
    patches = _surface_patches_outdoor()
    areas = np.array([p[1] for p in patches])
    cdf = np.cumsum(areas) / areas.sum()
    idx = np.arange(start, start + n, dtype=np.uint64)
    sel = np.searchsorted(cdf, uniform24(seed, idx, 3).astype(np.float64), side="right")
    sel = np.minimum(sel, len(patches) - 1)
    u = uniform24(seed, idx, 0).astype(np.float64)
    v = uniform24(seed, idx, 1).astype(np.float64)
    x = np.empty(n); y = np.empty(n); z = np.empty(n)
    g = sel == 0
    x[g] = u[g] * 200.0; y[g] = v[g] * 200.0; z[g] = _ground_z(x[g], y[g])
    kinds = np.array([0 if p[0] == "ground" else (1 if p[0] == "wall" else 2) for p in patches])
    par = np.zeros((len(patches), 5))
    for i, p in enumerate(patches):
        if p[2] is not None:
            par[i, :len(p[2])] = p[2]
    w = kinds[sel] == 1
    pw = par[sel[w]]
    t = (u[w] - 0.5) * pw[:, 3]
    x[w] = pw[:, 0] + t * np.cos(pw[:, 2]); y[w] = pw[:, 1] + t * np.sin(pw[:, 2])
    z[w] = _ground_z(x[w], y[w]) + v[w] * pw[:, 4]
    c = kinds[sel] == 2
    pc = par[sel[c]]
    ang = u[c] * 2 * np.pi
    x[c] = pc[:, 0] + pc[:, 2] * np.cos(ang); y[c] = pc[:, 1] + pc[:, 2] * np.sin(ang)
    z[c] = _ground_z(pc[:, 0], pc[:, 1]) + v[c] * pc[:, 3]
    xyz = np.stack([x, y, z], axis=1)
    if noise_sigma > 0:
        for d in range(3):
            xyz[:, d] += noise_sigma * gaussian(seed, idx, 8 + 2 * d)
    return _finish(xyz)


def indoor_scene(n, seed, noise_sigma, start=0, rooms=10):
    """rooms x rooms grid of 8 x 8 x 3 m rooms: floor, ceiling, 4 walls each (config C5)."""
    idx = np.arange(start, start + n, dtype=np.uint64)
    face_area = np.array([64.0, 64.0, 24.0, 24.0, 24.0, 24.0])
    cdf = np.cumsum(face_area) / face_area.sum()
    room = np.minimum((uniform24(seed, idx, 4).astype(np.float64) * rooms * rooms).astype(np.int64), rooms * rooms - 1)
    face = np.minimum(np.searchsorted(cdf, uniform24(seed, idx, 3).astype(np.float64), side="right"), 5)
    u = uniform24(seed, idx, 0).astype(np.float64)
    v = uniform24(seed, idx, 1).astype(np.float64)
    ox = (room % rooms) * 8.0
    oy = (room // rooms) * 8.0
    x = np.where(face < 2, u * 8.0, np.where(face == 2, 0.02, np.where(face == 3, 7.98, u * 8.0)))
    y = np.where(face < 2, v * 8.0, np.where((face == 2) | (face == 3), u * 8.0, np.where(face == 4, 0.02, 7.98)))
    z = np.where(face == 0, 0.0, np.where(face == 1, 3.0, v * 3.0))
    xyz = np.stack([x + ox, y + oy, z], axis=1)
    if noise_sigma > 0:
        for d in range(3):
            xyz[:, d] += noise_sigma * gaussian(seed, idx, 8 + 2 * d)
    return _finish(xyz)


def _site_buildings():
    rng = np.random.RandomState(777)
    out = []
    for _ in range(300):
        cx, cy = rng.uniform(30, 970, 2)
        ang = rng.uniform(0, np.pi)
        out.append((cx, cy, ang, rng.uniform(10, 60), rng.uniform(4, 25)))
    return out


def _site_z(x, y):
    return 12.0 * np.sin(x / 160.0) + 8.0 * np.sin(y / 115.0 + 0.7) + 5.0 * np.sin((x + y) / 47.0) + 0.4 * np.sin(x / 3.1) * np.cos(y / 2.7)


def site_scene(n, seed, noise_sigma, start=0):
    """Site-scale survey (not a BASELINE config; exercises the sparse cell table): a 1 km x 1 km
    terrain with ~50 m of relief plus 300 building facades, area-uniform sampling."""
    b = _site_buildings()
    areas = np.array([1000.0 * 1000.0] + [p[3] * p[4] for p in b])
    cdf = np.cumsum(areas) / areas.sum()
    idx = np.arange(start, start + n, dtype=np.uint64)
    sel = np.minimum(np.searchsorted(cdf, uniform24(seed, idx, 3).astype(np.float64), side="right"), len(areas) - 1)
    u = uniform24(seed, idx, 0).astype(np.float64)
    v = uniform24(seed, idx, 1).astype(np.float64)
    # 24-bit uniforms over 1 km are 6e-5 m apart: add a second draw for the low bits
    u = u + uniform24(seed, idx, 5).astype(np.float64) * 2.0 ** -24
    v = v + uniform24(seed, idx, 6).astype(np.float64) * 2.0 ** -24
    par = np.zeros((len(areas), 5))
    par[1:] = np.array(b)
    p = par[sel]
    g = sel == 0
    t = (u - 0.5) * p[:, 3]
    x = np.where(g, u * 1000.0, p[:, 0] + t * np.cos(p[:, 2]))
    y = np.where(g, v * 1000.0, p[:, 1] + t * np.sin(p[:, 2]))
    z = _site_z(x, y) + np.where(g, 0.0, v * p[:, 4])
    xyz = np.stack([x, y, z], axis=1)
    if noise_sigma > 0:
        for d in range(3):
            xyz[:, d] += noise_sigma * gaussian(seed, idx, 8 + 2 * d)
    return _finish(xyz)


# ---- the five BASELINE.json configs ----------------------------------------------------
CONFIGS = {
    "C1": dict(desc="100k vs 100k uniform box, 0.1 m voxel, AC+CD only", n_est=100_000, n_gt=100_000, kind="box",
               tau=[0.2, 0.1, 0.08, 0.05, 0.01], nn_radius=0.1, vmd_voxel_size=0.1, mme=False, gt_mme=False,
               awd=False),
    "C2": dict(desc="1M vs 1M uniform box, full AC/COM/CD/MME/AWD/SCS", n_est=1_000_000, n_gt=1_000_000, kind="box",
               tau=[0.2, 0.1, 0.08, 0.05, 0.01], nn_radius=0.1, vmd_voxel_size=0.25, mme=True, gt_mme=True,
               awd=True),
    "C3": dict(desc="10M est vs 10M GT, 0.2 m voxel, trunc_dist 0.5 m", n_est=10_000_000, n_gt=10_000_000, kind="box",
               tau=[0.5, 0.3, 0.2, 0.1, 0.05], nn_radius=0.1, vmd_voxel_size=0.2, mme=True, gt_mme=False, awd=True),
    "C4": dict(desc="50M est vs 20M GT outdoor surfaces, MME radius 1.0 m", n_est=50_000_000, n_gt=20_000_000,
               kind="outdoor", tau=[0.5, 0.3, 0.2, 0.1, 0.05], nn_radius=1.0, vmd_voxel_size=3.0, mme=True,
               gt_mme=False, awd=True),
    "C5": dict(desc="200M vs 200M dense indoor", n_est=200_000_000, n_gt=200_000_000, kind="indoor",
               tau=[0.2, 0.1, 0.08, 0.05, 0.01], nn_radius=0.1, vmd_voxel_size=2.0, mme=True, gt_mme=False,
               awd=True),
    # not a BASELINE config: site-scale extent for the sparse cell table
    "S1": dict(desc="100M vs 100M site-scale terrain, 1 km x 1 km x 50 m (sparse cell table)", n_est=100_000_000,
               n_gt=100_000_000, kind="site", tau=[0.5, 0.3, 0.2, 0.1, 0.05], nn_radius=0.5, vmd_voxel_size=3.0, mme=True,
               gt_mme=False, awd=True),
}
EST_NOISE_SIGMA = 0.01
GT_SURFACE_NOISE_SIGMA = 0.002


def make_pair(name, scale=1.0):
    """(est, gt, cfg) for a BASELINE config; scale < 1 shrinks both clouds at constant density (parity-test sizes)."""
    cfg = dict(CONFIGS[name])
    n_est = max(1, int(round(cfg["n_est"] * scale)))
    n_gt = max(1, int(round(cfg["n_gt"] * scale)))
    cfg["n_est"], cfg["n_gt"] = n_est, n_gt
    if cfg["kind"] == "box":
        side = box_side_for_density(n_gt)
        cfg["side"] = side
        gt = uniform_box(n_gt, side, GT_SEED)
        est = uniform_box(n_est, side, EST_SEED, noise_sigma=EST_NOISE_SIGMA)
    elif cfg["kind"] == "outdoor":
        gt = outdoor_scene(n_gt, GT_SEED, GT_SURFACE_NOISE_SIGMA)
        est = outdoor_scene(n_est, EST_SEED, EST_NOISE_SIGMA)
    elif cfg["kind"] == "site":
        gt = site_scene(n_gt, GT_SEED, GT_SURFACE_NOISE_SIGMA)
        est = site_scene(n_est, EST_SEED, EST_NOISE_SIGMA)
    else:
        gt = indoor_scene(n_gt, GT_SEED, GT_SURFACE_NOISE_SIGMA)
        est = indoor_scene(n_est, EST_SEED, EST_NOISE_SIGMA)
    return est, gt, cfg
