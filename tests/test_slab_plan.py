"""The planner of the slab layout (me_plan_slab_cut: the pure host arithmetic every rank runs on the replicated cloud's
plane histogram) — CPU tests, no device needed."""
import ctypes as C

import numpy as np
import pytest

from cloud_map_evaluation_b200 import _lib


def _cut(py, pz, m, world, halo=4):
    L = _lib.load()
    py = np.ascontiguousarray(py, dtype=np.uint64)
    pz = np.ascontiguousarray(pz, dtype=np.uint64)
    axis, share = C.c_int32(-1), C.c_double(0)
    b = (C.c_int32 * (world + 1))()
    u64 = C.POINTER(C.c_uint64)
    rc = L.me_plan_slab_cut(py.ctypes.data_as(u64), len(py), pz.ctypes.data_as(u64), len(pz), m, world, halo, C.byref(axis), b,
                            C.byref(share))
    assert rc == 0
    return axis.value, list(b), share.value


def _numpy_rule(planes, m, world, halo):
    """restatement of the cut of one axis: (bounds, busiest share) or None"""
    nl = len(planes) // m
    if nl < 2 * world:
        return None
    cum = np.concatenate([[0], np.cumsum(planes[:nl * m].astype(np.int64))])
    total = int(cum[-1])
    if total == 0:
        return None
    b = [0] * (world + 1)
    b[world] = nl
    for r in range(1, world):
        want = total // world * r
        lo, hi = b[r - 1] + 1, nl - (world - r)
        l = lo
        while l < hi and cum[l * m] < want:
            l += 1
        if l > lo and want - cum[(l - 1) * m] < cum[l * m] - want:
            l -= 1
        b[r] = l
    worst = max(int(cum[min(nl * m, b[r + 1] * m + halo)] - cum[max(0, b[r] * m - halo)]) for r in range(world))
    return b, worst / total


def test_library_loads_without_a_device_and_plans():
    axis, b, share = _cut(np.full(184, 1000), np.full(184, 1000), 4, 8)
    assert axis in (1, 2) and b[0] == 0 and b[8] == 46
    assert all(b[r + 1] > b[r] for r in range(8))
    assert max(b[r + 1] - b[r] for r in range(8)) - min(b[r + 1] - b[r] for r in range(8)) <= 1
    assert 0.125 < share < 0.2          # 6 layers of 46 plus 8 halo planes of 184


@pytest.mark.parametrize("world", [2, 3, 4, 8])
@pytest.mark.parametrize("seed", [0, 1, 2])
def test_cut_equals_the_rule_and_is_a_partition(world, seed):
    rs = np.random.RandomState(seed)
    m = int(rs.choice([1, 4, 6]))
    ny, nz = int(rs.randint(2 * world, 90)) * m, int(rs.randint(1, 40)) * m
    py = rs.randint(0, 5000, ny).astype(np.uint64)
    pz = np.zeros(nz, np.uint64)
    pz[rs.randint(0, nz, 3 * nz)] += np.uint64(py.sum() // (3 * nz) + 1)       # lumpy along z
    pz[0] += np.uint64(int(py.sum()) - int(pz.sum())) if int(py.sum()) > int(pz.sum()) else np.uint64(0)
    axis, b, share = _cut(py, pz, m, world)
    cands = {1: _numpy_rule(py, m, world, 4), 2: _numpy_rule(pz, m, world, 4)}
    best = min((c[1], a) for a, c in cands.items() if c is not None)
    if best[0] > 0.75:
        assert axis == 0
        return
    # y is examined first and wins ties
    exp_axis = 1 if cands[1] is not None and cands[1][1] <= best[0] else 2
    assert axis == exp_axis
    assert b == cands[axis][0] and share == pytest.approx(cands[axis][1], rel=1e-15)
    assert b[0] == 0 and all(b[r + 1] > b[r] for r in range(world))      # every rank owns at least one layer, no gaps


def test_flat_scene_is_cut_along_y():
    """a terrestrial scan: 67 voxel layers along y, 5 along z with the ground in two of them"""
    m = 6
    py = np.full(67 * m, 500, np.uint64)
    pz = np.zeros(5 * m, np.uint64)
    pz[m - 2:m + 2] = 67 * m * 500 // 4
    axis8, b8, share8 = _cut(py, pz, m, 8)
    assert axis8 == 1 and share8 < 0.2
    axis2, _, share2 = _cut(py, pz, m, 2)          # two ranks: y (0.51) beats z (the ground's planes sit inside the halo)
    assert axis2 == 1 and share2 < 0.55


def test_scenes_that_cannot_be_cut():
    assert _cut(np.full(12, 10), np.full(12, 10), 4, 2)[0] == 0           # 3 voxel layers per axis < 2 x world
    one = np.zeros(64, np.uint64)
    one[10] = 1000                                                         # everything in one plane: the busiest rank holds it all
    assert _cut(one, one, 1, 4)[0] == 0
    assert _cut(np.zeros(64, np.uint64), np.zeros(64, np.uint64), 1, 4)[0] == 0
    L = _lib.load()
    assert L.me_plan_slab_cut(None, 0, None, 0, 1, 2, 4, None, None, None) != 0
