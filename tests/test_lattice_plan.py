"""One planning step of the cell lattice (me_plan_lattice: pure host arithmetic, no device) — the invariants the kernels
rely on, and the table kind (dense / sparse) the BASELINE configs get."""
import ctypes as C
import math

import numpy as np
import pytest

from cloud_map_evaluation_b200 import _abi as A
from cloud_map_evaluation_b200 import _lib, synth


def _plan(bmin, bmax, n, v=0.0, h=0.0, other=None, other_n=0, budget=0, sparse=True):
    L = _lib.load()
    dp = C.POINTER(C.c_double)
    a = (C.c_double * 3)(*bmin)
    b = (C.c_double * 3)(*bmax)
    oa = (C.c_double * 3)(*other[0]) if other else None
    ob = (C.c_double * 3)(*other[1]) if other else None
    out = A.me_lattice_plan()
    rc = L.me_plan_lattice(a, b, n, C.cast(oa, dp) if other else None, C.cast(ob, dp) if other else None, other_n, v, h, budget,
                           1 if sparse else 0, C.byref(out))
    return rc, out


def _check_invariants(p, bmin, bmax, v):
    assert p.h == pytest.approx(p.v / p.m, rel=1e-15)
    for a in range(3):
        assert p.dims[a] == p.nvox[a] * p.m
        assert p.nvox[a] == math.floor(bmax[a] / p.v) - math.floor(bmin[a] / p.v) + 1      # voxel_calculator.cpp:241-245
    if v > 0:
        assert p.v == v
    if not p.sparse:
        assert p.ncells == p.dims[0] * p.dims[1] * p.dims[2]


def test_c3_box_gets_the_dense_table_of_design_2():
    side = synth.box_side_for_density(10_000_000)
    rc, p = _plan([0, 0, 0], [side] * 3, 10_000_000, v=0.2)
    assert rc == 0 and not p.sparse and p.m == 4 and p.h == pytest.approx(0.05)
    assert all(180 <= d <= 192 for d in p.dims)
    _check_invariants(p, [0, 0, 0], [side] * 3, 0.2)


def test_c5_rooms_dense_at_the_volume_edge_sparse_at_the_refined_edge():
    """200 M points on the walls of 100 rooms (80 m x 80 m x 3 m): the first step (2 points per cell of the box VOLUME) fits
    a dense table; the occupancy refinement (surfaces: ~43 points per occupied cell -> h = 1.2 cm) asks for 10^10 cells, over
    the 2^31 budget -> sparse table.  This is why C5 stays on the replicated layout."""
    bmin, bmax = [0, 0, 0], [80, 80, 3]
    rc, p1 = _plan(bmin, bmax, 200_000_000, v=2.0)
    assert rc == 0 and not p1.sparse and 0.05 < p1.h < 0.065 and p1.ncells < 2 ** 31
    rc, p2 = _plan(bmin, bmax, 200_000_000, v=2.0, h=p1.h * math.sqrt(2.0 / 43.0))
    assert rc == 0 and p2.sparse and p2.h < 0.0135
    assert float(p2.dims[0]) * p2.dims[1] * p2.dims[2] > 2 ** 31
    _check_invariants(p1, bmin, bmax, 2.0)
    _check_invariants(p2, bmin, bmax, 2.0)
    # round-1 behaviour (no sparse table): the cells are coarsened until the dense table fits
    rc, p3 = _plan(bmin, bmax, 200_000_000, v=2.0, h=p2.h, sparse=False)
    assert rc == 0 and not p3.sparse and p3.h > 1.25 * p2.h and p3.ncells <= 2 ** 31


def test_site_scale_scene_is_sparse_and_small_scenes_dense():
    rc, p = _plan([0, 0, -20], [1000, 1000, 30], 100_000_000, v=3.0, h=0.1)
    assert rc == 0 and p.sparse and p.h == pytest.approx(0.1)
    rc, p = _plan([0, 0, 0], [2.5, 2.5, 2.5], 200_000, v=0.2)
    assert rc == 0 and not p.sparse


def test_budget_cells_per_voxel_cap_and_shared_spec():
    rc, p = _plan([0, 0, 0], [10, 10, 10], 1_000_000, v=1.0, h=0.01, budget=1_000_000, sparse=False)
    assert rc == 0 and p.ncells <= 1_000_000 and p.h >= 0.01
    # duplicates / collinear points would ask for ever smaller cells: m is capped at 512 cells per voxel edge
    rc, p = _plan([0, 0, 0], [9, 9, 9], 1000, v=3.0, h=1e-5)
    assert rc == 0 and p.m <= 512
    # the second cloud must fit under the same (v, m): a far-away ground truth forces coarser cells on a dense table
    rc, alone = _plan([0, 0, 0], [10, 10, 10], 1_000_000, v=1.0, h=0.02, sparse=False)
    rc2, shared = _plan([0, 0, 0], [10, 10, 10], 1_000_000, v=1.0, h=0.02, other=([0, 0, 0], [400, 400, 40]), other_n=1_000_000,
                        sparse=False)
    assert rc == 0 and rc2 == 0 and shared.m < alone.m and shared.v == alone.v == 1.0
    # free voxel size (no voxel stage): cells of the wanted edge
    rc, p = _plan([-5, -5, -5], [5, 5, 5], 1_000_000, h=0.05)
    assert rc == 0 and p.m == 1 and p.h == pytest.approx(0.05)
    _check_invariants(p, [-5, -5, -5], [5, 5, 5], 0.0)


def test_negative_coordinates_and_errors():
    bmin, bmax = [-7.3, -0.1, -120.5], [-1.2, 33.0, -100.0]
    rc, p = _plan(bmin, bmax, 500_000, v=0.25)
    assert rc == 0
    _check_invariants(p, bmin, bmax, 0.25)
    assert _plan([0, 0, 0], [1, 1, 1], 0)[0] != 0
    assert _plan([1, 0, 0], [0, 1, 1], 10)[0] != 0
    # |voxel index| must fit an int, as the reference casts (voxel_calculator.cpp:242)
    assert _plan([0, 0, 0], [1e12, 1, 1], 10, v=1e-3)[0] != 0
