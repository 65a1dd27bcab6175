"""Host-side rendering helpers of the re-hosted driver (cloud_map_evaluation_b200/host/render.hpp), compiled into a tiny
harness: jet colour map, entropy / distance colouring and the binary xyz+rgb PCD writer against a Python restatement."""
import os
import subprocess

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))

HARNESS = r'''
#include "render.hpp"
#include <cstdio>
#include <cstdlib>
int main(int argc, char **argv) {
  // stdin: n, then n rows "x y z entropy sqdist"; argv[1], argv[2]: output PCDs; argv[3]: dis
  int n = 0;
  if (scanf("%d", &n) != 1) return 1;
  std::vector<double> xyz(3 * (size_t)n), ent(n), d2(n);
  for (int i = 0; i < n; ++i)
    if (scanf("%lf %lf %lf %lf %lf", &xyz[3 * i], &xyz[3 * i + 1], &xyz[3 * i + 2], &ent[i], &d2[i]) != 5) return 2;
  if (!render::write_pcd(argv[1], render::color_by_entropy(xyz, ent))) return 3;
  if (!render::write_pcd(argv[2], render::color_by_distance(xyz, d2, atof(argv[3])))) return 4;
  return 0;
}
'''


def _jet_u8(v):
    def interp(x, y0, x0, y1, x1):
        return np.where(x < x0, y0, np.where(x > x1, y1, (x - x0) * (y1 - y0) / (x1 - x0) + y0))

    def base(x):
        return np.where(x <= -0.75, 0.0, np.where(x <= -0.25, interp(x, 0.0, -0.75, 1.0, -0.25),
                        np.where(x <= 0.25, 1.0, np.where(x <= 0.75, interp(x, 1.0, 0.25, 0.0, 0.75), 0.0))))
    rgb = np.stack([base(v * 2 - 1.5), base(v * 2 - 1.0), base(v * 2 - 0.5)], axis=1)
    return np.round(np.clip(rgb, 0, 1) * 255).astype(np.uint8)


def _read(path):
    raw = open(path, "rb").read()
    head, data = raw.split(b"DATA binary\n", 1)
    text = head.decode()
    n = int([l for l in text.splitlines() if l.startswith("POINTS")][0].split()[1])
    assert "FIELDS x y z rgb" in text and "SIZE 4 4 4 4" in text and "TYPE F F F F" in text and f"WIDTH {n}" in text
    assert len(data) == 16 * n
    rec = np.frombuffer(data, dtype=np.dtype([("xyz", "<f4", 3), ("rgb", "<u4")]), count=n)
    rgb = np.stack([(rec["rgb"] >> 16) & 255, (rec["rgb"] >> 8) & 255, rec["rgb"] & 255], axis=1).astype(np.uint8)
    return rec["xyz"].copy(), rgb


def test_render_helpers(tmp_path):
    src = tmp_path / "harness.cpp"
    src.write_text(HARNESS)
    exe = tmp_path / "harness"
    subprocess.check_call(["/usr/bin/g++", "-O1", "-std=c++17", "-I", os.path.join(ROOT, "cloud_map_evaluation_b200", "host"),
                           "-o", str(exe), str(src)])
    rng = np.random.RandomState(2)
    n = 500
    xyz = rng.normal(0, 30, (n, 3))
    ent = -rng.uniform(5.0, 9.0, n)
    ent[rng.rand(n) < 0.2] = 0.0                               # invalid points are left out of the entropy map
    d2 = rng.uniform(0, 0.08, n)
    d2[:5] = np.nan                                            # no neighbour: painted at the clipping distance
    dis = 0.05
    text = f"{n}\n" + "\n".join(" ".join(repr(float(v)) for v in (*xyz[i], ent[i], d2[i])) for i in range(n))
    a, b = tmp_path / "ent.pcd", tmp_path / "dist.pcd"
    subprocess.run([str(exe), str(a), str(b), repr(dis)], input=text.replace("nan", "nan"), text=True, check=True)
    pts, rgb = _read(str(a))
    keep = ent != 0
    np.testing.assert_array_equal(pts, xyz[keep].astype(np.float32))
    nz = ent[keep]
    max_abs, min_abs = abs(nz.min()), abs(nz.max())
    norm = (np.abs(nz) - min_abs) / (max_abs - min_abs)
    norm = (np.log(norm + 0.1) - np.log(0.1)) / (np.log(1.1) - np.log(0.1))
    np.testing.assert_array_equal(rgb, _jet_u8(norm))
    pts, rgb = _read(str(b))
    np.testing.assert_array_equal(pts, xyz.astype(np.float32))
    dd = np.where(np.isnan(d2), dis, np.minimum(d2, dis))
    np.testing.assert_array_equal(rgb, _jet_u8(dd / dis))
    # the jet map itself: blue at 0, red at 1, green-ish in the middle
    assert _jet_u8(np.array([0.0]))[0].tolist() == [0, 0, 128] and _jet_u8(np.array([1.0]))[0].tolist() == [128, 0, 0]
