"""The torch restatement of the synthetic generators (bench input for C4/C5 on the device) against synth.py."""
import numpy as np
import pytest

from cloud_map_evaluation_b200 import synth

torch = pytest.importorskip("torch")
from cloud_map_evaluation_b200 import synth_torch  # noqa: E402


def _ulp32(a):
    return np.spacing(np.abs(a).astype(np.float32)).astype(np.float64)


def test_uniform24_bit_identical():
    idx = np.arange(5, 5 + 20000, dtype=np.uint64)
    for seed, lane in ((synth.GT_SEED, 0), (synth.EST_SEED, 3), (7, 9)):
        a = synth.uniform24(seed, idx, lane)
        b = synth_torch.uniform24(seed, torch.arange(5, 5 + 20000, dtype=torch.int64), lane).numpy()
        np.testing.assert_array_equal(a, b)


@pytest.mark.parametrize("name,scale", [("C2", 0.02), ("C4", 0.0005), ("C5", 0.0001), ("S1", 0.0002)])
def test_make_pair_matches_numpy(name, scale):
    est, gt, cfg = synth.make_pair(name, scale=scale)
    t_est, t_gt, t_cfg = synth_torch.make_pair(name, scale=scale, device="cpu")
    assert t_cfg["n_est"] == cfg["n_est"] and t_cfg["n_gt"] == cfg["n_gt"]
    for a, b in ((est, t_est.numpy()), (gt, t_gt.numpy())):
        assert a.shape == b.shape
        # libm differences (log / sin / cos) may move a coordinate by one fp32 ulp; everything else is identical
        d = np.abs(a - b)
        assert np.all(d <= _ulp32(a) * 1.0000001)
        assert np.mean(d == 0) > 0.999


def test_chunking_is_transparent():
    a = synth_torch.indoor_scene(5000, 3, 0.01, chunk=1 << 24)
    b = synth_torch.indoor_scene(5000, 3, 0.01, chunk=777)
    assert torch.equal(a, b)
