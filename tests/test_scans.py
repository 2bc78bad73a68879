"""The synthetic workloads of SURVEY.md 8(d) (vdb_mapping_b200/scans.py): shapes, determinism and the properties the
benchmarks rely on (NaN fraction, range clipping share, off-grid moving origin)."""
import numpy as np
import pytest

from vdb_mapping_b200 import scans


@pytest.mark.parametrize("cfg", [1, 2, 3, 4])
def test_scan_shape_determinism_and_nan_fraction(cfg):
    c = scans.CONFIGS[cfg]
    pts, origin = scans.make_scan(cfg, 2)
    pts2, origin2 = scans.make_scan(cfg, 2)
    assert pts.dtype == np.float32 and pts.shape[0] == c.n_points and pts.shape[1] in (3, 4)
    assert np.array_equal(np.nan_to_num(pts, nan=-7.0), np.nan_to_num(pts2, nan=-7.0)) and np.array_equal(origin, origin2)
    nan = np.isnan(pts[:, :3]).any(axis=1).mean()
    assert 0.005 < nan < 0.02, nan  # ~1 % NaN points exercise VDBMapping.hpp:505-510
    other, _ = scans.make_scan(cfg, 3)
    assert not np.array_equal(np.nan_to_num(pts, nan=-7.0), np.nan_to_num(other, nan=-7.0))


def test_cfg2_sequence_moves_off_grid_and_clips_some_rays():
    c = scans.CONFIGS[2]
    _, o0 = scans.make_scan(2, 0)
    pts, o5 = scans.make_scan(2, 5)
    assert np.allclose(o5 - o0, [0.137 * 5, 0.061 * 5, 0.013 * 5])
    assert np.all(np.abs(np.fmod(o5, c.resolution)) > 1e-9)  # off-grid: the fmod branch of worldToIndex is taken
    d = np.linalg.norm(pts[:, :3].astype(np.float64) - o5, axis=1)
    clipped = np.nanmean(d > c.max_range)
    assert 0.0 < clipped < 0.5  # some rays reach beyond max_range (free-space-only rays), most do not


def test_multi_sensor_variants_differ():
    a, _ = scans.make_scan(2, 0, sensor=0)
    b, _ = scans.make_scan(2, 0, sensor=1)
    assert a.shape == b.shape and not np.array_equal(np.nan_to_num(a), np.nan_to_num(b))
