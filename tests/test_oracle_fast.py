"""CPU checks of the oracle's fast_mode / raytrace restatement (oracle/vdbm_oracle.cpp, "fast_mode / raytrace support":
tools::VolumeRayIntersector restated, PARITY UNPINNED at the OpenVDB boundary). These pin the restatement against what the
reference's call sites (VDBMapping.hpp:577-602, 675-721) must produce in situations that can be worked out by hand."""
import numpy as np

from helpers import CFG_GTEST, popcount64
from oracle.oracle import OracleOccupancyVDBMapping


def _map(res=0.1, max_range=20.0, cfg=CFG_GTEST):
    o = OracleOccupancyVDBMapping(res)
    assert o.setConfig(max_range, *cfg) == 0
    o.addInputSource("s", max_range, 0)
    return o


def _voxels(ls, value=False):
    """set of (x, y, z) of the active (or value == true) voxels of a bool leaf set"""
    out = set()
    masks = ls.valmask if value else ls.active
    for org, m in zip(ls.origins, masks):
        for w in range(8):
            word = int(m[w])
            while word:
                b = (word & -word).bit_length() - 1
                word &= word - 1
                out.add((int(org[0]) + w, int(org[1]) + (b >> 3), int(org[2]) + (b & 7)))
    return out


def _wall(x, half=6, step=0.1):
    return np.array([[x, i * step, j * step] for i in range(-half, half + 1) for j in range(-half, half + 1)], dtype=np.float32)


def test_fast_mode_on_an_empty_map_sets_end_points_only():
    o = _map()
    o.setFastMode(True)
    pts = _wall(2.0)
    o.accumulateUpdate(pts, np.zeros(3), "s")          # V:522: nothing is cast into an empty map, V:533-536 still applies
    u = o.exportUpdateGrid("s")
    assert _voxels(u) == _voxels(u, value=True)
    assert len(_voxels(u)) == len({(20, i, j) for i in range(-6, 7) for j in range(-6, 7)})
    assert o.stats()["visits"] == 0


def test_fast_mode_touches_only_occupied_voxels_on_the_ray():
    """A wall in the voxel plane x = 20 (one hit activates with this config), then fast-mode rays THROUGH it to x = 60:
    every ray marks exactly its wall voxel (a miss) and its new end point (a hit); free space stays untouched."""
    o = _map()
    origin = np.zeros(3)
    o.insertPointCloud(_wall(2.0), origin, "s")
    wall = {(20, i, j) for i in range(-6, 7) for j in range(-6, 7)}
    o.setFastMode(True)
    visits_before = o.stats()["visits"]
    through = np.array([[6.0, 0.0, 0.0], [6.0, 0.0, 0.1], [0.0, 6.0, 0.0]], dtype=np.float32)   # two through the wall, one past it
    o.accumulateUpdate(through, origin, "s")
    u = o.exportUpdateGrid("s")
    act, hit = _voxels(u), _voxels(u, value=True)
    assert hit == {(60, 0, 0), (60, 0, 1), (0, 60, 0)}
    # the ray to (60, 0, 1) leaves z = 0 for z = 1 half way (x = 30), i.e. behind the wall: it crosses the wall at (20, 0, 0) too
    assert act - hit == {(20, 0, 0)}
    assert (act - hit) <= wall
    assert o.stats()["visits"] - visits_before == 2    # two setActiveState calls of V:598, one per crossing ray
    o.integrateUpdate()
    assert o.probe([20, 0, 0])[1] is True or o.probe([20, 0, 0])[1] == 1   # 2.197 - 2.197 = 0 is not below thres_min: still active


def test_fast_mode_marks_are_a_subset_of_the_normal_marks_and_of_the_occupied_set():
    rng = np.random.default_rng(4)
    a, b = _map(), _map()
    d = rng.normal(size=(1500, 3)); d /= np.linalg.norm(d, axis=1, keepdims=True)
    room = (d * 3.0).astype(np.float32)
    origin = np.array([0.02, 0.03, 0.01])
    for m in (a, b):
        m.insertPointCloud(room, origin, "s")
    occupied = set()
    mp = a.exportMap()
    for org, m in zip(mp.origins, mp.active):
        for w in range(8):
            word = int(m[w])
            while word:
                bit = (word & -word).bit_length() - 1
                word &= word - 1
                occupied.add((int(org[0]) + w, int(org[1]) + (bit >> 3), int(org[2]) + (bit & 7)))
    assert len(occupied) > 500
    b.setFastMode(True)
    longer = (d * 5.0).astype(np.float32)
    a.accumulateUpdate(longer, origin, "s"); b.accumulateUpdate(longer, origin, "s")
    ua, ub = a.exportUpdateGrid("s"), b.exportUpdateGrid("s")
    na, nb = _voxels(ua), _voxels(ub)
    hits = _voxels(ub, value=True)
    assert hits == _voxels(ua, value=True)
    assert (nb - hits) <= occupied
    # the span DDA restarts from ray(t0) with its own roundings, so it may leave the origin DDA's path by a voxel at a tie;
    # nearly all marks must coincide with the normal path
    assert len((nb - hits) - na) <= max(2, len(nb) // 100)
    assert len(nb - hits) > 300                         # the rays really crossed the room's walls


def test_raytrace_known_answers():
    o = _map()
    o.insertPointCloud(_wall(2.0), np.zeros(3), "s")
    ok, e = o.raytrace([[0.05, 0.05, 0.05]], [[1.0, 0.0, 0.0]], 5.0)
    assert ok[0] and np.allclose(e[0], [2.0, 0.0, 0.0])          # indexToWorld(voxel 20, 0, 0): the voxel's lower corner
    ok, e = o.raytrace([[0.05, 0.05, 0.05]], [[3.0, 0.0, 0.0]], 5.0)   # the direction is normalised first (V:691)
    assert ok[0] and np.allclose(e[0], [2.0, 0.0, 0.0])
    ok, e = o.raytrace([[50.0, 50.0, 50.0]], [[0.0, 0.0, 1.0]], 5.0)   # nowhere near the map: origin + direction * length (V:719)
    assert not ok[0] and np.allclose(e[0], [50.0, 50.0, 55.0])
    ok, e = o.raytrace([[0.05, 0.05, 0.05]], [[1.0, 0.0, 0.0]], 1.0)   # stops short of the wall inside mapped free space:
    assert ok[0] and e[0][0] < 2.0                                     # "success" (a node span was found), end = where the DDA ran out
    empty = _map()
    ok, e = empty.raytrace([[0.0, 0.0, 0.0]], [[1.0, 0.0, 0.0]], 2.0)
    assert not ok[0] and np.allclose(e[0], [2.0, 0.0, 0.0])


def test_infinite_points_are_dropped_like_nan():
    """+-inf coordinates are undefined behaviour in the reference; the documented rule of this build is
    'dropped and counted like NaN' - on the device (tests/test_gpu_parity.py::test_edge_inputs) and in the checker."""
    o = _map()
    pts = np.array([[np.inf, 0, 0], [0, -np.inf, 0], [1.0, 0, 0], [np.nan, 0, 0], [0, 0, np.inf]], dtype=np.float32)
    o.accumulateUpdate(pts, np.zeros(3), "s")
    st = o.stats()
    assert st["rays"] == 5 and st["nan_skipped"] == 4
    u = o.exportUpdateGrid("s")
    assert _voxels(u, value=True) == {(10, 0, 0)}
    o.accumulateUpdate(pts[2:3], np.array([np.inf, 0.0, 0.0]), "s")     # an infinite ORIGIN drops every point
    assert o.stats()["nan_skipped"] == 5
