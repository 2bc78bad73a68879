"""One process, several GPUs: vdbm_group_* (include/vdbm_b200.h) and the device-side ray split (vdbm_ray_sector_set).
The union of the shards must be the reference's ONE map: compared bit for bit with the oracle and, by checksum, with the same
scans on a single handle. Shards may share a device, so the whole machinery (worker threads, direct peer wiring, sector
plan, fused exchange, device-side waits) also runs on a one-GPU box; with >= 2 GPUs the same tests run across NVLink."""
import ctypes as C

import numpy as np
import pytest

from helpers import CFG_ROS, assert_leafsets_equal

pytestmark = pytest.mark.gpu


def _n_gpus():
    import torch
    return torch.cuda.device_count()


def _device_sets():
    sets = [[0], [0, 0], [0, 0, 0]]
    n = _n_gpus()
    if n >= 2:
        sets.append([0, 1])
    if n >= 4:
        sets.append([0, 1, 2, 3])
    if n >= 8:
        sets.append(list(range(8)))
    return sets


def _oracle(res, max_range, cfg):
    from oracle.oracle import OracleOccupancyVDBMapping
    o = OracleOccupancyVDBMapping(res)
    assert o.setConfig(max_range, *cfg) == 0
    o.addInputSource("s", max_range, 0)
    return o


@pytest.mark.parametrize("devices", _device_sets(), ids=lambda d: "dev" + "".join(map(str, d)))
def test_group_union_is_the_one_map(devices):
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping, OccupancyVDBMappingGroup
    res, rng_max = 0.1, 6.0
    grp = OccupancyVDBMappingGroup(res, devices)
    one = OccupancyVDBMapping(res)
    o = _oracle(res, rng_max, CFG_ROS)
    for m in (grp, one):
        assert m.setConfig(rng_max, *CFG_ROS) == 0
        m.addInputSource("s", rng_max, 0)
    for k in range(4):
        pts, origin = scans.small_scan(300 + k, n=20000, scale=4.0)
        origin = origin + np.array([0.21 * k, -0.17 * k, 0.02 * k])
        pts[5] = np.nan
        for m in (grp, one, o):
            m.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(grp.exportMap(), o.exportMap(), f"union of {len(devices)} shards vs oracle")
    assert grp.checksum() == one.mapChecksum()
    sg, so = grp.stats(), o.stats()
    for key in ("rays", "nan_skipped", "clipped", "visits", "voxel_updates", "state_changes"):
        assert sg[key] == so[key], key
    if len(devices) > 1:
        counts = [grp.shard(i).mapLeafCount() for i in range(len(devices))]
        assert all(c > 0 for c in counts) and sum(counts) == len(o.exportMap())
        # ownership follows the plan: every leaf of shard i is owned by rank i
        c, rb, ob = grp.plan()
        assert np.all(np.diff(rb) > 0) and np.all(np.diff(ob) > 0)
    # resetMap: empty shards, a new plan from the next first scan, same result again
    grp.resetMap(); o.resetMap()
    pts, origin = scans.small_scan(77, n=15000, scale=3.0)
    grp.insertPointCloud(pts, origin + 1.5, "s"); o.insertPointCloud(pts, origin + 1.5, "s")
    assert_leafsets_equal(grp.exportMap(), o.exportMap(), "after resetMap")
    grp.close()


def test_group_full_size_scan_matches_single_handle():
    """BASELINE configs[1] (262,144-point OS1-128 scans) on as many GPUs as the box has (2 shards on one GPU otherwise):
    checksum of the sharded map == the single-handle map, scan after scan."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping, OccupancyVDBMappingGroup
    c = scans.CONFIGS[2]
    n = _n_gpus()
    devices = list(range(n)) if n >= 2 else [0, 0]
    grp = OccupancyVDBMappingGroup(c.resolution, devices)
    one = OccupancyVDBMapping(c.resolution)
    for m in (grp, one):
        assert m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max) == 0
        m.addInputSource("s", c.max_range, 0)
    for k in range(3):
        pts, origin = scans.make_scan(2, k)
        grp.insertPointCloud(pts, origin, "s"); one.insertPointCloud(pts, origin, "s")
        assert grp.checksum() == one.mapChecksum(), f"scan {k}"
    st = grp.stats()
    assert st["map_leaves"] == one.mapLeafCount() and st["rays"] == 3 * 262144
    grp.close()


def test_ray_sector_filter_partitions_a_scan_exactly():
    """vdbm_ray_sector_set: three handles, the same cloud, sectors 0 / 1 / 2 of the same bounds: every ray is cast by exactly
    one of them (counters add up) and the OR of the three update grids is the unsplit update grid."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    import vdb_mapping_b200._lib as L
    res, rng_max = 0.1, 6.0
    pts, origin = scans.small_scan(5, n=30000, scale=4.0)
    pts[7] = np.nan
    full = OccupancyVDBMapping(res)
    parts = [OccupancyVDBMapping(res) for _ in range(3)]
    bounds = np.array([0.3, 1.7, 2.9])
    for i, m in enumerate([full] + parts):
        assert m.setConfig(rng_max, *CFG_ROS) == 0
        m.addInputSource("s", rng_max, 0)
        if i:
            assert m._L.vdbm_ray_sector_set(m._h, 3, i - 1, bounds.ctypes.data_as(C.POINTER(C.c_double))) == L.VDBM_OK
        m.accumulateUpdate(pts, origin, "s")
    uf = full.exportUpdateGrid("s")
    acc = {}
    for m in parts:
        u = m.exportUpdateGrid("s")
        assert 0 < len(u) < len(uf)
        for o_, a, v in zip(map(tuple, u.origins), u.active, u.valmask):
            if o_ in acc:
                acc[o_] = (acc[o_][0] | a, acc[o_][1] | v)
            else:
                acc[o_] = (a.copy(), v.copy())
    assert len(acc) == len(uf)
    for o_, a, v in zip(map(tuple, uf.origins), uf.active, uf.valmask):
        assert np.array_equal(acc[o_][0], a) and np.array_equal(acc[o_][1], v)
    sf = full.stats()
    for key in ("nan_skipped", "clipped", "visits"):
        assert sum(m.stats()[key] for m in parts) == sf[key], key
    # switching the filter off again
    m = parts[0]
    assert m._L.vdbm_ray_sector_set(m._h, 0, 0, None) == L.VDBM_OK
    m.integrateUpdate(); m.accumulateUpdate(pts, origin, "s")
    assert_leafsets_equal(m.exportUpdateGrid("s"), uf, "filter off")
