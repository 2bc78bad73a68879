"""Pins the CPU oracle on the reference's own known-answer tests (/root/reference/tests/mapping.cpp),
transcribed in tests/golden/mapping_kats.json, and cross-checks it against the independent
pure-Python restatement in tests/pyref.py."""
import json
import math
import os

import numpy as np
import pytest

from oracle.oracle import OracleOccupancyVDBMapping
import pyref
from helpers import CFG_GTEST, CFG_ROS, assert_leafsets_equal, popcount64

HERE = os.path.dirname(os.path.abspath(__file__))
KATS = json.load(open(os.path.join(HERE, "golden", "mapping_kats.json")))


def _const(name, cfg):
    if name == "log_hit":
        return np.float32(math.log(cfg["prob_hit"]) - math.log(1 - cfg["prob_hit"]))
    if name == "log_miss":
        return np.float32(math.log(cfg["prob_miss"]) - math.log(1 - cfg["prob_miss"]))
    return np.float32(name)


def kat_points(case):
    if "points" in case:
        return np.array(case["points"], dtype=np.float32)
    res = case["resolution"]
    # the reference test computes k * resolution in double, PointXYZ narrows to float
    return np.array([[np.float32(k * res) for k in p] for p in case["points_in_resolutions"]], dtype=np.float32)


def run_kat(make_map, case, cfg):
    """Drives any object with the OccupancyVDBMapping-like surface through one reference test case."""
    m = make_map(case["resolution"])
    pts = kat_points(case)
    if "pre_config_insert" in case:  # tests/mapping.cpp:10-15: insert before setConfig/addInputSource is a no-op
        pre = case["pre_config_insert"]
        m.insertPointCloud(np.array(pre["points"], dtype=np.float32), pre["origin"], "test")
        for c, v, _ in pre["expect"]:
            assert m.probe(c)[0] == _const(v, cfg)
    assert m.setConfig(case["max_range"], cfg["prob_hit"], cfg["prob_miss"], cfg["prob_thres_min"], cfg["prob_thres_max"]) == 0
    m.addInputSource("test", case["max_range"], 0)
    assert m.insertPointCloud(pts, case["origin"], "test") is True
    for c, v, a in case["expect"]:
        val, on = m.probe(c)
        assert np.float32(val) == _const(v, cfg), (case["name"], c, val)
        if a is not None:
            assert on == a, (case["name"], c)
    if "after_reset_expect" in case:
        m.resetMap()
        for c, v, _ in case["after_reset_expect"]:
            assert m.probe(c)[0] == _const(v, cfg)
    return m


@pytest.mark.parametrize("case", KATS["cases"], ids=[c["name"] for c in KATS["cases"]])
def test_oracle_reference_kats(case):
    run_kat(OracleOccupancyVDBMapping, case, KATS["config"])


def test_logodds_constants():
    # constants of the gtest config and of the typical ROS config
    m = OracleOccupancyVDBMapping(0.1)
    assert m.setConfig(10, 0.9, 0.1, 0.49, 0.51) == 0
    lo = m.logodds()
    assert lo[0] == np.float32(float.fromhex("0x1.193ea8p+1"))
    assert lo[1] == -lo[0]
    assert lo[4] == np.float32(float.fromhex("0x1.261672p+2")) and lo[5] == -lo[4]
    assert m.setConfig(10, 0.7, 0.4, 0.12, 0.97) == 0
    lo = m.logodds()
    assert lo[0] == np.float32(float.fromhex("0x1.b1d106p-1"))
    assert lo[1] == np.float32(-float.fromhex("0x1.9f323ep-2"))


def test_setconfig_validation():
    m = OracleOccupancyVDBMapping(0.1)
    assert m.setConfig(-1, 0.9, 0.1, 0.49, 0.51) == 1      # VDBMapping.hpp:1458-1463
    assert m.setConfig(10, 0.9, 0.6, 0.49, 0.51) == 2      # OccupancyVDBMapping.hpp:65-70
    assert m.setConfig(10, 0.4, 0.1, 0.49, 0.51) == 2      # OccupancyVDBMapping.hpp:71-76


def test_world_to_index_half_voxel_rule():
    # VDBMapping.hpp:612-631: +res/2 only when fmod(c,res) != 0
    m = OracleOccupancyVDBMapping(0.1)
    for w in [0.0, 0.5, -0.5, 0.05, 0.04, 0.06, -0.04, -0.06, 0.7, 1e-9, -1e-9, 12.34, -7.77, 0.1, 0.2, 0.30000000000000004]:
        got = m.worldToIndex([w, w, w])
        exp = pyref.world_to_index([w, w, w], 0.1)
        assert tuple(int(x) for x in got) == exp, w


def _rand_cloud(rng, n, scale):
    p = rng.normal(size=(n, 3)) * scale
    return p.astype(np.float32)


@pytest.mark.parametrize("cfg", [(0.9, 0.1, 0.49, 0.51), (0.7, 0.4, 0.12, 0.97)])
@pytest.mark.parametrize("quirk", [True, False])
def test_oracle_vs_pyref_random(cfg, quirk):
    """Multi-scan random clouds: update grid, change grid and map must be identical to pyref."""
    rng = np.random.default_rng(7)
    res, max_range = 0.1, 2.0
    hit, miss, tmin, tmax = cfg
    o = OracleOccupancyVDBMapping(res)
    o.setProbeQuirk(quirk)
    assert o.setConfig(max_range, hit, miss, tmin, tmax) == 0
    o.addInputSource("s", max_range, 0)
    p = pyref.PyMap(res, max_range, hit, miss, tmin, tmax, quirk=quirk)
    for k in range(6):
        pts = _rand_cloud(rng, 60, 1.2)
        pts[::17] = np.nan
        origin = [0.137 * k, 0.061 * k, 0.013 * k]
        assert o.accumulateUpdate(pts, origin, "s") == 0
        upd_o = pyref.leafset_to_voxels(o.exportUpdateGrid("s"))
        upd_p, chg_p = p.insert(pts, origin)
        assert {k_: v[1] for k_, v in upd_o.items()} == upd_p
        assert all(v[0] for v in upd_o.values())
        o.integrateUpdate()
        chg_o = pyref.leafset_to_voxels(o.exportLastChange("s"))
        assert {k_: v[1] for k_, v in chg_o.items()} == chg_p
        map_o = pyref.leafset_to_voxels(o.exportMap())
        map_p = {v: (a, val) for v, (val, a) in p.vox.items() if a or val != 0}
        assert map_o == map_p
        assert len(o.exportUpdateGrid("s")) == 0
    assert o.mapLeafCount() == len(p.leaves)


def test_dda_invariants():
    """visits == 1 + |dx|+|dy|+|dz|, last voxel == end voxel, incl. exact-tie rays."""
    rng = np.random.default_rng(11)
    ends = [(5, 5, 5), (-5, -5, -5), (3, 9, 0), (1, 3, 9), (0, 0, 7), (-4, 4, 2), (100, 1, 0), (7, 7, 1)]
    ends += [tuple(int(x) for x in rng.integers(-40, 40, 3)) for _ in range(300)]
    for e in ends:
        o = tuple(int(x) for x in rng.integers(-5, 5, 3))
        ee = tuple(o[a] + e[a] for a in range(3))
        vox = pyref.dda_voxels(o, ee)
        if ee == o:
            assert vox == []
            continue
        assert len(vox) == 1 + sum(abs(x) for x in e)
        assert vox[0] == o and vox[-1] == ee
        assert len(set(vox)) == len(vox)


def test_sections_vs_bruteforce():
    rng = np.random.default_rng(3)
    o = OracleOccupancyVDBMapping(0.1)
    o.setConfig(3.0, 0.9, 0.1, 0.49, 0.51)
    o.addInputSource("s", 3.0, 0)
    for k in range(3):
        o.insertPointCloud(_rand_cloud(rng, 80, 1.5), [0.05 * k, 0, 0], "s")
    full_map = pyref.leafset_to_voxels(o.exportMap())
    mn, mx = (-9, -4, -3), (6, 11, 5)
    inside = lambda v: all(mn[a] <= v[a] <= mx[a] for a in range(3))
    sparse_u = pyref.leafset_to_voxels(o.getMapSectionUpdateGrid(mn, mx, False))
    assert sparse_u == {v: (True, True) for v, (a, _) in full_map.items() if a and inside(v)}
    sparse_f = pyref.leafset_to_voxels(o.getMapSectionGrid(mn, mx, False))
    assert sparse_f == {v: (True, np.float32(1.0)) for v, (a, _) in full_map.items() if a and inside(v)}
    full_f = pyref.leafset_to_voxels(o.getMapSectionGrid(mn, mx, True))
    assert full_f == {v: av for v, av in full_map.items() if inside(v)}
    full_u = pyref.leafset_to_voxels(o.getMapSectionUpdateGrid(mn, mx, True))
    assert full_u == {v: (a, bool(val != 0)) for v, (a, val) in full_map.items() if inside(v)}


def test_zero_and_negative_source_range_inserts_nothing():
    # VDBMapping.hpp:331: max_range <= 0 on the source => nothing is raycast, not even endpoints
    o = OracleOccupancyVDBMapping(0.1)
    o.setConfig(0.0, 0.9, 0.1, 0.49, 0.51)
    o.addInputSource("s", 0.0, 0)   # 0 -> falls back to m_max_range (0) -> nothing
    o.insertPointCloud(np.array([[0, 0, 0.5]], np.float32), [0, 0, 0], "s")
    assert o.mapLeafCount() == 0
    o.addInputSource("n", -1.0, 0)
    o.insertPointCloud(np.array([[0, 0, 0.5]], np.float32), [0, 0, 0], "n")
    assert o.mapLeafCount() == 0


def _map_dict(m):
    return pyref.leafset_to_voxels(m.exportMap())


def _two_maps(seed_a, seed_b):
    """sender map A and receiver map B with different content"""
    rng = np.random.default_rng(seed_a)
    maps = []
    for seed in (seed_a, seed_b):
        m = OracleOccupancyVDBMapping(0.1)
        m.setConfig(3.0, 0.9, 0.1, 0.49, 0.51)
        m.addInputSource("s", 3.0, 0)
        r = np.random.default_rng(seed)
        for k in range(3):
            m.insertPointCloud(_rand_cloud(r, 80, 1.5), [0.05 * k, 0.02 * seed, 0], "s")
        maps.append(m)
    return maps


def test_apply_section_update_grid_vs_bruteforce():
    """applyMapSectionUpdateGrid (VDBMapping.hpp:1058-1085): deactivate active voxels in the box, activate the section's."""
    a, b = _two_maps(1, 2)
    mn, mx = (-9, -4, -3), (6, 11, 5)
    inside = lambda v: all(mn[i] <= v[i] <= mx[i] for i in range(3))
    section = a.getMapSectionUpdateGrid(mn, mx, False)
    before = _map_dict(b)
    b.applyMapSectionUpdateGrid(mn, mx, section)
    want = {}
    for v, (act, val) in before.items():
        want[v] = (act and not inside(v), val)
    for v in pyref.leafset_to_voxels(section):
        want[v] = (True, want.get(v, (False, np.float32(0)))[1])
    want = {v: av for v, av in want.items() if av[0] or av[1] != 0}
    assert _map_dict(b) == want


@pytest.mark.parametrize("quirk", [False, True])
def test_apply_section_grid_vs_bruteforce(quirk):
    """applyMapSectionGrid (VDBMapping.hpp:1022-1047): section leaves overwrite map leaves completely; with the
    value-all tile visits every missing leaf slot of the section's internal nodes clears its origin voxel in the map."""
    a, b = _two_maps(3, 4)
    mn, mx = (-9, -4, -3), (6, 11, 5)
    section = a.getMapSectionGrid(mn, mx, True)
    before = _map_dict(b)
    leaves_before = {tuple(o) for o in b.exportMap().origins}
    b.applyMapSectionGrid(section, tile_quirk=quirk)
    want = dict(before)
    sec_leaves = {tuple(int(x) for x in o) for o in section.origins}
    if quirk:
        i1 = {tuple((c >> 7) << 7 for c in o) for o in sec_leaves}
        i2 = {tuple((c >> 12) << 12 for c in o) for o in sec_leaves}
        for blk in i1:  # every leaf slot of an I1 node of the section without a section leaf: its origin voxel
            for i in range(16):
                for j in range(16):
                    for k in range(16):
                        o = (blk[0] + 8 * i, blk[1] + 8 * j, blk[2] + 8 * k)
                        if o not in sec_leaves and o in leaves_before:
                            want.pop(o, None)
        for blk in i2:  # every I1 slot of an I2 node without child: origin voxel of that 128^3 block
            for i in range(32):
                for j in range(32):
                    for k in range(32):
                        o = (blk[0] + 128 * i, blk[1] + 128 * j, blk[2] + 128 * k)
                        if tuple((c >> 7) << 7 for c in o) not in i1 and o in leaves_before:
                            want.pop(o, None)
    for o in sec_leaves:  # section leaves replace map leaves voxel for voxel
        for v in [v for v in want if tuple((c >> 3) << 3 for c in v) == o]:
            del want[v]
    want.update(pyref.leafset_to_voxels(section))
    assert _map_dict(b) == want


# ---- remote-mapping deltas and direct edits -------------------------------------------
def _voxel_dict(ls):
    """{(x, y, z): (value, active)} of a float leaf set / {(x,y,z): hit} of a bool one (active voxels only for bool)."""
    out = {}
    for i in range(len(ls)):
        ox, oy, oz = (int(v) for v in ls.origins[i])
        for w in range(8):
            word = int(ls.active[i, w])
            vw = int(ls.valmask[i, w]) if ls.valmask is not None else 0
            for b in range(64):
                on = (word >> b) & 1
                c = (ox + w, oy + (b >> 3), oz + (b & 7))
                if ls.values is not None:
                    out[c] = (np.float32(ls.values[i, w * 64 + b]), bool(on))
                elif on:
                    out[c] = bool((vw >> b) & 1)
    return out


def _mk_oracle(res=0.1, rng=4.0, cfg=CFG_ROS):
    m = OracleOccupancyVDBMapping(res)
    assert m.setConfig(rng, *cfg) == 0
    m.addInputSource("s", rng)
    return m


def test_reduced_update_is_lossless_and_matches_ray_ends():
    """Level 2: the reduced grid holds exactly the rays' end voxels (value = not clipped), and re-raycasting it on a
    receiver reproduces the sender's update grid, change grid and map bit for bit."""
    from vdb_mapping_b200 import scans
    snd, rcv = _mk_oracle(), _mk_oracle()
    for k in range(4):
        pts, origin = scans.small_scan(40 + k, n=1500, scale=3.0)
        origin = origin + np.array([0.137 * k, 0.061 * k, 0.013 * k])
        snd.accumulateUpdate(pts, origin, "s")
        full = snd.exportUpdateGrid("s")
        red, o = snd.createUpdate("s", 2)
        assert np.array_equal(o, origin)
        # brute force: end voxel of every finite ray
        expect = {}
        oi = snd.worldToIndex(origin)
        for p in pts[:, :3].astype(np.float64):
            if not np.all(np.isfinite(p)):
                continue
            d = p - origin
            ln = np.sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2])
            clipped = ln > 4.0
            e = origin + (d / ln) * 4.0 if clipped else p
            c = tuple(int(v) for v in snd.worldToIndex(e))
            expect[c] = expect.get(c, False) or (not clipped)
        assert _voxel_dict(red) == expect
        snd.integrateUpdate()
        rcv.applyUpdate("s", 2, red, o)
        assert_leafsets_equal(rcv.exportLastChange("s"), snd.exportLastChange("s"), f"change scan {k}")
        assert popcount64(red.active) < popcount64(full.active) // 20
    assert_leafsets_equal(rcv.exportMap(), snd.exportMap(), "receiver map")


def test_raw_and_overwrite_updates():
    """Level 0 (raw update grid -> updateMap) reproduces the sender's map; level 1 (change grid -> overwrite) reproduces
    its ACTIVE SET with values pinned to the clamping bounds (setNodeToOccupied / setNodeToFree)."""
    from vdb_mapping_b200 import scans
    snd, r0, r1 = _mk_oracle(cfg=CFG_GTEST), _mk_oracle(cfg=CFG_GTEST), _mk_oracle(cfg=CFG_GTEST)
    lo = snd.logodds()
    for k in range(5):
        pts, origin = scans.small_scan(60 + k, n=1200, scale=2.5)
        snd.accumulateUpdate(pts, origin, "s")
        raw, _ = snd.createUpdate("s", 0)
        snd.integrateUpdate()
        chg, _ = snd.createUpdate("s", 1)
        r0.applyUpdate("s", 0, raw)
        r1.applyUpdate("s", 1, chg)
    assert_leafsets_equal(r0.exportMap(), snd.exportMap(), "level-0 receiver")
    a, b = _voxel_dict(snd.exportMap()), _voxel_dict(r1.exportMap())
    assert {c for c, (v, on) in a.items() if on} == {c for c, (v, on) in b.items() if on}
    for c, (v, on) in b.items():
        assert (v == lo[4] and on) or (v == lo[5] and not on) or (v == 0 and not on)


def test_points_set_brute_force():
    m = _mk_oracle(res=0.1)
    lo = m.logodds()
    rng = np.random.default_rng(5)
    add = (rng.uniform(-3, 3, size=(500, 3))).astype(np.float32)
    add[::50] = np.round(add[::50] * 10) / 10  # points on voxel boundaries: plain floor(p / res), no half-voxel shift
    m.addPointsToGrid(add)
    rem = np.concatenate([add[:100], rng.uniform(-3, 3, size=(100, 3)).astype(np.float32)])
    m.removePointsFromGrid(rem)
    expect = {}
    inv = 1.0 / 0.1
    for p in add.astype(np.float64):
        expect[tuple(int(np.floor(x * inv)) for x in p)] = (lo[4], True)
    for p in rem.astype(np.float64):
        expect[tuple(int(np.floor(x * inv)) for x in p)] = (lo[5], False)
    got = {c: v for c, v in _voxel_dict(m.exportMap()).items() if v != (np.float32(0), False)}
    assert got == expect


def test_artificial_areas():
    """Walls = castRayIntoGrid per height level into the artificial grid; every updateMap then forces them active;
    restoreMapIntegrity re-derives the flags from the values."""
    res = 0.1
    m = _mk_oracle(res=res, cfg=CFG_GTEST)
    poly = [np.array([[0.52, 0.5, 0.0], [2.31, 0.77, 0.0], [1.9, 2.2, 0.0]]), np.array([[-1.0, -1.0, 0.2], [-2.0, -1.5, 0.2]])]
    m.addArtificialAreas(poly, -0.35, 0.55)
    art = _voxel_dict(m.exportArtificialAreaGrid())
    expect = set()
    for pg in poly:
        for i in range(len(pg)):
            s, e = pyref.world_to_index(pg[i], res), pyref.world_to_index(pg[(i + 1) % len(pg)], res)
            for lv in range(int(-0.35 / res), int(0.55 / res)):
                expect |= set(pyref.dda_voxels((s[0], s[1], s[2] + lv), (e[0], e[1], e[2] + lv)))
    assert set(art) == expect and len(expect) > 100
    assert m.mapLeafCount() == 0  # nothing happens to the map until the next updateMap
    from vdb_mapping_b200 import scans
    pts, origin = scans.small_scan(3, n=800, scale=2.5)
    m.insertPointCloud(pts, origin, "s")
    mp = _voxel_dict(m.exportMap())
    assert all(mp[c][1] for c in expect)
    m.restoreMapIntegrity()
    assert len(m.exportArtificialAreaGrid()) == 0
    lo = m.logodds()
    mp2 = _voxel_dict(m.exportMap())
    for c in expect:
        assert mp2[c][1] == bool(mp2[c][0] > lo[3])


def test_flat_probe_counts_match_the_oracle():
    """The bench-only 'optimistic CPU' probe (oracle/flat_probe.cpp) walks the same voxels and ends with the same map
    occupancy as the oracle, single- and multi-threaded."""
    from oracle.oracle import FlatProbe
    from vdb_mapping_b200 import scans
    o = _mk_oracle(res=0.1, rng=4.0, cfg=CFG_ROS)
    f1, f3 = FlatProbe(0.1, o.logodds()), FlatProbe(0.1, o.logodds())
    for k in range(4):
        pts, origin = scans.small_scan(80 + k, n=2500, scale=2.5)
        origin = origin + 0.037 * k
        o.insertPointCloud(pts, origin, "s")
        f1.insert(pts, origin, 4.0, threads=1)
        f3.insert(pts, origin, 4.0, threads=3)
    so = o.stats()
    active = popcount64(o.exportMap().active)
    for f in (f1, f3):
        sf = f.stats()
        assert sf["visits"] == so["visits"] and sf["voxel_updates"] == so["voxel_updates"]
        assert sf["map_leaves"] == o.mapLeafCount() and sf["active_voxels"] == active


def test_oracle_vs_pyref_fuzz():
    """Property test over resolutions, ranges, far-away / negative origins and points that sit exactly on voxel
    boundaries (the fmod branch of worldToIndex): the oracle's update grid equals the pure-Python restatement's."""
    from hypothesis import given, settings, strategies as st, HealthCheck

    @settings(max_examples=40, deadline=None, suppress_health_check=list(HealthCheck))
    @given(res=st.sampled_from([0.02, 0.05, 0.1, 0.25, 1.0]), rng_m=st.sampled_from([0.6, 1.5, 4.0]),
           ox=st.floats(-50, 50), oy=st.floats(-50, 50), oz=st.floats(-5, 5), seed=st.integers(0, 2**31 - 1),
           snap=st.booleans())
    def run(res, rng_m, ox, oy, oz, seed, snap):
        rng = np.random.default_rng(seed)
        origin = np.array([ox, oy, oz])
        if snap:
            origin = np.round(origin / res) * res  # origin on a voxel corner
        pts = (origin + rng.uniform(-1.5 * rng_m, 1.5 * rng_m, size=(12, 3)) * np.array([1, 1, 0.4])).astype(np.float32)
        if snap:
            pts[::2] = (np.round(pts[::2].astype(np.float64) / res) * res).astype(np.float32)
        pts[3] = pts[5]  # duplicate ray
        pts[7] = origin.astype(np.float32)  # (almost) zero-length ray
        o = OracleOccupancyVDBMapping(res)
        assert o.setConfig(rng_m, 0.7, 0.4, 0.12, 0.97) == 0
        o.addInputSource("s", rng_m, 0)
        assert o.accumulateUpdate(pts, origin, "s") == 0
        upd_o = pyref.leafset_to_voxels(o.exportUpdateGrid("s"))
        upd_p = pyref.raycast(pts, origin, res, rng_m)
        assert {k: v[1] for k, v in upd_o.items()} == upd_p

    run()
