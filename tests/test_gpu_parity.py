"""GPU parity tests: the CUDA path (through the C ABI) against the CPU oracle on the same inputs.
Bar: bit-exact leaf sets — voxel coordinates, active masks, hit masks, change masks, f32 log-odds bit patterns."""
import json
import os

import numpy as np
import pytest

from helpers import CFG_GTEST, CFG_ROS, assert_leafsets_equal, popcount64
from test_oracle_kat import KATS, run_kat

pytestmark = pytest.mark.gpu


def _mk(kind, res, **kw):
    if kind == "gpu":
        from vdb_mapping_b200.mapping import OccupancyVDBMapping
        return OccupancyVDBMapping(res, **kw)
    from oracle.oracle import OracleOccupancyVDBMapping
    m = OracleOccupancyVDBMapping(res)
    if "replicate_probe_quirk" in kw:
        m.setProbeQuirk(kw["replicate_probe_quirk"])
    return m


def _pair(res, max_range, cfg, sources=("s",), **kw):
    g, o = _mk("gpu", res, **kw), _mk("cpu", res, **{k: v for k, v in kw.items() if k == "replicate_probe_quirk"})
    for m in (g, o):
        assert m.setConfig(max_range, *cfg[:2], cfg[2], cfg[3]) == 0
        for s in sources:
            m.addInputSource(s, max_range, 0)
    return g, o


@pytest.mark.parametrize("case", KATS["cases"], ids=[c["name"] for c in KATS["cases"]])
def test_reference_kats_on_gpu(case):
    """The reference's own gtest cases (tests/mapping.cpp) through the CUDA path."""
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    run_kat(OccupancyVDBMapping, case, KATS["config"])


def test_logodds_constants_match():
    g, o = _pair(0.1, 10.0, CFG_ROS)
    assert np.array_equal(g.logodds().view(np.uint32), o.logodds().view(np.uint32))
    assert g.setConfig(-1, 0.9, 0.1, 0.49, 0.51) == 1
    assert g.setConfig(10, 0.9, 0.6, 0.49, 0.51) == 2
    assert g.setConfig(10, 0.4, 0.1, 0.49, 0.51) == 2


@pytest.mark.parametrize("cfg", [CFG_GTEST, CFG_ROS], ids=["gtest_cfg", "ros_cfg"])
@pytest.mark.parametrize("quirk", [True, False], ids=["quirk", "noquirk"])
def test_small_sequences_bit_exact(cfg, quirk):
    """Random multi-scan sequences with ties, axis-aligned, clipped, zero-length, NaN rays and a moving origin:
    update grid, change grid and map identical after every scan."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, cfg, replicate_probe_quirk=quirk)
    for k in range(8):
        pts, origin = scans.small_scan(100 + k, n=3000, scale=2.5)
        origin = origin + np.array([0.137 * k, 0.061 * k, 0.013 * k])
        assert g.accumulateUpdate(pts, origin, "s") == 0 and o.accumulateUpdate(pts, origin, "s") == 0
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"update grid scan {k}")
        g.integrateUpdate(keep_change=True)
        o.integrateUpdate()
        assert_leafsets_equal(g.exportLastChange("s"), o.exportLastChange("s"), f"change grid scan {k}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map scan {k}")
        assert len(g.exportUpdateGrid("s")) == 0
    sg, so = g.stats(), o.stats()
    for key in ("rays", "clipped", "visits", "voxel_updates", "state_changes"):
        assert sg[key] == so[key], key
    assert sg["map_leaves"] == o.mapLeafCount()


def test_accumulation_over_several_scans_and_sources():
    """Several accumulateUpdate calls before one integrate (dedup across scans) and two sources applied in
    std::map key order (VDBMapping.hpp:380)."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.05, 3.0, CFG_GTEST, sources=("b_src", "a_src"))
    for rep in range(3):
        for i, s in enumerate(("b_src", "a_src", "b_src")):
            pts, origin = scans.small_scan(7 * rep + i, n=2000, scale=1.5)
            g.accumulateUpdate(pts, origin, s)
            o.accumulateUpdate(pts, origin, s)
        for s in ("a_src", "b_src"):
            assert_leafsets_equal(g.exportUpdateGrid(s), o.exportUpdateGrid(s), f"update {s}")
        g.integrateUpdate(keep_change=True)
        o.integrateUpdate()
        for s in ("a_src", "b_src"):
            assert_leafsets_equal(g.exportLastChange(s), o.exportLastChange(s), f"change {s}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")


def test_sources_on_their_own_handles_integrate_in_place():
    """vdbm_integrate_from: every source raycasts on a raycast-only handle of its own (one-leaf map), the map's handle runs
    updateMap per source in key order on the holders' grids where they lie. Map, change grids and counters equal the oracle's
    two-source run; the holders are left with fresh update grids."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    g, o = _pair(0.05, 3.0, CFG_GTEST, sources=("b_src", "a_src"))
    holders = {}
    for s in ("a_src", "b_src"):
        h = OccupancyVDBMapping(0.05, map_capacity_leaves=1)
        assert h.setConfig(3.0, 0.7, 0.4, 0.12, 0.97) == 0          # only the range matters to a raycast
        h.addInputSource(s, 3.0, 0)
        holders[s] = h
    for rep in range(3):
        for i, s in enumerate(("b_src", "a_src", "b_src")):
            pts, origin = scans.small_scan(70 * rep + i, n=2500, scale=1.5)
            holders[s].accumulateUpdate(pts, origin, s)
            o.accumulateUpdate(pts, origin, s)
        for s in ("a_src", "b_src"):
            assert_leafsets_equal(holders[s].exportUpdateGrid(s), o.exportUpdateGrid(s), f"update {s}")
        for s in ("a_src", "b_src"):                                # std::map key order, VDBMapping.hpp:380
            g.integrateFrom(holders[s], s, keep_change=True)
        o.integrateUpdate()
        for s in ("a_src", "b_src"):
            assert_leafsets_equal(holders[s].exportLastChange(s), o.exportLastChange(s), f"change {s}")   # stays with the holder
            assert len(holders[s].exportUpdateGrid(s)) == 0
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map {rep}")
    st, so = g.stats(), o.stats()
    assert st["voxel_updates"] == so["voxel_updates"] and st["rays"] == 0
    assert sum(h.stats()["visits"] for h in holders.values()) == so["visits"]
    # the map's own grid of a source integrates through the same call (holder == map); unknown source / foreign resolution refused
    pts, origin = scans.small_scan(999, n=1500, scale=1.5)
    g.accumulateUpdate(pts, origin, "a_src"); o.accumulateUpdate(pts, origin, "a_src")
    g.integrateFrom(g, "a_src"); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after holder == map")
    from vdb_mapping_b200.mapping import VdbmError
    with pytest.raises(VdbmError):
        g.integrateFrom(holders["a_src"], "nobody")
    other = OccupancyVDBMapping(0.1, map_capacity_leaves=1)
    with pytest.raises(VdbmError):
        g.integrateFrom(other, "a_src")
    for h in list(holders.values()) + [other]:
        h.close()


def test_clamping_and_state_flips_after_repeated_hits():
    """Same scan 12 times: values run into +-clamp, flags flip once; bit-exact all the way."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 5.0, CFG_ROS)
    pts, origin = scans.small_scan(5, n=1500, scale=2.0)
    for k in range(12):
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        chg = g.updateMap("s")
        o.integrateUpdate()
        assert_leafsets_equal(chg, o.exportLastChange("s"), f"change {k}")
    mg = g.exportMap()
    assert_leafsets_equal(mg, o.exportMap(), "map")
    lo = g.logodds()
    assert mg.values.max() == lo[4] and mg.values.min() == lo[5]


def test_degenerate_miss_noop_config():
    """prob_miss = 0.5 (log-odds 0) with thres_min > 0.5: OpenVDB's tile probe does not create leaves for misses."""
    from vdb_mapping_b200 import scans
    cfg = (0.9, 0.5, 0.6, 0.7)
    g, o = _pair(0.1, 3.0, cfg)
    for k in range(3):
        pts, origin = scans.small_scan(40 + k, n=800, scale=1.5)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        chg = g.updateMap("s")
        o.integrateUpdate()
        assert_leafsets_equal(chg, o.exportLastChange("s"), f"change {k}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map {k}")


def test_edge_inputs():
    g, o = _pair(0.1, 2.0, CFG_GTEST)
    empty = np.zeros((0, 4), np.float32)
    allnan = np.full((64, 4), np.nan, np.float32)
    same = np.zeros((5, 4), np.float32)                       # zero-length rays: endpoint only
    inf = np.array([[np.inf, 0, 0, 1], [0, -np.inf, 0, 1]], np.float32)
    one = np.array([[0.31, -0.2, 0.77, 1]], np.float32)
    for pts in (empty, allnan, same, one):
        assert g.accumulateUpdate(pts, [0, 0, 0], "s") == 0 and o.accumulateUpdate(pts, [0, 0, 0], "s") == 0
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "update")
    g.accumulateUpdate(inf, [0, 0, 0], "s"); o.accumulateUpdate(inf, [0, 0, 0], "s")   # dropped (undefined in the reference)
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "update after inf")
    assert g.stats()["nan_skipped"] == o.stats()["nan_skipped"] == 64 + 2
    # NaN origin: every point skipped (VDBMapping.hpp:505-510)
    g.accumulateUpdate(one, [np.nan, 0, 0], "s"); o.accumulateUpdate(one, [np.nan, 0, 0], "s")
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "nan origin")
    g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")
    # unknown source: no-op with status 1, unconfigured map: status 2
    assert g.accumulateUpdate(one, [0, 0, 0], "nope") == 1
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    fresh = OccupancyVDBMapping(0.1)
    fresh.addInputSource("s", 1.0)
    assert fresh.accumulateUpdate(one, [0, 0, 0], "s") == 2
    assert fresh.insertPointCloud(one, [0, 0, 0], "s") is True and fresh.mapLeafCount() == 0
    # source max_range <= 0: nothing inserted (VDBMapping.hpp:331)
    g.addInputSource("neg", -1.0)
    assert g.accumulateUpdate(one, [0, 0, 0], "neg") == 0 and len(g.exportUpdateGrid("neg")) == 0


def test_hash_growth_paths():
    """Tiny initial capacities force update-hash overflow + replay, and map pool / hash growth."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.05, 6.0, CFG_GTEST, update_capacity_leaves=64, map_capacity_leaves=32)
    for k in range(3):
        pts, origin = scans.small_scan(900 + k, n=6000, scale=3.0)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"update {k}")
        g.integrateUpdate(); o.integrateUpdate()
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map {k}")
    st = g.stats()
    assert st["update_capacity"] > 8 * 512 and st["map_capacity"] > 32   # 64 leaves -> 8 bricks of 512 leaves initially
    assert st["visits"] == o.stats()["visits"]          # replay must not double count


def test_sections_and_dirty_export():
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_GTEST)
    for k in range(3):
        pts, origin = scans.small_scan(300 + k, n=4000, scale=2.5)
        g.insertPointCloud(pts, origin, "s"); o.insertPointCloud(pts, origin, "s")
    for mn, mx in [((-9, -4, -3), (6, 11, 5)), ((-100, -100, -100), (100, 100, 100)), ((0, 0, 0), (0, 0, 0)),
                   ((3, 3, 3), (2, 2, 2)), ((-17, 5, -8), (-9, 23, 7))]:
        for full in (False, True):
            assert_leafsets_equal(g.getMapSectionUpdateGrid(mn, mx, full), o.getMapSectionUpdateGrid(mn, mx, full), f"section U {mn} {full}")
            assert_leafsets_equal(g.getMapSectionGrid(mn, mx, full), o.getMapSectionGrid(mn, mx, full), f"section F {mn} {full}")
    # dirty export: after a full export, only leaves touched by the next scan come back, with current content
    full = g.exportMap()
    pts, origin = scans.small_scan(999, n=500, scale=1.0)
    g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
    touched = o.exportUpdateGrid("s")
    g.integrateUpdate(); o.integrateUpdate()
    dirty = g.exportMap(dirty_only=True)
    now = o.exportMap()
    idx = {tuple(x): i for i, x in enumerate(now.origins)}
    sel = [idx[tuple(x)] for x in dirty.origins]
    assert np.array_equal(dirty.values.view(np.uint32), now.values[sel].view(np.uint32))
    assert np.array_equal(dirty.active, now.active[sel])
    # dirty = touched leaves; the contract only promises that nothing MODIFIED is missing: a touched leaf that is not reported
    # must be bit for bit what it was before the scan
    was = {tuple(x): i for i, x in enumerate(full.origins)}
    reported, touched_set = {tuple(x) for x in dirty.origins}, {tuple(x) for x in touched.origins}
    assert reported <= touched_set and len(reported) > 0
    for key in touched_set - reported:
        i, j = was[key], idx[key]
        assert np.array_equal(full.values[i].view(np.uint32), now.values[j].view(np.uint32)) and np.array_equal(full.active[i], now.active[j])
    assert len(g.exportMap(dirty_only=True)) == 0


def test_mirror_stream_follows_the_map():
    """vdbm_map_mirror (the shim's host mirror): a host copy kept up to date ONLY from the chunked stream, addressed through the
    pool-index table, equals the oracle's map after every scan; chunks are index-ordered, an aborted transfer loses nothing, a
    reset starts a new generation."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS)
    host = {}      # pool index -> [origin, values, active]
    seen_chunks = []

    def sink(idx, origins, values, active):
        assert np.all(np.diff(idx.astype(np.int64)) > 0)
        seen_chunks.append(len(idx))
        for i, l in enumerate(idx):
            if int(l) in host:
                assert tuple(host[int(l)][0]) == tuple(origins[i])     # an index never changes its leaf
            host[int(l)] = [origins[i].copy(), values[i].copy(), active[i].copy()]
        return 0

    def host_leafset():
        items = sorted(host.values(), key=lambda t: tuple(t[0]))
        return (np.array([t[0] for t in items]), np.array([t[1] for t in items]), np.array([t[2] for t in items]))

    gen0 = g.mapGeneration()
    last = 0
    for k in range(5):
        pts, origin = scans.small_scan(400 + k, n=6000, scale=2.5)
        origin = origin + np.array([0.21 * k, -0.13 * k, 0.02 * k])
        g.insertPointCloud(pts, origin, "s"); o.insertPointCloud(pts, origin, "s")
        seen_chunks.clear()
        n = g.mirrorMap(sink, chunk_leaves=(64, 100, 1 << 15, 7, 0)[k])       # many small chunks, one big chunk, default
        assert n == sum(seen_chunks) and n > 0
        if k == 0:
            assert max(seen_chunks) <= 64 and len(seen_chunks) >= 3            # the rotation of the three buffers was exercised
        ref = o.exportMap()
        ho, hv, ha = host_leafset()
        assert np.array_equal(ho, ref.origins) and np.array_equal(ha, ref.active)
        assert np.array_equal(hv.view(np.uint32), ref.values.view(np.uint32))
        assert sorted(host) == list(range(len(host))) and len(host) >= last    # dense pool indices, append-only
        last = len(host)
        assert g.mirrorMap(sink) == 0                                          # nothing is dirty any more
    # abort after the first chunk: the rest stays dirty and arrives with the next call
    pts, origin = scans.small_scan(999, n=6000, scale=2.5)
    g.insertPointCloud(pts, origin, "s"); o.insertPointCloud(pts, origin, "s")
    from vdb_mapping_b200.mapping import VdbmError
    calls = []

    def quitter(idx, origins, values, active):
        calls.append(len(idx))
        return 1

    with pytest.raises(VdbmError):
        g.mirrorMap(quitter, chunk_leaves=32)
    assert len(calls) == 1
    n = g.mirrorMap(sink, chunk_leaves=50)
    assert n >= calls[0]
    ref = o.exportMap()
    ho, hv, ha = host_leafset()
    assert np.array_equal(ho, ref.origins) and np.array_equal(ha, ref.active) and np.array_equal(hv.view(np.uint32), ref.values.view(np.uint32))
    # the dirty export and the mirror share the flags
    g.insertPointCloud(pts, origin, "s")
    assert len(g.exportMap(dirty_only=True)) > 0 and g.mirrorMap(sink) == 0
    g.resetMap()
    assert g.mapGeneration() == gen0 + 1 and g.mirrorMap(sink) == 0


def test_update_grid_import_roundtrip_and_reset():
    """updateMap(grid produced elsewhere): export on one map, import on another, maps end identical."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS)
    g2, _ = _pair(0.1, 4.0, CFG_ROS)
    pts, origin = scans.small_scan(77, n=3000, scale=2.0)
    g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
    upd = g.exportUpdateGrid("s")
    g2.importUpdate("s", upd.origins, upd.active, upd.valmask)
    g2.importUpdate("s", upd.origins[::2], upd.active[::2], upd.valmask[::2])   # OR is idempotent
    assert_leafsets_equal(g2.exportUpdateGrid("s"), upd, "imported update grid")
    g.integrateUpdate(); g2.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g2.exportMap(), o.exportMap(), "map via import")
    g.resetMap(); o.resetMap()
    assert len(g.exportMap()) == 0 and g.probe((0, 0, 1)) == (0.0, False)
    g.insertPointCloud(pts, origin, "s"); o.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after reset")


def test_raycast_with_explicit_range():
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 1.0, CFG_ROS)
    pts, origin = scans.small_scan(12, n=2000, scale=2.0)
    assert g.raycastPointCloud(pts, origin, 2.5, "s")
    o2 = _mk("cpu", 0.1)
    o2.setConfig(2.5, *CFG_ROS); o2.addInputSource("s", 2.5)
    o2.accumulateUpdate(pts, origin, "s")
    assert_leafsets_equal(g.exportUpdateGrid("s"), o2.exportUpdateGrid("s"), "explicit range")


@pytest.mark.parametrize("cfg_id", [1, 3, 2])
def test_baseline_configs_one_scan_bit_exact(cfg_id):
    """BASELINE.json configs 1-3 at full size: one scan, update grid + map bit-exact vs the oracle, then a second
    scan from a moved origin (RMW on existing leaves)."""
    from vdb_mapping_b200 import scans
    c = scans.CONFIGS[cfg_id]
    g, o = _pair(c.resolution, c.max_range, (c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max))
    for k in range(2):
        pts, origin = scans.make_scan(cfg_id, k)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        ug, uo = g.exportUpdateGrid("s"), o.exportUpdateGrid("s")
        assert_leafsets_equal(ug, uo, f"cfg{cfg_id} update grid scan {k}")
        g.integrateUpdate(); o.integrateUpdate()
        assert g.stats()["last_voxel_updates"] == popcount64(uo.active)
    assert_leafsets_equal(g.exportMap(), o.exportMap(), f"cfg{cfg_id} map")
    sg, so = g.stats(), o.stats()
    for key in ("rays", "clipped", "visits", "voxel_updates"):
        assert sg[key] == so[key], key


def test_full_size_properties_cfg2_sequence():
    """Size-independent properties at BASELINE size (no oracle): idempotence of the update grid under re-insertion
    of the same scan, permutation invariance, visits == sum(1+|d|_1), hit voxels are active."""
    from vdb_mapping_b200 import scans
    c = scans.CONFIGS[2]
    cfg = (c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    g = OccupancyVDBMapping(c.resolution); g.setConfig(c.max_range, *cfg); g.addInputSource("s", c.max_range)
    pts, origin = scans.make_scan(2, 3)
    g.accumulateUpdate(pts, origin, "s")
    a = g.exportUpdateGrid("s")
    g.accumulateUpdate(pts, origin, "s")                      # idempotent OR
    b = g.exportUpdateGrid("s")
    assert_leafsets_equal(a, b, "idempotence")
    g.updateMap("s")
    perm = np.random.default_rng(0).permutation(pts.shape[0])
    g.accumulateUpdate(pts[perm], origin, "s")                # order independence
    assert_leafsets_equal(a, g.exportUpdateGrid("s"), "permutation invariance")
    assert np.all((a.valmask & ~a.active) == 0)               # every hit voxel is active
    st = g.stats()
    assert st["last_visits"] > 1.4e8 and st["last_touched_leaves"] == len(a)


def _sender_receiver(seed_a, seed_b):
    from vdb_mapping_b200 import scans
    pairs = []
    for seed in (seed_a, seed_b):
        g, o = _pair(0.1, 4.0, CFG_GTEST)
        for k in range(3):
            pts, origin = scans.small_scan(seed * 10 + k, n=3000, scale=2.5)
            g.insertPointCloud(pts, origin, "s"); o.insertPointCloud(pts, origin, "s")
        pairs.append((g, o))
    return pairs


@pytest.mark.parametrize("box", [((-9, -4, -3), (6, 11, 5)), ((-40, -40, -40), (40, 40, 40)), ((2, 2, 2), (2, 2, 2))])
def test_apply_map_section_update_grid(box):
    """Receiver side: applyMapSectionUpdateGrid, VDBMapping.hpp:1058-1085."""
    (ga, oa), (gb, ob) = _sender_receiver(61, 62)
    mn, mx = box
    section = ga.getMapSectionUpdateGrid(mn, mx, False)
    assert_leafsets_equal(section, oa.getMapSectionUpdateGrid(mn, mx, False), "section")
    gb.applyMapSectionUpdateGrid(mn, mx, section)
    ob.applyMapSectionUpdateGrid(mn, mx, section)
    assert_leafsets_equal(gb.exportMap(), ob.exportMap(), "map after applyMapSectionUpdateGrid")
    # the touched leaves are reported dirty (host mirror); a later scan still integrates bit-exactly
    from vdb_mapping_b200 import scans
    pts, origin = scans.small_scan(999, n=2000, scale=2.0)
    gb.insertPointCloud(pts, origin, "s"); ob.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(gb.exportMap(), ob.exportMap(), "map after a following scan")


@pytest.mark.parametrize("quirk", [True, False], ids=["tilequirk", "noquirk"])
@pytest.mark.parametrize("full", [True, False], ids=["full", "sparse"])
def test_apply_map_section_grid(full, quirk):
    """Receiver side: applyMapSectionGrid, VDBMapping.hpp:1022-1047, incl. the value-all tile visits."""
    (ga, oa), (gb, ob) = _sender_receiver(71, 72)
    mn, mx = (-12, -7, -6), (9, 14, 8)
    section = ga.getMapSectionGrid(mn, mx, full)
    gb.applyMapSectionGrid(section, tile_quirk=quirk)
    ob.applyMapSectionGrid(section, tile_quirk=quirk)
    assert_leafsets_equal(gb.exportMap(), ob.exportMap(), "map after applyMapSectionGrid")


@pytest.mark.parametrize("seg_len", ["257", "1024", "auto"])
def test_long_rays_are_split_into_segments_exactly(seg_len, monkeypatch):
    """Rays longer than 1.5 x the segment length are traversed as independent segments whose start states come
    from a per-axis pre-pass; the result must stay bit-identical to the sequential oracle (ties, negative directions,
    axis-aligned and clipped long rays, zero components)."""
    if seg_len != "auto":
        monkeypatch.setenv("VDBM_SEG_LEN", seg_len)   # forced split; "auto" plans it from the previous scan
    rng = np.random.default_rng(123)
    res, max_range = 0.01, 60.0
    g, o = _pair(res, max_range, CFG_GTEST)
    for k in range(3):
        n = 600
        d = rng.normal(size=(n, 3))
        d /= np.linalg.norm(d, axis=1, keepdims=True)
        r = rng.uniform(5.0, 80.0, size=(n, 1))                 # 500 .. 8000 voxels, some clipped at 60 m
        pts = d * r
        pts[:20, 1:] = 0.0                                       # axis-aligned long rays (two disabled axes)
        pts[20:40, 2] = 0.0                                      # planar rays
        pts[40:60] = np.round(pts[40:60] * 8) / 8                # lattice end points: exact ties in the DDA
        pts[60:70] = np.array([[40.0, 40.0, 40.0]]) * rng.choice([-1, 1], size=(10, 3))   # perfect diagonals
        origin = np.array([0.0031 * k, -0.0042 * k, 0.0017 * k])
        cloud = np.ones((n, 4), np.float32); cloud[:, :3] = (pts + origin).astype(np.float32)
        g.accumulateUpdate(cloud, origin, "s"); o.accumulateUpdate(cloud, origin, "s")
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"update grid scan {k}")
        g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")
    assert g.stats()["visits"] == o.stats()["visits"]


# ---- remote-mapping deltas: createUpdate / applyUpdate ---------------------------------------
@pytest.mark.parametrize("cfg", [CFG_GTEST, CFG_ROS], ids=["gtest_cfg", "ros_cfg"])
def test_create_and_apply_update_levels(cfg):
    """Sender GPU vs sender oracle: the three update kinds are identical. Receiver GPUs fed with the GPU-made updates
    vs receiver oracles fed with the oracle-made ones: maps and change grids identical; the level-2 (reduced) receiver
    additionally equals the SENDER's map bit for bit."""
    from vdb_mapping_b200 import scans
    gs, os_ = _pair(0.1, 4.0, cfg)
    g0, o0 = _pair(0.1, 4.0, cfg)
    g1, o1 = _pair(0.1, 4.0, cfg)
    g2, o2 = _pair(0.1, 4.0, cfg)
    for k in range(6):
        pts, origin = scans.small_scan(300 + k, n=3000, scale=2.5)
        origin = origin + np.array([0.137 * k, 0.061 * k, 0.013 * k])
        gs.accumulateUpdate(pts, origin, "s"); os_.accumulateUpdate(pts, origin, "s")
        raw_g, og = gs.createUpdate("s", 0); raw_o, oo = os_.createUpdate("s", 0)
        red_g, og2 = gs.createUpdate("s", 2); red_o, _ = os_.createUpdate("s", 2)
        assert np.array_equal(og, origin) and np.array_equal(og2, origin) and np.array_equal(oo, origin)
        assert_leafsets_equal(raw_g, raw_o, f"raw update {k}")
        assert_leafsets_equal(red_g, red_o, f"reduced update {k}")
        assert_leafsets_equal(gs.exportUpdateGrid("s"), raw_o, "createUpdate must not disturb the accumulated grid")
        gs.integrateUpdate(keep_change=True); os_.integrateUpdate()
        chg_g, _ = gs.createUpdate("s", 1); chg_o, _ = os_.createUpdate("s", 1)
        assert_leafsets_equal(chg_g, chg_o, f"overwrite grid {k}")
        c0 = g0.applyUpdate(0, raw_g, want_change=True); o0.applyUpdate("s", 0, raw_o)
        assert_leafsets_equal(c0, o0.exportLastChange("s"), f"level-0 change {k}")
        g1.applyUpdate(1, chg_g); o1.applyUpdate("s", 1, chg_o)
        c2 = g2.applyUpdate(2, red_g, origin=og2, want_change=True); o2.applyUpdate("s", 2, red_o, origin)
        assert_leafsets_equal(c2, o2.exportLastChange("s"), f"level-2 change {k}")
        assert_leafsets_equal(c2, chg_g, f"level-2 receiver change == sender change {k}")
        assert_leafsets_equal(g1.exportMap(), o1.exportMap(), f"level-1 receiver map {k}")
    sender = gs.exportMap()
    assert_leafsets_equal(sender, os_.exportMap(), "sender map")
    assert_leafsets_equal(g0.exportMap(), sender, "level-0 receiver == sender")
    assert_leafsets_equal(g2.exportMap(), sender, "level-2 receiver == sender")
    assert_leafsets_equal(g2.exportMap(), o2.exportMap(), "level-2 receiver vs oracle")


def test_reduced_update_full_size_roundtrip():
    """cfg1 at full size: sender -> reduced update -> receiver, three scans. The receiver's map equals the sender's, and
    the reduced update carries < 1/50 of the raw update's voxels."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    c = scans.CONFIGS[1]
    maps = []
    for _ in range(2):
        m = OccupancyVDBMapping(c.resolution)
        m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        m.addInputSource("s", c.max_range)
        maps.append(m)
    snd, rcv = maps
    for k in range(3):
        pts, origin = scans.make_scan(1, k)
        snd.accumulateUpdate(pts, origin, "s")
        raw_voxels = snd.stats()["last_touched_leaves"]
        red, o = snd.createUpdate("s", 2)
        assert 0 < popcount64(red.active) <= pts.shape[0]
        snd.integrateUpdate(keep_change=False)
        rcv.applyUpdate(2, red, origin=o)
        assert raw_voxels > 0
        assert snd.stats()["last_voxel_updates"] == rcv.stats()["last_voxel_updates"]
        assert popcount64(red.active) * 50 < snd.stats()["last_voxel_updates"]
    assert_leafsets_equal(rcv.exportMap(), snd.exportMap(), "receiver map after 3 reduced updates")


def test_reduced_update_requires_last_accumulate_and_handles_empty():
    from vdb_mapping_b200.mapping import OccupancyVDBMapping, VdbmError
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS, sources=("a", "b"))
    with pytest.raises(VdbmError):
        g.createUpdate("a", 2)  # nothing accumulated yet
    pts, origin = scans.small_scan(1, n=500, scale=2.0)
    g.accumulateUpdate(pts, origin, "a")
    with pytest.raises(VdbmError):
        g.createUpdate("b", 2)  # the reduced update belongs to source a's accumulate
    red, og = g.createUpdate("a", 2)
    assert len(red) > 0
    # all-NaN cloud: empty reduced update, applying it is a no-op
    nan = np.full((10, 3), np.nan, dtype=np.float32)
    g.integrateUpdate()
    g.accumulateUpdate(nan, origin, "a")
    red2, _ = g.createUpdate("a", 2)
    assert len(red2) == 0
    r = OccupancyVDBMapping(0.1); r.setConfig(4.0, *CFG_ROS)
    r.applyUpdate(2, red2, origin=origin)
    assert r.mapLeafCount() == 0
    r.applyUpdate(2, red, origin=og)
    assert r.mapLeafCount() > 0


# ---- direct edits + artificial areas (the tail of updateMap V:785-789) ---------------------------
def test_add_and_remove_points():
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS)
    rng = np.random.default_rng(11)
    add = rng.uniform(-3, 3, size=(5000, 3)).astype(np.float32)
    add[::25] = np.round(add[::25] * 10) / 10
    add[7] = np.nan
    for m in (g, o):
        m.addPointsToGrid(add)
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "after addPointsToGrid")
    pts, origin = scans.small_scan(5, n=2000, scale=2.5)
    for m in (g, o):
        m.insertPointCloud(pts, origin, "s")
    rem = np.concatenate([add[:1500], rng.uniform(-3, 3, size=(1500, 3)).astype(np.float32)])
    for m in (g, o):
        m.removePointsFromGrid(rem)
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "after removePointsFromGrid")
    for m in (g, o):
        m.insertPointCloud(pts, origin + 0.05, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "scan after edits")


@pytest.mark.parametrize("cfg", [CFG_GTEST, CFG_ROS, (0.7, 0.4, 0.12, 0.45)], ids=["gtest_cfg", "ros_cfg", "neg_thres_max"])
def test_artificial_areas(cfg):
    """addArtificialAreas -> artificial grid; every updateMap then re-activates the walls; restoreMapIntegrity. The third
    config has thres_max < 0 (prob 0.45), where restoring a wall voxel in a missing leaf CREATES the leaf."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, cfg)
    polys = [np.array([[0.52, 0.5, 0.0, 1], [2.31, 0.77, 0.0, 1], [1.9, 2.2, 0.0, 1], [0.4, 1.9, 0.1, 1]]),
             np.array([[-1.0, -1.0, 0.2, 1], [-2.0, -1.5, 0.2, 1]]), np.array([[6.0, 6.0, 1.0, 1]])]
    for m in (g, o):
        m.addArtificialAreas(polys, -0.35, 0.55)
    art = g.exportArtificialAreaGrid()
    assert len(art) > 0
    assert_leafsets_equal(art, o.exportArtificialAreaGrid(), "artificial grid")
    assert g.mapLeafCount() == 0
    for k in range(3):
        pts, origin = scans.small_scan(20 + k, n=2500, scale=2.5)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        g.integrateUpdate(keep_change=True); o.integrateUpdate()
        assert_leafsets_equal(g.exportLastChange("s"), o.exportLastChange("s"), f"change {k}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map with walls {k}")
    # replacing the areas restores the old walls first
    for m in (g, o):
        m.addArtificialAreas(polys[1:], -0.2, 0.3)
    assert_leafsets_equal(g.exportArtificialAreaGrid(), o.exportArtificialAreaGrid(), "second artificial grid")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after replacing areas")
    pts, origin = scans.small_scan(30, n=2500, scale=2.5)
    for m in (g, o):
        m.insertPointCloud(pts, origin, "s")
        m.restoreMapIntegrity()
    assert len(g.exportArtificialAreaGrid()) == 0
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after restoreMapIntegrity")
    # resetMap keeps the artificial grid (V:174-186 does not touch it)
    for m in (g, o):
        m.addArtificialAreas(polys[:1], -0.2, 0.3)
        m.resetMap()
        m.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "walls survive resetMap")


# ---- asynchronous scan pipeline (vdbm_insert_async) ------------------------------------------------------------------
@pytest.mark.parametrize("small_tables", [False, True], ids=["default_capacity", "tiny_capacity"])
def test_async_pipeline_matches_the_oracle(small_tables):
    """insertPointCloudAsync queues whole scans (upload on a copy stream, device-side capacity guard, one host sync per
    scan, taken by the NEXT call). Whatever mix of queued / redone / synchronous scans results - tiny tables force
    redos - the map must equal the oracle's after every flush, and the counters must agree."""
    from vdb_mapping_b200 import scans
    kw = dict(update_capacity_leaves=1024, map_capacity_leaves=256) if small_tables else {}
    res = 0.05 if small_tables else 0.1
    g, o = _pair(res, 4.0, CFG_ROS, **kw)
    # 300 k short rays per scan (clipped at 4 m): many more rays than resident DDA lanes and no ray longer than twice a
    # lane's share, i.e. scans the planner sends down the QUEUED path (3,000-ray clouds never are). Tiny tables: a few
    # scans of small extent keep the 16-brick update hash small, then a scan of the full 4 m extent overflows it while it is
    # already queued behind the device-side guard -> refused -> redone by the finishing call (with the next raycast in flight).
    def cloud(k):
        n = 300000 + 2000 * (k % 3)
        if small_tables and k in (0, 1, 2, 3, 5, 6):
            # small extent, no long rays at all (ray length 0.2 .. 1.0 m): the planner keeps the NEXT scan on the queued path
            rng = np.random.default_rng(900 + k)
            d = rng.normal(size=(n, 3))
            d /= np.linalg.norm(d, axis=1, keepdims=True)
            origin = rng.normal(size=3) * 0.3
            return (origin + d * rng.uniform(0.2, 1.0, size=(n, 1))).astype(np.float32), origin
        return scans.small_scan(500 + k, n=n, scale=2.5 + 0.3 * (k % 4))

    for k in range(12):
        pts, origin = cloud(k)
        origin = origin + np.array([0.137 * k, 0.061 * k, 0.013 * k])
        g.insertPointCloudAsync(pts, origin, "s")
        pts[:] = np.nan  # the host buffer is free again as soon as the call returns
        pts2, _ = cloud(k)
        o.insertPointCloud(pts2, origin, "s")
        if k % 4 == 3:
            assert_leafsets_equal(g.exportMap(), o.exportMap(), f"map after scan {k}")  # exportMap finishes the queued scan
    g.flush()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "final map")
    sg, so = g.stats(), o.stats()
    for key in ("rays", "clipped", "visits", "voxel_updates"):
        assert sg[key] == so[key], key
    assert sg["map_leaves"] == o.mapLeafCount()
    # the paths this test is about were really taken: scans queued with their raycast half overlapping the previous scan's
    # update, and (tiny tables) scans the device-side guard refused and the finishing call redid - with a raycast in flight
    pc = g.pipelineCounts()
    assert pc["queued"] + pc["redone"] + pc["synchronous"] == 12
    assert pc["overlapped"] >= 4, pc
    if small_tables:
        assert pc["redone"] >= 1, pc


def test_async_pipeline_full_size_and_interleaved_calls():
    """cfg2 at full size (the scans the queued path is made for): async inserts interleaved with sections and a synchronous
    insert of a second source equal the all-synchronous run."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    c = scans.CONFIGS[2]
    maps = []
    for _ in range(2):
        m = OccupancyVDBMapping(c.resolution)
        m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        m.addInputSource("s", c.max_range); m.addInputSource("t", c.max_range)
        maps.append(m)
    a, b = maps
    for k in range(10):
        pts, origin = scans.make_scan(2, k)
        a.insertPointCloudAsync(pts, origin, "s")
        b.insertPointCloud(pts, origin, "s")
        if k == 4:
            lo, hi = np.array([-80, -80, -20], np.int32), np.array([80, 80, 20], np.int32)
            assert_leafsets_equal(a.getMapSectionUpdateGrid(lo, hi), b.getMapSectionUpdateGrid(lo, hi), "section mid-pipeline")
        if k == 6:
            p2, o2 = scans.make_scan(2, 100)
            a.insertPointCloud(p2[:5000], o2, "t"); b.insertPointCloud(p2[:5000], o2, "t")
    assert_leafsets_equal(a.exportMap(), b.exportMap(), "async vs sync map")
    assert a.stats()["visits"] == b.stats()["visits"] and a.stats()["voxel_updates"] == b.stats()["voxel_updates"]
    pc = a.pipelineCounts()
    assert pc["queued"] >= 4 and pc["overlapped"] >= 3, pc


def test_prefetched_cloud_is_used_and_stale_prefetch_is_ignored():
    """vdbm_prefetch uploads the next cloud on the copy stream; vdbm_accumulate with the same pointer uses that copy, with
    another pointer it copies normally. Results equal the oracle either way."""
    import torch
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS)
    clouds = [scans.small_scan(700 + k, n=2500, scale=2.5) for k in range(4)]
    pinned = [torch.from_numpy(np.ascontiguousarray(np.concatenate([p[:, :3], np.ones((len(p), 1), np.float32)], axis=1))).pin_memory()
              for p, _ in clouds]
    g.prefetchRaw(pinned[0].data_ptr(), pinned[0].shape[0])
    for k, (pts, origin) in enumerate(clouds):
        if k == 2:
            g.prefetchRaw(pinned[3].data_ptr(), pinned[3].shape[0])  # stale: the next accumulate passes another pointer
        g.accumulateRaw(pinned[k].data_ptr(), pinned[k].shape[0], origin, "s")
        if k == 0:
            g.prefetchRaw(pinned[1].data_ptr(), pinned[1].shape[0])
        g.integrateUpdate(keep_change=False)
        o.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")


def test_large_exports_are_views_that_outlive_calls_and_the_map():
    """Leaf sets >= 16 MB come back as zero-copy numpy views of the library's pinned staging memory. They must stay valid
    across later exports (which then use other memory) and after the map is closed."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    c = scans.CONFIGS[1]
    m = OccupancyVDBMapping(c.resolution)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    m.addInputSource("s", c.max_range)
    pts, origin = scans.make_scan(1, 0)
    m.insertPointCloud(pts, origin, "s")
    a = m.exportMap()
    assert a.values.nbytes >= (16 << 20) and a.values.base is not None  # a view, not a copy
    snap = a.values.copy(); snap_act = a.active.copy()
    b = m.exportMap()            # second large export while the first still borrows the staging buffer
    assert np.array_equal(b.values.view(np.uint32), snap.view(np.uint32))
    pts2, origin2 = scans.make_scan(1, 1)
    m.insertPointCloud(pts2, origin2, "s")
    c2 = m.exportMap()
    assert len(c2) >= len(a)
    m.close()                    # the map goes away first
    assert np.array_equal(a.values.view(np.uint32), snap.view(np.uint32)) and np.array_equal(a.active, snap_act)
    keep = a.values[5]
    del a, b, c2
    import gc; gc.collect()
    assert np.array_equal(keep.view(np.uint32), snap[5].view(np.uint32))


def test_prefetch_mixed_with_queued_scans_does_not_clobber_the_cloud_in_flight():
    """A prefetch issued while a queued scan is still reading its staging buffer must go to the OTHER buffer, also when the
    next call is another vdbm_insert_async rather than the accumulate the prefetch was meant for."""
    import torch
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    c = scans.CONFIGS[1]
    g = OccupancyVDBMapping(c.resolution); ref = OccupancyVDBMapping(c.resolution)
    for m in (g, ref):
        m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        m.addInputSource("s", c.max_range)
    clouds = [scans.make_scan(1, k) for k in range(6)]
    pinned = [torch.from_numpy(np.ascontiguousarray(np.concatenate([p[:, :3], np.ones((len(p), 1), np.float32)], axis=1))).pin_memory()
              for p, _ in clouds]
    junk = torch.full_like(pinned[0], 7.0).pin_memory()
    for k, (pts, origin) in enumerate(clouds):
        g.insertRawAsync(pinned[k].data_ptr(), pinned[k].shape[0], origin, "s")
        g.prefetchRaw(junk.data_ptr(), junk.shape[0])  # never consumed
        ref.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), ref.exportMap(), "map")


@pytest.mark.parametrize("tbs", ["1", "auto"])
def test_test_before_set_marking_is_exact(tbs, monkeypatch):
    """Scans whose rays overlap heavily (depth camera) load a mask word and skip the RED when its bits are already set
    (raycast_dda_kernel<4>). Forced on for random LiDAR-like scans, and chosen automatically on BASELINE config 3 at full
    size (47 visits per distinct voxel): update grid, change grid and map stay bit-identical to the oracle."""
    from vdb_mapping_b200 import scans
    if tbs != "auto":
        monkeypatch.setenv("VDBM_DDA_TBS", tbs)
        g, o = _pair(0.1, 4.0, CFG_GTEST)
        for k in range(5):
            pts, origin = scans.small_scan(900 + k, n=4000, scale=2.5)
            g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
            assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"update grid {k}")
            g.integrateUpdate(keep_change=True); o.integrateUpdate()
            assert_leafsets_equal(g.exportLastChange("s"), o.exportLastChange("s"), f"change {k}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")
        return
    c = scans.CONFIGS[3]
    g, o = _pair(c.resolution, c.max_range, (c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max))
    for k in range(3):  # the first scan runs the plain kernel, the overlap ratio it reports switches the next ones over
        pts, origin = scans.make_scan(3, k)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"cfg3 update grid {k}")
        g.integrateUpdate(keep_change=False); o.integrateUpdate()
        st = g.stats()
        assert st["last_visits"] > 8 * st["last_voxel_updates"]
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "cfg3 map")


def test_many_invalid_points_do_not_starve_very_short_rays():
    """Regression: the longest-first sort only looks at key bits [4, 20), so rays with 1..15 marks used to tie with the
    empty work items (NaN points). With more empty items than resident lanes (a depth frame with > 60 % invalid pixels),
    every lane retired on a zero key before the short rays behind them were fetched, and those rays were lost."""
    g, o = _pair(0.1, 4.0, CFG_ROS)
    rng = np.random.default_rng(3)
    n_nan, n_short, n_long = 260_000, 120_000, 20_000
    origin = np.array([0.03, 0.02, 0.01])
    short = (origin + rng.uniform(-0.45, 0.45, size=(n_short, 3))).astype(np.float32)   # <= ~13 marks each
    long_ = (origin + rng.uniform(-3.0, 3.0, size=(n_long, 3))).astype(np.float32)
    pts = np.concatenate([np.full((n_nan, 3), np.nan, np.float32), short, long_])
    assert g.accumulateUpdate(pts, origin, "s") == 0 and o.accumulateUpdate(pts, origin, "s") == 0
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "update grid")
    g.integrateUpdate(keep_change=False); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")
    assert g.stats()["visits"] == o.stats()["visits"]


def test_prefetch_is_one_shot_and_async_insert_of_unknown_source_still_integrates():
    import torch
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS, sources=("s",))
    for m in (g, o):
        m.addInputSource("off", -1.0, 0)  # negative range: accumulateUpdate ignores the cloud (VDBMapping.hpp:331)
    (pa, origin), (pb, _) = scans.small_scan(810, n=2000, scale=2.5), scans.small_scan(811, n=2000, scale=2.5)
    buf = torch.ones((2000, 4), dtype=torch.float32).pin_memory()
    buf[:, :3] = torch.from_numpy(pa[:, :3])
    g.prefetchRaw(buf.data_ptr(), 2000)
    g.accumulateRaw(buf.data_ptr(), 2000, origin, "off")       # ignored call: must consume the prefetched copy anyway
    g.synchronize()
    buf[:, :3] = torch.from_numpy(pb[:, :3])                   # same address, new cloud
    g.accumulateRaw(buf.data_ptr(), 2000, origin, "s")
    o.accumulateUpdate(pa, origin, "off"); o.accumulateUpdate(pb, origin, "s")
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "update grid must come from the refilled buffer")
    # insertPointCloud on an unknown source: accumulateUpdate complains, integrateUpdate still runs (V:399-406)
    g.insertPointCloudAsync(pa, origin, "nope"); o.insertPointCloud(pa, origin, "nope")
    assert len(g.exportUpdateGrid("s")) == 0
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after the integrate triggered by the unknown-source insert")


def test_cfg4_sector_at_real_resolution_bit_exact():
    """BASELINE configs[3] at its own resolution / range (0.02 m, 100 m): a 1/8 azimuth sector of the 1M-point hall scan
    that CONTAINS a door (rays clipped at 100 m: ~5,000-voxel rays, split into segments by the planner on "auto"), two
    scans from a moving origin. Update grid, map and counters bit-exact vs the oracle."""
    from vdb_mapping_b200 import scans, dist as vdist
    c = scans.CONFIGS[4]
    g, o = _pair(c.resolution, c.max_range, (c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max))
    for k in range(2):
        pts, origin = scans.make_scan(4, k)
        ang = vdist.diamond_angle(pts[:, 0].astype(np.float64) - origin[0], pts[:, 1].astype(np.float64) - origin[1])
        sel = np.nonzero(~np.isfinite(ang) | ((ang >= 0.0) & (ang < 0.5)))[0]   # 1/8 of the circle around the door at az 0.3 rad
        sector = np.ascontiguousarray(pts[sel])
        assert 100_000 < len(sector) < 200_000
        g.accumulateUpdate(sector, origin, "s"); o.accumulateUpdate(sector, origin, "s")
        assert g.stats()["clipped"] > 0
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"cfg4 sector update grid scan {k}")
        g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "cfg4 sector map")
    sg, so = g.stats(), o.stats()
    for key in ("rays", "clipped", "visits", "voxel_updates"):
        assert sg[key] == so[key], key
    # the longest (door) rays really were long: the planner had something to split
    assert sg["visits"] / max(1, sg["rays"]) > 500


def test_cfg5_remote_mapping_full_scan_vs_oracle():
    """BASELINE configs[4] at full size: one cfg2 scan (262,144 rays) through createUpdate(2) on the sender and
    applyUpdate(2) on a second GPU map; the reduced update and the RECEIVER's map are bit-exact vs the oracle's receiver."""
    from vdb_mapping_b200 import scans
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    from oracle.oracle import OracleOccupancyVDBMapping
    c = scans.CONFIGS[2]
    cfg = (c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    snd, o_snd = _pair(c.resolution, c.max_range, cfg)
    rcv = OccupancyVDBMapping(c.resolution)
    o_rcv = OracleOccupancyVDBMapping(c.resolution)
    for m in (rcv, o_rcv):
        m.setConfig(c.max_range, *cfg)
        m.addInputSource("s", c.max_range)
    pts, origin = scans.make_scan(2, 0)
    snd.accumulateUpdate(pts, origin, "s"); o_snd.accumulateUpdate(pts, origin, "s")
    red, og = snd.createUpdate("s", 2)
    red_o, og_o = o_snd.createUpdate("s", 2)
    assert_leafsets_equal(red, red_o, "reduced (level 2) update")
    assert np.array_equal(og, og_o)
    snd.integrateUpdate(keep_change=False); o_snd.integrateUpdate()
    rcv.applyUpdate(2, red, origin=og)
    o_rcv.applyUpdate("s", 2, red_o, og_o)
    assert_leafsets_equal(rcv.exportMap(), o_rcv.exportMap(), "receiver map (GPU) vs receiver map (oracle)")
    assert_leafsets_equal(rcv.exportMap(), snd.exportMap(), "receiver map vs sender map")


# ---- fast_mode, raytrace, loadMap import, explicit rays, single walls ----------------------------
def _room_scan(seed, n, origin, radius=3.0, noise=0.01, dtype=np.float32):
    """points on a sphere-ish room around `origin` (hits repeat from scan to scan -> occupied voxels build up)"""
    rng = np.random.default_rng(seed)
    d = rng.normal(size=(n, 3))
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    r = radius * (1.0 + 0.25 * np.sin(3 * d[:, 0]) * np.cos(2 * d[:, 1]))
    return (np.asarray(origin) + d * r[:, None] + rng.normal(scale=noise, size=d.shape)).astype(dtype)


@pytest.mark.parametrize("cfg", [CFG_GTEST, CFG_ROS], ids=["gtest_cfg", "ros_cfg"])
def test_fast_mode_bit_exact(cfg):
    """Config::fast_mode (V:1466): castRayIntoGridFast V:577-602 against the oracle's restatement of VolumeRayIntersector.
    Phase 1 builds a map in normal mode (free-space leaves everywhere), phase 2 switches both to fast mode: rays that pass
    through occupied voxels, rays that stop short, clipped rays, NaN, zero-length rays, moving origin, two accumulates per
    integrate. Phase 3 is a map built in fast mode from the first scan (empty map: only end points, V:522)."""
    g, o = _pair(0.1, 6.0, cfg)
    origin = np.array([0.031, -0.012, 0.2])
    for k in range(5):
        pts = _room_scan(5, 6000, origin)
        for m in (g, o):
            m.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map before fast mode")
    assert popcount64(o.exportMap().active) > 1000   # there are occupied voxels for the fast rays to find
    for m in (g, o):
        m.setFastMode(True)
    for k in range(4):
        og = origin + np.array([0.21 * k, -0.13 * k, 0.05 * k])
        pts = _room_scan(40 + k, 5000, og, radius=[4.5, 2.0, 3.0, 7.5][k])   # through the wall / short / on it / clipped
        pts[11] = np.nan
        pts[12] = og.astype(np.float32)                                       # zero-length ray
        assert g.accumulateUpdate(pts, og, "s") == 0 and o.accumulateUpdate(pts, og, "s") == 0
        if k == 1:   # a second cloud in the same accumulation period
            pts2 = _room_scan(77, 2000, og, radius=5.0)
            g.accumulateUpdate(pts2, og, "s"); o.accumulateUpdate(pts2, og, "s")
        ug, uo = g.exportUpdateGrid("s"), o.exportUpdateGrid("s")
        assert len(uo) > 0
        assert_leafsets_equal(ug, uo, f"fast-mode update grid {k}")
        g.integrateUpdate(keep_change=True); o.integrateUpdate()
        assert_leafsets_equal(g.exportLastChange("s"), o.exportLastChange("s"), f"fast-mode change grid {k}")
        assert_leafsets_equal(g.exportMap(), o.exportMap(), f"fast-mode map {k}")
    sg, so = g.stats(), o.stats()
    for key in ("rays", "nan_skipped", "clipped", "visits", "voxel_updates", "state_changes"):
        assert sg[key] == so[key], key
    # phase 3: fast mode from an empty map
    for m in (g, o):
        m.resetMap()
    for k in range(5):
        pts = _room_scan(5 if k < 4 else 6, 6000, origin, radius=3.0 if k < 4 else 5.0)
        g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
        assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), f"fast-only update grid {k}")
        g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "fast-only map")
    # a reduced (level 2) update is refused in fast mode (it describes castRayIntoGrid scans)
    from vdb_mapping_b200.mapping import VdbmError
    g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
    with pytest.raises(VdbmError):
        g.createUpdate("s", 2)
    # and back to normal mode, on top of the fast-mode accumulate that is still in the update grid
    for m in (g, o):
        m.setFastMode(False)
    g.accumulateUpdate(pts, origin, "s"); o.accumulateUpdate(pts, origin, "s")
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "normal-mode accumulate on top of a fast-mode one")


def test_fast_mode_large_coordinates_and_many_nodes():
    """Rays far from the world origin crossing many 128^3 and several 4096^3 blocks (0.02 m voxels, 90 m rays): the coarse node
    sets grow, negative coordinates, long spans."""
    g, o = _pair(0.02, 100.0, CFG_ROS)
    origin = np.array([-351.237, 812.001, -3.3])
    rng = np.random.default_rng(3)
    d = rng.normal(size=(300, 3)); d[:, 2] *= 0.1
    d /= np.linalg.norm(d, axis=1, keepdims=True)
    pts = (origin + d * rng.uniform(20, 90, size=(300, 1))).astype(np.float32)
    for k in range(2):
        for m in (g, o):
            m.insertPointCloud(pts, origin, "s")
    for m in (g, o):
        m.setFastMode(True)
    pts2 = (origin + d * 95.0).astype(np.float32)   # through every end point of the first scans
    g.accumulateUpdate(pts2, origin, "s"); o.accumulateUpdate(pts2, origin, "s")
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "fast-mode update grid, long rays")
    g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map")
    assert g.stats()["visits"] == o.stats()["visits"]


def test_raytrace_matches_the_oracle():
    """Batch raytrace V:675-721: hits, misses, rays that start inside / outside the map's bounding box, axis-aligned and
    zero-component directions, unnormalised directions, an empty map. End points compared as fp64 bit patterns."""
    g, o = _pair(0.1, 6.0, CFG_GTEST)
    rng = np.random.default_rng(9)
    n = 4000
    org = rng.uniform(-1.0, 1.0, size=(n, 3))
    org[:200] = rng.uniform(-30, 30, size=(200, 3))           # outside the mapped room
    dirs = rng.normal(size=(n, 3)) * rng.uniform(0.1, 5.0, size=(n, 1))
    dirs[200:260, 1:] = 0.0                                    # axis-aligned
    dirs[260:300, 2] = 0.0
    lens = rng.uniform(0.5, 12.0, size=n)
    ok_g, e_g = g.raytrace(org, dirs, lens)
    ok_o, e_o = o.raytrace(org, dirs, lens)
    assert not ok_g.any() and not ok_o.any()
    assert np.array_equal(e_g.view(np.uint64), e_o.view(np.uint64)), "end points on an empty map"
    origin = np.array([0.031, -0.012, 0.2])
    for k in range(3):
        pts = _room_scan(5, 8000, origin)
        for m in (g, o):
            m.insertPointCloud(pts, origin, "s")
    ok_g, e_g = g.raytrace(org, dirs, lens)
    ok_o, e_o = o.raytrace(org, dirs, lens)
    assert ok_o.sum() > n // 2 and (~ok_o).sum() > 20
    assert np.array_equal(ok_g, ok_o), "success flags"
    assert np.array_equal(e_g.view(np.uint64), e_o.view(np.uint64)), "end points"
    # single-ray form used by the reference's other overload (V:643-664)
    ok1, e1 = g.raytrace(origin, [1.0, 0.2, 0.0], 10.0)
    ok2, e2 = o.raytrace(origin, [1.0, 0.2, 0.0], 10.0)
    assert ok1[0] and ok2[0] and np.array_equal(e1, e2)


def test_map_import_replaces_and_merges():
    """loadMap V:263-284 (device part, vdbm_map_import): a map exported from one handle and imported into another is the
    same map and keeps integrating identically; replace=False overwrites leaf for leaf."""
    from vdb_mapping_b200 import scans
    g, o = _pair(0.1, 4.0, CFG_ROS)
    for k in range(3):
        pts, origin = scans.small_scan(60 + k, n=3000, scale=2.5)
        for m in (g, o):
            m.insertPointCloud(pts, origin, "s")
    saved = g.exportMap()
    g2, _ = _pair(0.1, 4.0, CFG_ROS)
    pts0, origin0 = scans.small_scan(1, n=500, scale=1.0)
    g2.insertPointCloud(pts0, origin0 + 7.0, "s")        # content that must disappear
    g2.importMap(saved, replace=True)
    assert_leafsets_equal(g2.exportMap(), saved, "imported map")
    assert g2.mapLeafCount() == len(saved)
    pts, origin = scans.small_scan(70, n=3000, scale=2.5)
    for m in (g, g2, o):
        m.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g2.exportMap(), o.exportMap(), "scan on top of an imported map")
    assert_leafsets_equal(g2.exportMap(), g.exportMap(), "imported handle == original handle")
    # merge: leaves of `part` overwrite, the rest stays
    part = saved
    g3, _ = _pair(0.1, 4.0, CFG_ROS)
    g3.insertPointCloud(pts0, origin0 + 7.0, "s")
    far = g3.exportMap()
    g3.importMap(part, replace=False)
    merged = g3.exportMap()
    assert len(merged) == len(far) + len(part)
    # fast mode / raytrace see the imported topology (coarse node sets are rebuilt)
    ok, _e = g2.raytrace(origin, [1.0, 0.0, 0.0], 3.0)
    ok_o, _e2 = o.raytrace(origin, [1.0, 0.0, 0.0], 3.0)
    assert ok[0] == ok_o[0]


def test_cast_index_rays_and_single_walls():
    """castRayIntoGrid V:550-566 on explicit voxel pairs, addArtificialWall V:1217 / addArtificialPolygon V:1198 without the
    restoreMapIntegrity of addArtificialAreas."""
    g, o = _pair(0.1, 4.0, CFG_ROS)
    rng = np.random.default_rng(21)
    starts = rng.integers(-40, 40, size=(500, 3)).astype(np.int32)
    ends = starts + rng.integers(-60, 60, size=(500, 3)).astype(np.int32)
    ends[:20] = starts[:20]                         # zero-length: nothing is marked
    ends[20:40, 1:] = starts[20:40, 1:]             # axis-aligned
    g.castRaysIntoGrid("s", starts, ends)
    assert o.castRaysIntoGrid("s", starts, ends) == 0
    assert_leafsets_equal(g.exportUpdateGrid("s"), o.exportUpdateGrid("s"), "explicit rays")
    g.integrateUpdate(); o.integrateUpdate()
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "map after explicit rays")
    for m in (g, o):
        m.addArtificialWall([0.52, 0.5, 0.0, 1], [2.31, 0.77, 0.0, 1], -0.35, 0.55)
        m.addArtificialPolygon(np.array([[-1.0, -1.0, 0.2, 1], [-2.0, -1.5, 0.2, 1], [-1.2, -2.4, 0.2, 1]]), -0.2, 0.3)
        m.addArtificialWall([0.52, 0.5, 0.0, 1], [0.52, 0.5, 0.0, 1], -0.35, 0.55)   # start == end: nothing
    assert_leafsets_equal(g.exportArtificialAreaGrid(), o.exportArtificialAreaGrid(), "walls accumulate without a restore")
    from vdb_mapping_b200 import scans
    pts, origin = scans.small_scan(3, n=2000, scale=2.5)
    for m in (g, o):
        m.insertPointCloud(pts, origin, "s")
    assert_leafsets_equal(g.exportMap(), o.exportMap(), "updateMap re-activates the walls")
