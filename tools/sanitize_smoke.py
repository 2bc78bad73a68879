"""Small end-to-end run for compute-sanitizer: scan path (sync + queued), fast_mode, raytrace, sections, the host
mirror stream, a source on its own handle (vdbm_integrate_from), a 2-shard group on one GPU.
    compute-sanitizer --tool memcheck python tools/sanitize_smoke.py"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdb_mapping_b200 import scans
from vdb_mapping_b200.mapping import OccupancyVDBMapping, OccupancyVDBMappingGroup

m = OccupancyVDBMapping(0.1)
m.setConfig(4.0, 0.7, 0.4, 0.12, 0.97)
m.addInputSource("s", 4.0)
for k in range(3):
    pts, origin = scans.small_scan(10 + k, n=20000, scale=2.5)
    m.insertPointCloud(pts, origin, "s")
for k in range(4):   # queued / overlapped path (many short rays)
    pts, origin = scans.small_scan(20 + k, n=200000, scale=2.5)
    m.insertPointCloudAsync(pts, origin, "s")
m.flush()
print("pipeline", m.pipelineCounts())
m.setFastMode(True)
pts, origin = scans.small_scan(30, n=20000, scale=3.0)
m.insertPointCloud(pts, origin, "s")
ok, e = m.raytrace(np.random.default_rng(0).uniform(-1, 1, (512, 3)), np.random.default_rng(1).normal(size=(512, 3)), 6.0)
m.setFastMode(False)
lo, hi = np.array([-20, -20, -10], np.int32), np.array([20, 20, 10], np.int32)
m.getMapSectionUpdateGrid(lo, hi); m.getMapSectionGrid(lo, hi, full=True)
full = m.exportMap()
m2 = OccupancyVDBMapping(0.1)
m2.setConfig(4.0, 0.7, 0.4, 0.12, 0.97); m2.addInputSource("s", 4.0)
m2.importMap(full)
assert m2.mapChecksum() == m.mapChecksum()
# host mirror stream (several chunks, an aborted transfer, the rest on the next call) and a source on its own raycast handle
got = []
m.mirrorMap(lambda idx, o, v, a: got.append(len(idx)) or 0, chunk_leaves=64)
pts, origin = scans.small_scan(31, n=20000, scale=2.5)
m.insertPointCloud(pts, origin, "s")
try:
    m.mirrorMap(lambda idx, o, v, a: 1, chunk_leaves=32)
except Exception:
    pass
n_rest = m.mirrorMap(lambda idx, o, v, a: 0, chunk_leaves=50)
h = OccupancyVDBMapping(0.1, map_capacity_leaves=1)
h.setConfig(4.0, 0.7, 0.4, 0.12, 0.97); h.addInputSource("s", 4.0)
pts, origin = scans.small_scan(32, n=20000, scale=2.5)
h.accumulateUpdate(pts, origin, "s")
m.integrateFrom(h, "s", keep_change=True)
print("mirror chunks", len(got), "rest", n_rest, "change leaves on the holder", len(h.exportLastChange("s")))
h.close()
g = OccupancyVDBMappingGroup(0.1, [0, 0])
g.setConfig(4.0, 0.7, 0.4, 0.12, 0.97); g.addInputSource("s", 4.0)
for k in range(3):
    pts, origin = scans.small_scan(40 + k, n=20000, scale=2.5)
    g.insertPointCloud(pts, origin, "s")
print("group leaves", g.stats()["map_leaves"], "hits", int(ok.sum()))
g.close(); m.close(); m2.close()
print("sanitize_smoke done")
