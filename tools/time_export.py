"""Times the host-mirror path (vdbm_map_export dirty_only=1) and sections after cfg2 scans (run on the GPU box)."""
import sys, time, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vdb_mapping_b200 import scans
from vdb_mapping_b200.mapping import OccupancyVDBMapping
import ctypes as C
c = scans.CONFIGS[2]
m = OccupancyVDBMapping(c.resolution); m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max); m.addInputSource("s", c.max_range)
L = m._L
for k in range(4):
    pts, o = scans.make_scan(2, k)
    t0 = time.perf_counter(); m.insertPointCloud(pts, o, "s"); t1 = time.perf_counter()
    out = C.c_void_p(); L.vdbm_map_export(m._h, 1, C.byref(out)); t2 = time.perf_counter()
    n = L.vdbm_leafset_size(out); L.vdbm_leafset_free(out); t3 = time.perf_counter()
    print(f"scan {k}: insert {1e3*(t1-t0):.2f} ms | dirty export {n} leaves ({n*2124/1e6:.0f} MB) {1e3*(t2-t1):.1f} ms | free {1e3*(t3-t2):.1f} ms")
t0 = time.perf_counter(); ls = m.getMapSectionUpdateGrid((-200, -200, -60), (200, 200, 60), False); t1 = time.perf_counter()
print(f"section 20x20x6 m sparse: {len(ls)} leaves {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); ls = m.getMapSectionGrid((-200, -200, -60), (200, 200, 60), True); t1 = time.perf_counter()
print(f"section 20x20x6 m full float: {len(ls)} leaves {1e3*(t1-t0):.2f} ms")
t0 = time.perf_counter(); m.accumulateUpdate(pts, o, "s"); ch = m.updateMap("s"); t1 = time.perf_counter()
print(f"accumulate + updateMap with change grid: {len(ch)} change leaves {1e3*(t1-t0):.2f} ms")
