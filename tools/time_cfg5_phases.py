"""Host wall time of every call of the remote-mapping pipeline (bench.py --workload cfg5), averaged over the timed scans."""
import os, sys, time
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np
from vdb_mapping_b200 import scans
from vdb_mapping_b200.mapping import OccupancyVDBMapping

c = scans.CONFIGS[2]
def mk():
    m = OccupancyVDBMapping(c.resolution)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    m.addInputSource("s", c.max_range)
    return m
snd, rcv = mk(), mk()
W, K = 5, 30
clouds = [scans.make_scan(2, k) for k in range(W + K)]
T = {}
def timed(name, fn, k):
    t0 = time.perf_counter(); r = fn(); dt = time.perf_counter() - t0
    if k >= W: T.setdefault(name, []).append(dt * 1e3)
    return r
for k, (pts, origin) in enumerate(clouds):
    timed("accumulate", lambda: snd.accumulateUpdate(pts, origin, "s"), k)
    red, o = timed("createUpdate2", lambda: snd.createUpdate("s", 2), k)
    timed("integrate", lambda: snd.integrateUpdate(keep_change=False), k)
    timed("applyUpdate2", lambda: rcv.applyUpdate(2, red, origin=o), k)
    if k % 10 == 9:
        lo = np.floor((origin - [10, 10, 3]) / c.resolution).astype(np.int32); hi = np.floor((origin + [10, 10, 3]) / c.resolution).astype(np.int32)
        sp = timed("section_sparse", lambda: snd.getMapSectionUpdateGrid(lo, hi, full=False), k)
        fu = timed("section_full", lambda: snd.getMapSectionGrid(lo, hi, full=True), k)
        print("section leaves", len(sp), len(fu))
print("reduced leaves", len(red), "voxels", int(np.unpackbits(red.active.view(np.uint8)).sum()))
print("createUpdate2 per call:", np.round(T["createUpdate2"], 1).tolist())
for n, v in T.items():
    print(f"{n:16s} mean {np.mean(v):8.3f} ms  min {np.min(v):8.3f}  n {len(v)}")
