"""Single-GPU timing of the fused bin-and-send kernel in loop-back mode (all 'peers' = own inbox): separates the
kernel's own cost from the NVLink write path. Run on the GPU box."""
import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdb_mapping_b200 import scans
from vdb_mapping_b200.mapping import OccupancyVDBMapping
for cfg, nr in ((2, 2), (2, 8), (4, 8)):
    c = scans.CONFIGS[cfg]
    m = OccupancyVDBMapping(c.resolution); m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max); m.addInputSource("s", c.max_range)
    m.exchangeCreate(0, nr, 1 << 23 if cfg == 4 else 1 << 20)
    m.exchangeConnect(None)
    for k in range(3):
        pts, o = scans.make_scan(cfg, k)
        m.accumulateUpdate(pts, o, "s")
        n = m.stats()["last_touched_leaves"]
        m.updatePush("s")
        t = m.exchangeTimings()
        print(f"cfg{cfg} ranks={nr} scan {k}: {n} leaves  push {t[0]:.3f} ms ({n*136/t[0]/1e6:.0f} GB/s)")
        m.integrateUpdate(False)
    m.close()
