"""One process, N GPUs through the C ABI (vdbm_group_*): wall time per insertPointCloud of pinned HOST clouds.
    python tools/bench_group.py --gpus 8 --workload cfg4 --steps 8 --warmup 3
Every shard uploads the whole cloud over its own PCIe link, casts its azimuth sector, exchanges, integrates; the call returns
when every shard has finished (host wall clock around the calls: the group synchronises inside). Prints one JSON line with
ms/scan, rays/s and the checksum comparison with a single handle (outside the timed region)."""
import argparse, ctypes as C, json, os, sys, time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from vdb_mapping_b200 import scans                      # noqa: E402
from vdb_mapping_b200 import _lib as L                  # noqa: E402
from vdb_mapping_b200.mapping import OccupancyVDBMapping, OccupancyVDBMappingGroup  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--workload", default="cfg4", choices=["cfg2", "cfg4"])
    ap.add_argument("--steps", type=int, default=8)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--no-check", action="store_true")
    a = ap.parse_args()
    cfg = int(a.workload[3:])
    c = scans.CONFIGS[cfg]
    lib = L.lib()
    n_steps = a.warmup + a.steps
    clouds = []
    for k in range(n_steps):
        pts, origin = scans.make_scan(cfg, k)
        p16 = np.ones((pts.shape[0], 4), dtype=np.float32); p16[:, :3] = pts[:, :3]
        ptr = lib.vdbm_host_alloc(C.c_size_t(p16.nbytes))
        C.memmove(ptr, p16.ctypes.data, p16.nbytes)
        clouds.append((ptr, p16.shape[0], origin, p16))
    grp = OccupancyVDBMappingGroup(c.resolution, list(range(a.gpus)), inbox_capacity_records=1 << 21)
    grp.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    grp.addInputSource("s", c.max_range)
    per = []
    for k in range(n_steps):
        ptr, n, origin, _ = clouds[k]
        t0 = time.perf_counter()
        grp.insertRaw(ptr, n, origin, "s")
        per.append(1e3 * (time.perf_counter() - t0))
    timed = per[a.warmup:]
    ms = float(np.mean(timed))
    st = grp.stats()
    shard_leaves = [grp.shard(i).mapLeafCount() for i in range(a.gpus)]
    shard_ms = [[round(grp.shard(i).stats()[key], 3) for key in ("last_accumulate_ms", "last_integrate_ms")] for i in range(a.gpus)]
    chk = grp.checksum()
    identical = None
    if not a.no_check:
        one = OccupancyVDBMapping(c.resolution, device=0)
        one.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        one.addInputSource("s", c.max_range)
        for ptr, n, origin, _ in clouds:
            one.insertRaw(ptr, n, origin, "s")
        identical = bool(one.mapChecksum() == chk)
        one.close()
    print(json.dumps({
        "tool": "bench_group", "api": "vdbm_group_insert (one process, worker thread per GPU, pinned host cloud, whole cloud uploaded to every GPU)",
        "workload": c.name, "n_gpus": a.gpus, "points_per_scan": clouds[0][1], "steps": a.steps, "warmup": a.warmup,
        "ms_per_scan_wall": ms, "ms_min": float(np.min(timed)), "ms_max": float(np.max(timed)), "rays_per_sec": clouds[0][1] / (ms * 1e-3),
        "first_scan_ms_incl_planning": per[0], "ms_all": [round(x, 2) for x in per], "map_leaves": st["map_leaves"], "leaves_per_shard": shard_leaves,
        "last_scan_kernel_ms_per_shard[raycast,update]": shard_ms, "sharded_map_identical_to_single_gpu": identical,
        "plan": None if a.gpus == 1 else {k: v.tolist() for k, v in zip(("center_leaf_xy", "ray_bounds", "ownership_bounds"), grp.plan())}}))
    grp.close()


if __name__ == "__main__":
    main()
