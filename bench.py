#!/usr/bin/env python
"""bench.py — scan-integration throughput of the B200-native vdb_mapping hot path.

  python bench.py [--gpus N] [--steps K] [--warmup W] [--impl ours|reference] [--workload cfg2|cfg1|cfg3|cfg4]

A "step" is one insertPointCloud (raycast into the update grid + updateMap) of one synthetic scan of the
BASELINE.json config (default configs[1]: Ouster OS1-128-style 262,144-point scan, 0.05 m voxels, 30 m range,
moving-sensor sequence). With N > 1 every rank owns one LiDAR of a merged multi-LiDAR rig (N x 262,144 points per
step, weak scaling), rays stay rank-local, the map is sharded by leaf key and update leaves are exchanged with an
NCCL all-to-all. Rank 0 prints ONE JSON line.

Legs of the default (--impl ours) arm:
  value : scans already resident in HBM, vdbm_accumulate_device + vdbm_integrate per step
  e2e   : the public call (vdbm_insert) on PINNED HOST buffers: H2D of the cloud and D2H of the per-step counters
          inside the timed region
  cpu_baseline : the CPU oracle (restated reference, OpenVDB-free, 1 thread like the reference's serial path) on the
          first scans of the same sequence (rank 0, N == 1 only)
--impl reference times that CPU path alone, each step on a bounded sample of the scan.
"""
from __future__ import annotations

import argparse
import json
import os
import statistics
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

from vdb_mapping_b200 import scans  # noqa: E402

REF_SAMPLE_STRIDE = 1  # reference arm: FULL scans of the same sequence per step (same config as our arm)
REF_MAX_STEPS = 30     # ... and a cap on the number of timed steps instead (2.7 s per cfg2 scan on one core); printed when it bites
STRONG_WARMUP, STRONG_STEPS = 2, 6  # the strong_cfg4 block of every line: one 1M-point scan per step split over all ranks


def measured_peaks():
    p = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(p):
        try:
            d = json.load(open(p))
            return float(d["hbm_gbs"]), "measured (MEASURED_PEAKS.json)"
        except Exception:
            pass
    return 6650.0, "fallback (B200_PROFILING.md)"


class ClockSampler:
    """SM clock + throttle reasons DURING the timed region, sampled in-process through NVML every ~2 ms
    (spawning `nvidia-smi -lms` from every rank measurably perturbs multi-GPU runs)."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}

    def __init__(self, gpu_index: int):
        self.gpu, self.rows, self.stop_flag, self.t, self.h = gpu_index, [], threading.Event(), None, None
        try:
            import pynvml
            self.nv = pynvml
            pynvml.nvmlInit()
            idx = gpu_index
            vis = os.environ.get("CUDA_VISIBLE_DEVICES")
            if vis:
                try:
                    idx = int(vis.split(",")[gpu_index])
                except Exception:
                    idx = gpu_index
            self.h = pynvml.nvmlDeviceGetHandleByIndex(idx)
            self.max_sm = pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM)
        except Exception:
            self.h = None

    def _loop(self):
        nv = self.nv
        while not self.stop_flag.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(self.h, nv.NVML_CLOCK_SM)
                rs = nv.nvmlDeviceGetCurrentClocksEventReasons(self.h) if hasattr(nv, "nvmlDeviceGetCurrentClocksEventReasons") \
                    else nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h)
                pw = nv.nvmlDeviceGetPowerUsage(self.h) / 1000.0
                self.rows.append((sm, rs, pw))
            except Exception:
                pass
            self.stop_flag.wait(0.002)

    def start(self):
        if self.h is None:
            return
        self.rows, self.stop_flag = [], threading.Event()
        self.t = threading.Thread(target=self._loop, daemon=True)
        self.t.start()

    def stop(self):
        if self.h is None or self.t is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["nvml unavailable"]}
        self.stop_flag.set()
        self.t.join(timeout=1)
        sm = [r[0] for r in self.rows]
        reasons = set()
        for _, rs, _ in self.rows:
            for bit, name in self.REASONS.items():
                if rs & bit:
                    reasons.add(name)
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": self.max_sm,
                "power_w_max": max((r[2] for r in self.rows), default=None), "samples": len(sm), "reasons": sorted(reasons)}


def measured_traffic(kernel: str):
    """DRAM bytes per launch of `kernel` from the newest committed ncu capture (profiles/traffic_*.json), or None.
    bench.py never runs under the profiler; the capture is taken separately with tools/profile_gpu.sh."""
    pdir = os.path.join(ROOT, "profiles")
    try:
        files = sorted(f for f in os.listdir(pdir) if f.startswith("traffic_") and f.endswith(".json"))
        if not files:
            return None, None
        d = json.load(open(os.path.join(pdir, files[-1])))
        for k, v in d.items():
            if k.startswith(kernel):
                return float(v["dram_bytes_per_launch"]), files[-1]
    except Exception:
        pass
    return None, None


def workload_cfg(name: str) -> int:
    """cfg5 (BASELINE configs[4]) = cfg2's scan sequence as REMOTE MAPPING: the sender raycasts, creates the
    level-2 (reduced) update, integrates; a second map applies that update (re-raycast + updateMap); every 10th scan a
    sparse and a full-leaf map section of a 20x20x6 m box around the sensor is extracted to the host."""
    return {"cfg1": 1, "cfg2": 2, "cfg3": 3, "cfg4": 4, "cfg5": 2}[name]


def section_box(origin, resolution: float):
    """Index bounding box (inclusive) of the 20 x 20 x 6 m box centred on the sensor."""
    half = np.array([10.0, 10.0, 3.0])
    lo = np.floor((np.asarray(origin) - half) / resolution).astype(np.int32)
    hi = np.floor((np.asarray(origin) + half) / resolution).astype(np.int32)
    return lo, hi


def workload_name(args, c) -> str:
    return c.name if args.workload != "cfg5" else "cfg5_remote_" + c.name.split("_", 1)[1] + "_level2+sections"


def alg_bytes(n_pts: int, leaves: int, change: bool = False) -> int:
    """SURVEY.md 8(d) / BASELINE.md 4: 16 B per point + per touched leaf: update masks written+read (256 B),
    map leaf values+mask read and written (4224 B) [+128 B change masks]."""
    return 16 * n_pts + (4608 if change else 4480) * leaves


# ------------------------------------------------------------------------------------------------------------------
def run_reference(args, rank, world):
    """CPU arm: the reference's own algorithm (oracle port; the reference itself needs OpenVDB and cannot be built
    here) on the host cores. The reference's scan-integration path is serial per input source (VDBMapping.hpp:499,
    561,764 — no TBB call anywhere), so 'all the host threads it can use' is 1."""
    if rank != 0:
        return
    from oracle.oracle import OracleOccupancyVDBMapping
    cfg = workload_cfg(args.workload)
    c = scans.CONFIGS[cfg]
    m = OracleOccupancyVDBMapping(c.resolution)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    m.addInputSource("s", c.max_range)
    stride = REF_SAMPLE_STRIDE
    steps_asked = args.steps
    if args.workload != "cfg3":
        args.steps = min(args.steps, REF_MAX_STEPS)
        args.warmup = min(args.warmup, 3)
    clouds = []
    for k in range(args.warmup + args.steps):
        pts, origin = scans.make_scan(cfg, k)
        # a CONTIGUOUS 1/stride sector of the scan (azimuth-major order), rotating with k: neighbouring rays stay
        # neighbours, so the per-ray cost (voxel dedup ratio, cache behaviour) is that of the full scan
        per = pts.shape[0] // stride
        lo = (k % stride) * per
        pts = np.ascontiguousarray(pts[lo:lo + per])
        clouds.append((pts, origin))
    mixed = args.workload == "cfg5"
    remote = None
    if mixed:
        remote = OracleOccupancyVDBMapping(c.resolution)
        remote.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        remote.addInputSource("s", c.max_range)

    def step(k):
        if not mixed:
            m.insertPointCloud(*clouds[k], "s")
            return
        # remote mapping (BASELINE configs[4]): sender raycast -> reduced update -> sender updateMap; the remote map
        # re-raycasts the reduced update; every 10th scan a sparse and a full-leaf section around the sensor
        m.accumulateUpdate(*clouds[k], "s")
        red, o = m.createUpdate("s", 2)
        m.integrateUpdate()
        remote.applyUpdate("s", 2, red, o)
        if k % 10 == 9:
            lo, hi = section_box(clouds[k][1], c.resolution)
            m.getMapSectionUpdateGrid(lo, hi, full=False)
            m.getMapSectionGrid(lo, hi, full=True)

    for k in range(args.warmup):
        step(k)
    s0 = m.stats()
    t0 = time.perf_counter()
    rays = 0
    for k in range(args.warmup, args.warmup + args.steps):
        step(k)
        rays += clouds[k][0].shape[0]
    dt = time.perf_counter() - t0
    s1 = m.stats()
    value = rays / dt
    sample = (f"full scans ({c.n_points} rays per step), full pipeline" if stride == 1 else
              f"a contiguous 1/{stride} azimuth sector of each scan ({clouds[0][0].shape[0]} of {c.n_points} rays per step, sector rotates with the step), full pipeline")
    if args.steps != steps_asked:
        sample += f"; timed steps capped at {args.steps} of the {steps_asked} asked for (a scan takes seconds on one core)"
    line = {
        "impl": "reference", "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * dt / args.steps, "higher_is_better": True, "scaling": "weak",
        "vs_baseline": None, "dtype": "f64 DDA + f32 log-odds", "data": "synthetic",
        "config": {"workload": workload_name(args, c), "description": c.description, "resolution_m": c.resolution, "max_range_m": c.max_range,
                   "points_per_scan": c.n_points, "sample": sample},
        "voxel_updates_per_sec": (s1["voxel_updates"] - s0["voxel_updates"]) / dt,
        "visits_per_sec": (s1["visits"] - s0["visits"]) / dt,
        "cpu_baseline": {"value": value, "unit": "rays/s", "cores": 1, "kind": "port", "sample": sample,
                         "note": "restated reference (OpenVDB-free oracle, tree+accessor cost model); the reference path is "
                                 "single-threaded per input source; host has %d logical cores" % (os.cpu_count() or 0),
                         "multi_source": (cpu_multi_source_leg(cfg, 4, stride) if not mixed else None),
                         "optimistic_flat_hash": (cpu_flat_probe_leg(cfg, 4, stride) if not mixed else None)},
        "e2e": {"value": value, "unit": "rays/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
    }
    emit(line)


def cpu_multi_source_leg(cfg: int, n_scans: int, stride: int = 1):
    """SURVEY.md 8(d) 'S-source variant': the reference's only parallelism is one accumulation thread per input source plus
    one integrator (VDBMapping.hpp:1373,1416). The scan is split into S azimuth sectors fed as S sources: S threads raycast
    concurrently into their own update grids, then updateMap runs per source, serially, under the map lock."""
    from concurrent.futures import ThreadPoolExecutor
    from oracle.oracle import OracleOccupancyVDBMapping
    c = scans.CONFIGS[cfg]
    S = max(1, min(4, (os.cpu_count() or 2) - 1))
    m = OracleOccupancyVDBMapping(c.resolution)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    ids = [f"s{i}" for i in range(S)]
    for i in ids:
        m.addInputSource(i, c.max_range)
    clouds = []
    for k in range(n_scans):
        pts, origin = scans.make_scan(cfg, k)
        per = pts.shape[0] // stride
        pts = pts[(k % stride) * per:(k % stride) * per + per]
        parts = [np.ascontiguousarray(pts[i * len(pts) // S:(i + 1) * len(pts) // S]) for i in range(S)]
        clouds.append((parts, origin))
    rays = 0
    with ThreadPoolExecutor(max_workers=S) as pool:
        t0 = time.perf_counter()
        for parts, origin in clouds:
            list(pool.map(lambda a: m.accumulateUpdate(a[1], origin, a[0]), zip(ids, parts)))
            m.integrateUpdate()
            rays += sum(len(p) for p in parts)
        dt = time.perf_counter() - t0
    return {"value": rays / dt, "unit": "rays/s", "sources": S, "threads": S, "ms_per_scan": 1e3 * dt / n_scans,
            "sample": f"{n_scans} scan(s)" + (f", a 1/{stride} azimuth sector each" if stride > 1 else "") + f", split into {S} sources",
            "note": "S accumulation threads + serial updateMap per source; voxels seen by several sources are updated once per source"}


def cpu_flat_probe_leg(cfg: int, n_scans: int, stride: int = 1):
    """SURVEY.md 8(d) 'optimistic CPU': oracle/flat_probe.cpp — same arithmetic on a flat leaf hash (no OpenVDB-style tree,
    no virtual calls), single-threaded and with all host threads. A yardstick, not the reference."""
    from oracle.oracle import FlatProbe, OracleOccupancyVDBMapping
    c = scans.CONFIGS[cfg]
    o = OracleOccupancyVDBMapping(c.resolution)
    o.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    lo = o.logodds()
    clouds = []
    for k in range(n_scans):
        pts, origin = scans.make_scan(cfg, k)
        per = pts.shape[0] // stride
        clouds.append((np.ascontiguousarray(pts[(k % stride) * per:(k % stride) * per + per]), origin))
    out = {}
    for name, threads in (("single_thread", 1), ("all_threads", max(1, os.cpu_count() or 1))):
        f = FlatProbe(c.resolution, lo)
        f.insert(clouds[0][0], clouds[0][1], c.max_range, threads=threads)  # untimed: sizes the hash tables
        st0 = f.stats()
        t0 = time.perf_counter()
        rays = 0
        for pts, origin in clouds:
            f.insert(pts, origin, c.max_range, threads=threads)
            rays += len(pts)
        dt = time.perf_counter() - t0
        st = f.stats()
        out[name] = {"value": rays / dt, "unit": "rays/s", "threads": threads, "ms_per_scan": 1e3 * dt / n_scans,
                     "visits_per_sec": (st["visits"] - st0["visits"]) / dt,
                     "voxel_updates_per_sec": (st["voxel_updates"] - st0["voxel_updates"]) / dt}
    out["sample"] = f"{n_scans} scan(s)" + (f", a 1/{stride} azimuth sector each" if stride > 1 else "")
    out["note"] = "flat 8^3-leaf hash + last-leaf cache instead of the OpenVDB-style tree; threads = private update grids, OR-merge, leaf-parallel update"
    return out


def cpu_baseline_leg(cfg: int, n_scans: int):
    from oracle.oracle import OracleOccupancyVDBMapping
    c = scans.CONFIGS[cfg]
    m = OracleOccupancyVDBMapping(c.resolution)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    m.addInputSource("s", c.max_range)
    clouds = [scans.make_scan(cfg, k) for k in range(n_scans)]
    t0 = time.perf_counter()
    for pts, origin in clouds:
        m.insertPointCloud(pts, origin, "s")
    dt = time.perf_counter() - t0
    st = m.stats()
    out = {"value": n_scans * c.n_points / dt, "unit": "rays/s", "cores": 1, "kind": "port",
           "sample": f"first {n_scans} full scans of the same sequence ({dt:.1f} s of CPU work)",
           "ms_per_scan": 1e3 * dt / n_scans, "voxel_updates_per_sec": st["voxel_updates"] / dt,
           "visits_per_sec": st["visits"] / dt, "host_logical_cores": os.cpu_count(),
           "note": "restated reference (OpenVDB-free oracle); serial like the reference's per-source path"}
    del m
    out["multi_source"] = cpu_multi_source_leg(cfg, 2)
    out["optimistic_flat_hash"] = cpu_flat_probe_leg(cfg, 2)
    return out



# ------------------------------------------------------------------------------------------------------------------
def _gather(world, dist, obj):
    if world == 1:
        return [obj]
    out = [None] * world
    dist.all_gather_object(out, obj)
    return out


def reference_checksum(c, local_rank, stream, scans_per_step):
    """Rank 0, outside every timed region: the same scans integrated on ONE GPU (one accumulate per cloud of a step, then
    one integrate) -> (checksum, leaves). scans_per_step: list over steps of lists of (points, origin)."""
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    m = OccupancyVDBMapping(c.resolution, device=local_rank, stream=stream.cuda_stream)
    m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
    m.addInputSource("s", c.max_range)
    for clouds in scans_per_step:
        for pts, origin in clouds:
            m.accumulateUpdate(pts, origin, "s")
        m.integrateUpdate(keep_change=False)
    out = m.mapChecksum()
    m.close()
    return out


def strong_cfg4_block(args, rank, world, local_rank, stream, dist):
    """BASELINE configs[3] / north_star's multi-GPU target, on every line: ONE 1,048,576-point scan (0.02 m, 100 m, bounded
    hall) per step, split over all ranks of the run (strong scaling; at N = 1 this is the single-GPU time the efficiency is
    measured against). Rays are split by azimuth sector around the sensor, the map is owned by the same sectors (ShardPlan
    mode 1). Two sets of bounds are planned from a dry run of scan 0, because the exchange separates the two phases: the ray
    sectors equalise the raycast cost (voxel visits + weighted touched leaves), the ownership sectors equalise the owned
    leaves (dist.plan_rays_and_ownership; weights fitted to measured per-rank kernel times). Only the leaves
    near the sensor and along sector borders cross NVLink. Outside the timed region the order-independent checksum of all
    shards is compared with rank 0 integrating the same full scans on one GPU."""
    import torch
    from vdb_mapping_b200 import dist as vdist
    from vdb_mapping_b200.mapping import OccupancyVDBMapping
    cfg, c = 4, scans.CONFIGS[4]
    dev = torch.device("cuda", local_rank)
    n_steps = STRONG_WARMUP + STRONG_STEPS
    full = [scans.make_scan(cfg, k) for k in range(n_steps)]
    plan = ray_plan = None
    with torch.cuda.stream(stream):
        if world > 1:
            holder = [None]
            if rank == 0:
                # dry run: which leaves does scan 0 touch? (their azimuth histogram weights the sector bounds)
                tmp = OccupancyVDBMapping(c.resolution, device=local_rank, stream=stream.cuda_stream)
                tmp.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
                tmp.addInputSource("s", c.max_range)
                tmp.accumulateUpdate(full[0][0], full[0][1], "s")
                leaf_origins = np.array(tmp.exportUpdateGrid("s").origins)
                tmp.close()
                holder[0] = vdist.plan_rays_and_ownership(full[0][0], full[0][1], c.resolution, c.max_range, world, leaf_origins)
                del leaf_origins
            dist.broadcast_object_list(holder, src=0)
            ray_plan, plan = holder[0]
            mine = [(np.ascontiguousarray(p[vdist.sector_rays(p, o, ray_plan, rank)]), o) for p, o in full]
        else:
            mine = full
        resident = [torch.from_numpy(p).to(dev) for p, _ in mine]
        m = OccupancyVDBMapping(c.resolution, device=local_rank, stream=stream.cuda_stream)
        m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
        m.addInputSource("s", c.max_range)
        eng = vdist.CudaEngine(m, "s")
        if world > 1:
            m.setShardPlan(plan)
            vdist.connect_peers(m, dist, capacity_records_per_sender=1 << 21)
        ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        ph = {"raycast": [], "push": [], "wait": [], "import": [], "update": []}
        touched, owned, st0 = [], [], None
        for k in range(n_steps):
            if k == STRONG_WARMUP:
                if world > 1:
                    dist.barrier()
                torch.cuda.synchronize()
                st0 = m.stats()
                ev0.record(stream)
            eng.accumulate_raw(resident[k].data_ptr(), mine[k][0].shape[0], mine[k][1], on_device=True)
            sa = m.stats()
            if world > 1:
                vdist.push_pull_and_integrate(eng)
            else:
                eng.integrate()
            if k >= STRONG_WARMUP:
                sb = m.stats()
                ph["raycast"].append(sa["last_accumulate_ms"]); ph["update"].append(sb["last_integrate_ms"])
                touched.append(sa["last_touched_leaves"]); owned.append(sb["last_touched_leaves"])
                if world > 1:
                    a, b, cc = m.exchangeTimings()
                    ph["push"].append(a); ph["wait"].append(b); ph["import"].append(cc)
        ev1.record(stream)
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()
        st1 = m.stats()
        ms = ev0.elapsed_time(ev1)
        chk = m.mapChecksum()
        m.close()
        del resident
    mean = lambda x: float(sum(x) / max(1, len(x)))
    mine_stats = {"ms": ms, "rays": st1["rays"] - st0["rays"], "visits": st1["visits"] - st0["visits"],
                  "raycast_ms": mean(ph["raycast"]), "update_ms": mean(ph["update"]),
                  "updates": st1["voxel_updates"] - st0["voxel_updates"], "chk": chk,
                  "phases": {k: mean(v) for k, v in ph.items()}, "touched": mean(touched), "owned": mean(owned)}
    allr = _gather(world, dist, mine_stats)
    identical = None
    if world > 1 and rank == 0:
        ref = reference_checksum(c, local_rank, stream, [[f] for f in full])
        tot = (sum(r["chk"][0] for r in allr) & ((1 << 64) - 1), sum(r["chk"][1] for r in allr))
        identical = bool(tot == ref)
    if rank != 0:
        return None
    ms_max = max(r["ms"] for r in allr)
    rays = sum(r["rays"] for r in allr)
    phases = {k: {"max": max(r["phases"][k] for r in allr), "mean": mean([r["phases"][k] for r in allr])} for k in ph}
    return {
        "workload": c.name, "scaling": "strong", "points_per_step": int(rays / STRONG_STEPS), "steps": STRONG_STEPS, "warmup": STRONG_WARMUP,
        "ms_per_step": ms_max / STRONG_STEPS, "rays_per_sec": rays / (ms_max * 1e-3),
        "voxel_updates_per_sec": sum(r["updates"] for r in allr) / (ms_max * 1e-3),
        "visits_per_sec": sum(r["visits"] for r in allr) / (ms_max * 1e-3),
        "per_rank_ms": phases,
        "leaves_touched_per_rank": [r["touched"] for r in allr], "leaves_owned_per_rank": [r["owned"] for r in allr],
        "raycast_ms_per_rank": [r["raycast_ms"] for r in allr], "update_ms_per_rank": [r["update_ms"] for r in allr],
        "visits_per_step_per_rank": [r["visits"] / STRONG_STEPS for r in allr],
        "map_checksum": "%016x" % (sum(r["chk"][0] for r in allr) & ((1 << 64) - 1)), "map_leaves": sum(r["chk"][1] for r in allr),
        "sharded_map_identical": identical,
        "ownership": (None if plan is None else {"kind": "azimuth sectors (ShardPlan mode 1)", "center_leaf_xy": [plan.cx, plan.cy], "bounds": plan.bounds,
                                                 "ray_sector_bounds": ray_plan.bounds}),
        "note": "device time (CUDA events) of the timed steps, max over ranks; efficiency at N GPUs = ms_per_step(N=1) / (N * ms_per_step(N)) "
                "from the driver's own N = 1 line; sharded_map_identical: sum of the shard checksums == the same scans on one GPU (rank 0)",
    }

# ------------------------------------------------------------------------------------------------------------------
def shim_block(cfg: int, clouds, n_lazy: int = 30, n_eager: int = 14, warmup: int = 4):
    """The same scans through the reference-compatible C++ CLASS (include/vdb_mapping/OccupancyVDBMapping.hpp ->
    insertPointCloud), i.e. what an unmodified consumer of vdb_mapping sees: tools/bench_shim.cpp, a separate process with
    no Python in it, clouds in ordinary (pageable) std::vector memory, wall clock. Two mirror policies of getGrid():
    Lazy (the scan is queued; the host grid is brought up to date when getGrid() is called) and Eager (default of the
    shim: every insert ends with the modified leaves copied into the host grid, so an accessor taken earlier stays live,
    tests/mapping.cpp:13-29). Runs after the timed legs; never inside a timed region."""
    import struct
    import subprocess
    import tempfile
    exe = os.path.join(ROOT, "tools", "build", "bench_shim")
    so = os.path.join(ROOT, "vdb_mapping_b200", "libvdbm_b200.so")
    try:
        if not os.path.exists(exe) or os.path.getmtime(exe) < os.path.getmtime(so):
            os.makedirs(os.path.dirname(exe), exist_ok=True)
            subprocess.run(["g++", "-std=c++17", "-O2", "-I", os.path.join(ROOT, "include"), os.path.join(ROOT, "tools", "bench_shim.cpp"),
                            "-o", exe, so, "-Wl,-rpath," + os.path.join("$ORIGIN", "..", "..", "vdb_mapping_b200"), "-lpthread"], check=True, capture_output=True)
        c = scans.CONFIGS[cfg]
        n = min(len(clouds), max(n_lazy, n_eager))
        out = {"api": "vdb_mapping::OccupancyVDBMapping::insertPointCloud (C++ shim over the C ABI, separate process, pageable clouds, wall clock)",
               "modes": "lazy / eager = MirrorMode of getGrid(); sources4 = each scan split into 4 azimuth sectors fed as 4 input sources from 4 threads "
                        "(accumulateUpdate x 4, then integrateUpdate; compare cpu_baseline.multi_source), every source on its own raycast handle; "
                        "sources4_shared = the same with all sources taking turns on the map's handle"}
        with tempfile.TemporaryDirectory() as tmp:
            path = os.path.join(tmp, "scans.bin")
            with open(path, "wb") as f:
                f.write(struct.pack("7d", c.resolution, c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max, n))
                for pts, origin in clouds[:n]:
                    p16 = np.ones((pts.shape[0], 4), dtype=np.float32)
                    p16[:, :3] = pts[:, :3]
                    f.write(struct.pack("3d", *origin)); f.write(struct.pack("I", p16.shape[0])); f.write(p16.tobytes())
            for mode, count in (("lazy", n_lazy), ("eager", n_eager), ("sources4", n_lazy), ("sources4_shared", n_lazy)):
                p = subprocess.run([exe, path, mode, str(warmup), str(min(n, count))], capture_output=True, text=True, timeout=600)
                rows = [l for l in p.stdout.splitlines() if l.startswith("{")]
                if p.returncode != 0 or not rows:
                    out[mode] = {"error": (p.stderr or p.stdout)[-300:]}
                    continue
                d = json.loads(rows[-1])
                out[mode] = {k: d[k] for k in ("scans", "ms_per_scan", "rays_per_sec", "map_leaves", "host_grid_active_voxels", "mirror_leaves_per_scan") if k in d}
        return out
    except Exception as e:  # the block is informative; it must never cost the bench line
        return {"error": repr(e)[:300]}


# ------------------------------------------------------------------------------------------------------------------
def run_ours(args, rank, world, local_rank):
    import torch
    import torch.distributed as dist
    from vdb_mapping_b200 import dist as vdist
    from vdb_mapping_b200.mapping import OccupancyVDBMapping

    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    cfg = workload_cfg(args.workload)
    c = scans.CONFIGS[cfg]
    n_steps = args.warmup + args.steps
    strong = (args.scaling == "strong") and world > 1
    mixed = args.workload == "cfg5"
    pipelined = (world == 1) and not mixed and not args.no_pipeline

    # ---- synthetic input: scan k of the sequence; with N > 1 rank r owns LiDAR r of the merged rig ----
    clouds = []
    for k in range(n_steps):
        if strong:
            pts, origin = scans.make_scan(cfg, k)
            # azimuth sectors with equal voxel visits: a rank then touches ~1/N of the leaves, not all of them
            pts = np.ascontiguousarray(pts[vdist.sector_split(pts, origin, c.resolution, c.max_range, rank, world)])
        else:
            pts, origin = scans.make_scan(cfg, k, sensor=rank) if cfg == 2 else scans.make_scan(cfg, k + 1000 * rank)
        clouds.append((pts, origin))
    n_pts = clouds[0][0].shape[0]
    n_pts_k = [p.shape[0] for p, _ in clouds]
    pinned = [torch.from_numpy(p).pin_memory() for p, _ in clouds]
    resident = [t.to(dev, non_blocking=True) for t in pinned]
    torch.cuda.synchronize()

    stream = torch.cuda.Stream(device=dev)
    sampler = ClockSampler(local_rank)

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    def run_leg(e2e: bool):
        """W warm-up + K timed steps on a fresh map. Returns dict with device time (CUDA events on the launching
        stream), counters and per-kernel times."""
        with torch.cuda.stream(stream):
            m = OccupancyVDBMapping(c.resolution, device=local_rank, stream=stream.cuda_stream)
            m.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
            m.addInputSource("s", c.max_range)
            eng = vdist.CudaEngine(m, "s")
            remote = None
            if mixed:
                remote = OccupancyVDBMapping(c.resolution, device=local_rank, stream=stream.cuda_stream)
                remote.setConfig(c.max_range, c.prob_hit, c.prob_miss, c.prob_thres_min, c.prob_thres_max)
            p2p = world > 1 and args.exchange == "p2p" and not mixed
            if p2p:
                vdist.connect_peers(m, dist, capacity_records_per_sender=(1 << 22) if cfg == 4 else (1 << 19))
            ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            acc_ms, prep_ms, int_ms, leaves, upd_leaves = [], [], [], [], []
            d2h_extra = [0, 0, 0]  # cfg5: bytes read back (reduced updates + sections), sections extracted, reduced-update bytes
            sent = recv = 0
            st0 = None
            for k in range(n_steps):
                if k == args.warmup:
                    if pipelined:
                        m.flush()
                    barrier()
                    st0 = m.stats()
                    launches0 = st0["gpu_launches"]
                    if not e2e and not os.environ.get("VDBM_BENCH_NO_SAMPLER"):
                        sampler.start()
                    t_wall0 = time.perf_counter()
                    ev0.record(stream)
                origin = clouds[k][1]
                if pipelined:
                    # insertPointCloud as a pipeline stage: the scan is queued (upload on a copy stream, no host round trip
                    # between raycast and updateMap) and finished by the next call; statistics lag one scan behind
                    m.insertRawAsync(pinned[k].data_ptr() if e2e else resident[k].data_ptr(), n_pts_k[k], origin, "s", on_device=not e2e)
                    if k >= args.warmup + 1:
                        s = m.stats()
                        acc_ms.append(s["last_accumulate_ms"]); prep_ms.append(s["last_prep_ms"]); leaves.append(s["last_touched_leaves"])
                        int_ms.append(s["last_integrate_ms"])
                    if k == n_steps - 1:
                        m.flush()
                        s = m.stats()
                        acc_ms.append(s["last_accumulate_ms"]); prep_ms.append(s["last_prep_ms"]); leaves.append(s["last_touched_leaves"])
                        int_ms.append(s["last_integrate_ms"])
                    continue
                if e2e:
                    eng.accumulate_raw(pinned[k].data_ptr(), n_pts_k[k], origin, on_device=False)
                else:
                    eng.accumulate_raw(resident[k].data_ptr(), n_pts_k[k], origin, on_device=True)
                if e2e and k + 1 < n_steps:
                    # the next cloud crosses PCIe while this scan's exchange / update runs
                    m.prefetchRaw(pinned[k + 1].data_ptr(), n_pts_k[k + 1])
                if k >= args.warmup:
                    s = m.stats()
                    acc_ms.append(s["last_accumulate_ms"]); prep_ms.append(s["last_prep_ms"]); leaves.append(s["last_touched_leaves"])
                if mixed:
                    red, o = m.createUpdate("s", 2)          # device: end voxels -> leaf set; D2H of the reduced grid
                    m.integrateUpdate(keep_change=False)
                    if k >= args.warmup:
                        int_ms.append(m.stats()["last_integrate_ms"])
                        d2h_extra[0] += red.origins.nbytes + red.active.nbytes + red.valmask.nbytes
                        d2h_extra[2] += red.origins.nbytes + red.active.nbytes + red.valmask.nbytes
                    remote.applyUpdate(2, red, origin=o)     # H2D of the reduced grid, re-raycast, updateMap
                    if k % 10 == 9:
                        lo, hi = section_box(origin, c.resolution)
                        sp = m.getMapSectionUpdateGrid(lo, hi, full=False)
                        fu = m.getMapSectionGrid(lo, hi, full=True)
                        if k >= args.warmup:
                            d2h_extra[0] += sp.origins.nbytes + sp.active.nbytes + sp.valmask.nbytes + fu.origins.nbytes + fu.active.nbytes + fu.values.nbytes
                            d2h_extra[1] += 1
                        del sp, fu  # the consumer is done with them: the library's staging buffer is free for the next export
                    continue
                if p2p:
                    vdist.push_pull_and_integrate(eng)
                    a = b = 0
                else:
                    a, b = vdist.exchange_and_integrate(eng, world, dist if world > 1 else None)
                if k >= args.warmup:
                    sent += a; recv += b
                    s = m.stats()
                    int_ms.append(s["last_integrate_ms"]); upd_leaves.append(s["last_touched_leaves"])
            ev1.record(stream)
            barrier()
            t_wall = time.perf_counter() - t_wall0
            clocks = sampler.stop() if not e2e else None
            st1 = m.stats()
            ms = ev0.elapsed_time(ev1)
            out = {"ms": ms, "wall_ms": 1e3 * t_wall, "clocks": clocks,
                   "rays": st1["rays"] - st0["rays"], "voxel_updates": st1["voxel_updates"] - st0["voxel_updates"],
                   "visits": st1["visits"] - st0["visits"], "launches": st1["gpu_launches"] - launches0,
                   "acc_ms": acc_ms, "prep_ms": prep_ms, "int_ms": int_ms, "leaves": leaves, "map_leaves": st1["map_leaves"],
                   "sent": sent, "recv": recv, "d2h_extra": d2h_extra, "upd_leaves": upd_leaves}
            if pipelined and not e2e:
                # Per-kernel times for the roofline, OUTSIDE the timed region: in the pipeline the raycast of scan k+1 runs next
                # to updateMap of scan k, so CUDA events around one kernel also see the other. A few more scans through the
                # synchronous calls (same kernels, nothing overlapped) give clean launch durations.
                p_acc, p_prep, p_int, p_leaves = [], [], [], []
                # (8 scans spread evenly over the timed range, so that their touched-leaf count is the sequence's, not its tail's)
                for k in sorted(set(int(round(x)) for x in np.linspace(args.warmup, n_steps - 1, 8))):
                    eng.accumulate_raw(resident[k].data_ptr(), n_pts_k[k], clouds[k][1], on_device=True)
                    s = m.stats()
                    p_acc.append(s["last_accumulate_ms"]); p_prep.append(s["last_prep_ms"]); p_leaves.append(s["last_touched_leaves"])
                    eng.integrate()
                    p_int.append(m.stats()["last_integrate_ms"])
                out["probe"] = {"acc_ms": p_acc, "prep_ms": p_prep, "int_ms": p_int, "leaves": p_leaves}
            if world > 1 and not mixed and not e2e:
                out["checksum"] = m.mapChecksum()  # parity witness of the sharded map, compared below (outside the timed region)
            if mixed:
                # outside the timed region: the remote map must equal the sender's (level-2 updates are lossless)
                a, b = m.exportMap(), remote.exportMap()
                out["remote_identical"] = bool(len(a) == len(b) and np.array_equal(a.origins, b.origins) and np.array_equal(a.active, b.active)
                                               and np.array_equal(a.values.view(np.uint32), b.values.view(np.uint32)))
                out["remote_launches"] = 0
                del a, b
                remote.close()
            m.close()
            return out

    if os.environ.get("VDBM_BENCH_E2E_FIRST"):
        res_e = run_leg(e2e=True)
        res_v = run_leg(e2e=False)
    else:
        res_v = run_leg(e2e=False)
        res_e = run_leg(e2e=True)
    if world > 1 and vdist.PROFILE:
        mean_ = lambda x: float(sum(x) / max(1, len(x)))
        log = np.array(vdist.PROFILE_LOG[-args.steps:]) if vdist.PROFILE_LOG else np.zeros((1, 6))
        print(f"[rank {rank}] value leg: step {res_v['ms'] / args.steps:.3f} ms  acc {mean_(res_v['acc_ms']):.3f}  int {mean_(res_v['int_ms']):.3f} | "
              f"e2e leg: step {res_e['ms'] / args.steps:.3f} ms acc {mean_(res_e['acc_ms']):.3f} int {mean_(res_e['int_ms']):.3f} phases {log.mean(axis=0).round(3).tolist()} "
              f"leaves {mean_(res_v['leaves']):.0f} map_leaves {res_v['map_leaves']}", file=sys.stderr, flush=True)

    # ---- parity witness of the sharded map of the value leg (outside every timed region) ----
    weak_identical = None
    if world > 1 and not mixed and "checksum" in res_v:
        allc = _gather(world, dist, res_v["checksum"])
        if rank == 0:
            if strong:
                per_step = [[scans.make_scan(cfg, k)] for k in range(n_steps)]
            else:
                per_step = [[(scans.make_scan(cfg, k, sensor=r) if cfg == 2 else scans.make_scan(cfg, k + 1000 * r)) for r in range(world)]
                            for k in range(n_steps)]
            ref = reference_checksum(c, local_rank, stream, per_step)
            weak_identical = bool((sum(x[0] for x in allc) & ((1 << 64) - 1), sum(x[1] for x in allc)) == ref)
            del per_step
    # ---- north_star's multi-GPU target on every line of the default workload ----
    strong_block = None
    if args.workload == "cfg2" and not strong and not args.no_strong_block:
        torch.cuda.empty_cache()
        strong_block = strong_cfg4_block(args, rank, world, local_rank, stream, dist if world > 1 else None)

    # ---- max over ranks of the device time; totals over ranks ----
    def reduce(vals, op):
        if world == 1:
            return vals
        t = torch.tensor(vals, dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=op)
        return t.tolist()

    ms_v, ms_e = reduce([res_v["ms"], res_e["ms"]], dist.ReduceOp.MAX if world > 1 else None)
    tot = reduce([res_v["rays"], res_v["voxel_updates"], res_v["visits"], res_e["rays"], res_v["launches"], res_v["sent"]],
                 dist.ReduceOp.SUM if world > 1 else None)
    if rank != 0:
        return
    rays_v, upd_v, vis_v, rays_e, launches, sent = tot
    K = args.steps
    value = rays_v / (ms_v * 1e-3)
    e2e_value = rays_e / (ms_e * 1e-3)
    peak, peak_src = measured_peaks()
    mean = lambda x: float(sum(x) / max(1, len(x)))
    L_step = mean(res_v["leaves"])  # touched leaves per scan of the TIMED steps
    probe = res_v.get("probe")
    if probe:
        # kernel durations (and the leaves those launches processed) come from the synchronous probe scans, see run_leg
        L = mean(probe["leaves"])
        t_prep, t_acc, t_int = mean(probe["prep_ms"]), mean(probe["acc_ms"]), mean(probe["int_ms"])
    else:
        L = L_step
        t_prep, t_acc, t_int = mean(res_v["prep_ms"]), mean(res_v["acc_ms"]), mean(res_v["int_ms"])
    t_dda = t_acc - t_prep
    b_alg = alg_bytes(n_pts, int(L))  # cfg5: roofline figures describe the SENDER's scan-step kernels only
    b_alg_step = alg_bytes(n_pts, int(L_step))
    t_kernels = t_acc + t_int
    # K2 alone: update masks read (128 B) + map leaf values+mask read and written (4224 B) per leaf THIS RANK integrates
    # (with N > 1 that is the rank's owned share after the exchange, not the leaves its own rays touched)
    L_upd = mean(res_v["upd_leaves"]) if res_v["upd_leaves"] else L
    upd_bytes = 4352 * L_upd
    if world > 1:
        # rank 0's own kernels: it raycasts its rays (16 B/point in, 128 B per touched leaf out) and integrates the leaves it OWNS
        b_alg = 16 * n_pts + 128 * L + upd_bytes
    tr_dda, tr_src = measured_traffic("raycast_dda_kernel")
    tr_upd, _ = measured_traffic("apply_update_kernel")
    traffic = (tr_dda + tr_upd) if (tr_dda is not None and tr_upd is not None and cfg == 2) else None
    roofline = {
        "bound": "hbm", "kernel": "scan step = prep_rays + raycast_dda + apply_update (SURVEY 8d definition)",
        "achieved": b_alg / (t_kernels * 1e-3) / 1e9, "peak": peak, "peak_source": peak_src, "unit": "GB/s",
        "frac": b_alg / (t_kernels * 1e-3) / 1e9 / peak, "traffic": traffic,
        "traffic_source": (f"profiles/{tr_src}: dram bytes of raycast_dda_kernel + apply_update_kernel per scan (ncu --set full, cfg2)" if traffic else None),
        "algorithmic_bytes_per_step": b_alg, "touched_leaves_per_step": L, "kernel_ms_per_step": t_kernels,
        "frac_of_step_time": b_alg_step / (ms_v / K * 1e-3) / 1e9 / peak,
        "algorithmic_bytes_per_timed_step": b_alg_step, "touched_leaves_per_timed_step": L_step,
        "kernel_times_from": ("%d extra scans through the synchronous calls after the timed region (the pipeline overlaps the raycast of scan k+1 "
                              "with updateMap of scan k, so events inside it would see both kernels); achieved / frac / algorithmic_bytes_per_step / "
                              "by_kernel describe those launches, frac_of_step_time = algorithmic bytes of the timed steps / ms_per_step"
                              % len(probe["acc_ms"])) if probe else "CUDA events of the timed steps",
        "by_kernel": {
            "prep_rays_kernel": {"ms": t_prep, "alg_bytes": 64 * n_pts, "achieved_gbs": 64 * n_pts / (t_prep * 1e-3) / 1e9 if t_prep else None},
            "raycast_dda_kernel": {"ms": t_dda, "alg_bytes": 48 * n_pts + 128 * L,
                                   "achieved_gbs": (48 * n_pts + 128 * L) / (t_dda * 1e-3) / 1e9 if t_dda else None,
                                   "visits_per_sec": (res_v["visits"] / K) / (t_dda * 1e-3) if t_dda else None,
                                   "note": "not HBM-bound: bounded by L2 atomic (RED) throughput and instruction issue"},
            "apply_update_kernel": {"ms": t_int, "alg_bytes": upd_bytes, "leaves": L_upd, "achieved_gbs": upd_bytes / (t_int * 1e-3) / 1e9 if t_int else None,
                                    "frac": upd_bytes / (t_int * 1e-3) / 1e9 / peak if t_int else None,
                                    "note": "rank 0's numbers" if world > 1 else None},
        },
    }
    line = {
        "metric": "rays_per_sec", "value": value, "unit": "rays/s", "n_gpus": world, "steps": K, "warmup": args.warmup,
        "ms_per_step": ms_v / K, "higher_is_better": True, "scaling": "strong" if strong else "weak", "vs_baseline": None,
        "dtype": "f64 DDA + f32 log-odds", "data": "synthetic",
        "config": {"workload": workload_name(args, c), "description": c.description, "resolution_m": c.resolution, "max_range_m": c.max_range,
                   "points_per_scan_per_gpu": n_pts, "points_per_step_total": int(rays_v / K),
                   "sequence": "moving sensor, scan k of the sequence per step, fresh map at step 0",
                   "parallelism": ("1 GPU" if world == 1 else (f"{world} GPUs: " + ("one scan split across ranks" if strong else "one LiDAR per GPU of a merged rig") +
                                                               ", map sharded by leaf key, " + ("peer-memory (NVLink) push" if args.exchange == "p2p" else "NCCL all-to-all") + " of update leaves")),
                   "l2": "no explicit flush: per-step working set (map leaves %.2f GB + update grid) exceeds the 126 MB L2 and every step has new input" % (res_v["map_leaves"] * 2112 / 1e9)},
        "voxel_updates_per_sec": upd_v / (ms_v * 1e-3), "visits_per_sec": vis_v / (ms_v * 1e-3),
        "e2e": {"value": e2e_value, "unit": "rays/s", "ms_per_step": ms_e / K, "h2d_bytes_per_step": 16 * n_pts,
                "d2h_bytes_per_step": 164 if not pipelined else 108,
                "api": ("vdbm_insert_async(host pinned cloud) per scan + vdbm_flush at the end == insertPointCloud as a pipeline stage "
                        "(upload overlaps the previous scan, one host synchronisation per scan)") if pipelined
                       else "vdbm_accumulate(host pinned cloud) + vdbm_integrate == insertPointCloud"},
        "gpu_launches": int(launches), "roofline": roofline, "clocks": res_v["clocks"],
        "pipelined": bool(pipelined),
        "wall_ms_per_step": res_v["wall_ms"] / K,
    }
    if mixed:
        line["remote_mapping"] = {
            "pipeline": "sender accumulate -> createUpdate(level 2) -> integrate; remote applyUpdate(level 2) on a second map of the same GPU; "
                        "every 10th scan getMapSectionUpdateGrid (sparse) + getMapSectionGrid (full) of a 20x20x6 m box, read back to the host",
            "remote_map_identical_to_sender": bool(res_v["remote_identical"] and res_e["remote_identical"]),
            "d2h_bytes_per_step_reduced_updates_and_sections": res_v["d2h_extra"][0] / K,
            "sections_per_step": res_v["d2h_extra"][1] / K,
            "note": "rays_per_sec counts the SENDER's rays; every ray is traversed twice (sender + remote); the roofline block describes the sender's scan-step kernels"}
        line["e2e"]["d2h_bytes_per_step"] = 164 + int(res_e["d2h_extra"][0] / K)
        line["e2e"]["h2d_bytes_per_step"] = 16 * n_pts + int(res_e["d2h_extra"][2] / K)  # cloud + the reduced update fed to the remote map
        line["e2e"]["api"] = "accumulate(host pinned cloud) + createUpdate(2) + integrate + remote applyUpdate(2) + periodic getMapSection*"
        line["config"]["parallelism"] = "1 GPU (sender and remote map)" if world == 1 else f"{world} independent sender/remote pairs, one per GPU"
        line["exchange"] = None
    line["strong_cfg4"] = strong_block
    if world > 1 and not mixed:
        line["sharded_map_identical"] = weak_identical
    if world > 1 and not mixed:
        line["exchange"] = {"kind": args.exchange, "records_sent_per_step": (sent / K) if args.exchange == "nccl" else None,
                            "bytes_sent_per_step": (136 * sent / K) if args.exchange == "nccl" else None,
                            "note": "p2p = one kernel bins the update leaves by owner and stores the 136-byte records into the owners' inboxes over NVLink (CUDA IPC), device-side epoch wait; nccl = count + record all-to-all"}
    if world == 1 and not mixed and not args.no_shim_block:
        line["shim"] = shim_block(cfg, clouds)
    if world == 1 and not args.no_cpu_baseline:
        line["cpu_baseline"] = cpu_baseline_leg(cfg, args.cpu_scans)
    else:
        line["cpu_baseline"] = None
    if world > 1 and vdist.PROFILE and not mixed:
        log = np.array(vdist.PROFILE_LOG[-K:])
        names = ["push_launch", "pull_wait_import_sync", "integrate", "dev_push", "dev_wait", "dev_import"] if args.exchange == "p2p" else \
            ["partition", "counts_a2a", "records_a2a_launch", "import", "integrate"]
        line["exchange"]["host_phase_ms_rank0"] = dict(zip(names, [float(x) for x in log.mean(axis=0)]))
        line["exchange"]["accumulate_ms_rank0"] = mean(res_e["acc_ms"])
    emit(line)


_REAL_STDOUT = None


def emit(line: dict):
    """Print the JSON line on the process's ORIGINAL stdout (fd 1 is redirected to stderr while the bench runs, so
    that library chatter such as 'NCCL version ...' cannot pollute the one-line contract)."""
    data = (json.dumps(line) + "\n").encode()
    if _REAL_STDOUT is not None:
        os.write(_REAL_STDOUT, data)
    else:
        sys.stdout.write(data.decode())
        sys.stdout.flush()


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=100)  # configs[1] is a 100-scan sequence
    ap.add_argument("--warmup", type=int, default=5)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--workload", default="cfg2", choices=["cfg1", "cfg2", "cfg3", "cfg4", "cfg5"])
    ap.add_argument("--scaling", default="weak", choices=["weak", "strong"])
    ap.add_argument("--exchange", default="p2p", choices=["p2p", "nccl"],
                    help="multi-GPU update-leaf exchange: fused peer-memory stores over NVLink (default) or NCCL all-to-all")
    ap.add_argument("--cpu-scans", type=int, default=4)  # ~11 s of single-thread CPU work for cfg2 (+ the multi-source and flat-hash legs)
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-strong-block", action="store_true", help="skip the strong_cfg4 block (1M-point scans split over all ranks)")
    ap.add_argument("--no-shim-block", action="store_true", help="skip the shim block (the same scans through the C++ class API)")
    ap.add_argument("--no-pipeline", action="store_true",
                    help="1 GPU: use the synchronous accumulate + integrate calls instead of vdbm_insert_async")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3

    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        run_reference(args, rank, world)
        return
    if world > 1:
        import torch
        import torch.distributed as dist
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        torch.cuda.set_device(local_rank)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local_rank))
    try:
        run_ours(args, rank, world, local_rank)
    finally:
        if world > 1:
            import torch.distributed as dist
            dist.destroy_process_group()


if __name__ == "__main__":
    main()
