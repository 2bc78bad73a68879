"""Multi-GPU scan integration: one process per GPU (torch.distributed), rays split across ranks, map sharded by
leaf key.

Per integrate, on every rank:
  1. raycast the rank's share of the cloud into its local device update grid          (K0 + K1)
  2. bin the touched update leaves by owner rank (vdbm_update_partition)               (136-byte leaf records)
  3. all-to-all the per-peer record counts, then the records themselves                (NCCL over NVLink / NVSwitch)
  4. OR the received records into the (now empty) local update grid (import)           -> only owned leaves remain
  5. updateMap on the owned shard                                                      (K2)

The exchange is the one real data-path collective of this path; everything else is rank-local. The host logic
(split sizes, exchange, import order) is backend-agnostic so that it can be exercised on CPU with gloo in tests
(tests/test_dist_gloo.py supplies a test double for the engine; the product engine is the CUDA library).
"""
from __future__ import annotations

import os
import time

import numpy as np

PROFILE = bool(os.environ.get("VDBM_DIST_PROFILE"))  # host-side phase times (ms) of every exchange, for tuning
PROFILE_LOG: list = []

RECORD_BYTES = 136  # uint64 key + 8 x uint64 active + 8 x uint64 value
RECORD_WORDS = 17

_MASK21 = (1 << 21) - 1
_BIAS = 1 << 20


def pack_leaf_key(origin) -> int:
    """Same 63-bit key as packLeafKey() in csrc/vdbm_device.cuh (leaf coords = origin >> 3, 21 bits per axis)."""
    x, y, z = (int(v) >> 3 for v in origin)
    return (((x + _BIAS) & _MASK21) << 42) | (((y + _BIAS) & _MASK21) << 21) | ((z + _BIAS) & _MASK21)


def _mix64(x: int) -> int:
    m = (1 << 64) - 1
    x ^= x >> 33
    x = (x * 0xff51afd7ed558ccd) & m
    x ^= x >> 33
    x = (x * 0xc4ceb9fe1a85ec53) & m
    x ^= x >> 33
    return x


def leaf_owner_py(origin, n_ranks: int) -> int:
    """Pure-Python twin of vdbm_leaf_owner (2x2x2 leaf bricks share an owner)."""
    key = pack_leaf_key(origin)
    brick = key & ~((1 << 42) | (1 << 21) | 1)
    return _mix64(brick ^ 0x9E3779B97F4A7C15) % n_ranks


# ---- sector ownership (ShardPlan mode 1): see ShardPlan in csrc/vdbm_device.cuh -------------------------------------
from dataclasses import dataclass, field


@dataclass
class ShardPlan:
    mode: int = 0                 # 0: hash ownership, 1: azimuth sectors around the leaf column (cx, cy)
    n_ranks: int = 1
    cx: int = 0
    cy: int = 0
    bounds: list = field(default_factory=list)  # ascending diamond angles in [0, 4), one per rank


def diamond_angle(dx, dy):
    """Monotone stand-in for atan2 on [0, 4): one IEEE division per point, identical on host, device and here."""
    dx = np.asarray(dx, dtype=np.float64)
    dy = np.asarray(dy, dtype=np.float64)
    with np.errstate(divide="ignore", invalid="ignore"):
        q1 = dy / (dx + dy)
        q2 = 1.0 - dx / (dy - dx)
        q3 = 2.0 - dy / (-dx - dy)
        q4 = 3.0 + dx / (dx - dy)
    a = np.where(dy >= 0.0, np.where(dx >= 0.0, q1, q2), np.where(dx < 0.0, q3, q4))
    return np.where((dx == 0.0) & (dy == 0.0), 0.0, a)


def sector_of(angles, bounds):
    """Index of the sector [bounds[r], bounds[r+1]) an angle falls into; below bounds[0] wraps to the last one."""
    b = np.asarray(bounds, dtype=np.float64)
    r = np.searchsorted(b, np.asarray(angles, dtype=np.float64), side="right") - 1
    return np.where(r < 0, len(b) - 1, r)


def leaf_owner_planned_py(origin, plan: ShardPlan) -> int:
    """Pure-Python twin of vdbm_leaf_owner_planned."""
    if plan.mode == 0:
        return leaf_owner_py(origin, plan.n_ranks)
    lx, ly = int(origin[0]) >> 3, int(origin[1]) >> 3
    return int(sector_of(diamond_angle(lx - plan.cx, ly - plan.cy), plan.bounds))


def ray_visits(points, origin, resolution: float, max_range: float):
    """voxel visits of every ray (1 + |d|_1 / res after range clipping; 0 for invalid points): the DDA's cost."""
    p = np.asarray(points, dtype=np.float64)[:, :3]
    d = p - np.asarray(origin, dtype=np.float64)[None, :]
    ln = np.sqrt((d * d).sum(axis=1))
    with np.errstate(invalid="ignore", divide="ignore"):
        scale = np.where((max_range > 0) & (ln > max_range), max_range / ln, 1.0)
    v = 1.0 + np.abs(d * scale[:, None]).sum(axis=1) / resolution
    return np.where(np.isfinite(v), v, 0.0)


def sector_histograms(points, origin, resolution: float, max_range: float, leaf_origins=None, n_bins: int = 8192):
    """Per azimuth bin (diamond angle, n_bins over [0, 4)) of one representative scan: the voxel visits of the rays that
    point into the bin, and the touched leaves whose column lies in it (leaf_origins: the scan's touched leaf origins, e.g.
    the update grid of a dry-run raycast; zeros without them). Returns (visits[n_bins], leaves[n_bins], (cx, cy) = leaf
    column of the sensor)."""
    o = np.asarray(origin, dtype=np.float64)
    cx, cy = int(np.floor(o[0] / resolution)) >> 3, int(np.floor(o[1] / resolution)) >> 3
    p = np.asarray(points, dtype=np.float64)[:, :3]
    ang = diamond_angle(p[:, 0] - o[0], p[:, 1] - o[1])
    ang = np.where(np.isfinite(ang), ang, 0.0)
    visits = np.bincount(np.minimum((ang * (n_bins / 4.0)).astype(np.int64), n_bins - 1),
                         weights=ray_visits(points, origin, resolution, max_range), minlength=n_bins).astype(np.float64)
    leaves = np.zeros(n_bins)
    if leaf_origins is not None and len(leaf_origins):
        lo = np.asarray(leaf_origins, dtype=np.int64)
        la = diamond_angle((lo[:, 0] >> 3) - cx, (lo[:, 1] >> 3) - cy)
        leaves = np.bincount(np.minimum((la * (n_bins / 4.0)).astype(np.int64), n_bins - 1), minlength=n_bins).astype(np.float64)
    return visits, leaves, (cx, cy)


# Cost weights fitted on B200 (8 ranks, 1M-point scans at 0.02 m, per-rank kernel times): raycast = 2.7 us per 1000 voxel
# visits + 0.18 us per touched leaf; updateMap = 0.71 us per owned leaf.
RAYCAST_LEAF_COST_IN_VISITS = 67.0
TOTAL_LEAF_COST_IN_VISITS = 330.0


def sector_costs(points, origin, resolution: float, max_range: float, leaf_origins=None,
                 leaf_cost_in_visits: float = TOTAL_LEAF_COST_IN_VISITS, n_bins: int = 8192):
    """visits + leaf_cost_in_visits x leaves per azimuth bin (see sector_histograms). Returns (cost[n_bins], (cx, cy))."""
    visits, leaves, center = sector_histograms(points, origin, resolution, max_range, leaf_origins, n_bins)
    return visits + leaf_cost_in_visits * leaves, center


def cut_sectors(cost, center, world: int) -> ShardPlan:
    """Sector bounds that give every rank the same share of `cost` (per azimuth bin). Deterministic."""
    n_bins = len(cost)
    c = np.cumsum(np.asarray(cost, dtype=np.float64) + 1e-9)
    total = float(c[-1])
    cuts = [0]
    for r in range(1, world):
        b = int(np.searchsorted(c, total * r / world)) + 1
        cuts.append(min(max(b, cuts[-1] + 1), n_bins - (world - r)))  # strictly ascending, room for the rest
    return ShardPlan(1, world, int(center[0]), int(center[1]), [4.0 * b / n_bins for b in cuts])


def refine_costs(cost, plan: ShardPlan, measured_ms):
    """Measured feedback for the planner: rank r needed measured_ms[r] for the sector it was given. The cost of the bins of
    every sector is rescaled so that the sector's total equals its measured time; cut_sectors() of the result moves the
    bounds towards equal TIME (the estimate's weights are only a first guess: the DDA's speed depends on how many distinct
    mask words it hits, the update's on the leaf count). One or two rounds converge."""
    n_bins = len(cost)
    cost = np.asarray(cost, dtype=np.float64)
    edges = [int(round(b * n_bins / 4.0)) for b in plan.bounds] + [n_bins]
    out = cost.copy()
    for r in range(plan.n_ranks):
        lo, hi = edges[r], edges[r + 1]
        tot = float(cost[lo:hi].sum())
        if tot > 0 and measured_ms[r] > 0:
            out[lo:hi] = cost[lo:hi] * (float(measured_ms[r]) / tot)
    return out


def plan_sectors(points, origin, resolution: float, max_range: float, world: int, leaf_origins=None,
                 leaf_cost_in_visits: float = TOTAL_LEAF_COST_IN_VISITS, n_bins: int = 8192) -> ShardPlan:
    """sector_costs + cut_sectors: ONE set of bounds for `world` ranks (rays and ownership) from one representative scan."""
    cost, center = sector_costs(points, origin, resolution, max_range, leaf_origins, leaf_cost_in_visits, n_bins)
    return cut_sectors(cost, center, world)


def plan_rays_and_ownership(points, origin, resolution: float, max_range: float, world: int, leaf_origins, n_bins: int = 8192):
    """TWO sets of sector bounds from one representative scan. The exchange sits between the ray casting and updateMap, so a
    step costs max over ranks (raycast) + max over ranks (update): both phases are balanced on their own -
      ray bounds : equal raycast cost (voxel visits + RAYCAST_LEAF_COST_IN_VISITS x touched leaves),
      ownership  : equal number of owned leaves (what updateMap streams).
    Where the two disagree the leaves between the two borders cross NVLink; everything else stays local.
    Returns (ray_plan, ownership_plan): split the rays with sector_rays(points, origin, ray_plan, rank), give the map
    ownership_plan (setShardPlan)."""
    visits, leaves, center = sector_histograms(points, origin, resolution, max_range, leaf_origins, n_bins)
    return (cut_sectors(visits + RAYCAST_LEAF_COST_IN_VISITS * leaves, center, world),
            cut_sectors(leaves + 1e-6 * visits, center, world))


def sector_rays(points, origin, plan: ShardPlan, rank: int):
    """Indices of the rays of `rank`: those whose direction around the SENSOR falls into the rank's sector."""
    p = np.asarray(points, dtype=np.float64)[:, :3]
    o = np.asarray(origin, dtype=np.float64)
    ang = diamond_angle(p[:, 0] - o[0], p[:, 1] - o[1])
    ang = np.where(np.isfinite(ang), ang, 0.0)   # NaN points: any rank will do (they are skipped), keep them on sector 0's owner
    return np.nonzero(sector_of(ang, plan.bounds) == rank)[0]


def split_points(n: int, rank: int, world: int):
    """Contiguous 1/world slice of a cloud (strong-scaling mode: one scan split across ranks)."""
    per = (n + world - 1) // world
    lo = min(n, rank * per)
    return lo, min(n, lo + per)


def sector_split(points, origin, resolution: float, max_range: float, rank: int, world: int):
    """Strong-scaling ray partition: rank r gets the rays of one AZIMUTH SECTOR around the sensor (so that it touches
    about 1/world of the map leaves instead of all of them), with sector widths chosen such that every rank gets
    about the same number of voxel visits. Deterministic on every rank. Returns the index array of this rank's rays."""
    p = np.asarray(points, dtype=np.float64)[:, :3]
    d = p - np.asarray(origin, dtype=np.float64)[None, :]
    az = np.arctan2(d[:, 1], d[:, 0])
    az = np.where(np.isfinite(az), az, 0.0)
    order = np.argsort(az, kind="stable")
    lo, hi = balanced_split(np.asarray(points)[order], origin, resolution, max_range, rank, world)
    return np.sort(order[lo:hi])


def balanced_split(points, origin, resolution: float, max_range: float, rank: int, world: int):
    """Contiguous slice of a cloud for `rank` such that every rank gets about the same number of voxel VISITS (the
    cost of a ray is 1 + |d|_1 voxel steps, known from its end point), not the same number of rays. Deterministic:
    every rank computes the same cut points from the same cloud. Returns (lo, hi)."""
    p = np.asarray(points, dtype=np.float64)[:, :3]
    d = p - np.asarray(origin, dtype=np.float64)[None, :]
    ln = np.sqrt((d * d).sum(axis=1))
    with np.errstate(invalid="ignore", divide="ignore"):
        scale = np.where((max_range > 0) & (ln > max_range), max_range / ln, 1.0)
    visits = 1.0 + np.abs(d * scale[:, None]).sum(axis=1) / resolution
    visits = np.where(np.isfinite(visits), visits, 0.0)
    c = np.cumsum(visits)
    total = c[-1] if len(c) else 0.0
    cuts = [0] + [int(np.searchsorted(c, total * r / world)) for r in range(1, world)] + [len(c)]
    return cuts[rank], cuts[rank + 1]


class CudaEngine:
    """The product engine: an OccupancyVDBMapping handle on this rank's GPU + torch tensors for the exchange."""

    def __init__(self, mapping, source_id: str):
        import torch
        self.torch = torch
        self.m = mapping
        self.src = source_id
        self.device = torch.device("cuda", torch.cuda.current_device())

    def accumulate(self, points, origin):
        return self.m.accumulateUpdate(points, origin, self.src)

    def accumulate_raw(self, ptr, n, origin, on_device):
        self.m.accumulateRaw(ptr, n, origin, self.src, on_device=on_device)

    def partition(self, world: int):
        """-> (counts int64[world] (host), send tensor int64 [(sum counts) * 17] on the device)"""
        torch = self.torch
        counts, ptr = self.m.partitionUpdate(self.src, world)
        total = int(counts.sum())
        if total == 0:
            return counts.astype(np.int64), torch.empty(0, dtype=torch.int64, device=self.device)

        class _Wrap:  # zero-copy view of the library-owned partition buffer
            __cuda_array_interface__ = {"shape": (total * RECORD_WORDS,), "typestr": "<i8", "data": (ptr, False), "version": 3}

        return counts.astype(np.int64), torch.as_tensor(_Wrap(), device=self.device)

    def new_recv(self, n_records: int):
        return self.torch.empty(n_records * RECORD_WORDS, dtype=self.torch.int64, device=self.device)

    def import_records(self, buf, n_records: int):
        # torch orders its NCCL collectives after/before work on torch's CURRENT stream only. If the library launches on
        # another stream (its own, or one that is not current), the received buffer must be complete before the import
        # kernel may read it: wait for the collective on the host.
        torch = self.torch
        if self.m.stream_handle == 0 or self.m.stream_handle != torch.cuda.current_stream().cuda_stream:
            torch.cuda.current_stream().synchronize()
        if n_records:
            self.m.importUpdateDevice(self.src, buf.data_ptr(), n_records)

    def integrate(self):
        self.m.integrateUpdate(keep_change=False)

    def counts_tensor(self, counts):
        return self.torch.as_tensor(np.asarray(counts, dtype=np.int64), device=self.device)


def connect_peers(mapping, dist, capacity_records_per_sender: int = 1 << 19):
    """One-time set-up of the peer-memory exchange: every rank allocates its inbox, the CUDA IPC handles are
    all-gathered through torch.distributed and every peer inbox is mapped. Returns True when the fused path is usable."""
    rank, world = dist.get_rank(), dist.get_world_size()
    mine = mapping.exchangeCreate(rank, world, capacity_records_per_sender)
    gathered = [None] * world
    dist.all_gather_object(gathered, mine)
    mapping.exchangeConnect(b"".join(gathered))
    dist.barrier()
    return True


def push_pull_and_integrate(engine):
    """Fused exchange: bin + NVLink stores into the owners' inboxes, device-side wait on the peers' epoch words, import,
    updateMap. No counts cross the host, no collective call on the data path."""
    t = [time.perf_counter()] if PROFILE else None
    engine.m.updatePush(engine.src)
    if PROFILE: t.append(time.perf_counter())
    if PROFILE:
        engine.m.updatePull(engine.src)
        t.append(time.perf_counter())
        engine.integrate()
    else:
        engine.m.updatePullIntegrate(engine.src)  # wait + import + updateMap behind the device-side guard: one host sync
    if PROFILE:
        t.append(time.perf_counter())
        # host: push (async launch), pull (wait + import + sync), integrate; device: push kernels, wait, import
        PROFILE_LOG.append([round(1e3 * (b - a), 3) for a, b in zip(t, t[1:])] + [round(x, 3) for x in engine.m.exchangeTimings()])


def exchange_and_integrate(engine, world: int, dist=None):
    """Steps 2-5 above. `engine` implements partition/new_recv/import_records/integrate/counts_tensor;
    `dist` is torch.distributed (or None / world == 1 for the single-rank short-cut).
    Returns (records sent to other ranks, records received from other ranks)."""
    if world == 1 or dist is None:
        engine.integrate()
        return 0, 0
    rank = dist.get_rank()
    t = [time.perf_counter()] if PROFILE else None
    counts, send = engine.partition(world)
    if PROFILE: t.append(time.perf_counter())
    send_counts = engine.counts_tensor(counts)
    recv_counts = send_counts.new_empty(world)
    dist.all_to_all_single(recv_counts, send_counts)
    recv_host = [int(v) for v in recv_counts.cpu().tolist()]
    if PROFILE: t.append(time.perf_counter())
    n_recv = sum(recv_host)
    recv = engine.new_recv(n_recv)
    dist.all_to_all_single(recv, send,
                           output_split_sizes=[c * RECORD_WORDS for c in recv_host],
                           input_split_sizes=[int(c) * RECORD_WORDS for c in counts])
    if PROFILE: t.append(time.perf_counter())
    engine.import_records(recv, n_recv)
    if PROFILE: t.append(time.perf_counter())
    engine.integrate()
    if PROFILE:
        t.append(time.perf_counter())
        PROFILE_LOG.append([round(1e3 * (b - a), 3) for a, b in zip(t, t[1:])])  # partition, counts, a2a launch, import, integrate
    return int(counts.sum() - counts[rank]), n_recv - recv_host[rank]


def sharded_insert(engine, points, origin, world: int, dist=None, mode: str = "own_cloud"):
    """insertPointCloud across `world` ranks.
    mode "own_cloud": `points` is this rank's own cloud (one sensor per GPU, weak scaling);
    mode "split":     `points` is the same full cloud on every rank, each rank raycasts its 1/world slice."""
    if mode == "split" and world > 1:
        lo, hi = split_points(points.shape[0], dist.get_rank(), world)
        points = points[lo:hi]
    engine.accumulate(points, origin)
    return exchange_and_integrate(engine, world, dist)
