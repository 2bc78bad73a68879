"""Host-side Python mirror of the reference's OccupancyVDBMapping interface for the scan-integration
path, over the C ABI of libvdbm_b200.so (include/vdbm_b200.h).

Method names, argument meaning and error behaviour follow
/root/reference/include/vdb_mapping/VDBMapping.hpp (insertPointCloud :399, accumulateUpdate :316,
integrateUpdate :375, updateMap :731, getGrid :799, getMapSection* :883-960, addInputSource :1352,
resetMap :174) and OccupancyVDBMapping.hpp (setConfig :59). Grids come back as LeafSet records
(leaf origin + OpenVDB-layout masks / values) instead of openvdb::Grid objects; the C++ shim in
include/vdb_mapping/ turns the same records into real grids.

There is no CPU path here: everything is executed by the CUDA kernels behind the ABI.
"""
from __future__ import annotations

import ctypes as C
from dataclasses import dataclass

import numpy as np

from . import _lib as L


@dataclass
class LeafSet:
    """Leaves sorted by origin (x, y, z). Layout as in OpenVDB: offset n=(x&7)<<6|(y&7)<<3|(z&7)."""
    origins: np.ndarray          # (n, 3) int32
    active: np.ndarray           # (n, 8) uint64
    valmask: np.ndarray | None   # (n, 8) uint64 (bool grids)
    values: np.ndarray | None    # (n, 512) float32 (float grids)

    def __len__(self):
        return int(self.origins.shape[0])


class _LeafsetOwner:
    """Frees a vdbm_leafset when the last numpy view of it is garbage-collected."""

    def __init__(self, lib, ptr):
        self._lib, self._ptr = lib, ptr

    def __del__(self):
        p, self._ptr = self._ptr, None
        if p:
            self._lib.vdbm_leafset_free(p)


class _View:
    """Array-interface wrapper: np.asarray(view) aliases library memory and keeps `owner` alive through .base."""

    def __init__(self, owner, address, shape, typestr):
        self.owner = owner
        self.__array_interface__ = {"data": (address, False), "shape": shape, "typestr": typestr, "version": 3}


class VdbmError(RuntimeError):
    def __init__(self, code, msg):
        super().__init__(f"vdbm status {code}: {msg}")
        self.code = code


def _pts16(points) -> np.ndarray:
    p = np.asarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be (n,3) or (n,4) float32")
    if p.shape[1] == 3:
        q = np.ones((p.shape[0], 4), dtype=np.float32)
        q[:, :3] = p
        p = q
    return np.ascontiguousarray(p)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OccupancyVDBMapping:
    """vdb_mapping::OccupancyVDBMapping on one B200 (device-resident map, CUDA kernels)."""

    def __init__(self, resolution: float, device: int = -1, replicate_probe_quirk: bool = True,
                 update_capacity_leaves: int = 0, map_capacity_leaves: int = 0, stream: int | None = None):
        self._L = L.lib()
        self.resolution = float(resolution)
        self.stream_handle = int(stream) if stream else 0  # 0 = library-owned stream
        p = L.VdbmParams(float(resolution), int(device), int(replicate_probe_quirk), int(update_capacity_leaves),
                         int(map_capacity_leaves), C.c_void_p(stream) if stream else None)
        h = C.c_void_p()
        rc = self._L.vdbm_create(C.byref(p), C.byref(h))
        if rc != L.VDBM_OK:
            raise VdbmError(rc, "vdbm_create failed (CUDA device required; there is no CPU fallback)")
        self._h = h

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.vdbm_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # ---- helpers -------------------------------------------------------------------------------------
    def _err(self) -> str:
        return (self._L.vdbm_last_error(self._h) or b"").decode()

    def _check(self, rc, allow=()):
        if rc != L.VDBM_OK and rc not in allow:
            raise VdbmError(rc, self._err())
        return rc

    def _take(self, ls_ptr, is_float: bool) -> LeafSet:
        """vdbm_leafset -> LeafSet. Small sets are copied and released at once. Sets of 16 MB and more are NOT copied: the
        numpy arrays are views of the library's (pinned) memory and keep the vdbm_leafset alive through their base object;
        it is released when the last array goes away."""
        n = int(self._L.vdbm_leafset_size(ls_ptr))
        big = n * (2048 if is_float else 128) >= (16 << 20)
        owner = _LeafsetOwner(self._L, ls_ptr) if big else None

        def arr(ptr, cols, dtype):
            if n == 0:
                return np.zeros((0, cols), dtype)
            if big:
                return np.asarray(_View(owner, C.cast(ptr, C.c_void_p).value, (n, cols), np.dtype(dtype).str))
            return np.ctypeslib.as_array(ptr, shape=(n, cols)).copy()

        try:
            origins = arr(self._L.vdbm_leafset_origins(ls_ptr), 3, np.int32)
            active = arr(self._L.vdbm_leafset_active(ls_ptr), 8, np.uint64)
            valmask = values = None
            if is_float:
                values = arr(self._L.vdbm_leafset_values(ls_ptr), 512, np.float32)
            else:
                valmask = arr(self._L.vdbm_leafset_valmask(ls_ptr), 8, np.uint64)
        finally:
            if not big:
                self._L.vdbm_leafset_free(ls_ptr)
        return LeafSet(origins, active, valmask, values)

    # ---- reference surface ---------------------------------------------------------------------------
    def setConfig(self, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max) -> int:
        """0 ok; 1 rejected by the base class (max_range < 0); 2 rejected by the occupancy checks."""
        rc = self._L.vdbm_set_config(self._h, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max)
        if rc == L.VDBM_ERR_BAD_CONFIG:
            return 1 if max_range < 0 else 2
        self._check(rc)
        return 0

    def logodds(self) -> np.ndarray:
        out = np.zeros(6, dtype=np.float32)
        self._check(self._L.vdbm_get_logodds(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return out

    def addInputSource(self, source_id: str, max_range: float, max_rate: float = 0.0):
        self._check(self._L.vdbm_source_add(self._h, source_id.encode(), float(max_range)))

    def resetMap(self):
        self._check(self._L.vdbm_reset(self._h))

    def accumulateUpdate(self, points, origin, source_id: str) -> int:
        """Returns 0 ok, 1 unknown source (no-op, like the reference's print+return), 2 not configured."""
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        rc = self._L.vdbm_accumulate(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o))
        if rc == L.VDBM_ERR_UNKNOWN_SOURCE:
            return 1
        if rc == L.VDBM_ERR_NOT_CONFIGURED:
            return 2
        self._check(rc)
        return 0

    def accumulateRaw(self, ptr: int, n: int, origin, source_id: str, on_device: bool = False, stride: int = 16):
        """accumulateUpdate on a raw pointer (pinned host or device memory); used by bench.py."""
        o = np.ascontiguousarray(origin, dtype=np.float64)
        f = self._L.vdbm_accumulate_device if on_device else self._L.vdbm_accumulate
        self._check(f(self._h, source_id.encode(), C.c_void_p(ptr), n, stride, _dp(o)))

    def raycastPointCloud(self, points, origin, raycast_range: float, source_id: str) -> bool:
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        rc = self._L.vdbm_raycast(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o), float(raycast_range))
        if rc == L.VDBM_ERR_NOT_CONFIGURED:
            return False
        self._check(rc)
        return True

    def integrateUpdate(self, keep_change: bool = True):
        self._check(self._L.vdbm_integrate(self._h, int(keep_change)))

    def integrateFrom(self, holder: "OccupancyVDBMapping", source_id: str, keep_change: bool = False):
        """vdbm_integrate_from: updateMap of ONE source whose update grid lives in `holder` (another handle of the same device
        and resolution, e.g. a raycast-only handle of that source), read in place; then a fresh update grid there."""
        self._check(self._L.vdbm_integrate_from(self._h, holder._h, source_id.encode(), int(keep_change)))

    def insertPointCloud(self, points, origin, source_id: str) -> bool:
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        rc = self._L.vdbm_insert(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o))
        self._check(rc, allow=(L.VDBM_ERR_UNKNOWN_SOURCE, L.VDBM_ERR_NOT_CONFIGURED))
        return True  # VDBMapping.hpp:405 always returns true

    def insertRaw(self, ptr: int, n: int, origin, source_id: str, stride: int = 16):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._check(self._L.vdbm_insert(self._h, source_id.encode(), C.c_void_p(ptr), n, stride, _dp(o)))

    def insertPointCloudAsync(self, points, origin, source_id: str) -> bool:
        """insertPointCloud as a pipeline stage (vdbm_insert_async): returns once the scan is queued; any later call on the
        map (or flush()) finishes it. Results are identical to insertPointCloud."""
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        rc = self._L.vdbm_insert_async(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o), 0)
        self._check(rc, allow=(L.VDBM_ERR_UNKNOWN_SOURCE, L.VDBM_ERR_NOT_CONFIGURED))
        return True

    def insertRawAsync(self, ptr: int, n: int, origin, source_id: str, on_device: bool = False, stride: int = 16):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._check(self._L.vdbm_insert_async(self._h, source_id.encode(), C.c_void_p(ptr), n, stride, _dp(o), int(on_device)))

    def prefetchRaw(self, ptr: int, n: int, stride: int = 16):
        """Upload the next (pinned) cloud on the copy stream now; accumulateRaw with the same pointer then skips the copy."""
        self._check(self._L.vdbm_prefetch(self._h, C.c_void_p(ptr), n, stride))

    def flush(self):
        self._check(self._L.vdbm_flush(self._h))

    def updateMap(self, source_id: str) -> LeafSet:
        """updateMap(source's accumulated update grid) -> change grid."""
        out = C.c_void_p()
        self._check(self._L.vdbm_update_map(self._h, source_id.encode(), C.byref(out)))
        return self._take(out, False)

    def exportUpdateGrid(self, source_id: str) -> LeafSet:
        out = C.c_void_p()
        self._check(self._L.vdbm_update_export(self._h, source_id.encode(), C.byref(out)))
        return self._take(out, False)

    def exportLastChange(self, source_id: str) -> LeafSet:
        out = C.c_void_p()
        self._check(self._L.vdbm_change_export(self._h, source_id.encode(), C.byref(out)))
        return self._take(out, False)

    def importUpdate(self, source_id: str, origins, active, valmask):
        o = np.ascontiguousarray(origins, dtype=np.int32)
        a = np.ascontiguousarray(active, dtype=np.uint64)
        v = np.ascontiguousarray(valmask, dtype=np.uint64)
        self._check(self._L.vdbm_update_import(self._h, source_id.encode(), o.shape[0],
                                               o.ctypes.data_as(C.POINTER(C.c_int32)),
                                               a.ctypes.data_as(C.POINTER(C.c_uint64)),
                                               v.ctypes.data_as(C.POINTER(C.c_uint64))))

    def exportMap(self, dirty_only: bool = False) -> LeafSet:
        """getGrid(): the map's leaves (values + active masks)."""
        out = C.c_void_p()
        self._check(self._L.vdbm_map_export(self._h, int(dirty_only), C.byref(out)))
        return self._take(out, True)

    def mirrorMap(self, sink, chunk_leaves: int = 0) -> int:
        """vdbm_map_mirror: streams the leaves modified since the previous mirror / dirty export to
        sink(leaf_index[n], origins[n,3], values[n,512], active[n,8]) chunk by chunk (the arrays are views that die with the
        call; a truthy return value aborts). Returns the number of leaves delivered."""
        raised = []

        def _sink(_user, n, idx, origins, values, active):
            n = int(n)
            try:
                r = sink(np.ctypeslib.as_array(idx, (n,)), np.ctypeslib.as_array(origins, (n, 3)),
                         np.ctypeslib.as_array(values, (n, 512)), np.ctypeslib.as_array(active, (n, 8)))
            except BaseException as e:  # an exception cannot cross the C frames: abort the transfer, re-raise afterwards
                raised.append(e)
                return 1
            return 1 if r else 0
        cb = L.MIRROR_SINK(_sink)
        done = C.c_uint64(0)
        rc = self._L.vdbm_map_mirror(self._h, int(chunk_leaves), cb, None, C.byref(done))
        if raised:
            raise raised[0]
        self._check(rc)
        return int(done.value)

    def mapGeneration(self) -> int:
        return int(self._L.vdbm_map_generation(self._h))

    def _section(self, bbmin, bbmax, full, result_float) -> LeafSet:
        mn = np.ascontiguousarray(bbmin, dtype=np.int32)
        mx = np.ascontiguousarray(bbmax, dtype=np.int32)
        out = C.c_void_p()
        i32p = C.POINTER(C.c_int32)
        self._check(self._L.vdbm_section(self._h, mn.ctypes.data_as(i32p), mx.ctypes.data_as(i32p), int(full),
                                         int(result_float), C.byref(out)))
        return self._take(out, bool(result_float))

    def getMapSectionUpdateGrid(self, bbmin, bbmax, full=False) -> LeafSet:
        return self._section(bbmin, bbmax, full, 0)

    def getMapSectionGrid(self, bbmin, bbmax, full=False) -> LeafSet:
        return self._section(bbmin, bbmax, full, 1)

    def applyMapSectionUpdateGrid(self, bbmin, bbmax, section: LeafSet):
        """VDBMapping.hpp:1058-1085 (bb_min / bb_max = the section grid's metadata)."""
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        mn = np.ascontiguousarray(bbmin, dtype=np.int32)
        mx = np.ascontiguousarray(bbmax, dtype=np.int32)
        o = np.ascontiguousarray(section.origins, dtype=np.int32)
        a = np.ascontiguousarray(section.active, dtype=np.uint64)
        self._check(self._L.vdbm_section_apply_update(self._h, mn.ctypes.data_as(i32p), mx.ctypes.data_as(i32p), o.shape[0],
                                                      o.ctypes.data_as(i32p), a.ctypes.data_as(u64p)))

    def applyMapSectionGrid(self, section: LeafSet, tile_quirk: bool = True):
        """VDBMapping.hpp:1022-1047."""
        i32p, u64p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
        o = np.ascontiguousarray(section.origins, dtype=np.int32)
        a = np.ascontiguousarray(section.active, dtype=np.uint64)
        v = np.ascontiguousarray(section.values, dtype=np.float32)
        self._check(self._L.vdbm_section_apply_grid(self._h, o.shape[0], o.ctypes.data_as(i32p), a.ctypes.data_as(u64p),
                                                    v.ctypes.data_as(f32p), int(tile_quirk)))

    # ---- remote-mapping deltas (level semantics in include/vdbm_b200.h) ----
    def createUpdate(self, source_id: str, level: int):
        """-> (LeafSet, origin of the source's last accumulate). level 0: raw update grid, 1: change/overwrite grid of the
        last integrate, 2: reduced update (ray end voxels, value = hit)."""
        out = C.c_void_p()
        o = np.zeros(3, dtype=np.float64)
        self._check(self._L.vdbm_update_create(self._h, source_id.encode(), int(level), C.byref(out), _dp(o)))
        return self._take(out, False), o

    def applyUpdate(self, level: int, update: LeafSet, origin=None, want_change: bool = False):
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        o = np.ascontiguousarray(update.origins, dtype=np.int32)
        a = np.ascontiguousarray(update.active, dtype=np.uint64)
        v = np.ascontiguousarray(update.valmask, dtype=np.uint64)
        og = np.ascontiguousarray(origin if origin is not None else [0, 0, 0], dtype=np.float64)
        out = C.c_void_p()
        self._check(self._L.vdbm_update_apply(self._h, int(level), o.shape[0], o.ctypes.data_as(i32p), a.ctypes.data_as(u64p),
                                              v.ctypes.data_as(u64p), _dp(og), C.byref(out) if want_change else None))
        return self._take(out, False) if want_change else None

    # ---- direct map edits + artificial areas ----
    def addPointsToGrid(self, points) -> bool:
        p = _pts16(points)
        self._check(self._L.vdbm_points_set(self._h, p.ctypes.data, p.shape[0], 16, 1))
        return True  # VDBMapping.hpp:446

    def removePointsFromGrid(self, points) -> bool:
        p = _pts16(points)
        self._check(self._L.vdbm_points_set(self._h, p.ctypes.data, p.shape[0], 16, 0))
        return True  # VDBMapping.hpp:428

    def addArtificialAreas(self, polygons, negative_height: float, positive_height: float):
        """polygons: list of (k_i, >=3) arrays of world points (VDBMapping.hpp:1175-1236)."""
        counts = np.asarray([len(p) for p in polygons], dtype=np.uint32)
        xyz = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64)[:, :3] for p in polygons])
                                   if len(polygons) else np.zeros((0, 3)))
        self._check(self._L.vdbm_artificial_areas_add(self._h, len(polygons), counts.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(xyz),
                                                      float(negative_height), float(positive_height)))

    def restoreMapIntegrity(self):
        self._check(self._L.vdbm_map_integrity_restore(self._h))

    def exportArtificialAreaGrid(self) -> LeafSet:
        out = C.c_void_p()
        self._check(self._L.vdbm_artificial_export(self._h, C.byref(out)))
        return self._take(out, False)

    def addArtificialWall(self, start, end, negative_height: float, positive_height: float):
        """VDBMapping.hpp:1217-1236 (no restoreMapIntegrity, unlike addArtificialAreas)."""
        xyz = np.ascontiguousarray(np.stack([np.asarray(start, dtype=np.float64)[:3], np.asarray(end, dtype=np.float64)[:3]]))
        counts = np.asarray([2], dtype=np.uint32)
        self._check(self._L.vdbm_artificial_walls_add(self._h, 1, counts.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(xyz), float(negative_height),
                                                       float(positive_height), 0))

    def addArtificialPolygon(self, polygon, negative_height: float, positive_height: float):
        """VDBMapping.hpp:1198-1207: one wall per edge, closing edge included."""
        xyz = np.ascontiguousarray(np.asarray(polygon, dtype=np.float64)[:, :3])
        counts = np.asarray([xyz.shape[0]], dtype=np.uint32)
        self._check(self._L.vdbm_artificial_walls_add(self._h, 1, counts.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(xyz), float(negative_height),
                                                       float(positive_height), 1))

    # ---- fast_mode (VDBMapping.hpp:577-602) and raytrace (VDBMapping.hpp:675-721) ----
    def setFastMode(self, on: bool):
        """Config::fast_mode (VDBMapping.hpp:1466): accumulate / insert use castRayIntoGridFast."""
        self._check(self._L.vdbm_set_fast_mode(self._h, int(bool(on))))

    def raytrace(self, origins, directions, max_lengths):
        """Batch raytrace (VDBMapping.hpp:675-721): returns (success bool[n], end_points float64[n, 3])."""
        o = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).reshape(-1, 3))
        d = np.ascontiguousarray(np.asarray(directions, dtype=np.float64).reshape(-1, 3))
        ln = np.ascontiguousarray(np.broadcast_to(np.asarray(max_lengths, dtype=np.float64), (o.shape[0],)))
        ok = np.zeros(o.shape[0], dtype=np.int32)
        e = np.zeros((o.shape[0], 3), dtype=np.float64)
        self._check(self._L.vdbm_raytrace(self._h, o.shape[0], _dp(o), _dp(d), _dp(ln), ok.ctypes.data_as(C.POINTER(C.c_int32)), _dp(e)))
        return ok.astype(bool), e

    def importMap(self, leaves: LeafSet, replace: bool = True):
        """loadMap (VDBMapping.hpp:263-284), device part: the leaf set becomes the map (replace) or overwrites it leaf for leaf."""
        i32p, u64p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
        o = np.ascontiguousarray(leaves.origins, dtype=np.int32)
        a = np.ascontiguousarray(leaves.active, dtype=np.uint64)
        v = np.ascontiguousarray(leaves.values, dtype=np.float32)
        self._check(self._L.vdbm_map_import(self._h, o.shape[0], o.ctypes.data_as(i32p), a.ctypes.data_as(u64p), v.ctypes.data_as(f32p), int(replace)))

    def castRaysIntoGrid(self, source_id: str, starts, ends):
        """castRayIntoGrid (VDBMapping.hpp:550-566) for explicit voxel index pairs, into the source's update grid."""
        r = np.ascontiguousarray(np.concatenate([np.asarray(starts, dtype=np.int32).reshape(-1, 3), np.asarray(ends, dtype=np.int32).reshape(-1, 3)], axis=1))
        self._check(self._L.vdbm_cast_index_rays(self._h, source_id.encode(), r.shape[0], r.ctypes.data_as(C.POINTER(C.c_int32))))

    def probe(self, coord):
        c = np.ascontiguousarray(coord, dtype=np.int32)
        v, a = C.c_float(0), C.c_int32(0)
        self._check(self._L.vdbm_probe(self._h, c.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(v), C.byref(a)))
        return float(np.float32(v.value)), bool(a.value)

    def stats(self) -> dict:
        s = L.VdbmStats()
        self._check(self._L.vdbm_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in L.VdbmStats._fields_ if n != "reserved"}

    def pipelineCounts(self) -> dict:
        """Scans of insertPointCloudAsync by path: queued / redone after a guard refusal / synchronous / overlapped."""
        out = np.zeros(4, dtype=np.uint64)
        self._check(self._L.vdbm_pipeline_counts(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return dict(zip(("queued", "redone", "synchronous", "overlapped"), (int(x) for x in out)))

    def mapLeafCount(self) -> int:
        return int(self.stats()["map_leaves"])

    def synchronize(self):
        self._check(self._L.vdbm_synchronize(self._h))

    # ---- multi-GPU plumbing (see dist.py) -------------------------------------------------------------
    def setShardPlan(self, plan):
        """plan: dist.ShardPlan (mode 0 = hash ownership, 1 = azimuth sectors). Same plan on every rank, before the first insert."""
        c = np.ascontiguousarray([plan.cx, plan.cy], dtype=np.int32)
        b = np.ascontiguousarray(plan.bounds if plan.mode else [0.0], dtype=np.float64)
        self._check(self._L.vdbm_shard_plan_set(self._h, int(plan.mode), int(plan.n_ranks), c.ctypes.data_as(C.POINTER(C.c_int32)), _dp(b)))

    def leafOwnerPlanned(self, origin, n_ranks: int) -> int:
        o = np.ascontiguousarray(origin, dtype=np.int32)
        return int(self._L.vdbm_leaf_owner_planned(self._h, o.ctypes.data_as(C.POINTER(C.c_int32)), n_ranks))

    def mapChecksum(self):
        """-> (order-independent 64-bit checksum of this handle's leaves, number of leaves); add over ranks mod 2^64."""
        out = np.zeros(2, dtype=np.uint64)
        self._check(self._L.vdbm_map_checksum(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return int(out[0]), int(out[1])

    def partitionUpdate(self, source_id: str, n_ranks: int):
        counts = np.zeros(n_ranks, dtype=np.uint64)
        ptr = C.c_void_p()
        self._check(self._L.vdbm_update_partition(self._h, source_id.encode(), n_ranks,
                                                  counts.ctypes.data_as(C.POINTER(C.c_uint64)), C.byref(ptr)))
        return counts, (ptr.value or 0)

    def importUpdateDevice(self, source_id: str, ptr: int, n: int):
        self._check(self._L.vdbm_update_import_device(self._h, source_id.encode(), C.c_void_p(ptr), int(n)))


    # peer-memory exchange (CUDA IPC over NVLink), see dist.py
    def exchangeCreate(self, rank: int, n_ranks: int, capacity_records_per_sender: int) -> bytes:
        buf = C.create_string_buffer(L.VDBM_IPC_HANDLE_BYTES)
        self._check(self._L.vdbm_exchange_create(self._h, rank, n_ranks, capacity_records_per_sender, buf))
        return buf.raw

    def exchangeConnect(self, all_handles: bytes | None):
        self._check(self._L.vdbm_exchange_connect(self._h, C.c_char_p(all_handles) if all_handles else None))

    def updatePush(self, source_id: str):
        self._check(self._L.vdbm_update_push(self._h, source_id.encode()))

    def updatePull(self, source_id: str):
        self._check(self._L.vdbm_update_pull(self._h, source_id.encode()))

    def updatePullIntegrate(self, source_id: str):
        self._check(self._L.vdbm_update_pull_integrate(self._h, source_id.encode()))

    def exchangeTimings(self):
        out = np.zeros(3, dtype=np.float32)
        self._check(self._L.vdbm_exchange_timings(self._h, out.ctypes.data_as(C.POINTER(C.c_float))))
        return [float(x) for x in out]  # push kernels, wait for peers, import + compaction (ms)


def leaf_owner(origin, n_ranks: int) -> int:
    o = np.ascontiguousarray(origin, dtype=np.int32)
    return int(L.lib().vdbm_leaf_owner(o.ctypes.data_as(C.POINTER(C.c_int32)), n_ranks))


class _ShardView(OccupancyVDBMapping):
    """A shard of a group seen through the single-map interface (exports, sections, probes, stats). Not owned: the group
    destroys it."""

    def __init__(self, lib, handle, resolution):  # noqa: D401 - no vdbm_create here
        self._L = lib
        self._h = handle
        self.resolution = resolution
        self.stream_handle = 0

    def close(self):
        self._h = None


class OccupancyVDBMappingGroup:
    """One process, several GPUs: vdbm_group_* of include/vdbm_b200.h (a sharded map whose union is the reference's one map).
    insertPointCloud gives every shard the whole cloud; each casts the rays of its azimuth sector, foreign update leaves go to
    their owners over NVLink peer memory, every shard integrates its own leaves."""

    def __init__(self, resolution: float, devices, replicate_probe_quirk: bool = True, update_capacity_leaves: int = 0,
                 map_capacity_leaves: int = 0, inbox_capacity_records: int = 0):
        self._L = L.lib()
        self.resolution = float(resolution)
        devs = np.ascontiguousarray(list(devices), dtype=np.int32)
        p = L.VdbmParams(float(resolution), -1, int(replicate_probe_quirk), int(update_capacity_leaves), int(map_capacity_leaves), None)
        h = C.c_void_p()
        rc = self._L.vdbm_group_create(C.byref(p), len(devs), devs.ctypes.data_as(C.POINTER(C.c_int32)), int(inbox_capacity_records), C.byref(h))
        if rc != L.VDBM_OK:
            raise VdbmError(rc, "vdbm_group_create failed (needs that many CUDA devices with peer access)")
        self._h = h
        self.n = len(devs)

    def close(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.vdbm_group_destroy(h)

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _check(self, rc, allow=()):
        if rc != L.VDBM_OK and rc not in allow:
            raise VdbmError(rc, (self._L.vdbm_group_last_error(self._h) or b"").decode())
        return rc

    def setConfig(self, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max) -> int:
        rc = self._L.vdbm_group_set_config(self._h, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max)
        if rc == L.VDBM_ERR_BAD_CONFIG:
            return 1 if max_range < 0 else 2
        self._check(rc)
        return 0

    def addInputSource(self, source_id: str, max_range: float, max_rate: float = 0.0):
        self._check(self._L.vdbm_group_source_add(self._h, source_id.encode(), float(max_range)))

    def resetMap(self):
        self._check(self._L.vdbm_group_reset(self._h))

    def insertPointCloud(self, points, origin, source_id: str) -> bool:
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        rc = self._L.vdbm_group_insert(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o))
        self._check(rc, allow=(L.VDBM_ERR_UNKNOWN_SOURCE, L.VDBM_ERR_NOT_CONFIGURED))
        return True

    def insertRaw(self, ptr: int, n: int, origin, source_id: str, stride: int = 16):
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._check(self._L.vdbm_group_insert(self._h, source_id.encode(), C.c_void_p(ptr), n, stride, _dp(o)))

    def setPlan(self, center_leaf_xy, ray_bounds, ownership_bounds):
        c = np.ascontiguousarray(center_leaf_xy, dtype=np.int32)
        rb = np.ascontiguousarray(ray_bounds, dtype=np.float64)
        ob = np.ascontiguousarray(ownership_bounds, dtype=np.float64)
        self._check(self._L.vdbm_group_plan_set(self._h, c.ctypes.data_as(C.POINTER(C.c_int32)), _dp(rb), _dp(ob)))

    def plan(self):
        c = np.zeros(2, dtype=np.int32)
        rb, ob = np.zeros(self.n), np.zeros(self.n)
        self._check(self._L.vdbm_group_plan_get(self._h, c.ctypes.data_as(C.POINTER(C.c_int32)), _dp(rb), _dp(ob)))
        return c, rb, ob

    def shard(self, i: int) -> OccupancyVDBMapping:
        return _ShardView(self._L, C.c_void_p(self._L.vdbm_group_shard(self._h, i)), self.resolution)

    def checksum(self):
        out = np.zeros(2, dtype=np.uint64)
        self._check(self._L.vdbm_group_checksum(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64))))
        return int(out[0]), int(out[1])

    def stats(self) -> dict:
        s = L.VdbmStats()
        self._check(self._L.vdbm_group_stats(self._h, C.byref(s)))
        return {n: getattr(s, n) for n, _ in L.VdbmStats._fields_ if n != "reserved"}

    def exportMap(self) -> LeafSet:
        """Union of the shards' (disjoint) leaf sets, sorted by origin like every export."""
        parts = [self.shard(i).exportMap() for i in range(self.n)]
        origins = np.concatenate([p.origins for p in parts])
        active = np.concatenate([p.active for p in parts])
        values = np.concatenate([p.values for p in parts])
        order = np.lexsort((origins[:, 2], origins[:, 1], origins[:, 0]))
        return LeafSet(origins[order], active[order], None, values[order])
