"""ctypes wrapper around oracle/libvdbm_oracle.so — the CPU ORACLE.

TEST INFRASTRUCTURE ONLY: importable from tests/, __graft_entry__.smoke() and bench.py's
cpu_baseline / --impl reference legs. The product package (vdb_mapping_b200/) never imports this.

The class mirrors the reference's OccupancyVDBMapping surface for the scan-integration path
(/root/reference/include/vdb_mapping/VDBMapping.hpp:316-406,731-799,883-960 and
OccupancyVDBMapping.hpp:59-117) so parity tests read like the reference's tests/mapping.cpp.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess
from dataclasses import dataclass

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_SO = os.path.join(_HERE, "libvdbm_oracle.so")


_FLAT_SO = os.path.join(_HERE, "libvdbm_flat_probe.so")


def build(force: bool = False) -> str:
    """Compile the oracle (and the flat-hash probe) with oracle/Makefile (g++). Returns the oracle .so path."""
    stale = False
    for so, src in ((_SO, "vdbm_oracle.cpp"), (_FLAT_SO, "flat_probe.cpp")):
        src = os.path.join(_HERE, src)
        stale = stale or not os.path.exists(so) or os.path.getmtime(so) < os.path.getmtime(src)
    if force or stale:
        subprocess.run(["make", "-C", _HERE, "-s"], check=True)
    return _SO


class FlatProbe:
    """The 'optimistic CPU' yardstick of oracle/flat_probe.cpp: same arithmetic on a flat leaf hash, optional threads."""

    def __init__(self, resolution: float, logodds6):
        build()
        self._L = C.CDLL(_FLAT_SO)
        self._L.flat_create.restype = C.c_void_p
        self._L.flat_create.argtypes = [C.c_double, C.POINTER(C.c_float)]
        self._L.flat_destroy.argtypes = [C.c_void_p]
        self._L.flat_insert.argtypes = [C.c_void_p, C.c_void_p, C.c_uint64, C.c_uint64, C.POINTER(C.c_double), C.c_double, C.c_int]
        self._L.flat_stats.argtypes = [C.c_void_p, C.POINTER(C.c_uint64)]
        lo = np.ascontiguousarray(logodds6, dtype=np.float32)
        self._h = self._L.flat_create(resolution, lo.ctypes.data_as(C.POINTER(C.c_float)))

    def __del__(self):
        if getattr(self, "_h", None):
            self._L.flat_destroy(self._h)
            self._h = None

    def insert(self, points, origin, max_range: float, threads: int = 1):
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._L.flat_insert(self._h, p.ctypes.data, p.shape[0], 16, _dp(o), float(max_range), int(threads))

    def stats(self) -> dict:
        out = np.zeros(4, dtype=np.uint64)
        self._L.flat_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        return dict(zip(["visits", "voxel_updates", "map_leaves", "active_voxels"], (int(x) for x in out)))


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_SO)
        vp, dbl, i32p, u64p, f32p = C.c_void_p, C.c_double, C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
        L.vdbo_create.restype = vp
        L.vdbo_create.argtypes = [dbl]
        L.vdbo_destroy.argtypes = [vp]
        L.vdbo_set_config.restype = C.c_int
        L.vdbo_set_config.argtypes = [vp, dbl, dbl, dbl, dbl, dbl]
        L.vdbo_set_probe_quirk.argtypes = [vp, C.c_int]
        L.vdbo_get_logodds.argtypes = [vp, f32p]
        L.vdbo_add_source.argtypes = [vp, C.c_char_p, dbl]
        L.vdbo_reset.argtypes = [vp]
        L.vdbo_world_to_index.argtypes = [vp, C.POINTER(dbl), i32p]
        L.vdbo_accumulate.restype = C.c_int
        L.vdbo_accumulate.argtypes = [vp, C.c_char_p, vp, C.c_uint64, C.c_uint64, C.POINTER(dbl)]
        L.vdbo_insert.restype = C.c_int
        L.vdbo_insert.argtypes = [vp, C.c_char_p, vp, C.c_uint64, C.c_uint64, C.POINTER(dbl)]
        L.vdbo_integrate.argtypes = [vp]
        L.vdbo_stats.argtypes = [vp, u64p]
        L.vdbo_export_prepare.restype = C.c_int64
        L.vdbo_export_prepare.argtypes = [vp, C.c_int, C.c_char_p, i32p, i32p, C.c_int]
        L.vdbo_export_fetch.argtypes = [vp, i32p, u64p, u64p, f32p]
        L.vdbo_probe.restype = C.c_int
        L.vdbo_probe.argtypes = [vp, i32p, f32p]
        L.vdbo_map_leaf_count.restype = C.c_uint64
        L.vdbo_map_leaf_count.argtypes = [vp]
        L.vdbo_update_import.restype = C.c_int
        L.vdbo_update_import.argtypes = [vp, C.c_char_p, C.c_uint64, i32p, u64p, u64p]
        L.vdbo_apply_section_update.restype = C.c_int
        L.vdbo_apply_section_update.argtypes = [vp, i32p, i32p, C.c_uint64, i32p, u64p]
        L.vdbo_apply_section_grid.restype = C.c_int
        L.vdbo_apply_section_grid.argtypes = [vp, C.c_uint64, i32p, u64p, f32p, C.c_int]
        L.vdbo_update_clear.restype = C.c_int
        L.vdbo_update_clear.argtypes = [vp, C.c_char_p]
        L.vdbo_update_apply.restype = C.c_int
        L.vdbo_update_apply.argtypes = [vp, C.c_char_p, C.c_int, C.c_uint64, i32p, u64p, u64p, C.POINTER(dbl)]
        L.vdbo_last_origin.argtypes = [vp, C.c_char_p, C.POINTER(dbl)]
        L.vdbo_points_set.argtypes = [vp, vp, C.c_uint64, C.c_uint64, C.c_int]
        L.vdbo_add_artificial_areas.argtypes = [vp, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(dbl), dbl, dbl]
        L.vdbo_restore_map_integrity.argtypes = [vp]
        L.vdbo_set_fast_mode.argtypes = [vp, C.c_int]
        L.vdbo_raytrace.argtypes = [vp, C.c_uint64, C.POINTER(dbl), C.POINTER(dbl), C.POINTER(dbl), i32p, C.POINTER(dbl)]
        L.vdbo_add_artificial_wall.argtypes = [vp, C.POINTER(dbl), C.POINTER(dbl), dbl, dbl]
        L.vdbo_cast_index_rays.restype = C.c_int
        L.vdbo_cast_index_rays.argtypes = [vp, C.c_char_p, C.c_uint64, i32p]
        _lib = L
    return _lib


@dataclass
class LeafSet:
    """Canonical leaf list, sorted by leaf origin (x, y, z)."""
    origins: np.ndarray          # (n, 3) int32, multiples of 8
    active: np.ndarray           # (n, 8) uint64 — word = x&7, bit = (y&7)<<3 | (z&7)
    valmask: np.ndarray | None   # (n, 8) uint64 for bool grids
    values: np.ndarray | None    # (n, 512) float32 for float grids, offset = (x&7)<<6|(y&7)<<3|(z&7)

    def __len__(self):
        return int(self.origins.shape[0])


def _pts16(points: np.ndarray) -> np.ndarray:
    """(n,3) or (n,4) float32 -> contiguous pcl::PointXYZ layout (16-byte stride)."""
    p = np.asarray(points, dtype=np.float32)
    if p.ndim != 2 or p.shape[1] not in (3, 4):
        raise ValueError("points must be (n,3) or (n,4) float32")
    if p.shape[1] == 3:
        q = np.zeros((p.shape[0], 4), dtype=np.float32)
        q[:, :3] = p
        q[:, 3] = 1.0
        p = q
    return np.ascontiguousarray(p)


def _dp(a):
    return a.ctypes.data_as(C.POINTER(C.c_double))


class OracleOccupancyVDBMapping:
    def __init__(self, resolution: float):
        self._L = lib()
        self._h = self._L.vdbo_create(float(resolution))
        self.resolution = float(resolution)

    def __del__(self):
        h, self._h = getattr(self, "_h", None), None
        if h:
            self._L.vdbo_destroy(h)

    # setConfig: 0 ok, 1 rejected by base (max_range < 0), 2 rejected by Occupancy (probabilities)
    def setConfig(self, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max) -> int:
        return self._L.vdbo_set_config(self._h, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max)

    def setProbeQuirk(self, on: bool):
        self._L.vdbo_set_probe_quirk(self._h, int(on))

    def logodds(self) -> np.ndarray:
        out = np.zeros(6, dtype=np.float32)
        self._L.vdbo_get_logodds(self._h, out.ctypes.data_as(C.POINTER(C.c_float)))
        return out

    def addInputSource(self, source_id: str, max_range: float, max_rate: float = 0.0):
        self._L.vdbo_add_source(self._h, source_id.encode(), float(max_range))

    def resetMap(self):
        self._L.vdbo_reset(self._h)

    def worldToIndex(self, w) -> np.ndarray:
        w = np.ascontiguousarray(w, dtype=np.float64)
        out = np.zeros(3, dtype=np.int32)
        self._L.vdbo_world_to_index(self._h, _dp(w), out.ctypes.data_as(C.POINTER(C.c_int32)))
        return out

    def accumulateUpdate(self, points, origin, source_id: str) -> int:
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        return self._L.vdbo_accumulate(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o))

    def integrateUpdate(self):
        self._L.vdbo_integrate(self._h)

    def insertPointCloud(self, points, origin, source_id: str) -> bool:
        p = _pts16(points)
        o = np.ascontiguousarray(origin, dtype=np.float64)
        self._L.vdbo_insert(self._h, source_id.encode(), p.ctypes.data, p.shape[0], 16, _dp(o))
        return True

    def stats(self) -> dict:
        out = np.zeros(6, dtype=np.uint64)
        self._L.vdbo_stats(self._h, out.ctypes.data_as(C.POINTER(C.c_uint64)))
        k = ["rays", "nan_skipped", "clipped", "visits", "voxel_updates", "state_changes"]
        return {a: int(b) for a, b in zip(k, out)}

    def _export(self, kind, source=None, bbmin=None, bbmax=None, full=False) -> LeafSet:
        i32p = C.POINTER(C.c_int32)
        mn = np.ascontiguousarray(bbmin if bbmin is not None else [0, 0, 0], dtype=np.int32)
        mx = np.ascontiguousarray(bbmax if bbmax is not None else [0, 0, 0], dtype=np.int32)
        n = self._L.vdbo_export_prepare(self._h, kind, (source or "").encode(), mn.ctypes.data_as(i32p),
                                        mx.ctypes.data_as(i32p), int(full))
        if n < 0:
            raise KeyError(f"export kind {kind} source {source!r} failed ({n})")
        origins = np.zeros((n, 3), dtype=np.int32)
        active = np.zeros((n, 8), dtype=np.uint64)
        is_float = kind in (0, 4)

        valmask = None if is_float else np.zeros((n, 8), dtype=np.uint64)
        values = np.zeros((n, 512), dtype=np.float32) if is_float else None
        self._L.vdbo_export_fetch(
            self._h, origins.ctypes.data_as(i32p), active.ctypes.data_as(C.POINTER(C.c_uint64)),
            valmask.ctypes.data_as(C.POINTER(C.c_uint64)) if valmask is not None else None,
            values.ctypes.data_as(C.POINTER(C.c_float)) if values is not None else None)
        return LeafSet(origins, active, valmask, values)

    def exportMap(self) -> LeafSet:
        return self._export(0)

    def exportUpdateGrid(self, source_id: str) -> LeafSet:
        return self._export(1, source_id)

    def exportLastChange(self, source_id: str) -> LeafSet:
        return self._export(2, source_id)

    def getMapSectionUpdateGrid(self, bbmin, bbmax, full=False) -> LeafSet:
        return self._export(3, None, bbmin, bbmax, full)

    def getMapSectionGrid(self, bbmin, bbmax, full=False) -> LeafSet:
        return self._export(4, None, bbmin, bbmax, full)

    def probe(self, coord):
        c = np.ascontiguousarray(coord, dtype=np.int32)
        v = C.c_float(0)
        on = self._L.vdbo_probe(self._h, c.ctypes.data_as(C.POINTER(C.c_int32)), C.byref(v))
        return float(np.float32(v.value)), bool(on)

    def mapLeafCount(self) -> int:
        return int(self._L.vdbo_map_leaf_count(self._h))

    def applyMapSectionUpdateGrid(self, bbmin, bbmax, section: "LeafSet"):
        """VDBMapping.hpp:1058-1085 with the section's bb_min / bb_max metadata passed explicitly."""
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        mn = np.ascontiguousarray(bbmin, dtype=np.int32); mx = np.ascontiguousarray(bbmax, dtype=np.int32)
        o = np.ascontiguousarray(section.origins, dtype=np.int32); a = np.ascontiguousarray(section.active, dtype=np.uint64)
        self._L.vdbo_apply_section_update(self._h, mn.ctypes.data_as(i32p), mx.ctypes.data_as(i32p), o.shape[0],
                                          o.ctypes.data_as(i32p), a.ctypes.data_as(u64p))

    def applyMapSectionGrid(self, section: "LeafSet", tile_quirk: bool = True):
        """VDBMapping.hpp:1022-1047."""
        i32p, u64p, f32p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float)
        o = np.ascontiguousarray(section.origins, dtype=np.int32); a = np.ascontiguousarray(section.active, dtype=np.uint64)
        v = np.ascontiguousarray(section.values, dtype=np.float32)
        self._L.vdbo_apply_section_grid(self._h, o.shape[0], o.ctypes.data_as(i32p), a.ctypes.data_as(u64p), v.ctypes.data_as(f32p), int(tile_quirk))

    def clearUpdateGrid(self, source_id: str) -> int:
        return self._L.vdbo_update_clear(self._h, source_id.encode())

    def importUpdate(self, source_id: str, origins, active, valmask) -> int:
        o = np.ascontiguousarray(origins, dtype=np.int32)
        a = np.ascontiguousarray(active, dtype=np.uint64)
        v = np.ascontiguousarray(valmask, dtype=np.uint64)
        return self._L.vdbo_update_import(self._h, source_id.encode(), o.shape[0],
                                          o.ctypes.data_as(C.POINTER(C.c_int32)),
                                          a.ctypes.data_as(C.POINTER(C.c_uint64)),
                                          v.ctypes.data_as(C.POINTER(C.c_uint64)))

    # ---- remote-mapping deltas: createUpdate / applyUpdate ----
    def createUpdate(self, source_id: str, level: int):
        """level 0: the source's raw update grid; 1: the change ("overwrite") grid of the last integrate; 2: the reduced
        update of the last accumulate (ray end voxels, value = hit). Returns (LeafSet, origin of that accumulate)."""
        o = np.zeros(3, dtype=np.float64)
        self._L.vdbo_last_origin(self._h, source_id.encode(), _dp(o))
        if level == 0:
            return self._export(1, source_id), o
        if level == 1:
            return self._export(2, source_id), o
        if level == 2:
            return self._export(1, source_id, full=2), o
        raise ValueError(level)

    def applyUpdate(self, source_id: str, level: int, update: "LeafSet", origin=None) -> int:
        i32p, u64p = C.POINTER(C.c_int32), C.POINTER(C.c_uint64)
        o = np.ascontiguousarray(update.origins, dtype=np.int32)
        a = np.ascontiguousarray(update.active, dtype=np.uint64)
        v = np.ascontiguousarray(update.valmask, dtype=np.uint64)
        og = np.ascontiguousarray(origin if origin is not None else [0, 0, 0], dtype=np.float64)
        return self._L.vdbo_update_apply(self._h, source_id.encode(), level, o.shape[0], o.ctypes.data_as(i32p),
                                         a.ctypes.data_as(u64p), v.ctypes.data_as(u64p), _dp(og))

    # ---- direct map edits + artificial areas ----
    def addPointsToGrid(self, points):
        p = _pts16(points)
        self._L.vdbo_points_set(self._h, p.ctypes.data, p.shape[0], 16, 1)

    def removePointsFromGrid(self, points):
        p = _pts16(points)
        self._L.vdbo_points_set(self._h, p.ctypes.data, p.shape[0], 16, 0)

    def addArtificialAreas(self, polygons, negative_height: float, positive_height: float):
        """polygons: list of (k_i, 3) arrays of world points (VDBMapping.hpp:1175-1236)."""
        counts = np.asarray([len(p) for p in polygons], dtype=np.uint32)
        xyz = np.ascontiguousarray(np.concatenate([np.asarray(p, dtype=np.float64)[:, :3] for p in polygons]) if len(polygons) else np.zeros((0, 3)))
        self._L.vdbo_add_artificial_areas(self._h, len(polygons), counts.ctypes.data_as(C.POINTER(C.c_uint32)), _dp(xyz),
                                          float(negative_height), float(positive_height))

    def restoreMapIntegrity(self):
        self._L.vdbo_restore_map_integrity(self._h)

    def exportArtificialAreaGrid(self) -> LeafSet:
        return self._export(5)

    def addArtificialWall(self, start, end, negative_height: float, positive_height: float):
        a = np.ascontiguousarray(np.asarray(start, dtype=np.float64)[:3])
        b = np.ascontiguousarray(np.asarray(end, dtype=np.float64)[:3])
        self._L.vdbo_add_artificial_wall(self._h, _dp(a), _dp(b), float(negative_height), float(positive_height))

    def addArtificialPolygon(self, polygon, negative_height: float, positive_height: float):
        """VDBMapping.hpp:1198-1207: one wall per edge, closing edge included."""
        pts = np.asarray(polygon, dtype=np.float64)[:, :3]
        for i in range(len(pts)):
            self.addArtificialWall(pts[i], pts[(i + 1) % len(pts)], negative_height, positive_height)

    # ---- fast_mode (VDBMapping.hpp:577-602) and raytrace (VDBMapping.hpp:675-721) ----
    def setFastMode(self, on: bool):
        self._L.vdbo_set_fast_mode(self._h, int(bool(on)))

    def raytrace(self, origins, directions, max_lengths):
        o = np.ascontiguousarray(np.asarray(origins, dtype=np.float64).reshape(-1, 3))
        d = np.ascontiguousarray(np.asarray(directions, dtype=np.float64).reshape(-1, 3))
        ln = np.ascontiguousarray(np.broadcast_to(np.asarray(max_lengths, dtype=np.float64), (o.shape[0],)))
        ok = np.zeros(o.shape[0], dtype=np.int32)
        e = np.zeros((o.shape[0], 3), dtype=np.float64)
        self._L.vdbo_raytrace(self._h, o.shape[0], _dp(o), _dp(d), _dp(ln), ok.ctypes.data_as(C.POINTER(C.c_int32)), _dp(e))
        return ok.astype(bool), e

    def castRaysIntoGrid(self, source_id: str, starts, ends):
        r = np.ascontiguousarray(np.concatenate([np.asarray(starts, dtype=np.int32).reshape(-1, 3), np.asarray(ends, dtype=np.int32).reshape(-1, 3)], axis=1))
        return self._L.vdbo_cast_index_rays(self._h, source_id.encode(), r.shape[0], r.ctypes.data_as(C.POINTER(C.c_int32)))
