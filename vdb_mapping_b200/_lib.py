"""ctypes binding of libvdbm_b200.so (the C ABI in include/vdbm_b200.h).

The product path has no CPU fallback: if the CUDA library is missing this module raises.
"""
from __future__ import annotations

import ctypes as C
import os

HERE = os.path.dirname(os.path.abspath(__file__))
SO_PATH = os.path.join(HERE, "libvdbm_b200.so")

VDBM_OK, VDBM_ERR_INVALID_ARG, VDBM_ERR_NOT_CONFIGURED, VDBM_ERR_UNKNOWN_SOURCE = 0, 1, 2, 3
VDBM_ERR_BAD_CONFIG, VDBM_ERR_CUDA, VDBM_ERR_OUT_OF_MEMORY, VDBM_ERR_COORD_RANGE = 4, 5, 6, 7
VDBM_IPC_HANDLE_BYTES = 128


class VdbmParams(C.Structure):
    _fields_ = [("resolution", C.c_double), ("device", C.c_int32), ("replicate_probe_quirk", C.c_int32),
                ("update_capacity_leaves", C.c_uint64), ("map_capacity_leaves", C.c_uint64), ("stream", C.c_void_p)]


class VdbmStats(C.Structure):
    _fields_ = [(n, C.c_uint64) for n in ("rays", "nan_skipped", "clipped", "visits", "voxel_updates", "state_changes",
                                          "map_leaves", "new_leaves", "last_touched_leaves", "last_voxel_updates",
                                          "last_visits")] + \
               [("last_accumulate_ms", C.c_float), ("last_integrate_ms", C.c_float), ("last_prep_ms", C.c_float),
                ("update_capacity", C.c_uint32), ("map_capacity", C.c_uint32), ("gpu_launches", C.c_uint32)]


# vdbm_mirror_sink: (user, n, leaf_index[n], origins[n][3], values[n][512], active[n][8]) -> 0 to go on
MIRROR_SINK = C.CFUNCTYPE(C.c_int, C.c_void_p, C.c_uint64, C.POINTER(C.c_uint32), C.POINTER(C.c_int32), C.POINTER(C.c_float),
                          C.POINTER(C.c_uint64))

_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(SO_PATH):
        raise ImportError(
            f"{SO_PATH} is missing: build it with `python -m vdb_mapping_b200.build` (nvcc, sm_100a). "
            "vdb_mapping_b200 has no CPU fallback.")
    L = C.CDLL(SO_PATH)
    vp, cp, dbl, i32, u64 = C.c_void_p, C.c_char_p, C.c_double, C.c_int32, C.c_uint64
    pvp = C.POINTER(C.c_void_p)
    i32p, u64p, f32p, dblp = C.POINTER(C.c_int32), C.POINTER(C.c_uint64), C.POINTER(C.c_float), C.POINTER(C.c_double)

    def sig(name, res, *args):
        f = getattr(L, name)
        f.restype = res
        f.argtypes = list(args)

    sig("vdbm_abi_version", C.c_int)
    sig("vdbm_create", C.c_int, C.POINTER(VdbmParams), pvp)
    sig("vdbm_destroy", None, vp)
    sig("vdbm_reset", C.c_int, vp)
    sig("vdbm_set_config", C.c_int, vp, dbl, dbl, dbl, dbl, dbl)
    sig("vdbm_get_logodds", C.c_int, vp, f32p)
    sig("vdbm_source_add", C.c_int, vp, cp, dbl)
    sig("vdbm_accumulate", C.c_int, vp, cp, vp, u64, u64, dblp)
    sig("vdbm_integrate_from", C.c_int, vp, vp, cp, C.c_int)
    sig("vdbm_accumulate_device", C.c_int, vp, cp, vp, u64, u64, dblp)
    sig("vdbm_raycast", C.c_int, vp, cp, vp, u64, u64, dblp, dbl)
    sig("vdbm_integrate", C.c_int, vp, C.c_int)
    sig("vdbm_insert", C.c_int, vp, cp, vp, u64, u64, dblp)
    sig("vdbm_insert_async", C.c_int, vp, cp, vp, u64, u64, dblp, C.c_int)
    sig("vdbm_flush", C.c_int, vp)
    sig("vdbm_prefetch", C.c_int, vp, vp, u64, u64)
    sig("vdbm_update_map", C.c_int, vp, cp, pvp)
    sig("vdbm_update_export", C.c_int, vp, cp, pvp)
    sig("vdbm_update_import", C.c_int, vp, cp, u64, i32p, u64p, u64p)
    sig("vdbm_change_export", C.c_int, vp, cp, pvp)
    sig("vdbm_map_export", C.c_int, vp, C.c_int, pvp)
    sig("vdbm_map_mirror", C.c_int, vp, u64, MIRROR_SINK, vp, C.POINTER(C.c_uint64))
    sig("vdbm_map_generation", C.c_uint64, vp)
    sig("vdbm_section", C.c_int, vp, i32p, i32p, C.c_int, C.c_int, pvp)
    sig("vdbm_section_apply_update", C.c_int, vp, i32p, i32p, u64, i32p, u64p)
    sig("vdbm_section_apply_grid", C.c_int, vp, u64, i32p, u64p, f32p, C.c_int)
    sig("vdbm_probe", C.c_int, vp, i32p, f32p, i32p)
    sig("vdbm_map_import", C.c_int, vp, u64, i32p, u64p, f32p, C.c_int)
    sig("vdbm_cast_index_rays", C.c_int, vp, cp, u64, i32p)
    sig("vdbm_update_create", C.c_int, vp, cp, C.c_int, pvp, dblp)
    sig("vdbm_update_apply", C.c_int, vp, C.c_int, u64, i32p, u64p, u64p, dblp, pvp)
    sig("vdbm_points_set", C.c_int, vp, vp, u64, u64, C.c_int)
    sig("vdbm_artificial_areas_add", C.c_int, vp, u64, C.POINTER(C.c_uint32), dblp, dbl, dbl)
    sig("vdbm_artificial_walls_add", C.c_int, vp, u64, C.POINTER(C.c_uint32), dblp, dbl, dbl, C.c_int)
    sig("vdbm_set_fast_mode", C.c_int, vp, C.c_int)
    sig("vdbm_raytrace", C.c_int, vp, u64, dblp, dblp, dblp, i32p, dblp)
    sig("vdbm_map_integrity_restore", C.c_int, vp)
    sig("vdbm_artificial_export", C.c_int, vp, pvp)
    sig("vdbm_leafset_size", u64, vp)
    sig("vdbm_leafset_origins", i32p, vp)
    sig("vdbm_leafset_active", u64p, vp)
    sig("vdbm_leafset_valmask", u64p, vp)
    sig("vdbm_leafset_values", f32p, vp)
    sig("vdbm_leafset_free", None, vp)
    sig("vdbm_leaf_owner", i32, i32p, i32)
    sig("vdbm_shard_plan_set", C.c_int, vp, i32, i32, i32p, dblp)
    sig("vdbm_ray_sector_set", C.c_int, vp, i32, i32, dblp)
    sig("vdbm_exchange_connect_peers", C.c_int, vp, pvp)
    sig("vdbm_group_create", C.c_int, C.POINTER(VdbmParams), i32, i32p, u64, pvp)
    sig("vdbm_group_destroy", None, vp)
    sig("vdbm_group_size", i32, vp)
    sig("vdbm_group_shard", vp, vp, i32)
    sig("vdbm_group_set_config", C.c_int, vp, dbl, dbl, dbl, dbl, dbl)
    sig("vdbm_group_source_add", C.c_int, vp, cp, dbl)
    sig("vdbm_group_reset", C.c_int, vp)
    sig("vdbm_group_insert", C.c_int, vp, cp, vp, u64, u64, dblp)
    sig("vdbm_group_plan_set", C.c_int, vp, i32p, dblp, dblp)
    sig("vdbm_group_plan_get", C.c_int, vp, i32p, dblp, dblp)
    sig("vdbm_group_checksum", C.c_int, vp, u64p)
    sig("vdbm_group_stats", C.c_int, vp, C.POINTER(VdbmStats))
    sig("vdbm_group_last_error", cp, vp)
    sig("vdbm_leaf_owner_planned", i32, vp, i32p, i32)
    sig("vdbm_map_checksum", C.c_int, vp, u64p)
    sig("vdbm_update_partition", C.c_int, vp, cp, i32, u64p, pvp)
    sig("vdbm_update_import_device", C.c_int, vp, cp, vp, u64)
    sig("vdbm_exchange_create", C.c_int, vp, i32, i32, u64, vp)
    sig("vdbm_exchange_connect", C.c_int, vp, vp)
    sig("vdbm_update_push", C.c_int, vp, cp)
    sig("vdbm_update_pull", C.c_int, vp, cp)
    sig("vdbm_update_pull_integrate", C.c_int, vp, cp)
    sig("vdbm_exchange_timings", C.c_int, vp, f32p)
    sig("vdbm_pipeline_counts", C.c_int, vp, u64p)
    sig("vdbm_stats", C.c_int, vp, C.POINTER(VdbmStats))
    sig("vdbm_last_error", cp, vp)
    sig("vdbm_synchronize", C.c_int, vp)
    sig("vdbm_host_alloc", vp, C.c_size_t)
    sig("vdbm_host_free", None, vp)
    _lib = L
    return L
