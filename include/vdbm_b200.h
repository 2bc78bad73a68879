/* vdbm_b200.h — C ABI of libvdbm_b200.so: the B200-native (sm_100a) scan-integration hot path of
 * vdb_mapping. Plain C types only (no torch / OpenVDB / PCL / Eigen in any signature).
 *
 * The reference (/root/reference) has no FFI: its boundary is the header-level C++ API of
 *   include/vdb_mapping/VDBMapping.hpp           (cited below as V:<line>)
 *   include/vdb_mapping/OccupancyVDBMapping.hpp   (cited below as O:<line>)
 * Every entry point names the reference member whose device work it replaces. The C++ shim in
 * include/vdb_mapping/ keeps the reference's class/member signatures and calls only this ABI;
 * INTEGRATION.md shows the binding.
 *
 * Conventions
 *  - every call returns a vdbm_status (0 = ok); nothing throws across the ABI;
 *    vdbm_last_error() gives a human-readable message for the last non-zero status on a handle.
 *  - handles are opaque, created/destroyed by paired calls, thread-compatible (not internally
 *    locked): the shim keeps the reference's std::shared_mutex discipline (V:327,378).
 *  - leaf layout mirrors OpenVDB: leaf origin = coord & ~7; voxel offset n = (x&7)<<6 | (y&7)<<3 |
 *    (z&7); bit masks are 8 x uint64, word n>>6 (= x&7), bit n&63. Leaf values are 512 x f32 in
 *    offset order. Exported leaf sets are sorted by origin (x, then y, then z).
 *  - there is NO CPU fallback: without a CUDA device vdbm_create fails with VDBM_ERR_CUDA.
 */
#ifndef VDBM_B200_H_INCLUDED
#define VDBM_B200_H_INCLUDED

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif
#if defined(__GNUC__)
#pragma GCC visibility push(default) /* the library is built with -fvisibility=hidden; only this ABI is exported */
#endif

#define VDBM_ABI_VERSION 1

typedef enum vdbm_status
{
  VDBM_OK                 = 0,
  VDBM_ERR_INVALID_ARG    = 1,
  VDBM_ERR_NOT_CONFIGURED = 2, /* V:478-482: raycast before setConfig -> reference prints + returns false */
  VDBM_ERR_UNKNOWN_SOURCE = 3, /* V:320-326: reference prints "Source not available" and returns */
  VDBM_ERR_BAD_CONFIG     = 4, /* V:1458-1463 / O:65-76: reference prints and keeps the old values */
  VDBM_ERR_CUDA           = 5,
  VDBM_ERR_OUT_OF_MEMORY  = 6,
  VDBM_ERR_COORD_RANGE    = 7  /* a voxel coordinate left the +-2^23 range the 63-bit leaf key covers */
} vdbm_status;

typedef struct vdbm_map vdbm_map;         /* device-resident map + per-source update grids */
typedef struct vdbm_leafset vdbm_leafset; /* host-side list of 8^3 leaves returned by exports */

typedef struct vdbm_params
{
  double resolution;               /* V:115 ctor argument */
  int32_t device;                  /* CUDA ordinal, -1 = current device */
  int32_t replicate_probe_quirk;   /* 1 (default): change grid reproduces OpenVDB's tile-probe side effect
                                      (SURVEY.md F9); 0: report only real active-flag flips */
  uint64_t update_capacity_leaves; /* initial slots of each per-source update-leaf hash (0 = 1<<20) */
  uint64_t map_capacity_leaves;    /* initial map leaf pool size (0 = 1<<19); both grow on demand */
  void* stream;                    /* cudaStream_t to launch on; NULL = library-owned stream */
} vdbm_params;

typedef struct vdbm_stats_t
{
  /* cumulative since create/reset */
  uint64_t rays;          /* points seen by the raycast (incl. NaN and clipped) */
  uint64_t nan_skipped;   /* V:505-510 (the B200 build also drops +-inf, undefined in the reference) */
  uint64_t clipped;       /* V:512-517 */
  uint64_t visits;        /* DDA voxel visits = setActiveState calls of V:563 */
  uint64_t voxel_updates; /* active update voxels applied by updateMap (V:764) */
  uint64_t state_changes; /* voxels reported in change grids (V:770-781) */
  uint64_t map_leaves;    /* 8^3 leaves currently in the map */
  uint64_t new_leaves;    /* map leaves created, cumulative */
  /* last call */
  uint64_t last_touched_leaves; /* update-grid leaves of the source(s) consumed by the last integrate/update_map,
                                   or currently accumulated after accumulate */
  uint64_t last_voxel_updates;
  uint64_t last_visits;
  float last_accumulate_ms; /* device time of the last accumulate: prep_rays + raycast_dda kernels (CUDA events) */
  float last_integrate_ms;  /* device time of the last integrate / update_map: apply_update kernel(s) */
  float last_prep_ms;       /* prep_rays kernel share of last_accumulate_ms */
  uint32_t update_capacity; /* current per-source update hash slots (max over sources) */
  uint32_t map_capacity;    /* current map leaf pool size */
  uint32_t gpu_launches;    /* kernels launched by the library since create (cumulative) */
} vdbm_stats_t;

/* ---- lifecycle --------------------------------------------------------------------------- */
/* VDBMapping(double resolution) V:115-134 + createVDBMap V:163-169 (background 0.0f, inactive). */
int vdbm_create(const vdbm_params* params, vdbm_map** out);
void vdbm_destroy(vdbm_map* map);
/* resetMap V:174-186: empty map and empty per-source update grids. */
int vdbm_reset(vdbm_map* map);
/* OccupancyVDBMapping::setConfig O:59-89 over VDBMapping::setConfig V:1456-1469. Log-odds constants are
 * computed on the host in double exactly like O:79-87 and narrowed to float. Returns VDBM_ERR_BAD_CONFIG
 * for max_range < 0 (nothing changes) and for prob_miss > 0.5 / prob_hit < 0.5 (max_range and the
 * "configured" flag ARE already updated then, like the reference). */
int vdbm_set_config(vdbm_map* map, double max_range, double prob_hit, double prob_miss, double prob_thres_min,
                    double prob_thres_max);
/* out[6] = logodds hit, miss, thres_min, thres_max, max, min (O:179-199) */
int vdbm_get_logodds(vdbm_map* map, float* out6);
/* addInputSource V:1352-1375 (device part: the per-source update grid; max_range == 0 -> config max_range).
 * Worker threads and rate limiting stay in the host shim. */
int vdbm_source_add(vdbm_map* map, const char* source_id, double max_range);

/* ---- the hot path ------------------------------------------------------------------------- */
/* accumulateUpdate V:316-346 -> raycastPointCloud V:466-539 (castRayIntoGrid V:550-566, worldToIndex
 * V:612-631) into the source's device update grid. `points` is a HOST buffer of n records with the
 * pcl::PointXYZ layout (3 x f32 at the start of each `stride_bytes` record, normally 16); pinned memory is
 * copied asynchronously. No-op (VDBM_OK) when the source's max_range <= 0 (V:331). */
int vdbm_accumulate(vdbm_map* map, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes,
                    const double origin[3]);
/* Same, `points` already resident in device memory (used for the kernel-only benchmark leg and for callers
 * that produce clouds on the GPU). */
int vdbm_accumulate_device(vdbm_map* map, const char* source_id, const void* d_points, uint64_t n,
                           uint64_t stride_bytes, const double origin[3]);
/* raycastPointCloud V:466-472 with an explicit raycast_range, into the named source's update grid
 * (the reference passes an UpdateGridT::Accessor; here the grid is addressed by its source). */
int vdbm_raycast(vdbm_map* map, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes,
                 const double origin[3], double raycast_range);
/* integrateUpdate V:375-387: updateMap V:731-792 for every source in std::map key order, then fresh update
 * grids. The change grids are discarded like V:382 unless keep_change != 0 (then vdbm_change_export works). */
int vdbm_integrate(vdbm_map* map, int keep_change);
/* updateMap V:731-792 for ONE source whose update grid lives in another handle of the same device and resolution (holder
 * may also be `map` itself), then a fresh update grid there (V:384). The reference's one parallel axis is a thread per input
 * source, each raycasting into its own grid (V:316-346, 1383-1411); a handle is thread-compatible, so a host that wants the
 * sources to raycast concurrently gives every source a raycast-only handle (own stream, staging and counters,
 * map_capacity_leaves = 1) and integrates them into the map's handle in std::map key order with this call: the map's kernels
 * read the holder's grid in place, nothing is copied. Neither handle may be used by another thread during the call. With
 * keep_change the change records stay with the holder (vdbm_change_export(holder, source_id)). */
int vdbm_integrate_from(vdbm_map* map, vdbm_map* holder, const char* source_id, int keep_change);
/* insertPointCloud V:399-406 = accumulate + integrate. */
int vdbm_insert(vdbm_map* map, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes,
                const double origin[3]);
/* insertPointCloud V:399-406 as a PIPELINE stage: same result as vdbm_insert, but the call returns once the scan is
 * queued. The cloud is uploaded on a copy stream while the previous scan still computes, the whole chain (prep, sort,
 * DDA, compaction, updateMap) is queued without a host round trip in between (a one-thread guard kernel takes the
 * capacity decisions on the device), and the single host synchronisation of a scan happens when the NEXT call on the
 * handle - any entry point except vdbm_stats / vdbm_last_error - finishes it. Scans the queued path cannot take (tables
 * that must grow, long rays that need segmentation, several sources holding data, artificial areas, the first scans of a
 * map) run synchronously inside the call or are redone by the finishing call; results are identical either way. A status
 * belonging to a queued scan (e.g. VDBM_ERR_COORD_RANGE) is returned by the call that finishes it: another entry point
 * then returns that status instead of doing its own work (call it again), while vdbm_insert_async still queues its scan
 * and returns the earlier scan's VDBM_ERR_COORD_RANGE as a warning. points_on_device != 0:
 * `points` is device memory and must stay valid until the scan is finished; a host buffer is free when the call returns. */
int vdbm_insert_async(vdbm_map* map, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes,
                      const double origin[3], int points_on_device);
/* Start uploading the NEXT cloud (pinned host memory) on the copy stream while earlier work is still running; a following
 * vdbm_accumulate called with the same pointer, count and stride uses the uploaded copy instead of copying on the critical
 * path. The host buffer must stay untouched until that vdbm_accumulate returns. Used by callers that cannot use
 * vdbm_insert_async because something happens between raycast and update (the multi-GPU exchange, createUpdate). */
int vdbm_prefetch(vdbm_map* map, const void* points, uint64_t n, uint64_t stride_bytes);
/* finish the queued scan, if any (every other entry point does this implicitly) */
int vdbm_flush(vdbm_map* map);
/* updateMap(UpdateGridT::Ptr) V:731-792 on ONE source's accumulated update grid; the grid is emptied.
 * If change != NULL it receives the returned change grid (active = flag flipped, value = it was a hit). */
int vdbm_update_map(vdbm_map* map, const char* source_id, vdbm_leafset** change);

/* ---- grids in and out (remote mapping, host mirror) ---------------------------------------- */
/* The source's raw update grid (InputSource::update_grid V:97): origin, active mask, value mask per leaf. */
int vdbm_update_export(vdbm_map* map, const char* source_id, vdbm_leafset** out);
/* OR a host update grid into the source's device update grid (for updateMap() called with a grid that was
 * produced elsewhere, e.g. byteArrayToGrid V:1328-1338 on the receiving side). */
int vdbm_update_import(vdbm_map* map, const char* source_id, uint64_t n_leaves, const int32_t* origins /*[n][3]*/,
                       const uint64_t* active /*[n][8]*/, const uint64_t* value /*[n][8]*/);
/* Change grid kept by the last vdbm_integrate(map, 1) for that source. */
int vdbm_change_export(vdbm_map* map, const char* source_id, vdbm_leafset** out);
/* getGrid V:799: map leaves (origin, 512 x f32, active mask). dirty_only != 0: only leaves modified since the
 * previous vdbm_map_export call (what an eager/lazy host mirror needs); 0: all leaves. */
int vdbm_map_export(vdbm_map* map, int dirty_only, vdbm_leafset** out);
/* getGrid V:799 for a host grid that FOLLOWS the device map (the shim's mirror, tests/mapping.cpp:13-29 read through an
 * accessor taken before the insert): every leaf modified since the previous vdbm_map_mirror / vdbm_map_export(dirty) call is
 * handed to `sink` in chunks of at most chunk_leaves (0 = 16384); the device gather and the D2H copy of the following
 * chunks run while sink works on the current one, so the host-side merge hides behind PCIe. Chunks arrive in ascending
 * leaf_index = the leaf's slot in the device pool: fixed for the lifetime of the leaf and dense (a new leaf takes the
 * next index) until the pool restarts (vdbm_reset, vdbm_map_import with replace: vdbm_map_generation changes), so a
 * consumer can keep a table leaf_index -> host leaf instead of looking leaves up by origin. The arrays are valid only
 * during the call. sink returns non-zero to abort: the chunk of that call and every later one stay dirty. n_leaves (may be
 * NULL) receives the number of leaves delivered. */
typedef int (*vdbm_mirror_sink)(void* user, uint64_t n, const uint32_t* leaf_index /*[n]*/, const int32_t* origins /*[n][3]*/,
                                const float* values /*[n][512]*/, const uint64_t* active /*[n][8]*/);
int vdbm_map_mirror(vdbm_map* map, uint64_t chunk_leaves, vdbm_mirror_sink sink, void* user, uint64_t* n_leaves);
uint64_t vdbm_map_generation(const vdbm_map* map);
/* getMapSection<T> V:921-960 with extractSparseLeaf V:999-1011 / extractFullLeaf V:970-989 on an INCLUSIVE
 * index bounding box (createIndexBoundingBox V:857-871 is host maths and stays in the shim).
 * result_float = 0 -> UpdateGridT result (active + value masks), 1 -> GridT result (active + 512 f32). */
int vdbm_section(vdbm_map* map, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float,
                 vdbm_leafset** out);
/* ---- receiver side of remote mapping ---------------------------------------------- */
/* applyMapSectionUpdateGrid V:1058-1085 (without the optional OpenVDB morphology smoothing): deactivate every active
 * map voxel inside the inclusive index box [bbmin, bbmax] (the section's bb_min / bb_max metadata, floored), then
 * activate every active voxel of the section (missing leaves are created with background values). */
int vdbm_section_apply_update(vdbm_map* map, const int32_t bbmin[3], const int32_t bbmax[3], uint64_t n_leaves,
                              const int32_t* origins /*[n][3]*/, const uint64_t* active /*[n][8]*/);
/* applyMapSectionGrid V:1022-1047: every section leaf replaces the map leaf voxel for voxel (value + state).
 * replicate_tile_quirk != 0 also reproduces the reference's visits of the section tree's inactive background tiles
 * (cbeginValueAll): map.setValueOff(tile origin, 0) clears one voxel per missing leaf slot / 128^3 slot wherever the
 * map has a leaf. */
int vdbm_section_apply_grid(vdbm_map* map, uint64_t n_leaves, const int32_t* origins /*[n][3]*/, const uint64_t* active /*[n][8]*/,
                            const float* values /*[n][512]*/, int replicate_tile_quirk);
/* ---- remote-mapping deltas: createUpdate / applyUpdate -------------------------------------
 * BASELINE north_star names createUpdate/applyUpdate with a "reduction level"; the reference revision under
 * /root/reference has no such members (its wrapper ships whole grids through gridToByteArray V:1310-1338), so the levels
 * are defined here from the grids the reference DOES produce:
 *   level 0  the source's raw update grid (InputSource::update_grid V:97)        -> apply = updateMap V:731-792
 *   level 1  the change ("overwrite") grid updateMap returned V:733,772,780     -> apply = every active voxel forced
 *            occupied (value bit) / free with setNodeToOccupied / setNodeToFree O:118-129
 *   level 2  the REDUCED update: the end voxel of every ray of the source's last accumulate call (active; value = the
 *            ray was not range-clipped, V:533-536) + that scan's origin. castRayIntoGrid V:550-566 depends only on the
 *            two voxel indices, so the receiver re-raycasts the set from the origin and obtains the sender's update grid
 *            bit for bit: ~one voxel per ray crosses the wire instead of hundreds.
 * vdbm_update_create: level 0/2 must be called between accumulate and integrate, level 1 after vdbm_integrate(map, 1)
 * or vdbm_update_map. origin_out (may be NULL) receives the origin of the source's last accumulate. */
int vdbm_update_create(vdbm_map* map, const char* source_id, int level, vdbm_leafset** out, double origin_out[3]);
/* Apply such a grid to THIS map (library-owned scratch grid; no input source involved). origin is required for level 2.
 * If change != NULL it receives the change grid of the updateMap run (levels 0 and 2; empty for level 1). */
int vdbm_update_apply(vdbm_map* map, int level, uint64_t n_leaves, const int32_t* origins /*[n][3]*/, const uint64_t* active /*[n][8]*/,
                      const uint64_t* value /*[n][8]*/, const double origin[3], vdbm_leafset** change);

/* ---- direct map edits and artificial areas ---------------------------------------------------- */
/* addPointsToGrid V:431-447 (occupied != 0: value = max log-odds, active) / removePointsFromGrid V:413-429 (value = min
 * log-odds, inactive) for the voxel Coord::floor(point / resolution) of every point (host buffer, pcl::PointXYZ layout). */
int vdbm_points_set(vdbm_map* map, const void* points, uint64_t n, uint64_t stride_bytes, int occupied);
/* addArtificialAreas V:1175-1236: restoreMapIntegrity, then one wall per polygon edge (closing edge included), each wall
 * = castRayIntoGrid(start + (0,0,i), end + (0,0,i)) for i in [(int)(negative_height/res), (int)(positive_height/res)) into
 * the artificial-area grid. xyz holds sum(counts) world points (x, y, z). From then on every updateMap ends by forcing
 * those voxels active (V:785-789). */
int vdbm_artificial_areas_add(vdbm_map* map, uint64_t n_polygons, const uint32_t* counts, const double* xyz, double negative_height,
                              double positive_height);
/* addArtificialPolygon V:1198-1207 (closed != 0: every polyline also gets its closing edge) / addArtificialWall V:1217-1236
 * (closed == 0, counts[p] == 2): walls added to the artificial-area grid WITHOUT the restoreMapIntegrity that
 * addArtificialAreas runs first. */
int vdbm_artificial_walls_add(vdbm_map* map, uint64_t n_polylines, const uint32_t* counts, const double* xyz, double negative_height,
                              double positive_height, int closed);
/* restoreMapIntegrity V:1152-1166: active = (value > thres_max) for every artificial-area voxel, then the grid is cleared. */
int vdbm_map_integrity_restore(vdbm_map* map);
/* the artificial-area grid (m_artificial_area_grid V:1493) as a bool leaf set */
int vdbm_artificial_export(vdbm_map* map, vdbm_leafset** out);

/* loadMap V:263-284 / loadMapFromPCD V:295-307 (device part): the given leaves (OpenVDB layout) become map leaves, leaf for
 * leaf (values and active mask replace what the map held there). replace != 0 first forgets every map leaf, like the
 * reference replacing m_vdb_grid. */
int vdbm_map_import(vdbm_map* map, uint64_t n_leaves, const int32_t* origins, const uint64_t* active, const float* values, int replace);
/* castRayIntoGrid V:550-566 for n explicit rays given as voxel index pairs rays6[n][6] = {start xyz, end xyz}: every voxel
 * of the DDA from start to end (both included; nothing when start == end) becomes active in the named source's update
 * grid. Same fp64 stepping as the scan path. */
int vdbm_cast_index_rays(vdbm_map* map, const char* source_id, uint64_t n_rays, const int32_t* rays6);
/* Config::fast_mode V:1466: on != 0 makes accumulate / raycast / insert use castRayIntoGridFast V:577-602 instead of
 * castRayIntoGrid: nothing is cast while the map is empty (V:522); otherwise the ray is intersected with the map's node
 * topology (tools::VolumeRayIntersector<FloatGrid>::hits, restated: hierarchical DDA over the 4096^3 / 128^3 / 8^3 node levels),
 * a voxel DDA runs over every hit span and only voxels that are ACTIVE in the map become active in the update grid; the end
 * point rule V:533-536 is unchanged. The intersector the reference rebuilds after every integrateUpdate (V:386, 1436-1449) is
 * the map itself here: identical whenever the map was last changed by integrateUpdate. */
int vdbm_set_fast_mode(vdbm_map* map, int on);
/* raytrace V:675-721 for n rays: origins / directions are world vectors [n][3], max_lengths [n]; success[n] receives 1 when
 * the ray meets a node span of the map (VolumeRayIntersector::march), end_points [n][3] the world position of the voxel the
 * fine DDA stopped at (first active voxel after the span start, V:707-712) or origin + direction * length on failure. */
int vdbm_raytrace(vdbm_map* map, uint64_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* success,
                  double* end_points);
/* GridT::Accessor::getValue / isValueOn for one voxel (tests/mapping.cpp:27-29). */
int vdbm_probe(vdbm_map* map, const int32_t xyz[3], float* value, int32_t* active);

/* ---- leaf sets ------------------------------------------------------------------------------
 * A leaf set is read through the accessors below and released with vdbm_leafset_free(). Small sets own their
 * (pageable) memory. A large set (>= 1 MB) borrows the handle's persistent pinned staging buffer when it is free
 * (page-locking per call would cost more than the copy itself): free it BEFORE the next large export on the same
 * handle if you want that export to be fast too (otherwise the next one simply uses pageable memory of its own). A set
 * that still borrows the buffer when the handle is destroyed takes ownership of it. */
uint64_t vdbm_leafset_size(const vdbm_leafset* s);
const int32_t* vdbm_leafset_origins(const vdbm_leafset* s); /* [n][3] */
const uint64_t* vdbm_leafset_active(const vdbm_leafset* s); /* [n][8] */
const uint64_t* vdbm_leafset_valmask(const vdbm_leafset* s); /* [n][8] or NULL (float results) */
const float* vdbm_leafset_values(const vdbm_leafset* s);    /* [n][512] or NULL (bool results) */
void vdbm_leafset_free(vdbm_leafset* s);

/* ---- multi-GPU: map sharded by leaf key, rays split across ranks -------------------- */
/* Owner rank of a leaf (origin must be a multiple of 8): mix64(morton-brick(leaf)) % n_ranks. Pure function. */
int32_t vdbm_leaf_owner(const int32_t origin[3], int32_t n_ranks);
/* Ownership plan of the sharded map (replaces nothing in the reference: V:375-387 integrates ONE map; here every rank
 * integrates its shard). mode 0 (default): the hash of vdbm_leaf_owner. mode 1: azimuth sectors around the leaf column
 * center_leaf_xy (leaf coordinates = voxel >> 3): rank r owns the leaves whose centre direction, measured as a diamond
 * angle in [0, 4) (dy/(|dx|+|dy|) per quadrant), lies in [bounds[r], bounds[r+1]); the last sector wraps around. Split the
 * rays of a scan by the same bounds around the sensor and nearly every touched leaf stays on the rank that touched it
 * (vdb_mapping_b200/dist.py plans the bounds from the first scan). Every rank must set the same plan; it cannot change
 * while the map holds leaves. */
int vdbm_shard_plan_set(vdbm_map* map, int32_t mode, int32_t n_ranks, const int32_t center_leaf_xy[2], const double* bounds);
/* Ray split on the device: with n_ranks > 1 every accumulate / raycast / insert of this handle casts only the points whose
 * direction around the SCAN ORIGIN (diamond angle of (x - origin.x, y - origin.y)) lies in sector `rank` of `bounds` (same
 * convention as the shard plan; invalid points belong to the sector that holds angle 0). Give every rank of a box the whole
 * cloud and the same bounds and each ray is cast exactly once. n_ranks <= 1 switches the filter off. */
int vdbm_ray_sector_set(vdbm_map* map, int32_t n_ranks, int32_t rank, const double* bounds);
/* owner under the handle's plan (mode 0: same as vdbm_leaf_owner) */
int32_t vdbm_leaf_owner_planned(vdbm_map* map, const int32_t origin[3], int32_t n_ranks);
/* Order-independent parity witness of a (sharded) map: out2[0] = sum over the leaves of this handle of a 64-bit hash of
 * (leaf key, active mask, 512 value bit patterns), out2[1] = number of leaves. Summed over all ranks (mod 2^64) it equals
 * the checksum of the same scans integrated on one GPU (getGrid() V:799 compared without moving the map). */
int vdbm_map_checksum(vdbm_map* map, uint64_t out2[2]);
/* Bin the source's accumulated update leaves by owner into a device buffer of 136-byte records
 * {uint64 key, uint64 active[8], uint64 value[8]} grouped by rank; counts[n_ranks] (host) receives the group
 * sizes; the update grid is emptied. *d_records stays valid until the next call on this handle. */
int vdbm_update_partition(vdbm_map* map, const char* source_id, int32_t n_ranks, uint64_t* counts,
                          const void** d_records);
/* OR device-resident records (as produced by vdbm_update_partition on any rank) into the source's grid. */
int vdbm_update_import_device(vdbm_map* map, const char* source_id, const void* d_records, uint64_t n_records);

/* ---- multi-GPU exchange over peer memory (one process per GPU on one NVLink/NVSwitch box) -------------------
 * Fused bin-and-send: ONE kernel bins the source's accumulated update leaves by owner rank and stores each 136-byte
 * record straight into the owner's inbox (peer memory mapped with CUDA IPC, stores travel over NVLink), then
 * publishes (epoch, count) words to the peers. The receiving side waits on those words ON THE DEVICE, so no counts
 * cross the host and no collective library call is on the data path.
 *   vdbm_exchange_create : allocate this rank's inbox (2 parities x n_ranks sender regions x capacity records) and
 *                          return its IPC handles (VDBM_IPC_HANDLE_BYTES bytes) for an all-gather by the caller
 *   vdbm_exchange_connect: map every peer's inbox from the gathered handles (n_ranks x VDBM_IPC_HANDLE_BYTES)
 *   vdbm_update_push     : bin the source's update leaves by owner and send the FOREIGN ones (zeroed locally); the leaves
 *                          this rank owns stay in its grid; asynchronous
 *   vdbm_update_pull     : wait for all senders of this epoch, OR their records into the source's grid */
#define VDBM_IPC_HANDLE_BYTES 128
int vdbm_exchange_create(vdbm_map* map, int32_t rank, int32_t n_ranks, uint64_t capacity_records_per_sender, void* handles_out);
int vdbm_exchange_connect(vdbm_map* map, const void* all_handles);
/* vdbm_exchange_connect for handles that live in THIS process (one per GPU): peers[r] is the handle created with rank r
 * (peers[own rank] is ignored); peer access is enabled and the inboxes are addressed directly, no CUDA IPC involved. */
int vdbm_exchange_connect_peers(vdbm_map* map, vdbm_map* const* peers);
int vdbm_update_push(vdbm_map* map, const char* source_id);
int vdbm_update_pull(vdbm_map* map, const char* source_id);
/* vdbm_update_pull followed by vdbm_integrate(map, 0) with one host synchronisation instead of two: the update kernels are
 * queued behind a device-side capacity guard (see vdbm_insert_async); when it refuses, the two calls run synchronously. */
int vdbm_update_pull_integrate(vdbm_map* map, const char* source_id);
/* device times (ms, CUDA events) of the last push / pull pair: out[0] bin+send kernels, out[1] wait for the peers'
 * epoch words (includes their raycast skew), out[2] import + leaf compaction */
int vdbm_exchange_timings(vdbm_map* map, float* out3);

/* ---- one process, several GPUs (SURVEY.md 8e from the C ABI: a ROS node is ONE process) ------------------------------
 * A vdbm_group is a sharded map on n_devices GPUs of one NVLink / NVSwitch box, driven from one process: one vdbm_map per
 * device (each with its own worker thread), inboxes wired with direct peer pointers, the same fused bin-and-send exchange as
 * the multi-process path. integrateUpdate V:375-387 integrates ONE map; here every shard integrates the leaves it owns and
 * the union of the shards is that map (vdbm_group_checksum == vdbm_map_checksum of the same scans on one GPU).
 *   vdbm_group_insert : insertPointCloud V:399-406 across the shards: every shard receives the whole cloud (host buffer, pinned
 *                       for speed) and casts the rays of its azimuth sector (vdbm_ray_sector_set), foreign update leaves are
 *                       pushed to their owners over NVLink, every shard runs updateMap on its own leaves. The sector plan
 *                       (ray sectors of equal raycast cost, ownership sectors of equal owned leaves) is cut from the FIRST scan
 *                       of an empty map (dry-run raycast on shard 0) unless vdbm_group_plan_set supplied one.
 *   vdbm_group_shard  : the per-device handle, for everything that is read shard by shard (vdbm_map_export, vdbm_section,
 *                       vdbm_probe, vdbm_stats ...): the shards hold disjoint leaf sets.
 * params->stream must be NULL (every shard owns its stream); params->device is ignored. inbox_capacity_records: update
 * leaves one shard may send to one other shard per scan (0 = 1 << 19). Calls on a group must come from one thread at a time. */
typedef struct vdbm_group vdbm_group;
int vdbm_group_create(const vdbm_params* params, int32_t n_devices, const int32_t* devices, uint64_t inbox_capacity_records, vdbm_group** out);
void vdbm_group_destroy(vdbm_group* group);
int32_t vdbm_group_size(const vdbm_group* group);
vdbm_map* vdbm_group_shard(vdbm_group* group, int32_t i);
int vdbm_group_set_config(vdbm_group* group, double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max);
int vdbm_group_source_add(vdbm_group* group, const char* source_id, double max_range);
int vdbm_group_reset(vdbm_group* group);
int vdbm_group_insert(vdbm_group* group, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3]);
/* explicit sector plan (diamond angles, see vdbm_shard_plan_set / vdbm_ray_sector_set); only while the map is empty */
int vdbm_group_plan_set(vdbm_group* group, const int32_t center_leaf_xy[2], const double* ray_bounds, const double* ownership_bounds);
int vdbm_group_plan_get(vdbm_group* group, int32_t center_leaf_xy[2], double* ray_bounds, double* ownership_bounds);
/* sum of the shard checksums / leaf counts (mod 2^64) */
int vdbm_group_checksum(vdbm_group* group, uint64_t out2[2]);
/* counters summed over the shards (times: max over the shards) */
int vdbm_group_stats(vdbm_group* group, vdbm_stats_t* out);
const char* vdbm_group_last_error(vdbm_group* group);

/* ---- diagnostics ---------------------------------------------------------------------------- */
int vdbm_stats(vdbm_map* map, vdbm_stats_t* out);
/* How the scans given to vdbm_insert_async travelled so far: out4 = { finished on the queued path, redone synchronously after
 * the device-side guard refused, run synchronously because they were not eligible, queued with their raycast half
 * overlapping the previous scan's updateMap }. Like vdbm_stats it does not finish a queued scan. */
int vdbm_pipeline_counts(vdbm_map* map, uint64_t out4[4]);
const char* vdbm_last_error(vdbm_map* map);
int vdbm_abi_version(void);
/* wait for all work queued on the handle's stream */
int vdbm_synchronize(vdbm_map* map);
/* pinned host memory helpers for callers that want truly asynchronous uploads */
void* vdbm_host_alloc(size_t bytes);
void vdbm_host_free(void* p);

#if defined(__GNUC__)
#pragma GCC visibility pop
#endif
#ifdef __cplusplus
}
#endif
#endif /* VDBM_B200_H_INCLUDED */
