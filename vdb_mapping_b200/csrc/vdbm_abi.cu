// vdbm_abi.cu — host side of libvdbm_b200.so: the C ABI declared in include/vdbm_b200.h.
// Owns device memory (update-leaf hashes, map hash + leaf pool), sequencing of the kernels in
// vdbm_kernels.cu, growth, and the export staging. No CPU compute path exists here: every grid
// operation is a kernel launch; without a CUDA device vdbm_create fails.
#include "../../include/vdbm_b200.h"
#include "vdbm_device.cuh"

#include <algorithm>
#include <chrono>
#include <cmath>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

using namespace vdbm;

struct vdbm_leafset
{
  uint64_t n        = 0;
  int32_t* origins  = nullptr;
  uint64_t* active  = nullptr;
  uint64_t* valmask = nullptr;
  float* values     = nullptr;
  bool pinned       = false;
  vdbm_map* lender  = nullptr; // non-null: the arrays live in the handle's persistent pinned staging buffer
  void* orphan_stage = nullptr; // the handle was destroyed while this set borrowed its staging buffer: the set owns it now
};

namespace {

struct Source
{
  std::string id;
  double max_range = 0.0;
  UpdateGrid g{};
  uint32_t cap       = 0;         // brick slots
  uint64_t prev_visits     = 0;   // voxel marks and longest ray of this source's previous scan: the segment length of
  uint32_t prev_max_visits = 0;   // the next scan is planned from them (consecutive scans of a sensor look alike)
  uint64_t prev_updates    = 0;   // distinct voxels of the previous scan: visits / updates = how heavily its rays overlap
  uint32_t n_bricks  = 0;         // host copies, valid after every synchronising call
  uint32_t n_entries = 0;         // touched leaves (compact list is always rebuilt after a grid write)
  // second update grid of the scan pipeline (vdbm_insert_async with overlap): the raycast of scan k+1 marks into one grid
  // while updateMap of scan k consumes the other; swapped with (g, cap, n_bricks, n_entries) by swapParity(). Lazy.
  UpdateGrid g_alt{};
  uint32_t cap_alt = 0, n_bricks_alt = 0, n_entries_alt = 0;
  LeafRecord* d_change = nullptr; // change records of the last update (device)
  uint32_t change_cap  = 0;
  uint32_t n_change    = 0;
};

} // namespace

struct vdbm_map
{
  vdbm_params params{};
  int device          = 0;
  cudaStream_t stream = nullptr;
  bool own_stream     = false;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr;
  std::string last_error;

  // config (VDBMapping.hpp:1456-1469, OccupancyVDBMapping.hpp:59-89)
  double max_range = 0.0;
  bool config_set  = false;
  LogOdds lo{};

  // device state
  Counters* d_ctr = nullptr;
  Counters* h_ctr = nullptr; // pinned
  MapTable mt{};
  uint32_t hcap     = 0;
  uint32_t n_leaves = 0; // host copy
  uint32_t* d_map_counters = nullptr; // [0] n_leaves
  uint32_t* h_small        = nullptr; // pinned scratch (>= 64 words)
  std::map<std::string, std::unique_ptr<Source> > sources; // std::map order == integrateUpdate order (V:380)
  std::unique_ptr<Source> scratch;    // library-owned bool grid for applyUpdate / reduced updates / point edits (lazy)
  std::unique_ptr<Source> artificial; // m_artificial_area_grid V:132,1493 (lazy; survives resetMap like the reference's)
  // end-voxel records of the last accumulate (the "reduced", level-2 update of that scan)
  int4* d_ends       = nullptr; // [rays_cap]
  Source* ends_src   = nullptr;
  uint64_t ends_n    = 0;
  double ends_origin[3] = {0, 0, 0};

  // staging
  uint8_t* d_points = nullptr;
  size_t points_cap = 0;
  RayRec* d_rays    = nullptr;
  SegRec* d_segs    = nullptr; // seg_cap segments (first rays_cap: one per ray)
  uint32_t* d_long  = nullptr; // 2 x rays_cap u32: long-ray list, extra-segment base per ray
  size_t seg_cap    = 0;
  uint32_t* d_sort  = nullptr; // 4 x seg_cap u32: keys in/out, idx in/out
  void* d_sort_tmp  = nullptr;
  size_t sort_tmp_bytes = 0;
  size_t rays_cap   = 0;
  uint64_t* d_near     = nullptr; // privatised near-field brick copies (always left zeroed)
  uint32_t* d_resolved = nullptr; // K2a output: map leaf index per touched update leaf
  size_t resolved_cap  = 0;
  LeafRecord* d_part = nullptr; // partition output
  size_t part_cap    = 0;
  int dda_grid       = 0;

  ShardPlan shard{}; // ownership of map leaves across ranks (mode 0 = hash; vdbm_shard_plan_set)
  int32_t sector_n = 0, sector_rank = 0; // ray split on the device (vdbm_ray_sector_set); sector_n <= 1: off
  double sector_bounds[kMaxRanks] = {};

  // fast_mode / raytrace (V:577-602, V:675-721): node levels above the leaves as coarse key sets, built lazily
  bool fast_mode = false; // Config::fast_mode V:1466
  CoarseSets coarse{};
  uint32_t coarse_cap1 = 0, coarse_cap2 = 0;
  uint32_t coarse_built = 0; // map leaves [0, coarse_built) are in the sets
  int32_t* d_bbox = nullptr; // [6] active bounding box of the map (min xyz, max xyz inclusive)

  // peer-memory exchange (multi-GPU)
  struct Exchange
  {
    ExchangePeers px{};
    uint64_t* inbox = nullptr;            // own inbox: masks [2*n_ranks][cap][16] then keys [2*n_ranks][cap]
    unsigned long long* ctrl = nullptr;   // own ctrl  [2][n_ranks]
    uint32_t* d_cursors = nullptr;        // [kMaxRanks] send cursors
    uint32_t* d_counts  = nullptr;        // [kMaxRanks] received counts of the current epoch
    std::vector<void*> opened;            // peer mappings to close
    uint32_t epoch = 0;
    bool created = false, connected = false;
    cudaEvent_t ev[5] = {nullptr, nullptr, nullptr, nullptr, nullptr}; // push begin/end, wait end, import end
    float ms[3] = {0, 0, 0};
  } ex;

  // asynchronous scan pipeline (vdbm_insert_async): at most one scan in flight
  struct Pending
  {
    bool active = false;
    Source* src = nullptr;
    uint64_t n = 0, stride = 0;
    const uint8_t* d_pts = nullptr; // the device-resident cloud of the scan (own staging buffer or the caller's)
    double origin[3] = {0, 0, 0};
    Counters before{};
    uint64_t rays_before = 0;
  } pending;
  cudaStream_t copy_stream = nullptr;
  cudaEvent_t ev_copy = nullptr, ev_done = nullptr, ev3 = nullptr;
  // overlap of scan k+1's raycast with scan k's updateMap: a second counter block, a second set of
  // timing events and a second stream for the update half of a queued scan. (d_ctr, h_ctr, ev0..ev3) always name the
  // CURRENT set; swapParity() exchanges them with the *_alt set together with the source's two update grids.
  Counters* d_ctr_alt = nullptr;
  Counters* h_ctr_alt = nullptr; // pinned
  cudaEvent_t ev_alt[4] = {nullptr, nullptr, nullptr, nullptr};
  cudaStream_t update_stream = nullptr;
  cudaEvent_t ev_ray = nullptr; // raycast half of the queued scan finished (and everything queued on `stream` before its update half)
  bool overlap = true;          // VDBM_OVERLAP=0 switches it off (experiments)
  bool ends_clobbered = false;
  bool ray_in_flight = false;   // the next scan's raycast half is already queued while finishPending() works on the previous scan
  int dda_grid_overlap = 0;     // DDA CTAs when the raycast half runs next to an update (fewer per SM: leaves room for the update CTAs)
  uint64_t async_overlapped = 0;
  uint8_t* d_points_async[2] = {nullptr, nullptr};
  size_t points_async_cap[2] = {0, 0};
  int async_buf              = 0;
  struct Prefetch { const void* host = nullptr; uint64_t n = 0, stride = 0; int buf = 0; bool valid = false; } prefetch;
  uint32_t async_expect      = 0; // touched leaves of the last finished scan: sizes the next deferred update
  uint64_t async_fast = 0, async_redone = 0, async_sync = 0; // scans that took the queued path / were redone / went synchronous

  // persistent pinned staging for large exports (page-locking hundreds of MB per call costs more than the copy)
  void* h_stage      = nullptr;
  size_t h_stage_cap = 0;
  bool h_stage_lent  = false;
  vdbm_leafset* h_stage_borrower = nullptr;

  // vdbm_map_mirror: rotating chunk buffers (device gather target + pinned host landing zone) and their "copied" events
  struct Mirror
  {
    static constexpr int kBufs = 3;
    uint8_t* d_buf[kBufs] = {nullptr, nullptr, nullptr};
    uint8_t* h_buf[kBufs] = {nullptr, nullptr, nullptr};
    cudaEvent_t ev[kBufs] = {nullptr, nullptr, nullptr};
    uint32_t cap_leaves   = 0;
  } mirror;
  uint64_t generation = 0; // bumped whenever the leaf pool restarts (vdbm_reset, vdbm_map_import with replace): leaf indices are void

  vdbm_stats_t stats{};
  Counters base{}; // counters at the last reset, to keep cumulative numbers across device counter resets
};

namespace {

#define CU_TRY(m, expr)                                                                                         \
  do                                                                                                            \
  {                                                                                                             \
    cudaError_t e__ = (expr);                                                                                   \
    if (e__ != cudaSuccess)                                                                                     \
    {                                                                                                           \
      (m)->last_error = std::string(#expr) + ": " + cudaGetErrorString(e__);                                    \
      return (e__ == cudaErrorMemoryAllocation) ? VDBM_ERR_OUT_OF_MEMORY : VDBM_ERR_CUDA;                       \
    }                                                                                                           \
  } while (0)

int fail(vdbm_map* m, int code, const std::string& msg)
{
  m->last_error = msg;
  return code;
}

uint32_t nextPow2(uint64_t v)
{
  uint64_t p = 1;
  while (p < v) p <<= 1;
  return uint32_t(std::min<uint64_t>(p, 1ull << 31));
}

// ---- update grid (brick hash) ------------------------------------------------------------------------
int allocUpdateGrid(vdbm_map* m, UpdateGrid& g, uint32_t cap)
{
  const size_t brick_bytes = size_t(kBrickLeaves) * 64;
  CU_TRY(m, cudaMalloc(&g.bkeys, size_t(cap) * 8));
  // one extra "trash" brick at index cap: where marks go when the hash is full (scan is replayed after growing)
  CU_TRY(m, cudaMalloc(&g.act, (size_t(cap) + 1) * brick_bytes));
  CU_TRY(m, cudaMalloc(&g.val, (size_t(cap) + 1) * brick_bytes));
  CU_TRY(m, cudaMalloc(&g.btouched, size_t(cap) * 4));
  CU_TRY(m, cudaMalloc(&g.entries, size_t(cap) * kBrickLeaves * 4));
  CU_TRY(m, cudaMalloc(&g.counters, 8));
  g.cap_mask = cap - 1;
  CU_TRY(m, cudaMemsetAsync(g.bkeys, 0xFF, size_t(cap) * 8, m->stream));
  CU_TRY(m, cudaMemsetAsync(g.act, 0, (size_t(cap) + 1) * brick_bytes, m->stream));
  CU_TRY(m, cudaMemsetAsync(g.val, 0, (size_t(cap) + 1) * brick_bytes, m->stream));
  CU_TRY(m, cudaMemsetAsync(g.counters, 0, 8, m->stream));
  return VDBM_OK;
}
void freeUpdateGrid(UpdateGrid& g)
{
  cudaFree(g.bkeys); cudaFree(g.act); cudaFree(g.val); cudaFree(g.btouched); cudaFree(g.entries); cudaFree(g.counters);
  g = UpdateGrid{};
}

// read the grid counters (bricks, entries) to the host; synchronises
int readGridCounters(vdbm_map* m, Source& s)
{
  CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s.g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  s.n_bricks  = m->h_small[8];
  s.n_entries = m->h_small[9];
  return VDBM_OK;
}

// double the brick table, move the occupied bricks, rebuild the leaf list
int growUpdateGrid(vdbm_map* m, Source& s)
{
  if (s.cap >= (1u << 22)) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "update brick hash cannot grow further");
  const uint32_t new_cap = s.cap * 2;
  UpdateGrid ng{};
  int rc = allocUpdateGrid(m, ng, new_cap);
  if (rc) return rc;
  launchRehashUpdate(s.g, std::min(s.n_bricks, s.cap), ng, m->d_ctr, m->stream);
  launchCompactLeaves(ng, m->stream);
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  freeUpdateGrid(s.g);
  s.g   = ng;
  s.cap = new_cap;
  m->stats.update_capacity = std::max(m->stats.update_capacity, s.cap * uint32_t(kBrickLeaves));
  return readGridCounters(m, s);
}

// drop everything accumulated in the grid (masks zeroed through the leaf list, bricks forgotten)
int clearUpdateGrid(vdbm_map* m, Source& s)
{
  launchClearEntries(s.g, s.n_entries, m->stream);
  launchResetBricks(s.g, s.n_bricks, m->stream);
  CU_TRY(m, cudaGetLastError());
  s.n_bricks = s.n_entries = 0;
  return VDBM_OK;
}

// ---- map ---------------------------------------------------------------------------------------------
int allocMapHash(vdbm_map* m, uint32_t hcap)
{
  CU_TRY(m, cudaMalloc(&m->mt.hkeys, size_t(hcap) * 8));
  CU_TRY(m, cudaMalloc(&m->mt.hvals, size_t(hcap) * 4));
  CU_TRY(m, cudaMemsetAsync(m->mt.hkeys, 0xFF, size_t(hcap) * 8, m->stream));
  m->mt.hcap_mask = hcap - 1;
  m->hcap         = hcap;
  return VDBM_OK;
}

int allocMapPool(vdbm_map* m, MapTable& t, uint32_t cap)
{
  CU_TRY(m, cudaMalloc(&t.leaf_keys, size_t(cap) * 8));
  CU_TRY(m, cudaMalloc(&t.leaf_mask, size_t(cap) * 64));
  CU_TRY(m, cudaMalloc(&t.leaf_vals, size_t(cap) * 2048));
  CU_TRY(m, cudaMalloc(&t.leaf_dirty, size_t(cap) * 4));
  CU_TRY(m, cudaMemsetAsync(t.leaf_dirty, 0, size_t(cap) * 4, m->stream));
  t.pool_cap = cap;
  return VDBM_OK;
}
void freeMapPool(MapTable& t)
{
  cudaFree(t.leaf_keys); cudaFree(t.leaf_mask); cudaFree(t.leaf_vals); cudaFree(t.leaf_dirty);
}

// make room for `extra` more leaves (pool) and keep the hash load factor <= 0.5
int ensureMapCapacity(vdbm_map* m, uint64_t extra)
{
  const uint64_t need = uint64_t(m->n_leaves) + extra;
  if (need > m->mt.pool_cap)
  {
    uint32_t new_cap = nextPow2(std::max<uint64_t>(need, uint64_t(m->mt.pool_cap) * 2));
    MapTable nt      = m->mt;
    int rc           = allocMapPool(m, nt, new_cap);
    if (rc) return rc;
    const size_t n = m->n_leaves;
    if (n)
    {
      CU_TRY(m, cudaMemcpyAsync(nt.leaf_keys, m->mt.leaf_keys, n * 8, cudaMemcpyDeviceToDevice, m->stream));
      CU_TRY(m, cudaMemcpyAsync(nt.leaf_mask, m->mt.leaf_mask, n * 64, cudaMemcpyDeviceToDevice, m->stream));
      CU_TRY(m, cudaMemcpyAsync(nt.leaf_vals, m->mt.leaf_vals, n * 2048, cudaMemcpyDeviceToDevice, m->stream));
      CU_TRY(m, cudaMemcpyAsync(nt.leaf_dirty, m->mt.leaf_dirty, n * 4, cudaMemcpyDeviceToDevice, m->stream));
    }
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    freeMapPool(m->mt);
    m->mt = nt;
  }
  if (need * 2 > m->hcap)
  {
    uint32_t new_h = nextPow2(need * 4);
    cudaFree(m->mt.hkeys);
    cudaFree(m->mt.hvals);
    int rc = allocMapHash(m, new_h);
    if (rc) return rc;
    launchRehashMap(m->mt, m->n_leaves, m->d_ctr, m->stream);
  }
  return VDBM_OK;
}

// D2H of the counter block + map counters, synchronising the stream
// cumulative statistics = counters at the last reset + BOTH device counter blocks (host copies; the block that is not
// current only changes through queued scans, whose finishing call refreshes its host copy)
void publishCumulative(vdbm_map* m)
{
  const Counters& c = *m->h_ctr;
  const Counters& o = *m->h_ctr_alt;
  m->stats.nan_skipped   = m->base.nan_skipped + c.nan_skipped + o.nan_skipped;
  m->stats.clipped       = m->base.clipped + c.clipped + o.clipped;
  m->stats.visits        = m->base.visits + c.visits + o.visits;
  m->stats.voxel_updates = m->base.voxel_updates + c.voxel_updates + o.voxel_updates;
  m->stats.state_changes = m->base.state_changes + c.state_changes + o.state_changes;
  m->stats.new_leaves    = m->base.new_leaves + c.new_leaves + o.new_leaves;
  m->stats.map_leaves    = m->n_leaves;
  m->stats.map_capacity  = m->mt.pool_cap;
  m->stats.gpu_launches  = launchCount();
}

int syncCounters(vdbm_map* m)
{
  CU_TRY(m, cudaMemcpyAsync(m->h_ctr, m->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(m->h_small, m->d_map_counters, 4, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  m->n_leaves = m->h_small[0];
  publishCumulative(m);
  return VDBM_OK;
}

Source* findSource(vdbm_map* m, const char* id)
{
  if (!id) return nullptr;
  auto it = m->sources.find(id);
  return it == m->sources.end() ? nullptr : it->second.get();
}

// host worldToIndex for the sensor origin (VDBMapping.hpp:612-631), same expression order as the reference
bool originIndex(double res, const double o[3], int32_t out[3])
{
  const double inv = 1.0 / res;
  for (int k = 0; k < 3; ++k)
  {
    double c = o[k];
    if (std::fmod(c, res)) c = c + (res / 2.0);
    const double fl = std::floor(c * inv);
    if (!(std::fabs(fl) < double(kVoxelLimit))) return false;
    out[k] = int32_t(fl);
  }
  return true;
}

// Rays of a depth camera overlap heavily (dozens of visits per distinct voxel): then the DDA tests a mask word before it
// issues the RED. Decided from the source's previous scan; VDBM_DDA_TBS=0/1 overrides (tests, experiments).
bool testBeforeSet(const Source& s)
{
  if (const char* e = getenv("VDBM_DDA_TBS")) return atoi(e) != 0;
  return s.prev_updates != 0 && s.prev_visits > 8 * s.prev_updates;
}

// ---- fast_mode / raytrace support: coarse node sets + active bounding box of the CURRENT map ----------------------------
// (the reference rebuilds its VolumeRayIntersector after every integrateUpdate V:386; here the structures follow the map
// lazily: new leaves are appended to the sets, the bounding box is re-reduced - one 64 B/leaf pass - on every use)
int allocCoarse(vdbm_map* m, uint32_t cap1, uint32_t cap2)
{
  cudaFree(m->coarse.k1); cudaFree(m->coarse.k2);
  m->coarse.k1 = m->coarse.k2 = nullptr;
  CU_TRY(m, cudaMalloc(&m->coarse.k1, size_t(cap1) * 8));
  CU_TRY(m, cudaMalloc(&m->coarse.k2, size_t(cap2) * 8));
  if (!m->coarse.counts) CU_TRY(m, cudaMalloc(&m->coarse.counts, 8));
  if (!m->d_bbox) CU_TRY(m, cudaMalloc(&m->d_bbox, 6 * sizeof(int32_t)));
  m->coarse_cap1  = cap1;
  m->coarse_cap2  = cap2;
  m->coarse.mask1 = cap1 - 1;
  m->coarse.mask2 = cap2 - 1;
  m->coarse_built = ~0u;
  return VDBM_OK;
}

// m->n_leaves must be current (every entry point that changes the map ends with syncCounters)
int ensureCoarse(vdbm_map* m)
{
  if (!m->coarse.k1)
  {
    int rc = allocCoarse(m, 1u << 14, 1u << 10);
    if (rc) return rc;
  }
  for (int attempt = 0; attempt < 24; ++attempt)
  {
    if (m->coarse_built == ~0u || m->coarse_built > m->n_leaves)
    {
      CU_TRY(m, cudaMemsetAsync(m->coarse.k1, 0xFF, size_t(m->coarse_cap1) * 8, m->stream));
      CU_TRY(m, cudaMemsetAsync(m->coarse.k2, 0xFF, size_t(m->coarse_cap2) * 8, m->stream));
      CU_TRY(m, cudaMemsetAsync(m->coarse.counts, 0, 8, m->stream));
      m->coarse_built = 0;
    }
    if (m->coarse_built < m->n_leaves) launchCoarseInsert(m->mt, m->coarse_built, m->n_leaves, m->coarse, m->d_ctr, m->stream);
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 16, m->coarse.counts, 8, cudaMemcpyDeviceToHost, m->stream));
    int rc = syncCounters(m);
    if (rc) return rc;
    m->coarse_built = m->n_leaves;
    const bool overflow = (m->h_ctr->flags & kFlagUpdateOverflow) != 0;
    const bool crowded1 = uint64_t(m->h_small[16]) * 2 > m->coarse_cap1, crowded2 = uint64_t(m->h_small[17]) * 2 > m->coarse_cap2;
    if (!overflow && !crowded1 && !crowded2) break;
    if (overflow)
    {
      m->h_ctr->flags &= ~kFlagUpdateOverflow;
      CU_TRY(m, cudaMemcpyAsync(&m->d_ctr->flags, &m->h_ctr->flags, sizeof(unsigned), cudaMemcpyHostToDevice, m->stream));
    }
    rc = allocCoarse(m, (overflow || crowded1) ? m->coarse_cap1 * 2 : m->coarse_cap1, (overflow || crowded2) ? m->coarse_cap2 * 2 : m->coarse_cap2);
    if (rc) return rc;
  }
  // RootNode::evalActiveBoundingBox(bbox, false): leaves that hold an active voxel
  const int32_t init[6] = {INT32_MAX, INT32_MAX, INT32_MAX, INT32_MIN, INT32_MIN, INT32_MIN};
  CU_TRY(m, cudaMemcpyAsync(m->d_bbox, init, sizeof(init), cudaMemcpyHostToDevice, m->stream));
  launchActiveBBox(m->mt, m->n_leaves, m->d_bbox, m->stream);
  return VDBM_OK;
}

// ---- the raycast (K0 + K1) on device-resident points ---------------------------------------------------
int raycastDevice(vdbm_map* m, Source& s, const uint8_t* d_points, uint64_t n, uint64_t stride, const double origin[3], double range,
                  bool index_mode = false)
{
  m->stats.rays += n;
  if (!index_mode && m->ends_src == &s) m->ends_src = nullptr; // the previous scan's reduced update is gone
  m->stats.last_visits = 0;
  if (n == 0) return VDBM_OK;
  if (n > 0xFFFFFFF0ull) return fail(m, VDBM_ERR_INVALID_ARG, "more than 2^32 points in one cloud");
  // VDBMapping.hpp:505-510: a NaN origin skips every point
  if (std::isnan(origin[0]) || std::isnan(origin[1]) || std::isnan(origin[2]) || std::isinf(origin[0]) || std::isinf(origin[1]) ||
      std::isinf(origin[2]))
  {
    m->base.nan_skipped += n;
    m->stats.nan_skipped += n;
    return VDBM_OK;
  }
  RaycastArgs a{};
  a.points = d_points;
  a.n      = n;
  a.stride = uint32_t(stride);
  for (int k = 0; k < 3; ++k) a.origin[k] = origin[k];
  if (!originIndex(m->params.resolution, origin, a.origin_idx)) return fail(m, VDBM_ERR_COORD_RANGE, "sensor origin outside the +-2^23 voxel range");
  a.range      = range;
  a.resolution = m->params.resolution;
  a.half_res   = m->params.resolution / 2.0;
  a.inv_res    = 1.0 / m->params.resolution;
  if (m->rays_cap < n)
  {
    cudaFree(m->d_rays); cudaFree(m->d_segs); cudaFree(m->d_long); cudaFree(m->d_sort); cudaFree(m->d_sort_tmp); cudaFree(m->d_ends);
    m->d_rays = nullptr; m->d_segs = nullptr; m->d_long = nullptr; m->d_sort = nullptr; m->d_sort_tmp = nullptr; m->d_ends = nullptr;
    m->rays_cap = 0;
    m->ends_src = nullptr;
    // grow geometrically: cloud sizes fluctuate from scan to scan and every reallocation synchronises the device
    const size_t rc = size_t(n) + n / 4 + 4096;
    // one segment per ray + a pool of extra segments for long rays (a ray that finds the pool empty stays whole)
    const size_t seg_cap = rc + 4 * rc + 65536;
    CU_TRY(m, cudaMalloc(&m->d_rays, rc * sizeof(RayRec)));
    CU_TRY(m, cudaMalloc(&m->d_ends, rc * sizeof(int4)));
    CU_TRY(m, cudaMalloc(&m->d_segs, seg_cap * sizeof(SegRec)));
    CU_TRY(m, cudaMalloc(&m->d_long, rc * 2 * sizeof(uint32_t)));
    CU_TRY(m, cudaMalloc(&m->d_sort, seg_cap * 4 * sizeof(uint32_t)));
    m->sort_tmp_bytes = sortRaysByLength(nullptr, 0, m->d_sort, m->d_sort + seg_cap, m->d_sort + 2 * seg_cap, m->d_sort + 3 * seg_cap,
                                         uint32_t(seg_cap), m->stream);
    CU_TRY(m, cudaMalloc(&m->d_sort_tmp, m->sort_tmp_bytes ? m->sort_tmp_bytes : 8));
    m->rays_cap = rc;
    m->seg_cap  = seg_cap;
  }
  if (!index_mode && m->sector_n > 1)
  {
    a.sector_n    = m->sector_n;
    a.sector_rank = m->sector_rank;
    for (int r = 0; r < m->sector_n; ++r) a.sector_bounds[r] = m->sector_bounds[r];
  }
  a.fast_mode = (m->fast_mode && !index_mode) ? 1u : 0u;
  if (a.fast_mode && m->n_leaves != 0)
  {
    int rcc = ensureCoarse(m);
    if (rcc) return rcc;
  }
  a.rays      = m->d_rays;
  a.ends       = index_mode ? nullptr : m->d_ends;
  a.index_mode = index_mode ? 1u : 0u;
  a.segs      = m->d_segs;
  a.seg_cap   = uint32_t(std::min<size_t>(m->seg_cap, 0xFFFFFFF0u));
  a.long_rays = m->d_long;
  a.seg_base  = m->d_long + m->rays_cap;
  // layout inside d_sort (stride = seg_cap): [keys_in | keys_out | idx_in | idx_out]
  a.sort_keys = m->d_sort;
  a.sort_idx  = m->d_sort + 2 * m->seg_cap;
  a.order     = m->d_sort + 3 * m->seg_cap;
  a.sorted_keys = m->d_sort + m->seg_cap;
  // Segment plan. A ray is sequential, so the DDA kernel cannot finish before its longest work item; splitting costs a
  // pre-pass and extra refills, so it is only done when the longest ray of the previous scan exceeded TWICE the average
  // load of a resident lane (few rays per lane: multi-GPU shards, very long rays), with segments of half that load.
  a.seg_len = 0;
  if (const char* e = getenv("VDBM_SEG_LEN")) a.seg_len = uint32_t(atoi(e)); // experiment override
  else if (s.prev_visits)
  {
    const uint64_t lanes    = uint64_t(m->dda_grid) * uint64_t(raycastDDABlock());
    const uint64_t per_lane = std::max<uint64_t>(1, s.prev_visits / lanes);
    if (uint64_t(s.prev_max_visits) > per_lane * 2) a.seg_len = uint32_t(std::min<uint64_t>(4096, std::max<uint64_t>(256, per_lane / 2)));
  }

  // counters before this attempt (needed if the scan has to be replayed after a hash overflow). Every ABI
  // call that changes device counters ends with syncCounters(), so the pinned host copy is current.
  int rc = VDBM_OK;
  const Counters before = *m->h_ctr;
  bool coord_range = false;
  for (int attempt = 0; attempt < 24; ++attempt)
  {
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->ray_cursor, 0, sizeof(unsigned), m->stream));
    CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_extra, 0, 3 * sizeof(unsigned), m->stream)); // n_extra, n_long, max_visits
    // extra-segment slots that end up unused must carry a zero sort key (the DDA kernel stops at the first zero key)
    CU_TRY(m, cudaMemsetAsync(m->d_sort + n, 0, (m->seg_cap - n) * sizeof(uint32_t), m->stream));
    // The DDA kernel marks z-slice mask words; leaves the grid already holds (an earlier accumulate of this period, or
    // the failed attempt being replayed) are in the x-slice layout every other kernel works on: turn them back first.
    if (a.fast_mode)
    {
      // castRayIntoGridFast V:577-602: end-voxel records from prep_rays, then one thread per ray through the map's node levels
      launchPrepRays(a, m->d_ctr, m->stream);
      CU_TRY(m, cudaEventRecord(m->ev2, m->stream));
      launchRaycastFast(a, s.g, m->mt, m->coarse, m->d_bbox, m->n_leaves == 0 ? 1u : 0u, m->d_ctr, m->stream);
      launchCompactLeaves(s.g, m->stream, /*cook=*/false);
    }
    else
    {
    launchUncookLeaves(s.g, s.n_entries, m->stream);
    launchPrepRays(a, m->d_ctr, m->stream);
    uint32_t n_long = 0, n_extra = 0;
    if (a.seg_len)
    {
      // how many rays were split, how many extra segments exist (one small read-back; the sort needs the exact count)
      CU_TRY(m, cudaMemcpyAsync(m->h_small + 12, &m->d_ctr->n_extra, 2 * sizeof(unsigned), cudaMemcpyDeviceToHost, m->stream));
      CU_TRY(m, cudaStreamSynchronize(m->stream));
      n_long  = m->h_small[13];
      n_extra = uint32_t(std::min<uint64_t>(m->h_small[12], a.seg_cap - n));
    }
    a.n_segs = uint32_t(n) + n_extra;
    launchLongRaySegments(a, n_long, m->stream);
    sortRaysByLength(m->d_sort_tmp, m->sort_tmp_bytes, a.sort_keys, m->d_sort + m->seg_cap, a.sort_idx, m->d_sort + 3 * m->seg_cap,
                     a.n_segs, m->stream);
    CU_TRY(m, cudaEventRecord(m->ev2, m->stream));
    launchRaycastDDA(a, s.g, m->d_near, m->d_ctr, m->dda_grid, m->stream, testBeforeSet(s));
    launchCompactLeaves(s.g, m->stream, /*cook=*/true);
    }
    CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
    CU_TRY(m, cudaGetLastError());
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s.g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
    rc = syncCounters(m); // one synchronisation per accumulate: counters, flags, brick and leaf counts
    if (rc) return rc;
    s.n_bricks  = m->h_small[8];
    s.n_entries = m->h_small[9];
    float ms = 0.f;
    cudaEventElapsedTime(&ms, m->ev0, m->ev1);
    m->stats.last_accumulate_ms = ms;
    cudaEventElapsedTime(&ms, m->ev0, m->ev2);
    m->stats.last_prep_ms = ms;
    const uint32_t flags = m->h_ctr->flags;
    if (getenv("VDBM_DEBUG"))
      std::fprintf(stderr, "[vdbm] raycast attempt %d n=%llu seg_len=%u flags=%u bricks=%u cap=%u entries=%u clipped=%llu (before %llu) visits=%llu (before %llu)\n", attempt,
                   (unsigned long long)n, a.seg_len, flags, s.n_bricks, s.cap, s.n_entries, (unsigned long long)m->h_ctr->clipped,
                   (unsigned long long)before.clipped, (unsigned long long)m->h_ctr->visits, (unsigned long long)before.visits);
    // Out-of-range end points are only REPORTED (the rays are dropped, like +-inf); they must not hide an overflow of the
    // brick hash raised by the same scan: complete the grid first (grow, replay), report afterwards.
    if (flags & kFlagCoordRange) coord_range = true;
    const bool overflow = (flags & kFlagUpdateOverflow) != 0;
    const bool crowded  = uint64_t(s.n_bricks) * 10 > uint64_t(s.cap) * 7;
    if (!overflow && !crowded) break;
    rc = growUpdateGrid(m, s);
    if (rc) return rc;
    if (!overflow) break; // table was only crowded: content is complete, no replay needed
    // overflow: some marks were dropped. Restore the counters and replay the scan (marking is an idempotent OR).
    Counters restored   = before;
    restored.flags      = 0;
    restored.ray_cursor = 0;
    CU_TRY(m, cudaMemcpyAsync(m->d_ctr, &restored, sizeof(Counters), cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
  }
  m->stats.last_visits         = m->h_ctr->visits - before.visits;
  m->stats.last_touched_leaves = s.n_entries;
  m->stats.update_capacity     = std::max(m->stats.update_capacity, s.cap * uint32_t(kBrickLeaves));
  if (!a.fast_mode)
  {
    s.prev_visits     = m->stats.last_visits;
    s.prev_max_visits = m->h_ctr->max_visits;
  }
  if (!index_mode)
  {
    m->ends_src = &s;
    m->ends_n   = n;
    for (int k = 0; k < 3; ++k) m->ends_origin[k] = origin[k];
  }
  if (coord_range)
  {
    // clear only this bit: whatever else was raised stays for its own handler
    m->h_ctr->flags &= ~kFlagCoordRange;
    CU_TRY(m, cudaMemcpyAsync(&m->d_ctr->flags, &m->h_ctr->flags, sizeof(unsigned), cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    m->last_error = "some ray end points were outside the +-2^23 voxel range and were dropped";
    return VDBM_ERR_COORD_RANGE;
  }
  return VDBM_OK;
}

int stagePoints(vdbm_map* m, const void* points, uint64_t n, uint64_t stride)
{
  const size_t bytes = size_t(n) * stride;
  if (bytes > m->points_cap)
  {
    cudaFree(m->d_points);
    m->d_points   = nullptr;
    m->points_cap = 0;
    const size_t want = bytes + bytes / 4 + 65536; // geometric growth: no reallocation for slightly larger clouds
    CU_TRY(m, cudaMalloc(&m->d_points, want));
    m->points_cap = want;
  }
  if (bytes) CU_TRY(m, cudaMemcpyAsync(m->d_points, points, bytes, cudaMemcpyHostToDevice, m->stream));
  return VDBM_OK;
}

// room for n entries in the K2a output array (geometric growth: reallocation synchronises the device)
int ensureResolved(vdbm_map* m, size_t n)
{
  if (m->resolved_cap >= n) return VDBM_OK;
  cudaFree(m->d_resolved);
  m->d_resolved   = nullptr;
  m->resolved_cap = 0;
  const size_t cap = n + n / 4 + 1024;
  CU_TRY(m, cudaMalloc(&m->d_resolved, cap * 4));
  m->resolved_cap = cap;
  return VDBM_OK;
}

// updateMap for one source (K2). want_change: keep change records on the device.
int updateMapInternal(vdbm_map* m, Source& s, bool want_change)
{
  s.n_change = 0;
  const uint32_t n = s.n_entries;
  m->stats.last_touched_leaves += n;
  if (n == 0) // VDBMapping.hpp:735-738 (a grid with bricks but no touched leaf is empty too)
  {
    if (s.n_bricks) launchResetBricks(s.g, s.n_bricks, m->stream);
    s.n_bricks = 0;
    return VDBM_OK;
  }
  const uint32_t n_art = m->artificial ? m->artificial->n_entries : 0;
  int rc = ensureMapCapacity(m, uint64_t(n) + n_art);
  if (rc) return rc;
  if (want_change && s.change_cap < n)
  {
    cudaFree(s.d_change);
    s.d_change   = nullptr;
    s.change_cap = 0;
    const uint32_t want = n + n / 4 + 1024;
    CU_TRY(m, cudaMalloc(&s.d_change, size_t(want) * sizeof(LeafRecord)));
    s.change_cap = want;
  }
  rc = ensureResolved(m, n);
  if (rc) return rc;
  CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_change, 0, sizeof(unsigned), m->stream));
  launchApplyUpdate(s.g, m->mt, m->lo, m->d_resolved, want_change ? s.d_change : nullptr, want_change ? s.change_cap : 0, m->d_ctr, n,
                    m->stream);
  launchResetBricks(s.g, s.n_bricks, m->stream); // fresh update grid (VDBMapping.hpp:384)
  if (n_art) launchGridActivate(m->artificial->g, n_art, m->mt, m->d_ctr, m->stream); // V:785-789
  CU_TRY(m, cudaGetLastError());
  s.n_bricks = s.n_entries = 0;
  return VDBM_OK;
}

// ---- leaf sets ---------------------------------------------------------------------------------------
void* hostAlloc(size_t bytes, bool pinned)
{
  if (bytes == 0) bytes = 8;
  void* p = nullptr;
  if (pinned)
  {
    if (cudaHostAlloc(&p, bytes, cudaHostAllocDefault) != cudaSuccess) return nullptr;
    return p;
  }
  return std::malloc(bytes);
}

vdbm_leafset* newLeafset(vdbm_map* m, uint64_t n, bool with_valmask, bool with_values)
{
  auto* ls = new vdbm_leafset();
  ls->n    = n;
  auto up  = [](size_t b) { return (b + 255) & ~size_t(255); };
  const size_t b_or = up(n * 12), b_ac = up(n * 64), b_vm = with_valmask ? up(n * 64) : 0, b_va = with_values ? up(n * 2048) : 0;
  const size_t total = b_or + b_ac + b_vm + b_va;
  // sets of 1 MB and more borrow the handle's persistent pinned staging buffer; everything else (and a second large
  // set while the buffer is lent) lives in plain pageable memory: page-locking per call costs more than it saves
  if (m && total >= (size_t(1) << 20) && !m->h_stage_lent)
  {
    if (m->h_stage_cap < total)
    {
      if (m->h_stage) cudaFreeHost(m->h_stage);
      m->h_stage     = nullptr;
      m->h_stage_cap = 0;
      const size_t want = total + total / 2;
      if (cudaHostAlloc(&m->h_stage, want, cudaHostAllocDefault) == cudaSuccess) m->h_stage_cap = want;
    }
    if (m->h_stage_cap >= total)
    {
      char* p     = static_cast<char*>(m->h_stage);
      ls->origins = reinterpret_cast<int32_t*>(p); p += b_or;
      ls->active  = reinterpret_cast<uint64_t*>(p); p += b_ac;
      if (with_valmask) { ls->valmask = reinterpret_cast<uint64_t*>(p); p += b_vm; }
      if (with_values) ls->values = reinterpret_cast<float*>(p);
      ls->pinned      = true;
      ls->lender      = m;
      m->h_stage_lent = true;
      m->h_stage_borrower = ls;
      return ls;
    }
  }
  ls->pinned  = false;
  ls->origins = static_cast<int32_t*>(hostAlloc(n * 12, ls->pinned));
  ls->active  = static_cast<uint64_t*>(hostAlloc(n * 64, ls->pinned));
  if (with_valmask) ls->valmask = static_cast<uint64_t*>(hostAlloc(n * 64, ls->pinned));
  if (with_values) ls->values = static_cast<float*>(hostAlloc(n * 2048, ls->pinned));
  return ls;
}

struct TempBuf
{
  void* p = nullptr;
  cudaStream_t s;
  explicit TempBuf(cudaStream_t st) : s(st) {}
  cudaError_t alloc(size_t bytes) { return cudaMallocAsync(&p, bytes ? bytes : 8, s); }
  ~TempBuf() { if (p) cudaFreeAsync(p, s); }
  template <typename T> T* as() { return static_cast<T*>(p); }
};

// sort (key, idx) pairs on the device; keys/idx are overwritten with the sorted result
int sortByKey(vdbm_map* m, uint64_t*& keys, uint32_t*& idx, uint64_t* keys_alt, uint32_t* idx_alt, uint32_t n)
{
  if (n == 0) return VDBM_OK;
  size_t bytes = sortPairs(nullptr, 0, keys, keys_alt, idx, idx_alt, n, m->stream);
  TempBuf tmp(m->stream);
  CU_TRY(m, tmp.alloc(bytes));
  sortPairs(tmp.p, bytes, keys, keys_alt, idx, idx_alt, n, m->stream);
  CU_TRY(m, cudaGetLastError());
  std::swap(keys, keys_alt);
  std::swap(idx, idx_alt);
  return VDBM_OK;
}

// export n LeafRecords that live on the device (unsorted) as a sorted bool leaf set: sort (key, index) on the device,
// split the records in that order, copy the three arrays straight into the leaf set
int recordsToLeafset(vdbm_map* m, const LeafRecord* d_recs, uint32_t n, vdbm_leafset** out)
{
  vdbm_leafset* ls = newLeafset(m, n, true, false);
  *out             = ls;
  if (n == 0) return VDBM_OK;
  TempBuf k0(m->stream), k1(m->stream), i0(m->stream), i1(m->stream), so(m->stream), sa(m->stream), sv(m->stream);
  CU_TRY(m, k0.alloc(size_t(n) * 8)); CU_TRY(m, k1.alloc(size_t(n) * 8));
  CU_TRY(m, i0.alloc(size_t(n) * 4)); CU_TRY(m, i1.alloc(size_t(n) * 4));
  CU_TRY(m, so.alloc(size_t(n) * 12)); CU_TRY(m, sa.alloc(size_t(n) * 64)); CU_TRY(m, sv.alloc(size_t(n) * 64));
  uint64_t* keys = k0.as<uint64_t>();
  uint32_t* idx  = i0.as<uint32_t>();
  launchRecordKeys(d_recs, n, keys, idx, m->stream);
  int rc = sortByKey(m, keys, idx, k1.as<uint64_t>(), i1.as<uint32_t>(), n);
  if (rc) return rc;
  launchSplitRecords(d_recs, idx, n, so.as<int32_t>(), sa.as<uint64_t>(), sv.as<uint64_t>(), m->stream);
  CU_TRY(m, cudaMemcpyAsync(ls->origins, so.p, size_t(n) * 12, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(ls->active, sa.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(ls->valmask, sv.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  return VDBM_OK;
}

// a bool grid's touched leaves as a sorted leaf set (origin, active mask, value mask)
int exportGrid(vdbm_map* m, Source& src, vdbm_leafset** out)
{
  Source* s        = &src;
  const uint32_t n = s->n_entries;
  vdbm_leafset* ls = newLeafset(m, n, true, false);
  *out             = ls;
  if (n == 0) return VDBM_OK;
  TempBuf k0(m->stream), k1(m->stream), i0(m->stream), i1(m->stream), recs(m->stream), so(m->stream), sa(m->stream), sv(m->stream);
  CU_TRY(m, k0.alloc(size_t(n) * 8)); CU_TRY(m, k1.alloc(size_t(n) * 8));
  CU_TRY(m, i0.alloc(size_t(n) * 4)); CU_TRY(m, i1.alloc(size_t(n) * 4));
  CU_TRY(m, recs.alloc(size_t(n) * sizeof(LeafRecord)));
  CU_TRY(m, so.alloc(size_t(n) * 12)); CU_TRY(m, sa.alloc(size_t(n) * 64)); CU_TRY(m, sv.alloc(size_t(n) * 64));
  uint64_t* keys = k0.as<uint64_t>();
  uint32_t* idx  = i0.as<uint32_t>();
  launchEntryKeys(s->g, n, keys, idx, m->stream);
  int rc = sortByKey(m, keys, idx, k1.as<uint64_t>(), i1.as<uint32_t>(), n);
  if (rc) return rc;
  launchGatherUpdate(s->g, n, keys, idx, recs.as<LeafRecord>(), m->stream);
  launchSplitRecords(recs.as<LeafRecord>(), nullptr, n, so.as<int32_t>(), sa.as<uint64_t>(), sv.as<uint64_t>(), m->stream);
  CU_TRY(m, cudaMemcpyAsync(ls->origins, so.p, size_t(n) * 12, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(ls->active, sa.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(ls->valmask, sv.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  return VDBM_OK;
}

// Run `mark` (kernels that OR bits into s.g; idempotent) until the brick hash neither overflowed nor is crowded;
// rebuilds the leaf list and the host counts. One synchronisation per attempt.
template <typename MarkFn>
int markIntoGrid(vdbm_map* m, Source& s, MarkFn mark)
{
  for (int attempt = 0; attempt < 24; ++attempt)
  {
    mark();
    launchCompactLeaves(s.g, m->stream);
    CU_TRY(m, cudaGetLastError());
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s.g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
    int rc = syncCounters(m);
    if (rc) return rc;
    s.n_bricks  = m->h_small[8];
    s.n_entries = m->h_small[9];
    const bool overflow = (m->h_ctr->flags & kFlagUpdateOverflow) != 0;
    const bool crowded  = uint64_t(s.n_bricks) * 10 > uint64_t(s.cap) * 7;
    if (!overflow && !crowded) break;
    if (overflow) CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    rc = growUpdateGrid(m, s);
    if (rc) return rc;
    if (!overflow) break;
  }
  m->stats.update_capacity = std::max(m->stats.update_capacity, s.cap * uint32_t(kBrickLeaves));
  return VDBM_OK;
}

int importRecords(vdbm_map* m, Source& s, const LeafRecord* d_records, uint64_t n)
{
  if (n == 0) return VDBM_OK;
  // optimistic: OR the records in; if the brick hash overflowed or got crowded, grow and replay (OR is idempotent)
  int rc = markIntoGrid(m, s, [&] { launchImportUpdate(s.g, d_records, n, m->d_ctr, m->stream); });
  m->stats.last_touched_leaves = s.n_entries;
  return rc;
}

// host leaf arrays (ABI layout) -> OR-ed into the grid of `s`. The arrays are copied as they are (three H2D copies) and
// unpacked by the import kernel; nothing is repacked on the host.
int importLeafArrays(vdbm_map* m, Source& s, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* value)
{
  if (n == 0) return VDBM_OK;
  TempBuf d_o(m->stream), d_a(m->stream), d_v(m->stream);
  CU_TRY(m, d_o.alloc(n * 12)); CU_TRY(m, d_a.alloc(n * 64));
  CU_TRY(m, cudaMemcpyAsync(d_o.p, origins, n * 12, cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaMemcpyAsync(d_a.p, active, n * 64, cudaMemcpyHostToDevice, m->stream));
  if (value)
  {
    CU_TRY(m, d_v.alloc(n * 64));
    CU_TRY(m, cudaMemcpyAsync(d_v.p, value, n * 64, cudaMemcpyHostToDevice, m->stream));
  }
  int rc = markIntoGrid(m, s, [&] { launchImportUpdateSoA(s.g, d_o.as<int32_t>(), d_a.as<uint64_t>(), value ? d_v.as<uint64_t>() : nullptr, n, m->d_ctr, m->stream); });
  m->stats.last_touched_leaves = s.n_entries;
  if (rc) return rc;
  if (m->h_ctr->flags & kFlagCoordRange)
  {
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    return fail(m, VDBM_ERR_COORD_RANGE, "leaf origin outside the +-2^23 voxel range");
  }
  return VDBM_OK;
}

// lazily created library-owned bool grids
int auxGrid(vdbm_map* m, std::unique_ptr<Source>& p, const char* name, uint32_t cap)
{
  if (p) return VDBM_OK;
  auto s = std::make_unique<Source>();
  s->id  = name;
  s->cap = cap;
  int rc = allocUpdateGrid(m, s->g, s->cap);
  if (rc) return rc;
  p = std::move(s);
  return VDBM_OK;
}
int scratchGrid(vdbm_map* m)
{
  int rc = auxGrid(m, m->scratch, "<scratch>", 256);
  if (rc) return rc;
  if (m->scratch->n_entries || m->scratch->n_bricks) return clearUpdateGrid(m, *m->scratch);
  return VDBM_OK;
}

// finish a call that ran updateMapInternal on `s`: timings, counters, optional change grid
int finishUpdate(vdbm_map* m, Source& s, uint64_t upd_before, vdbm_leafset** change)
{
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  int rc = syncCounters(m);
  if (rc) return rc;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  m->stats.last_integrate_ms  = ms;
  m->stats.last_voxel_updates = m->stats.voxel_updates - upd_before;
  if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
  if (change)
  {
    s.n_change = std::min(m->h_ctr->n_change, s.change_cap);
    return recordsToLeafset(m, s.d_change, s.n_change, change);
  }
  return VDBM_OK;
}


// ---- asynchronous scan pipeline ---------------------------------------------------------------------------------
// vdbm_insert_async queues  H2D (copy stream) -> prep -> sort -> DDA -> compaction -> guard -> resolve -> apply -> reset
// and returns. The only host synchronisation of a scan happens when the NEXT call (or any other entry point) finishes
// it: the cloud of scan k+1 crosses PCIe while scan k computes, and the accumulate -> integrate round trip through the
// host disappears. Capacity decisions the host would have taken in between are taken by update_guard_kernel; when it
// refuses (growth needed, overflow, out-of-range points) the update kernels are no-ops and finishPending() simply redoes
// the scan on the synchronous path from the cloud that is still resident (marking is an idempotent OR).
int finishPending(vdbm_map* m)
{
  if (!m->pending.active) return VDBM_OK;
  vdbm_map::Pending pd = m->pending;
  m->pending.active    = false;
  CU_TRY(m, cudaEventSynchronize(m->ev_done));
  Source& s            = *pd.src;
  const Counters& c    = *m->h_ctr; // snapshot copied at the end of the queued chain
  m->n_leaves          = m->h_small[0];
  if (c.deferred_skip)
  {
    // redo synchronously: restore the counters of before the scan, then the ordinary accumulate + integrate
    ++m->async_redone;
    Counters restored   = pd.before;
    restored.flags      = 0;
    restored.ray_cursor = 0;
    CU_TRY(m, cudaMemcpyAsync(m->d_ctr, &restored, sizeof(Counters), cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    *m->h_ctr     = restored;
    m->stats.rays = pd.rays_before;
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s.g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    s.n_bricks  = m->h_small[8];
    s.n_entries = m->h_small[9];
    int rc = raycastDevice(m, s, pd.d_pts, pd.n, pd.stride, pd.origin, s.max_range);
    if (m->ray_in_flight) { m->ends_clobbered = true; m->ends_src = nullptr; } // the staging buffers held the NEXT scan's rays
    if (rc != VDBM_OK && rc != VDBM_ERR_COORD_RANGE) return rc;
    int rc2 = vdbm_integrate(m, 0);
    m->async_expect = uint32_t(m->stats.last_touched_leaves);
    return rc2 ? rc2 : rc;
  }
  ++m->async_fast;
  // the queued path ran to completion: publish what accumulate + integrate would have published
  const uint64_t upd_before = m->stats.voxel_updates;
  publishCumulative(m);
  m->stats.last_visits         = c.visits - pd.before.visits;
  m->stats.last_touched_leaves = c.deferred_entries;
  m->stats.last_voxel_updates  = m->stats.voxel_updates - upd_before;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1); m->stats.last_accumulate_ms = ms;
  cudaEventElapsedTime(&ms, m->ev0, m->ev2); m->stats.last_prep_ms = ms;
  cudaEventElapsedTime(&ms, m->ev1, m->ev3); m->stats.last_integrate_ms = ms;
  s.prev_visits     = m->stats.last_visits;
  s.prev_max_visits = c.max_visits;
  s.prev_updates    = m->stats.last_voxel_updates;
  s.n_bricks = s.n_entries = 0;
  s.n_change        = 0;
  m->async_expect   = c.deferred_entries;
  // the end-voxel records of this scan (its reduced update) are still in the staging buffer unless the NEXT scan's raycast
  // half has already been queued behind it, or a redo re-used the buffer in between
  if (m->ray_in_flight || m->ends_clobbered) m->ends_src = nullptr;
  else
  {
    m->ends_src = &s; m->ends_n = pd.n;
    for (int k = 0; k < 3; ++k) m->ends_origin[k] = pd.origin[k];
  }
  m->ends_clobbered = false;
  return VDBM_OK;
}

// every entry point runs on the handle's device whatever the calling thread's current device is (one process may drive
// several handles on several GPUs: vdbm_group_*)
inline void enterDevice(vdbm_map* m)
{
  int cur = -1;
  if (cudaGetDevice(&cur) != cudaSuccess || cur != m->device) cudaSetDevice(m->device);
}

#define VDBM_ENTER(m)                          \
  do                                           \
  {                                            \
    enterDevice(m);                            \
    if ((m)->pending.active)                   \
    {                                          \
      int rc_enter__ = finishPending(m);       \
      if (rc_enter__) return rc_enter__;       \
    }                                          \
  } while (0)

int ensureAsyncStaging(vdbm_map* m, int buf, size_t bytes)
{
  if (bytes <= m->points_async_cap[buf]) return VDBM_OK;
  cudaFree(m->d_points_async[buf]);
  m->d_points_async[buf]   = nullptr;
  m->points_async_cap[buf] = 0;
  const size_t want = bytes + bytes / 4 + 65536;
  CU_TRY(m, cudaMalloc(&m->d_points_async[buf], want));
  m->points_async_cap[buf] = want;
  return VDBM_OK;
}

// May this scan take the queued path? Everything the synchronous path decides on the host between kernels must be
// unnecessary: one source holding data, no segmentation planned, no artificial areas, staging already large enough.
bool asyncEligible(vdbm_map* m, Source& s, uint64_t n, const double origin[3])
{
  if (n == 0 || n > 0xFFFFFFF0ull || !(s.max_range > 0) || !m->config_set || m->fast_mode || m->sector_n > 1) return false;
  for (int k = 0; k < 3; ++k)
    if (!std::isfinite(origin[k])) return false;
  for (auto& kv : m->sources)
    if (kv.second->n_entries || kv.second->n_bricks) return false;
  if (m->artificial && m->artificial->n_entries) return false;
  if (m->rays_cap < n || m->async_expect == 0) return false; // first scans: let the synchronous path size everything
  if (getenv("VDBM_SEG_LEN")) return false;
  if (s.prev_visits)
  {
    const uint64_t lanes    = uint64_t(m->dda_grid) * uint64_t(raycastDDABlock());
    const uint64_t per_lane = std::max<uint64_t>(1, s.prev_visits / lanes);
    if (uint64_t(s.prev_max_visits) > per_lane * 2) return false; // long rays: the segment planner needs a read-back
  }
  return true;
}

} // namespace

// ======================================================================================================
extern "C" {

int vdbm_abi_version(void) { return VDBM_ABI_VERSION; }

int vdbm_create(const vdbm_params* params, vdbm_map** out)
{
  if (!params || !out || !(params->resolution > 0.0)) return VDBM_ERR_INVALID_ARG;
  *out = nullptr;
  int n_dev = 0;
  if (cudaGetDeviceCount(&n_dev) != cudaSuccess || n_dev == 0)
  {
    std::fprintf(stderr, "vdbm_b200: no CUDA device available; this library has no CPU fallback\n");
    return VDBM_ERR_CUDA;
  }
  // a failure anywhere below releases whatever was allocated so far (vdbm_destroy copes with a half-built handle)
  std::unique_ptr<vdbm_map, void (*)(vdbm_map*)> m(new vdbm_map(), &vdbm_destroy);
  m->params = *params;
  if (params->device >= 0)
  {
    if (cudaSetDevice(params->device) != cudaSuccess) return VDBM_ERR_CUDA;
  }
  cudaGetDevice(&m->device);
  {
    // stream-ordered temporaries (exports, sections, imports) come from the device's default memory pool: keep up to
    // 2 GB of freed blocks cached instead of returning everything to the driver at every synchronisation (the default
    // threshold is 0, which turns each export into a round of driver allocations)
    cudaMemPool_t pool = nullptr;
    if (cudaDeviceGetDefaultMemPool(&pool, m->device) == cudaSuccess && pool)
    {
      uint64_t keep = uint64_t(2) << 30;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
    }
  }
  if (params->stream) m->stream = static_cast<cudaStream_t>(params->stream);
  else
  {
    if (cudaStreamCreateWithFlags(&m->stream, cudaStreamNonBlocking) != cudaSuccess) return VDBM_ERR_CUDA;
    m->own_stream = true;
  }
  vdbm_map* mm = m.get();
  CU_TRY(mm, cudaEventCreate(&m->ev0));
  CU_TRY(mm, cudaEventCreate(&m->ev1));
  CU_TRY(mm, cudaEventCreate(&m->ev2));
  CU_TRY(mm, cudaEventCreate(&m->ev3));
  CU_TRY(mm, cudaEventCreateWithFlags(&m->ev_copy, cudaEventDisableTiming));
  CU_TRY(mm, cudaEventCreateWithFlags(&m->ev_done, cudaEventDisableTiming));
  CU_TRY(mm, cudaStreamCreateWithFlags(&m->copy_stream, cudaStreamNonBlocking));
  CU_TRY(mm, cudaMalloc(&m->d_ctr, sizeof(Counters)));
  CU_TRY(mm, cudaMemsetAsync(m->d_ctr, 0, sizeof(Counters), m->stream));
  CU_TRY(mm, cudaHostAlloc(&m->h_ctr, sizeof(Counters), cudaHostAllocDefault));
  CU_TRY(mm, cudaHostAlloc(&m->h_small, 64 * sizeof(uint32_t), cudaHostAllocDefault));
  CU_TRY(mm, cudaMalloc(&m->d_ctr_alt, sizeof(Counters)));
  CU_TRY(mm, cudaMemsetAsync(m->d_ctr_alt, 0, sizeof(Counters), m->stream));
  CU_TRY(mm, cudaHostAlloc(&m->h_ctr_alt, sizeof(Counters), cudaHostAllocDefault));
  std::memset(m->h_ctr_alt, 0, sizeof(Counters));
  for (auto& e : m->ev_alt) CU_TRY(mm, cudaEventCreate(&e));
  CU_TRY(mm, cudaEventCreateWithFlags(&m->ev_ray, cudaEventDisableTiming));
  {
    int lo = 0, hi = 0; // the update half of a queued scan fills the SM resources the persistent DDA CTAs leave free
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    int prio = hi;
    if (const char* e = getenv("VDBM_UPDATE_PRIORITY")) prio = atoi(e) ? hi : lo; // experiments
    CU_TRY(mm, cudaStreamCreateWithPriority(&m->update_stream, cudaStreamNonBlocking, prio));
  }
  if (const char* e = getenv("VDBM_OVERLAP")) m->overlap = atoi(e) != 0;
  // pinned memory is recycled, not zeroed: a stale counter block of an earlier handle would become the "before" state a
  // replayed first scan restores
  std::memset(m->h_ctr, 0, sizeof(Counters));
  std::memset(m->h_small, 0, 64 * sizeof(uint32_t));
  CU_TRY(mm, cudaMalloc(&m->d_map_counters, 8));
  CU_TRY(mm, cudaMemsetAsync(m->d_map_counters, 0, 8, m->stream));
  m->mt.n_leaves = m->d_map_counters;
  const uint32_t pool = nextPow2(params->map_capacity_leaves ? params->map_capacity_leaves : (1u << 19));
  int rc = allocMapPool(mm, m->mt, pool);
  if (rc) return rc;
  rc = allocMapHash(mm, nextPow2(uint64_t(pool) * 2));
  if (rc) return rc;
  CU_TRY(mm, cudaMalloc(&m->d_near, nearCopiesBytes()));
  CU_TRY(mm, cudaMemsetAsync(m->d_near, 0, nearCopiesBytes(), m->stream));
  m->lo.replicate_quirk = params->replicate_probe_quirk ? 1u : 0u;
  m->dda_grid           = raycastDDAGrid(m->device);
  m->dda_grid_overlap   = m->dda_grid;
  if (const char* e = getenv("VDBM_OVERLAP_DDA_CTAS_PER_SM"))
  {
    int sms = 148;
    cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, m->device);
    m->dda_grid_overlap = std::min(m->dda_grid, std::max(1, atoi(e)) * sms);
  }
  CU_TRY(mm, cudaStreamSynchronize(m->stream));
  *out = m.release();
  return VDBM_OK;
}

void vdbm_destroy(vdbm_map* m)
{
  if (!m) return;
  if (m->pending.active) finishPending(m);
  cudaStreamSynchronize(m->stream);
  cudaStreamSynchronize(m->copy_stream);
  if (m->update_stream) { cudaStreamSynchronize(m->update_stream); cudaStreamDestroy(m->update_stream); }
  cudaFree(m->d_ctr_alt);
  cudaFreeHost(m->h_ctr_alt);
  for (cudaEvent_t e : m->ev_alt) if (e) cudaEventDestroy(e);
  if (m->ev_ray) cudaEventDestroy(m->ev_ray);
  cudaFree(m->d_points_async[0]); cudaFree(m->d_points_async[1]);
  for (cudaEvent_t e : {m->ev3, m->ev_copy, m->ev_done})
    if (e) cudaEventDestroy(e);
  if (m->copy_stream) cudaStreamDestroy(m->copy_stream);
  for (auto& kv : m->sources)
  {
    freeUpdateGrid(kv.second->g);
    freeUpdateGrid(kv.second->g_alt);
    cudaFree(kv.second->d_change);
  }
  for (Source* aux : {m->scratch.get(), m->artificial.get()})
    if (aux) { freeUpdateGrid(aux->g); cudaFree(aux->d_change); }
  cudaFree(m->d_ends);
  cudaFree(m->coarse.k1); cudaFree(m->coarse.k2); cudaFree(m->coarse.counts); cudaFree(m->d_bbox);
  freeMapPool(m->mt);
  cudaFree(m->mt.hkeys); cudaFree(m->mt.hvals);
  cudaFree(m->d_ctr); cudaFree(m->d_map_counters); cudaFree(m->d_points); cudaFree(m->d_rays); cudaFree(m->d_part);
  cudaFree(m->d_sort); cudaFree(m->d_sort_tmp); cudaFree(m->d_resolved); cudaFree(m->d_near); cudaFree(m->d_segs); cudaFree(m->d_long);
  for (void* p : m->ex.opened) cudaIpcCloseMemHandle(p);
  cudaFree(m->ex.inbox); cudaFree(m->ex.ctrl); cudaFree(m->ex.d_cursors); cudaFree(m->ex.d_counts);
  for (auto& e : m->ex.ev) if (e) cudaEventDestroy(e);
  cudaFreeHost(m->h_ctr); cudaFreeHost(m->h_small);
  if (m->h_stage_lent && m->h_stage_borrower)
  {
    // a leaf set still borrows the staging buffer (e.g. numpy views of a large export): hand the buffer over to it
    m->h_stage_borrower->lender       = nullptr;
    m->h_stage_borrower->orphan_stage = m->h_stage;
    m->h_stage                        = nullptr;
  }
  for (int b = 0; b < vdbm_map::Mirror::kBufs; ++b)
  {
    cudaFree(m->mirror.d_buf[b]);
    if (m->mirror.h_buf[b]) cudaFreeHost(m->mirror.h_buf[b]);
    if (m->mirror.ev[b]) cudaEventDestroy(m->mirror.ev[b]);
  }
  if (m->h_stage) cudaFreeHost(m->h_stage);
  for (cudaEvent_t e : {m->ev0, m->ev1, m->ev2})
    if (e) cudaEventDestroy(e);
  if (m->own_stream && m->stream) cudaStreamDestroy(m->stream);
  delete m;
}

int vdbm_reset(vdbm_map* m)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  // resetMap V:174-186: new empty map, new empty update grids
  CU_TRY(m, cudaMemsetAsync(m->mt.hkeys, 0xFF, size_t(m->hcap) * 8, m->stream));
  CU_TRY(m, cudaMemsetAsync(m->mt.leaf_dirty, 0, size_t(m->mt.pool_cap) * 4, m->stream));
  CU_TRY(m, cudaMemsetAsync(m->d_map_counters, 0, 8, m->stream));
  m->n_leaves = 0;
  m->coarse_built = ~0u; // the coarse node sets describe the old map: rebuilt from scratch on next use
  ++m->generation;
  for (auto& kv : m->sources)
  {
    Source& s = *kv.second;
    int rc    = clearUpdateGrid(m, s);
    if (rc) return rc;
    s.n_change = 0;
  }
  m->ends_src = nullptr;
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  m->stats.map_leaves = 0;
  return VDBM_OK;
}

int vdbm_set_config(vdbm_map* m, double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (max_range < 0.0) return fail(m, VDBM_ERR_BAD_CONFIG, "Max range invalid. Range cannot be negative."); // V:1458-1463
  m->max_range  = max_range;
  m->config_set = true; // V:1468 — before the derived checks, like the reference
  if (prob_miss > 0.5) return fail(m, VDBM_ERR_BAD_CONFIG, "Probability for a miss should be below 0.5"); // O:65-70
  if (prob_hit < 0.5) return fail(m, VDBM_ERR_BAD_CONFIG, "Probability for a hit should be above 0.5");   // O:71-76
  // O:79-87, evaluated in double on the host exactly like the reference, then narrowed
  m->lo.miss      = static_cast<float>(std::log(prob_miss) - std::log(1 - prob_miss));
  m->lo.hit       = static_cast<float>(std::log(prob_hit) - std::log(1 - prob_hit));
  m->lo.thres_min = static_cast<float>(std::log(prob_thres_min) - std::log(1 - prob_thres_min));
  m->lo.thres_max = static_cast<float>(std::log(prob_thres_max) - std::log(1 - prob_thres_max));
  m->lo.max_lo    = static_cast<float>(std::log(0.99) - std::log(0.01));
  m->lo.min_lo    = static_cast<float>(std::log(0.01) - std::log(0.99));
  // OpenVDB InternalNode tile probe of a miss on an (0.0f, inactive) background tile, run with state=true
  {
    volatile float pv = 0.0f + m->lo.miss;
    bool pa           = true;
    if (pv < m->lo.thres_min)
    {
      pa = false;
      if (pv < m->lo.min_lo) pv = m->lo.min_lo;
    }
    m->lo.miss_probe_flips     = pa ? 0u : 1u;
    m->lo.miss_probe_no_create = (!pa && pv == 0.0f) ? 1u : 0u;
  }
  return VDBM_OK;
}

int vdbm_get_logodds(vdbm_map* m, float* out6)
{
  if (!m || !out6) return VDBM_ERR_INVALID_ARG;
  out6[0] = m->lo.hit; out6[1] = m->lo.miss; out6[2] = m->lo.thres_min; out6[3] = m->lo.thres_max;
  out6[4] = m->lo.max_lo; out6[5] = m->lo.min_lo;
  return VDBM_OK;
}

int vdbm_source_add(vdbm_map* m, const char* source_id, double max_range)
{
  if (!m || !source_id) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  auto it = m->sources.find(source_id);
  if (it != m->sources.end())
  {
    // the reference overwrites the map entry with a fresh InputSource (V:1374)
    Source& s = *it->second;
    int rc    = clearUpdateGrid(m, s);
    if (rc) return rc;
    s.max_range = (max_range == 0) ? m->max_range : max_range;
    return VDBM_OK;
  }
  auto s       = std::make_unique<Source>();
  s->id        = source_id;
  s->max_range = (max_range == 0) ? m->max_range : max_range; // V:1356-1363
  // capacity hint is in leaves; the grid is a hash of 512-leaf bricks (64 KB each). Default: 4096 bricks = 256 MB.
  s->cap = m->params.update_capacity_leaves ? nextPow2(std::max<uint64_t>(8, m->params.update_capacity_leaves / 64)) : 4096u;
  int rc = allocUpdateGrid(m, s->g, s->cap);
  if (rc) return rc;
  m->stats.update_capacity = std::max(m->stats.update_capacity, s->cap * uint32_t(kBrickLeaves));
  m->sources[source_id]    = std::move(s);
  return VDBM_OK;
}

int vdbm_raycast(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3],
                 double raycast_range)
{
  if (!m || (!points && n) || !origin || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  if (!m->config_set) return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
  int rc = stagePoints(m, points, n, stride_bytes);
  if (rc) return rc;
  return raycastDevice(m, *s, m->d_points, n, stride_bytes, origin, raycast_range);
}

int vdbm_accumulate(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  if (!m || (!points && n) || !origin || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  // a prefetched copy is good for exactly ONE accumulate call, whatever that call ends up doing (an ignored call must
  // not leave a stale copy behind that a later call with a refilled buffer at the same address would pick up)
  const auto pf     = m->prefetch;
  m->prefetch.valid = false;
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : "")); // V:320-326
  if (!(s->max_range > 0)) return VDBM_OK;                                                                             // V:331
  if (!m->config_set) return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
  if (pf.valid && pf.host == points && pf.n == n && pf.stride == stride_bytes)
  {
    // the cloud was uploaded ahead of time by vdbm_prefetch (copy stream): no H2D on the critical path
    m->async_buf = pf.buf;
    CU_TRY(m, cudaEventSynchronize(m->ev_copy));
    return raycastDevice(m, *s, m->d_points_async[pf.buf], n, stride_bytes, origin, s->max_range);
  }
  int rc = stagePoints(m, points, n, stride_bytes);
  if (rc) return rc;
  return raycastDevice(m, *s, m->d_points, n, stride_bytes, origin, s->max_range);
}

int vdbm_accumulate_device(vdbm_map* m, const char* source_id, const void* d_points, uint64_t n, uint64_t stride_bytes,
                           const double origin[3])
{
  if (!m || (!d_points && n) || !origin || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  if (!(s->max_range > 0)) return VDBM_OK;
  if (!m->config_set) return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
  return raycastDevice(m, *s, static_cast<const uint8_t*>(d_points), n, stride_bytes, origin, s->max_range);
}

int vdbm_integrate(vdbm_map* m, int keep_change)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  m->stats.last_touched_leaves = 0;
  const uint64_t upd_before    = m->stats.voxel_updates;
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  Source* only   = nullptr; // the one source that holds data, if there is exactly one (its ray-overlap ratio is then known)
  int with_data  = 0;
  for (auto& kv : m->sources)
    if (kv.second->n_entries) { only = kv.second.get(); ++with_data; }
  for (auto& kv : m->sources) // std::map key order, V:380
  {
    int rc = updateMapInternal(m, *kv.second, keep_change != 0);
    if (rc) return rc;
    if (m->sources.size() > 1 || keep_change)
    {
      // the next source's capacity check needs the new leaf count; change export needs n_change
      rc = syncCounters(m);
      if (rc) return rc;
      kv.second->n_change = std::min(m->h_ctr->n_change, kv.second->change_cap);
    }
  }
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  int rc = syncCounters(m);
  if (rc) return rc;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  m->stats.last_integrate_ms  = ms;
  m->stats.last_voxel_updates = m->stats.voxel_updates - upd_before;
  if (with_data == 1) only->prev_updates = m->stats.last_voxel_updates;
  if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
  return VDBM_OK;
}

int vdbm_integrate_from(vdbm_map* m, vdbm_map* holder, const char* source_id, int keep_change)
{
  if (!m || !holder) return VDBM_ERR_INVALID_ARG;
  if (holder != m)
  {
    if (holder->device != m->device || holder->params.resolution != m->params.resolution)
      return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_integrate_from: the holder must live on the same device and have the same resolution");
    VDBM_ENTER(holder); // a scan still queued there is finished first
    // accumulate calls are synchronous, so this returns at once; it orders anything else the caller queued there before us
    CU_TRY(holder, cudaStreamSynchronize(holder->stream));
  }
  VDBM_ENTER(m);
  Source* s = findSource(holder, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  m->stats.last_touched_leaves = 0;
  const uint64_t upd_before    = m->stats.voxel_updates;
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  int rc = updateMapInternal(m, *s, keep_change != 0); // this map's stream, counters and leaf pool; the holder's grid
  if (rc) return rc;
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  rc = syncCounters(m); // the grid is consumed and reset when this returns: the holder may accumulate again
  if (rc) return rc;
  if (keep_change) s->n_change = std::min(m->h_ctr->n_change, s->change_cap);
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  m->stats.last_integrate_ms  = ms;
  m->stats.last_voxel_updates = m->stats.voxel_updates - upd_before;
  s->prev_updates             = m->stats.last_voxel_updates;
  if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
  return VDBM_OK;
}

int vdbm_insert(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  // insertPointCloud V:399-406: accumulateUpdate (whatever it does) then integrateUpdate
  int rc_acc = vdbm_accumulate(m, source_id, points, n, stride_bytes, origin);
  if (rc_acc == VDBM_ERR_INVALID_ARG || rc_acc == VDBM_ERR_CUDA || rc_acc == VDBM_ERR_OUT_OF_MEMORY) return rc_acc;
  int rc_int = vdbm_integrate(m, 0);
  return rc_int ? rc_int : rc_acc;
}

int vdbm_prefetch(vdbm_map* m, const void* points, uint64_t n, uint64_t stride_bytes)
{
  if (!m || (!points && n) || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  // no VDBM_ENTER: the point is to run while earlier work is still in flight. The staging buffer used is the one the
  // queued / last scan does NOT read.
  const int buf      = m->async_buf ^ 1;
  const size_t bytes = size_t(n) * stride_bytes;
  m->prefetch.valid  = false;
  if (bytes == 0) return VDBM_OK;
  int rc = ensureAsyncStaging(m, buf, bytes);
  if (rc) return rc;
  CU_TRY(m, cudaMemcpyAsync(m->d_points_async[buf], points, bytes, cudaMemcpyHostToDevice, m->copy_stream));
  CU_TRY(m, cudaEventRecord(m->ev_copy, m->copy_stream));
  m->prefetch.host = points; m->prefetch.n = n; m->prefetch.stride = stride_bytes; m->prefetch.buf = buf; m->prefetch.valid = true;
  return VDBM_OK; // async_buf (the buffer the latest scan reads) only moves when a scan actually uses this copy
}

int vdbm_flush(vdbm_map* m)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  return finishPending(m);
}

// (d_ctr, h_ctr, ev0..ev3) of the handle and (g, cap, n_bricks, n_entries) of the source <-> their alternates
static void swapParity(vdbm_map* m, Source& s)
{
  std::swap(m->d_ctr, m->d_ctr_alt);
  std::swap(m->h_ctr, m->h_ctr_alt);
  std::swap(m->ev0, m->ev_alt[0]); std::swap(m->ev1, m->ev_alt[1]); std::swap(m->ev2, m->ev_alt[2]); std::swap(m->ev3, m->ev_alt[3]);
  std::swap(s.g, s.g_alt);
  std::swap(s.cap, s.cap_alt); std::swap(s.n_bricks, s.n_bricks_alt); std::swap(s.n_entries, s.n_entries_alt);
}

// Raycast half of a queued scan on m->stream, into the CURRENT set: prep -> sort -> DDA -> merge -> compaction (+ cook)
static int queueRaycastHalf(vdbm_map* m, Source& s, const uint8_t* d_pts, uint64_t n, uint64_t stride_bytes, const double origin[3], int dda_grid)
{
  RaycastArgs a{};
  a.points = d_pts;
  a.n      = n;
  a.stride = uint32_t(stride_bytes);
  for (int k = 0; k < 3; ++k) a.origin[k] = origin[k];
  if (!originIndex(m->params.resolution, origin, a.origin_idx)) return fail(m, VDBM_ERR_COORD_RANGE, "sensor origin outside the +-2^23 voxel range");
  a.range      = s.max_range;
  a.resolution = m->params.resolution;
  a.half_res   = m->params.resolution / 2.0;
  a.inv_res    = 1.0 / m->params.resolution;
  a.rays       = m->d_rays;
  a.ends       = m->d_ends;
  a.index_mode = 0;
  a.segs       = m->d_segs;
  a.seg_cap    = uint32_t(std::min<size_t>(m->seg_cap, 0xFFFFFFF0u));
  a.long_rays  = m->d_long;
  a.seg_base   = m->d_long + m->rays_cap;
  a.sort_keys  = m->d_sort;
  a.sort_idx   = m->d_sort + 2 * m->seg_cap;
  a.order      = m->d_sort + 3 * m->seg_cap;
  a.sorted_keys = m->d_sort + m->seg_cap;
  a.seg_len    = 0;
  a.n_segs     = uint32_t(n);
  CU_TRY(m, cudaMemsetAsync(&m->d_ctr->ray_cursor, 0, sizeof(unsigned), m->stream));
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_extra, 0, 3 * sizeof(unsigned), m->stream));
  CU_TRY(m, cudaMemsetAsync(m->d_sort + n, 0, (m->seg_cap - n) * sizeof(uint32_t), m->stream));
  launchPrepRays(a, m->d_ctr, m->stream);
  sortRaysByLength(m->d_sort_tmp, m->sort_tmp_bytes, a.sort_keys, m->d_sort + m->seg_cap, a.sort_idx, m->d_sort + 3 * m->seg_cap, a.n_segs,
                   m->stream);
  CU_TRY(m, cudaEventRecord(m->ev2, m->stream));
  launchRaycastDDA(a, s.g, m->d_near, m->d_ctr, dda_grid, m->stream, testBeforeSet(s));
  launchCompactLeaves(s.g, m->stream, /*cook=*/true); // the grid was empty (asyncEligible): nothing to uncook before
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  CU_TRY(m, cudaGetLastError());
  return VDBM_OK;
}

// A raycast half was queued into the current set but its scan cannot be completed (an error in between): wait for it,
// forget what it marked and what it counted.
static int abandonQueuedRaycast(vdbm_map* m, Source& s, const Counters& before)
{
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  int rc = readGridCounters(m, s);
  if (rc) return rc;
  rc = clearUpdateGrid(m, s);
  if (rc) return rc;
  Counters restored   = before;
  restored.flags      = 0;
  restored.ray_cursor = 0;
  CU_TRY(m, cudaMemcpyAsync(m->d_ctr, &restored, sizeof(Counters), cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  *m->h_ctr = restored;
  return VDBM_OK;
}

int vdbm_insert_async(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3],
                      int points_on_device)
{
  if (!m || (!points && n) || !origin || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  enterDevice(m);
  Source* sp = findSource(m, source_id);
  m->prefetch.valid = false;
  // 1. start the upload of THIS cloud while the previous scan may still be computing (its buffer is the other one)
  const int buf      = m->async_buf ^ 1;
  const size_t bytes = size_t(n) * stride_bytes;
  const uint8_t* d_pts = static_cast<const uint8_t*>(points);
  bool uploaded = false;
  if (sp && !points_on_device && bytes && bytes <= m->points_async_cap[buf])
  {
    CU_TRY(m, cudaMemcpyAsync(m->d_points_async[buf], points, bytes, cudaMemcpyHostToDevice, m->copy_stream));
    CU_TRY(m, cudaEventRecord(m->ev_copy, m->copy_stream));
    uploaded = true;
  }
  // 2a. OVERLAP: the raycast half of this scan does not read the map, so it is queued - into the source's OTHER update grid,
  // counting into the other counter block - BEFORE the previous scan is waited for: its DDA (instruction-bound) then runs
  // next to the previous scan's updateMap (HBM-bound), which sits on the update stream.
  bool ray_queued = false;
  Counters before{};
  uint64_t rays_before = 0;
  cudaStream_t upd_stream = m->overlap ? m->update_stream : m->stream;
  if (m->overlap && m->pending.active && sp && m->pending.src == sp && (points_on_device || uploaded) && asyncEligible(m, *sp, n, origin))
  {
    Source& s = *sp;
    if (!s.g_alt.bkeys)
    {
      s.cap_alt = s.cap;
      int rca   = allocUpdateGrid(m, s.g_alt, s.cap_alt);
      if (rca) return rca;
    }
    swapParity(m, s);
    before      = *m->h_ctr;
    rays_before = m->stats.rays;
    if (!points_on_device)
    {
      d_pts = m->d_points_async[buf];
      CU_TRY(m, cudaStreamWaitEvent(m->stream, m->ev_copy, 0));
    }
    const int rcq = queueRaycastHalf(m, s, d_pts, n, stride_bytes, origin, m->dda_grid_overlap);
    swapParity(m, s); // the previous scan's set is the current one again: finishPending() works on it
    if (rcq) return rcq; // (only the origin check can fail, before anything is queued)
    ray_queued          = true;
    m->ray_in_flight    = true;
  }
  // 2b. finish the previous scan (the one synchronisation per scan)
  int rc = finishPending(m);
  m->ray_in_flight = false;
  // out-of-range points of the PREVIOUS scan are a warning, not a reason to drop this one: remember it, carry on, and
  // return it at the end unless this scan has something of its own to report
  int rc_prev = VDBM_OK;
  if (rc == VDBM_ERR_COORD_RANGE) { rc_prev = rc; rc = VDBM_OK; }
  if (rc)
  {
    if (ray_queued) { swapParity(m, *sp); abandonQueuedRaycast(m, *sp, before); swapParity(m, *sp); }
    return rc;
  }
  if (!sp)
  {
    // like vdbm_insert: accumulateUpdate complains and returns (V:320-326), integrateUpdate still runs (V:404)
    rc = vdbm_integrate(m, 0);
    if (rc) return rc;
    return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  }
  Source& s = *sp;
  if (!points_on_device)
  {
    if (!uploaded && bytes)
    {
      rc = ensureAsyncStaging(m, buf, bytes);
      if (rc) return rc;
      CU_TRY(m, cudaMemcpyAsync(m->d_points_async[buf], points, bytes, cudaMemcpyHostToDevice, m->copy_stream));
      CU_TRY(m, cudaEventRecord(m->ev_copy, m->copy_stream));
    }
    d_pts = m->d_points_async[buf];
    // the caller's buffer is free again when this call returns; by now the copy has normally long finished (it ran
    // while finishPending waited for the previous scan)
    if (bytes) CU_TRY(m, cudaEventSynchronize(m->ev_copy));
    m->async_buf = buf;
  }
  if (ray_queued)
  {
    swapParity(m, s); // this scan's set (grid + counter block + events) becomes the current one
    ++m->async_overlapped;
  }
  else
  {
    // 3. not eligible for the queued path: the ordinary synchronous insert on the uploaded cloud
    if (!asyncEligible(m, s, n, origin))
    {
      ++m->async_sync;
      if (!(s.max_range > 0)) return vdbm_integrate(m, 0); // V:331 then integrateUpdate
      if (!m->config_set)
      {
        vdbm_integrate(m, 0);
        return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
      }
      int rc_acc = raycastDevice(m, s, d_pts, n, stride_bytes, origin, s.max_range);
      if (rc_acc != VDBM_OK && rc_acc != VDBM_ERR_COORD_RANGE) return rc_acc;
      int rc_int = vdbm_integrate(m, 0);
      m->async_expect = uint32_t(std::max<uint64_t>(1, m->stats.last_touched_leaves));
      return rc_int ? rc_int : (rc_acc ? rc_acc : rc_prev);
    }
    before      = *m->h_ctr;
    rays_before = m->stats.rays;
  }
  // 4. queue the scan (its raycast half may already be running)
  const uint32_t expect = m->async_expect + m->async_expect / 2 + 65536; // leaves the update is sized for (guard checks the truth)
  rc = ensureMapCapacity(m, expect);
  if (!rc) rc = ensureResolved(m, expect);
  if (!rc && !ray_queued) rc = queueRaycastHalf(m, s, d_pts, n, stride_bytes, origin, m->dda_grid);
  if (rc)
  {
    if (ray_queued) abandonQueuedRaycast(m, s, before);
    return rc;
  }
  m->pending.before      = before;
  m->pending.rays_before = rays_before;
  m->stats.rays += n;
  m->ends_src = nullptr;
  // update half: guard -> resolve -> apply -> reset, behind everything queued on the main stream so far (the raycast half and,
  // possibly, a reallocation of the map tables)
  if (upd_stream != m->stream)
  {
    CU_TRY(m, cudaEventRecord(m->ev_ray, m->stream));
    CU_TRY(m, cudaStreamWaitEvent(upd_stream, m->ev_ray, 0));
  }
  launchApplyUpdateDeferred(s.g, m->mt, m->lo, m->d_resolved, uint32_t(std::min<size_t>(m->resolved_cap, 0xFFFFFFFFu)), m->d_ctr, expect, upd_stream);
  CU_TRY(m, cudaEventRecord(m->ev3, upd_stream));
  CU_TRY(m, cudaGetLastError());
  CU_TRY(m, cudaMemcpyAsync(m->h_ctr, m->d_ctr, sizeof(Counters), cudaMemcpyDeviceToHost, upd_stream));
  CU_TRY(m, cudaMemcpyAsync(m->h_small, m->d_map_counters, 4, cudaMemcpyDeviceToHost, upd_stream));
  CU_TRY(m, cudaEventRecord(m->ev_done, upd_stream));
  m->pending.active = true;
  m->pending.src    = &s;
  m->pending.n      = n;
  m->pending.stride = stride_bytes;
  m->pending.d_pts  = d_pts;
  for (int k = 0; k < 3; ++k) m->pending.origin[k] = origin[k];
  return rc_prev;
}

int vdbm_update_map(vdbm_map* m, const char* source_id, vdbm_leafset** change)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  m->stats.last_touched_leaves = 0;
  const uint64_t upd_before    = m->stats.voxel_updates;
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  int rc = updateMapInternal(m, *s, change != nullptr);
  if (rc) return rc;
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  rc = syncCounters(m);
  if (rc) return rc;
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  m->stats.last_integrate_ms  = ms;
  m->stats.last_voxel_updates = m->stats.voxel_updates - upd_before;
  if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
  if (change)
  {
    s->n_change = std::min(m->h_ctr->n_change, s->change_cap);
    return recordsToLeafset(m, s->d_change, s->n_change, change);
  }
  return VDBM_OK;
}

int vdbm_change_export(vdbm_map* m, const char* source_id, vdbm_leafset** out)
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  return recordsToLeafset(m, s->d_change, s->n_change, out);
}

int vdbm_update_export(vdbm_map* m, const char* source_id, vdbm_leafset** out)
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  return exportGrid(m, *s, out);
}

int vdbm_update_import(vdbm_map* m, const char* source_id, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* value)
{
  if (!m || (n && (!origins || !active || !value))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  if (n == 0) return VDBM_OK;
  return importLeafArrays(m, *s, n, origins, active, value);
}

int vdbm_update_import_device(vdbm_map* m, const char* source_id, const void* d_records, uint64_t n)
{
  if (!m || (n && !d_records)) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  return importRecords(m, *s, static_cast<const LeafRecord*>(d_records), n);
}

// ---- remote-mapping deltas: createUpdate / applyUpdate --------------------------------------
int vdbm_update_create(vdbm_map* m, const char* source_id, int level, vdbm_leafset** out, double origin_out[3])
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  if (origin_out)
    for (int k = 0; k < 3; ++k) origin_out[k] = (m->ends_src == s) ? m->ends_origin[k] : 0.0;
  if (level == 0) return exportGrid(m, *s, out);
  if (level == 1) return recordsToLeafset(m, s->d_change, s->n_change, out);
  if (level != 2) return fail(m, VDBM_ERR_INVALID_ARG, "update level must be 0, 1 or 2");
  if (m->fast_mode)
    return fail(m, VDBM_ERR_INVALID_ARG, "level 2 (reduced) updates describe castRayIntoGrid scans; in fast_mode the update grid depends on the map: use level 0");
  if (m->ends_src != s)
    return fail(m, VDBM_ERR_INVALID_ARG, "no reduced update available: level 2 describes the source's LAST accumulate call on this handle");
  int rc = scratchGrid(m);
  if (rc) return rc;
  Source& sc = *m->scratch;
  rc = markIntoGrid(m, sc, [&] { launchMarkEnds(m->d_ends, m->ends_n, sc.g, m->d_ctr, m->stream); });
  if (rc) return rc;
  rc = exportGrid(m, sc, out);
  if (rc) return rc;
  return clearUpdateGrid(m, sc);
}

int vdbm_update_apply(vdbm_map* m, int level, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* value,
                      const double origin[3], vdbm_leafset** change)
{
  if (!m || (n && (!origins || !active || !value)) || level < 0 || level > 2 || (level == 2 && !origin)) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (change) *change = nullptr;
  if (!m->config_set) return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
  int rc = scratchGrid(m);
  if (rc) return rc;
  Source& sc = *m->scratch;
  rc = importLeafArrays(m, sc, n, origins, active, value);
  if (rc) return rc;
  m->stats.last_touched_leaves = 0;
  const uint64_t upd_before    = m->stats.voxel_updates;
  if (level == 1)
  {
    // overwrite grid: every active voxel is forced occupied (value bit) or free, O:118-129
    rc = ensureMapCapacity(m, sc.n_entries);
    if (rc) return rc;
    launchOverwrite(sc.g, sc.n_entries, m->mt, m->lo, m->d_ctr, m->stream);
    rc = clearUpdateGrid(m, sc);
    if (rc) return rc;
    rc = syncCounters(m);
    if (rc) return rc;
    if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
    if (change) *change = newLeafset(m, 0, true, false);
    return VDBM_OK;
  }
  if (level == 2 && sc.n_entries)
  {
    // reduced grid -> end-voxel records -> the sender's update grid again (castRayIntoGrid depends on voxel indices only)
    uint64_t total = 0;
    for (uint64_t i = 0; i < n * 8; ++i) total += uint64_t(__builtin_popcountll(active[i]));
    if (total > 0xFFFFFFF0ull) return fail(m, VDBM_ERR_INVALID_ARG, "more than 2^32 rays in one reduced update");
    if (total * 16 > m->points_cap)
    {
      cudaFree(m->d_points);
      m->d_points   = nullptr;
      m->points_cap = 0;
      const size_t want = total * 16 + total * 4 + 65536;
      CU_TRY(m, cudaMalloc(&m->d_points, want));
      m->points_cap = want;
    }
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_out, 0, sizeof(unsigned), m->stream));
    launchExpandEnds(sc.g, sc.n_entries, reinterpret_cast<int4*>(m->d_points), uint32_t(total), m->d_ctr, m->stream);
    rc = clearUpdateGrid(m, sc);
    if (rc) return rc;
    rc = syncCounters(m);
    if (rc) return rc;
    const uint64_t n_rays = std::min<uint64_t>(m->h_ctr->n_out, total);
    rc = raycastDevice(m, sc, m->d_points, n_rays, 16, origin, 0.0, /*index_mode=*/true);
    if (rc) return rc;
  }
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  rc = updateMapInternal(m, sc, change != nullptr);
  if (rc) return rc;
  return finishUpdate(m, sc, upd_before, change);
}

// ---- direct map edits -------------------------------------------------------------------------
int vdbm_points_set(vdbm_map* m, const void* points, uint64_t n, uint64_t stride_bytes, int occupied)
{
  if (!m || (!points && n) || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (n == 0) return VDBM_OK;
  if (n > 0xFFFFFFF0ull) return fail(m, VDBM_ERR_INVALID_ARG, "more than 2^32 points in one cloud");
  if (!m->config_set) return fail(m, VDBM_ERR_NOT_CONFIGURED, "Map not properly configured. Did you call setConfig method?");
  int rc = stagePoints(m, points, n, stride_bytes);
  if (rc) return rc;
  rc = scratchGrid(m);
  if (rc) return rc;
  Source& sc = *m->scratch;
  TempBuf ends(m->stream);
  CU_TRY(m, ends.alloc(size_t(n) * sizeof(int4)));
  launchPointsToEnds(m->d_points, n, uint32_t(stride_bytes), 1.0 / m->params.resolution, occupied, ends.as<int4>(), m->d_ctr, m->stream);
  rc = markIntoGrid(m, sc, [&] { launchMarkEnds(ends.as<int4>(), n, sc.g, m->d_ctr, m->stream); });
  if (rc) return rc;
  const bool range_err = (m->h_ctr->flags & kFlagCoordRange) != 0;
  if (range_err) CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
  rc = ensureMapCapacity(m, sc.n_entries);
  if (rc) return rc;
  launchOverwrite(sc.g, sc.n_entries, m->mt, m->lo, m->d_ctr, m->stream);
  rc = clearUpdateGrid(m, sc);
  if (rc) return rc;
  rc = syncCounters(m);
  if (rc) return rc;
  if (m->h_ctr->flags & kFlagMapOverflow) return fail(m, VDBM_ERR_OUT_OF_MEMORY, "map hash / leaf pool overflow");
  if (range_err) return fail(m, VDBM_ERR_COORD_RANGE, "some points were outside the +-2^23 voxel range and were dropped");
  return VDBM_OK;
}

int vdbm_map_integrity_restore(vdbm_map* m)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (!m->artificial || m->artificial->n_entries == 0) return VDBM_OK;
  Source& ar = *m->artificial;
  int rc     = ensureMapCapacity(m, (0.0f > m->lo.thres_max) ? ar.n_entries : 0);
  if (rc) return rc;
  launchRestoreState(ar.g, ar.n_entries, m->mt, m->lo, m->d_ctr, m->stream);
  rc = clearUpdateGrid(m, ar); // V:1164
  if (rc) return rc;
  return syncCounters(m);
}

// addArtificialPolygon V:1198-1207 / addArtificialWall V:1217-1236 into the artificial-area grid: one wall per consecutive point
// pair of every polyline (plus the closing edge when closed), end points through worldToIndex V:1224-1226
static int addWalls(vdbm_map* m, uint64_t n_polygons, const uint32_t* counts, const double* xyz, double negative_height, double positive_height,
                    bool closed)
{
  int rc = auxGrid(m, m->artificial, "<artificial>", 64);
  if (rc) return rc;
  Source& ar = *m->artificial;
  std::vector<int32_t> walls;
  size_t base = 0;
  for (uint64_t p = 0; p < n_polygons; ++p)
  {
    const uint32_t n_edges = closed ? counts[p] : (counts[p] ? counts[p] - 1 : 0);
    for (uint32_t i = 0; i < n_edges; ++i)
    {
      int32_t a[3], b[3];
      if (!originIndex(m->params.resolution, xyz + 3 * (base + i), a) || !originIndex(m->params.resolution, xyz + 3 * (base + (i + 1) % counts[p]), b))
        return fail(m, VDBM_ERR_COORD_RANGE, "artificial area corner outside the +-2^23 voxel range");
      walls.insert(walls.end(), {a[0], a[1], a[2], b[0], b[1], b[2]});
    }
    base += counts[p];
  }
  const uint32_t n_walls = uint32_t(walls.size() / 6);
  if (n_walls == 0) return VDBM_OK;
  const int32_t neg_index = int32_t(negative_height / m->params.resolution); // V:1228-1229: C cast, truncation
  const int32_t pos_index = int32_t(positive_height / m->params.resolution);
  TempBuf d(m->stream);
  CU_TRY(m, d.alloc(walls.size() * 4));
  CU_TRY(m, cudaMemcpyAsync(d.p, walls.data(), walls.size() * 4, cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  rc = markIntoGrid(m, ar, [&] { launchWallDDA(d.as<int32_t>(), n_walls, neg_index, pos_index, ar.g, m->d_ctr, m->stream); });
  if (rc) return rc;
  if (m->h_ctr->flags & kFlagCoordRange)
  {
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    return fail(m, VDBM_ERR_COORD_RANGE, "an artificial wall left the +-2^23 voxel range");
  }
  return VDBM_OK;
}

int vdbm_artificial_areas_add(vdbm_map* m, uint64_t n_polygons, const uint32_t* counts, const double* xyz, double negative_height,
                              double positive_height)
{
  if (!m || (n_polygons && (!counts || !xyz))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  int rc = vdbm_map_integrity_restore(m); // V:1181
  if (rc) return rc;
  return addWalls(m, n_polygons, counts, xyz, negative_height, positive_height, true);
}

int vdbm_artificial_walls_add(vdbm_map* m, uint64_t n_polylines, const uint32_t* counts, const double* xyz, double negative_height,
                              double positive_height, int closed)
{
  if (!m || (n_polylines && (!counts || !xyz))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  return addWalls(m, n_polylines, counts, xyz, negative_height, positive_height, closed != 0);
}

int vdbm_artificial_export(vdbm_map* m, vdbm_leafset** out)
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (!m->artificial)
  {
    *out = newLeafset(m, 0, true, false);
    return VDBM_OK;
  }
  return exportGrid(m, *m->artificial, out);
}

int vdbm_map_export(vdbm_map* m, int dirty_only, vdbm_leafset** out)
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  int rc = syncCounters(m);
  if (rc) return rc;
  const uint32_t n_all = m->n_leaves;
  TempBuf dl(m->stream);
  uint32_t n = n_all;
  if (dirty_only)
  {
    // dirty leaves = flag set by apply_update since the previous export; collect (and clear) them on the device
    CU_TRY(m, dl.alloc(size_t(std::max(1u, n_all)) * 4));
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_out, 0, sizeof(unsigned), m->stream));
    launchCollectDirty(m->mt, n_all, dl.as<uint32_t>(), m->d_ctr, m->stream);
    CU_TRY(m, cudaGetLastError());
    rc = syncCounters(m);
    if (rc) return rc;
    n = std::min(m->h_ctr->n_out, n_all);
  }
  else if (n_all)
    CU_TRY(m, cudaMemsetAsync(m->mt.leaf_dirty, 0, size_t(n_all) * 4, m->stream)); // a full export also leaves nothing dirty
  vdbm_leafset* ls = newLeafset(m, n, false, true);
  *out             = ls;
  if (n)
  {
    TempBuf k0(m->stream), k1(m->stream), i0(m->stream), i1(m->stream), so(m->stream), sa(m->stream), sv(m->stream);
    CU_TRY(m, k0.alloc(size_t(n) * 8)); CU_TRY(m, k1.alloc(size_t(n) * 8));
    CU_TRY(m, i0.alloc(size_t(n) * 4)); CU_TRY(m, i1.alloc(size_t(n) * 4));
    CU_TRY(m, so.alloc(size_t(n) * 12)); CU_TRY(m, sa.alloc(size_t(n) * 64)); CU_TRY(m, sv.alloc(size_t(n) * 2048));
    uint64_t* keys = k0.as<uint64_t>();
    uint32_t* idx  = i0.as<uint32_t>();
    launchKeysFromIdx(m->mt.leaf_keys, dirty_only ? dl.as<uint32_t>() : nullptr, n, keys, idx, m->stream);
    rc = sortByKey(m, keys, idx, k1.as<uint64_t>(), i1.as<uint32_t>(), n);
    if (rc) return rc;
    launchGatherMap(m->mt, n, idx, so.as<int32_t>(), sa.as<uint64_t>(), sv.as<float>(), m->stream);
    CU_TRY(m, cudaMemcpyAsync(ls->origins, so.p, size_t(n) * 12, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaMemcpyAsync(ls->active, sa.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaMemcpyAsync(ls->values, sv.p, size_t(n) * 2048, cudaMemcpyDeviceToHost, m->stream));
  }
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  return VDBM_OK;
}

uint64_t vdbm_map_generation(const vdbm_map* m) { return m ? m->generation : 0; }

// Streams the dirty leaves to the consumer chunk by chunk: gather (HBM -> device chunk buffer) and the D2H copy of the
// next chunks are queued while the consumer works on the current one; three buffers rotate. Bytes per leaf over PCIe:
// 2048 values + 64 active mask + 12 origin + 4 pool index.
int vdbm_map_mirror(vdbm_map* m, uint64_t chunk_leaves, vdbm_mirror_sink sink, void* user, uint64_t* n_leaves_out)
{
  if (!m || !sink) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (n_leaves_out) *n_leaves_out = 0;
  static const bool profile = getenv("VDBM_MIRROR_PROFILE") != nullptr; // experiments: where a mirror call spends its time
  using clk = std::chrono::steady_clock;
  auto ms_since = [](clk::time_point t) { return std::chrono::duration<double, std::milli>(clk::now() - t).count(); };
  const clk::time_point t_begin = clk::now();
  double ms_wait = 0, ms_sink = 0;
  int rc = syncCounters(m);
  if (rc) return rc;
  const uint32_t n_all = m->n_leaves;
  if (n_all == 0) return VDBM_OK;
  TempBuf dl(m->stream), ds(m->stream), tmp(m->stream);
  CU_TRY(m, dl.alloc(size_t(n_all) * 4));
  CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_out, 0, sizeof(unsigned), m->stream));
  launchCollectDirty(m->mt, n_all, dl.as<uint32_t>(), m->d_ctr, m->stream);
  CU_TRY(m, cudaGetLastError());
  rc = syncCounters(m);
  if (rc) return rc;
  const uint32_t n = std::min(m->h_ctr->n_out, n_all);
  if (n == 0) return VDBM_OK;
  // pool-index order: deterministic, new leaves arrive in creation order, the gather walks the pool front to back
  CU_TRY(m, ds.alloc(size_t(n) * 4));
  const size_t sort_bytes = sortKeys32(nullptr, 0, dl.as<uint32_t>(), ds.as<uint32_t>(), n, m->stream);
  CU_TRY(m, tmp.alloc(sort_bytes));
  sortKeys32(tmp.p, sort_bytes, dl.as<uint32_t>(), ds.as<uint32_t>(), n, m->stream);
  CU_TRY(m, cudaGetLastError());
  const uint32_t* d_idx = ds.as<uint32_t>();

  vdbm_map::Mirror& mr = m->mirror;
  const uint32_t want  = uint32_t(std::min<uint64_t>(chunk_leaves ? chunk_leaves : 16384, 1u << 20));
  constexpr size_t kLeafBytes = 2048 + 64 + 12 + 4;
  if (mr.cap_leaves < want)
  {
    for (int b = 0; b < vdbm_map::Mirror::kBufs; ++b)
    {
      cudaFree(mr.d_buf[b]);
      if (mr.h_buf[b]) cudaFreeHost(mr.h_buf[b]);
      mr.d_buf[b] = mr.h_buf[b] = nullptr;
    }
    mr.cap_leaves = 0;
    for (int b = 0; b < vdbm_map::Mirror::kBufs; ++b)
    {
      CU_TRY(m, cudaMalloc(&mr.d_buf[b], size_t(want) * kLeafBytes));
      CU_TRY(m, cudaHostAlloc(reinterpret_cast<void**>(&mr.h_buf[b]), size_t(want) * kLeafBytes, cudaHostAllocDefault));
      if (!mr.ev[b]) CU_TRY(m, cudaEventCreateWithFlags(&mr.ev[b], cudaEventDisableTiming));
    }
    mr.cap_leaves = want;
  }
  const uint32_t C = want; // chunk size of THIS call (buffers may be larger)
  // chunk buffer layout (same on both sides): values [C][512] | active [C][8] | origins [C][3] | index [C]
  auto valuesOf  = [&](uint8_t* p) { return reinterpret_cast<float*>(p); };
  auto activeOf  = [&](uint8_t* p) { return reinterpret_cast<uint64_t*>(p + size_t(C) * 2048); };
  auto originsOf = [&](uint8_t* p) { return reinterpret_cast<int32_t*>(p + size_t(C) * (2048 + 64)); };
  auto indexOf   = [&](uint8_t* p) { return reinterpret_cast<uint32_t*>(p + size_t(C) * (2048 + 64 + 12)); };
  const uint32_t n_chunks = (n + C - 1) / C;
  auto enqueue = [&](uint32_t k) -> cudaError_t {
    const int b        = int(k % vdbm_map::Mirror::kBufs);
    const uint32_t off = k * C, cnt = std::min(C, n - off);
    uint8_t *d = mr.d_buf[b], *h = mr.h_buf[b];
    launchGatherMap(m->mt, cnt, d_idx + off, originsOf(d), activeOf(d), valuesOf(d), m->stream);
    cudaError_t e = cudaGetLastError();
    if (e == cudaSuccess) e = cudaMemcpyAsync(valuesOf(h), valuesOf(d), size_t(cnt) * 2048, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(activeOf(h), activeOf(d), size_t(cnt) * 64, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(originsOf(h), originsOf(d), size_t(cnt) * 12, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaMemcpyAsync(indexOf(h), d_idx + off, size_t(cnt) * 4, cudaMemcpyDeviceToHost, m->stream);
    if (e == cudaSuccess) e = cudaEventRecord(mr.ev[b], m->stream);
    return e;
  };
  const double ms_prepare = ms_since(t_begin);
  for (uint32_t k = 0; k < std::min<uint32_t>(n_chunks, vdbm_map::Mirror::kBufs); ++k) CU_TRY(m, enqueue(k));
  for (uint32_t k = 0; k < n_chunks; ++k)
  {
    const int b        = int(k % vdbm_map::Mirror::kBufs);
    const uint32_t off = k * C, cnt = std::min(C, n - off);
    clk::time_point t0 = clk::now();
    CU_TRY(m, cudaEventSynchronize(mr.ev[b]));
    ms_wait += ms_since(t0);
    uint8_t* h = mr.h_buf[b];
    t0 = clk::now();
    const int sink_rc = sink(user, cnt, indexOf(h), originsOf(h), valuesOf(h), activeOf(h));
    ms_sink += ms_since(t0);
    if (sink_rc != 0)
    {
      // the consumer gave up: what it has not seen stays dirty for the next call
      cudaStreamSynchronize(m->stream);
      launchMarkDirty(m->mt, d_idx + off, n - off, m->stream);
      CU_TRY(m, cudaStreamSynchronize(m->stream));
      return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_map_mirror: the sink aborted the transfer");
    }
    if (n_leaves_out) *n_leaves_out += cnt;
    if (k + vdbm_map::Mirror::kBufs < n_chunks) CU_TRY(m, enqueue(k + vdbm_map::Mirror::kBufs));
  }
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  if (profile)
    std::fprintf(stderr, "vdbm_map_mirror: %u leaves (%.1f MB) in %u chunks: total %.2f ms = prepare %.2f + waiting for copies %.2f + sink %.2f\n", n,
                 double(n) * kLeafBytes / 1e6, n_chunks, ms_since(t_begin), ms_prepare, ms_wait, ms_sink);
  return VDBM_OK;
}

int vdbm_section(vdbm_map* m, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float, vdbm_leafset** out)
{
  if (!m || !out || !bbmin || !bbmax) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  int rc = syncCounters(m);
  if (rc) return rc;
  // upper bound of result leaves: leaves overlapping the box (clamped to the map size)
  uint64_t box_leaves = 1;
  for (int k = 0; k < 3; ++k)
  {
    if (bbmax[k] < bbmin[k]) { box_leaves = 0; break; }
    const uint64_t span = uint64_t((int64_t(bbmax[k]) >> 3) - (int64_t(bbmin[k]) >> 3) + 1);
    box_leaves          = std::min<uint64_t>(box_leaves * span, m->n_leaves);
  }
  const uint32_t cap = uint32_t(std::min<uint64_t>(box_leaves, m->n_leaves));
  if (cap == 0)
  {
    *out = newLeafset(m, 0, !result_float, result_float != 0);
    return VDBM_OK;
  }
  TempBuf dk(m->stream), da(m->stream), dv(m->stream), df(m->stream);
  CU_TRY(m, dk.alloc(size_t(cap) * 8));
  CU_TRY(m, da.alloc(size_t(cap) * 64));
  if (result_float) CU_TRY(m, df.alloc(size_t(cap) * 2048));
  else CU_TRY(m, dv.alloc(size_t(cap) * 64));
  CU_TRY(m, cudaMemsetAsync(&m->d_ctr->n_out, 0, sizeof(unsigned), m->stream));
  launchSection(m->mt, m->n_leaves, bbmin, bbmax, full, result_float, dk.as<uint64_t>(), da.as<uint64_t>(), dv.as<uint64_t>(),
                df.as<float>(), cap, m->d_ctr, m->stream);
  CU_TRY(m, cudaGetLastError());
  rc = syncCounters(m);
  if (rc) return rc;
  const uint32_t n = std::min(m->h_ctr->n_out, cap);
  vdbm_leafset* ls = newLeafset(m, n, !result_float, result_float != 0);
  *out             = ls;
  if (n == 0) return VDBM_OK;
  // sort (key, row) on the device, permute the rows into key order, copy straight into the leaf set
  TempBuf k1(m->stream), i0(m->stream), i1(m->stream), so(m->stream), pa(m->stream), pv(m->stream), pf(m->stream);
  CU_TRY(m, k1.alloc(size_t(n) * 8)); CU_TRY(m, i0.alloc(size_t(n) * 4)); CU_TRY(m, i1.alloc(size_t(n) * 4));
  CU_TRY(m, so.alloc(size_t(n) * 12)); CU_TRY(m, pa.alloc(size_t(n) * 64));
  if (result_float) CU_TRY(m, pf.alloc(size_t(n) * 2048));
  else CU_TRY(m, pv.alloc(size_t(n) * 64));
  uint64_t* keys = dk.as<uint64_t>();
  uint32_t* idx  = i0.as<uint32_t>();
  launchKeysFromIdx(dk.as<uint64_t>(), nullptr, n, k1.as<uint64_t>(), idx, m->stream); // identity permutation (+ key copy)
  keys              = k1.as<uint64_t>();
  uint64_t* keys_alt = dk.as<uint64_t>();
  rc = sortByKey(m, keys, idx, keys_alt, i1.as<uint32_t>(), n);
  if (rc) return rc;
  launchPermuteSection(n, keys, idx, da.as<uint64_t>(), result_float ? nullptr : dv.as<uint64_t>(), result_float ? df.as<float>() : nullptr,
                       so.as<int32_t>(), pa.as<uint64_t>(), pv.as<uint64_t>(), pf.as<float>(), m->stream);
  CU_TRY(m, cudaGetLastError());
  CU_TRY(m, cudaMemcpyAsync(ls->origins, so.p, size_t(n) * 12, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(ls->active, pa.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  if (result_float) CU_TRY(m, cudaMemcpyAsync(ls->values, pf.p, size_t(n) * 2048, cudaMemcpyDeviceToHost, m->stream));
  else CU_TRY(m, cudaMemcpyAsync(ls->valmask, pv.p, size_t(n) * 64, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  return VDBM_OK;
}

namespace {
// leaf keys of host leaf origins (validated); sorted-unique parent block keys at a coarser alignment
int originsToKeys(vdbm_map* m, uint64_t n, const int32_t* origins, std::vector<uint64_t>& keys)
{
  keys.resize(n);
  for (uint64_t i = 0; i < n; ++i)
  {
    for (int k = 0; k < 3; ++k)
      if (std::abs(int64_t(origins[3 * i + k])) >= kVoxelLimit) return fail(m, VDBM_ERR_COORD_RANGE, "leaf origin outside the +-2^23 voxel range");
    keys[i] = packLeafKey(origins[3 * i] >> 3, origins[3 * i + 1] >> 3, origins[3 * i + 2] >> 3);
  }
  return VDBM_OK;
}
std::vector<uint64_t> parentBlocks(uint64_t n, const int32_t* origins, int log2dim)
{
  std::vector<uint64_t> b(n);
  for (uint64_t i = 0; i < n; ++i)
    b[i] = packLeafKey(((origins[3 * i] >> log2dim) << log2dim) >> 3, ((origins[3 * i + 1] >> log2dim) << log2dim) >> 3,
                       ((origins[3 * i + 2] >> log2dim) << log2dim) >> 3);
  std::sort(b.begin(), b.end());
  b.erase(std::unique(b.begin(), b.end()), b.end());
  return b;
}
} // namespace

int vdbm_section_apply_update(vdbm_map* m, const int32_t bbmin[3], const int32_t bbmax[3], uint64_t n, const int32_t* origins,
                              const uint64_t* active)
{
  if (!m || !bbmin || !bbmax || (n && (!origins || !active))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  int rc = syncCounters(m);
  if (rc) return rc;
  // validate the section and make room BEFORE the box is deactivated: a rejected call must leave the map untouched
  std::vector<uint64_t> keys;
  if (n)
  {
    rc = originsToKeys(m, n, origins, keys);
    if (rc) return rc;
    rc = ensureMapCapacity(m, n);
    if (rc) return rc;
  }
  launchSectionDeactivate(m->mt, m->n_leaves, bbmin, bbmax, m->stream);
  if (n)
  {
    TempBuf dk(m->stream), da(m->stream);
    CU_TRY(m, dk.alloc(n * 8));
    CU_TRY(m, da.alloc(n * 64));
    CU_TRY(m, cudaMemcpyAsync(dk.p, keys.data(), n * 8, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(da.p, active, n * 64, cudaMemcpyHostToDevice, m->stream));
    launchSectionActivate(m->mt, dk.as<uint64_t>(), da.as<uint64_t>(), uint32_t(n), m->d_ctr, m->stream);
    CU_TRY(m, cudaGetLastError());
    rc = syncCounters(m); // the staged host vectors must outlive the copies
    if (rc) return rc;
  }
  return syncCounters(m);
}

int vdbm_section_apply_grid(vdbm_map* m, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int tile_quirk)
{
  if (!m || (n && (!origins || !active || !values))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (n == 0) return VDBM_OK;
  int rc = syncCounters(m);
  if (rc) return rc;
  std::vector<uint64_t> keys;
  rc = originsToKeys(m, n, origins, keys);
  if (rc) return rc;
  rc = ensureMapCapacity(m, n);
  if (rc) return rc;
  TempBuf dk(m->stream), da(m->stream), dv(m->stream), db0(m->stream), db1(m->stream), ds(m->stream);
  CU_TRY(m, dk.alloc(n * 8));
  CU_TRY(m, da.alloc(n * 64));
  CU_TRY(m, dv.alloc(n * 2048));
  CU_TRY(m, cudaMemcpyAsync(dk.p, keys.data(), n * 8, cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaMemcpyAsync(da.p, active, n * 64, cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaMemcpyAsync(dv.p, values, n * 2048, cudaMemcpyHostToDevice, m->stream));
  std::vector<uint64_t> sorted_keys, i1, i2;
  if (tile_quirk)
  {
    // tiles first (they only touch voxels outside the section leaves, so the order does not matter)
    sorted_keys = keys;
    std::sort(sorted_keys.begin(), sorted_keys.end());
    i1 = parentBlocks(n, origins, 7);   // 128^3 blocks = Internal<4> nodes of the section tree
    i2 = parentBlocks(n, origins, 12);  // 4096^3 blocks = Internal<5> nodes
    CU_TRY(m, ds.alloc(sorted_keys.size() * 8));
    CU_TRY(m, db0.alloc(i1.size() * 8));
    CU_TRY(m, db1.alloc(i2.size() * 8));
    CU_TRY(m, cudaMemcpyAsync(ds.p, sorted_keys.data(), sorted_keys.size() * 8, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(db0.p, i1.data(), i1.size() * 8, cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaMemcpyAsync(db1.p, i2.data(), i2.size() * 8, cudaMemcpyHostToDevice, m->stream));
    launchSectionTileQuirk(m->mt, db0.as<uint64_t>(), uint32_t(i1.size()), 0, ds.as<uint64_t>(), uint32_t(sorted_keys.size()), m->stream);
    launchSectionTileQuirk(m->mt, db1.as<uint64_t>(), uint32_t(i2.size()), 1, db0.as<uint64_t>(), uint32_t(i1.size()), m->stream);
  }
  launchSectionApplyGrid(m->mt, dk.as<uint64_t>(), da.as<uint64_t>(), dv.as<float>(), uint32_t(n), m->d_ctr, m->stream);
  CU_TRY(m, cudaGetLastError());
  return syncCounters(m);
}

int vdbm_map_import(vdbm_map* m, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int replace)
{
  if (!m || (n && (!origins || !active || !values))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (replace)
  {
    // loadMap V:263-284 replaces m_vdb_grid: forget every map leaf (the update grids of the sources are left alone)
    CU_TRY(m, cudaMemsetAsync(m->mt.hkeys, 0xFF, size_t(m->hcap) * 8, m->stream));
    CU_TRY(m, cudaMemsetAsync(m->mt.leaf_dirty, 0, size_t(m->mt.pool_cap) * 4, m->stream));
    CU_TRY(m, cudaMemsetAsync(m->d_map_counters, 0, 8, m->stream));
    m->n_leaves = 0;
    m->coarse_built = ~0u;
    ++m->generation;
    int rc = syncCounters(m);
    if (rc) return rc;
  }
  // leaf for leaf: values and active mask of the given leaf replace the map's (a leaf without any content is not created)
  return vdbm_section_apply_grid(m, n, origins, active, values, 0);
}

int vdbm_cast_index_rays(vdbm_map* m, const char* source_id, uint64_t n_rays, const int32_t* rays6)
{
  if (!m || (n_rays && !rays6) || n_rays > 0xFFFFFFF0ull) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  if (n_rays == 0) return VDBM_OK;
  TempBuf d(m->stream);
  CU_TRY(m, d.alloc(n_rays * 24));
  CU_TRY(m, cudaMemcpyAsync(d.p, rays6, n_rays * 24, cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  // castRayIntoGrid V:550-566 for explicit voxel pairs: the wall kernel with one height level is exactly that loop
  int rc = markIntoGrid(m, *s, [&] { launchWallDDA(d.as<int32_t>(), uint32_t(n_rays), 0, 1, s->g, m->d_ctr, m->stream); });
  if (rc) return rc;
  if (m->h_ctr->flags & kFlagCoordRange)
  {
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    return fail(m, VDBM_ERR_COORD_RANGE, "a ray left the +-2^23 voxel range");
  }
  return VDBM_OK;
}

int vdbm_set_fast_mode(vdbm_map* m, int on)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (on && (m->ex.connected || m->sector_n > 1))
    return fail(m, VDBM_ERR_INVALID_ARG, "fast_mode probes the map along every ray: it needs the whole map on one device (this handle is a shard)");
  m->fast_mode = on != 0;
  return VDBM_OK;
}

int vdbm_raytrace(vdbm_map* m, uint64_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* success,
                  double* end_points)
{
  if (!m || (n && (!origins || !directions || !max_lengths || !success || !end_points)) || n > 0xFFFFFFF0ull) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (n == 0) return VDBM_OK;
  int rc = syncCounters(m);
  if (rc) return rc;
  if (m->n_leaves != 0)
  {
    rc = ensureCoarse(m);
    if (rc) return rc;
  }
  else if (!m->d_bbox)
  {
    rc = allocCoarse(m, 1u << 14, 1u << 10);
    if (rc) return rc;
  }
  TempBuf in(m->stream), out(m->stream);
  CU_TRY(m, in.alloc(n * 7 * sizeof(double)));
  CU_TRY(m, out.alloc(n * (3 * sizeof(double) + sizeof(int32_t))));
  double* d_o = in.as<double>();
  double* d_d = d_o + 3 * n;
  double* d_l = d_d + 3 * n;
  double* d_e = out.as<double>();
  int32_t* d_ok = reinterpret_cast<int32_t*>(d_e + 3 * n);
  CU_TRY(m, cudaMemcpyAsync(d_o, origins, n * 3 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaMemcpyAsync(d_d, directions, n * 3 * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  CU_TRY(m, cudaMemcpyAsync(d_l, max_lengths, n * sizeof(double), cudaMemcpyHostToDevice, m->stream));
  launchRaytrace(n, d_o, d_d, d_l, m->params.resolution, 1.0 / m->params.resolution, m->mt, m->coarse, m->d_bbox, m->n_leaves == 0 ? 1u : 0u, d_ok,
                 d_e, m->d_ctr, m->stream);
  CU_TRY(m, cudaMemcpyAsync(end_points, d_e, n * 3 * sizeof(double), cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaMemcpyAsync(success, d_ok, n * sizeof(int32_t), cudaMemcpyDeviceToHost, m->stream));
  rc = syncCounters(m);
  if (rc) return rc;
  if (m->h_ctr->flags & kFlagCoordRange)
  {
    m->h_ctr->flags &= ~kFlagCoordRange;
    CU_TRY(m, cudaMemcpyAsync(&m->d_ctr->flags, &m->h_ctr->flags, sizeof(unsigned), cudaMemcpyHostToDevice, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    return fail(m, VDBM_ERR_COORD_RANGE, "some rays left the +-2^23 voxel range and were reported as misses");
  }
  return VDBM_OK;
}

int vdbm_probe(vdbm_map* m, const int32_t xyz[3], float* value, int32_t* active)
{
  if (!m || !xyz || !value || !active) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  TempBuf d(m->stream);
  CU_TRY(m, d.alloc(16));
  launchProbe(m->mt, xyz[0], xyz[1], xyz[2], d.as<float>(), reinterpret_cast<int32_t*>(d.as<float>() + 1), m->stream);
  CU_TRY(m, cudaMemcpyAsync(m->h_small + 16, d.p, 8, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  std::memcpy(value, m->h_small + 16, 4);
  *active = int32_t(m->h_small[17]);
  return VDBM_OK;
}

uint64_t vdbm_leafset_size(const vdbm_leafset* s) { return s ? s->n : 0; }
const int32_t* vdbm_leafset_origins(const vdbm_leafset* s) { return s ? s->origins : nullptr; }
const uint64_t* vdbm_leafset_active(const vdbm_leafset* s) { return s ? s->active : nullptr; }
const uint64_t* vdbm_leafset_valmask(const vdbm_leafset* s) { return s ? s->valmask : nullptr; }
const float* vdbm_leafset_values(const vdbm_leafset* s) { return s ? s->values : nullptr; }
void vdbm_leafset_free(vdbm_leafset* s)
{
  if (!s) return;
  if (s->lender)
  {
    s->lender->h_stage_lent     = false; // arrays belong to the handle's staging buffer
    s->lender->h_stage_borrower = nullptr;
    delete s;
    return;
  }
  if (s->orphan_stage)
  {
    cudaFreeHost(s->orphan_stage);
    delete s;
    return;
  }
  auto rel = [&](void* p) {
    if (!p) return;
    if (s->pinned) cudaFreeHost(p);
    else std::free(p);
  };
  rel(s->origins); rel(s->active); rel(s->valmask); rel(s->values);
  delete s;
}

int32_t vdbm_leaf_owner(const int32_t origin[3], int32_t n_ranks)
{
  if (!origin || n_ranks <= 0) return -1;
  return leafOwner(packLeafKey(origin[0] >> 3, origin[1] >> 3, origin[2] >> 3), n_ranks);
}

int vdbm_shard_plan_set(vdbm_map* m, int32_t mode, int32_t n_ranks, const int32_t center_leaf_xy[2], const double* bounds)
{
  if (!m || (mode != 0 && mode != 1) || n_ranks < 1 || n_ranks > kMaxRanks) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  ShardPlan sp{};
  sp.mode = mode; sp.n_ranks = n_ranks;
  if (mode == 1)
  {
    if (!center_leaf_xy || !bounds) return VDBM_ERR_INVALID_ARG;
    sp.cx = center_leaf_xy[0]; sp.cy = center_leaf_xy[1];
    for (int r = 0; r < n_ranks; ++r)
    {
      if (!(bounds[r] >= 0.0 && bounds[r] < 4.0) || (r && !(bounds[r] > bounds[r - 1])))
        return fail(m, VDBM_ERR_INVALID_ARG, "sector bounds must be ascending diamond angles in [0, 4)");
      sp.bounds[r] = bounds[r];
    }
  }
  if (m->n_leaves != 0 && std::memcmp(&sp, &m->shard, sizeof(sp)) != 0)
    return fail(m, VDBM_ERR_INVALID_ARG, "the shard plan cannot change while the map holds leaves (reset the map first)");
  m->shard = sp;
  return VDBM_OK;
}

int vdbm_ray_sector_set(vdbm_map* m, int32_t n_ranks, int32_t rank, const double* bounds)
{
  if (!m || n_ranks < 0 || n_ranks > kMaxRanks || (n_ranks > 1 && (!bounds || rank < 0 || rank >= n_ranks))) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (n_ranks > 1 && m->fast_mode) return fail(m, VDBM_ERR_INVALID_ARG, "fast_mode needs the whole map on one device: switch it off before sharding");
  if (n_ranks > 1)
    for (int r = 0; r < n_ranks; ++r)
      if (!(bounds[r] >= 0.0 && bounds[r] < 4.0) || (r && !(bounds[r] > bounds[r - 1])))
        return fail(m, VDBM_ERR_INVALID_ARG, "sector bounds must be ascending diamond angles in [0, 4)");
  m->sector_n    = n_ranks > 1 ? n_ranks : 0;
  m->sector_rank = rank;
  for (int r = 0; r < m->sector_n; ++r) m->sector_bounds[r] = bounds[r];
  return VDBM_OK;
}

int32_t vdbm_leaf_owner_planned(vdbm_map* m, const int32_t origin[3], int32_t n_ranks)
{
  if (!m || !origin || n_ranks <= 0) return -1;
  ShardPlan plan = m->shard;
  if (plan.mode == 0) plan.n_ranks = n_ranks;
  return leafOwnerPlanned(packLeafKey(origin[0] >> 3, origin[1] >> 3, origin[2] >> 3), plan);
}

int vdbm_map_checksum(vdbm_map* m, uint64_t out2[2])
{
  if (!m || !out2) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  int rc = syncCounters(m);
  if (rc) return rc;
  TempBuf d(m->stream);
  CU_TRY(m, d.alloc(16));
  CU_TRY(m, cudaMemsetAsync(d.p, 0, 16, m->stream));
  launchMapChecksum(m->mt, m->n_leaves, d.as<unsigned long long>(), m->stream);
  CU_TRY(m, cudaGetLastError());
  CU_TRY(m, cudaMemcpyAsync(m->h_small + 32, d.p, 16, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  std::memcpy(out2, m->h_small + 32, 16);
  return VDBM_OK;
}

int vdbm_update_partition(vdbm_map* m, const char* source_id, int32_t n_ranks, uint64_t* counts, const void** d_records)
{
  if (!m || !counts || !d_records || n_ranks <= 0 || n_ranks > 32) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  const uint32_t n = s->n_entries;
  for (int r = 0; r < n_ranks; ++r) counts[r] = 0;
  *d_records = m->d_part;
  if (n == 0)
  {
    if (s->n_bricks) launchResetBricks(s->g, s->n_bricks, m->stream);
    s->n_bricks = 0;
    return VDBM_OK;
  }
  if (m->part_cap < n)
  {
    cudaFree(m->d_part);
    m->d_part   = nullptr;
    m->part_cap = 0;
    const size_t want = size_t(n) + n / 4 + 1024;
    CU_TRY(m, cudaMalloc(&m->d_part, want * sizeof(LeafRecord)));
    m->part_cap = want;
  }
  *d_records = m->d_part;
  TempBuf d(m->stream);
  CU_TRY(m, d.alloc(64 * 4));
  uint32_t* d_counts = d.as<uint32_t>();
  uint32_t* d_cursor = d_counts + 32;
  CU_TRY(m, cudaMemsetAsync(d_counts, 0, 64 * 4, m->stream));
  ShardPlan plan = m->shard;
  if (plan.mode == 0) plan.n_ranks = n_ranks;
  else if (plan.n_ranks != n_ranks) return fail(m, VDBM_ERR_INVALID_ARG, "n_ranks differs from the shard plan set with vdbm_shard_plan_set");
  launchPartition(s->g, n, plan, d_counts, d_cursor, m->d_part, 0, m->stream);
  CU_TRY(m, cudaMemcpyAsync(m->h_small + 24, d_counts, 32 * 4, cudaMemcpyDeviceToHost, m->stream));
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  uint32_t off[32];
  uint32_t acc = 0;
  for (int r = 0; r < 32; ++r)
  {
    off[r] = acc;
    if (r < n_ranks)
    {
      counts[r] = m->h_small[24 + r];
      acc += m->h_small[24 + r];
    }
  }
  CU_TRY(m, cudaMemcpyAsync(d_cursor, off, 32 * 4, cudaMemcpyHostToDevice, m->stream));
  launchPartition(s->g, n, plan, d_counts, d_cursor, m->d_part, 1, m->stream);
  launchResetBricks(s->g, s->n_bricks, m->stream);
  CU_TRY(m, cudaGetLastError());
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  s->n_bricks = s->n_entries = 0;
  return VDBM_OK;
}

int vdbm_exchange_create(vdbm_map* m, int32_t rank, int32_t n_ranks, uint64_t cap, void* handles_out)
{
  if (!m || !handles_out || n_ranks < 1 || n_ranks > kMaxRanks || rank < 0 || rank >= n_ranks || cap == 0 || cap > 0xFFFFFFu)
    return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  if (m->ex.created) return fail(m, VDBM_ERR_INVALID_ARG, "exchange already created on this handle");
  auto& ex = m->ex;
  const size_t inbox_bytes = inboxBytes(2u * uint32_t(n_ranks), uint32_t(cap));
  CU_TRY(m, cudaMalloc(&ex.inbox, inbox_bytes));
  CU_TRY(m, cudaMalloc(&ex.ctrl, size_t(2) * kMaxRanks * sizeof(unsigned long long)));
  CU_TRY(m, cudaMalloc(&ex.d_cursors, kMaxRanks * sizeof(uint32_t)));
  CU_TRY(m, cudaMalloc(&ex.d_counts, kMaxRanks * sizeof(uint32_t)));
  CU_TRY(m, cudaMemset(ex.ctrl, 0, size_t(2) * kMaxRanks * sizeof(unsigned long long)));
  CU_TRY(m, cudaMemset(ex.d_counts, 0, kMaxRanks * sizeof(uint32_t)));
  ex.px.cap = uint32_t(cap); ex.px.n_ranks = n_ranks; ex.px.rank = rank;
  cudaIpcMemHandle_t h[2];
  CU_TRY(m, cudaIpcGetMemHandle(&h[0], ex.inbox));
  CU_TRY(m, cudaIpcGetMemHandle(&h[1], ex.ctrl));
  static_assert(sizeof(h) == VDBM_IPC_HANDLE_BYTES, "two CUDA IPC handles");
  std::memcpy(handles_out, h, sizeof(h));
  for (auto& e : ex.ev) CU_TRY(m, cudaEventCreate(&e));
  ex.created = true;
  return VDBM_OK;
}

int vdbm_exchange_timings(vdbm_map* m, float* out3)
{
  if (!m || !out3) return VDBM_ERR_INVALID_ARG;
  if (m->ex.created && m->ex.epoch > 0 && cudaEventSynchronize(m->ex.ev[1]) == cudaSuccess)
    cudaEventElapsedTime(&m->ex.ms[0], m->ex.ev[0], m->ex.ev[1]); // valid even when no pull followed (profiling)
  out3[0] = m->ex.ms[0]; out3[1] = m->ex.ms[1]; out3[2] = m->ex.ms[2];
  return VDBM_OK;
}

int vdbm_exchange_connect(vdbm_map* m, const void* all_handles)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  auto& ex = m->ex;
  if (!ex.created) return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_exchange_create first");
  if (!all_handles)
  {
    // loop-back (single process, for profiling the bin-and-send kernel without NVLink): every "peer" is this rank's inbox.
    // Only the region of sender `rank` is ever written, so the data of all owners land in the same region: counts and
    // contents are meaningless, timings are not.
    for (int r = 0; r < ex.px.n_ranks; ++r) { ex.px.inbox[r] = ex.inbox; ex.px.ctrl[r] = ex.ctrl; }
    ex.connected = true;
    return VDBM_OK;
  }
  const auto* hs = static_cast<const unsigned char*>(all_handles);
  for (int r = 0; r < ex.px.n_ranks; ++r)
  {
    if (r == ex.px.rank)
    {
      ex.px.inbox[r] = ex.inbox;
      ex.px.ctrl[r]  = ex.ctrl;
      continue;
    }
    cudaIpcMemHandle_t h[2];
    std::memcpy(h, hs + size_t(r) * VDBM_IPC_HANDLE_BYTES, sizeof(h));
    void *pi = nullptr, *pc = nullptr;
    CU_TRY(m, cudaIpcOpenMemHandle(&pi, h[0], cudaIpcMemLazyEnablePeerAccess));
    CU_TRY(m, cudaIpcOpenMemHandle(&pc, h[1], cudaIpcMemLazyEnablePeerAccess));
    ex.opened.push_back(pi);
    ex.opened.push_back(pc);
    ex.px.inbox[r] = static_cast<uint64_t*>(pi);
    ex.px.ctrl[r]  = static_cast<unsigned long long*>(pc);
  }
  ex.connected = true;
  return VDBM_OK;
}

int vdbm_exchange_connect_peers(vdbm_map* m, vdbm_map* const* peers)
{
  if (!m || !peers) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  auto& ex = m->ex;
  if (!ex.created) return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_exchange_create first");
  for (int r = 0; r < ex.px.n_ranks; ++r)
  {
    vdbm_map* p = (r == ex.px.rank) ? m : peers[r];
    if (!p || !p->ex.created || p->ex.px.n_ranks != ex.px.n_ranks || p->ex.px.cap != ex.px.cap || p->ex.px.rank != r)
      return fail(m, VDBM_ERR_INVALID_ARG, "peer handle missing or its exchange was created with another rank / size / capacity");
    if (p->device != m->device)
    {
      int can = 0;
      CU_TRY(m, cudaDeviceCanAccessPeer(&can, m->device, p->device));
      if (!can) return fail(m, VDBM_ERR_CUDA, "no peer access between the devices of this group");
      const cudaError_t e = cudaDeviceEnablePeerAccess(p->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled)
      {
        m->last_error = std::string("cudaDeviceEnablePeerAccess: ") + cudaGetErrorString(e);
        return VDBM_ERR_CUDA;
      }
      cudaGetLastError(); // clear "already enabled"
    }
    ex.px.inbox[r] = p->ex.inbox;
    ex.px.ctrl[r]  = p->ex.ctrl;
  }
  ex.connected = true;
  return VDBM_OK;
}

int vdbm_update_push(vdbm_map* m, const char* source_id)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  auto& ex = m->ex;
  if (!ex.connected) return fail(m, VDBM_ERR_INVALID_ARG, "exchange not connected");
  ShardPlan plan = m->shard;
  if (plan.mode == 0) plan.n_ranks = ex.px.n_ranks;
  else if (plan.n_ranks != ex.px.n_ranks) return fail(m, VDBM_ERR_INVALID_ARG, "shard plan and exchange disagree on the number of ranks");
  ex.epoch += 1;
  CU_TRY(m, cudaEventRecord(ex.ev[0], m->stream));
  launchPushUpdate(s->g, s->n_entries, ex.px, plan, ex.epoch & 1u, ex.epoch, ex.d_cursors, m->d_ctr, m->stream);
  CU_TRY(m, cudaEventRecord(ex.ev[1], m->stream));
  CU_TRY(m, cudaGetLastError());
  // The leaves this rank owns stay in the grid (the foreign ones were sent and zeroed): the host counts are upper bounds
  // until vdbm_update_pull* rebuilds the leaf list on top of the imported records.
  return VDBM_OK;
}

int vdbm_update_pull(vdbm_map* m, const char* source_id)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  auto& ex = m->ex;
  if (!ex.connected || ex.epoch == 0) return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_update_push first");
  launchWaitPeers(ex.ctrl, ex.px.n_ranks, ex.epoch & 1u, ex.epoch, ex.d_counts, m->d_ctr, m->stream);
  CU_TRY(m, cudaEventRecord(ex.ev[2], m->stream));
  for (int attempt = 0; attempt < 24; ++attempt)
  {
    launchPullUpdate(s->g, ex.inbox, ex.ctrl, ex.px.cap, ex.px.n_ranks, ex.epoch & 1u, ex.epoch, ex.d_counts, m->d_ctr, m->stream);
    launchCompactLeaves(s->g, m->stream);
    CU_TRY(m, cudaEventRecord(ex.ev[3], m->stream));
    CU_TRY(m, cudaGetLastError());
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s->g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
    int rc = syncCounters(m);
    if (rc) return rc;
    s->n_bricks  = m->h_small[8];
    s->n_entries = m->h_small[9];
    const uint32_t flags = m->h_ctr->flags;
    if (flags & (kFlagExchangeOverflow | kFlagExchangeTimeout))
    {
      CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
      return fail(m, VDBM_ERR_OUT_OF_MEMORY, (flags & kFlagExchangeTimeout) ? "exchange: a peer did not publish its epoch in time"
                                                                            : "exchange: inbox capacity_records_per_sender too small");
    }
    const bool overflow = (flags & kFlagUpdateOverflow) != 0;
    const bool crowded  = uint64_t(s->n_bricks) * 10 > uint64_t(s->cap) * 7;
    if (!overflow && !crowded) break;
    if (overflow) CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    rc = growUpdateGrid(m, *s);
    if (rc) return rc;
    if (!overflow) break; // the inbox still holds this epoch's records: a replay of the pull is idempotent
  }
  m->stats.last_touched_leaves = s->n_entries;
  cudaEventElapsedTime(&ex.ms[0], ex.ev[0], ex.ev[1]);
  cudaEventElapsedTime(&ex.ms[1], ex.ev[1], ex.ev[2]);
  cudaEventElapsedTime(&ex.ms[2], ex.ev[2], ex.ev[3]);
  return VDBM_OK;
}

// vdbm_update_pull + vdbm_integrate with ONE host synchronisation: wait for the peers, OR their records in, rebuild the
// leaf list and run the update behind the device-side guard (see vdbm_insert_async). If the guard refuses (the import
// overflowed or crowded the brick hash, or the map must grow) the host falls back to the two synchronous calls; the
// inbox still holds this epoch's records and importing them again is idempotent.
int vdbm_update_pull_integrate(vdbm_map* m, const char* source_id)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  Source* s = findSource(m, source_id);
  if (!s) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, std::string("Source not available: ") + (source_id ? source_id : ""));
  auto& ex = m->ex;
  if (!ex.connected || ex.epoch == 0) return fail(m, VDBM_ERR_INVALID_ARG, "vdbm_update_push first");
  bool other_data = false;
  for (auto& kv : m->sources)
    if (kv.second.get() != s && (kv.second->n_entries || kv.second->n_bricks)) other_data = true;
  const bool fused = !other_data && !(m->artificial && m->artificial->n_entries) && m->async_expect != 0;
  if (!fused)
  {
    int rc = vdbm_update_pull(m, source_id);
    if (rc) return rc;
    rc = vdbm_integrate(m, 0);
    m->async_expect = uint32_t(std::max<uint64_t>(1, m->stats.last_touched_leaves));
    return rc;
  }
  const uint32_t expect = m->async_expect + m->async_expect / 2 + 65536;
  int rc = ensureMapCapacity(m, expect);
  if (rc) return rc;
  rc = ensureResolved(m, expect);
  if (rc) return rc;
  const uint64_t upd_before = m->stats.voxel_updates;
  launchWaitPeers(ex.ctrl, ex.px.n_ranks, ex.epoch & 1u, ex.epoch, ex.d_counts, m->d_ctr, m->stream);
  CU_TRY(m, cudaEventRecord(ex.ev[2], m->stream));
  launchPullUpdate(s->g, ex.inbox, ex.ctrl, ex.px.cap, ex.px.n_ranks, ex.epoch & 1u, ex.epoch, ex.d_counts, m->d_ctr, m->stream);
  launchCompactLeaves(s->g, m->stream);
  CU_TRY(m, cudaEventRecord(ex.ev[3], m->stream));
  CU_TRY(m, cudaEventRecord(m->ev0, m->stream));
  launchApplyUpdateDeferred(s->g, m->mt, m->lo, m->d_resolved, uint32_t(std::min<size_t>(m->resolved_cap, 0xFFFFFFFFu)), m->d_ctr, expect, m->stream);
  CU_TRY(m, cudaEventRecord(m->ev1, m->stream));
  CU_TRY(m, cudaGetLastError());
  rc = syncCounters(m);
  if (rc) return rc;
  cudaEventElapsedTime(&ex.ms[0], ex.ev[0], ex.ev[1]);
  cudaEventElapsedTime(&ex.ms[1], ex.ev[1], ex.ev[2]);
  cudaEventElapsedTime(&ex.ms[2], ex.ev[2], ex.ev[3]);
  const uint32_t flags = m->h_ctr->flags;
  if (flags & (kFlagExchangeOverflow | kFlagExchangeTimeout))
  {
    CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    return fail(m, VDBM_ERR_OUT_OF_MEMORY, (flags & kFlagExchangeTimeout) ? "exchange: a peer did not publish its epoch in time"
                                                                          : "exchange: inbox capacity_records_per_sender too small");
  }
  if (m->h_ctr->deferred_skip)
  {
    // redo on the synchronous path: the records are still in the inbox, the OR is idempotent
    if (flags & kFlagUpdateOverflow) CU_TRY(m, cudaMemsetAsync(&m->d_ctr->flags, 0, sizeof(unsigned), m->stream));
    CU_TRY(m, cudaMemcpyAsync(m->h_small + 8, s->g.counters, 8, cudaMemcpyDeviceToHost, m->stream));
    CU_TRY(m, cudaStreamSynchronize(m->stream));
    s->n_bricks  = m->h_small[8];
    s->n_entries = m->h_small[9];
    if (flags & kFlagUpdateOverflow)
    {
      rc = growUpdateGrid(m, *s);
      if (rc) return rc;
    }
    rc = vdbm_update_pull(m, source_id);
    if (rc) return rc;
    rc = vdbm_integrate(m, 0);
    m->async_expect = uint32_t(std::max<uint64_t>(1, m->stats.last_touched_leaves));
    return rc;
  }
  float ms = 0.f;
  cudaEventElapsedTime(&ms, m->ev0, m->ev1);
  m->stats.last_integrate_ms   = ms;
  m->stats.last_touched_leaves = m->h_ctr->deferred_entries;
  m->stats.last_voxel_updates  = m->stats.voxel_updates - upd_before;
  m->async_expect              = std::max(1u, m->h_ctr->deferred_entries);
  s->n_bricks = s->n_entries = 0;
  s->n_change = 0;
  return VDBM_OK;
}

int vdbm_pipeline_counts(vdbm_map* m, uint64_t out4[4])
{
  if (!m || !out4) return VDBM_ERR_INVALID_ARG;
  out4[0] = m->async_fast;
  out4[1] = m->async_redone;
  out4[2] = m->async_sync;
  out4[3] = m->async_overlapped;
  return VDBM_OK;
}

int vdbm_stats(vdbm_map* m, vdbm_stats_t* out)
{
  if (!m || !out) return VDBM_ERR_INVALID_ARG;
  m->stats.gpu_launches = launchCount();
  *out                  = m->stats;
  return VDBM_OK;
}

const char* vdbm_last_error(vdbm_map* m) { return m ? m->last_error.c_str() : "null handle"; }

int vdbm_synchronize(vdbm_map* m)
{
  if (!m) return VDBM_ERR_INVALID_ARG;
  VDBM_ENTER(m);
  CU_TRY(m, cudaStreamSynchronize(m->stream));
  return VDBM_OK;
}

void* vdbm_host_alloc(size_t bytes)
{
  void* p = nullptr;
  if (cudaHostAlloc(&p, bytes ? bytes : 8, cudaHostAllocDefault) != cudaSuccess) return nullptr;
  return p;
}
void vdbm_host_free(void* p)
{
  if (p) cudaFreeHost(p);
}

} // extern "C"
