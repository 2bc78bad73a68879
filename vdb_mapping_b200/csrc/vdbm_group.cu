// vdbm_group.cu — ONE process driving a sharded map on several GPUs of one NVLink / NVSwitch box.
//
// The reference integrates one map in one process (integrateUpdate, VDBMapping.hpp:375-387), and a ROS node that links the
// drop-in library is one process too: the multi-process layout of vdb_mapping_b200/dist.py (one rank per GPU, CUDA IPC,
// torch.distributed for the rendezvous) is out of its reach. A vdbm_group is the same machine behind the C ABI:
//   * one vdbm_map per device, each driven by its own worker thread (CUDA's current device is per thread);
//   * the shards' inboxes are wired with direct peer pointers (cudaDeviceEnablePeerAccess), the exchange itself is the
//     fused bin-and-send / device-side wait / import of vdbm_update_push + vdbm_update_pull_integrate, unchanged;
//   * every shard receives the WHOLE cloud and keeps the rays of its azimuth sector on the device (vdbm_ray_sector_set):
//     no host-side split, no per-rank cloud;
//   * the sector plan (ray sectors with equal raycast cost, ownership sectors with equal owned leaves) is cut from the first
//     scan of an empty map, from a dry-run raycast on shard 0 - the C++ twin of dist.plan_rays_and_ownership().
// Only the public C ABI of vdbm_b200.h is used here.
#include "vdbm_b200.h"

#include <algorithm>
#include <cmath>
#include <condition_variable>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include <cuda_runtime.h>

namespace {

constexpr int kMaxGroup = 16;
constexpr int kBins     = 8192;
// cost weights fitted on B200 (see dist.py): raycast = visits + 67 x touched leaves; update = owned leaves
constexpr double kRaycastLeafCostInVisits = 67.0;
const char* const kPlanSource             = "\x01vdbm_plan";

double diamondAngleHost(double dx, double dy)
{
  if (dx == 0.0 && dy == 0.0) return 0.0;
  if (dy >= 0.0) return dx >= 0.0 ? dy / (dx + dy) : 1.0 - dx / (dy - dx);
  return dx < 0.0 ? 2.0 - dy / (-dx - dy) : 3.0 + dx / (dx - dy);
}
int angleBin(double a)
{
  if (!(a >= 0.0 && a < 4.0)) a = 0.0;
  return std::min(int(a * (kBins / 4.0)), kBins - 1);
}

// dist.cut_sectors: bounds that give every rank the same share of `cost`
void cutSectors(const std::vector<double>& cost, int world, double* bounds)
{
  std::vector<double> c(kBins);
  double acc = 0.0;
  for (int i = 0; i < kBins; ++i) { acc += cost[i] + 1e-9; c[i] = acc; }
  const double total = acc;
  int prev = 0;
  bounds[0] = 0.0;
  for (int r = 1; r < world; ++r)
  {
    int b = int(std::lower_bound(c.begin(), c.end(), total * r / world) - c.begin()) + 1;
    b     = std::min(std::max(b, prev + 1), kBins - (world - r));
    bounds[r] = 4.0 * b / kBins;
    prev      = b;
  }
}

} // namespace

struct vdbm_group
{
  int n = 0;
  std::vector<vdbm_map*> shards;
  std::vector<int> devices;
  double resolution = 0.0, max_range = 0.0;
  std::vector<std::pair<std::string, double> > sources; // id, max_range as given
  bool planned = false;
  vdbm_stats_t dry_run{}; // what the planning dry runs added to shard 0's ray counters (not part of the map's history)
  double ray_bounds[kMaxGroup] = {}, own_bounds[kMaxGroup] = {};
  int32_t center[2]            = {0, 0};
  std::string last_error;

  // worker threads: one per shard, all run the same task with their shard index
  std::vector<std::thread> threads;
  std::mutex mu;
  std::condition_variable cv_go, cv_done;
  std::function<int(int)> task;
  uint64_t generation = 0;
  int remaining       = 0;
  bool stop           = false;
  std::vector<int> rc;
  // shards that share a GPU (test configuration): a cudaMalloc / cudaFree in one shard's accumulate waits for the whole
  // device, i.e. for another shard's wait kernel, which in turn waits for the first shard's push -> all shards finish their
  // accumulate before any of them starts the exchange
  bool shared_device = false;
  int bar_count      = 0;
  uint64_t bar_gen   = 0;
  void barrier()
  {
    std::unique_lock<std::mutex> lk(mu);
    const uint64_t gen = bar_gen;
    if (++bar_count == n)
    {
      bar_count = 0;
      ++bar_gen;
      cv_done.notify_all();
    }
    else cv_done.wait(lk, [&] { return bar_gen != gen; });
  }

  void workerLoop(int i)
  {
    cudaSetDevice(devices[i]);
    uint64_t seen = 0;
    for (;;)
    {
      std::function<int(int)> fn;
      {
        std::unique_lock<std::mutex> lk(mu);
        cv_go.wait(lk, [&] { return stop || generation != seen; });
        if (stop) return;
        seen = generation;
        fn   = task;
      }
      const int r = fn(i);
      {
        std::lock_guard<std::mutex> lk(mu);
        rc[i] = r;
        if (--remaining == 0) cv_done.notify_all();
      }
    }
  }
  // run fn(shard index) on every worker; returns the first non-zero status (shard order)
  int runAll(std::function<int(int)> fn)
  {
    {
      std::lock_guard<std::mutex> lk(mu);
      task      = std::move(fn);
      remaining = n;
      ++generation;
    }
    cv_go.notify_all();
    std::unique_lock<std::mutex> lk(mu);
    cv_done.wait(lk, [&] { return remaining == 0; });
    for (int i = 0; i < n; ++i)
      if (rc[i])
      {
        last_error = std::string("shard ") + std::to_string(i) + ": " + vdbm_last_error(shards[i]);
        return rc[i];
      }
    return VDBM_OK;
  }
};

namespace {

double sourceRange(const vdbm_group* g, const char* id)
{
  for (auto& s : g->sources)
    if (s.first == id) return s.second == 0.0 ? g->max_range : s.second;
  return -1.0;
}

// Sector plan from one representative scan (the first one of an empty map): dry-run raycast on shard 0 for the touched
// leaves, histograms of voxel visits (by ray direction around the sensor) and of touched leaves (by leaf column around the
// sensor's leaf column), two cuts. Twin of dist.plan_rays_and_ownership().
int planFromScan(vdbm_group* g, const char* source_id, const void* points, uint64_t n, uint64_t stride, const double origin[3])
{
  const double range = sourceRange(g, source_id);
  if (range < 0.0) return VDBM_OK; // unknown source: the insert itself reports it
  vdbm_map* m0 = g->shards[0];
  std::vector<double> visits(kBins, 0.0), leaves(kBins, 0.0);
  const double res = g->resolution;
  const int32_t cx = int32_t(std::floor(origin[0] / res)) >> 3, cy = int32_t(std::floor(origin[1] / res)) >> 3;
  const auto* base = static_cast<const unsigned char*>(points);
  for (uint64_t i = 0; i < n; ++i)
  {
    float p[3];
    std::memcpy(p, base + i * stride, sizeof(p));
    const double d[3] = {double(p[0]) - origin[0], double(p[1]) - origin[1], double(p[2]) - origin[2]};
    const double len  = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
    const double sc   = (range > 0.0 && len > range) ? range / len : 1.0;
    double v          = 1.0 + (std::fabs(d[0]) + std::fabs(d[1]) + std::fabs(d[2])) * sc / res;
    if (!std::isfinite(v)) v = 0.0;
    double a = diamondAngleHost(d[0], d[1]);
    visits[angleBin(a)] += v;
  }
  int rc = vdbm_source_add(m0, kPlanSource, range > 0.0 ? range : 1.0);
  if (rc) return rc;
  vdbm_stats_t before{}, after{};
  vdbm_stats(m0, &before);
  rc = vdbm_raycast(m0, kPlanSource, points, n, stride, origin, range);
  vdbm_stats(m0, &after);
  g->dry_run.rays += after.rays - before.rays;
  g->dry_run.nan_skipped += after.nan_skipped - before.nan_skipped;
  g->dry_run.clipped += after.clipped - before.clipped;
  g->dry_run.visits += after.visits - before.visits;
  if (rc == VDBM_OK || rc == VDBM_ERR_COORD_RANGE)
  {
    vdbm_leafset* ls = nullptr;
    rc               = vdbm_update_export(m0, kPlanSource, &ls);
    if (rc == VDBM_OK)
    {
      const uint64_t nl = vdbm_leafset_size(ls);
      const int32_t* lo = vdbm_leafset_origins(ls);
      for (uint64_t i = 0; i < nl; ++i) leaves[angleBin(diamondAngleHost(double((lo[3 * i] >> 3) - cx), double((lo[3 * i + 1] >> 3) - cy)))] += 1.0;
      vdbm_leafset_free(ls);
    }
  }
  vdbm_source_add(m0, kPlanSource, 1.0); // re-adding a source empties its update grid
  if (rc != VDBM_OK && rc != VDBM_ERR_COORD_RANGE) return rc;
  std::vector<double> ray_cost(kBins), own_cost(kBins);
  for (int i = 0; i < kBins; ++i)
  {
    ray_cost[i] = visits[i] + kRaycastLeafCostInVisits * leaves[i];
    own_cost[i] = leaves[i] + 1e-6 * visits[i];
  }
  cutSectors(ray_cost, g->n, g->ray_bounds);
  cutSectors(own_cost, g->n, g->own_bounds);
  g->center[0] = cx;
  g->center[1] = cy;
  for (int i = 0; i < g->n; ++i)
  {
    rc = vdbm_shard_plan_set(g->shards[i], 1, g->n, g->center, g->own_bounds);
    if (rc) { g->last_error = vdbm_last_error(g->shards[i]); return rc; }
    rc = vdbm_ray_sector_set(g->shards[i], g->n, i, g->ray_bounds);
    if (rc) { g->last_error = vdbm_last_error(g->shards[i]); return rc; }
  }
  g->planned = true;
  return VDBM_OK;
}

} // namespace

extern "C" {

int vdbm_group_create(const vdbm_params* params, int32_t n_devices, const int32_t* devices, uint64_t inbox_capacity_records, vdbm_group** out)
{
  if (!params || !out || n_devices < 1 || n_devices > kMaxGroup || !devices || params->stream) return VDBM_ERR_INVALID_ARG;
  *out = nullptr;
  int have = 0;
  if (cudaGetDeviceCount(&have) != cudaSuccess) return VDBM_ERR_CUDA;
  for (int i = 0; i < n_devices; ++i)
  {
    // (a device may appear more than once: several shards then share one GPU, which exercises the whole machinery on a
    // single-GPU box; there is no speed to gain from it)
    if (devices[i] < 0 || devices[i] >= have) return VDBM_ERR_INVALID_ARG;
  }
  int prev_device = 0;
  cudaGetDevice(&prev_device);
  auto* g       = new vdbm_group();
  g->n          = n_devices;
  g->resolution = params->resolution;
  g->devices.assign(devices, devices + n_devices);
  g->rc.assign(n_devices, 0);
  for (int i = 0; i < n_devices; ++i)
    for (int k = 0; k < i; ++k)
      if (devices[k] == devices[i]) g->shared_device = true;
  int rc = VDBM_OK;
  for (int i = 0; i < n_devices && rc == VDBM_OK; ++i)
  {
    vdbm_params p = *params;
    p.device      = devices[i];
    vdbm_map* m   = nullptr;
    rc            = vdbm_create(&p, &m);
    if (rc == VDBM_OK) g->shards.push_back(m);
  }
  if (rc == VDBM_OK && n_devices > 1)
  {
    unsigned char handles[VDBM_IPC_HANDLE_BYTES];
    const uint64_t cap = inbox_capacity_records ? inbox_capacity_records : (uint64_t(1) << 19);
    for (int i = 0; i < n_devices && rc == VDBM_OK; ++i) rc = vdbm_exchange_create(g->shards[i], i, n_devices, cap, handles);
    for (int i = 0; i < n_devices && rc == VDBM_OK; ++i) rc = vdbm_exchange_connect_peers(g->shards[i], g->shards.data());
  }
  cudaSetDevice(prev_device);
  if (rc != VDBM_OK)
  {
    for (vdbm_map* m : g->shards) vdbm_destroy(m);
    delete g;
    return rc;
  }
  for (int i = 0; i < n_devices; ++i) g->threads.emplace_back(&vdbm_group::workerLoop, g, i);
  *out = g;
  return VDBM_OK;
}

void vdbm_group_destroy(vdbm_group* g)
{
  if (!g) return;
  {
    std::lock_guard<std::mutex> lk(g->mu);
    g->stop = true;
  }
  g->cv_go.notify_all();
  for (auto& t : g->threads) t.join();
  // no shard may free its inbox while a peer could still write to it: everything is idle once every stream is drained
  for (vdbm_map* m : g->shards) vdbm_synchronize(m);
  for (vdbm_map* m : g->shards) vdbm_destroy(m);
  delete g;
}

int32_t vdbm_group_size(const vdbm_group* g) { return g ? g->n : 0; }
vdbm_map* vdbm_group_shard(vdbm_group* g, int32_t i) { return (g && i >= 0 && i < g->n) ? g->shards[i] : nullptr; }
const char* vdbm_group_last_error(vdbm_group* g) { return g ? g->last_error.c_str() : "null group"; }

int vdbm_group_set_config(vdbm_group* g, double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max)
{
  if (!g) return VDBM_ERR_INVALID_ARG;
  int rc = VDBM_OK;
  for (int i = 0; i < g->n; ++i)
  {
    const int r = vdbm_set_config(g->shards[i], max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max);
    if (r && !rc) { rc = r; g->last_error = vdbm_last_error(g->shards[i]); }
  }
  if (max_range >= 0.0) g->max_range = max_range; // V:1458-1465: a negative range changes nothing
  return rc;
}

int vdbm_group_source_add(vdbm_group* g, const char* source_id, double max_range)
{
  if (!g || !source_id) return VDBM_ERR_INVALID_ARG;
  for (int i = 0; i < g->n; ++i)
  {
    const int rc = vdbm_source_add(g->shards[i], source_id, max_range);
    if (rc) { g->last_error = vdbm_last_error(g->shards[i]); return rc; }
  }
  for (auto& s : g->sources)
    if (s.first == source_id) { s.second = max_range; return VDBM_OK; }
  g->sources.emplace_back(source_id, max_range);
  return VDBM_OK;
}

int vdbm_group_reset(vdbm_group* g)
{
  if (!g) return VDBM_ERR_INVALID_ARG;
  const int rc = g->runAll([g](int i) {
    const int r = vdbm_reset(g->shards[i]);
    vdbm_ray_sector_set(g->shards[i], 0, 0, nullptr); // the next first scan is planned from an unfiltered dry run
    return r;
  });
  g->planned = false; // an empty map may be re-planned from its next first scan
  return rc;
}

int vdbm_group_plan_set(vdbm_group* g, const int32_t center_leaf_xy[2], const double* ray_bounds, const double* ownership_bounds)
{
  if (!g || !center_leaf_xy || !ray_bounds || !ownership_bounds) return VDBM_ERR_INVALID_ARG;
  if (g->n == 1) return VDBM_OK;
  for (int i = 0; i < g->n; ++i)
  {
    int rc = vdbm_shard_plan_set(g->shards[i], 1, g->n, center_leaf_xy, ownership_bounds);
    if (!rc) rc = vdbm_ray_sector_set(g->shards[i], g->n, i, ray_bounds);
    if (rc) { g->last_error = vdbm_last_error(g->shards[i]); return rc; }
  }
  std::copy(ray_bounds, ray_bounds + g->n, g->ray_bounds);
  std::copy(ownership_bounds, ownership_bounds + g->n, g->own_bounds);
  g->center[0] = center_leaf_xy[0];
  g->center[1] = center_leaf_xy[1];
  g->planned   = true;
  return VDBM_OK;
}

int vdbm_group_plan_get(vdbm_group* g, int32_t center_leaf_xy[2], double* ray_bounds, double* ownership_bounds)
{
  if (!g || !g->planned) return VDBM_ERR_INVALID_ARG;
  if (center_leaf_xy) { center_leaf_xy[0] = g->center[0]; center_leaf_xy[1] = g->center[1]; }
  if (ray_bounds) std::copy(g->ray_bounds, g->ray_bounds + g->n, ray_bounds);
  if (ownership_bounds) std::copy(g->own_bounds, g->own_bounds + g->n, ownership_bounds);
  return VDBM_OK;
}

int vdbm_group_insert(vdbm_group* g, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  if (!g || !source_id || !origin || (n && !points) || stride_bytes < 12) return VDBM_ERR_INVALID_ARG;
  if (g->n == 1)
  {
    const int rc = vdbm_insert(g->shards[0], source_id, points, n, stride_bytes, origin);
    if (rc) g->last_error = vdbm_last_error(g->shards[0]);
    return rc;
  }
  if (!g->planned && n != 0 && std::isfinite(origin[0]) && std::isfinite(origin[1]) && std::isfinite(origin[2]))
  {
    const int rc = planFromScan(g, source_id, points, n, stride_bytes, origin);
    if (rc) return rc;
  }
  if (!g->planned)
  {
    // nothing to plan from (empty cloud / NaN origin): integrate whatever the sources hold, shard by shard
    return g->runAll([g](int i) { return vdbm_integrate(g->shards[i], 0); });
  }
  return g->runAll([=](int i) {
    vdbm_map* m = g->shards[i];
    // every step runs on every shard whatever the previous one returned: a shard that stops early would leave its peers
    // waiting for an epoch word that never comes
    const int r0 = vdbm_accumulate(m, source_id, points, n, stride_bytes, origin);
    if (g->shared_device) g->barrier();
    if (r0 == VDBM_ERR_UNKNOWN_SOURCE) return r0; // the same on every shard: nobody pushes
    const int r1 = vdbm_update_push(m, source_id);
    if (g->shared_device)
    {
      // Shards on ONE device (the test tier): a peer that still allocates (cudaMalloc waits for the device to drain) could never
      // publish its epoch while this shard's wait kernel spins on that same device, and the wait would run into its time-out.
      // Let every push finish before anybody waits. With one device per shard the device-side wait is the point of the design.
      vdbm_synchronize(m);
      g->barrier();
    }
    const int r2 = r1 ? r1 : vdbm_update_pull_integrate(m, source_id);
    if (r2) return r2;
    return (r0 == VDBM_ERR_COORD_RANGE || r0 == VDBM_ERR_NOT_CONFIGURED) ? r0 : (r0 ? r0 : VDBM_OK);
  });
}

int vdbm_group_checksum(vdbm_group* g, uint64_t out2[2])
{
  if (!g || !out2) return VDBM_ERR_INVALID_ARG;
  out2[0] = out2[1] = 0;
  for (int i = 0; i < g->n; ++i)
  {
    uint64_t c[2] = {0, 0};
    const int rc  = vdbm_map_checksum(g->shards[i], c);
    if (rc) { g->last_error = vdbm_last_error(g->shards[i]); return rc; }
    out2[0] += c[0];
    out2[1] += c[1];
  }
  return VDBM_OK;
}

int vdbm_group_stats(vdbm_group* g, vdbm_stats_t* out)
{
  if (!g || !out) return VDBM_ERR_INVALID_ARG;
  std::memset(out, 0, sizeof(*out));
  for (int i = 0; i < g->n; ++i)
  {
    vdbm_stats_t s;
    const int rc = vdbm_stats(g->shards[i], &s);
    if (rc) return rc;
    if (i == 0) out->rays = s.rays; // every shard sees the whole cloud; each ray is cast by exactly one of them
    out->nan_skipped += s.nan_skipped;
    out->clipped += s.clipped;
    out->visits += s.visits;
    out->voxel_updates += s.voxel_updates;
    out->state_changes += s.state_changes;
    out->map_leaves += s.map_leaves;
    out->new_leaves += s.new_leaves;
    out->last_touched_leaves += s.last_touched_leaves;
    out->last_voxel_updates += s.last_voxel_updates;
    out->last_visits += s.last_visits;
    out->last_accumulate_ms = std::max(out->last_accumulate_ms, s.last_accumulate_ms);
    out->last_integrate_ms  = std::max(out->last_integrate_ms, s.last_integrate_ms);
    out->last_prep_ms       = std::max(out->last_prep_ms, s.last_prep_ms);
    out->update_capacity    = std::max(out->update_capacity, s.update_capacity);
    out->map_capacity += s.map_capacity;
    out->gpu_launches = s.gpu_launches; // process-wide counter
  }
  out->rays -= g->dry_run.rays;
  out->nan_skipped -= g->dry_run.nan_skipped;
  out->clipped -= g->dry_run.clipped;
  out->visits -= g->dry_run.visits;
  return VDBM_OK;
}

} // extern "C"
