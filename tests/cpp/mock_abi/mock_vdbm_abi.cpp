// mock_vdbm_abi.cpp — TEST INFRASTRUCTURE, never built into or loaded by the product.
//
// The subset of the C ABI (include/vdbm_b200.h) that the drop-in C++ classes call, implemented on the CPU oracle
// (oracle/vdbm_oracle.cpp, linked into the same test executable). It exists so that the HOST LOGIC of the shim - locks and
// threads, the per-source raycast handles and their bookkeeping, the sharded mode, the mirror tables, both backends of
// detail/backend.hpp - runs in the CPU test tier (`pytest -m "not gpu"`), where no GPU exists; the same test programs run
// against the real libvdbm_b200.so in the GPU tier. Nothing here says anything about the CUDA path: parity of the product is
// only ever claimed from the `-m gpu` tests, which call the real library.
//
// Fidelity: semantics of every call follow the oracle's restatement of the reference; handles of a "group" are modelled
// as one map plus empty shards (the union of the shards is the map, which is all the shim relies on); "device" records are
// host memory; entry points the shim never calls abort with a message.
#include "../../../include/vdbm_b200.h"

#include <array>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <map>
#include <set>
#include <string>
#include <vector>

extern "C" {
// the oracle's C API (oracle/vdbm_oracle.cpp)
void* vdbo_create(double resolution);
void vdbo_destroy(void* h);
int vdbo_set_config(void* h, double max_range, double prob_hit, double prob_miss, double thres_min, double thres_max);
void vdbo_get_logodds(void* h, float* out);
void vdbo_add_source(void* h, const char* id, double max_range);
void vdbo_reset(void* h);
int vdbo_accumulate(void* h, const char* id, const void* pts, uint64_t n, uint64_t stride, const double* origin);
void vdbo_integrate(void* h);
void vdbo_stats(void* h, uint64_t* out);
int64_t vdbo_export_prepare(void* hh, int kind, const char* source, const int32_t* bbmin, const int32_t* bbmax, int full);
void vdbo_export_fetch(void* hh, int32_t* origins, uint64_t* active, uint64_t* valmask, float* values);
uint64_t vdbo_map_leaf_count(void* hh);
int vdbo_update_import(void* hh, const char* source, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* valmask);
int vdbo_apply_section_update(void* hh, const int32_t* bbmin, const int32_t* bbmax, uint64_t n, const int32_t* origins, const uint64_t* active);
int vdbo_apply_section_grid(void* hh, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int tile_quirk);
int vdbo_update_apply(void* hh, const char* source, int level, uint64_t n, const int32_t* origins, const uint64_t* active,
                      const uint64_t* valmask, const double* origin);
void vdbo_last_origin(void* hh, const char* source, double* out);
void vdbo_points_set(void* hh, const void* pts, uint64_t n, uint64_t stride, int occupied);
void vdbo_add_artificial_areas(void* hh, uint64_t n_poly, const uint32_t* counts, const double* xyz, double negative_height, double positive_height);
void vdbo_set_fast_mode(void* hh, int on);
void vdbo_raytrace(void* hh, uint64_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* successes,
                   double* end_points);
void vdbo_add_artificial_wall(void* hh, const double* start, const double* end, double negative_height, double positive_height);
int vdbo_cast_index_rays(void* hh, const char* source, uint64_t n, const int32_t* rays6);
void vdbo_restore_map_integrity(void* hh);
int vdbo_update_clear(void* hh, const char* source);
}

struct vdbm_leafset
{
  std::vector<int32_t> origins;
  std::vector<uint64_t> active, valmask;
  std::vector<float> values;
  uint64_t n = 0;
};

namespace {
constexpr const char* kApplyScratch = "\x02mock_apply_scratch";
struct Record // what crosses between handles in vdbm_update_partition / vdbm_update_import_device (opaque to the shim)
{
  int32_t origin[3];
  uint64_t active[8], value[8];
};
struct Delivered
{
  uint32_t index;
  std::array<uint64_t, 8> active;
  std::vector<float> values;
};
} // namespace

struct vdbm_map
{
  void* o = nullptr;
  std::string err;
  uint64_t generation = 0;
  std::map<std::array<int32_t, 3>, Delivered> delivered; // leaves as last handed to the host (dirty = differs from this)
  uint32_t next_index = 0;
  std::vector<Record> part;
  std::set<std::string> sources;
  double last_range = 0.0;
  bool has_apply_scratch = false;
};

struct vdbm_group
{
  std::vector<vdbm_map*> shards; // shards[0] holds the map, the others stay empty: the union of the shards is the map
  std::string err;
};

namespace {
int fail(vdbm_map* m, int code, const std::string& msg)
{
  m->err = msg;
  return code;
}
[[noreturn]] void notMocked(const char* name)
{
  std::fprintf(stderr, "mock_vdbm_abi: %s is not part of the mock (the shim does not call it)\n", name);
  std::abort();
}
vdbm_leafset* exportKind(vdbm_map* m, int kind, const char* source, const int32_t* mn, const int32_t* mx, int full, bool with_values)
{
  const int32_t zero[3] = {0, 0, 0};
  const int64_t n = vdbo_export_prepare(m->o, kind, source ? source : "", mn ? mn : zero, mx ? mx : zero, full);
  auto* ls = new vdbm_leafset();
  if (n <= 0) return ls;
  ls->n = uint64_t(n);
  ls->origins.resize(size_t(n) * 3);
  ls->active.resize(size_t(n) * 8);
  if (with_values) ls->values.resize(size_t(n) * 512);
  else ls->valmask.resize(size_t(n) * 8);
  vdbo_export_fetch(m->o, ls->origins.data(), ls->active.data(), with_values ? nullptr : ls->valmask.data(), with_values ? ls->values.data() : nullptr);
  return ls;
}
// the leaves of the map that differ from what the host was last given (or all of them), in pool-index order
vdbm_leafset* collectMap(vdbm_map* m, bool dirty_only, std::vector<uint32_t>* index_out)
{
  vdbm_leafset* all = exportKind(m, 0, nullptr, nullptr, nullptr, 0, true);
  struct Pick { uint32_t index; uint64_t i; };
  std::vector<Pick> picks;
  for (uint64_t i = 0; i < all->n; ++i)
  {
    const std::array<int32_t, 3> key = {all->origins[3 * i], all->origins[3 * i + 1], all->origins[3 * i + 2]};
    auto it = m->delivered.find(key);
    bool changed = it == m->delivered.end();
    if (changed)
    {
      Delivered d;
      d.index = m->next_index++;
      it      = m->delivered.emplace(key, std::move(d)).first;
    }
    else
      changed = std::memcmp(it->second.active.data(), &all->active[8 * i], 64) != 0 ||
                std::memcmp(it->second.values.data(), &all->values[512 * i], 2048) != 0;
    if (changed)
    {
      std::memcpy(it->second.active.data(), &all->active[8 * i], 64);
      it->second.values.assign(all->values.begin() + 512 * i, all->values.begin() + 512 * (i + 1));
    }
    if (changed || !dirty_only) picks.push_back({it->second.index, i});
  }
  if (index_out)
  {
    // the mirror delivers in pool-index order; the leaf-set exports keep the canonical origin order
    for (size_t a = 1; a < picks.size(); ++a)
      for (size_t b = a; b > 0 && picks[b].index < picks[b - 1].index; --b) std::swap(picks[b], picks[b - 1]);
  }
  auto* out = new vdbm_leafset();
  out->n    = picks.size();
  for (const Pick& p : picks)
  {
    out->origins.insert(out->origins.end(), all->origins.begin() + 3 * p.i, all->origins.begin() + 3 * (p.i + 1));
    out->active.insert(out->active.end(), all->active.begin() + 8 * p.i, all->active.begin() + 8 * (p.i + 1));
    out->values.insert(out->values.end(), all->values.begin() + 512 * p.i, all->values.begin() + 512 * (p.i + 1));
    if (index_out) index_out->push_back(p.index);
  }
  delete all;
  return out;
}
void restartPool(vdbm_map* m)
{
  m->delivered.clear();
  m->next_index = 0;
  ++m->generation;
}
int mapAccumulateRc(int rc) { return rc == 1 ? VDBM_ERR_UNKNOWN_SOURCE : (rc == 2 ? VDBM_ERR_NOT_CONFIGURED : VDBM_OK); }
} // namespace

extern "C" {

int vdbm_abi_version(void) { return VDBM_ABI_VERSION; }

int vdbm_create(const vdbm_params* params, vdbm_map** out)
{
  if (!params || !out || !(params->resolution > 0.0)) return VDBM_ERR_INVALID_ARG;
  auto* m = new vdbm_map();
  m->o    = vdbo_create(params->resolution);
  *out    = m;
  return VDBM_OK;
}
void vdbm_destroy(vdbm_map* m)
{
  if (!m) return;
  vdbo_destroy(m->o);
  delete m;
}
int vdbm_reset(vdbm_map* m)
{
  vdbo_reset(m->o);
  restartPool(m);
  return VDBM_OK;
}
int vdbm_set_config(vdbm_map* m, double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max)
{
  return vdbo_set_config(m->o, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max) ? fail(m, VDBM_ERR_BAD_CONFIG, "bad config") : VDBM_OK;
}
int vdbm_get_logodds(vdbm_map* m, float* out6)
{
  vdbo_get_logodds(m->o, out6);
  return VDBM_OK;
}
int vdbm_source_add(vdbm_map* m, const char* source_id, double max_range)
{
  vdbo_add_source(m->o, source_id, max_range); // re-adding gives the source a fresh update grid, like the reference
  m->sources.insert(source_id);
  return VDBM_OK;
}
int vdbm_accumulate(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  const int rc = mapAccumulateRc(vdbo_accumulate(m->o, source_id, points, n, stride_bytes, origin));
  return rc ? fail(m, rc, rc == VDBM_ERR_UNKNOWN_SOURCE ? "Source not available" : "not configured") : VDBM_OK;
}
int vdbm_raycast(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3], double raycast_range)
{
  vdbo_add_source(m->o, source_id, raycast_range); // the shim only ever raycasts with an explicit range into its (empty) scratch source
  return vdbm_accumulate(m, source_id, points, n, stride_bytes, origin);
}
int vdbm_integrate(vdbm_map* m, int /*keep_change*/)
{
  vdbo_integrate(m->o);
  return VDBM_OK;
}
int vdbm_integrate_from(vdbm_map* m, vdbm_map* holder, const char* source_id, int /*keep_change*/)
{
  if (!m || !holder) return VDBM_ERR_INVALID_ARG;
  if (!holder->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available");
  if (!m->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available"); // (the mock applies through the map's twin of the source)
  vdbm_leafset* upd = exportKind(holder, 1, source_id, nullptr, nullptr, 0, false);
  const double zero[3] = {0, 0, 0};
  vdbo_update_apply(m->o, source_id, 0, upd->n, upd->origins.data(), upd->active.data(), upd->valmask.data(), zero); // updateMap R:731 with that grid
  vdbo_update_clear(holder->o, source_id);
  delete upd;
  return VDBM_OK;
}
int vdbm_insert(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  const int rc = vdbm_accumulate(m, source_id, points, n, stride_bytes, origin);
  vdbo_integrate(m->o);
  return rc;
}
int vdbm_insert_async(vdbm_map* m, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3], int)
{
  return vdbm_insert(m, source_id, points, n, stride_bytes, origin);
}
int vdbm_flush(vdbm_map*) { return VDBM_OK; }
int vdbm_synchronize(vdbm_map*) { return VDBM_OK; }

int vdbm_update_export(vdbm_map* m, const char* source_id, vdbm_leafset** out)
{
  if (!m->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available");
  *out = exportKind(m, 1, source_id, nullptr, nullptr, 0, false);
  return VDBM_OK;
}
int vdbm_change_export(vdbm_map* m, const char* source_id, vdbm_leafset** out)
{
  if (!m->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available");
  *out = exportKind(m, 2, source_id, nullptr, nullptr, 0, false);
  return VDBM_OK;
}
int vdbm_update_import(vdbm_map* m, const char* source_id, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* value)
{
  return vdbo_update_import(m->o, source_id, n, origins, active, value) ? fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available") : VDBM_OK;
}
int vdbm_update_map(vdbm_map* m, const char* source_id, vdbm_leafset** change)
{
  if (!m->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available");
  vdbm_leafset* upd = exportKind(m, 1, source_id, nullptr, nullptr, 0, false);
  const double zero[3] = {0, 0, 0};
  vdbo_update_apply(m->o, source_id, 0, upd->n, upd->origins.data(), upd->active.data(), upd->valmask.data(), zero);
  vdbo_update_clear(m->o, source_id);
  delete upd;
  if (change) *change = exportKind(m, 2, source_id, nullptr, nullptr, 0, false);
  return VDBM_OK;
}
int vdbm_update_partition(vdbm_map* m, const char* source_id, int32_t n_ranks, uint64_t* counts, const void** d_records)
{
  if (n_ranks != 1 || !m->sources.count(source_id)) return fail(m, VDBM_ERR_INVALID_ARG, "mock: one rank, known source");
  vdbm_leafset* upd = exportKind(m, 1, source_id, nullptr, nullptr, 0, false);
  m->part.resize(upd->n);
  for (uint64_t i = 0; i < upd->n; ++i)
  {
    std::memcpy(m->part[i].origin, &upd->origins[3 * i], 12);
    std::memcpy(m->part[i].active, &upd->active[8 * i], 64);
    std::memcpy(m->part[i].value, &upd->valmask[8 * i], 64);
  }
  counts[0]  = upd->n;
  *d_records = m->part.data();
  delete upd;
  vdbo_update_clear(m->o, source_id);
  return VDBM_OK;
}
int vdbm_update_import_device(vdbm_map* m, const char* source_id, const void* d_records, uint64_t n)
{
  const Record* r = static_cast<const Record*>(d_records);
  std::vector<int32_t> o;
  std::vector<uint64_t> a, v;
  for (uint64_t i = 0; i < n; ++i)
  {
    o.insert(o.end(), r[i].origin, r[i].origin + 3);
    a.insert(a.end(), r[i].active, r[i].active + 8);
    v.insert(v.end(), r[i].value, r[i].value + 8);
  }
  return vdbm_update_import(m, source_id, n, o.data(), a.data(), v.data());
}
int vdbm_update_create(vdbm_map* m, const char* source_id, int level, vdbm_leafset** out, double origin_out[3])
{
  if (!m->sources.count(source_id)) return fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available");
  if (level < 0 || level > 2) return fail(m, VDBM_ERR_INVALID_ARG, "update level must be 0, 1 or 2");
  if (origin_out) vdbo_last_origin(m->o, source_id, origin_out);
  *out = level == 1 ? exportKind(m, 2, source_id, nullptr, nullptr, 0, false) : exportKind(m, 1, source_id, nullptr, nullptr, level == 2 ? 2 : 0, false);
  if (level == 2 && (*out)->n == 0) // the real library refuses when the source's last accumulate did not run on this handle
  {
    delete *out;
    *out = nullptr;
    return fail(m, VDBM_ERR_INVALID_ARG, "no reduced update available: level 2 describes the source's LAST accumulate call on this handle");
  }
  return VDBM_OK;
}
int vdbm_update_apply(vdbm_map* m, int level, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* value, const double origin[3],
                      vdbm_leafset** change)
{
  if (!m->has_apply_scratch)
  {
    vdbo_add_source(m->o, kApplyScratch, 1.0);
    m->has_apply_scratch = true;
  }
  const double zero[3] = {0, 0, 0};
  if (vdbo_update_apply(m->o, kApplyScratch, level, n, origins, active, value, origin ? origin : zero)) return fail(m, VDBM_ERR_INVALID_ARG, "bad level");
  if (change) *change = (level == 1) ? nullptr : exportKind(m, 2, kApplyScratch, nullptr, nullptr, 0, false);
  return VDBM_OK;
}

int vdbm_map_export(vdbm_map* m, int dirty_only, vdbm_leafset** out)
{
  *out = collectMap(m, dirty_only != 0, nullptr);
  return VDBM_OK;
}
uint64_t vdbm_map_generation(const vdbm_map* m) { return m ? m->generation : 0; }
int vdbm_map_mirror(vdbm_map* m, uint64_t chunk_leaves, vdbm_mirror_sink sink, void* user, uint64_t* n_leaves)
{
  if (!m || !sink) return VDBM_ERR_INVALID_ARG;
  if (n_leaves) *n_leaves = 0;
  std::vector<uint32_t> index;
  vdbm_leafset* ls = collectMap(m, true, &index);
  const uint64_t C = chunk_leaves ? chunk_leaves : 16384;
  int rc           = VDBM_OK;
  for (uint64_t off = 0; off < ls->n && rc == VDBM_OK; off += C)
  {
    const uint64_t cnt = std::min<uint64_t>(C, ls->n - off);
    if (sink(user, cnt, index.data() + off, ls->origins.data() + 3 * off, ls->values.data() + 512 * off, ls->active.data() + 8 * off) != 0)
    {
      // what the consumer has not taken stays dirty: forget that it was delivered
      for (uint64_t i = off; i < ls->n; ++i)
      {
        auto it = m->delivered.find({ls->origins[3 * i], ls->origins[3 * i + 1], ls->origins[3 * i + 2]});
        if (it != m->delivered.end()) it->second.values.assign(512, -12345.0f);
      }
      rc = fail(m, VDBM_ERR_INVALID_ARG, "vdbm_map_mirror: the sink aborted the transfer");
    }
    else if (n_leaves) *n_leaves += cnt;
  }
  delete ls;
  return rc;
}
int vdbm_section(vdbm_map* m, const int32_t bbmin[3], const int32_t bbmax[3], int full, int result_float, vdbm_leafset** out)
{
  *out = exportKind(m, result_float ? 4 : 3, nullptr, bbmin, bbmax, full, result_float != 0);
  return VDBM_OK;
}
int vdbm_section_apply_update(vdbm_map* m, const int32_t bbmin[3], const int32_t bbmax[3], uint64_t n, const int32_t* origins, const uint64_t* active)
{
  vdbo_apply_section_update(m->o, bbmin, bbmax, n, origins, active);
  return VDBM_OK;
}
int vdbm_section_apply_grid(vdbm_map* m, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int replicate_tile_quirk)
{
  vdbo_apply_section_grid(m->o, n, origins, active, values, replicate_tile_quirk);
  return VDBM_OK;
}
int vdbm_map_import(vdbm_map* m, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int replace)
{
  if (replace)
  {
    // only the MAP is replaced (loadMap R:263-284); the oracle's reset also renews the sources' update grids, which the
    // shim's callers leave empty at this point
    vdbo_reset(m->o);
    restartPool(m);
  }
  vdbo_apply_section_grid(m->o, n, origins, active, values, 0);
  return VDBM_OK;
}
int vdbm_points_set(vdbm_map* m, const void* points, uint64_t n, uint64_t stride_bytes, int occupied)
{
  vdbo_points_set(m->o, points, n, stride_bytes, occupied);
  return VDBM_OK;
}
int vdbm_artificial_areas_add(vdbm_map* m, uint64_t n_polygons, const uint32_t* counts, const double* xyz, double negative_height, double positive_height)
{
  vdbo_add_artificial_areas(m->o, n_polygons, counts, xyz, negative_height, positive_height);
  return VDBM_OK;
}
int vdbm_artificial_walls_add(vdbm_map* m, uint64_t n_polylines, const uint32_t* counts, const double* xyz, double negative_height, double positive_height,
                              int closed)
{
  size_t base = 0;
  for (uint64_t p = 0; p < n_polylines; ++p)
  {
    const uint32_t c = counts[p];
    for (uint32_t k = 0; k + 1 < c; ++k) vdbo_add_artificial_wall(m->o, xyz + 3 * (base + k), xyz + 3 * (base + k + 1), negative_height, positive_height);
    if (closed && c > 2) vdbo_add_artificial_wall(m->o, xyz + 3 * (base + c - 1), xyz + 3 * base, negative_height, positive_height);
    base += c;
  }
  return VDBM_OK;
}
int vdbm_map_integrity_restore(vdbm_map* m)
{
  vdbo_restore_map_integrity(m->o);
  return VDBM_OK;
}
int vdbm_cast_index_rays(vdbm_map* m, const char* source_id, uint64_t n_rays, const int32_t* rays6)
{
  return vdbo_cast_index_rays(m->o, source_id, n_rays, rays6) ? fail(m, VDBM_ERR_UNKNOWN_SOURCE, "Source not available") : VDBM_OK;
}
int vdbm_set_fast_mode(vdbm_map* m, int on)
{
  vdbo_set_fast_mode(m->o, on);
  return VDBM_OK;
}
int vdbm_raytrace(vdbm_map* m, uint64_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* success, double* end_points)
{
  vdbo_raytrace(m->o, n, origins, directions, max_lengths, success, end_points);
  return VDBM_OK;
}

uint64_t vdbm_leafset_size(const vdbm_leafset* s) { return s ? s->n : 0; }
const int32_t* vdbm_leafset_origins(const vdbm_leafset* s) { return s->origins.data(); }
const uint64_t* vdbm_leafset_active(const vdbm_leafset* s) { return s->active.data(); }
const uint64_t* vdbm_leafset_valmask(const vdbm_leafset* s) { return s->valmask.data(); }
const float* vdbm_leafset_values(const vdbm_leafset* s) { return s->values.data(); }
void vdbm_leafset_free(vdbm_leafset* s) { delete s; }

int vdbm_stats(vdbm_map* m, vdbm_stats_t* out)
{
  std::memset(out, 0, sizeof(*out));
  uint64_t s[6];
  vdbo_stats(m->o, s);
  out->rays = s[0]; out->nan_skipped = s[1]; out->clipped = s[2]; out->visits = s[3]; out->voxel_updates = s[4]; out->state_changes = s[5];
  out->map_leaves = vdbo_map_leaf_count(m->o);
  return VDBM_OK;
}
const char* vdbm_last_error(vdbm_map* m) { return m ? m->err.c_str() : "null handle"; }

// ---- group: one real map + empty shards ---------------------------------------------------------------------------------
int vdbm_group_create(const vdbm_params* params, int32_t n_devices, const int32_t* devices, uint64_t, vdbm_group** out)
{
  if (!params || n_devices <= 0 || !devices || !out) return VDBM_ERR_INVALID_ARG;
  auto* g = new vdbm_group();
  for (int32_t i = 0; i < n_devices; ++i)
  {
    vdbm_map* m = nullptr;
    vdbm_create(params, &m);
    g->shards.push_back(m);
  }
  *out = g;
  return VDBM_OK;
}
void vdbm_group_destroy(vdbm_group* g)
{
  if (!g) return;
  for (vdbm_map* m : g->shards) vdbm_destroy(m);
  delete g;
}
int32_t vdbm_group_size(const vdbm_group* g) { return g ? int32_t(g->shards.size()) : 0; }
vdbm_map* vdbm_group_shard(vdbm_group* g, int32_t i) { return g->shards[size_t(i)]; }
int vdbm_group_set_config(vdbm_group* g, double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max)
{
  int rc = VDBM_OK;
  for (vdbm_map* m : g->shards) rc = vdbm_set_config(m, max_range, prob_hit, prob_miss, prob_thres_min, prob_thres_max);
  return rc;
}
int vdbm_group_source_add(vdbm_group* g, const char* source_id, double max_range)
{
  for (vdbm_map* m : g->shards) vdbm_source_add(m, source_id, max_range);
  return VDBM_OK;
}
int vdbm_group_reset(vdbm_group* g)
{
  for (vdbm_map* m : g->shards) vdbm_reset(m);
  return VDBM_OK;
}
int vdbm_group_insert(vdbm_group* g, const char* source_id, const void* points, uint64_t n, uint64_t stride_bytes, const double origin[3])
{
  const int rc = vdbm_insert(g->shards[0], source_id, points, n, stride_bytes, origin);
  if (rc) g->err = g->shards[0]->err;
  return rc;
}
int vdbm_group_stats(vdbm_group* g, vdbm_stats_t* out) { return vdbm_stats(g->shards[0], out); }
const char* vdbm_group_last_error(vdbm_group* g) { return g ? g->err.c_str() : "null group"; }

// ---- not needed by the shim ---------------------------------------------------------------------------------------------
int vdbm_accumulate_device(vdbm_map*, const char*, const void*, uint64_t, uint64_t, const double*) { notMocked("vdbm_accumulate_device"); }
int vdbm_prefetch(vdbm_map*, const void*, uint64_t, uint64_t) { notMocked("vdbm_prefetch"); }
int vdbm_artificial_export(vdbm_map*, vdbm_leafset**) { notMocked("vdbm_artificial_export"); }
int vdbm_probe(vdbm_map*, const int32_t*, float*, int32_t*) { notMocked("vdbm_probe"); }
int vdbm_pipeline_counts(vdbm_map*, uint64_t*) { notMocked("vdbm_pipeline_counts"); }
void* vdbm_host_alloc(size_t bytes) { return std::malloc(bytes); }
void vdbm_host_free(void* p) { std::free(p); }

} // extern "C"
