// flat_probe.cpp — "optimistic CPU" probe: TEST / BENCH INFRASTRUCTURE ONLY, never on the product path.
//
// The oracle (vdbm_oracle.cpp) restates the reference faithfully, tree, accessor and virtual calls included, so that
// its cost model is the reference's. This file answers a different question: how fast could a CPU go on the same
// arithmetic if the OpenVDB tree were replaced by what the GPU path uses — a flat open-addressing hash of 8^3 leaves
// with a last-leaf cache — and the rays were spread over T threads (private update grids, OR-merged; leaf-parallel
// update)? bench.py reports it beside the oracle so that the GPU/CPU ratio is never inflated by a slow baseline.
// Same fp64 sequences as VDBMapping.hpp:466-566,612-631 and the node ops of OccupancyVDBMapping.hpp:92-117 (no change
// grid, no tile-probe quirk: those cost nothing measurable). Its voxel-visit and voxel-update counts must equal the
// oracle's (tests/test_oracle_kat.py::test_flat_probe_counts_match_the_oracle).
#include <algorithm>
#include <cfloat>
#include <chrono>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace {

inline uint64_t mix64(uint64_t x)
{
  x ^= x >> 33; x *= 0xff51afd7ed558ccdULL; x ^= x >> 33; x *= 0xc4ceb9fe1a85ec53ULL; x ^= x >> 33;
  return x;
}
inline uint64_t leafKey(int32_t x, int32_t y, int32_t z)
{
  const uint64_t b = 1u << 20;
  return (uint64_t(uint32_t((x >> 3) + b) & 0x1FFFFF) << 42) | (uint64_t(uint32_t((y >> 3) + b) & 0x1FFFFF) << 21) |
         uint64_t(uint32_t((z >> 3) + b) & 0x1FFFFF);
}

template <typename LeafT>
struct FlatHash
{
  std::vector<uint64_t> keys;
  std::vector<uint32_t> idx;
  std::vector<LeafT> leaves;
  std::vector<uint64_t> leaf_keys;
  uint64_t mask = 0;
  explicit FlatHash(size_t cap_log2 = 16) { rehash(size_t(1) << cap_log2); }
  void rehash(size_t cap)
  {
    keys.assign(cap, ~uint64_t(0));
    idx.assign(cap, 0);
    mask = cap - 1;
    for (uint32_t i = 0; i < leaf_keys.size(); ++i)
    {
      uint64_t h = mix64(leaf_keys[i]) & mask;
      while (keys[h] != ~uint64_t(0)) h = (h + 1) & mask;
      keys[h] = leaf_keys[i];
      idx[h]  = i;
    }
  }
  void clear()
  {
    std::fill(keys.begin(), keys.end(), ~uint64_t(0));
    leaves.clear();
    leaf_keys.clear();
  }
  LeafT* find(uint64_t key)
  {
    uint64_t h = mix64(key) & mask;
    while (keys[h] != ~uint64_t(0))
    {
      if (keys[h] == key) return &leaves[idx[h]];
      h = (h + 1) & mask;
    }
    return nullptr;
  }
  // returns index (stable across growth); created = true when the leaf is new (zero-initialised)
  uint32_t touch(uint64_t key, bool& created)
  {
    uint64_t h = mix64(key) & mask;
    while (keys[h] != ~uint64_t(0))
    {
      if (keys[h] == key) { created = false; return idx[h]; }
      h = (h + 1) & mask;
    }
    created = true;
    if ((leaf_keys.size() + 1) * 2 > keys.size())
    {
      leaf_keys.push_back(key);
      leaves.emplace_back();
      rehash(keys.size() * 2);
      return uint32_t(leaf_keys.size() - 1);
    }
    keys[h] = key;
    idx[h]  = uint32_t(leaf_keys.size());
    leaf_keys.push_back(key);
    leaves.emplace_back();
    return idx[h];
  }
};

struct UpdLeaf { uint64_t act[8] = {0, 0, 0, 0, 0, 0, 0, 0}, val[8] = {0, 0, 0, 0, 0, 0, 0, 0}; };
struct MapLeaf
{
  float v[512];
  uint64_t m[8];
  MapLeaf() { std::memset(v, 0, sizeof(v)); std::memset(m, 0, sizeof(m)); }
};

struct Probe
{
  double res, inv_res;
  float lo[6] = {0, 0, 0, 0, 0, 0}; // hit, miss, thres_min, thres_max, max, min
  FlatHash<MapLeaf> map{18};
  std::vector<FlatHash<UpdLeaf> > upd; // one private update grid per thread
  uint64_t visits = 0, updates = 0;
  explicit Probe(double r) : res(r), inv_res(1.0 / r) {}

  int32_t w2i(double c) const
  {
    if (std::fmod(c, res)) c = c + (res / 2.0);
    return int32_t(std::floor(c * inv_res));
  }

  // rays [i0, i1) into grid g; returns voxel visits
  uint64_t raycast(FlatHash<UpdLeaf>& g, const uint8_t* pts, size_t i0, size_t i1, size_t stride, const double origin[3], double range)
  {
    const int32_t o[3] = {w2i(origin[0]), w2i(origin[1]), w2i(origin[2])};
    uint64_t n_vis = 0, cur_key = ~uint64_t(0);
    uint32_t cur = 0;
    bool created;
    for (size_t i = i0; i < i1; ++i)
    {
      float p[3];
      std::memcpy(p, pts + i * stride, sizeof(p));
      double e[3] = {double(p[0]), double(p[1]), double(p[2])};
      if (std::isnan(e[0]) || std::isnan(e[1]) || std::isnan(e[2])) continue;
      bool clipped = false;
      if (range > 0.0)
      {
        const double d[3] = {e[0] - origin[0], e[1] - origin[1], e[2] - origin[2]};
        const double len  = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (len > range)
        {
          for (int a = 0; a < 3; ++a) e[a] = origin[a] + (d[a] / len) * range;
          clipped = true;
        }
      }
      const int32_t en[3] = {w2i(e[0]), w2i(e[1]), w2i(e[2])};
      int32_t v[3]        = {o[0], o[1], o[2]};
      if (!(en[0] == o[0] && en[1] == o[1] && en[2] == o[2]))
      {
        double next[3], delta[3];
        int32_t step[3];
        for (int a = 0; a < 3; ++a)
        {
          const double dir = double(en[a]) - double(o[a]);
          if (dir == 0.0) { step[a] = 0; next[a] = DBL_MAX; delta[a] = DBL_MAX; }
          else
          {
            const double inv = 1.0 / dir;
            step[a]  = inv > 0 ? 1 : -1;
            delta[a] = double(step[a]) * inv;
            next[a]  = 0.0 + (double(v[a] + (inv > 0 ? 1 : 0)) - (double(o[a]) + 0.5)) * inv;
          }
        }
        bool more;
        do
        {
          const uint64_t key = leafKey(v[0], v[1], v[2]);
          if (key != cur_key) { cur_key = key; cur = g.touch(key, created); }
          g.leaves[cur].act[v[0] & 7] |= uint64_t(1) << (((v[1] & 7) << 3) | (v[2] & 7));
          ++n_vis;
          const int axis = (next[0] < next[1] && next[0] < next[2]) ? 0 : ((next[1] < next[2]) ? 1 : 2);
          const double t = next[axis];
          next[axis] += delta[axis];
          v[axis] += step[axis];
          more = (t <= 1.0);
        } while (more);
      }
      if (!clipped)
      {
        const uint64_t key = leafKey(en[0], en[1], en[2]);
        if (key != cur_key) { cur_key = key; cur = g.touch(key, created); }
        const uint64_t bit = uint64_t(1) << (((en[1] & 7) << 3) | (en[2] & 7));
        g.leaves[cur].act[en[0] & 7] |= bit;
        g.leaves[cur].val[en[0] & 7] |= bit;
      }
    }
    return n_vis;
  }

  uint64_t applyLeaf(const UpdLeaf& u, MapLeaf& ml) const
  {
    uint64_t n = 0;
    for (int w = 0; w < 8; ++w)
    {
      uint64_t a = u.act[w];
      while (a)
      {
        const int b = __builtin_ctzll(a);
        a &= a - 1;
        float& val       = ml.v[w * 64 + b];
        const uint64_t m = uint64_t(1) << b;
        if (u.val[w] & m)
        {
          val += lo[0];
          if (val > lo[3]) { ml.m[w] |= m; if (val > lo[4]) val = lo[4]; }
        }
        else
        {
          val += lo[1];
          if (val < lo[2]) { ml.m[w] &= ~m; if (val < lo[5]) val = lo[5]; }
        }
        ++n;
      }
    }
    return n;
  }

  void insert(const uint8_t* pts, size_t n, size_t stride, const double origin[3], double range, int threads)
  {
    threads = std::max(1, threads);
    if (int(upd.size()) < threads) upd.resize(threads, FlatHash<UpdLeaf>(16));
    std::vector<uint64_t> vis(threads, 0);
    if (threads == 1) vis[0] = raycast(upd[0], pts, 0, n, stride, origin, range);
    else
    {
      std::vector<std::thread> th;
      for (int t = 0; t < threads; ++t)
        th.emplace_back([&, t] { vis[t] = raycast(upd[t], pts, n * t / threads, n * (t + 1) / threads, stride, origin, range); });
      for (auto& x : th) x.join();
      // OR-merge the private grids into grid 0 (serial: it is a small fraction of the work)
      bool created;
      for (int t = 1; t < threads; ++t)
        for (size_t i = 0; i < upd[t].leaf_keys.size(); ++i)
        {
          UpdLeaf& d = upd[0].leaves[upd[0].touch(upd[t].leaf_keys[i], created)];
          for (int w = 0; w < 8; ++w) { d.act[w] |= upd[t].leaves[i].act[w]; d.val[w] |= upd[t].leaves[i].val[w]; }
        }
    }
    for (uint64_t v : vis) visits += v;
    // updateMap: create the map leaves serially (hash inserts), then apply leaf-parallel
    FlatHash<UpdLeaf>& g = upd[0];
    std::vector<uint32_t> target(g.leaf_keys.size());
    bool created;
    for (size_t i = 0; i < g.leaf_keys.size(); ++i) target[i] = map.touch(g.leaf_keys[i], created);
    std::vector<uint64_t> cnt(threads, 0);
    auto work = [&](int t) {
      uint64_t c = 0;
      for (size_t i = g.leaf_keys.size() * t / threads; i < g.leaf_keys.size() * (t + 1) / threads; ++i) c += applyLeaf(g.leaves[i], map.leaves[target[i]]);
      cnt[t] = c;
    };
    if (threads == 1) work(0);
    else
    {
      std::vector<std::thread> th;
      for (int t = 0; t < threads; ++t) th.emplace_back(work, t);
      for (auto& x : th) x.join();
    }
    for (uint64_t c : cnt) updates += c;
    for (int t = 0; t < threads; ++t) upd[t].clear();
  }
};

} // namespace

extern "C" {
void* flat_create(double resolution, const float* logodds6)
{
  auto* p = new Probe(resolution);
  std::memcpy(p->lo, logodds6, sizeof(p->lo));
  return p;
}
void flat_destroy(void* h) { delete static_cast<Probe*>(h); }
void flat_insert(void* h, const void* pts, uint64_t n, uint64_t stride, const double* origin, double range, int threads)
{
  static_cast<Probe*>(h)->insert(static_cast<const uint8_t*>(pts), n, stride, origin, range, threads);
}
// out[4] = visits, voxel updates, map leaves, active map voxels
void flat_stats(void* h, uint64_t* out)
{
  auto* p = static_cast<Probe*>(h);
  out[0]  = p->visits;
  out[1]  = p->updates;
  out[2]  = p->map.leaf_keys.size();
  uint64_t on = 0;
  for (const MapLeaf& l : p->map.leaves)
    for (int w = 0; w < 8; ++w) on += uint64_t(__builtin_popcountll(l.m[w]));
  out[3] = on;
}
}
