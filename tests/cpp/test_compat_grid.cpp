// CPU-only check of the stand-in host grid's bulk leaf update (the eager mirror of the shim uses it): sorted, unsorted and
// repeated origins must end in exactly the grid that per-leaf updates produce.
#include <algorithm>
#include <array>
#include <cstdio>
#include <random>
#include <vector>

#define VDBM_FORCE_COMPAT 1
#include <vdb_mapping/detail/backend.hpp>

#include "mini_gtest.h"

using Backend = vdb_mapping::detail::Backend<float>;

static int fill(std::vector<std::int32_t>& origins, std::vector<float>& values, std::vector<std::uint64_t>& active, std::mt19937& rng, int n,
                bool sorted)
{
  std::vector<std::array<std::int32_t, 3> > o(n);
  for (auto& c : o) c = {std::int32_t(rng() % 64 - 32) * 8, std::int32_t(rng() % 64 - 32) * 8, std::int32_t(rng() % 16 - 8) * 8};
  std::sort(o.begin(), o.end());
  o.erase(std::unique(o.begin(), o.end()), o.end()); // an export never lists a leaf twice (the payload copies run threaded)
  if (!sorted) std::shuffle(o.begin(), o.end(), rng);
  for (auto& c : o)
  {
    origins.insert(origins.end(), c.begin(), c.end());
    for (int k = 0; k < 512; ++k) values.push_back(float(rng() % 1000) * 0.01f);
    for (int w = 0; w < 8; ++w) active.push_back((std::uint64_t(rng()) << 32) | rng());
  }
  return int(o.size());
}

TEST(CompatGrid, BulkLeafUpdateEqualsPerLeafUpdates)
{
  std::mt19937 rng(42);
  for (int round = 0; round < 6; ++round)
  {
    Backend::GridT bulk, single;
    for (int pass = 0; pass < 3; ++pass) // later passes hit existing leaves and add new ones in between
    {
      std::vector<std::int32_t> origins;
      std::vector<float> values;
      std::vector<std::uint64_t> active;
      const int want = (round % 2 == 0) ? 6000 : 300; // above and below the threading threshold
      const int n    = fill(origins, values, active, rng, want, /*sorted=*/(round + pass) % 3 != 0);
      Backend::putMapLeaves(bulk, std::uint64_t(n), origins.data(), values.data(), active.data());
      for (int i = 0; i < n; ++i) Backend::putMapLeaf(single, origins.data() + 3 * i, values.data() + 512 * i, active.data() + 8 * i);
    }
    EXPECT_EQ(bulk.leafCount(), single.leafCount());
    EXPECT_EQ(bulk.activeVoxelCount(), single.activeVoxelCount());
    bool same = true;
    auto a = bulk.leaves().begin();
    auto b = single.leaves().begin();
    for (; a != bulk.leaves().end() && b != single.leaves().end(); ++a, ++b)
    {
      if (!(a->first == b->first)) same = false;
      for (int k = 0; k < 512 && same; ++k) same = a->second.values[k] == b->second.values[k];
      for (int w = 0; w < 8 && same; ++w) same = a->second.active[w] == b->second.active[w];
    }
    EXPECT_TRUE(same);
  }
}

// The chunked mirror path: leaves addressed through the "device pool index -> host leaf" table, payloads copied on the
// persistent worker pool. Indices are dense and append-only like the device pool's; chunks repeat old indices and add new ones.
TEST(CompatGrid, IndexedChunksOnTheWorkerPoolEqualPerLeafUpdates)
{
  std::mt19937 rng(7);
  Backend::GridT indexed, single;
  std::vector<Backend::MapLeafT*> table;
  std::vector<std::array<std::int32_t, 3> > pool; // pool index -> origin
  for (int chunk = 0; chunk < 12; ++chunk)
  {
    std::vector<std::int32_t> origins;
    std::vector<float> values;
    std::vector<std::uint64_t> active;
    const int n = fill(origins, values, active, rng, chunk % 3 == 0 ? 5000 : 200, true);
    std::vector<std::pair<std::uint32_t, int> > order; // (pool index, position in this chunk)
    for (int i = 0; i < n; ++i)
    {
      const std::array<std::int32_t, 3> o = {origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]};
      auto it = std::find(pool.begin(), pool.end(), o);
      if (it == pool.end()) { pool.push_back(o); it = pool.end() - 1; }
      order.emplace_back(std::uint32_t(it - pool.begin()), i);
    }
    std::sort(order.begin(), order.end()); // the ABI delivers a chunk in ascending pool index
    std::vector<std::uint32_t> index;
    std::vector<std::int32_t> o2;
    std::vector<float> v2;
    std::vector<std::uint64_t> a2;
    for (auto& [idx, i] : order)
    {
      index.push_back(idx);
      o2.insert(o2.end(), origins.begin() + 3 * i, origins.begin() + 3 * i + 3);
      v2.insert(v2.end(), values.begin() + 512 * i, values.begin() + 512 * (i + 1));
      a2.insert(a2.end(), active.begin() + 8 * i, active.begin() + 8 * (i + 1));
    }
    Backend::putMapLeavesIndexed(indexed, table, std::uint64_t(n), index.data(), o2.data(), v2.data(), a2.data());
    for (int i = 0; i < n; ++i) Backend::putMapLeaf(single, origins.data() + 3 * i, values.data() + 512 * i, active.data() + 8 * i);
  }
  EXPECT_EQ(indexed.leafCount(), single.leafCount());
  EXPECT_EQ(indexed.leafCount(), pool.size());
  bool same = true;
  auto a = indexed.leaves().begin();
  auto b = single.leaves().begin();
  for (; a != indexed.leaves().end() && b != single.leaves().end(); ++a, ++b)
  {
    if (!(a->first == b->first)) same = false;
    for (int k = 0; k < 512 && same; ++k) same = a->second.values[k] == b->second.values[k];
    for (int w = 0; w < 8 && same; ++w) same = a->second.active[w] == b->second.active[w];
  }
  EXPECT_TRUE(same);
  // clear() removes leaves: the epoch tells the shim to drop its table
  const std::uint64_t e0 = Backend::gridEpoch(indexed);
  indexed.clear();
  EXPECT_TRUE(Backend::gridEpoch(indexed) != e0);
  // the pool survives many small jobs and odd sizes
  std::vector<int> hit(100000, 0);
  for (int rep = 0; rep < 50; ++rep) vdb_mapping::detail::WorkerPool::instance().run(1000 + 1979 * rep, [&](std::uint64_t i) { hit[i] += 1; });
  long total = 0;
  for (int h : hit) total += h;
  long expect = 0;
  for (int rep = 0; rep < 50; ++rep) expect += 1000 + 1979 * rep;
  EXPECT_EQ(total, expect);
}

int main() { return RUN_ALL_TESTS(); }
