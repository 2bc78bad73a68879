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

int main() { return RUN_ALL_TESTS(); }
