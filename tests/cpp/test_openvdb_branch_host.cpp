// CPU-only: the host-side services of the shim's OPENVDB branch (detail/backend.hpp under VDBM_HAVE_OPENVDB: leaf transfer,
// iteration, persistence, wire bytes, morphology, PCD io) executed against the API-shaped stand-ins of tests/cpp/stubs.
// No device call is made. What is checked is the branch's own logic (e.g. that a leaf written through
// buffer().data() / setValueMask reads back through cbeginLeaf / getValueMask in the ABI layout), not OpenVDB.
#define VDBM_HAVE_OPENVDB 1
#include <vdb_mapping/detail/backend.hpp>

#include <random>

#include "mini_gtest.h"

using B = vdb_mapping::detail::Backend<float>;

TEST(OpenVdbBranch, LeafTransferRoundTripInAbiLayout)
{
  std::mt19937 rng(3);
  const int n = 3000;
  std::vector<std::int32_t> origins;
  std::vector<float> values(std::size_t(n) * 512);
  std::vector<std::uint64_t> active(std::size_t(n) * 8);
  std::vector<std::uint32_t> index(n);
  for (int i = 0; i < n; ++i)
  {
    origins.insert(origins.end(), {(i % 40 - 20) * 8, (i / 40 - 30) * 8, (i % 3) * 8});
    index[i] = std::uint32_t(i);
  }
  for (auto& v : values) v = float(rng() % 2000) * 0.01f - 10.0f;
  for (auto& a : active) a = (std::uint64_t(rng()) << 32) | rng();
  auto grid = B::createMapGrid(0.05);
  std::vector<B::MapLeafT*> table;
  B::putMapLeavesIndexed(*grid, table, n, index.data(), origins.data(), values.data(), active.data());
  // a second chunk overwrites half of the leaves through the table (no tree lookup) and adds nothing new
  for (int i = 0; i < n; i += 2) values[std::size_t(i) * 512 + 7] = 99.0f;
  std::vector<std::uint32_t> idx2;
  std::vector<std::int32_t> o2;
  std::vector<float> v2;
  std::vector<std::uint64_t> a2;
  for (int i = 0; i < n; i += 2)
  {
    idx2.push_back(std::uint32_t(i));
    o2.insert(o2.end(), origins.begin() + 3 * i, origins.begin() + 3 * i + 3);
    v2.insert(v2.end(), values.begin() + std::size_t(i) * 512, values.begin() + std::size_t(i + 1) * 512);
    a2.insert(a2.end(), active.begin() + std::size_t(i) * 8, active.begin() + std::size_t(i + 1) * 8);
  }
  B::putMapLeavesIndexed(*grid, table, idx2.size(), idx2.data(), o2.data(), v2.data(), a2.data());
  EXPECT_EQ(table.size() >= std::size_t(n), true);
  std::size_t seen = 0, bad = 0;
  B::forEachMapLeaf(*grid, [&](const std::int32_t o[3], const float* v, const std::uint64_t* a) {
    ++seen;
    int i = -1;
    for (int k = 0; k < n && i < 0; ++k)
      if (origins[3 * k] == o[0] && origins[3 * k + 1] == o[1] && origins[3 * k + 2] == o[2]) i = k;
    if (i < 0 || !std::equal(v, v + 512, values.begin() + std::size_t(i) * 512) || !std::equal(a, a + 8, active.begin() + std::size_t(i) * 8)) ++bad;
  });
  EXPECT_EQ(seen, std::size_t(n));
  EXPECT_EQ(bad, std::size_t(0));
  // voxel view agrees with the leaf view: offset (x&7)<<6 | (y&7)<<3 | (z&7), mask word = offset >> 6
  auto acc = grid->getAccessor();
  const openvdb::Coord c(origins[0] + 3, origins[1] + 5, origins[2] + 6);
  const unsigned off = (3u << 6) | (5u << 3) | 6u;
  EXPECT_EQ(acc.getValue(c), values[off]);
  EXPECT_EQ(acc.isValueOn(c), bool((active[off >> 6] >> (off & 63)) & 1u));
}

TEST(OpenVdbBranch, UpdateLeavesSectionsMetaAndPersistence)
{
  auto upd = B::createUpdateGrid(0.1);
  const std::int32_t origin[3] = {8, -16, 0};
  std::uint64_t a[8] = {0x5, 0, 0, 0, 0, 0, 0, 0x8000000000000000ull}, v[8] = {0x4, 0, 0, 0, 0, 0, 0, 0};
  B::putUpdateLeaf(*upd, origin, a, v);
  EXPECT_EQ(upd->activeVoxelCount(), std::uint64_t(3));
  EXPECT_TRUE(upd->getAccessor().getValue(openvdb::Coord(8, -16, 2)));  // offset 2: active with value true
  EXPECT_FALSE(upd->getAccessor().getValue(openvdb::Coord(8, -16, 0))); // offset 0: active, value false
  int leaves = 0;
  B::forEachUpdateLeaf(*upd, [&](const std::int32_t o[3], const std::uint64_t* aa, const std::uint64_t* vv) {
    ++leaves;
    EXPECT_EQ(o[1], -16);
    EXPECT_EQ(aa[0], a[0]);
    EXPECT_EQ(aa[7], a[7]);
    EXPECT_EQ(vv[0], v[0]);
  });
  EXPECT_EQ(leaves, 1);
  const std::int32_t mn[3] = {-3, 4, 5}, mx[3] = {10, 11, 12};
  std::int32_t gmn[3], gmx[3];
  B::setSectionMeta(*upd, mn, mx);
  B::getSectionMeta(*upd, gmn, gmx);
  EXPECT_EQ(gmn[0], -3);
  EXPECT_EQ(gmx[2], 12);
  // wire bytes and files (OpenVDB's io::Stream / io::File in a real build)
  auto back = B::stringToGrid<B::UpdateGridT>(B::gridToString<B::UpdateGridT>(upd));
  EXPECT_TRUE(back != nullptr);
  EXPECT_EQ(back->activeVoxelCount(), std::uint64_t(3));
  B::getSectionMeta(*back, gmn, gmx);
  EXPECT_EQ(gmx[1], 11);
  auto map = B::createMapGrid(0.1);
  B::setVoxel(*map, openvdb::Coord(3, -9, 20), 1.5f, true);
  B::setVoxel(*map, openvdb::Coord(4, -9, 20), -2.0f, false);
  EXPECT_TRUE(B::writeGridFile("/tmp/vdbm_openvdb_branch_test.vdb", map));
  auto loaded = B::readGridFile("/tmp/vdbm_openvdb_branch_test.vdb");
  EXPECT_TRUE(loaded != nullptr);
  EXPECT_EQ(loaded->activeVoxelCount(), std::uint64_t(1));
  EXPECT_EQ(loaded->getAccessor().getValue(openvdb::Coord(4, -9, 20)), -2.0f);
  openvdb::CoordBBox bb;
  EXPECT_TRUE(B::activeBBox(*loaded, bb));
  EXPECT_EQ(bb.min().x(), 3);
  B::dilateActive(*loaded, 1); // 26-neighbourhood, R:1122-1129
  EXPECT_EQ(loaded->activeVoxelCount(), std::uint64_t(27));
  B::erodeActive(*loaded, 1);
  EXPECT_EQ(loaded->activeVoxelCount(), std::uint64_t(1));
  int active_voxels = 0;
  B::forEachActiveVoxel(*loaded, [&](const openvdb::Coord& c, const float& val) {
    ++active_voxels;
    EXPECT_EQ(c.z(), 20);
    EXPECT_EQ(val, 1.5f);
  });
  EXPECT_EQ(active_voxels, 1);
  B::prune(*loaded);
}

TEST(OpenVdbBranch, PcdThroughPclsInterface)
{
  B::PointCloudT c, bin, asc;
  for (int i = 0; i < 100; ++i) c.points.emplace_back(0.1f * i, -0.3f * i, 7.25f);
  EXPECT_TRUE(vdb_mapping::detail::writePCD("/tmp/vdbm_openvdb_branch_bin.pcd", c)); // binary, as saveMapToPCD's helper writes it
  EXPECT_TRUE(B::loadPCD("/tmp/vdbm_openvdb_branch_bin.pcd", bin));
  EXPECT_EQ(bin.points.size(), std::size_t(100));
  EXPECT_EQ(bin.points[99].y, c.points[99].y);
  EXPECT_TRUE(B::savePCD("/tmp/vdbm_openvdb_branch_asc.pcd", c));
  EXPECT_TRUE(B::loadPCD("/tmp/vdbm_openvdb_branch_asc.pcd", asc));
  EXPECT_EQ(asc.points[42].x, c.points[42].x);
  EXPECT_FALSE(B::loadPCD("/tmp/vdbm_definitely_not_there.pcd", asc));
}

int main() { return RUN_ALL_TESTS(); }
