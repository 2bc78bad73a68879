// C++ drop-in check: the reference's known-answer scenarios (/root/reference/tests/mapping.cpp) driven through the
// SHIM headers (include/vdb_mapping/*.hpp) -> C ABI -> CUDA kernels, written against the same public API the
// reference's tests use (OccupancyVDBMapping, PointCloudT, Config, getGrid()->getAccessor(), openvdb::Coord).
// Table-driven instead of one TEST per scenario; plus API cases the reference does not test (updateMap with a
// caller grid, raycastPointCloud into an accessor, sections, lazy mirror).
#include <algorithm>
#include <array>
#include <chrono>
#include <cmath>
#include <map>
#include <thread>
#include <vector>
#include <vdb_mapping/OccupancyVDBMapping.hpp>

#include "mini_gtest.h"

using vdb_mapping::Config;
using vdb_mapping::OccupancyVDBMapping;

static Config gtestConfig(double max_range)
{
  Config conf;
  conf.max_range           = max_range;
  conf.fast_mode           = false;
  conf.accumulation_period = 0.0; // the reference leaves it uninitialised; 0 = only explicit integrates
  conf.prob_hit            = 0.9;
  conf.prob_miss           = 0.1;
  conf.prob_thres_max      = 0.51;
  conf.prob_thres_min      = 0.49;
  return conf;
}
static float logOdds(double p) { return static_cast<float>(std::log(p) - std::log(1 - p)); }


// centres of all active voxels, like saveMapToPCD R:218-256 writes them
static void detail_forEachActive(OccupancyVDBMapping& map, OccupancyVDBMapping::PointCloudT& cloud)
{
  auto grid = map.getGrid();
  vdb_mapping::detail::Backend<float>::forEachActiveVoxel(*grid, [&](const openvdb::Coord& c, const float&) {
    const openvdb::Vec3d w = grid->indexToWorld(c);
    cloud.points.emplace_back(static_cast<float>(w.x() + 0.05), static_cast<float>(w.y() + 0.05), static_cast<float>(w.z() + 0.05));
  });
}

// leaves of a float grid, whatever the backend (openvdb::Grid has tree().leafCount(), the stand-in grid leafCount())
template <typename G>
static std::size_t leafCountOf(const G& grid)
{
  std::size_t n = 0;
  vdb_mapping::detail::Backend<float>::forEachMapLeaf(grid, [&](const std::int32_t*, const float*, const std::uint64_t*) { ++n; });
  return n;
}

struct Expect { int z; int kind; bool check_flag; bool flag; }; // kind: 0 untouched, 1 miss, 2 hit

static void runAxisCase(double resolution, double max_range, double z_in_resolutions, std::initializer_list<Expect> exp)
{
  OccupancyVDBMapping map(resolution);
  const Config conf = gtestConfig(max_range);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0, 0, z_in_resolutions * resolution);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  EXPECT_TRUE(map.insertPointCloud(cloud, origin, "test"));
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  for (const Expect& e : exp)
  {
    const openvdb::Coord c(0, 0, e.z);
    const float want = e.kind == 0 ? 0.0f : (e.kind == 1 ? logOdds(conf.prob_miss) : logOdds(conf.prob_hit));
    EXPECT_EQ(acc.getValue(c), want);
    if (e.check_flag) EXPECT_EQ(acc.isValueOn(c), e.flag);
  }
}

TEST(Shim, InsertBeforeConfigIsANoOpAndAccessorStaysLive)
{
  OccupancyVDBMapping map(1);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0, 0, 1);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  map.insertPointCloud(cloud, origin, "test"); // no source, no config: nothing happens
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 1)), 0.0f);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  map.insertPointCloud(cloud, origin, "test");
  // read through the accessor obtained BEFORE the insert (eager mirror keeps the grid object alive and current)
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 0)), logOdds(0.1));
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 1)), logOdds(0.9));
}

TEST(Shim, PositiveAxisRay)
{
  runAxisCase(0.1, 10, 5, {{0, 1, true, false}, {1, 1, true, false}, {2, 1, true, false}, {3, 1, true, false}, {4, 1, true, false}, {5, 2, true, true}});
}
TEST(Shim, NegativeAxisRay)
{
  runAxisCase(0.1, 10, -5, {{-1, 1, true, false}, {-2, 1, true, false}, {-3, 1, true, false}, {-4, 1, true, false}, {-5, 2, true, true}});
}
TEST(Shim, MaxRangeRayFreesTheClippedVoxelOnly)
{
  runAxisCase(0.1, 0.5, 7, {{0, 1, true, false}, {1, 1, true, false}, {2, 1, true, false}, {3, 1, true, false}, {4, 1, true, false}, {5, 1, true, false}, {6, 0, true, false}});
}

TEST(Shim, ResetMapGivesAFreshGrid)
{
  OccupancyVDBMapping map(1);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0, 0, 1);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  map.insertPointCloud(cloud, origin, "test");
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 1)), logOdds(0.9));
  map.resetMap();
  acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 1)), 0.0f);
}

TEST(Shim, RaycastIntoAccessorThenUpdateMapReturnsChangeGrid)
{
  OccupancyVDBMapping map(0.1);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0.35f, 0.0f, 0.0f);
  cloud->points.emplace_back(0.0f, -0.25f, 0.0f);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  OccupancyVDBMapping::UpdateGridT::Ptr upd = OccupancyVDBMapping::UpdateGridT::create(false);
  OccupancyVDBMapping::UpdateGridT::Accessor uacc = upd->getAccessor();
  EXPECT_TRUE(map.raycastPointCloud(cloud, origin, 10.0, uacc));
  EXPECT_TRUE(uacc.isValueOn(openvdb::Coord(0, 0, 0)));
  EXPECT_FALSE(uacc.getValue(openvdb::Coord(0, 0, 0)));      // free-space voxel: active, value false
  EXPECT_TRUE(uacc.getValue(openvdb::Coord(3, 0, 0)));       // endpoint of the first ray: hit
  EXPECT_TRUE(uacc.getValue(openvdb::Coord(0, -2, 0)));      // endpoint of the second ray (worldToIndex half-voxel rule)
  OccupancyVDBMapping::UpdateGridT::Ptr change = map.updateMap(upd);
  OccupancyVDBMapping::UpdateGridT::Accessor cacc = change->getAccessor();
  EXPECT_TRUE(cacc.isValueOn(openvdb::Coord(3, 0, 0)));      // hit flipped inactive -> active
  EXPECT_TRUE(cacc.getValue(openvdb::Coord(3, 0, 0)));
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(3, 0, 0)), logOdds(0.9));
  EXPECT_EQ(acc.getValue(openvdb::Coord(1, 0, 0)), logOdds(0.1));
  // a section around the first endpoint contains exactly that active voxel
  Eigen::Matrix<double, 3, 1> mn(0.25, -0.05, -0.05), mx(0.39, 0.05, 0.05);
  auto section = map.getMapSectionUpdateGrid(mn, mx, Eigen::Matrix<double, 4, 4>::Identity(), false);
  EXPECT_EQ(section->activeVoxelCount(), std::uint64_t(1));
  EXPECT_TRUE(section->getAccessor().isValueOn(openvdb::Coord(3, 0, 0)));
}

TEST(Shim, SectionsTravelFromOneMapToAnother)
{
  // remote mapping round trip: sender map -> getMapSection*Grid -> receiver applyMapSection*Grid
  const Config conf = gtestConfig(10);
  OccupancyVDBMapping sender(0.1), receiver(0.1);
  for (OccupancyVDBMapping* m : {&sender, &receiver})
  {
    m->setConfig(conf);
    m->addInputSource("test", conf.max_range, 0);
  }
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0.45f, 0.0f, 0.0f);
  cloud->points.emplace_back(0.0f, 0.35f, 0.1f);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  sender.insertPointCloud(cloud, origin, "test");
  Eigen::Matrix<double, 3, 1> mn(-1, -1, -1), mx(1, 1, 1);
  const auto I4 = Eigen::Matrix<double, 4, 4>::Identity();
  auto full     = sender.getMapSectionGrid(mn, mx, I4, true);
  receiver.applyMapSectionGrid(full);
  OccupancyVDBMapping::GridT::Accessor racc = receiver.getGrid()->getAccessor();
  EXPECT_EQ(racc.getValue(openvdb::Coord(4, 0, 0)), logOdds(0.9));
  EXPECT_TRUE(racc.isValueOn(openvdb::Coord(4, 0, 0)));
  EXPECT_EQ(racc.getValue(openvdb::Coord(2, 0, 0)), logOdds(0.1));
  // update-grid section: receiver gets exactly the sender's active voxels inside the box
  OccupancyVDBMapping third(0.1);
  third.setConfig(conf);
  third.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr other(new OccupancyVDBMapping::PointCloudT);
  other->points.emplace_back(-0.25f, 0.0f, 0.0f);
  third.insertPointCloud(other, origin, "test");            // active voxel (-2,0,0) inside the box -> must be deactivated
  auto sparse = sender.getMapSectionUpdateGrid(mn, mx, I4, false);
  third.applyMapSectionUpdateGrid(sparse);
  OccupancyVDBMapping::GridT::Accessor tacc = third.getGrid()->getAccessor();
  EXPECT_FALSE(tacc.isValueOn(openvdb::Coord(-2, 0, 0)));
  EXPECT_EQ(tacc.getValue(openvdb::Coord(-2, 0, 0)), logOdds(0.9)); // value kept, only the flag changes
  EXPECT_TRUE(tacc.isValueOn(openvdb::Coord(4, 0, 0)));
  EXPECT_EQ(tacc.getValue(openvdb::Coord(4, 0, 0)), 0.0f);          // activated with the background value
}

TEST(Shim, LazyMirrorSyncsOnGetGrid)
{
  OccupancyVDBMapping map(0.1);
  map.setMirrorMode(vdb_mapping::MirrorMode::Lazy);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0, 0, 0.5f);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  map.insertPointCloud(cloud, origin, "test");
  map.insertPointCloud(cloud, origin, "test");
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(0, 0, 5)), logOdds(0.9) + logOdds(0.9));
  vdbm_stats_t st;
  EXPECT_TRUE(map.deviceStats(st));
  EXPECT_EQ(st.rays, std::uint64_t(2));
}

TEST(Shim, ReducedUpdateReproducesTheSenderOnARemoteMap)
{
  // createUpdate(level 2) ships one voxel per ray + the origin; applyUpdate re-raycasts it: the remote map equals the
  // sender's. Level 1 (the change grid) pins the flipped voxels to the clamping bounds.
  const Config conf = gtestConfig(10);
  OccupancyVDBMapping sender(0.1), remote(0.1), overwritten(0.1);
  for (OccupancyVDBMapping* m : {&sender, &remote, &overwritten})
  {
    m->setConfig(conf);
    m->addInputSource("test", conf.max_range, 0);
  }
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0.45f, 0.13f, 0.0f);
  cloud->points.emplace_back(-0.2f, 0.35f, 0.1f);
  cloud->points.emplace_back(30.0f, 1.0f, 0.5f); // clipped at max range: carves free space, no hit
  Eigen::Matrix<double, 3, 1> origin(0.013, 0.02, 0.0), got_origin(9, 9, 9);
  sender.accumulateUpdate(cloud, origin, "test");
  auto reduced = sender.createUpdate("test", 2, &got_origin);
  auto raw     = sender.createUpdate("test", 0);
  EXPECT_EQ(got_origin.x(), origin.x());
  EXPECT_EQ(got_origin.z(), origin.z());
  EXPECT_EQ(reduced->activeVoxelCount(), std::uint64_t(3));
  EXPECT_TRUE(raw->activeVoxelCount() > 100);
  sender.integrateUpdate();
  auto change = remote.applyUpdate(reduced, 2, got_origin);
  EXPECT_TRUE(change->getAccessor().getValue(openvdb::Coord(4, 1, 0))); // hit flipped to active
  OccupancyVDBMapping::GridT::Accessor sacc = sender.getGrid()->getAccessor();
  OccupancyVDBMapping::GridT::Accessor racc = remote.getGrid()->getAccessor();
  EXPECT_EQ(sender.getGrid()->activeVoxelCount(), remote.getGrid()->activeVoxelCount());
  for (int x = -3; x <= 100; ++x)
    for (int y = -1; y <= 5; ++y)
      for (int z = -1; z <= 2; ++z)
      {
        const openvdb::Coord c(x, y, z);
        EXPECT_EQ(sacc.getValue(c), racc.getValue(c));
        EXPECT_EQ(sacc.isValueOn(c), racc.isValueOn(c));
      }
  overwritten.applyUpdate(change, 1);
  OccupancyVDBMapping::GridT::Accessor oacc = overwritten.getGrid()->getAccessor();
  EXPECT_TRUE(oacc.isValueOn(openvdb::Coord(4, 1, 0)));
  EXPECT_EQ(oacc.getValue(openvdb::Coord(4, 1, 0)), logOdds(0.99));
}

TEST(Shim, PointEditsAndArtificialAreas)
{
  OccupancyVDBMapping map(0.1);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::PointCloudT::Ptr pts(new OccupancyVDBMapping::PointCloudT);
  pts->points.emplace_back(0.55f, 0.0f, 0.0f);
  pts->points.emplace_back(0.2f, 0.0f, 0.0f); // exactly on a voxel boundary: plain floor(p / res) -> voxel 2
  EXPECT_TRUE(map.addPointsToGrid(pts));
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(5, 0, 0)));
  EXPECT_EQ(acc.getValue(openvdb::Coord(5, 0, 0)), logOdds(0.99));
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(2, 0, 0)));
  EXPECT_TRUE(map.removePointsFromGrid(pts));
  EXPECT_FALSE(acc.isValueOn(openvdb::Coord(5, 0, 0)));
  EXPECT_EQ(acc.getValue(openvdb::Coord(5, 0, 0)), logOdds(0.01));
  // a wall from (1,1) to (1,2) m, heights [-0.1, 0.2): invisible until the next updateMap, then always active
  std::vector<std::vector<Eigen::Matrix<double, 4, 1> > > areas(1);
  Eigen::Matrix<double, 4, 1> a, b;
  a[0] = 1.0; a[1] = 1.0; a[2] = 0.0; a[3] = 1.0;
  b[0] = 1.0; b[1] = 2.0; b[2] = 0.0; b[3] = 1.0;
  areas[0] = {a, b};
  map.addArtificialAreas(areas, -0.1, 0.2);
  EXPECT_FALSE(acc.isValueOn(openvdb::Coord(10, 15, 0)));
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0.0f, 0.0f, 0.5f);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  map.insertPointCloud(cloud, origin, "test");
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(10, 15, 0)));
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(10, 15, 1)));
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(10, 15, -1)));
  EXPECT_FALSE(acc.isValueOn(openvdb::Coord(10, 15, 2)));
  EXPECT_EQ(acc.getValue(openvdb::Coord(10, 15, 0)), 0.0f);
  map.restoreMapIntegrity(); // value 0 is not above the occupancy threshold -> inactive again
  EXPECT_FALSE(acc.isValueOn(openvdb::Coord(10, 15, 0)));
}


// ---- the rest of the reference's public surface ------------------------------------------------
static OccupancyVDBMapping::PointCloudT::Ptr wallCloud(float x, int n_side, float step)
{
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  for (int i = -n_side; i <= n_side; ++i)
    for (int j = -n_side; j <= n_side; ++j) cloud->points.emplace_back(x, i * step, j * step);
  return cloud;
}

TEST(Shim, RaytraceFindsTheWallAndFastModeOnlyTouchesOccupiedVoxels)
{
  OccupancyVDBMapping map(0.1);
  Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  // world 2.0 -> voxel 20 and world 0.0 -> voxel 0 under the reference's worldToIndex rule (R:612-631)
  Eigen::Matrix<double, 3, 1> origin(0.0, 0.0, 0.0);
  map.insertPointCloud(wallCloud(2.0f, 10, 0.1f), origin, "test"); // a wall in the voxel plane x = 20; one hit is enough with this config
  bool success = false;
  openvdb::Vec3d end;
  map.raytrace(openvdb::Vec3d(0.05, 0.05, 0.05), openvdb::Vec3d(1, 0, 0), 5.0, success, end);
  EXPECT_TRUE(success);
  EXPECT_TRUE(std::fabs(end.x() - 2.0) < 1e-9); // indexToWorld of voxel 20 = its lower corner
  // behind the sensor: the ray misses the active bounding box, but the reference ignores setIndexRay's result (R:700) and
  // marches the unclipped ray through the free-space leaf it starts in -> "success" with an end point behind the sensor
  map.raytrace(openvdb::Vec3d(0.05, 0.05, 0.05), openvdb::Vec3d(-1, 0, 0), 5.0, success, end);
  EXPECT_TRUE(end.x() < 0.05);
  // far away from everything mapped: a miss, end point = origin + direction * length
  map.raytrace(openvdb::Vec3d(50.0, 50.0, 50.0), openvdb::Vec3d(0, 0, 1), 5.0, success, end);
  EXPECT_FALSE(success);
  EXPECT_TRUE(std::fabs(end.z() - 55.0) < 1e-9);
  std::vector<openvdb::Vec3d> origins(3, openvdb::Vec3d(0.05, 0.05, 0.05)), dirs = {openvdb::Vec3d(1, 0, 0), openvdb::Vec3d(2, 0.1, 0), openvdb::Vec3d(0, 0, -1)}, ends;
  std::vector<double> lens = {5.0, 5.0, 5.0};
  std::vector<bool> oks;
  map.raytrace(origins, dirs, lens, oks, ends);
  EXPECT_EQ(oks.size(), std::size_t(3));
  EXPECT_TRUE(oks[0] && oks[1]);

  // fast mode: a ray THROUGH the wall frees the wall voxel it passes, but creates no free-space voxels behind it
  conf.fast_mode = true;
  map.setConfig(conf);
  const std::size_t leaves_before = leafCountOf(*map.getGrid());
  OccupancyVDBMapping::PointCloudT::Ptr through(new OccupancyVDBMapping::PointCloudT);
  through->points.emplace_back(6.0f, 0.0f, 0.0f);
  map.insertPointCloud(through, origin, "test");
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_EQ(acc.getValue(openvdb::Coord(20, 0, 0)), logOdds(0.9) + logOdds(0.1)); // hit, then the fast-mode miss
  EXPECT_EQ(acc.getValue(openvdb::Coord(40, 0, 0)), 0.0f);                        // behind the wall: untouched
  EXPECT_EQ(acc.getValue(openvdb::Coord(60, 0, 0)), logOdds(0.9));                // the new end point
  EXPECT_TRUE(leafCountOf(*map.getGrid()) <= leaves_before + 1);
}

TEST(Shim, SaveLoadRoundTripAndPcdImport)
{
  OccupancyVDBMapping map(0.1);
  Config conf             = gtestConfig(10);
  conf.map_directory_path = "/tmp/vdbm_shim_test_";
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  Eigen::Matrix<double, 3, 1> origin(0.0, 0.0, 0.0);
  map.insertPointCloud(wallCloud(1.5f, 6, 0.1f), origin, "test"); // voxel plane x = 15
  const auto before_active = map.getGrid()->activeVoxelCount();
  EXPECT_TRUE(before_active > 100);
  // save through gridToByteArray / byteArrayToGrid (wire codec R:1310-1338) and through a file
  std::vector<uint8_t> bytes = map.gridToByteArray<OccupancyVDBMapping::GridT>(map.getGrid());
  EXPECT_TRUE(bytes.size() > 16);
  OccupancyVDBMapping::GridT::Ptr back = map.byteArrayToGrid<OccupancyVDBMapping::GridT>(bytes);
  EXPECT_EQ(back->activeVoxelCount(), before_active);
  EXPECT_EQ(leafCountOf(*back), leafCountOf(*map.getGrid()));
  const std::string raw = "a string that should survive the zstd codec  a string that should survive the zstd codec";
  EXPECT_EQ(map.decompressByteArray(map.compressString(raw)), raw);
  EXPECT_TRUE(map.saveMapToPCD());
  // a PCD written by saveMapToPCD read back into a fresh map (loadMapFromPCD R:295 -> createMapFromPointCloud O:136)
  OccupancyVDBMapping::PointCloudT::Ptr centres(new OccupancyVDBMapping::PointCloudT);
  detail_forEachActive(map, *centres);
  const std::string pcd = "/tmp/vdbm_shim_test_points.pcd";
  EXPECT_TRUE(vdb_mapping::detail::writePCD(pcd, *centres));
  OccupancyVDBMapping other(0.1);
  other.setConfig(conf);
  other.addInputSource("test", conf.max_range, 0);
  EXPECT_TRUE(other.loadMapFromPCD(pcd, /*set_background=*/false, /*clear_map=*/true));
  EXPECT_EQ(other.getGrid()->activeVoxelCount(), before_active);
  OccupancyVDBMapping::GridT::Accessor oacc = other.getGrid()->getAccessor();
  EXPECT_TRUE(oacc.isValueOn(openvdb::Coord(15, 0, 0)));
  EXPECT_EQ(oacc.getValue(openvdb::Coord(15, 0, 0)), logOdds(0.99));
  // the loaded map lives on the DEVICE too: a scan integrates into it
  OccupancyVDBMapping::PointCloudT::Ptr one(new OccupancyVDBMapping::PointCloudT);
  one->points.emplace_back(1.5f, 0.0f, 0.0f);
  other.insertPointCloud(one, origin, "test");
  EXPECT_EQ(oacc.getValue(openvdb::Coord(15, 0, 0)), logOdds(0.99)); // already at the clamp
  EXPECT_EQ(oacc.getValue(openvdb::Coord(7, 0, 0)), logOdds(0.1));
  EXPECT_FALSE(other.loadMapFromPCD("/tmp/definitely_not_there.pcd", false, true));
}

TEST(Shim, ExplicitRaysWallsAndTypedSections)
{
  OccupancyVDBMapping map(0.1);
  const Config conf = gtestConfig(10);
  map.setConfig(conf);
  map.addInputSource("test", conf.max_range, 0);
  OccupancyVDBMapping::UpdateGridT::Ptr grid = OccupancyVDBMapping::UpdateGridT::create(false);
  OccupancyVDBMapping::UpdateGridT::Accessor uacc = grid->getAccessor();
  map.castRayIntoGrid(openvdb::Coord(0, 0, 0), openvdb::Coord(5, 0, 0), uacc);
  for (int x = 0; x <= 5; ++x) EXPECT_TRUE(uacc.isValueOn(openvdb::Coord(x, 0, 0)));
  EXPECT_FALSE(uacc.isValueOn(openvdb::Coord(6, 0, 0)));
  map.castRayIntoGrid(openvdb::Coord(3, 3, 3), openvdb::Coord(3, 3, 3), uacc); // start == end: nothing (R:559)
  EXPECT_FALSE(uacc.isValueOn(openvdb::Coord(3, 3, 3)));

  Eigen::Matrix<double, 4, 1> a(1.0, 1.0, 0.0, 1.0), b(1.0, 2.0, 0.0, 1.0);
  map.addArtificialWall(a, b, -0.1, 0.2);
  std::vector<Eigen::Matrix<double, 4, 1> > tri = {Eigen::Matrix<double, 4, 1>(-1.0, -1.0, 0.0, 1.0), Eigen::Matrix<double, 4, 1>(-2.0, -1.0, 0.0, 1.0),
                                                   Eigen::Matrix<double, 4, 1>(-2.0, -2.0, 0.0, 1.0)};
  map.addArtificialPolygon(tri, 0.0, 0.1);
  OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
  cloud->points.emplace_back(0.0f, 0.0f, 0.5f);
  Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  map.insertPointCloud(cloud, origin, "test");
  OccupancyVDBMapping::GridT::Accessor acc = map.getGrid()->getAccessor();
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(10, 15, 0)));   // the first wall survived the second call (no restore in between)
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(-15, -10, 0))); // polygon edge 0
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(-20, -15, 0))); // polygon edge 1
  EXPECT_TRUE(acc.isValueOn(openvdb::Coord(-15, -15, 0))); // the closing edge (diagonal)

  Eigen::Matrix<double, 3, 1> lo(-3, -3, -1), hi(3, 3, 1);
  Eigen::Matrix<double, 4, 4> tf = Eigen::Matrix<double, 4, 4>::Identity();
  auto ug = map.getMapSection<OccupancyVDBMapping::UpdateGridT>(lo, hi, tf, false);
  auto fg = map.getMapSection<OccupancyVDBMapping::GridT>(lo, hi, tf, true);
  EXPECT_TRUE(ug->activeVoxelCount() > 0);
  EXPECT_EQ(ug->activeVoxelCount(), fg->activeVoxelCount());
  const openvdb::BBoxd wb = map.createWorldBoundingBox(lo, hi, tf);
  EXPECT_EQ(wb.min().x(), -3.0);
  EXPECT_EQ(wb.max().z(), 1.0);
  map.morphologicalCloseMap<OccupancyVDBMapping::UpdateGridT>(ug, 1); // runs; result checked in test_compat_grid
}

// a quarter of a synthetic room scan per source (deterministic, no <random>: the two maps must see identical clouds)
static OccupancyVDBMapping::PointCloudT::Ptr sectorCloud(int sector, int n, unsigned seed)
{
  OccupancyVDBMapping::PointCloudT::Ptr c(new OccupancyVDBMapping::PointCloudT);
  unsigned state = seed * 2654435761u + 12345u;
  auto rnd = [&] { state = state * 1664525u + 1013904223u; return float((state >> 8) & 0xFFFF) / 65536.0f; };
  for (int i = 0; i < n; ++i)
  {
    const float az = (float(sector) + rnd()) * 1.5707963f, el = (rnd() - 0.5f) * 0.8f, r = 2.0f + 9.0f * rnd();
    c->points.emplace_back(r * std::cos(el) * std::cos(az), r * std::cos(el) * std::sin(az), r * std::sin(el));
  }
  return c;
}

static bool gridsIdentical(OccupancyVDBMapping& a, OccupancyVDBMapping& b)
{
  // every leaf, every voxel value (free space is inactive but carries log-odds) and every active flag
  using B = vdb_mapping::detail::Backend<float>;
  struct Leaf { std::array<float, 512> v; std::array<std::uint64_t, 8> m; };
  std::map<std::array<std::int32_t, 3>, Leaf> la;
  B::forEachMapLeaf(*a.getGrid(), [&](const std::int32_t o[3], const float* v, const std::uint64_t* m) {
    Leaf& l = la[{o[0], o[1], o[2]}];
    std::copy(v, v + 512, l.v.begin());
    std::copy(m, m + 8, l.m.begin());
  });
  bool same = true;
  std::size_t n = 0;
  B::forEachMapLeaf(*b.getGrid(), [&](const std::int32_t o[3], const float* v, const std::uint64_t* m) {
    ++n;
    auto it = la.find({o[0], o[1], o[2]});
    if (it == la.end() || !std::equal(v, v + 512, it->second.v.begin()) || !std::equal(m, m + 8, it->second.m.begin())) same = false;
  });
  return same && n == la.size() && n > 50;
}

TEST(Shim, SourcesOnTheirOwnHandlesAccumulateConcurrentlyAndGiveTheSameMap)
{
  // SourceConcurrency: four sources fed from four threads. PerSource (what Auto picks for > 1 source): every source raycasts on
  // its own device handle, integrateUpdate gathers the update leaves device-to-device. Shared: the threads take turns on the
  // map's handle. Same clouds -> the maps must be identical voxel for voxel (updateMap runs per source in key order in both).
  const Config conf = gtestConfig(8);
  OccupancyVDBMapping per_source(0.1), shared(0.1);
  shared.setSourceConcurrency(vdb_mapping::SourceConcurrency::Shared);
  const char* ids[4] = {"lidar_a", "lidar_b", "lidar_c", "lidar_d"};
  for (OccupancyVDBMapping* m : {&per_source, &shared})
  {
    m->setConfig(conf);
    m->setMirrorMode(vdb_mapping::MirrorMode::Lazy);
    for (const char* id : ids) m->addInputSource(id, id == ids[3] ? 5.0 : 0.0, 0); // one source with its own (shorter) range
  }
  for (int scan = 0; scan < 3; ++scan)
  {
    const Eigen::Matrix<double, 3, 1> origin(0.03 * scan, -0.02 * scan, 0.01);
    std::vector<OccupancyVDBMapping::PointCloudT::Ptr> clouds;
    for (int s = 0; s < 4; ++s) clouds.push_back(sectorCloud(s, 6000, 17u * unsigned(scan) + unsigned(s)));
    for (OccupancyVDBMapping* m : {&per_source, &shared})
    {
      std::vector<std::thread> th;
      for (int s = 0; s < 4; ++s) th.emplace_back([&, s] { m->accumulateUpdate(clouds[s], origin, ids[s]); });
      for (auto& t : th) t.join();
      if (scan == 1)
      {
        // deltas while the data still sits on the sources' own handles: raw grid and reduced update of one source
        auto raw = m->createUpdate(ids[2], 0);
        EXPECT_TRUE(raw->activeVoxelCount() > 6000);
        if (m == &per_source)
        {
          // (on a shared handle level 2 describes the LAST accumulate of the handle, whichever source's thread came last)
          Eigen::Matrix<double, 3, 1> o(9, 9, 9);
          auto reduced = m->createUpdate(ids[2], 2, &o);
          EXPECT_EQ(o.x(), origin.x());
          EXPECT_TRUE(reduced->activeVoxelCount() > 3000 && reduced->activeVoxelCount() <= 6000);
          EXPECT_TRUE(raw->activeVoxelCount() > reduced->activeVoxelCount());
        }
      }
      m->integrateUpdate();
    }
    EXPECT_TRUE(gridsIdentical(per_source, shared));
  }
  // insertPointCloud (accumulate + integrate) of one source while the others are idle, then a reset
  auto extra = sectorCloud(1, 3000, 99u);
  for (OccupancyVDBMapping* m : {&per_source, &shared}) EXPECT_TRUE(m->insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0, 0, 0), ids[1]));
  EXPECT_TRUE(gridsIdentical(per_source, shared));
  vdbm_stats_t sp, ss;
  EXPECT_TRUE(per_source.deviceStats(sp) && shared.deviceStats(ss));
  EXPECT_EQ(sp.voxel_updates, ss.voxel_updates); // the map's handle integrated the same update voxels
  EXPECT_EQ(sp.rays, ss.rays);                   // the raycast counters include the sources' own handles
  EXPECT_EQ(sp.visits, ss.visits);
  EXPECT_EQ(sp.clipped, ss.clipped);
  per_source.resetMap();
  EXPECT_EQ(per_source.getGrid()->activeVoxelCount(), std::uint64_t(0));
  per_source.insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0, 0, 0), ids[1]);
  shared.resetMap();
  shared.insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0, 0, 0), ids[1]);
  EXPECT_TRUE(gridsIdentical(per_source, shared));
}

TEST(Shim, EagerMirrorFollowsLargeScansThroughTheChunkedStream)
{
  // the default (Eager) mirror on scans that touch thousands of leaves, with a chunk size small enough for many chunks per
  // insert: an accessor taken BEFORE the inserts reads the same values as a lazily mirrored twin afterwards.
  const Config conf = gtestConfig(12);
  OccupancyVDBMapping eager(0.05), lazy(0.05);
  lazy.setMirrorMode(vdb_mapping::MirrorMode::Lazy);
  eager.setMirrorChunkLeaves(512);
  for (OccupancyVDBMapping* m : {&eager, &lazy})
  {
    m->setConfig(conf);
    m->addInputSource("lidar", 0.0, 0);
  }
  OccupancyVDBMapping::GridT::Accessor early = eager.getGrid()->getAccessor();
  for (int scan = 0; scan < 4; ++scan)
  {
    OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
    for (int s = 0; s < 4; ++s)
    {
      const auto part = sectorCloud(s, 5000, 5u * unsigned(scan) + unsigned(s));
      cloud->points.insert(cloud->points.end(), part->points.begin(), part->points.end());
    }
    const Eigen::Matrix<double, 3, 1> origin(0.1 * scan, 0.05 * scan, 0.0);
    eager.insertPointCloud(cloud, origin, "lidar");
    lazy.insertPointCloud(cloud, origin, "lidar");
  }
  EXPECT_TRUE(gridsIdentical(eager, lazy));
  std::uint64_t n = 0, bad = 0;
  vdb_mapping::detail::Backend<float>::forEachActiveVoxel(*lazy.getGrid(), [&](const openvdb::Coord& c, const float& v) {
    ++n;
    if (!early.isValueOn(c) || early.getValue(c) != v) ++bad;
  });
  EXPECT_TRUE(n > 1000);
  EXPECT_EQ(bad, std::uint64_t(0));
  eager.resetMap(); // new generation of the device pool, new host grid: the leaf table must not survive
  eager.insertPointCloud(sectorCloud(0, 2000, 1u), Eigen::Matrix<double, 3, 1>(0, 0, 0), "lidar");
  lazy.resetMap();
  lazy.insertPointCloud(sectorCloud(0, 2000, 1u), Eigen::Matrix<double, 3, 1>(0, 0, 0), "lidar");
  EXPECT_TRUE(gridsIdentical(eager, lazy));
}

TEST(Shim, MapShardedOverSeveralDevicesEqualsTheSingleDeviceMap)
{
  // setDevices(): one process, the map sharded by azimuth sector over several device handles (vdbm_group_*). The shards may
  // share a GPU, which is how this runs on a one-GPU box; tests/test_group.py and the 2 / 8-GPU logs under profiles/ cover
  // real device sets. Host grid, sections and counters must equal the single-device map's.
  const Config conf = gtestConfig(12);
  OccupancyVDBMapping sharded(0.05), single(0.05);
  EXPECT_TRUE(sharded.setDevices({0, 0, 0}));
  EXPECT_FALSE(sharded.setDevices({0}));                       // once only
  for (OccupancyVDBMapping* m : {&sharded, &single})
  {
    m->setConfig(conf);
    m->addInputSource("lidar", 0.0, 0);
  }
  OccupancyVDBMapping late(0.05);
  late.setConfig(conf);
  EXPECT_FALSE(late.setDevices({0, 0}));                       // after setConfig: refused, the map stays on one device
  OccupancyVDBMapping::GridT::Accessor early = sharded.getGrid()->getAccessor(); // Eager mirror fed by all shards
  for (int scan = 0; scan < 4; ++scan)
  {
    OccupancyVDBMapping::PointCloudT::Ptr cloud(new OccupancyVDBMapping::PointCloudT);
    for (int s = 0; s < 4; ++s)
    {
      const auto part = sectorCloud(s, 5000, 31u * unsigned(scan) + unsigned(s));
      cloud->points.insert(cloud->points.end(), part->points.begin(), part->points.end());
    }
    const Eigen::Matrix<double, 3, 1> origin(0.07 * scan, -0.04 * scan, 0.02);
    EXPECT_TRUE(sharded.insertPointCloud(cloud, origin, "lidar"));
    EXPECT_TRUE(single.insertPointCloud(cloud, origin, "lidar"));
    EXPECT_TRUE(gridsIdentical(sharded, single));
  }
  std::uint64_t n = 0, bad = 0;
  vdb_mapping::detail::Backend<float>::forEachActiveVoxel(*single.getGrid(), [&](const openvdb::Coord& c, const float& v) {
    ++n;
    if (!early.isValueOn(c) || early.getValue(c) != v) ++bad;
  });
  EXPECT_TRUE(n > 1000);
  EXPECT_EQ(bad, std::uint64_t(0));
  // sections are gathered from all shards
  const Eigen::Matrix<double, 3, 1> lo(-3.0, -2.0, -1.0), hi(2.5, 3.0, 1.5);
  const auto tf = Eigen::Matrix<double, 4, 4>::Identity();
  for (bool full : {false, true})
  {
    auto a = sharded.getMapSectionGrid(lo, hi, tf, full);
    auto b = single.getMapSectionGrid(lo, hi, tf, full);
    EXPECT_EQ(a->activeVoxelCount(), b->activeVoxelCount());
    EXPECT_EQ(leafCountOf(*a), leafCountOf(*b));
    auto ua = sharded.getMapSectionUpdateGrid(lo, hi, tf, full);
    auto ub = single.getMapSectionUpdateGrid(lo, hi, tf, full);
    EXPECT_EQ(ua->activeVoxelCount(), ub->activeVoxelCount());
    EXPECT_TRUE(b->activeVoxelCount() > 0);
  }
  vdbm_stats_t a, b;
  EXPECT_TRUE(sharded.deviceStats(a) && single.deviceStats(b));
  EXPECT_EQ(a.rays, b.rays);
  EXPECT_EQ(a.voxel_updates, b.voxel_updates);
  EXPECT_EQ(a.map_leaves, b.map_leaves);
  // what needs the whole map on one device says so and leaves the map alone
  bool ok = true;
  openvdb::Vec3d end;
  sharded.raytrace(openvdb::Vec3d(0, 0, 0), openvdb::Vec3d(1, 0, 0), 5.0, ok, end);
  EXPECT_FALSE(ok);
  EXPECT_EQ(sharded.createUpdate("lidar", 0)->activeVoxelCount(), std::uint64_t(0));
  EXPECT_TRUE(gridsIdentical(sharded, single));
  // accumulateUpdate integrates its cloud at once on a sharded map; reset gives an empty map on every shard
  const auto extra = sectorCloud(2, 3000, 77u);
  sharded.accumulateUpdate(extra, Eigen::Matrix<double, 3, 1>(0, 0, 0), "lidar");
  sharded.integrateUpdate();
  single.insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0, 0, 0), "lidar");
  EXPECT_TRUE(gridsIdentical(sharded, single));
  sharded.resetMap();
  EXPECT_EQ(sharded.getGrid()->activeVoxelCount(), std::uint64_t(0));
  single.resetMap();
  sharded.insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0.1, 0, 0), "lidar");
  single.insertPointCloud(extra, Eigen::Matrix<double, 3, 1>(0.1, 0, 0), "lidar");
  EXPECT_TRUE(gridsIdentical(sharded, single));
}

#ifdef VDBM_TEST_ON_MOCK
// Scenarios that only run in the CPU tier (programs linked against tests/cpp/mock_abi): host logic of the shim whose timing
// has not been exercised on the real library in this round, kept out of the GPU tier on purpose.

TEST(ShimHostLogic, ThreadedAccumulationAndPeriodicIntegrationLikeTheReferenceNodes)
{
  // addDataToAccumulate R:355-370 -> the source's accumulation thread R:1383-1411 -> the integration thread R:1416-1430
  // (accumulation_period > 0): what the ROS wrappers use. One cloud per source at a time, each waited for, so that the
  // result is comparable with direct inserts.
  Config conf              = gtestConfig(8);
  conf.accumulation_period = 0.03; // seconds
  OccupancyVDBMapping threaded(0.1), direct(0.1);
  const char* ids[2] = {"front", "rear"};
  for (OccupancyVDBMapping* m : {&threaded, &direct})
  {
    m->setMirrorMode(vdb_mapping::MirrorMode::Lazy);
    m->setConfig(conf);
    for (const char* id : ids) m->addInputSource(id, 0.0, 0);
  }
  std::uint64_t expected_updates = 0;
  for (int round = 0; round < 3; ++round)
    for (int s = 0; s < 2; ++s)
    {
      const auto cloud = sectorCloud(s, 3000, 11u * unsigned(round) + unsigned(s));
      const Eigen::Matrix<double, 3, 1> origin(0.02 * round, 0.01 * s, 0.0);
      direct.insertPointCloud(cloud, origin, ids[s]);
      vdbm_stats_t want;
      direct.deviceStats(want);
      expected_updates = want.voxel_updates;
      threaded.addDataToAccumulate(cloud, origin, ids[s]);
      bool done = false;
      for (int spin = 0; spin < 400 && !done; ++spin) // <= 8 s
      {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        vdbm_stats_t got;
        done = threaded.deviceStats(got) && got.voxel_updates == expected_updates;
      }
      EXPECT_TRUE(done);
    }
  EXPECT_TRUE(gridsIdentical(threaded, direct));
}

TEST(ShimHostLogic, SourceConcurrencyCanChangeBetweenAccumulates)
{
  // a source with update leaves on BOTH handles (accumulated on the map's handle, then on its own) is ONE update grid in
  // the reference: integrateUpdate gathers it on the map's handle first, so a voxel seen by both clouds is updated once
  const Config conf = gtestConfig(8);
  OccupancyVDBMapping switching(0.1), plain(0.1);
  plain.setSourceConcurrency(vdb_mapping::SourceConcurrency::Shared);
  switching.setSourceConcurrency(vdb_mapping::SourceConcurrency::Shared);
  for (OccupancyVDBMapping* m : {&switching, &plain})
  {
    m->setMirrorMode(vdb_mapping::MirrorMode::Lazy);
    m->setConfig(conf);
    m->addInputSource("a", 0.0, 0);
    m->addInputSource("b", 0.0, 0);
  }
  const Eigen::Matrix<double, 3, 1> origin(0, 0, 0);
  for (int round = 0; round < 3; ++round)
  {
    const auto c1 = sectorCloud(0, 3000, 5u + unsigned(round)), c2 = sectorCloud(0, 3000, 50u + unsigned(round)), c3 = sectorCloud(1, 3000, 500u + unsigned(round));
    for (OccupancyVDBMapping* m : {&switching, &plain})
    {
      m->accumulateUpdate(c1, origin, "a");                                                       // on the map's handle
      if (m == &switching) m->setSourceConcurrency(vdb_mapping::SourceConcurrency::PerSource);
      m->accumulateUpdate(c2, origin, "a");                                                       // same source, own handle
      m->accumulateUpdate(c3, origin, "b");
      m->integrateUpdate();
      if (m == &switching) m->setSourceConcurrency(vdb_mapping::SourceConcurrency::Shared);
    }
    EXPECT_TRUE(gridsIdentical(switching, plain));
  }
  vdbm_stats_t a, b;
  EXPECT_TRUE(switching.deviceStats(a) && plain.deviceStats(b));
  EXPECT_EQ(a.voxel_updates, b.voxel_updates);
  EXPECT_EQ(a.rays, b.rays);
}
#endif // VDBM_TEST_ON_MOCK

int main() { return RUN_ALL_TESTS(); }
