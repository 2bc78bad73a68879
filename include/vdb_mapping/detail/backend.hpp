// backend.hpp — selects the types behind the reference's aliases (VDBMapping.hpp:81-89) and provides the few
// leaf-level operations the shim needs to move grids across the C ABI:
//   real libraries (OpenVDB >= 8.3, PCL, Eigen) when their headers exist  -> openvdb::Grid<Tree4<...>>, pcl::PointCloud, Eigen
//   otherwise (this image)                                               -> vdb_mapping/compat/compat_types.hpp
// Define VDBM_FORCE_COMPAT to force the stand-ins.
#ifndef VDB_MAPPING_DETAIL_BACKEND_HPP_INCLUDED
#define VDB_MAPPING_DETAIL_BACKEND_HPP_INCLUDED

#include <algorithm>
#include <cstdint>
#include <cstring>
#include <thread>
#include <vector>

namespace vdb_mapping {
namespace detail {
// f(i) for i in [0, n) on a few threads (host-side payload copies of the grid mirror); serial for small n
template <typename F>
inline void parallelFor(std::uint64_t n, F&& f)
{
  const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
  const unsigned T  = n < 4096 ? 1u : std::min(8u, hw);
  if (T == 1)
  {
    for (std::uint64_t i = 0; i < n; ++i) f(i);
    return;
  }
  std::vector<std::thread> th;
  for (unsigned t = 0; t < T; ++t)
    th.emplace_back([&, t] {
      for (std::uint64_t i = n * t / T; i < n * (t + 1) / T; ++i) f(i);
    });
  for (auto& x : th) x.join();
}
} // namespace detail
} // namespace vdb_mapping

#if !defined(VDBM_FORCE_COMPAT) && defined(__has_include)
#if __has_include(<openvdb/openvdb.h>) && __has_include(<pcl/point_types.h>) && __has_include(<Eigen/Core>)
#define VDBM_HAVE_OPENVDB 1
#endif
#endif

#ifdef VDBM_HAVE_OPENVDB
// ------------------------------------------------------------------------------------------------------
// Real OpenVDB / PCL / Eigen. NOTE: this branch cannot be compiled in the build image (the libraries are
// absent there); it is kept small and uses only long-stable OpenVDB API
// (Tree::touchLeaf, LeafNode::buffer().data(), LeafNode::setValueMask / getValueMask, NodeMask::getWord).
// ------------------------------------------------------------------------------------------------------
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <openvdb/openvdb.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>

namespace vdb_mapping {
namespace detail {

template <typename TData>
struct Backend
{
  using PointT      = pcl::PointXYZ;
  using PointCloudT = pcl::PointCloud<PointT>;
  using GridT       = openvdb::Grid<typename openvdb::tree::Tree4<TData, 5, 4, 3>::Type>;
  using UpdateGridT = openvdb::Grid<openvdb::tree::Tree4<bool, 1, 4, 3>::Type>;

  static typename GridT::Ptr createMapGrid(double resolution)
  {
    typename GridT::Ptr g = GridT::create(TData());
    g->setTransform(openvdb::math::Transform::createLinearTransform(resolution));
    g->setGridClass(openvdb::GRID_LEVEL_SET);
    return g;
  }
  static UpdateGridT::Ptr createUpdateGrid(double resolution)
  {
    UpdateGridT::Ptr g = UpdateGridT::create(false);
    g->setTransform(openvdb::math::Transform::createLinearTransform(resolution));
    return g;
  }
  // write one exported map leaf (512 f32 + 8 mask words) into the host grid
  static void putMapLeaf(GridT& grid, const std::int32_t origin[3], const float* values, const std::uint64_t* active)
  {
    auto* leaf = grid.tree().touchLeaf(openvdb::Coord(origin[0], origin[1], origin[2]));
    std::memcpy(leaf->buffer().data(), values, 512 * sizeof(float));
    typename GridT::TreeType::LeafNodeType::NodeMaskType mask;
    for (int w = 0; w < 8; ++w) mask.template getWord<openvdb::Index64>(w) = active[w];
    leaf->setValueMask(mask);
  }
  // many exported map leaves at once: the tree is touched serially (node creation is not thread-safe), the 2 KB payloads
  // are then copied by a few threads (distinct leaves)
  static void putMapLeaves(GridT& grid, std::uint64_t n, const std::int32_t* origins, const float* values, const std::uint64_t* active)
  {
    using LeafT = typename GridT::TreeType::LeafNodeType;
    std::vector<LeafT*> dst(n);
    for (std::uint64_t i = 0; i < n; ++i) dst[i] = grid.tree().touchLeaf(openvdb::Coord(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]));
    parallelFor(n, [&](std::uint64_t i) {
      std::memcpy(dst[i]->buffer().data(), values + 512 * i, 512 * sizeof(float));
      typename LeafT::NodeMaskType mask;
      for (int w = 0; w < 8; ++w) mask.template getWord<openvdb::Index64>(w) = active[8 * i + w];
      dst[i]->setValueMask(mask);
    });
  }
  static void putUpdateLeaf(UpdateGridT& grid, const std::int32_t origin[3], const std::uint64_t* active, const std::uint64_t* value)
  {
    auto* leaf = grid.tree().touchLeaf(openvdb::Coord(origin[0], origin[1], origin[2]));
    for (unsigned n = 0; n < 512; ++n)
    {
      const bool a = (active[n >> 6] >> (n & 63)) & 1u, v = (value[n >> 6] >> (n & 63)) & 1u;
      if (a) leaf->setValueOn(n, v || leaf->getValue(n));
      else if (v) leaf->setValueOnly(n, true);
    }
  }
  // enumerate the leaves of an update grid: f(origin[3], active[8], value[8])
  template <typename F>
  static void forEachUpdateLeaf(const UpdateGridT& grid, F&& f)
  {
    for (auto it = grid.tree().cbeginLeaf(); it; ++it)
    {
      const openvdb::Coord o = it->origin();
      std::int32_t origin[3] = {o.x(), o.y(), o.z()};
      std::uint64_t active[8], value[8];
      for (int w = 0; w < 8; ++w)
      {
        active[w] = it->getValueMask().template getWord<openvdb::Index64>(w);
        value[w]  = it->buffer().getWord(w); // LeafBuffer<bool,3> stores the values as a NodeMask
      }
      f(origin, active, value);
    }
  }
  // enumerate the leaves of a float grid: f(origin[3], values[512], active[8])
  template <typename F>
  static void forEachMapLeaf(const GridT& grid, F&& f)
  {
    for (auto it = grid.tree().cbeginLeaf(); it; ++it)
    {
      const openvdb::Coord o = it->origin();
      std::int32_t origin[3] = {o.x(), o.y(), o.z()};
      std::uint64_t active[8];
      for (int w = 0; w < 8; ++w) active[w] = it->getValueMask().template getWord<openvdb::Index64>(w);
      f(origin, it->buffer().data(), active);
    }
  }
  template <typename G>
  static void getSectionMeta(const G& g, std::int32_t mn[3], std::int32_t mx[3])
  {
    const openvdb::Coord a = openvdb::Coord::floor(g.template metaValue<openvdb::Vec3d>("bb_min"));
    const openvdb::Coord b = openvdb::Coord::floor(g.template metaValue<openvdb::Vec3d>("bb_max"));
    for (int k = 0; k < 3; ++k) { mn[k] = a[k]; mx[k] = b[k]; }
  }
  static void setSectionMeta(UpdateGridT& g, const std::int32_t mn[3], const std::int32_t mx[3])
  {
    g.insertMeta("bb_min", openvdb::Vec3DMetadata(openvdb::Vec3d(mn[0], mn[1], mn[2])));
    g.insertMeta("bb_max", openvdb::Vec3DMetadata(openvdb::Vec3d(mx[0], mx[1], mx[2])));
  }
  static void setSectionMeta(GridT& g, const std::int32_t mn[3], const std::int32_t mx[3])
  {
    g.insertMeta("bb_min", openvdb::Vec3DMetadata(openvdb::Vec3d(mn[0], mn[1], mn[2])));
    g.insertMeta("bb_max", openvdb::Vec3DMetadata(openvdb::Vec3d(mx[0], mx[1], mx[2])));
  }
};

} // namespace detail
} // namespace vdb_mapping

#else
// ------------------------------------------------------------------------------------------------------
// Stand-in types (this image): same spelling as the real ones through namespace aliases.
// ------------------------------------------------------------------------------------------------------
#include "vdb_mapping/compat/compat_types.hpp"

namespace pcl     = vdbm_compat::pcl;
namespace Eigen   = vdbm_compat::Eigen;
namespace openvdb = vdbm_compat::openvdb;

namespace vdb_mapping {
namespace detail {

template <typename TData>
struct Backend
{
  using PointT      = pcl::PointXYZ;
  using PointCloudT = pcl::PointCloud<PointT>;
  using GridT       = openvdb::HostGrid<TData>;
  using UpdateGridT = openvdb::HostGrid<bool>;

  static typename GridT::Ptr createMapGrid(double resolution)
  {
    typename GridT::Ptr g = GridT::create(TData());
    g->setVoxelSize(resolution);
    return g;
  }
  static typename UpdateGridT::Ptr createUpdateGrid(double resolution)
  {
    typename UpdateGridT::Ptr g = UpdateGridT::create(false);
    g->setVoxelSize(resolution);
    return g;
  }
  static void putMapLeaf(GridT& grid, const std::int32_t origin[3], const float* values, const std::uint64_t* active)
  {
    auto& leaf = grid.touchLeaf(openvdb::Coord(origin[0], origin[1], origin[2]));
    std::memcpy(leaf.values, values, 512 * sizeof(float));
    std::memcpy(leaf.active, active, 8 * sizeof(std::uint64_t));
  }
  static void putMapLeaves(GridT& grid, std::uint64_t n, const std::int32_t* origins, const float* values, const std::uint64_t* active)
  {
    std::vector<openvdb::HostLeaf<float>*> dst;
    grid.touchLeaves(n, origins, dst);
    parallelFor(n, [&](std::uint64_t i) {
      std::memcpy(dst[i]->values, values + 512 * i, 512 * sizeof(float));
      std::memcpy(dst[i]->active, active + 8 * i, 8 * sizeof(std::uint64_t));
    });
  }
  static void putUpdateLeaf(UpdateGridT& grid, const std::int32_t origin[3], const std::uint64_t* active, const std::uint64_t* value)
  {
    auto& leaf = grid.touchLeaf(openvdb::Coord(origin[0], origin[1], origin[2]));
    for (int w = 0; w < 8; ++w)
    {
      leaf.active[w] |= active[w];
      leaf.valmask[w] |= value[w];
    }
  }
  template <typename F>
  static void forEachUpdateLeaf(const UpdateGridT& grid, F&& f)
  {
    for (auto& kv : grid.leaves())
    {
      std::int32_t origin[3] = {kv.first[0], kv.first[1], kv.first[2]};
      f(origin, kv.second.active, kv.second.valmask);
    }
  }
  template <typename F>
  static void forEachMapLeaf(const GridT& grid, F&& f)
  {
    for (auto& kv : grid.leaves())
    {
      std::int32_t origin[3] = {kv.first[0], kv.first[1], kv.first[2]};
      f(origin, kv.second.values, kv.second.active);
    }
  }
  template <typename G>
  static void getSectionMeta(const G& g, std::int32_t mn[3], std::int32_t mx[3])
  {
    const openvdb::Coord a = openvdb::Coord::floor(g.metaValue("bb_min")), b = openvdb::Coord::floor(g.metaValue("bb_max"));
    for (int k = 0; k < 3; ++k) { mn[k] = a[k]; mx[k] = b[k]; }
  }
  template <typename G>
  static void setSectionMeta(G& g, const std::int32_t mn[3], const std::int32_t mx[3])
  {
    g.insertMeta("bb_min", openvdb::Vec3d(mn[0], mn[1], mn[2]));
    g.insertMeta("bb_max", openvdb::Vec3d(mx[0], mx[1], mx[2]));
  }
};

} // namespace detail
} // namespace vdb_mapping
#endif

#endif
