// backend.hpp — selects the types behind the reference's aliases (VDBMapping.hpp:81-89) and provides the few
// leaf-level operations the shim needs to move grids across the C ABI:
//   real libraries (OpenVDB >= 8.3, PCL, Eigen) when their headers exist  -> openvdb::Grid<Tree4<...>>, pcl::PointCloud, Eigen
//   otherwise (this image)                                               -> vdb_mapping/compat/compat_types.hpp
// Define VDBM_FORCE_COMPAT to force the stand-ins.
#ifndef VDB_MAPPING_DETAIL_BACKEND_HPP_INCLUDED
#define VDB_MAPPING_DETAIL_BACKEND_HPP_INCLUDED

#include <algorithm>
#include <condition_variable>
#include <cstdint>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <mutex>
#include <string>
#include <thread>
#include <type_traits>
#include <vector>

#if defined(__SSE2__) && !defined(VDBM_NO_STREAMING_COPY)
#include <emmintrin.h>
#endif

#include "vdb_mapping/detail/host_io.hpp"

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

// 2 KB leaf payload, pinned staging buffer -> leaf of the host grid. The destination is written once and not read again
// soon, so on x86-64 the copy uses non-temporal stores when the leaf buffer is 16-byte aligned: no read-for-ownership of
// the destination lines (a third of the memory traffic of the mirror update, which is bound by host memory bandwidth,
// not by PCIe: profiles/README.md). Callers fence once per batch (streamingCopyFence).
inline void copyLeafPayload(void* dst, const void* src, std::size_t bytes)
{
#if defined(__SSE2__) && !defined(VDBM_NO_STREAMING_COPY)
  if (((reinterpret_cast<std::uintptr_t>(dst) | reinterpret_cast<std::uintptr_t>(src) | bytes) & 15u) == 0)
  {
    __m128i* d       = static_cast<__m128i*>(dst);
    const __m128i* s = static_cast<const __m128i*>(src);
    for (std::size_t i = 0; i < bytes / 16; i += 4)
    {
      const __m128i a = _mm_load_si128(s + i), b = _mm_load_si128(s + i + 1), c = _mm_load_si128(s + i + 2), e = _mm_load_si128(s + i + 3);
      _mm_stream_si128(d + i, a);
      _mm_stream_si128(d + i + 1, b);
      _mm_stream_si128(d + i + 2, c);
      _mm_stream_si128(d + i + 3, e);
    }
    return;
  }
#endif
  std::memcpy(dst, src, bytes);
}
inline void streamingCopyFence()
{
#if defined(__SSE2__) && !defined(VDBM_NO_STREAMING_COPY)
  _mm_sfence();
#endif
}

// The same on a persistent pool: the mirror of the device map (VDBMapping::syncMirrorLocked) merges a chunk of leaves every
// few hundred microseconds while the next chunk crosses PCIe; starting threads per chunk would cost more than the copies.
// Static ranges, the caller works too; one job at a time (callers are serialised by the shim's device mutex anyway).
class WorkerPool
{
public:
  static WorkerPool& instance()
  {
    static WorkerPool pool;
    return pool;
  }
  template <typename F>
  void run(std::uint64_t n, F&& f)
  {
    const unsigned T = unsigned(m_threads.size()) + 1;
    if (n < 1024 || T == 1)
    {
      for (std::uint64_t i = 0; i < n; ++i) f(i);
      streamingCopyFence();
      return;
    }
    std::lock_guard<std::mutex> one_job(m_job_mutex);
    std::function<void(unsigned)> body = [&](unsigned t) {
      for (std::uint64_t i = n * t / T; i < n * (t + 1) / T; ++i) f(i);
      streamingCopyFence(); // copyLeafPayload's non-temporal stores are visible before the job counts as done
    };
    {
      std::lock_guard<std::mutex> lk(m_mutex);
      m_body    = &body;
      m_pending = T - 1;
      ++m_ticket;
    }
    m_wake.notify_all();
    body(0);
    std::unique_lock<std::mutex> lk(m_mutex);
    m_done.wait(lk, [&] { return m_pending == 0; });
    m_body = nullptr;
  }

private:
  WorkerPool()
  {
    const unsigned hw = std::max(1u, std::thread::hardware_concurrency());
    unsigned T        = std::min(16u, hw);
    if (const char* e = std::getenv("VDBM_MIRROR_THREADS")) T = std::max(1, std::atoi(e)); // experiments
    for (unsigned t = 1; t < T; ++t) m_threads.emplace_back([this, t] { loop(t); });
  }
  ~WorkerPool()
  {
    {
      std::lock_guard<std::mutex> lk(m_mutex);
      m_stop = true;
    }
    m_wake.notify_all();
    for (auto& th : m_threads) th.join();
  }
  void loop(unsigned t)
  {
    std::uint64_t seen = 0;
    for (;;)
    {
      std::function<void(unsigned)>* body = nullptr;
      {
        std::unique_lock<std::mutex> lk(m_mutex);
        m_wake.wait(lk, [&] { return m_stop || m_ticket != seen; });
        if (m_stop) return;
        seen = m_ticket;
        body = m_body;
      }
      (*body)(t);
      {
        std::lock_guard<std::mutex> lk(m_mutex);
        if (--m_pending == 0) m_done.notify_one();
      }
    }
  }
  std::vector<std::thread> m_threads;
  std::mutex m_mutex, m_job_mutex;
  std::condition_variable m_wake, m_done;
  std::function<void(unsigned)>* m_body = nullptr;
  std::uint64_t m_ticket = 0;
  unsigned m_pending     = 0;
  bool m_stop            = false;
};
} // namespace detail
} // namespace vdb_mapping

// Which types the shim hands out. The CMake package decides and exports the decision on the vdb_mapping::vdb_mapping
// target: VDBM_HAVE_OPENVDB=1 when OpenVDB, PCL and Eigen were all FOUND AND LINKED, VDBM_FORCE_COMPAT=1 otherwise - so a
// box that has the headers but no linkable package never compiles against OpenVDB and then fails to link. Only a consumer
// that uses the headers without the package (plain -I) falls back to header detection.
#if !defined(VDBM_HAVE_OPENVDB) && !defined(VDBM_FORCE_COMPAT) && defined(__has_include)
#if __has_include(<openvdb/openvdb.h>) && __has_include(<pcl/point_types.h>) && __has_include(<Eigen/Core>)
#define VDBM_HAVE_OPENVDB 1
#endif
#endif
#if defined(VDBM_FORCE_COMPAT)
#undef VDBM_HAVE_OPENVDB
#endif

#ifdef VDBM_HAVE_OPENVDB
// ------------------------------------------------------------------------------------------------------
// Real OpenVDB / PCL / Eigen. NOTE: the libraries are absent from the build image, so this branch is
// compiled and run there against headers with their public API shape (tests/cpp/stubs); it is kept small and uses only
// long-stable OpenVDB API (Tree::touchLeaf, LeafNode::buffer().data(), LeafNode::setValueMask / getValueMask,
// NodeMask::getWord).
// ------------------------------------------------------------------------------------------------------
#include <Eigen/Core>
#include <Eigen/Geometry>
#include <openvdb/io/Stream.h>
#include <openvdb/openvdb.h>
#include <openvdb/tools/Morphology.h>
#include <pcl/io/pcd_io.h>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
#include <sstream>

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
  // One chunk of vdbm_map_mirror: `table` maps the device pool index of a leaf to its host leaf, so a leaf is looked up in
  // the tree only the first time it is seen (leaf nodes stay where they are while the tree is only grown; a consumer that
  // restructures the mirror grid - prune, clear, merge - must call VDBMapping::invalidateMirrorTable()). The origin check
  // catches a table that no longer matches the tree.
  using MapLeafT = typename GridT::TreeType::LeafNodeType;
  static std::uint64_t gridEpoch(const GridT&) { return 0; }
  static void putMapLeavesIndexed(GridT& grid, std::vector<MapLeafT*>& table, std::uint64_t n, const std::uint32_t* index,
                                  const std::int32_t* origins, const float* values, const std::uint64_t* active)
  {
    std::vector<MapLeafT*> dst(n);
    for (std::uint64_t i = 0; i < n; ++i)
    {
      const openvdb::Coord o(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
      if (index[i] >= table.size()) table.resize(std::max<std::size_t>(index[i] + 1, table.size() * 2), nullptr);
      MapLeafT*& slot = table[index[i]];
      if (!slot || slot->origin() != o) slot = grid.tree().touchLeaf(o);
      dst[i] = slot;
    }
    WorkerPool::instance().run(n, [&](std::uint64_t i) {
      copyLeafPayload(dst[i]->buffer().data(), values + 512 * i, 512 * sizeof(float));
      typename MapLeafT::NodeMaskType mask;
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

  // ---- host-side grid services behind the persistence / morphology / codec members (not on the scan path) ----
  static void setVoxel(GridT& grid, const openvdb::Coord& c, const TData& v, bool on)
  {
    auto acc = grid.getAccessor();
    if (on) acc.setValueOn(c, v);
    else acc.setValueOff(c, v);
  }
  template <typename F>
  static void forEachActiveVoxel(const GridT& grid, F&& f)
  {
    for (auto it = grid.cbeginValueOn(); it; ++it) f(it.getCoord(), it.getValue());
  }
  static bool activeBBox(const GridT& grid, openvdb::CoordBBox& bb)
  {
    bb = grid.evalActiveVoxelBoundingBox();
    return !bb.empty();
  }
  static void prune(GridT& grid) { grid.pruneGrid(); }
  template <typename G>
  static void dilateActive(G& grid, int iterations) // R:1122-1129
  {
    openvdb::tools::dilateActiveValues(grid.tree(), iterations, openvdb::tools::NN_FACE_EDGE_VERTEX, openvdb::tools::EXPAND_TILES, true);
  }
  template <typename G>
  static void erodeActive(G& grid, int iterations) // R:1139-1146
  {
    openvdb::tools::erodeActiveValues(grid.tree(), iterations, openvdb::tools::NN_FACE_EDGE_VERTEX, openvdb::tools::EXPAND_TILES, true);
  }
  static bool writeGridFile(const std::string& path, const typename GridT::Ptr& grid) // R:201-208
  {
    openvdb::io::File file_handle(path);
    openvdb::GridPtrVec grids;
    grids.push_back(grid);
    file_handle.write(grids);
    file_handle.close();
    return true;
  }
  static typename GridT::Ptr readGridFile(const std::string& path) // R:265-279: the LAST grid of the file
  {
    openvdb::io::File file_handle(path);
    file_handle.open();
    openvdb::GridBase::Ptr base_grid;
    for (openvdb::io::File::NameIterator name_iter = file_handle.beginName(); name_iter != file_handle.endName(); ++name_iter)
      base_grid = file_handle.readGrid(name_iter.gridName());
    file_handle.close();
    return openvdb::gridPtrCast<GridT>(base_grid);
  }
  template <typename G>
  static std::string gridToString(const typename G::Ptr& grid) // R:1312-1316
  {
    openvdb::GridPtrVec grids;
    grids.push_back(grid);
    std::ostringstream oss(std::ios_base::binary);
    openvdb::io::Stream(oss).write(grids);
    return oss.str();
  }
  template <typename G>
  static typename G::Ptr stringToGrid(const std::string& bytes) // R:1330-1337
  {
    std::istringstream iss(bytes);
    openvdb::io::Stream strm(iss);
    openvdb::GridPtrVecPtr grids = strm.getGrids();
    return openvdb::gridPtrCast<G>(grids->front());
  }
  static bool savePCD(const std::string& path, const PointCloudT& cloud) { return pcl::io::savePCDFile(path, cloud) == 0; }
  static bool loadPCD(const std::string& path, PointCloudT& cloud) { return pcl::io::loadPCDFile<PointT>(path, cloud) != -1; }
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
  // One chunk of vdbm_map_mirror (see the OpenVDB branch): device pool index -> host leaf table; std::map nodes never move,
  // and HostGrid::structureEpoch() tells the shim when leaves were removed.
  using MapLeafT = openvdb::HostLeaf<TData>;
  static std::uint64_t gridEpoch(const GridT& grid) { return grid.structureEpoch(); }
  static void putMapLeavesIndexed(GridT& grid, std::vector<MapLeafT*>& table, std::uint64_t n, const std::uint32_t* index,
                                  const std::int32_t* origins, const float* values, const std::uint64_t* active)
  {
    std::vector<MapLeafT*> dst(n);
    for (std::uint64_t i = 0; i < n; ++i)
    {
      if (index[i] >= table.size()) table.resize(std::max<std::size_t>(index[i] + 1, table.size() * 2), nullptr);
      MapLeafT*& slot = table[index[i]];
      if (!slot) slot = &grid.touchLeaf(openvdb::Coord(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]));
      dst[i] = slot;
    }
    WorkerPool::instance().run(n, [&](std::uint64_t i) {
      copyLeafPayload(dst[i]->values, values + 512 * i, 512 * sizeof(float));
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

  // ---- host-side grid services behind the persistence / morphology / codec members (not on the scan path) ----
  static void setVoxel(GridT& grid, const openvdb::Coord& c, const TData& v, bool on)
  {
    auto acc = grid.getAccessor();
    if (on) acc.setValueOn(c, v);
    else acc.setValueOff(c, v);
  }
  template <typename G, typename F>
  static void forEachActiveVoxelOf(const G& grid, F&& f)
  {
    for (auto& kv : grid.leaves())
      for (unsigned n = 0; n < 512; ++n)
        if ((kv.second.active[n >> 6] >> (n & 63)) & 1u)
          f(openvdb::Coord(kv.first[0] + int(n >> 6), kv.first[1] + int((n >> 3) & 7), kv.first[2] + int(n & 7)), kv.second.get(n));
  }
  template <typename F>
  static void forEachActiveVoxel(const GridT& grid, F&& f) { forEachActiveVoxelOf(grid, f); }
  static bool activeBBox(const GridT& grid, openvdb::CoordBBox& bb)
  {
    bool any = false;
    forEachActiveVoxel(grid, [&](const openvdb::Coord& c, const TData&) {
      if (!any) { bb = openvdb::CoordBBox(c, c); any = true; return; }
      for (int k = 0; k < 3; ++k)
      {
        if (c[k] < bb.mn.c[k]) bb.mn.c[k] = c[k];
        if (c[k] > bb.mx.c[k]) bb.mx.c[k] = c[k];
      }
    });
    return any;
  }
  static void prune(GridT&) {} // the stand-in grid has no tiles to collapse into
  // tools::dilateActiveValues(NN_FACE_EDGE_VERTEX): every voxel with an active 26-neighbour becomes active (value untouched)
  template <typename G>
  static void dilateActive(G& grid, int iterations)
  {
    for (int it = 0; it < iterations; ++it)
    {
      std::vector<openvdb::Coord> on;
      forEachActiveVoxelOf(grid, [&](const openvdb::Coord& c, const typename G::ValueType&) { on.push_back(c); });
      auto acc = grid.getAccessor();
      for (const auto& c : on)
        for (int dx = -1; dx <= 1; ++dx)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz)
              if (dx || dy || dz) acc.setActiveState(c.offsetBy(dx, dy, dz), true);
    }
  }
  // tools::erodeActiveValues(NN_FACE_EDGE_VERTEX): an active voxel with an inactive 26-neighbour becomes inactive
  template <typename G>
  static void erodeActive(G& grid, int iterations)
  {
    for (int it = 0; it < iterations; ++it)
    {
      std::vector<openvdb::Coord> off;
      auto acc = grid.getAccessor();
      forEachActiveVoxelOf(grid, [&](const openvdb::Coord& c, const typename G::ValueType&) {
        for (int dx = -1; dx <= 1; ++dx)
          for (int dy = -1; dy <= 1; ++dy)
            for (int dz = -1; dz <= 1; ++dz)
              if ((dx || dy || dz) && !acc.isValueOn(c.offsetBy(dx, dy, dz))) { off.push_back(c); return; }
      });
      for (const auto& c : off) acc.setActiveState(c, false);
    }
  }
  // flat leaf stream of host_io.hpp (NOT the .vdb format, see there)
  static std::string gridToStringImpl(const GridT& grid)
  {
    return leafStreamWrite(grid, 0u, grid.leafCount(), [&](auto&& put) {
      for (auto& kv : grid.leaves()) { put(kv.first.c, 12); put(kv.second.values, 2048); put(kv.second.active, 64); }
    });
  }
  static std::string gridToStringImpl(const UpdateGridT& grid)
  {
    return leafStreamWrite(grid, 1u, grid.leafCount(), [&](auto&& put) {
      for (auto& kv : grid.leaves()) { put(kv.first.c, 12); put(kv.second.active, 64); put(kv.second.valmask, 64); }
    });
  }
  template <typename G>
  static std::string gridToString(const typename G::Ptr& grid) { return gridToStringImpl(*grid); }
  template <typename G>
  static typename G::Ptr stringToGrid(const std::string& bytes)
  {
    constexpr std::uint32_t want = std::is_same<typename G::ValueType, bool>::value ? 1u : 0u;
    typename G::Ptr g = G::create(typename G::ValueType());
    if (bytes.size() < 76 || bytes.compare(0, 8, "VDBMLEAF") != 0) return g;
    const char* p = bytes.data() + 8;
    std::uint32_t kind; double vs; std::uint64_t n; double bb[6];
    std::memcpy(&kind, p, 4); p += 4; std::memcpy(&vs, p, 8); p += 8; std::memcpy(&n, p, 8); p += 8; std::memcpy(bb, p, 48); p += 48;
    if (kind != want) return g;
    g->setVoxelSize(vs);
    g->insertMeta("bb_min", openvdb::Vec3d(bb[0], bb[1], bb[2]));
    g->insertMeta("bb_max", openvdb::Vec3d(bb[3], bb[4], bb[5]));
    const std::size_t rec = 12 + (want ? 128 : 2112);
    if (bytes.size() < 76 + n * rec) return g;
    for (std::uint64_t i = 0; i < n; ++i, p += rec)
    {
      std::int32_t o[3];
      std::memcpy(o, p, 12);
      auto& leaf = g->touchLeaf(openvdb::Coord(o[0], o[1], o[2]));
      readLeafPayload(leaf, p + 12);
    }
    return g;
  }
  static void readLeafPayload(openvdb::HostLeaf<float>& leaf, const char* p) { std::memcpy(leaf.values, p, 2048); std::memcpy(leaf.active, p + 2048, 64); }
  static void readLeafPayload(openvdb::HostLeaf<bool>& leaf, const char* p) { std::memcpy(leaf.active, p, 64); std::memcpy(leaf.valmask, p + 64, 64); }
  static bool writeGridFile(const std::string& path, const typename GridT::Ptr& grid)
  {
    std::ofstream f(path, std::ios::binary);
    if (!f) return false;
    const std::string bytes = gridToString<GridT>(grid);
    f.write(bytes.data(), std::streamsize(bytes.size()));
    return bool(f);
  }
  static typename GridT::Ptr readGridFile(const std::string& path)
  {
    std::ifstream f(path, std::ios::binary);
    if (!f) return nullptr;
    std::stringstream ss;
    ss << f.rdbuf();
    return stringToGrid<GridT>(ss.str());
  }
  static bool savePCD(const std::string& path, const PointCloudT& cloud) { return writePCD(path, cloud); }
  static bool loadPCD(const std::string& path, PointCloudT& cloud) { return readPCD(path, cloud); }
};

} // namespace detail
} // namespace vdb_mapping
#endif

#endif
