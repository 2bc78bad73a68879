// VDBMapping.hpp — B200 drop-in for the scan-integration members of vdb_mapping::VDBMapping<TData, TConfig>.
//
// Same namespace, class template, aliases and member signatures as the reference header
// (/root/reference/include/vdb_mapping/VDBMapping.hpp, cited below as R:<line>), so that code written against the
// reference — its ROS/ROS2 wrappers, its tests/mapping.cpp — compiles unchanged. Nothing of the algorithm lives
// here: every grid operation is a call into the C ABI of libvdbm_b200.so (include/vdbm_b200.h), which runs the
// CUDA kernels. This header only
//   * keeps the host-side plumbing of the reference: input sources, the accumulation / integration worker threads,
//     the shared map mutex with writer priority (R:133,1343-1430,1534-1540);
//   * mirrors the device-resident map into a host grid object (`getGrid()`, R:799). Mirror modes:
//       Eager (default)  every integrate copies the leaves it modified back, so a `GridT::Accessor` obtained
//                        BEFORE an insert sees the new values afterwards (what tests/mapping.cpp:13-29 relies on);
//       Lazy             nothing is copied until getGrid() is called (throughput mode: the scan-integration
//                        path then never leaves the GPU).
//   * converts between host grids and the leaf records of the ABI where the reference passes grids around
//     (raycastPointCloud's accessor, updateMap's argument and result, getMapSection*).
//
// Scope: insertPointCloud, accumulateUpdate, addDataToAccumulate, integrateUpdate,
// raycastPointCloud, updateMap, getGrid, getMapSection*, applyMapSection*, createIndexBoundingBox, addInputSource,
// setConfig, resetMap, getMapMutex, worldToIndex, addPointsToGrid, removePointsFromGrid, addArtificialAreas,
// restoreMapIntegrity, plus createUpdate / applyUpdate (remote-mapping deltas; see vdbm_b200.h for the level semantics).
// The remaining public members of the reference are host code on the mirror grid: saveMap /
// saveMapToPCD / loadMap / loadMapFromPCD (a loaded map is imported into the device map with vdbm_map_import),
// morphological*Map, the zstd codec and gridToByteArray / byteArrayToGrid, createWorldBoundingBox, getMapSection<T>,
// addArtificialPolygon / addArtificialWall, castRayIntoGrid (vdbm_cast_index_rays), raytrace and fast_mode
// (vdbm_raytrace / the fast raycast; OpenVDB's VolumeRayIntersector is restated).
// The device arithmetic implements the OccupancyVDBMapping node operations (TData = float); the protected virtual
// update*Node hooks of the reference cannot be honoured on the device and are therefore not part of this class.
#ifndef VDB_MAPPING_VDB_MAPPING_H_INCLUDED
#define VDB_MAPPING_VDB_MAPPING_H_INCLUDED

#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <ctime>
#include <iomanip>
#include <iostream>
#include <map>
#include <memory>
#include <mutex>
#include <optional>
#include <shared_mutex>
#include <sstream>
#include <string>
#include <thread>
#include <type_traits>
#include <utility>
#include <vector>

#include "vdb_mapping/detail/backend.hpp"
#include "vdbm_b200.h"

namespace vdb_mapping {

/*! Configuration parameters, field for field as R:64-71. */
struct BaseConfig
{
  double max_range;
  bool fast_mode;
  double accumulation_period;
  std::string map_directory_path;
};

/*! How the host grid returned by getGrid() follows the device-resident map. */
enum class MirrorMode
{
  Eager,
  Lazy
};

/*! Where the rays of an input source are cast (not part of the reference API). The reference's one parallel axis is a
 *  thread per input source, each raycasting into its own update grid under a shared lock (R:316-346, 1383-1411). A C ABI
 *  handle is thread-compatible, not thread-safe, so sources that share the map's handle take turns on the device. With
 *  PerSource every source owns a second, raycast-only handle (own stream, staging buffers and counters, a one-leaf map):
 *  accumulateUpdate() of different sources then runs concurrently - host side and on the GPU, where the DDA launches of small
 *  clouds co-reside - and integrateUpdate() moves the accumulated update leaves device-to-device into the map's handle
 *  (vdbm_update_partition -> vdbm_update_import_device) before updateMap. Auto = PerSource as soon as more than one source
 *  is registered (costs one update grid, 0.5 GB at the default capacity, per source). fast_mode always uses the map's
 *  handle: castRayIntoGridFast reads the map. */
enum class SourceConcurrency
{
  Auto,
  Shared,
  PerSource
};

template <typename TData, typename TConfig = BaseConfig>
class VDBMapping
{
  static_assert(std::is_same<TData, float>::value, "the B200 path implements the float (occupancy log-odds) map");
  using BackendT = detail::Backend<TData>;

public:
  using PointT      = typename BackendT::PointT;
  using PointCloudT = typename BackendT::PointCloudT;
  using GridT       = typename BackendT::GridT;
  using UpdateGridT = typename BackendT::UpdateGridT;

  /*! Host-side record of one input source (R:91-103); its update grid lives on the device. */
  struct InputSource
  {
    std::string source_id;
    double max_range;
    std::mutex update_grid_mutex;
    std::mutex input_data_mutex;
    std::optional<std::pair<typename PointCloudT::ConstPtr, Eigen::Matrix<double, 3, 1> > > input_data;
    std::chrono::milliseconds max_input_period;
    std::condition_variable data_available_cv;
    // SourceConcurrency::PerSource (all three guarded by update_grid_mutex or the exclusive map lock)
    bool shared_holds_data    = false;   // update leaves accumulated on the MAP's handle and not yet integrated
    vdbm_map* raycaster       = nullptr; // raycast-only device handle of this source
    double raycaster_range    = -1.0;    // config range it was last configured with
    bool raycaster_holds_data = false;   // update leaves accumulated there and not yet moved to the map's handle
    bool last_cast_on_raycaster = false; // where the source's most recent accumulate ran (its ray end voxels stay there)
  };

  VDBMapping()                  = delete;
  VDBMapping(const VDBMapping&) = delete;
  VDBMapping& operator=(const VDBMapping&) = delete;

  /*! R:115-134. The device map is created here; the integration thread starts like in the reference. */
  explicit VDBMapping(const double resolution)
    : m_resolution(resolution)
    , m_config_set(false)
  {
    openvdb::initialize(); // R:122 (the grid types of this path are OpenVDB's own FloatGrid / a bool Tree4<1,4,3>)
    m_map_mutex = std::make_shared<std::shared_mutex>();
    vdbm_params p{};
    p.resolution            = resolution;
    p.device                = -1;
    p.replicate_probe_quirk = 1;
    const int rc            = vdbm_create(&p, &m_device_map);
    if (rc != VDBM_OK)
    {
      // no CPU fallback exists: the reference semantics cannot be provided without the device
      std::cerr << "vdb_mapping (B200): could not create the device map (status " << rc << "); a CUDA device is required."
                << std::endl;
      m_device_map = nullptr;
    }
    m_vdb_grid           = createVDBMap(m_resolution);
    m_integration_thread = std::thread(&VDBMapping::integrationThread, this);
  }

  /*! R:139-154 */
  virtual ~VDBMapping()
  {
    m_thread_stop_signal = true;
    for (auto& [source_id, worker_thread] : m_worker_threads)
    {
      if (auto src = findSource(source_id)) src->data_available_cv.notify_all();
      if (worker_thread.joinable()) worker_thread.join();
    }
    if (m_integration_thread.joinable()) m_integration_thread.join();
    for (auto& kv : sourcesSnapshot())
      if (kv.second->raycaster) vdbm_destroy(kv.second->raycaster);
    if (m_device_map) vdbm_destroy(m_device_map);
    if (m_device_group) vdbm_group_destroy(m_device_group);
  }

  /*! R:163-169 */
  typename GridT::Ptr createVDBMap(double resolution) { return BackendT::createMapGrid(resolution); }

  /*! R:174-186 */
  void resetMap()
  {
    std::unique_lock map_lock(*m_map_mutex);
    if (m_device_group)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      reportGroup(vdbm_group_reset(m_device_group));
    }
    if (m_device_map)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      vdbm_reset(m_device_map);
      for (auto& kv : sourcesSnapshot()) // "new empty update grids" R:181-185, wherever they live
      {
        kv.second->shared_holds_data = false;
        if (kv.second->raycaster)
        {
          vdbm_reset(kv.second->raycaster);
          kv.second->raycaster_holds_data = false;
        }
      }
    }
    m_vdb_grid->clear();
    m_vdb_grid = createVDBMap(m_resolution);
    map_lock.unlock();
  }

  /*! Mirror policy of getGrid(); see the header comment. Not part of the reference API. */
  void setMirrorMode(MirrorMode mode) { m_mirror_mode = mode; }
  /*! See SourceConcurrency. Not part of the reference API. */
  void setSourceConcurrency(SourceConcurrency mode) { m_source_concurrency = mode; }

  /*! Not part of the reference API: shard the map over several GPUs of one NVLink box, driven from this one process
   *  (vdbm_group_*): every device raycasts the rays of its azimuth sector and owns the map leaves of a sector,
   *  update leaves travel between the devices over NVLink, the union of the shards is bit for bit the map one GPU would hold.
   *  Call it right after construction (before setConfig / addInputSource). What a sharded map offers: setConfig,
   *  addInputSource, insertPointCloud (accumulateUpdate integrates its cloud at once: the group works scan by scan),
   *  getGrid, getMapSection*, resetMap, saveMap*, deviceStats. Everything that needs the whole map on one device
   *  (raycastPointCloud / updateMap with caller grids, createUpdate / applyUpdate, applyMapSection*, point edits, artificial
   *  areas, loadMap*, raytrace, fast_mode) reports that it is unavailable and does nothing.
   *  A device may be listed more than once (several shards on one GPU: how the test tier runs it on a single GPU). */
  bool setDevices(const std::vector<int>& devices)
  {
    std::unique_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    if (m_config_set || sourceCount() != 0 || m_device_group)
    {
      std::cerr << "vdb_mapping (B200): setDevices must be called once, before setConfig and addInputSource" << std::endl;
      return false;
    }
    if (devices.empty()) return false;
    vdbm_params p{};
    p.resolution            = m_resolution;
    p.device                = -1;
    p.replicate_probe_quirk = 1;
    std::vector<std::int32_t> ids(devices.begin(), devices.end());
    vdbm_group* group = nullptr;
    const int rc      = vdbm_group_create(&p, static_cast<std::int32_t>(ids.size()), ids.data(), 0, &group);
    if (rc != VDBM_OK || !group)
    {
      std::cerr << "vdb_mapping (B200): could not create the device group (status " << rc << "); the map stays on one device" << std::endl;
      return false;
    }
    if (m_device_map) vdbm_destroy(m_device_map);
    m_device_map   = nullptr; // every single-handle member is a no-op from here on; the sharded ones use the group
    m_device_group = group;
    m_shard_tables.assign(ids.size(), {});
    m_shard_generations.assign(ids.size(), ~std::uint64_t(0));
    return true;
  }

  /*! R:316-346. Unknown source: message + return; source range <= 0: nothing is raycast. */
  void accumulateUpdate(const typename PointCloudT::ConstPtr& cloud,
                        const Eigen::Matrix<double, 3, 1>& origin,
                        const std::string source_id)
  {
    const std::shared_ptr<InputSource> source = findSource(source_id);
    if (!source)
    {
      std::cout << "Tried to accumulate update for " << source_id << ". Source not available" << std::endl;
      return;
    }
    if (m_device_group)
    {
      // the group works scan by scan (raycast, exchange and updateMap are one call): this cloud is integrated by itself
      insertSharded(cloud, origin, source_id);
      return;
    }
    std::shared_lock map_lock(*m_map_mutex);
    std::unique_lock update_grid_lock(source->update_grid_mutex);
    if (!m_device_map || !cloud) return;
    const double o[3] = {origin.x(), origin.y(), origin.z()};
    if (vdbm_map* raycaster = sourceRaycaster(source_id, *source))
    {
      // this source's own handle: no other thread touches it (update_grid_mutex), nothing is shared with the map's handle
      const int rc = vdbm_accumulate(raycaster, source_id.c_str(), cloud->points.data(), cloud->points.size(), sizeof(PointT), o);
      source->raycaster_holds_data   = true;
      source->last_cast_on_raycaster = true;
      reportOn(raycaster, rc);
      return;
    }
    // The C ABI handle is thread-compatible, not thread-safe: sources that share the map's handle take turns on the device
    // (they still overlap their host work).
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    // the ABI wants the pcl::PointXYZ records as they lie in the cloud (16-byte stride)
    const int rc = vdbm_accumulate(m_device_map, source_id.c_str(), cloud->points.data(), cloud->points.size(), sizeof(PointT), o);
    source->shared_holds_data      = true;
    source->last_cast_on_raycaster = false;
    report(rc);
  }

  /*! R:355-370: latest-wins hand-off to the source's accumulation thread. */
  void addDataToAccumulate(const typename PointCloudT::ConstPtr& cloud,
                           const Eigen::Matrix<double, 3, 1>& origin,
                           const std::string source_id)
  {
    const std::shared_ptr<InputSource> source = findSource(source_id);
    if (!source)
    {
      std::cout << "Tried to add data for accumulation of " << source_id << ". Source not available" << std::endl;
      return;
    }
    std::unique_lock lock(source->input_data_mutex);
    source->input_data = std::make_pair(cloud, origin);
    source->data_available_cv.notify_all();
  }

  /*! R:375-387: exclusive map lock (writer priority flag), updateMap for every source in key order. */
  void integrateUpdate()
  {
    m_map_mutex_requested = true;
    std::unique_lock map_lock(*m_map_mutex);
    m_map_mutex_requested = false;
    if (!m_device_map) return;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    bool on_raycasters = false; // exclusive map lock: no accumulation is running, the sources' flags are stable
    const auto sources = sourcesSnapshot(); // key order, R:380
    for (auto& kv : sources) on_raycasters = on_raycasters || kv.second->raycaster_holds_data;
    if (!on_raycasters) report(vdbm_integrate(m_device_map, 0));
    else
    {
      // updateMap per source in key order (R:380), each grid read where it lies; a source with leaves on BOTH handles (the
      // mode changed between two accumulates) is one update grid in the reference: gathered on the map's handle first
      for (auto& kv : sources)
      {
        InputSource& src = *kv.second;
        if (src.raycaster_holds_data && src.shared_holds_data) collectRaycasterLocked(kv.first, src);
        if (src.raycaster_holds_data)
        {
          src.raycaster_holds_data = false;
          report(vdbm_integrate_from(m_device_map, src.raycaster, kv.first.c_str(), 0));
        }
        else if (src.shared_holds_data) report(vdbm_integrate_from(m_device_map, m_device_map, kv.first.c_str(), 0));
      }
    }
    for (auto& kv : sources) kv.second->shared_holds_data = false;
    if (m_mirror_mode == MirrorMode::Eager) syncMirrorLocked();
    else m_mirror_stale = true;
  }

  /*! R:399-406 */
  bool insertPointCloud(const typename PointCloudT::ConstPtr& cloud,
                        const Eigen::Matrix<double, 3, 1>& origin,
                        const std::string source_id)
  {
    if (m_device_group)
    {
      if (!findSource(source_id))
      {
        std::cout << "Tried to accumulate update for " << source_id << ". Source not available" << std::endl;
        return true;
      }
      insertSharded(cloud, origin, source_id);
      return true;
    }
    if (m_mirror_mode == MirrorMode::Lazy && m_device_map && cloud)
    {
      // throughput mode: the scan is queued as one pipeline stage (vdbm_insert_async: upload overlapped with the previous
      // scan, no host round trip between raycast and updateMap); getGrid() / any other member finishes it
      const std::shared_ptr<InputSource> source = findSource(source_id);
      if (source && !source->raycaster_holds_data && !wantsRaycaster())
      {
        m_map_mutex_requested = true;
        std::unique_lock map_lock(*m_map_mutex);
        m_map_mutex_requested = false;
        std::unique_lock update_grid_lock(source->update_grid_mutex);
        std::lock_guard<std::mutex> device_lock(m_device_mutex);
        const double o[3] = {origin.x(), origin.y(), origin.z()};
        report(vdbm_insert_async(m_device_map, source_id.c_str(), cloud->points.data(), cloud->points.size(), sizeof(PointT), o, 0));
        source->last_cast_on_raycaster = false;
        m_mirror_stale = true;
        return true;
      }
    }
    accumulateUpdate(cloud, origin, source_id);
    integrateUpdate();
    return true;
  }

  /*! R:466-539 (fast_mode follows Config::fast_mode on the device; the optional host intersector argument has no meaning
   *  here and is not declared). The rays are cast on the device into a scratch
   *  source; the resulting update leaves are then merged into the grid behind `update_grid_acc`. */
  bool raycastPointCloud(const typename PointCloudT::ConstPtr& cloud,
                         const Eigen::Matrix<double, 3, 1>& origin,
                         const double raycast_range,
                         typename UpdateGridT::Accessor& update_grid_acc)
  {
    if (shardedUnavailable("raycastPointCloud")) return false;
    if (!m_config_set)
    {
      std::cerr << "Map not properly configured. Did you call setConfig method?" << std::endl;
      return false;
    }
    if (!m_device_map || !cloud) return false;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    ensureScratchSource();
    ScratchClear clear_on_exit{m_device_map}; // whatever happens below, nothing stays behind for the next integrateUpdate
    const double o[3] = {origin.x(), origin.y(), origin.z()};
    const int rc = report(vdbm_raycast(m_device_map, kScratchSource, cloud->points.data(), cloud->points.size(), sizeof(PointT), o, raycast_range));
    if (rc != VDBM_OK && rc != VDBM_ERR_COORD_RANGE) return false; // out-of-range points are dropped and reported, the rest counts
    vdbm_leafset* ls = nullptr;
    if (report(vdbm_update_export(m_device_map, kScratchSource, &ls)) != VDBM_OK) return false;
    mergeIntoAccessor(ls, update_grid_acc);
    vdbm_leafset_free(ls);
    return true;
  }

  /*! R:612-631 */
  openvdb::Coord worldToIndex(const openvdb::Vec3d& world_coordinate) const
  {
    double c[3] = {world_coordinate.x(), world_coordinate.y(), world_coordinate.z()};
    openvdb::Int32 idx[3];
    const double inv = 1.0 / m_resolution;
    for (int a = 0; a < 3; ++a)
    {
      if (std::fmod(c[a], m_resolution)) c[a] = c[a] + (m_resolution / 2.0);
      idx[a] = openvdb::Int32(std::floor(c[a] * inv));
    }
    return openvdb::Coord(idx[0], idx[1], idx[2]);
  }

  /*! R:731-792: applies a caller-provided update grid and returns the change grid. */
  typename UpdateGridT::Ptr updateMap(const typename UpdateGridT::Ptr& temp_grid)
  {
    typename UpdateGridT::Ptr change = BackendT::createUpdateGrid(m_resolution);
    if (shardedUnavailable("updateMap")) return change;
    if (!m_device_map || !temp_grid || temp_grid->empty()) return change;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    ensureScratchSource();
    std::vector<std::int32_t> origins;
    std::vector<std::uint64_t> active, value;
    BackendT::forEachUpdateLeaf(*temp_grid, [&](const std::int32_t o[3], const std::uint64_t* a, const std::uint64_t* v) {
      origins.insert(origins.end(), o, o + 3);
      active.insert(active.end(), a, a + 8);
      value.insert(value.end(), v, v + 8);
    });
    ScratchClear clear_on_exit{m_device_map};
    if (report(vdbm_update_import(m_device_map, kScratchSource, origins.size() / 3, origins.data(), active.data(), value.data())) != VDBM_OK)
      return change;
    vdbm_leafset* ls = nullptr;
    if (report(vdbm_update_map(m_device_map, kScratchSource, &ls)) != VDBM_OK) return change;
    const std::uint64_t n = vdbm_leafset_size(ls);
    for (std::uint64_t i = 0; i < n; ++i)
      BackendT::putUpdateLeaf(*change, vdbm_leafset_origins(ls) + 3 * i, vdbm_leafset_active(ls) + 8 * i, vdbm_leafset_valmask(ls) + 8 * i);
    vdbm_leafset_free(ls);
    if (m_mirror_mode == MirrorMode::Eager) syncMirrorLocked();
    else m_mirror_stale = true;
    return change;
  }

  /*! R:799. The returned grid object is the live host mirror of the device map. */
  typename GridT::Ptr getGrid() const
  {
    if (m_mirror_stale)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      const_cast<VDBMapping*>(this)->syncMirrorLocked();
    }
    return m_vdb_grid;
  }

  /*! R:857-871 with R:810-847: index bounding box of a box given in a reference frame. The 8 corners are stored as
   *  float (pcl::PointXYZ), transformed with the 4x4 double matrix, min/max taken, then worldToIndex + floor. */
  openvdb::CoordBBox createIndexBoundingBox(const Eigen::Matrix<double, 3, 1>& min_boundary,
                                            const Eigen::Matrix<double, 3, 1>& max_boundary,
                                            const Eigen::Matrix<double, 4, 4>& map_to_reference_tf) const
  {
    float lo[3], hi[3];
    transformedCorners(min_boundary, max_boundary, map_to_reference_tf, lo, hi);
    const double inv = 1.0 / m_resolution;
    openvdb::Vec3d mn(lo[0] * inv, lo[1] * inv, lo[2] * inv), mx(hi[0] * inv, hi[1] * inv, hi[2] * inv);
    return openvdb::CoordBBox(openvdb::Coord::floor(mn), openvdb::Coord::floor(mx));
  }

  /*! R:883-891 */
  typename UpdateGridT::Ptr getMapSectionUpdateGrid(const Eigen::Matrix<double, 3, 1>& min_boundary,
                                                    const Eigen::Matrix<double, 3, 1>& max_boundary,
                                                    const Eigen::Matrix<double, 4, 4>& map_to_reference_tf,
                                                    const bool full_grid = false) const
  {
    typename UpdateGridT::Ptr out = BackendT::createUpdateGrid(m_resolution);
    const openvdb::CoordBBox bb   = createIndexBoundingBox(min_boundary, max_boundary, map_to_reference_tf);
    const std::int32_t mn[3] = {bb.min().x(), bb.min().y(), bb.min().z()}, mx[3] = {bb.max().x(), bb.max().y(), bb.max().z()};
    std::shared_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    forEachMapHandle([&](vdbm_map* handle, std::size_t) { // the shards of a sharded map hold disjoint leaf sets
      vdbm_leafset* ls = nullptr;
      if (vdbm_section(handle, mn, mx, full_grid ? 1 : 0, 0, &ls) != VDBM_OK) return;
      const std::uint64_t n = vdbm_leafset_size(ls);
      for (std::uint64_t i = 0; i < n; ++i)
        BackendT::putUpdateLeaf(*out, vdbm_leafset_origins(ls) + 3 * i, vdbm_leafset_active(ls) + 8 * i, vdbm_leafset_valmask(ls) + 8 * i);
      vdbm_leafset_free(ls);
    });
    BackendT::setSectionMeta(*out, mn, mx); // R:955-958
    return out;
  }

  /*! R:902-909 */
  typename GridT::Ptr getMapSectionGrid(const Eigen::Matrix<double, 3, 1>& min_boundary,
                                        const Eigen::Matrix<double, 3, 1>& max_boundary,
                                        const Eigen::Matrix<double, 4, 4>& map_to_reference_tf,
                                        const bool full_grid = false) const
  {
    typename GridT::Ptr out     = BackendT::createMapGrid(m_resolution);
    const openvdb::CoordBBox bb = createIndexBoundingBox(min_boundary, max_boundary, map_to_reference_tf);
    const std::int32_t mn[3] = {bb.min().x(), bb.min().y(), bb.min().z()}, mx[3] = {bb.max().x(), bb.max().y(), bb.max().z()};
    std::shared_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    forEachMapHandle([&](vdbm_map* handle, std::size_t) {
      vdbm_leafset* ls = nullptr;
      if (vdbm_section(handle, mn, mx, full_grid ? 1 : 0, 1, &ls) != VDBM_OK) return;
      const std::uint64_t n = vdbm_leafset_size(ls);
      for (std::uint64_t i = 0; i < n; ++i)
        BackendT::putMapLeaf(*out, vdbm_leafset_origins(ls) + 3 * i, vdbm_leafset_values(ls) + 512 * i, vdbm_leafset_active(ls) + 8 * i);
      vdbm_leafset_free(ls);
    });
    BackendT::setSectionMeta(*out, mn, mx);
    return out;
  }

  /*! R:1022-1047. The optional morphological smoothing (OpenVDB tools::dilate/erodeActiveValues, R:1094-1147) is not
   *  part of this build: smooth_map = true is reported and ignored. */
  void applyMapSectionGrid(const typename GridT::Ptr section, bool smooth_map = false, int smoothing_iterations = 2)
  {
    if (shardedUnavailable("applyMapSectionGrid")) return;
    (void)smoothing_iterations;
    if (smooth_map) std::cerr << "vdb_mapping (B200): morphological smoothing is not supported; section applied unsmoothed" << std::endl;
    if (!m_device_map || !section) return;
    std::vector<std::int32_t> origins;
    std::vector<std::uint64_t> active;
    std::vector<float> values;
    BackendT::forEachMapLeaf(*section, [&](const std::int32_t o[3], const float* v, const std::uint64_t* a) {
      origins.insert(origins.end(), o, o + 3);
      values.insert(values.end(), v, v + 512);
      active.insert(active.end(), a, a + 8);
    });
    std::unique_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_section_apply_grid(m_device_map, origins.size() / 3, origins.data(), active.data(), values.data(), 1));
    if (m_mirror_mode == MirrorMode::Eager) syncMirrorLocked();
    else m_mirror_stale = true;
  }

  /*! R:1058-1085 (same remark on smoothing). The box comes from the section's bb_min / bb_max metadata. */
  void applyMapSectionUpdateGrid(const typename UpdateGridT::Ptr section, bool smooth_map = false, int smoothing_iterations = 2)
  {
    if (shardedUnavailable("applyMapSectionUpdateGrid")) return;
    (void)smoothing_iterations;
    if (smooth_map) std::cerr << "vdb_mapping (B200): morphological smoothing is not supported; section applied unsmoothed" << std::endl;
    if (!m_device_map || !section) return;
    std::int32_t mn[3], mx[3];
    BackendT::getSectionMeta(*section, mn, mx);
    std::vector<std::int32_t> origins;
    std::vector<std::uint64_t> active;
    BackendT::forEachUpdateLeaf(*section, [&](const std::int32_t o[3], const std::uint64_t* a, const std::uint64_t*) {
      origins.insert(origins.end(), o, o + 3);
      active.insert(active.end(), a, a + 8);
    });
    std::unique_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_section_apply_update(m_device_map, mn, mx, origins.size() / 3, origins.data(), active.data()));
    if (m_mirror_mode == MirrorMode::Eager) syncMirrorLocked();
    else m_mirror_stale = true;
  }

  /*! R:413-429 */
  bool removePointsFromGrid(const typename PointCloudT::ConstPtr& cloud) { return setPoints(cloud, 0); }
  /*! R:431-447 */
  bool addPointsToGrid(const typename PointCloudT::ConstPtr& cloud) { return setPoints(cloud, 1); }

  /*! R:1152-1166 */
  void restoreMapIntegrity()
  {
    if (shardedUnavailable("restoreMapIntegrity")) return;
    if (!m_device_map) return;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_map_integrity_restore(m_device_map));
    m_artificial_areas_present = false;
    afterMapWriteLocked();
  }

  /*! R:1175-1189 (the walls reach the map with the next updateMap, exactly like R:785-789) */
  void addArtificialAreas(const std::vector<std::vector<Eigen::Matrix<double, 4, 1> > >& artificial_areas,
                          const double negative_height,
                          const double positive_height)
  {
    if (!m_device_map) return;
    std::vector<std::uint32_t> counts;
    if (shardedUnavailable("addArtificialAreas")) return;
    std::vector<double> xyz;
    for (const auto& area : artificial_areas)
    {
      counts.push_back(static_cast<std::uint32_t>(area.size()));
      for (const auto& p : area) xyz.insert(xyz.end(), {p[0], p[1], p[2]});
    }
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_artificial_areas_add(m_device_map, counts.size(), counts.data(), xyz.data(), negative_height, positive_height));
    m_artificial_areas_present = true;
    afterMapWriteLocked(); // the implied restoreMapIntegrity may have changed flags
  }

  /*! Remote-mapping delta of one source (north_star "createUpdate"; levels in vdbm_b200.h): 0 raw update grid, 1 change
   *  grid of the last integrate, 2 reduced update (ray end voxels). `origin` receives the scan origin (needed by
   *  applyUpdate for level 2). Call levels 0 / 2 between accumulateUpdate and integrateUpdate. */
  typename UpdateGridT::Ptr createUpdate(const std::string& source_id, int level, Eigen::Matrix<double, 3, 1>* origin = nullptr)
  {
    typename UpdateGridT::Ptr out = BackendT::createUpdateGrid(m_resolution);
    if (shardedUnavailable("createUpdate")) return out;
    if (!m_device_map) return out;
    const std::shared_ptr<InputSource> source = findSource(source_id);
    std::unique_lock<std::mutex> update_grid_lock;
    if (source) update_grid_lock = std::unique_lock<std::mutex>(source->update_grid_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    vdbm_leafset* ls = nullptr;
    double o[3]      = {0, 0, 0};
    vdbm_map* from   = m_device_map;
    if (source && source->raycaster)
    {
      // the ray end voxels of the last accumulate (level 2) stay with the handle that cast them; a raw grid (level 0) may be
      // spread over both handles and is gathered in the map's first
      if (level == 2 && source->last_cast_on_raycaster) from = source->raycaster;
      else if (level == 0) collectRaycasterLocked(source_id, *source);
    }
    if (reportOn(from, vdbm_update_create(from, source_id.c_str(), level, &ls, o)) != VDBM_OK) return out;
    if (origin) *origin = Eigen::Matrix<double, 3, 1>(o[0], o[1], o[2]);
    const std::uint64_t n = vdbm_leafset_size(ls);
    for (std::uint64_t i = 0; i < n; ++i)
      BackendT::putUpdateLeaf(*out, vdbm_leafset_origins(ls) + 3 * i, vdbm_leafset_active(ls) + 8 * i, vdbm_leafset_valmask(ls) + 8 * i);
    vdbm_leafset_free(ls);
    return out;
  }

  /*! Applies a grid made by createUpdate(level) on another map; returns the change grid (empty for level 1). */
  typename UpdateGridT::Ptr applyUpdate(const typename UpdateGridT::Ptr& update, int level,
                                        const Eigen::Matrix<double, 3, 1>& origin = Eigen::Matrix<double, 3, 1>(0, 0, 0))
  {
    typename UpdateGridT::Ptr change = BackendT::createUpdateGrid(m_resolution);
    if (shardedUnavailable("applyUpdate")) return change;
    if (!m_device_map || !update) return change;
    std::vector<std::int32_t> origins;
    std::vector<std::uint64_t> active, value;
    BackendT::forEachUpdateLeaf(*update, [&](const std::int32_t o[3], const std::uint64_t* a, const std::uint64_t* v) {
      origins.insert(origins.end(), o, o + 3);
      active.insert(active.end(), a, a + 8);
      value.insert(value.end(), v, v + 8);
    });
    const double o[3] = {origin.x(), origin.y(), origin.z()};
    std::unique_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    vdbm_leafset* ls = nullptr;
    if (report(vdbm_update_apply(m_device_map, level, origins.size() / 3, origins.data(), active.data(), value.data(), o, &ls)) != VDBM_OK)
      return change;
    const std::uint64_t n = ls ? vdbm_leafset_size(ls) : 0;
    for (std::uint64_t i = 0; i < n; ++i)
      BackendT::putUpdateLeaf(*change, vdbm_leafset_origins(ls) + 3 * i, vdbm_leafset_active(ls) + 8 * i, vdbm_leafset_valmask(ls) + 8 * i);
    if (ls) vdbm_leafset_free(ls);
    afterMapWriteLocked();
    return change;
  }


  // ------------------------------------------------------------------------------------------------------------------
  // Persistence (R:193-307). Host code on the mirror grid; the device map is the authority, so the mirror is brought up
  // to date first and a loaded grid is imported into the device map.
  // ------------------------------------------------------------------------------------------------------------------
  /*! R:193-211 */
  bool saveMap() const
  {
    const std::string map_name = m_map_directory_path + timestampString() + "_map.vdb";
    std::cout << map_name << std::endl;
    typename GridT::Ptr grid = getGrid();
    std::shared_lock map_lock(*m_map_mutex);
    return BackendT::writeGridFile(map_name, grid);
  }

  /*! R:218-256: the centres of all active voxels as a PCD file */
  bool saveMapToPCD()
  {
    const std::string pcd_path = m_map_directory_path + timestampString() + "_active_values_map.pcd";
    typename PointCloudT::Ptr cloud(new PointCloudT);
    typename GridT::Ptr grid = getGrid();
    {
      std::shared_lock map_lock(*m_map_mutex);
      cloud->points.reserve(grid->activeVoxelCount());
      BackendT::forEachActiveVoxel(*grid, [&](const openvdb::Coord& c, const TData&) {
        openvdb::Vec3d w = grid->indexToWorld(c);
        cloud->points.emplace_back(static_cast<float>(w.x() + 0.5 * m_resolution), static_cast<float>(w.y() + 0.5 * m_resolution),
                                   static_cast<float>(w.z() + 0.5 * m_resolution));
      });
    }
    cloud->points.shrink_to_fit();
    cloud->width  = static_cast<std::uint32_t>(cloud->points.size());
    cloud->height = 1;
    if (!BackendT::savePCD(pcd_path, *cloud))
    {
      std::cerr << "Could not write PCD file." << std::endl;
      return false;
    }
    std::cout << "Wrote pcd to: " << pcd_path << std::endl;
    return true;
  }

  /*! R:263-284: m_vdb_grid is replaced by the (last) grid of the file; here the device map is replaced with it too */
  bool loadMap(const std::string& file_path)
  {
    if (shardedUnavailable("loadMap")) return false;
    typename GridT::Ptr loaded = BackendT::readGridFile(file_path);
    if (!loaded) return false;
    std::unique_lock map_lock(*m_map_mutex);
    m_vdb_grid->clear();
    m_vdb_grid = loaded;
    uploadMirrorLocked();
    return true;
  }

  /*! R:295-307 */
  bool loadMapFromPCD(const std::string& file_path, const bool set_background, const bool clear_map)
  {
    if (shardedUnavailable("loadMapFromPCD")) return false;
    typename PointCloudT::Ptr cloud(new PointCloudT);
    if (!BackendT::loadPCD(file_path, *cloud))
    {
      std::cerr << "Could not open PCD file" << std::endl;
      return false;
    }
    if (!clear_map) getGrid(); // the points are added to the CURRENT map: the mirror must hold it
    std::unique_lock map_lock(*m_map_mutex);
    createMapFromPointCloud(cloud, set_background, clear_map);
    uploadMirrorLocked();
    return true;
  }

  /*! R:550-566 for one ray given by voxel indices, marked into the grid behind `update_grid_acc` (device DDA, same fp64
   *  stepping as the scan path; the host accessor receives the result). */
  void castRayIntoGrid(const openvdb::Coord& ray_origin_index, const openvdb::Coord& ray_end_index,
                       typename UpdateGridT::Accessor& update_grid_acc) const
  {
    if (shardedUnavailable("castRayIntoGrid")) return;
    if (!m_device_map) return;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    const_cast<VDBMapping*>(this)->ensureScratchSource();
    ScratchClear clear_on_exit{m_device_map};
    const std::int32_t ray[6] = {ray_origin_index.x(), ray_origin_index.y(), ray_origin_index.z(), ray_end_index.x(), ray_end_index.y(), ray_end_index.z()};
    if (report(vdbm_cast_index_rays(m_device_map, kScratchSource, 1, ray)) != VDBM_OK) return;
    vdbm_leafset* ls = nullptr;
    if (report(vdbm_update_export(m_device_map, kScratchSource, &ls)) != VDBM_OK) return;
    mergeIntoAccessor(ls, update_grid_acc);
    vdbm_leafset_free(ls);
  }

  /*! R:643-664 */
  void raytrace(const openvdb::Vec3d& ray_origin_world, const openvdb::Vec3d& ray_direction, const double max_ray_lengths, bool& success,
                openvdb::Vec3d& end_point)
  {
    std::vector<openvdb::Vec3d> origins = {ray_origin_world}, directions = {ray_direction}, end_points;
    std::vector<double> lengths = {max_ray_lengths};
    std::vector<bool> successes;
    raytrace(origins, directions, lengths, successes, end_points);
    success   = successes[0];
    end_point = end_points[0];
  }

  /*! R:675-721: first active map voxel along every ray (batch, one kernel launch: vdbm_raytrace) */
  void raytrace(const std::vector<openvdb::Vec3d>& ray_origins_world, const std::vector<openvdb::Vec3d>& ray_directions,
                const std::vector<double>& max_ray_lengths, std::vector<bool>& successes, std::vector<openvdb::Vec3d>& end_points)
  {
    const std::size_t n = ray_origins_world.size();
    successes.assign(n, false);
    end_points.resize(n);
    if (shardedUnavailable("raytrace")) return;
    if (!m_device_map || n == 0) return;
    std::vector<double> o(3 * n), d(3 * n), e(3 * n);
    std::vector<std::int32_t> ok(n);
    for (std::size_t i = 0; i < n; ++i)
      for (int k = 0; k < 3; ++k) { o[3 * i + k] = ray_origins_world[i][k]; d[3 * i + k] = ray_directions[i][k]; }
    std::shared_lock map_lock(*m_map_mutex);
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    if (report(vdbm_raytrace(m_device_map, n, o.data(), d.data(), max_ray_lengths.data(), ok.data(), e.data())) != VDBM_OK) return;
    for (std::size_t i = 0; i < n; ++i)
    {
      successes[i]  = ok[i] != 0;
      end_points[i] = openvdb::Vec3d(e[3 * i], e[3 * i + 1], e[3 * i + 2]);
    }
  }

  /*! R:810-847: world bounding box of a box given in a reference frame (corners stored as float like pcl::PointXYZ,
   *  transformed with the double matrix, min / max taken) */
  openvdb::BBoxd createWorldBoundingBox(const Eigen::Matrix<double, 3, 1>& min_boundary, const Eigen::Matrix<double, 3, 1>& max_boundary,
                                        const Eigen::Matrix<double, 4, 4>& map_to_reference_tf) const
  {
    float lo[3], hi[3];
    transformedCorners(min_boundary, max_boundary, map_to_reference_tf, lo, hi);
    return openvdb::BBoxd(openvdb::Vec3d(lo[0], lo[1], lo[2]), openvdb::Vec3d(hi[0], hi[1], hi[2]));
  }

  /*! R:921-960 */
  template <typename TResultGrid>
  typename TResultGrid::Ptr getMapSection(const Eigen::Matrix<double, 3, 1>& min_boundary, const Eigen::Matrix<double, 3, 1>& max_boundary,
                                          const Eigen::Matrix<double, 4, 4>& map_to_reference_tf, const bool full_grid = false) const
  {
    if constexpr (std::is_same<TResultGrid, UpdateGridT>::value)
      return getMapSectionUpdateGrid(min_boundary, max_boundary, map_to_reference_tf, full_grid);
    else
      return getMapSectionGrid(min_boundary, max_boundary, map_to_reference_tf, full_grid);
  }

  /*! R:1094-1147 on a host grid (OpenVDB's tools with OpenVDB installed, a 26-neighbourhood restatement otherwise) */
  template <typename TGrid>
  void morphologicalCloseMap(typename TGrid::Ptr grid, int iterations)
  {
    morphologicalDilateMap<TGrid>(grid, iterations);
    morphologicalErodeMap<TGrid>(grid, iterations);
  }
  template <typename TGrid>
  void morphologicalOpenMap(typename TGrid::Ptr grid, int iterations)
  {
    morphologicalErodeMap<TGrid>(grid, iterations);
    morphologicalDilateMap<TGrid>(grid, iterations);
  }
  template <typename TGrid>
  void morphologicalDilateMap(typename TGrid::Ptr grid, int iterations) { BackendT::dilateActive(*grid, iterations); }
  template <typename TGrid>
  void morphologicalErodeMap(typename TGrid::Ptr grid, int iterations) { BackendT::erodeActive(*grid, iterations); }

  /*! R:1198-1207 */
  void addArtificialPolygon(const std::vector<Eigen::Matrix<double, 4, 1> >& polygon, const double negative_height, const double positive_height)
  {
    if (shardedUnavailable("addArtificialPolygon")) return;
    if (!m_device_map || polygon.empty()) return;
    const std::uint32_t count = static_cast<std::uint32_t>(polygon.size());
    std::vector<double> xyz;
    for (const auto& p : polygon) xyz.insert(xyz.end(), {p[0], p[1], p[2]});
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_artificial_walls_add(m_device_map, 1, &count, xyz.data(), negative_height, positive_height, /*closed=*/1));
    m_artificial_areas_present = true;
  }

  /*! R:1217-1236 */
  void addArtificialWall(const Eigen::Matrix<double, 4, 1>& start, const Eigen::Matrix<double, 4, 1>& end, const double negative_height,
                         const double positive_height)
  {
    if (shardedUnavailable("addArtificialWall")) return;
    if (!m_device_map) return;
    const std::uint32_t count = 2;
    const double xyz[6]       = {start[0], start[1], start[2], end[0], end[1], end[2]};
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_artificial_walls_add(m_device_map, 1, &count, xyz, negative_height, positive_height, /*closed=*/0));
    m_artificial_areas_present = true;
  }

  /*! R:1246-1268 */
  std::vector<uint8_t> compressString(const std::string& string) const { return detail::zstdCompress(string, m_compression_level); }
  /*! R:1277-1300 */
  std::string decompressByteArray(const std::vector<uint8_t>& byte_array) const { return detail::zstdDecompress(byte_array); }
  /*! R:1310-1317 (bytes are OpenVDB's stream format only when OpenVDB is installed, see detail/host_io.hpp) */
  template <typename TGrid>
  std::vector<uint8_t> gridToByteArray(typename TGrid::Ptr grid)
  {
    return compressString(BackendT::template gridToString<TGrid>(grid));
  }
  /*! R:1328-1338 */
  template <typename TGrid>
  typename TGrid::Ptr byteArrayToGrid(std::vector<uint8_t> byte_array)
  {
    return BackendT::template stringToGrid<TGrid>(decompressByteArray(byte_array));
  }

  /*! R:1436-1449. fast_mode keeps no host-side intersector: the fast raycast probes the device map directly. */
  void updateVolumeRayIntersectors() {}

  /*! R:1343 */
  std::shared_ptr<std::shared_mutex> getMapMutex() { return m_map_mutex; }

  /*! R:1352-1375. The source is registered BEFORE its worker thread starts (the reference starts the thread first). */
  void addInputSource(std::string source_id, double max_range, double max_rate)
  {
    const double range = (max_range == 0) ? m_max_range : max_range;
    const auto period  = (max_rate <= 0) ? std::chrono::milliseconds(0) : std::chrono::milliseconds((int)(1000.0 / max_rate));
    if (m_device_group)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      reportGroup(vdbm_group_source_add(m_device_group, source_id.c_str(), max_range));
    }
    if (m_device_map)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      report(vdbm_source_add(m_device_map, source_id.c_str(), max_range)); // a re-added source starts with an empty update grid
    }
    const std::shared_ptr<InputSource> existing = findSource(source_id);
    if (existing)
    {
      // Re-adding a source (the reference overwrites the map entry, R:1374): its worker thread is blocked on THIS object's
      // condition variable, so the object stays and only its parameters change; pending input is dropped like the
      // reference's fresh InputSource would.
      {
        std::unique_lock lock(existing->input_data_mutex);
        existing->max_range        = range;
        existing->max_input_period = period;
        existing->input_data.reset();
      }
      std::unique_lock grid_lock(existing->update_grid_mutex);
      if (existing->raycaster)
      {
        vdbm_source_add(existing->raycaster, source_id.c_str(), range); // empty grid, new (resolved, R:1356-1363) range
        existing->raycaster_holds_data = false;
      }
      existing->shared_holds_data = false;
      return;
    }
    auto s              = std::make_shared<InputSource>();
    s->source_id        = source_id;
    s->max_range        = range;
    s->max_input_period = period;
    {
      // published under the sources lock: the integration and accumulation threads look sources up while a node is still
      // registering its sensors (the reference mutates the std::map unguarded, R:1374 against R:380 / R:1388)
      std::unique_lock<std::shared_mutex> sources_lock(m_sources_mutex);
      m_input_sources[source_id] = s;
    }
    m_worker_threads[source_id] = std::thread(&VDBMapping::accumulationThread, this, source_id);
  }

  /*! R:1456-1469 */
  virtual void setConfig(const TConfig& config)
  {
    if (config.max_range < 0.0)
    {
      std::cerr << "Max range of " << config.max_range << " invalid. Range cannot be negative." << std::endl;
      return;
    }
    m_max_range           = config.max_range;
    m_map_directory_path  = config.map_directory_path;
    m_fast_mode           = config.fast_mode;
    m_accumulation_period = (int)(config.accumulation_period * 1000);
    m_config_set          = true;
    if (m_device_group && m_fast_mode)
    {
      std::cerr << "vdb_mapping (B200): fast_mode is not available while the map is sharded over several devices; normal raycasting is used"
                << std::endl;
      m_fast_mode = false;
    }
    if (m_device_map)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      report(vdbm_set_fast_mode(m_device_map, m_fast_mode ? 1 : 0)); // R:333 / R:520: castRayIntoGridFast from now on
    }
  }

  /*! Counters of the device path (rays, visits, voxel updates, kernel times); not part of the reference API. */
  bool deviceStats(vdbm_stats_t& out) const
  {
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex); // not while another thread is inside a call on the handle
      if (m_device_group) return vdbm_group_stats(m_device_group, &out) == VDBM_OK; // summed over the shards
      if (!m_device_map || vdbm_stats(m_device_map, &out) != VDBM_OK) return false;
    }
    // rays cast on the sources' own handles (SourceConcurrency) belong to this map's raycast counters
    for (auto& kv : sourcesSnapshot())
    {
      std::unique_lock<std::mutex> grid_lock(kv.second->update_grid_mutex);
      vdbm_stats_t s;
      if (!kv.second->raycaster || vdbm_stats(kv.second->raycaster, &s) != VDBM_OK) continue;
      out.rays += s.rays;
      out.nan_skipped += s.nan_skipped;
      out.clipped += s.clipped;
      out.visits += s.visits;
      out.gpu_launches = std::max(out.gpu_launches, s.gpu_launches); // one process-wide launch counter
    }
    return true;
  }

protected:
  static constexpr const char* kScratchSource = "\x01vdbm_scratch";
  /*! empties the scratch source's update grid on scope exit without touching the map (re-adding a source clears its grid) */
  struct ScratchClear
  {
    vdbm_map* m;
    ~ScratchClear() { if (m) vdbm_source_add(m, kScratchSource, 1.0); }
  };

  int report(int rc) const
  {
    // the reference reports problems on std::cout / std::cerr and carries on; so does the shim
    if (rc != VDBM_OK && rc != VDBM_ERR_UNKNOWN_SOURCE && m_device_map)
      std::cerr << "vdb_mapping (B200): " << vdbm_last_error(m_device_map) << std::endl;
    return rc;
  }


  int reportOn(vdbm_map* handle, int rc) const
  {
    if (rc != VDBM_OK && rc != VDBM_ERR_UNKNOWN_SOURCE && handle) std::cerr << "vdb_mapping (B200): " << vdbm_last_error(handle) << std::endl;
    return rc;
  }

  /*! the registered sources are looked up by the accumulation threads, the integration thread and every caller while
   *  addInputSource may still be adding one: lookups copy the shared_ptr under a reader lock */
  std::shared_ptr<InputSource> findSource(const std::string& source_id) const
  {
    std::shared_lock<std::shared_mutex> sources_lock(m_sources_mutex);
    auto it = m_input_sources.find(source_id);
    return it == m_input_sources.end() ? std::shared_ptr<InputSource>() : it->second;
  }
  std::vector<std::pair<std::string, std::shared_ptr<InputSource> > > sourcesSnapshot() const
  {
    std::shared_lock<std::shared_mutex> sources_lock(m_sources_mutex);
    return std::vector<std::pair<std::string, std::shared_ptr<InputSource> > >(m_input_sources.begin(), m_input_sources.end());
  }
  std::size_t sourceCount() const
  {
    std::shared_lock<std::shared_mutex> sources_lock(m_sources_mutex);
    return m_input_sources.size();
  }

  int reportGroup(int rc) const
  {
    if (rc != VDBM_OK && rc != VDBM_ERR_UNKNOWN_SOURCE && m_device_group)
      std::cerr << "vdb_mapping (B200): " << vdbm_group_last_error(m_device_group) << std::endl;
    return rc;
  }

  /*! members that need the whole map on one device say so when the map is sharded (setDevices) and do nothing */
  bool shardedUnavailable(const char* member) const
  {
    if (!m_device_group) return false;
    std::cerr << "vdb_mapping (B200): " << member << " is not available while the map is sharded over several devices" << std::endl;
    return true;
  }

  /*! f(handle, shard index) for the map's handle, or for every shard of a sharded map */
  template <typename F>
  void forEachMapHandle(F&& f) const
  {
    if (m_device_group)
      for (std::int32_t i = 0; i < vdbm_group_size(m_device_group); ++i) f(vdbm_group_shard(m_device_group, i), std::size_t(i));
    else if (m_device_map) f(m_device_map, std::size_t(0));
  }

  /*! insertPointCloud R:399-406 on a sharded map: raycast on every device, exchange over NVLink, updateMap on every shard */
  void insertSharded(const typename PointCloudT::ConstPtr& cloud, const Eigen::Matrix<double, 3, 1>& origin, const std::string& source_id)
  {
    if (!cloud) return;
    m_map_mutex_requested = true;
    std::unique_lock map_lock(*m_map_mutex);
    m_map_mutex_requested = false;
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    const double o[3] = {origin.x(), origin.y(), origin.z()};
    reportGroup(vdbm_group_insert(m_device_group, source_id.c_str(), cloud->points.data(), cloud->points.size(), sizeof(PointT), o));
    afterMapWriteLocked();
  }

  /*! does accumulateUpdate() cast this map's rays on per-source handles right now? (see SourceConcurrency) */
  bool wantsRaycaster() const
  {
    if (!m_device_map || !m_config_set || m_fast_mode || m_source_concurrency == SourceConcurrency::Shared) return false;
    return m_source_concurrency == SourceConcurrency::PerSource || sourceCount() > 1;
  }

  /*! the source's raycast-only handle, created and (re)configured on demand; nullptr = use the map's handle.
   *  Caller holds src.update_grid_mutex. */
  vdbm_map* sourceRaycaster(const std::string& source_id, InputSource& src)
  {
    if (!wantsRaycaster()) return nullptr;
    if (!src.raycaster)
    {
      vdbm_params p{};
      p.resolution            = m_resolution;
      p.device                = -1;
      p.replicate_probe_quirk = 1;
      p.map_capacity_leaves   = 1; // it never integrates
      if (vdbm_create(&p, &src.raycaster) != VDBM_OK)
      {
        src.raycaster = nullptr;
        return nullptr;
      }
      // only the range matters to a raycast (R:1456-1469); the probabilities just have to be valid. The source gets the range
      // addInputSource resolved for it (R:1356-1363: 0 = the config range at THAT time), like its twin on the map's handle.
      if (vdbm_set_config(src.raycaster, m_max_range, 0.7, 0.4, 0.12, 0.97) != VDBM_OK ||
          vdbm_source_add(src.raycaster, source_id.c_str(), src.max_range) != VDBM_OK)
      {
        vdbm_destroy(src.raycaster);
        src.raycaster = nullptr;
        return nullptr;
      }
      src.raycaster_range = m_max_range;
    }
    if (src.raycaster_range != m_max_range)
    {
      if (vdbm_set_config(src.raycaster, m_max_range, 0.7, 0.4, 0.12, 0.97) != VDBM_OK) return nullptr;
      src.raycaster_range = m_max_range;
    }
    return src.raycaster;
  }

  /*! moves what the source accumulated on its own handle into its update grid on the map's handle, device to device.
   *  Caller holds m_device_mutex and either the exclusive map lock or src.update_grid_mutex. */
  void collectRaycasterLocked(const std::string& source_id, InputSource& src)
  {
    if (!src.raycaster || !src.raycaster_holds_data) return;
    src.raycaster_holds_data = false;
    std::uint64_t count   = 0;
    const void* d_records = nullptr;
    if (reportOn(src.raycaster, vdbm_update_partition(src.raycaster, source_id.c_str(), 1, &count, &d_records)) != VDBM_OK) return;
    if (count) report(vdbm_update_import_device(m_device_map, source_id.c_str(), d_records, count));
    if (count) src.shared_holds_data = true;
  }

  static std::string timestampString()
  {
    auto timestamp     = std::chrono::system_clock::now();
    std::time_t now_tt = std::chrono::system_clock::to_time_t(timestamp);
    std::tm tm         = *std::localtime(&now_tt);
    std::stringstream sstime;
    sstime << std::put_time(&tm, "%Y-%m-%d_%H-%M-%S");
    return sstime.str();
  }

  /*! R:815-843: the 8 corners as float points, transformed (pcl::transformPointCloud with a double matrix:
   *  float(M(r,0)*x + M(r,1)*y + M(r,2)*z + M(r,3))), then min / max per axis */
  static void transformedCorners(const Eigen::Matrix<double, 3, 1>& min_boundary, const Eigen::Matrix<double, 3, 1>& max_boundary,
                                 const Eigen::Matrix<double, 4, 4>& tf, float lo[3], float hi[3])
  {
    bool first = true;
    for (int k = 0; k < 8; ++k)
    {
      const float p[3] = {static_cast<float>((k & 4) ? max_boundary.x() : min_boundary.x()),
                          static_cast<float>((k & 2) ? max_boundary.y() : min_boundary.y()),
                          static_cast<float>((k & 1) ? max_boundary.z() : min_boundary.z())};
      for (int r = 0; r < 3; ++r)
      {
        const float q = static_cast<float>(tf(r, 0) * p[0] + tf(r, 1) * p[1] + tf(r, 2) * p[2] + tf(r, 3));
        if (first || q < lo[r]) lo[r] = q;
        if (first || q > hi[r]) hi[r] = q;
      }
      first = false;
    }
  }

  /*! the whole host grid becomes the device map (after loadMap / createMapFromPointCloud; caller holds the map lock) */
  void uploadMirrorLocked()
  {
    if (!m_device_map) return;
    std::vector<std::int32_t> origins;
    std::vector<std::uint64_t> active;
    std::vector<float> values;
    BackendT::forEachMapLeaf(*m_vdb_grid, [&](const std::int32_t o[3], const float* v, const std::uint64_t* a) {
      origins.insert(origins.end(), o, o + 3);
      values.insert(values.end(), v, v + 512);
      active.insert(active.end(), a, a + 8);
    });
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_map_import(m_device_map, origins.size() / 3, origins.data(), active.data(), values.data(), 1));
    vdbm_leafset* ls = nullptr; // nothing is dirty now: the mirror already equals the device map
    if (vdbm_map_export(m_device_map, 1, &ls) == VDBM_OK) vdbm_leafset_free(ls);
    m_mirror_stale = false;
  }

  // R:1472-1476: the node operations of the reference. On the B200 path updateMap runs them inside apply_update_kernel
  // for the occupancy semantics (TData = float); the virtuals stay for host-side users (createMapFromPointCloud,
  // subclasses that call them directly). Overriding them does NOT change what the device computes.
  virtual bool updateFreeNode(TData& /*voxel_value*/, bool& /*active*/) { return false; }
  virtual bool updateOccupiedNode(TData& /*voxel_value*/, bool& /*active*/) { return false; }
  virtual bool setNodeToFree(TData& /*voxel_value*/, bool& /*active*/) { return false; }
  virtual bool setNodeToOccupied(TData& /*voxel_value*/, bool& /*active*/) { return false; }
  virtual bool setNodeState(TData& /*voxel_value*/, bool& /*active*/) { return false; }
  /*! R:1478-1480 */
  virtual void createMapFromPointCloud(const typename PointCloudT::Ptr& /*cloud*/, const bool /*set_background*/, const bool /*clear_map*/)
  {
    std::cerr << "Not implemented for data type" << std::endl;
  }

  /*! after a device-side map write (caller holds m_device_mutex) */
  void afterMapWriteLocked()
  {
    if (m_mirror_mode == MirrorMode::Eager) syncMirrorLocked();
    else m_mirror_stale = true;
  }

  bool setPoints(const typename PointCloudT::ConstPtr& cloud, int occupied)
  {
    if (shardedUnavailable("addPointsToGrid / removePointsFromGrid")) return true;
    if (!m_device_map || !cloud) return true; // the reference returns true unconditionally (R:428,446)
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    report(vdbm_points_set(m_device_map, cloud->points.data(), cloud->points.size(), sizeof(PointT), occupied));
    afterMapWriteLocked();
    return true;
  }

  void ensureScratchSource()
  {
    if (!m_scratch_ready && m_device_map)
    {
      vdbm_source_add(m_device_map, kScratchSource, 1.0);
      m_scratch_ready = true;
    }
  }

  void mergeIntoAccessor(vdbm_leafset* ls, typename UpdateGridT::Accessor& acc) const
  {
    const std::uint64_t n = vdbm_leafset_size(ls);
    for (std::uint64_t i = 0; i < n; ++i)
    {
      const std::int32_t* o  = vdbm_leafset_origins(ls) + 3 * i;
      const std::uint64_t* a = vdbm_leafset_active(ls) + 8 * i;
      const std::uint64_t* v = vdbm_leafset_valmask(ls) + 8 * i;
      for (unsigned k = 0; k < 512; ++k)
      {
        if (!((a[k >> 6] >> (k & 63)) & 1u)) continue;
        const openvdb::Coord c(o[0] + int(k >> 6), o[1] + int((k >> 3) & 7), o[2] + int(k & 7));
        if ((v[k >> 6] >> (k & 63)) & 1u) acc.setValueOn(c, true); // R:535
        else acc.setActiveState(c, true);                          // R:563
      }
    }
  }

  /*! copies the leaves modified since the last call from the device into the host grid (caller holds the map lock
   *  or is the only user) */
  void syncMirrorLocked()
  {
    m_mirror_stale = false;
    if (!m_device_map && !m_device_group) return;
    if (m_shard_tables.empty())
    {
      m_shard_tables.assign(1, {});
      m_shard_generations.assign(1, ~std::uint64_t(0));
    }
    // the tables "device pool index -> host leaf" (one per shard: every shard has its own pool) are only good for one grid
    // object, one generation of that pool and as long as no leaf was removed from the host grid
    const std::uint64_t epoch = BackendT::gridEpoch(*m_vdb_grid);
    const bool grid_changed   = m_mirror_table_grid != m_vdb_grid.get() || m_mirror_table_epoch != epoch;
    m_mirror_table_grid       = m_vdb_grid.get();
    m_mirror_table_epoch      = epoch;
    forEachMapHandle([&](vdbm_map* handle, std::size_t shard) {
      const std::uint64_t generation = vdbm_map_generation(handle);
      if (grid_changed || m_shard_generations[shard] != generation) m_shard_tables[shard].clear();
      m_shard_generations[shard] = generation;
      m_mirror_sink_shard        = shard;
      reportOn(handle, vdbm_map_mirror(handle, m_mirror_chunk_leaves, &VDBMapping::mirrorSink, this, nullptr));
    });
  }

  /*! vdbm_mirror_sink: one chunk of modified leaves lands in the host grid while the next chunk is still on its way */
  static int mirrorSink(void* user, std::uint64_t n, const std::uint32_t* leaf_index, const std::int32_t* origins, const float* values,
                        const std::uint64_t* active)
  {
    VDBMapping* self = static_cast<VDBMapping*>(user);
    BackendT::putMapLeavesIndexed(*self->m_vdb_grid, self->m_shard_tables[self->m_mirror_sink_shard], n, leaf_index, origins, values, active);
    return 0;
  }

public:
  /*! Not part of the reference API: forget the leaf table of the host mirror. Needed only by a consumer that removes or
   *  merges leaf nodes of the grid returned by getGrid() (prune, clear ...) while the map keeps running. */
  void invalidateMirrorTable()
  {
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    for (auto& table : m_shard_tables) table.clear();
  }
  /*! Not part of the reference API: leaves per chunk of the mirror transfer (0 = library default). */
  void setMirrorChunkLeaves(std::uint64_t n) { m_mirror_chunk_leaves = n; }

protected:

  /*! R:1383-1411 */
  void accumulationThread(std::string source_id)
  {
    while (!m_config_set && !m_thread_stop_signal) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    while (!m_thread_stop_signal)
    {
      std::shared_ptr<InputSource> src = findSource(source_id);
      if (!src) // started before its source was published: try again
      {
        std::this_thread::sleep_for(std::chrono::milliseconds(1));
        continue;
      }
      auto wake_time = std::chrono::high_resolution_clock::now() + src->max_input_period;
      std::unique_lock lock(src->input_data_mutex);
      src->data_available_cv.wait(lock, [&] { return src->input_data || m_thread_stop_signal; });
      if (m_thread_stop_signal) break;
      if (!m_map_mutex_requested)
      {
        auto measurement = *src->input_data;
        src->input_data.reset();
        lock.unlock();
        accumulateUpdate(measurement.first, measurement.second, source_id);
        std::this_thread::sleep_until(wake_time);
      }
      else
      {
        lock.unlock();
        std::this_thread::yield();
      }
    }
  }

  /*! R:1416-1430. An accumulation period of 0 (or an uninitialised one, as in the reference's tests) only lets the
   *  explicit integrateUpdate()/insertPointCloud() calls integrate. */
  void integrationThread()
  {
    while (!m_config_set && !m_thread_stop_signal) std::this_thread::sleep_for(std::chrono::milliseconds(10));
    while (!m_thread_stop_signal)
    {
      const int period_ms = m_accumulation_period; // one read: setConfig may change it while this thread runs
      if (period_ms <= 0 || period_ms > 3600000)
      {
        std::this_thread::sleep_for(std::chrono::milliseconds(20));
        continue;
      }
      auto wake_time = std::chrono::high_resolution_clock::now() + std::chrono::milliseconds(period_ms);
      integrateUpdate();
      std::this_thread::sleep_until(wake_time);
    }
  }

  vdbm_map* m_device_map = nullptr;
  vdbm_group* m_device_group = nullptr; // setDevices(): the map is sharded over several GPUs (then m_device_map is null)
  typename GridT::Ptr m_vdb_grid;
  double m_max_range = 0.0;
  double m_resolution;
  bool m_fast_mode          = false;
  std::atomic<int> m_accumulation_period{0}; // written by setConfig, read by the integration thread (a plain int in the reference, R:1416 / 1466: a data race there)
  std::string m_map_directory_path;
  std::atomic<bool> m_config_set;
  MirrorMode m_mirror_mode = MirrorMode::Eager;
  SourceConcurrency m_source_concurrency = SourceConcurrency::Auto;
  mutable std::atomic<bool> m_mirror_stale{false}; // set under the device mutex, peeked at by getGrid() without it
  std::vector<std::vector<typename BackendT::MapLeafT*> > m_shard_tables; // per map handle: device pool index -> leaf of m_vdb_grid
  std::vector<std::uint64_t> m_shard_generations;
  std::size_t m_mirror_sink_shard     = 0;
  const void* m_mirror_table_grid     = nullptr;
  std::uint64_t m_mirror_table_epoch  = 0;
  std::uint64_t m_mirror_chunk_leaves = 0;
  bool m_scratch_ready        = false;
  bool m_artificial_areas_present = false; // R:1529
  int m_compression_level         = 1;     // R:1524
  mutable std::mutex m_device_mutex; // serialises calls into the MAP's (thread-compatible) C ABI handle; sources may own handles of their own
  mutable std::shared_ptr<std::shared_mutex> m_map_mutex;
  mutable std::atomic<bool> m_map_mutex_requested{false};
  std::map<std::string, std::shared_ptr<InputSource> > m_input_sources;
  mutable std::shared_mutex m_sources_mutex; // guards the STRUCTURE of m_input_sources (see findSource)
  std::atomic<bool> m_thread_stop_signal{false};
  std::map<std::string, std::thread> m_worker_threads;
  std::thread m_integration_thread;
};

} // namespace vdb_mapping

#endif /* VDB_MAPPING_VDB_MAPPING_H_INCLUDED */
