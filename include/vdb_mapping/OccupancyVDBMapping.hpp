// OccupancyVDBMapping.hpp — B200 drop-in for vdb_mapping::OccupancyVDBMapping (float log-odds occupancy map).
// Reference: /root/reference/include/vdb_mapping/OccupancyVDBMapping.hpp (cited as O:<line>).
//
// The reference implements the hit/miss update as protected virtual node operations (O:92-117) called once per
// voxel; here those operations run inside the CUDA update kernel (apply_update_kernel), and this class only
// validates the configuration exactly like O:59-89 and hands the probabilities to the device library, which
// derives the log-odds constants with the reference's expression static_cast<float>(log(p) - log(1 - p)).
#ifndef VDB_MAPPING_OCCUPANCY_VDB_MAPPING_H_INCLUDED
#define VDB_MAPPING_OCCUPANCY_VDB_MAPPING_H_INCLUDED

#include "vdb_mapping/VDBMapping.hpp"

namespace vdb_mapping {

/*! O:38-44 */
struct Config : BaseConfig
{
  double prob_hit;
  double prob_miss;
  double prob_thres_min;
  double prob_thres_max;
};

class OccupancyVDBMapping : public VDBMapping<float, Config>
{
public:
  explicit OccupancyVDBMapping(const double resolution)
    : VDBMapping<float, Config>(resolution)
  {
  }

  /*! O:59-89. The base class accepts the range first (and flags the map as configured, like the reference does
   *  before its own checks); invalid probabilities are reported and leave the log-odds untouched. */
  inline void setConfig(const Config& config) override
  {
    VDBMapping::setConfig(config);
    // The device takes the same decisions in the same order (vdbm_set_config: a negative range changes nothing; otherwise
    // range and the "configured" flag are accepted BEFORE the probability checks, O:61 then O:65-76), so host and device
    // never disagree on whether the map is configured.
    int device_rc = VDBM_OK;
    if (m_device_group)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      device_rc = vdbm_group_set_config(m_device_group, config.max_range, config.prob_hit, config.prob_miss, config.prob_thres_min, config.prob_thres_max);
    }
    if (m_device_map)
    {
      std::lock_guard<std::mutex> device_lock(m_device_mutex);
      device_rc = vdbm_set_config(m_device_map, config.max_range, config.prob_hit, config.prob_miss, config.prob_thres_min, config.prob_thres_max);
    }
    if (config.prob_miss > 0.5)
    {
      std::cerr << "Probability for a miss should be below 0.5 but is " << config.prob_miss << std::endl;
      return;
    }
    if (config.prob_hit < 0.5)
    {
      std::cerr << "Probability for a hit should be above 0.5 but is " << config.prob_hit << std::endl;
      return;
    }
    vdbm_map* any_handle = m_device_group ? vdbm_group_shard(m_device_group, 0) : m_device_map; // the shards share one configuration
    if (config.max_range < 0.0 || !any_handle || device_rc != VDBM_OK) return; // base already complained / no device
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    float lo[6];
    vdbm_get_logodds(any_handle, lo);
    m_logodds_hit = lo[0]; m_logodds_miss = lo[1]; m_logodds_thres_min = lo[2]; m_logodds_thres_max = lo[3];
    m_max_logodds = lo[4]; m_min_logodds = lo[5];
    m_config_set  = true;
  }

protected:
  // O:92-135: the node operations, host versions (the device runs the same arithmetic inside apply_update_kernel /
  // overwrite_kernel / restore_state_kernel; these serve host-side callers such as createMapFromPointCloud)
  inline bool updateFreeNode(float& voxel_value, bool& active) override
  {
    voxel_value += m_logodds_miss;
    if (voxel_value < m_logodds_thres_min)
    {
      active = false;
      if (voxel_value < m_min_logodds) voxel_value = m_min_logodds;
    }
    return true;
  }
  inline bool updateOccupiedNode(float& voxel_value, bool& active) override
  {
    voxel_value += m_logodds_hit;
    if (voxel_value > m_logodds_thres_max)
    {
      active = true;
      if (voxel_value > m_max_logodds) voxel_value = m_max_logodds;
    }
    return true;
  }
  inline bool setNodeToFree(float& voxel_value, bool& active) override
  {
    voxel_value = m_min_logodds;
    active      = false;
    return true;
  }
  inline bool setNodeToOccupied(float& voxel_value, bool& active) override
  {
    voxel_value = m_max_logodds;
    active      = true;
    return true;
  }
  inline bool setNodeState(float& voxel_value, bool& active) override
  {
    active = voxel_value > m_logodds_thres_max;
    return true;
  }

  /*! O:136-174: occupied voxels from a cloud (plain transform, no +res/2 rule), optionally the inactive rest of their
   *  bounding box set to the minimum. Runs on the host grid; the caller (loadMapFromPCD) then imports it into the device map. */
  inline void createMapFromPointCloud(const PointCloudT::Ptr& cloud, const bool set_background, const bool clear_map) override
  {
    if (clear_map) m_vdb_grid->clear();
    typename GridT::Accessor acc = m_vdb_grid->getAccessor();
    for (const auto& point : cloud->points)
    {
      const openvdb::Vec3d index_coord = m_vdb_grid->worldToIndex(openvdb::Vec3d(point.x, point.y, point.z));
      acc.setValueOn(openvdb::Coord::floor(index_coord), m_max_logodds);
    }
    openvdb::CoordBBox bbox;
    if (set_background && detail::Backend<float>::activeBBox(*m_vdb_grid, bbox))
    {
      for (int x = bbox.min().x(); x <= bbox.max().x(); ++x)
        for (int y = bbox.min().y(); y <= bbox.max().y(); ++y)
          for (int z = bbox.min().z(); z <= bbox.max().z(); ++z)
          {
            const openvdb::Coord c(x, y, z);
            if (!acc.isValueOn(c)) acc.setValueOff(c, m_min_logodds);
          }
    }
    detail::Backend<float>::prune(*m_vdb_grid);
  }

  // O:179-199 (kept for subclasses that read them; the device holds the authoritative copies)
  float m_logodds_hit = 0, m_logodds_miss = 0, m_logodds_thres_min = 0, m_logodds_thres_max = 0, m_max_logodds = 0, m_min_logodds = 0;
};

} // namespace vdb_mapping

#endif /* VDB_MAPPING_OCCUPANCY_VDB_MAPPING_H_INCLUDED */
