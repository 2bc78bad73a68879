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
    if (config.max_range < 0.0 || !m_device_map) return; // base already complained / no device
    std::lock_guard<std::mutex> device_lock(m_device_mutex);
    vdbm_set_config(m_device_map, config.max_range, config.prob_hit, config.prob_miss, config.prob_thres_min, config.prob_thres_max);
    float lo[6];
    vdbm_get_logodds(m_device_map, lo);
    m_logodds_hit = lo[0]; m_logodds_miss = lo[1]; m_logodds_thres_min = lo[2]; m_logodds_thres_max = lo[3];
    m_max_logodds = lo[4]; m_min_logodds = lo[5];
    m_config_set  = true;
  }

protected:
  // O:179-199 (kept for subclasses that read them; the device holds the authoritative copies)
  float m_logodds_hit = 0, m_logodds_miss = 0, m_logodds_thres_min = 0, m_logodds_thres_max = 0, m_max_logodds = 0, m_min_logodds = 0;
};

} // namespace vdb_mapping

#endif /* VDB_MAPPING_OCCUPANCY_VDB_MAPPING_H_INCLUDED */
