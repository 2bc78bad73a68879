// bench_shim.cpp — throughput of the scan-integration path through the REFERENCE-COMPATIBLE C++ class API
// (include/vdb_mapping/OccupancyVDBMapping.hpp: insertPointCloud / getGrid), i.e. what an unmodified consumer of
// vdb_mapping sees after switching headers. Reads scans written by tools/bench_shim.py (one file: header, then per scan
// origin[3] f64 + n u32 + n x pcl::PointXYZ), runs them in MirrorMode::Lazy (default: the pipelined insert) or Eager, and
// prints one JSON line.
#include <algorithm>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <string>
#include <thread>
#include <vector>

#include <vdb_mapping/OccupancyVDBMapping.hpp>

using vdb_mapping::Config;
using vdb_mapping::OccupancyVDBMapping;

int main(int argc, char** argv)
{
  if (argc < 2) { std::fprintf(stderr, "usage: bench_shim scans.bin [eager|lazy] [warmup] [max_scans]\n"); return 2; }
  const std::string mode = argc > 2 ? argv[2] : "lazy";
  const bool eager = mode == "eager";
  // "sources4" / "sources4_shared": every scan split into 4 azimuth sectors fed as 4 input sources from 4 threads (the
  // reference's one parallel axis, R:1373), then integrateUpdate; per-source raycast handles (SourceConcurrency::Auto) or
  // all sources on the map's handle
  const bool multi = mode.rfind("sources4", 0) == 0;
  // "groupN": the map sharded over devices 0 .. N-1 (setDevices; one process, no Python), lazy mirror
  const int group_n = mode.rfind("group", 0) == 0 ? std::atoi(mode.c_str() + 5) : 0;
  const int warmup = argc > 3 ? std::atoi(argv[3]) : 5;
  FILE* f = std::fopen(argv[1], "rb");
  if (!f) { std::perror("open"); return 2; }
  double hdr[7]; // resolution, max_range, prob_hit, prob_miss, thres_min, thres_max, n_scans
  if (std::fread(hdr, sizeof(double), 7, f) != 7) return 2;
  int n_scans = int(hdr[6]);
  if (argc > 4) n_scans = std::min(n_scans, std::atoi(argv[4]));
  std::vector<OccupancyVDBMapping::PointCloudT::Ptr> clouds;
  std::vector<Eigen::Matrix<double, 3, 1> > origins;
  for (int k = 0; k < n_scans; ++k)
  {
    double o[3];
    unsigned n = 0;
    if (std::fread(o, sizeof(double), 3, f) != 3 || std::fread(&n, sizeof(unsigned), 1, f) != 1) return 2;
    OccupancyVDBMapping::PointCloudT::Ptr c(new OccupancyVDBMapping::PointCloudT);
    c->points.resize(n);
    static_assert(sizeof(OccupancyVDBMapping::PointT) == 16, "pcl::PointXYZ layout");
    if (std::fread(c->points.data(), 16, n, f) != n) return 2;
    clouds.push_back(c);
    origins.emplace_back(o[0], o[1], o[2]);
  }
  std::fclose(f);

  OccupancyVDBMapping map(hdr[0]);
  if (group_n > 0)
  {
    std::vector<int> devices;
    const bool same_device = std::getenv("VDBM_GROUP_ONE_DEVICE") != nullptr; // experiments: all shards on device 0
    for (int d = 0; d < group_n; ++d) devices.push_back(same_device ? 0 : d);
    if (!map.setDevices(devices)) return 3;
  }
  Config conf;
  conf.max_range = hdr[1]; conf.prob_hit = hdr[2]; conf.prob_miss = hdr[3]; conf.prob_thres_min = hdr[4]; conf.prob_thres_max = hdr[5];
  conf.fast_mode = false; conf.accumulation_period = 0.0;
  map.setConfig(conf);
  map.addInputSource("lidar", conf.max_range, 0);
  map.setMirrorMode(eager ? vdb_mapping::MirrorMode::Eager : vdb_mapping::MirrorMode::Lazy);
  if (const char* e = std::getenv("VDBM_MIRROR_CHUNK")) map.setMirrorChunkLeaves(std::strtoull(e, nullptr, 10)); // experiments
  const char* ids[4] = {"lidar_0", "lidar_1", "lidar_2", "lidar_3"};
  std::vector<std::vector<OccupancyVDBMapping::PointCloudT::Ptr> > parts;
  if (multi)
  {
    for (const char* id : ids) map.addInputSource(id, conf.max_range, 0);
    if (mode == "sources4_shared") map.setSourceConcurrency(vdb_mapping::SourceConcurrency::Shared);
    for (auto& c : clouds)
    {
      parts.emplace_back();
      const std::size_t n = c->points.size();
      for (int s = 0; s < 4; ++s)
      {
        OccupancyVDBMapping::PointCloudT::Ptr p(new OccupancyVDBMapping::PointCloudT);
        p->points.assign(c->points.begin() + n * s / 4, c->points.begin() + n * (s + 1) / 4);
        parts.back().push_back(p);
      }
    }
  }

  unsigned long long rays = 0;
  std::chrono::steady_clock::time_point t0;
  for (int k = 0; k < n_scans; ++k)
  {
    if (k == warmup) { map.getGrid(); t0 = std::chrono::steady_clock::now(); }
    if (multi)
    {
      std::thread th[4];
      for (int s = 0; s < 4; ++s) th[s] = std::thread([&, s] { map.accumulateUpdate(parts[k][s], origins[k], ids[s]); });
      for (auto& t : th) t.join();
      map.integrateUpdate();
    }
    else
    {
      const auto t_scan = std::chrono::steady_clock::now();
      map.insertPointCloud(clouds[k], origins[k], "lidar");
      if (std::getenv("VDBM_BENCH_TRACE"))
        std::fprintf(stderr, "scan %d: %.3f ms\n", k, 1e3 * std::chrono::duration<double>(std::chrono::steady_clock::now() - t_scan).count());
    }
    if (k >= warmup) rays += clouds[k]->points.size();
  }
  vdbm_stats_t st;
  map.deviceStats(st);
  // finish the last queued scan without copying the map back: a section of an empty box goes through the ABI
  map.getMapSectionUpdateGrid(Eigen::Matrix<double, 3, 1>(1e6, 1e6, 1e6), Eigen::Matrix<double, 3, 1>(1e6, 1e6, 1e6), Eigen::Matrix<double, 4, 4>::Identity());
  const double dt = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
  const auto grid = map.getGrid(); // lazy mode: the one full mirror copy, outside the timed region
  map.deviceStats(st);
  std::printf("{\"api\": \"vdb_mapping::OccupancyVDBMapping::insertPointCloud (C++ shim, %s)\", \"scans\": %d, \"rays_per_sec\": %.1f, "
              "\"ms_per_scan\": %.4f, \"map_leaves\": %llu, \"host_grid_active_voxels\": %llu}\n",
              mode.c_str(), n_scans - warmup, double(rays) / dt, 1e3 * dt / (n_scans - warmup), (unsigned long long)st.map_leaves,
              (unsigned long long)grid->activeVoxelCount());
  return 0;
}
