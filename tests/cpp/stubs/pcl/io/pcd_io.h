// Stand-in with the API shape of <pcl/io/pcd_io.h> (tests/cpp/stubs/README.md): ASCII PCD v0.7, x y z fields only.
#pragma once
#include <fstream>
#include <sstream>
#include <string>
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
namespace pcl {
namespace io {
template <typename PointT>
int savePCDFile(const std::string& file_name, const PointCloud<PointT>& cloud, bool /*binary_mode*/ = false)
{
  std::ofstream f(file_name);
  if (!f) return -1;
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH " << cloud.points.size()
    << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << cloud.points.size() << "\nDATA ascii\n";
  f.precision(9);
  for (const auto& p : cloud.points) f << p.x << ' ' << p.y << ' ' << p.z << '\n';
  return f ? 0 : -1;
}
template <typename PointT>
int loadPCDFile(const std::string& file_name, PointCloud<PointT>& cloud)
{
  std::ifstream f(file_name);
  if (!f) return -1;
  std::string line;
  bool data = false;
  cloud.points.clear();
  while (std::getline(f, line))
  {
    if (!data)
    {
      if (line.rfind("DATA", 0) == 0)
      {
        if (line.find("ascii") == std::string::npos) return -1;
        data = true;
      }
      continue;
    }
    std::istringstream ls(line);
    PointT p;
    if (ls >> p.x >> p.y >> p.z) cloud.points.push_back(p);
  }
  cloud.width  = static_cast<std::uint32_t>(cloud.points.size());
  cloud.height = 1;
  return data ? 0 : -1;
}
} // namespace io
} // namespace pcl
