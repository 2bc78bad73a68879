// Stand-in with the API shape of <pcl/io/pcd_io.h> (tests/cpp/stubs/README.md): PCD v0.7 with x y z float fields, ascii (written and read)
// and binary (read).
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
  // header, then "DATA ascii" (one "x y z" line per point) or "DATA binary" (FIELDS x y z, SIZE 4 4 4: packed float triples)
  std::ifstream f(file_name, std::ios::binary);
  if (!f) return -1;
  std::string line, data;
  std::size_t points = 0;
  cloud.points.clear();
  while (std::getline(f, line))
  {
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "POINTS") ls >> points;
    else if (key == "FIELDS")
    {
      std::string a, b, c, more;
      ls >> a >> b >> c;
      if (a != "x" || b != "y" || c != "z" || (ls >> more)) return -1; // the stand-in only knows x y z clouds
    }
    else if (key == "DATA")
    {
      ls >> data;
      break;
    }
  }
  if (data == "binary")
  {
    for (std::size_t i = 0; i < points; ++i)
    {
      float xyz[3];
      if (!f.read(reinterpret_cast<char*>(xyz), sizeof(xyz))) return -1;
      cloud.points.emplace_back(xyz[0], xyz[1], xyz[2]);
    }
  }
  else if (data == "ascii")
  {
    while (std::getline(f, line))
    {
      std::istringstream ls(line);
      PointT p;
      if (ls >> p.x >> p.y >> p.z) cloud.points.push_back(p);
    }
  }
  else return -1;
  cloud.width  = static_cast<std::uint32_t>(cloud.points.size());
  cloud.height = 1;
  return 0;
}
} // namespace io
} // namespace pcl
