// host_io.hpp — host-side helpers behind the persistence / codec members of the drop-in VDBMapping
// (reference: /root/reference/include/vdb_mapping/VDBMapping.hpp, cited as R:<line>):
//   * zstd through the system's libzstd.so, bound at run time (no zstd.h in the build image; with the library absent the
//     codec degrades exactly like the reference does on a zstd error: the bytes travel uncompressed, R:1258-1263);
//   * a minimal PCD (v0.7) reader / writer for pcl::PointXYZ clouds when PCL itself is not installed;
//   * a flat leaf stream for grids when OpenVDB's io is not installed (NOT the .vdb format: files and byte arrays
//     written by the stand-in backend can only be read by the stand-in backend; with OpenVDB present backend.hpp uses
//     openvdb::io::File / io::Stream and the bytes are the reference's).
// None of this is on the scan-integration path.
#ifndef VDB_MAPPING_DETAIL_HOST_IO_HPP_INCLUDED
#define VDB_MAPPING_DETAIL_HOST_IO_HPP_INCLUDED

#include <cstdint>
#include <cstdio>
#include <cstring>
#include <dlfcn.h>
#include <fstream>
#include <iostream>
#include <sstream>
#include <string>
#include <vector>

namespace vdb_mapping {
namespace detail {

// ---- zstd, bound lazily ------------------------------------------------------------------------------------------
struct Zstd
{
  using compress_bound_t = std::size_t (*)(std::size_t);
  using compress_t       = std::size_t (*)(void*, std::size_t, const void*, std::size_t, int);
  using decompress_t     = std::size_t (*)(void*, std::size_t, const void*, std::size_t);
  using is_error_t       = unsigned (*)(std::size_t);
  using error_name_t     = const char* (*)(std::size_t);
  using content_size_t   = unsigned long long (*)(const void*, std::size_t);
  compress_bound_t compressBound = nullptr;
  compress_t compress            = nullptr;
  decompress_t decompress        = nullptr;
  is_error_t isError             = nullptr;
  error_name_t getErrorName      = nullptr;
  content_size_t getDecompressedSize = nullptr;
  bool ok = false;

  static const Zstd& get()
  {
    static const Zstd z = [] {
      Zstd r;
      void* h = nullptr;
      for (const char* name : {"libzstd.so.1", "libzstd.so"})
        if ((h = dlopen(name, RTLD_NOW | RTLD_GLOBAL))) break;
      if (!h) return r;
      r.compressBound       = reinterpret_cast<compress_bound_t>(dlsym(h, "ZSTD_compressBound"));
      r.compress            = reinterpret_cast<compress_t>(dlsym(h, "ZSTD_compress"));
      r.decompress          = reinterpret_cast<decompress_t>(dlsym(h, "ZSTD_decompress"));
      r.isError             = reinterpret_cast<is_error_t>(dlsym(h, "ZSTD_isError"));
      r.getErrorName        = reinterpret_cast<error_name_t>(dlsym(h, "ZSTD_getErrorName"));
      r.getDecompressedSize = reinterpret_cast<content_size_t>(dlsym(h, "ZSTD_getDecompressedSize"));
      r.ok = r.compressBound && r.compress && r.decompress && r.isError && r.getErrorName && r.getDecompressedSize;
      return r;
    }();
    return z;
  }
};

/*! R:1246-1268 */
inline std::vector<std::uint8_t> zstdCompress(const std::string& string, int level)
{
  std::vector<std::uint8_t> uncompressed(string.begin(), string.end());
  const Zstd& z = Zstd::get();
  if (!z.ok)
  {
    std::cerr << "Compression using ZSTD failed: libzstd not available , sending uncompressed byte array" << std::endl;
    return uncompressed;
  }
  const std::size_t len = z.compressBound(uncompressed.size());
  std::vector<std::uint8_t> compressed(len);
  const std::size_t ret = z.compress(compressed.data(), len, uncompressed.data(), uncompressed.size(), level);
  if (z.isError(ret))
  {
    std::cerr << "Compression using ZSTD failed: " << z.getErrorName(ret) << " , sending uncompressed byte array" << std::endl;
    return uncompressed;
  }
  compressed.resize(ret);
  return compressed;
}

/*! R:1277-1300 */
inline std::string zstdDecompress(const std::vector<std::uint8_t>& byte_array)
{
  const Zstd& z = Zstd::get();
  if (z.ok)
  {
    const unsigned long long len = z.getDecompressedSize(byte_array.data(), byte_array.size());
    std::vector<std::uint8_t> uncompressed(len);
    const std::size_t size = z.decompress(uncompressed.data(), len, byte_array.data(), byte_array.size());
    if (!z.isError(size)) return std::string(uncompressed.begin(), uncompressed.begin() + size);
    std::cerr << "Could not decompress map using ZSTD failed: " << z.getErrorName(size) << " , returning raw data" << std::endl;
  }
  return std::string(byte_array.begin(), byte_array.end());
}

// ---- PCD v0.7, x y z float32 ---------------------------------------------------------------------------------------
template <typename CloudT>
inline bool writePCD(const std::string& path, const CloudT& cloud)
{
  std::ofstream f(path, std::ios::binary);
  if (!f) return false;
  const std::size_t n = cloud.points.size();
  f << "# .PCD v0.7 - Point Cloud Data file format\nVERSION 0.7\nFIELDS x y z\nSIZE 4 4 4\nTYPE F F F\nCOUNT 1 1 1\nWIDTH " << n
    << "\nHEIGHT 1\nVIEWPOINT 0 0 0 1 0 0 0\nPOINTS " << n << "\nDATA binary\n";
  for (const auto& p : cloud.points)
  {
    const float xyz[3] = {p.x, p.y, p.z};
    f.write(reinterpret_cast<const char*>(xyz), sizeof(xyz));
  }
  return bool(f);
}

template <typename CloudT>
inline bool readPCD(const std::string& path, CloudT& cloud)
{
  std::ifstream f(path, std::ios::binary);
  if (!f) return false;
  std::string line, data;
  std::vector<std::string> fields;
  std::vector<int> sizes;
  std::size_t points = 0;
  while (std::getline(f, line))
  {
    std::istringstream ls(line);
    std::string key;
    ls >> key;
    if (key == "FIELDS") { std::string t; while (ls >> t) fields.push_back(t); }
    else if (key == "SIZE") { int t; while (ls >> t) sizes.push_back(t); }
    else if (key == "POINTS") ls >> points;
    else if (key == "DATA") { ls >> data; break; }
  }
  int off[3] = {-1, -1, -1}, stride = 0;
  for (std::size_t i = 0; i < fields.size() && i < sizes.size(); ++i)
  {
    for (int a = 0; a < 3; ++a)
      if (fields[i] == std::string(1, char('x' + a)) && sizes[i] == 4) off[a] = stride;
    stride += sizes[i];
  }
  if (off[0] < 0 || off[1] < 0 || off[2] < 0) return false;
  cloud.points.clear();
  cloud.points.reserve(points);
  if (data == "ascii")
  {
    for (std::size_t i = 0; i < points && std::getline(f, line); ++i)
    {
      std::istringstream ls(line);
      std::vector<float> v;
      float t;
      while (ls >> t) v.push_back(t);
      if (v.size() < fields.size()) return false;
      float xyz[3] = {0, 0, 0};
      for (std::size_t k = 0; k < fields.size(); ++k)
        for (int a = 0; a < 3; ++a)
          if (fields[k] == std::string(1, char('x' + a))) xyz[a] = v[k];
      cloud.points.emplace_back(xyz[0], xyz[1], xyz[2]);
    }
  }
  else if (data == "binary")
  {
    std::vector<char> rec(static_cast<std::size_t>(stride));
    for (std::size_t i = 0; i < points && f.read(rec.data(), stride); ++i)
    {
      float xyz[3];
      for (int a = 0; a < 3; ++a) std::memcpy(&xyz[a], rec.data() + off[a], 4);
      cloud.points.emplace_back(xyz[0], xyz[1], xyz[2]);
    }
  }
  else return false; // binary_compressed: needs PCL
  cloud.width  = static_cast<std::uint32_t>(cloud.points.size());
  cloud.height = 1;
  return true;
}

// ---- flat leaf stream (stand-in backend only) ------------------------------------------------------------------------
// "VDBMLEAF" | u32 kind (0 float, 1 bool) | f64 voxel size | u64 n | f64 bb_min[3], bb_max[3] | n x { i32 origin[3], payload }
// payload: float grid 512 x f32 + 8 x u64 active; bool grid 8 x u64 active + 8 x u64 values
template <typename GridT, typename ForEachLeaf>
inline std::string leafStreamWrite(const GridT& grid, std::uint32_t kind, std::uint64_t n_leaves, ForEachLeaf&& for_each_leaf)
{
  std::string out("VDBMLEAF");
  auto put = [&](const void* p, std::size_t n) { out.append(static_cast<const char*>(p), n); };
  const double vs = grid.voxelSize();
  put(&kind, 4); put(&vs, 8); put(&n_leaves, 8);
  const auto mn = grid.metaValue("bb_min"), mx = grid.metaValue("bb_max");
  for (int k = 0; k < 3; ++k) { const double v = mn[k]; put(&v, 8); }
  for (int k = 0; k < 3; ++k) { const double v = mx[k]; put(&v, 8); }
  for_each_leaf(put);
  return out;
}

} // namespace detail
} // namespace vdb_mapping

#endif
