// compat_types.hpp — minimal stand-ins for the third-party types that appear in the public signatures of
// vdb_mapping::VDBMapping (reference: /root/reference/include/vdb_mapping/VDBMapping.hpp:32-55,81-89), used ONLY
// when OpenVDB / PCL / Eigen are not installed (this image). With the real libraries present,
// vdb_mapping/detail/backend.hpp selects them instead and this file is not included.
//
// They implement just what the scan-integration path and the reference's tests touch:
//   pcl::PointXYZ (16-byte record), pcl::PointCloud<T>{points, Ptr, ConstPtr}
//   Eigen::Matrix<double,3,1> (x(),y(),z()), Eigen::Matrix<double,4,4> (operator()(r,c), Identity())
//   openvdb::Coord / Vec3d / CoordBBox, and a leaf-hashed host grid with OpenVDB's leaf layout and the
//   Grid::getAccessor() / Accessor::{getValue,isValueOn,setValueOn,setActiveState} subset.
#ifndef VDB_MAPPING_COMPAT_TYPES_HPP_INCLUDED
#define VDB_MAPPING_COMPAT_TYPES_HPP_INCLUDED

#include <array>
#include <cmath>
#include <cstdint>
#include <cstring>
#include <map>
#include <memory>
#include <iterator>
#include <string>
#include <unordered_map>
#include <vector>

namespace vdbm_compat {

// ---------------------------------------------------------------- pcl
namespace pcl {
struct alignas(16) PointXYZ
{
  float x, y, z, pad;
  PointXYZ() : x(0), y(0), z(0), pad(1.0f) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_), pad(1.0f) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ is a 16-byte record");

template <typename PointT>
struct PointCloud
{
  using Ptr      = std::shared_ptr<PointCloud<PointT> >;
  using ConstPtr = std::shared_ptr<const PointCloud<PointT> >;
  std::vector<PointT> points;
  std::uint32_t width = 0, height = 1;
  std::size_t size() const { return points.size(); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};
} // namespace pcl

// ---------------------------------------------------------------- Eigen
namespace Eigen {
template <typename T, int R, int C>
struct Matrix
{
  std::array<T, R * C> m{};
  Matrix() = default;
  Matrix(T a, T b, T c)
  {
    static_assert(R * C == 3, "3-vector constructor");
    m = {a, b, c};
  }
  Matrix(T a, T b, T c, T d)
  {
    static_assert(R * C == 4, "4-vector constructor");
    m = {a, b, c, d};
  }
  T& operator()(int r, int c) { return m[r * C + c]; }
  const T& operator()(int r, int c) const { return m[r * C + c]; }
  T& operator[](int i) { return m[i]; }
  const T& operator[](int i) const { return m[i]; }
  T x() const { return m[0]; }
  T y() const { return m[1]; }
  T z() const { return m[2]; }
  static Matrix Identity()
  {
    Matrix r;
    for (int i = 0; i < (R < C ? R : C); ++i) r(i, i) = T(1);
    return r;
  }
};
} // namespace Eigen

// ---------------------------------------------------------------- openvdb
namespace openvdb {
using Int32 = std::int32_t;
struct Vec3d
{
  double v[3];
  Vec3d() : v{0, 0, 0} {}
  Vec3d(double a, double b, double c) : v{a, b, c} {}
  double x() const { return v[0]; }
  double y() const { return v[1]; }
  double z() const { return v[2]; }
  double& operator[](int i) { return v[i]; }
  double operator[](int i) const { return v[i]; }
  Vec3d operator+(const Vec3d& o) const { return Vec3d(v[0] + o.v[0], v[1] + o.v[1], v[2] + o.v[2]); }
  Vec3d operator*(double s) const { return Vec3d(v[0] * s, v[1] * s, v[2] * s); }
  double length() const { return std::sqrt(v[0] * v[0] + v[1] * v[1] + v[2] * v[2]); }
};
struct Coord
{
  Int32 c[3];
  Coord() : c{0, 0, 0} {}
  Coord(Int32 x, Int32 y, Int32 z) : c{x, y, z} {}
  Int32 x() const { return c[0]; }
  Int32 y() const { return c[1]; }
  Int32 z() const { return c[2]; }
  Int32 operator[](int i) const { return c[i]; }
  Coord operator+(const Coord& o) const { return Coord(c[0] + o.c[0], c[1] + o.c[1], c[2] + o.c[2]); }
  Coord offsetBy(Int32 dx, Int32 dy, Int32 dz) const { return Coord(c[0] + dx, c[1] + dy, c[2] + dz); }
  bool operator==(const Coord& o) const { return c[0] == o.c[0] && c[1] == o.c[1] && c[2] == o.c[2]; }
  bool operator!=(const Coord& o) const { return !(*this == o); }
  bool operator<(const Coord& o) const
  {
    return c[0] != o.c[0] ? c[0] < o.c[0] : (c[1] != o.c[1] ? c[1] < o.c[1] : c[2] < o.c[2]);
  }
  static Coord floor(const Vec3d& p) { return Coord(Int32(std::floor(p[0])), Int32(std::floor(p[1])), Int32(std::floor(p[2]))); }
};
struct CoordBBox
{
  Coord mn, mx;
  CoordBBox() = default;
  CoordBBox(const Coord& a, const Coord& b) : mn(a), mx(b) {}
  const Coord& min() const { return mn; }
  const Coord& max() const { return mx; }
  bool hasOverlap(const CoordBBox& b) const
  {
    for (int i = 0; i < 3; ++i)
      if (mx[i] < b.mn[i] || mn[i] > b.mx[i]) return false;
    return true;
  }
  bool isInside(const Coord& p) const
  {
    for (int i = 0; i < 3; ++i)
      if (p[i] < mn[i] || p[i] > mx[i]) return false;
    return true;
  }
};
struct BBoxd
{
  Vec3d mn, mx;
  BBoxd() = default;
  BBoxd(const Vec3d& a, const Vec3d& b) : mn(a), mx(b) {}
  const Vec3d& min() const { return mn; }
  const Vec3d& max() const { return mx; }
};
inline void initialize() {}

// One 8^3 leaf with OpenVDB's layout: offset n = (x&7)<<6 | (y&7)<<3 | (z&7), mask word n>>6, bit n&63.
template <typename ValueT>
struct HostLeaf;
template <>
struct HostLeaf<float>
{
  alignas(64) float values[512]; // cache-line aligned like an OpenVDB leaf buffer: the mirror update streams into it
  std::uint64_t active[8];
  HostLeaf()
  {
    std::memset(values, 0, sizeof(values));
    std::memset(active, 0, sizeof(active));
  }
  float get(unsigned n) const { return values[n]; }
  void set(unsigned n, float v) { values[n] = v; }
};
template <>
struct HostLeaf<bool>
{
  std::uint64_t valmask[8];
  std::uint64_t active[8];
  HostLeaf()
  {
    std::memset(valmask, 0, sizeof(valmask));
    std::memset(active, 0, sizeof(active));
  }
  bool get(unsigned n) const { return (valmask[n >> 6] >> (n & 63)) & 1u; }
  void set(unsigned n, bool v)
  {
    if (v) valmask[n >> 6] |= std::uint64_t(1) << (n & 63);
    else valmask[n >> 6] &= ~(std::uint64_t(1) << (n & 63));
  }
};

// Host-side sparse grid: std::map of leaves keyed by leaf origin (ordered like OpenVDB's root table).
template <typename ValueT>
class HostGrid
{
public:
  using Ptr       = std::shared_ptr<HostGrid<ValueT> >;
  using ConstPtr  = std::shared_ptr<const HostGrid<ValueT> >;
  using ValueType = ValueT;
  using LeafT     = HostLeaf<ValueT>;
  using LeafMap   = std::map<Coord, LeafT>;

  class Accessor
  {
  public:
    explicit Accessor(HostGrid* g) : m_grid(g) {}
    ValueT getValue(const Coord& xyz) const
    {
      const LeafT* l = m_grid->probeLeaf(xyz);
      return l ? l->get(offset(xyz)) : m_grid->background();
    }
    bool isValueOn(const Coord& xyz) const
    {
      const LeafT* l = m_grid->probeLeaf(xyz);
      if (!l) return false;
      const unsigned n = offset(xyz);
      return (l->active[n >> 6] >> (n & 63)) & 1u;
    }
    void setValueOn(const Coord& xyz, const ValueT& v)
    {
      LeafT& l         = m_grid->touchLeaf(xyz);
      const unsigned n = offset(xyz);
      l.set(n, v);
      l.active[n >> 6] |= std::uint64_t(1) << (n & 63);
    }
    void setValueOff(const Coord& xyz, const ValueT& v)
    {
      LeafT& l         = m_grid->touchLeaf(xyz);
      const unsigned n = offset(xyz);
      l.set(n, v);
      l.active[n >> 6] &= ~(std::uint64_t(1) << (n & 63));
    }
    void setValueOnly(const Coord& xyz, const ValueT& v) { m_grid->touchLeaf(xyz).set(offset(xyz), v); }
    void setActiveState(const Coord& xyz, bool on)
    {
      LeafT& l         = m_grid->touchLeaf(xyz);
      const unsigned n = offset(xyz);
      if (on) l.active[n >> 6] |= std::uint64_t(1) << (n & 63);
      else l.active[n >> 6] &= ~(std::uint64_t(1) << (n & 63));
    }
    HostGrid* grid() const { return m_grid; }

  private:
    static unsigned offset(const Coord& c) { return (unsigned(c[0] & 7) << 6) | (unsigned(c[1] & 7) << 3) | unsigned(c[2] & 7); }
    HostGrid* m_grid;
  };

  explicit HostGrid(const ValueT& background = ValueT()) : m_background(background) {}
  static Ptr create(const ValueT& background = ValueT()) { return std::make_shared<HostGrid<ValueT> >(background); }
  Accessor getAccessor() { return Accessor(this); }
  const ValueT& background() const { return m_background; }
  bool empty() const { return m_leaves.empty(); }
  void clear()
  {
    m_leaves.clear();
    ++m_epoch;
  }
  // changes whenever leaves were REMOVED (pointers to leaves handed out earlier are void)
  std::uint64_t structureEpoch() const { return m_epoch; }
  std::size_t leafCount() const { return m_leaves.size(); }
  std::uint64_t activeVoxelCount() const
  {
    std::uint64_t n = 0;
    for (auto& kv : m_leaves)
      for (int w = 0; w < 8; ++w) n += std::uint64_t(__builtin_popcountll(kv.second.active[w]));
    return n;
  }
  void setVoxelSize(double s) { m_voxel_size = s; }
  double voxelSize() const { return m_voxel_size; }
  Vec3d worldToIndex(const Vec3d& w) const
  {
    const double inv = 1.0 / m_voxel_size; // ScaleMap::applyInverseMap multiplies by the stored inverse
    return Vec3d(w[0] * inv, w[1] * inv, w[2] * inv);
  }
  Vec3d indexToWorld(const Coord& c) const { return Vec3d(c[0] * m_voxel_size, c[1] * m_voxel_size, c[2] * m_voxel_size); }
  void insertMeta(const std::string& name, const Vec3d& v) { m_meta[name] = v; }
  Vec3d metaValue(const std::string& name) const
  {
    auto it = m_meta.find(name);
    return it == m_meta.end() ? Vec3d() : it->second;
  }

  const LeafT* probeLeaf(const Coord& xyz) const
  {
    auto it = m_leaves.find(Coord(xyz[0] & ~7, xyz[1] & ~7, xyz[2] & ~7));
    return it == m_leaves.end() ? nullptr : &it->second;
  }
  LeafT& touchLeaf(const Coord& xyz) { return m_leaves[Coord(xyz[0] & ~7, xyz[1] & ~7, xyz[2] & ~7)]; }
  // touchLeaf for many leaf origins at once. Exports arrive sorted by origin (x, then y, then z), the map's own order, so
  // one merge walk over the map replaces n logarithmic look-ups; unsorted input still works (falls back to lower_bound).
  void touchLeaves(std::uint64_t n, const std::int32_t* origins, std::vector<LeafT*>& out)
  {
    out.resize(n);
    auto it = m_leaves.begin();
    for (std::uint64_t i = 0; i < n; ++i)
    {
      const Coord key(origins[3 * i] & ~7, origins[3 * i + 1] & ~7, origins[3 * i + 2] & ~7);
      if (it != m_leaves.begin() && !(std::prev(it)->first < key)) it = m_leaves.lower_bound(key); // went backwards
      while (it != m_leaves.end() && it->first < key) ++it;
      if (it == m_leaves.end() || key < it->first) it = m_leaves.emplace_hint(it, key, LeafT());
      out[i] = &it->second;
    }
  }
  const LeafMap& leaves() const { return m_leaves; }
  LeafMap& leaves() { return m_leaves; }

private:
  ValueT m_background;
  double m_voxel_size = 1.0;
  LeafMap m_leaves;
  std::uint64_t m_epoch = 0;
  std::map<std::string, Vec3d> m_meta;
};

using FloatGrid = HostGrid<float>;
using BoolGrid  = HostGrid<bool>;
} // namespace openvdb

} // namespace vdbm_compat

#endif
