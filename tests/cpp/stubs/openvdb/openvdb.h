// Stand-in with the API shape of <openvdb/openvdb.h> (see tests/cpp/stubs/README.md): the subset of OpenVDB's public
// interface that include/vdb_mapping uses in its VDBM_HAVE_OPENVDB branch, over a std::map of 8^3 leaves. Test
// infrastructure: it type-checks and exercises that branch where the real library is not installed.
#pragma once
#include <cmath>
#include <cstdint>
#include <cstring>
#include <iostream>
#include <map>
#include <memory>
#include <stdexcept>
#include <string>
#include <typeinfo>
#include <vector>

namespace openvdb {
using Index   = std::uint32_t;
using Index32 = std::uint32_t;
using Index64 = std::uint64_t;
using Int32   = std::int32_t;
using Real    = double;

namespace math {
template <typename T>
class Vec3
{
public:
  Vec3() : v_{T(0), T(0), T(0)} {}
  Vec3(T x, T y, T z) : v_{x, y, z} {}
  T& x() { return v_[0]; }
  T& y() { return v_[1]; }
  T& z() { return v_[2]; }
  const T& x() const { return v_[0]; }
  const T& y() const { return v_[1]; }
  const T& z() const { return v_[2]; }
  T& operator[](int i) { return v_[i]; }
  const T& operator[](int i) const { return v_[i]; }
  Vec3 operator+(const Vec3& o) const { return Vec3(v_[0] + o.v_[0], v_[1] + o.v_[1], v_[2] + o.v_[2]); }
  Vec3 operator-(const Vec3& o) const { return Vec3(v_[0] - o.v_[0], v_[1] - o.v_[1], v_[2] - o.v_[2]); }
  Vec3 operator*(T s) const { return Vec3(v_[0] * s, v_[1] * s, v_[2] * s); }
  T length() const { return std::sqrt(v_[0] * v_[0] + v_[1] * v_[1] + v_[2] * v_[2]); }
  bool operator==(const Vec3& o) const { return v_[0] == o.v_[0] && v_[1] == o.v_[1] && v_[2] == o.v_[2]; }

private:
  T v_[3];
};

class Coord
{
public:
  using ValueType = Int32;
  Coord() : c_{0, 0, 0} {}
  explicit Coord(Int32 xyz) : c_{xyz, xyz, xyz} {}
  Coord(Int32 x, Int32 y, Int32 z) : c_{x, y, z} {}
  Int32 x() const { return c_[0]; }
  Int32 y() const { return c_[1]; }
  Int32 z() const { return c_[2]; }
  Int32& x() { return c_[0]; }
  Int32& y() { return c_[1]; }
  Int32& z() { return c_[2]; }
  Int32 operator[](int i) const { return c_[i]; }
  Int32& operator[](int i) { return c_[i]; }
  const Int32* data() const { return c_; }
  Coord operator+(const Coord& o) const { return Coord(c_[0] + o.c_[0], c_[1] + o.c_[1], c_[2] + o.c_[2]); }
  Coord operator-(const Coord& o) const { return Coord(c_[0] - o.c_[0], c_[1] - o.c_[1], c_[2] - o.c_[2]); }
  Coord offsetBy(Int32 dx, Int32 dy, Int32 dz) const { return Coord(c_[0] + dx, c_[1] + dy, c_[2] + dz); }
  bool operator==(const Coord& o) const { return c_[0] == o.c_[0] && c_[1] == o.c_[1] && c_[2] == o.c_[2]; }
  bool operator!=(const Coord& o) const { return !(*this == o); }
  bool operator<(const Coord& o) const { return c_[0] != o.c_[0] ? c_[0] < o.c_[0] : (c_[1] != o.c_[1] ? c_[1] < o.c_[1] : c_[2] < o.c_[2]); }
  template <typename T>
  static Coord floor(const Vec3<T>& p) { return Coord(Int32(std::floor(p[0])), Int32(std::floor(p[1])), Int32(std::floor(p[2]))); }
  static Coord min() { return Coord(INT32_MIN); }
  static Coord max() { return Coord(INT32_MAX); }

private:
  Int32 c_[3];
};

class CoordBBox
{
public:
  CoordBBox() : min_(Coord::max()), max_(Coord::min()) {} // empty, like OpenVDB's default box
  CoordBBox(const Coord& mn, const Coord& mx) : min_(mn), max_(mx) {}
  const Coord& min() const { return min_; }
  const Coord& max() const { return max_; }
  Coord& min() { return min_; }
  Coord& max() { return max_; }
  bool empty() const { return min_[0] > max_[0] || min_[1] > max_[1] || min_[2] > max_[2]; }
  operator bool() const { return !empty(); }
  bool isInside(const Coord& p) const
  {
    for (int i = 0; i < 3; ++i)
      if (p[i] < min_[i] || p[i] > max_[i]) return false;
    return true;
  }
  bool hasOverlap(const CoordBBox& b) const
  {
    for (int i = 0; i < 3; ++i)
      if (max_[i] < b.min_[i] || min_[i] > b.max_[i]) return false;
    return true;
  }
  void expand(const Coord& p)
  {
    for (int i = 0; i < 3; ++i)
    {
      if (p[i] < min_[i]) min_[i] = p[i];
      if (p[i] > max_[i]) max_[i] = p[i];
    }
  }

private:
  Coord min_, max_;
};

template <typename VecT>
class BBox
{
public:
  BBox() = default;
  BBox(const VecT& mn, const VecT& mx) : min_(mn), max_(mx) {}
  const VecT& min() const { return min_; }
  const VecT& max() const { return max_; }

private:
  VecT min_, max_;
};

// linear (uniform scale) transform only: what createLinearTransform(voxel_size) returns
class Transform
{
public:
  using Ptr = std::shared_ptr<Transform>;
  static Ptr createLinearTransform(double voxel_size = 1.0)
  {
    Ptr t(new Transform);
    t->scale_ = voxel_size;
    t->inv_   = 1.0 / voxel_size; // ScaleMap stores the inverse and multiplies by it
    return t;
  }
  Vec3<double> voxelSize() const { return Vec3<double>(scale_, scale_, scale_); }
  Vec3<double> worldToIndex(const Vec3<double>& w) const { return Vec3<double>(w[0] * inv_, w[1] * inv_, w[2] * inv_); }
  Vec3<double> indexToWorld(const Vec3<double>& i) const { return Vec3<double>(i[0] * scale_, i[1] * scale_, i[2] * scale_); }
  Vec3<double> indexToWorld(const Coord& c) const { return Vec3<double>(c[0] * scale_, c[1] * scale_, c[2] * scale_); }

private:
  double scale_ = 1.0, inv_ = 1.0;
};
} // namespace math

using math::Coord;
using math::CoordBBox;
using Vec3d = math::Vec3<double>;
using Vec3f = math::Vec3<float>;
using Vec3R = Vec3d;
using BBoxd = math::BBox<Vec3d>;

namespace util {
template <Index Log2Dim>
class NodeMask
{
public:
  static const Index SIZE       = 1u << (3 * Log2Dim);
  static const Index WORD_COUNT = SIZE >> 6;
  using Word                    = Index64;
  NodeMask() { setOff(); }
  void setOff() { std::memset(words_, 0, sizeof(words_)); }
  void setOn() { std::memset(words_, 0xFF, sizeof(words_)); }
  bool isOn(Index n) const { return (words_[n >> 6] >> (n & 63)) & Word(1); }
  bool isOff(Index n) const { return !isOn(n); }
  void setOn(Index n) { words_[n >> 6] |= Word(1) << (n & 63); }
  void setOff(Index n) { words_[n >> 6] &= ~(Word(1) << (n & 63)); }
  void set(Index n, bool on) { on ? setOn(n) : setOff(n); }
  bool isOff() const
  {
    for (Index w = 0; w < WORD_COUNT; ++w)
      if (words_[w]) return false;
    return true;
  }
  Index countOn() const
  {
    Index c = 0;
    for (Index w = 0; w < WORD_COUNT; ++w) c += Index(__builtin_popcountll(words_[w]));
    return c;
  }
  template <typename WordT>
  WordT getWord(Index n) const
  {
    static_assert(sizeof(WordT) == sizeof(Word), "stub: 64-bit words only");
    return WordT(words_[n]);
  }
  template <typename WordT>
  WordT& getWord(Index n)
  {
    static_assert(sizeof(WordT) == sizeof(Word), "stub: 64-bit words only");
    return reinterpret_cast<WordT&>(words_[n]);
  }
  bool operator==(const NodeMask& o) const { return std::memcmp(words_, o.words_, sizeof(words_)) == 0; }

private:
  Word words_[WORD_COUNT];
};
} // namespace util

namespace tree {
// LeafBuffer<T>: a plain array; LeafBuffer<bool>: the values ARE a NodeMask (OpenVDB's bool specialisation)
template <typename T, Index Log2Dim>
class LeafBuffer
{
public:
  static const Index SIZE = 1u << (3 * Log2Dim);
  LeafBuffer() { for (Index i = 0; i < SIZE; ++i) data_[i] = T(); }
  explicit LeafBuffer(const T& v) { for (Index i = 0; i < SIZE; ++i) data_[i] = v; }
  const T& getValue(Index i) const { return data_[i]; }
  void setValue(Index i, const T& v) { data_[i] = v; }
  T* data() { return data_; }
  const T* data() const { return data_; }

private:
  alignas(32) T data_[SIZE];
};
template <Index Log2Dim>
class LeafBuffer<bool, Log2Dim>
{
public:
  using WordType = Index64;
  LeafBuffer() = default;
  explicit LeafBuffer(bool v) { if (v) mask_.setOn(); }
  bool getValue(Index i) const { return mask_.isOn(i); }
  void setValue(Index i, bool v) { mask_.set(i, v); }
  WordType getWord(Index n) const { return mask_.template getWord<WordType>(n); }
  WordType& getWord(Index n) { return mask_.template getWord<WordType>(n); }

private:
  util::NodeMask<Log2Dim> mask_;
};

template <typename T, Index Log2Dim>
class LeafNode
{
public:
  using ValueType    = T;
  using BuildType    = T;
  using Buffer       = LeafBuffer<T, Log2Dim>;
  using NodeMaskType = util::NodeMask<Log2Dim>;
  static const Index LOG2DIM = Log2Dim, DIM = 1u << Log2Dim, NUM_VALUES = 1u << (3 * Log2Dim), SIZE = NUM_VALUES;
  LeafNode() = default;
  LeafNode(const Coord& xyz, const T& background, bool active = false) : buffer_(background), origin_(xyz[0] & ~Int32(DIM - 1), xyz[1] & ~Int32(DIM - 1), xyz[2] & ~Int32(DIM - 1))
  {
    if (active) mask_.setOn();
  }
  static Index coordToOffset(const Coord& xyz)
  {
    return (Index(xyz[0] & Int32(DIM - 1)) << (2 * Log2Dim)) | (Index(xyz[1] & Int32(DIM - 1)) << Log2Dim) | Index(xyz[2] & Int32(DIM - 1));
  }
  Coord offsetToGlobalCoord(Index n) const
  {
    return Coord(origin_[0] + Int32(n >> (2 * Log2Dim)), origin_[1] + Int32((n >> Log2Dim) & (DIM - 1)), origin_[2] + Int32(n & (DIM - 1)));
  }
  const Coord& origin() const { return origin_; }
  Buffer& buffer() { return buffer_; }
  const Buffer& buffer() const { return buffer_; }
  const NodeMaskType& getValueMask() const { return mask_; }
  NodeMaskType& getValueMask() { return mask_; }
  const NodeMaskType& valueMask() const { return mask_; }
  void setValueMask(const NodeMaskType& m) { mask_ = m; }
  bool isValueOn(Index n) const { return mask_.isOn(n); }
  bool isValueOn(const Coord& xyz) const { return mask_.isOn(coordToOffset(xyz)); }
  T getValue(Index n) const { return buffer_.getValue(n); }
  T getValue(const Coord& xyz) const { return buffer_.getValue(coordToOffset(xyz)); }
  void setValueOnly(Index n, const T& v) { buffer_.setValue(n, v); }
  void setValueOnly(const Coord& xyz, const T& v) { buffer_.setValue(coordToOffset(xyz), v); }
  void setValueOn(Index n) { mask_.setOn(n); }
  void setValueOn(Index n, const T& v) { buffer_.setValue(n, v); mask_.setOn(n); }
  void setValueOn(const Coord& xyz, const T& v) { setValueOn(coordToOffset(xyz), v); }
  void setValueOff(Index n) { mask_.setOff(n); }
  void setValueOff(Index n, const T& v) { buffer_.setValue(n, v); mask_.setOff(n); }
  void setValueOff(const Coord& xyz, const T& v) { setValueOff(coordToOffset(xyz), v); }
  void setActiveState(Index n, bool on) { mask_.set(n, on); }
  void setActiveState(const Coord& xyz, bool on) { mask_.set(coordToOffset(xyz), on); }
  Index64 onVoxelCount() const { return mask_.countOn(); }
  bool isEmpty() const { return mask_.isOff(); }

private:
  Buffer buffer_;
  NodeMaskType mask_;
  Coord origin_;
};

// The stand-in tree: no internal nodes, no tiles - a sorted map leaf origin -> leaf (leaf nodes never move).
template <typename LeafT>
class LeafMapTree
{
public:
  using LeafNodeType = LeafT;
  using ValueType    = typename LeafT::ValueType;
  using BuildType    = ValueType;
  using Ptr          = std::shared_ptr<LeafMapTree>;
  using MapType      = std::map<Coord, std::unique_ptr<LeafT> >;

  class LeafCIter
  {
  public:
    LeafCIter(typename MapType::const_iterator it, typename MapType::const_iterator end) : it_(it), end_(end) {}
    operator bool() const { return it_ != end_; }
    LeafCIter& operator++() { ++it_; return *this; }
    const LeafT* operator->() const { return it_->second.get(); }
    const LeafT& operator*() const { return *it_->second; }
    const LeafT* getLeaf() const { return it_->second.get(); }

  private:
    typename MapType::const_iterator it_, end_;
  };

  explicit LeafMapTree(const ValueType& background = ValueType()) : background_(background) {}
  const ValueType& background() const { return background_; }
  static Coord leafOrigin(const Coord& xyz) { return Coord(xyz[0] & ~Int32(LeafT::DIM - 1), xyz[1] & ~Int32(LeafT::DIM - 1), xyz[2] & ~Int32(LeafT::DIM - 1)); }
  LeafT* touchLeaf(const Coord& xyz)
  {
    const Coord o = leafOrigin(xyz);
    auto it       = leaves_.find(o);
    if (it == leaves_.end()) it = leaves_.emplace(o, std::unique_ptr<LeafT>(new LeafT(o, background_, false))).first;
    return it->second.get();
  }
  LeafT* probeLeaf(const Coord& xyz)
  {
    auto it = leaves_.find(leafOrigin(xyz));
    return it == leaves_.end() ? nullptr : it->second.get();
  }
  const LeafT* probeConstLeaf(const Coord& xyz) const
  {
    auto it = leaves_.find(leafOrigin(xyz));
    return it == leaves_.end() ? nullptr : it->second.get();
  }
  const LeafT* probeLeaf(const Coord& xyz) const { return probeConstLeaf(xyz); }
  LeafCIter cbeginLeaf() const { return LeafCIter(leaves_.begin(), leaves_.end()); }
  Index32 leafCount() const { return Index32(leaves_.size()); }
  Index64 activeVoxelCount() const
  {
    Index64 n = 0;
    for (auto& kv : leaves_) n += kv.second->onVoxelCount();
    return n;
  }
  bool empty() const { return leaves_.empty(); }
  void clear() { leaves_.clear(); }
  const ValueType& getValue(const Coord& xyz) const
  {
    const LeafT* l = probeConstLeaf(xyz);
    tmp_           = l ? l->getValue(xyz) : background_;
    return tmp_;
  }
  bool isValueOn(const Coord& xyz) const
  {
    const LeafT* l = probeConstLeaf(xyz);
    return l && l->isValueOn(xyz);
  }
  // drop leaves that hold nothing but inactive background (what prune() can remove without tiles)
  void prune()
  {
    for (auto it = leaves_.begin(); it != leaves_.end();)
    {
      bool removable = it->second->isEmpty();
      for (Index n = 0; removable && n < LeafT::NUM_VALUES; ++n) removable = it->second->getValue(n) == background_;
      it = removable ? leaves_.erase(it) : std::next(it);
    }
  }
  MapType& stubLeaves() { return leaves_; } // stub-internal (io / morphology stand-ins)
  const MapType& stubLeaves() const { return leaves_; }

private:
  ValueType background_;
  MapType leaves_;
  mutable ValueType tmp_ = ValueType();
};

// Tree4<T, N1, N2, N3>::Type: the level layout above the leaves does not exist in the stand-in
template <typename T, Index N1 = 5, Index N2 = 4, Index N3 = 3>
struct Tree4
{
  using Type = LeafMapTree<LeafNode<T, N3> >;
};

template <typename TreeT>
class ValueAccessor
{
public:
  using ValueType = typename TreeT::ValueType;
  using LeafT     = typename TreeT::LeafNodeType;
  explicit ValueAccessor(TreeT& tree) : tree_(&tree) {}
  const ValueType& getValue(const Coord& xyz) const { return tree_->getValue(xyz); }
  bool isValueOn(const Coord& xyz) const { return tree_->isValueOn(xyz); }
  void setValue(const Coord& xyz, const ValueType& v) { tree_->touchLeaf(xyz)->setValueOn(xyz, v); }
  void setValueOn(const Coord& xyz, const ValueType& v) { tree_->touchLeaf(xyz)->setValueOn(xyz, v); }
  void setValueOn(const Coord& xyz) { tree_->touchLeaf(xyz)->setActiveState(xyz, true); }
  void setValueOff(const Coord& xyz, const ValueType& v) { tree_->touchLeaf(xyz)->setValueOff(xyz, v); }
  void setValueOnly(const Coord& xyz, const ValueType& v) { tree_->touchLeaf(xyz)->setValueOnly(xyz, v); }
  void setActiveState(const Coord& xyz, bool on) { tree_->touchLeaf(xyz)->setActiveState(xyz, on); }
  LeafT* touchLeaf(const Coord& xyz) { return tree_->touchLeaf(xyz); }
  const LeafT* probeConstLeaf(const Coord& xyz) const { return tree_->probeConstLeaf(xyz); }
  TreeT& tree() const { return *tree_; }

private:
  TreeT* tree_;
};
} // namespace tree

// ---- metadata ------------------------------------------------------------------------------------------------------
class Metadata
{
public:
  using Ptr = std::shared_ptr<Metadata>;
  virtual ~Metadata() = default;
  virtual Ptr copy() const = 0;
};
template <typename T>
class TypedMetadata : public Metadata
{
public:
  TypedMetadata() = default;
  explicit TypedMetadata(const T& v) : value_(v) {}
  const T& value() const { return value_; }
  Metadata::Ptr copy() const override { return Metadata::Ptr(new TypedMetadata<T>(value_)); }

private:
  T value_ = T();
};
using Vec3DMetadata  = TypedMetadata<Vec3d>;
using StringMetadata = TypedMetadata<std::string>;

class MetaMap
{
public:
  void insertMeta(const std::string& name, const Metadata& value) { meta_[name] = value.copy(); }
  template <typename T>
  const T& metaValue(const std::string& name) const
  {
    auto it = meta_.find(name);
    const TypedMetadata<T>* m = it == meta_.end() ? nullptr : dynamic_cast<const TypedMetadata<T>*>(it->second.get());
    if (!m) throw std::runtime_error("openvdb stub: no metadata named " + name); // OpenVDB throws LookupError / TypeError
    return m->value();
  }
  const std::map<std::string, Metadata::Ptr>& stubMeta() const { return meta_; }

private:
  std::map<std::string, Metadata::Ptr> meta_;
};

enum GridClass
{
  GRID_UNKNOWN = 0,
  GRID_LEVEL_SET,
  GRID_FOG_VOLUME,
  GRID_STAGGERED
};

class GridBase : public MetaMap
{
public:
  using Ptr      = std::shared_ptr<GridBase>;
  using ConstPtr = std::shared_ptr<const GridBase>;
  virtual ~GridBase() = default;
  void setName(const std::string& n) { name_ = n; }
  const std::string& getName() const { return name_; }
  void setGridClass(GridClass c) { class_ = c; }
  GridClass getGridClass() const { return class_; }
  void setTransform(math::Transform::Ptr t) { transform_ = t; }
  const math::Transform& transform() const { return *transform_; }
  math::Transform::Ptr transformPtr() { return transform_; }
  Vec3d voxelSize() const { return transform_->voxelSize(); }
  Vec3d worldToIndex(const Vec3d& w) const { return transform_->worldToIndex(w); }
  Vec3d indexToWorld(const Vec3d& i) const { return transform_->indexToWorld(i); }
  Vec3d indexToWorld(const Coord& c) const { return transform_->indexToWorld(c); }
  virtual bool empty() const                  = 0;
  virtual void clear()                        = 0;
  virtual Index64 activeVoxelCount() const    = 0;
  virtual const char* stubValueTypeTag() const = 0; // stub-internal: io stand-ins
  virtual void stubWrite(std::ostream&) const  = 0;

protected:
  std::string name_;
  GridClass class_                = GRID_UNKNOWN;
  math::Transform::Ptr transform_ = math::Transform::createLinearTransform(1.0);
};

template <typename TreeT>
class Grid : public GridBase
{
public:
  using Ptr           = std::shared_ptr<Grid>;
  using ConstPtr      = std::shared_ptr<const Grid>;
  using TreeType      = TreeT;
  using TreePtrType   = typename TreeT::Ptr;
  using ValueType     = typename TreeT::ValueType;
  using Accessor      = tree::ValueAccessor<TreeT>;
  using ConstAccessor = tree::ValueAccessor<const TreeT>;

  // iterates the active voxels in leaf order, offset order within a leaf (OpenVDB: ValueOnCIter)
  class ValueOnCIter
  {
  public:
    using MapIt = typename TreeT::MapType::const_iterator;
    ValueOnCIter(MapIt it, MapIt end) : it_(it), end_(end) { settle(); }
    operator bool() const { return it_ != end_; }
    ValueOnCIter& operator++()
    {
      ++n_;
      settle();
      return *this;
    }
    Coord getCoord() const { return it_->second->offsetToGlobalCoord(n_); }
    ValueType getValue() const { return it_->second->getValue(n_); }
    ValueType operator*() const { return getValue(); }

  private:
    void settle()
    {
      while (it_ != end_)
      {
        while (n_ < TreeT::LeafNodeType::NUM_VALUES && !it_->second->isValueOn(n_)) ++n_;
        if (n_ < TreeT::LeafNodeType::NUM_VALUES) return;
        ++it_;
        n_ = 0;
      }
    }
    MapIt it_, end_;
    Index n_ = 0;
  };

  Grid() : tree_(new TreeT()) {}
  explicit Grid(const ValueType& background) : tree_(new TreeT(background)) {}
  static Ptr create(const ValueType& background) { return Ptr(new Grid(background)); }
  static Ptr create() { return Ptr(new Grid()); }
  TreeT& tree() { return *tree_; }
  const TreeT& tree() const { return *tree_; }
  const TreeT& constTree() const { return *tree_; }
  TreePtrType treePtr() { return tree_; }
  Accessor getAccessor() { return Accessor(*tree_); }
  const ValueType& background() const { return tree_->background(); }
  bool empty() const override { return tree_->empty(); }
  void clear() override { tree_->clear(); }
  Index64 activeVoxelCount() const override { return tree_->activeVoxelCount(); }
  ValueOnCIter cbeginValueOn() const { return ValueOnCIter(tree_->stubLeaves().begin(), tree_->stubLeaves().end()); }
  CoordBBox evalActiveVoxelBoundingBox() const
  {
    CoordBBox bb;
    for (ValueOnCIter it = cbeginValueOn(); it; ++it) bb.expand(it.getCoord());
    return bb;
  }
  void pruneGrid(float /*tolerance*/ = 0.0f) { tree_->prune(); }
  const char* stubValueTypeTag() const override { return typeid(ValueType).name(); }
  void stubWrite(std::ostream& os) const override;
  static Ptr stubRead(std::istream& is);

private:
  TreePtrType tree_;
};

using FloatTree = tree::Tree4<float, 5, 4, 3>::Type;
using BoolTree  = tree::Tree4<bool, 5, 4, 3>::Type;
using FloatGrid = Grid<FloatTree>;
using BoolGrid  = Grid<BoolTree>;

using GridPtrVec    = std::vector<GridBase::Ptr>;
using GridPtrVecPtr = std::shared_ptr<GridPtrVec>;

template <typename GridType>
inline typename GridType::Ptr gridPtrCast(const GridBase::Ptr& grid)
{
  return std::dynamic_pointer_cast<GridType>(grid);
}

inline void initialize() {}
inline void uninitialize() {}

// ---- flat byte stream behind the io stand-ins (NOT the .vdb format) -----------------------------------------------------
namespace stub_io {
template <typename T>
inline void put(std::ostream& os, const T& v) { os.write(reinterpret_cast<const char*>(&v), sizeof(T)); }
template <typename T>
inline bool get(std::istream& is, T& v) { return bool(is.read(reinterpret_cast<char*>(&v), sizeof(T))); }
inline void putString(std::ostream& os, const std::string& s)
{
  put<std::uint32_t>(os, std::uint32_t(s.size()));
  os.write(s.data(), std::streamsize(s.size()));
}
inline bool getString(std::istream& is, std::string& s)
{
  std::uint32_t n = 0;
  if (!get(is, n) || n > (1u << 20)) return false;
  s.resize(n);
  return n == 0 || bool(is.read(&s[0], n));
}
} // namespace stub_io

template <typename TreeT>
void Grid<TreeT>::stubWrite(std::ostream& os) const
{
  using namespace stub_io;
  putString(os, getName());
  put<double>(os, voxelSize()[0]);
  put<std::int32_t>(os, std::int32_t(getGridClass()));
  // Vec3d metadata only (bb_min / bb_max of map sections)
  std::uint32_t n_meta = 0;
  for (auto& kv : stubMeta())
    if (dynamic_cast<const Vec3DMetadata*>(kv.second.get())) ++n_meta;
  put(os, n_meta);
  for (auto& kv : stubMeta())
    if (auto* m = dynamic_cast<const Vec3DMetadata*>(kv.second.get()))
    {
      putString(os, kv.first);
      for (int k = 0; k < 3; ++k) put<double>(os, m->value()[k]);
    }
  put<ValueType>(os, background());
  put<std::uint64_t>(os, tree_->leafCount());
  for (auto& kv : tree_->stubLeaves())
  {
    for (int k = 0; k < 3; ++k) put<std::int32_t>(os, kv.first[k]);
    for (Index w = 0; w < 8; ++w) put<Index64>(os, kv.second->getValueMask().template getWord<Index64>(w));
    for (Index n = 0; n < TreeT::LeafNodeType::NUM_VALUES; ++n) put<ValueType>(os, kv.second->getValue(n));
  }
}

template <typename TreeT>
typename Grid<TreeT>::Ptr Grid<TreeT>::stubRead(std::istream& is)
{
  using namespace stub_io;
  std::string name;
  double voxel = 1.0;
  std::int32_t cls = 0;
  std::uint32_t n_meta = 0;
  if (!getString(is, name) || !get(is, voxel) || !get(is, cls) || !get(is, n_meta)) return Ptr();
  std::vector<std::pair<std::string, Vec3d> > meta;
  for (std::uint32_t i = 0; i < n_meta; ++i)
  {
    std::string key;
    double v[3];
    if (!getString(is, key) || !get(is, v[0]) || !get(is, v[1]) || !get(is, v[2])) return Ptr();
    meta.emplace_back(key, Vec3d(v[0], v[1], v[2]));
  }
  ValueType background = ValueType();
  std::uint64_t n_leaves = 0;
  if (!get(is, background) || !get(is, n_leaves)) return Ptr();
  Ptr g = create(background);
  g->setName(name);
  g->setTransform(math::Transform::createLinearTransform(voxel));
  g->setGridClass(GridClass(cls));
  for (auto& kv : meta) g->insertMeta(kv.first, Vec3DMetadata(kv.second));
  for (std::uint64_t i = 0; i < n_leaves; ++i)
  {
    std::int32_t o[3];
    if (!get(is, o[0]) || !get(is, o[1]) || !get(is, o[2])) return Ptr();
    auto* leaf = g->tree().touchLeaf(Coord(o[0], o[1], o[2]));
    typename TreeT::LeafNodeType::NodeMaskType mask;
    for (Index w = 0; w < 8; ++w)
      if (!get(is, mask.template getWord<Index64>(w))) return Ptr();
    leaf->setValueMask(mask);
    for (Index n = 0; n < TreeT::LeafNodeType::NUM_VALUES; ++n)
    {
      ValueType v;
      if (!get(is, v)) return Ptr();
      leaf->setValueOnly(n, v);
    }
  }
  return g;
}

namespace stub_io {
inline void writeGrids(std::ostream& os, const GridPtrVec& grids)
{
  os.write("VDBSTUB1", 8);
  put<std::uint32_t>(os, std::uint32_t(grids.size()));
  for (auto& g : grids)
  {
    putString(os, g->stubValueTypeTag());
    g->stubWrite(os);
  }
}
inline GridPtrVecPtr readGrids(std::istream& is)
{
  GridPtrVecPtr out(new GridPtrVec);
  char magic[8];
  std::uint32_t n = 0;
  if (!is.read(magic, 8) || std::memcmp(magic, "VDBSTUB1", 8) != 0 || !get(is, n)) return out;
  for (std::uint32_t i = 0; i < n; ++i)
  {
    std::string tag;
    if (!getString(is, tag)) break;
    GridBase::Ptr g;
    if (tag == typeid(float).name()) g = FloatGrid::stubRead(is);
    else if (tag == typeid(bool).name()) g = BoolGrid::stubRead(is);
    if (!g) break;
    out->push_back(g);
  }
  return out;
}
} // namespace stub_io
} // namespace openvdb

#include <openvdb/io/File.h>
