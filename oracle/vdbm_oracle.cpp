// vdbm_oracle.cpp — CPU ORACLE for the scan-integration hot path of vdb_mapping.
//
// *** TEST INFRASTRUCTURE ONLY. *** Nothing under vdb_mapping_b200/ (the product) may include,
// link, import or execute this file. Only tests/, __graft_entry__.smoke() and bench.py's
// cpu_baseline / --impl reference legs use it, as the checker and as the timed CPU baseline.
//
// What it is: an OpenVDB-free restatement of the reference algorithm
//   /root/reference/include/vdb_mapping/VDBMapping.hpp          (cited below as V:<line>)
//   /root/reference/include/vdb_mapping/OccupancyVDBMapping.hpp  (cited below as O:<line>)
// The arithmetic the reference delegates to OpenVDB (un-vendored dependency, unpinned,
// README.md:10-13 says ">= 8.3, recommended v9.0.0") is restated from OpenVDB's published
// algorithm: math/DDA.h (DDA<RayT,0>::init/step), math/Math.h (MinIndex), math/Ray.h,
// math/Maps.h (ScaleMap::applyInverseMap), math/Coord.h (Coord::floor), math/Vec3.h
// (length/unit), tree/LeafNode.h + LeafNodeBool.h (offset/mask layout), tree/InternalNode.h +
// tree/RootNode.h (modifyValueAndActiveStateAndCache incl. the inverted-state tile probe),
// tree/ValueAccessor.h (3-level node cache), tree/TreeIterator.h (value-on order).
//
// Parity pinning: PINNED on the five known-answer tests of /root/reference/tests/mapping.cpp
// (tests/test_oracle_kat.py, tests/golden/mapping_kats.json). Everything those tests do not
// cover (diagonal tie-breaking, multi-ray dedup, change grid, sections) is "PARITY UNPINNED at
// the OpenVDB boundary": the reference itself cannot be compiled in this image (no OpenVDB /
// PCL / Eigen headers, no network), so oracle/_ref cannot be built. See DESIGN.md section 3.
//
// Storage mirrors OpenVDB's cost model on purpose (so that it doubles as the timed CPU
// baseline): map = Root(std::map) -> Internal<5> -> Internal<4> -> Leaf<3> (float),
// update grid = Root -> Internal<1> -> Internal<4> -> Leaf<3> (bool), both accessed through a
// ValueAccessor3-style cache of the last leaf / internal nodes.
//
// Build: see oracle/Makefile (-O3 -DNDEBUG -ffp-contract=off, no -march=native: the x86-64
// baseline has no FMA, which is what a distro/ROS Release build of the reference executes).

#include <algorithm>
#include <cfloat>
#include <climits>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <map>
#include <memory>
#include <string>
#include <vector>

namespace vo {

// --------------------------------------------------------------------------------------------
// Coord (openvdb::math::Coord): 3 x int32, lexicographic x,y,z ordering (RootNode key order).
// --------------------------------------------------------------------------------------------
struct Coord
{
  int32_t v[3];
  Coord() : v{0, 0, 0} {}
  Coord(int32_t x, int32_t y, int32_t z) : v{x, y, z} {}
  int32_t operator[](int i) const { return v[i]; }
  int32_t& operator[](int i) { return v[i]; }
  bool operator==(const Coord& o) const { return v[0] == o.v[0] && v[1] == o.v[1] && v[2] == o.v[2]; }
  bool operator!=(const Coord& o) const { return !(*this == o); }
  bool operator<(const Coord& o) const
  {
    if (v[0] != o.v[0]) return v[0] < o.v[0];
    if (v[1] != o.v[1]) return v[1] < o.v[1];
    return v[2] < o.v[2];
  }
  Coord masked(int32_t m) const { return Coord(v[0] & m, v[1] & m, v[2] & m); }
};

// --------------------------------------------------------------------------------------------
// NodeMask<Log2Dim> (openvdb::util::NodeMask): SIZE bits in 64-bit words, word n>>6, bit n&63.
// --------------------------------------------------------------------------------------------
template <int LOG2DIM>
struct NodeMask
{
  static constexpr uint32_t SIZE  = 1u << (3 * LOG2DIM);
  static constexpr uint32_t WORDS = SIZE >= 64 ? SIZE / 64 : 1;
  uint64_t w[WORDS];
  NodeMask() { std::memset(w, 0, sizeof(w)); }
  bool isOn(uint32_t n) const { return (w[n >> 6] >> (n & 63)) & 1u; }
  void setOn(uint32_t n) { w[n >> 6] |= uint64_t(1) << (n & 63); }
  void setOff(uint32_t n) { w[n >> 6] &= ~(uint64_t(1) << (n & 63)); }
  void set(uint32_t n, bool on) { on ? setOn(n) : setOff(n); }
  bool isOff() const
  {
    for (uint32_t i = 0; i < WORDS; ++i)
      if (w[i]) return false;
    return true;
  }
  // first set bit at position >= start, or SIZE
  uint32_t findNextOn(uint32_t start) const
  {
    if (start >= SIZE) return SIZE;
    uint32_t wi = start >> 6;
    uint64_t b  = w[wi] & (~uint64_t(0) << (start & 63));
    while (true)
    {
      if (b) return (wi << 6) + uint32_t(__builtin_ctzll(b));
      if (++wi >= WORDS) return SIZE;
      b = w[wi];
    }
  }
};

// --------------------------------------------------------------------------------------------
// Leaf nodes. Offset n = ((x&7)<<6)|((y&7)<<3)|(z&7)  (LeafNode::coordToOffset).
// --------------------------------------------------------------------------------------------
inline uint32_t leafOffset(const Coord& c)
{
  return (uint32_t(c[0] & 7) << 6) | (uint32_t(c[1] & 7) << 3) | uint32_t(c[2] & 7);
}
inline Coord leafOffsetToLocal(uint32_t n) { return Coord(int32_t(n >> 6), int32_t((n >> 3) & 7), int32_t(n & 7)); }

struct FloatLeaf
{
  static constexpr int TOTAL = 3;
  using ValueT               = float;
  Coord origin;
  NodeMask<3> vmask;
  float buf[512];
  FloatLeaf(const Coord& xyz, float bg, bool active) : origin(xyz.masked(~7))
  {
    for (int i = 0; i < 512; ++i) buf[i] = bg;
    if (active)
      for (auto& x : vmask.w) x = ~uint64_t(0);
  }
};

struct BoolLeaf
{
  static constexpr int TOTAL = 3;
  using ValueT               = bool;
  Coord origin;
  NodeMask<3> vmask; // active states
  NodeMask<3> buf;   // bool values (LeafNode<bool,3> stores values as a second bit mask)
  BoolLeaf(const Coord& xyz, bool bg, bool active) : origin(xyz.masked(~7))
  {
    if (bg)
      for (auto& x : buf.w) x = ~uint64_t(0);
    if (active)
      for (auto& x : vmask.w) x = ~uint64_t(0);
  }
};

// --------------------------------------------------------------------------------------------
// InternalNode<Child, Log2Dim> (openvdb::tree::InternalNode): child mask + value mask + a
// union table of child pointers / tile values.
// --------------------------------------------------------------------------------------------
template <typename ChildT, int LOG2DIM>
struct Internal
{
  using ValueT                    = typename ChildT::ValueT;
  using Child                     = ChildT;
  static constexpr int TOTAL      = LOG2DIM + ChildT::TOTAL;
  static constexpr uint32_t NUM   = 1u << (3 * LOG2DIM);
  static constexpr int32_t DIMM1  = (int32_t(1) << TOTAL) - 1;
  union Slot
  {
    ChildT* child;
    ValueT tile;
  };
  Coord origin;
  NodeMask<LOG2DIM> childMask;
  NodeMask<LOG2DIM> valueMask;
  Slot nodes[NUM];

  Internal(const Coord& xyz, ValueT bg, bool active) : origin(xyz.masked(~DIMM1))
  {
    for (uint32_t i = 0; i < NUM; ++i) nodes[i].tile = bg;
    if (active)
      for (auto& x : valueMask.w) x = ~uint64_t(0);
  }
  ~Internal()
  {
    for (uint32_t n = childMask.findNextOn(0); n < NUM; n = childMask.findNextOn(n + 1)) delete nodes[n].child;
  }
  Internal(const Internal&)            = delete;
  Internal& operator=(const Internal&) = delete;

  static uint32_t coordToOffset(const Coord& c)
  {
    return (uint32_t((c[0] & DIMM1) >> ChildT::TOTAL) << (2 * LOG2DIM)) |
           (uint32_t((c[1] & DIMM1) >> ChildT::TOTAL) << LOG2DIM) | uint32_t((c[2] & DIMM1) >> ChildT::TOTAL);
  }
};

using FloatI1 = Internal<FloatLeaf, 4>;  // 16^3 leaves  (128^3 voxels)
using FloatI2 = Internal<FloatI1, 5>;    // 32^3 of those (4096^3 voxels)  -> Tree4<float,5,4,3>, V:88
using BoolI1  = Internal<BoolLeaf, 4>;   // 16^3 leaves
using BoolI2  = Internal<BoolI1, 1>;     // 2^3 of those (256^3 voxels)    -> Tree4<bool,1,4,3>, V:89

// --------------------------------------------------------------------------------------------
// Tree = RootNode(std::map keyed by child origin) + background.
// --------------------------------------------------------------------------------------------
template <typename I2T>
struct Tree
{
  using I2    = I2T;
  using I1    = typename I2T::Child;
  using Leaf  = typename I1::Child;
  using Value = typename Leaf::ValueT;
  std::map<Coord, I2*> table; // RootNode::mTable: only children on this path (no root tiles)
  Value background;
  explicit Tree(Value bg) : background(bg) {}
  ~Tree() { clear(); }
  Tree(const Tree&)            = delete;
  Tree& operator=(const Tree&) = delete;
  void clear()
  {
    for (auto& kv : table) delete kv.second;
    table.clear();
  }
  bool empty() const { return table.empty(); }
  static Coord rootKey(const Coord& c) { return c.masked(~I2::DIMM1); }

  template <typename F>
  void forEachLeaf(F&& f) const // cbeginLeaf order: root key order, then offset order per level
  {
    for (auto& kv : table)
    {
      const I2* n2 = kv.second;
      for (uint32_t a = n2->childMask.findNextOn(0); a < I2::NUM; a = n2->childMask.findNextOn(a + 1))
      {
        const I1* n1 = n2->nodes[a].child;
        for (uint32_t b = n1->childMask.findNextOn(0); b < I1::NUM; b = n1->childMask.findNextOn(b + 1))
          f(*n1->nodes[b].child);
      }
    }
  }
  size_t leafCount() const
  {
    size_t n = 0;
    forEachLeaf([&](const Leaf&) { ++n; });
    return n;
  }
};

using FloatTree = Tree<FloatI2>;
using BoolTree  = Tree<BoolI2>;

// --------------------------------------------------------------------------------------------
// ValueAccessor3-style accessor: caches the last leaf, I1 and I2 node (keys = node origins).
// --------------------------------------------------------------------------------------------
template <typename TreeT>
struct Accessor
{
  using I2   = typename TreeT::I2;
  using I1   = typename TreeT::I1;
  using Leaf = typename TreeT::Leaf;
  using V    = typename TreeT::Value;
  TreeT* tree;
  Coord k0, k1, k2;
  Leaf* n0 = nullptr;
  I1* n1   = nullptr;
  I2* n2   = nullptr;
  explicit Accessor(TreeT& t) : tree(&t) {}

  bool hashed0(const Coord& c) const { return n0 && (c[0] & ~7) == k0[0] && (c[1] & ~7) == k0[1] && (c[2] & ~7) == k0[2]; }
  bool hashed1(const Coord& c) const
  {
    return n1 && (c[0] & ~I1::DIMM1) == k1[0] && (c[1] & ~I1::DIMM1) == k1[1] && (c[2] & ~I1::DIMM1) == k1[2];
  }
  bool hashed2(const Coord& c) const
  {
    return n2 && (c[0] & ~I2::DIMM1) == k2[0] && (c[1] & ~I2::DIMM1) == k2[1] && (c[2] & ~I2::DIMM1) == k2[2];
  }
  void insert(const Coord& c, Leaf* l) { n0 = l; k0 = c.masked(~7); }
  void insert(const Coord& c, I1* n) { n1 = n; k1 = c.masked(~I1::DIMM1); }
  void insert(const Coord& c, I2* n) { n2 = n; k2 = c.masked(~I2::DIMM1); }

  // ---- read path (probe without allocation) ----
  const Leaf* probeLeaf(const Coord& c)
  {
    if (hashed0(c)) return n0;
    I1* a = nullptr;
    if (hashed1(c)) a = n1;
    else
    {
      I2* b = nullptr;
      if (hashed2(c)) b = n2;
      else
      {
        auto it = tree->table.find(TreeT::rootKey(c));
        if (it == tree->table.end()) return nullptr;
        b = it->second;
        insert(c, b);
      }
      uint32_t n = I2::coordToOffset(c);
      if (!b->childMask.isOn(n)) return nullptr;
      a = b->nodes[n].child;
      insert(c, a);
    }
    uint32_t n = I1::coordToOffset(c);
    if (!a->childMask.isOn(n)) return nullptr;
    Leaf* l = a->nodes[n].child;
    insert(c, l);
    return l;
  }

  // probeConstNode<NodeT> of ValueAccessor3 for the two internal levels, and isValueOn (voxel flag, or the flag of the tile
  // the descent ends at; a root miss is the inactive background)
  I2* probeI2(const Coord& c)
  {
    if (hashed2(c)) return n2;
    auto it = tree->table.find(TreeT::rootKey(c));
    if (it == tree->table.end()) return nullptr;
    insert(c, it->second);
    return it->second;
  }
  I1* probeI1(const Coord& c)
  {
    if (hashed1(c)) return n1;
    I2* b = probeI2(c);
    if (!b) return nullptr;
    const uint32_t n = I2::coordToOffset(c);
    if (!b->childMask.isOn(n)) return nullptr;
    insert(c, b->nodes[n].child);
    return b->nodes[n].child;
  }
  bool isValueOn(const Coord& c)
  {
    if (hashed0(c)) return n0->vmask.isOn(leafOffset(c));
    I2* b = probeI2(c);
    if (!b) return false;
    const uint32_t n = I2::coordToOffset(c);
    if (!b->childMask.isOn(n)) return b->valueMask.isOn(n);
    I1* a = b->nodes[n].child;
    insert(c, a);
    const uint32_t k = I1::coordToOffset(c);
    if (!a->childMask.isOn(k)) return a->valueMask.isOn(k);
    Leaf* l = a->nodes[k].child;
    insert(c, l);
    return l->vmask.isOn(leafOffset(c));
  }

  // ---- touch path used by setActiveState(true)/setValueOn on the bool tree: tiles are always
  //      (background, inactive) here, and the requested state is "on", so a child is always
  //      created (InternalNode::setActiveStateAndCache / setValueAndCache). ----
  Leaf* touchLeaf(const Coord& c)
  {
    if (hashed0(c)) return n0;
    I1* a = nullptr;
    if (hashed1(c)) a = n1;
    else
    {
      I2* b = nullptr;
      if (hashed2(c)) b = n2;
      else
      {
        Coord key = TreeT::rootKey(c);
        auto it   = tree->table.find(key);
        if (it == tree->table.end())
        {
          b = new I2(c, tree->background, false);
          tree->table[key] = b;
        }
        else b = it->second;
        insert(c, b);
      }
      uint32_t n = I2::coordToOffset(c);
      if (!b->childMask.isOn(n))
      {
        V tile = b->nodes[n].tile;
        bool on = b->valueMask.isOn(n);
        b->nodes[n].child = new I1(c, tile, on);
        b->childMask.setOn(n);
        b->valueMask.setOff(n);
      }
      a = b->nodes[n].child;
      insert(c, a);
    }
    uint32_t n = I1::coordToOffset(c);
    if (!a->childMask.isOn(n))
    {
      V tile = a->nodes[n].tile;
      bool on = a->valueMask.isOn(n);
      a->nodes[n].child = new Leaf(c, tile, on);
      a->childMask.setOn(n);
      a->valueMask.setOff(n);
    }
    Leaf* l = a->nodes[n].child;
    insert(c, l);
    return l;
  }
};

// bool-tree write ops used by the raycast (V:535, V:563)
inline void setActiveStateOn(Accessor<BoolTree>& acc, const Coord& c) { acc.touchLeaf(c)->vmask.setOn(leafOffset(c)); }
// GridT::Accessor::setActiveState(xyz, true) on the float map (V:788, V:1082): tiles are inactive background, so a
// child is created down to the leaf, the value is left alone
inline void setActiveStateOn(Accessor<FloatTree>& acc, const Coord& c) { acc.touchLeaf(c)->vmask.setOn(leafOffset(c)); }
inline void setValueOnTrue(Accessor<BoolTree>& acc, const Coord& c)
{
  BoolLeaf* l = acc.touchLeaf(c);
  uint32_t n  = leafOffset(c);
  l->buf.setOn(n);
  l->vmask.setOn(n);
}

// --------------------------------------------------------------------------------------------
// modifyValueAndActiveState on the float tree, with OpenVDB's tile probe:
//  InternalNode::modifyValueAndActiveStateAndCache: if the slot is a tile, the op is first run
//  on copies (tileValue, !tileState); a child is created iff the result differs from the tile;
//  then the op runs for real in the child. RootNode creates a missing child without a probe.
// --------------------------------------------------------------------------------------------
template <typename Op>
inline void modifyInLeaf(FloatLeaf* l, const Coord& c, Op& op)
{
  uint32_t n = leafOffset(c);
  bool state = l->vmask.isOn(n);
  op(l->buf[n], state);
  l->vmask.set(n, state);
}

template <typename Op>
inline void modifyInI1(Accessor<FloatTree>& acc, FloatI1* a, const Coord& c, Op& op)
{
  uint32_t n    = FloatI1::coordToOffset(c);
  bool hasChild = a->childMask.isOn(n);
  if (!hasChild)
  {
    const bool tileState = a->valueMask.isOn(n);
    const float tileVal  = a->nodes[n].tile;
    bool modifiedState   = !tileState;
    float modifiedVal    = tileVal;
    op(modifiedVal, modifiedState);
    if (modifiedState != tileState || !(modifiedVal == tileVal))
    {
      hasChild          = true;
      a->nodes[n].child = new FloatLeaf(c, tileVal, tileState);
      a->childMask.setOn(n);
      a->valueMask.setOff(n);
    }
  }
  if (hasChild)
  {
    FloatLeaf* l = a->nodes[n].child;
    acc.insert(c, l);
    modifyInLeaf(l, c, op);
  }
}

template <typename Op>
inline void modifyInI2(Accessor<FloatTree>& acc, FloatI2* b, const Coord& c, Op& op)
{
  uint32_t n    = FloatI2::coordToOffset(c);
  bool hasChild = b->childMask.isOn(n);
  if (!hasChild)
  {
    const bool tileState = b->valueMask.isOn(n);
    const float tileVal  = b->nodes[n].tile;
    bool modifiedState   = !tileState;
    float modifiedVal    = tileVal;
    op(modifiedVal, modifiedState);
    if (modifiedState != tileState || !(modifiedVal == tileVal))
    {
      hasChild          = true;
      b->nodes[n].child = new FloatI1(c, tileVal, tileState);
      b->childMask.setOn(n);
      b->valueMask.setOff(n);
    }
  }
  if (hasChild)
  {
    FloatI1* a = b->nodes[n].child;
    acc.insert(c, a);
    modifyInI1(acc, a, c, op);
  }
}

template <typename Op>
inline void modifyValueAndActiveState(Accessor<FloatTree>& acc, const Coord& c, Op& op)
{
  if (acc.hashed0(c)) { modifyInLeaf(acc.n0, c, op); return; }
  if (acc.hashed1(c)) { modifyInI1(acc, acc.n1, c, op); return; }
  if (acc.hashed2(c)) { modifyInI2(acc, acc.n2, c, op); return; }
  FloatTree* t = acc.tree;
  Coord key    = FloatTree::rootKey(c);
  auto it      = t->table.find(key);
  FloatI2* b;
  if (it == t->table.end())
  {
    b             = new FloatI2(c, t->background, false);
    t->table[key] = b;
  }
  else b = it->second;
  acc.insert(c, b);
  modifyInI2(acc, b, c, op);
}

// --------------------------------------------------------------------------------------------
// fast_mode / raytrace support: math::Ray<double>, math::DDA<RayT, Log2Dim>, math::VolumeHDDA and
// tools::VolumeRayIntersector<FloatGrid> restated from OpenVDB's published algorithm (math/Ray.h, math/DDA.h,
// tools/RayIntersector.h; v9.0 is what the reference's README recommends). PARITY UNPINNED: the reference holds no
// test or vector for either caller (V:577-602, V:675-721) and OpenVDB is absent from this image.
//   Ray(eye, dir, t0, t1): direction stored as given (not normalised), invDir = 1/dir, ray(t) = eye + dir * t
//   Ray::clip(bbox): slab test against the CoordBBox corners, shrinking [t0, t1]; the span is left alone on a miss
//   DDA<Ray, Log2Dim>::init(ray, t0, t1): voxel = floor(ray(t0)) & ~(DIM-1), per axis next = t0 + (voxel [+ DIM] - pos) * inv,
//     delta = +-DIM * inv, a zero direction component disables the axis (DBL_MAX); step(): axis = MinIndex(next),
//     time = next[axis], next[axis] += delta[axis], voxel[axis] += +-DIM, "more" while time <= t1;
//     next() = Min(t1, next[0], next[1], next[2])
//   VolumeHDDA<Tree, Ray, Level>: DDA over the nodes of that level; a child node -> recurse with the ray restricted to
//     [time(), next()]; an active tile opens a span; anything else closes the open span (kept if longer than Delta = 1e-9);
//     at level 0 a LEAF (whatever its voxels hold) or an active tile opens the span
//   VolumeRayIntersector(grid): topology copy of the tree, bbox = root.evalActiveBoundingBox(visit_voxels = false) with
//     max += 1; setIndexRay: ray.clip(bbox) (return value unused by the reference's fast path); hits(): all spans;
//     march(t0, t1): the first span, only attempted when the ray's own span is valid
// --------------------------------------------------------------------------------------------
constexpr double kDelta = 1e-9; // math::Delta<double>::value()

struct TimeSpan
{
  double t0, t1;
  bool valid(double eps = kDelta) const { return (t1 - t0) > eps; }
  void set(double a, double b) { t0 = a; t1 = b; }
};

struct Ray
{
  double eye[3], dir[3], inv[3];
  TimeSpan span;
  Ray(const double e[3], const double d[3], double t0, double t1)
  {
    for (int a = 0; a < 3; ++a) { eye[a] = e[a]; dir[a] = d[a]; inv[a] = 1.0 / d[a]; }
    span.set(t0, t1);
  }
  void at(double t, double out[3]) const
  {
    for (int a = 0; a < 3; ++a) out[a] = eye[a] + dir[a] * t; // mEye + mDir * time
  }
  bool valid() const { return span.valid(); }
  void setTimes(double t0, double t1) { span.set(t0, t1); }
  // Ray::intersects(bbox, t0, t1) + Ray::clip
  bool clip(const int32_t bmin[3], const int32_t bmax[3])
  {
    double t0 = span.t0, t1 = span.t1;
    for (int i = 0; i < 3; ++i)
    {
      double a = (double(bmin[i]) - eye[i]) * inv[i];
      double b = (double(bmax[i]) - eye[i]) * inv[i];
      if (a > b) std::swap(a, b);
      if (a > t0) t0 = a;
      if (b < t1) t1 = b;
      if (t0 > t1) return false;
    }
    span.set(t0, t1);
    return true;
  }
};

inline int minIndex3(const double n[3])
{
  static const int table[8] = {2, 1, 9, 1, 2, 9, 0, 0}; // math::MinIndex
  return table[(int(n[0] < n[1]) << 2) + (int(n[0] < n[2]) << 1) + int(n[1] < n[2])];
}

template <int LOG2DIM>
struct DDA
{
  static constexpr int32_t DIM = int32_t(1) << LOG2DIM;
  double mT0, mT1, mNext[3], mDelta[3];
  Coord mVoxel;
  int32_t mStep[3];
  void init(const Ray& ray) { init(ray, ray.span.t0, ray.span.t1); }
  void init(const Ray& ray, double startTime, double maxTime)
  {
    mT0 = startTime;
    mT1 = maxTime;
    double pos[3];
    ray.at(mT0, pos);
    for (int a = 0; a < 3; ++a) mVoxel[a] = int32_t(std::floor(pos[a])) & ~(DIM - 1);
    for (int a = 0; a < 3; ++a)
    {
      if (ray.dir[a] == 0.0) // math::isZero
      {
        mStep[a]  = 0;
        mNext[a]  = DBL_MAX;
        mDelta[a] = DBL_MAX;
      }
      else if (ray.inv[a] > 0)
      {
        mStep[a]  = DIM;
        mNext[a]  = mT0 + (double(mVoxel[a] + DIM) - pos[a]) * ray.inv[a];
        mDelta[a] = double(mStep[a]) * ray.inv[a];
      }
      else
      {
        mStep[a]  = -DIM;
        mNext[a]  = mT0 + (double(mVoxel[a]) - pos[a]) * ray.inv[a];
        mDelta[a] = double(mStep[a]) * ray.inv[a];
      }
    }
  }
  bool step()
  {
    const int a = minIndex3(mNext);
    mT0         = mNext[a];
    mNext[a] += mDelta[a];
    mVoxel[a] += mStep[a];
    return mT0 <= mT1;
  }
  double time() const { return mT0; }
  double maxTime() const { return mT1; }
  double next() const { return std::min(std::min(mT1, mNext[0]), std::min(mNext[1], mNext[2])); } // math::Min(a, b, c, d)
};

// VolumeHDDA<FloatTree, Ray, 2 / 1 / 0>. SpanFn(const TimeSpan&) -> true = terminate (march), false = keep going (hits).
struct VolumeHDDA
{
  template <typename SpanFn>
  static bool level0(Ray& ray, Accessor<FloatTree>& acc, TimeSpan& t, SpanFn& fn)
  {
    DDA<3> dda;
    dda.init(ray);
    do
    {
      if (acc.probeLeaf(dda.mVoxel) || acc.isValueOn(dda.mVoxel))
      {
        if (t.t0 < 0) t.t0 = dda.time();
      }
      else if (t.t0 >= 0)
      {
        t.t1 = dda.time();
        if (t.valid() && fn(t)) return true;
        t.set(-1, -1);
      }
    } while (dda.step());
    if (t.t0 >= 0) t.t1 = dda.maxTime();
    return false;
  }
  template <typename SpanFn>
  static bool level1(Ray& ray, Accessor<FloatTree>& acc, TimeSpan& t, SpanFn& fn)
  {
    DDA<7> dda;
    dda.init(ray);
    do
    {
      if (acc.probeI1(dda.mVoxel) != nullptr)
      {
        ray.setTimes(dda.time(), dda.next());
        if (level0(ray, acc, t, fn)) return true;
      }
      else if (acc.isValueOn(dda.mVoxel))
      {
        if (t.t0 < 0) t.t0 = dda.time();
      }
      else if (t.t0 >= 0)
      {
        t.t1 = dda.time();
        if (t.valid() && fn(t)) return true;
        t.set(-1, -1);
      }
    } while (dda.step());
    if (t.t0 >= 0) t.t1 = dda.maxTime();
    return false;
  }
  template <typename SpanFn>
  static bool level2(Ray& ray, Accessor<FloatTree>& acc, TimeSpan& t, SpanFn& fn)
  {
    DDA<12> dda;
    dda.init(ray);
    do
    {
      if (acc.probeI2(dda.mVoxel) != nullptr)
      {
        ray.setTimes(dda.time(), dda.next());
        if (level1(ray, acc, t, fn)) return true;
      }
      else if (acc.isValueOn(dda.mVoxel))
      {
        if (t.t0 < 0) t.t0 = dda.time();
      }
      else if (t.t0 >= 0)
      {
        t.t1 = dda.time();
        if (t.valid() && fn(t)) return true;
        t.set(-1, -1);
      }
    } while (dda.step());
    if (t.t0 >= 0) t.t1 = dda.maxTime();
    return false;
  }
};

struct VolumeRayIntersector
{
  // The reference's intersector owns a topology COPY of the map taken after the last integrateUpdate (V:1436-1449); the map
  // cannot change while accumulateUpdate / raytrace hold the shared lock, so reading the live tree is the same thing whenever
  // the map was last changed by integrateUpdate. Other writers (sections, point edits) leave the reference with a stale
  // copy; that staleness is NOT restated (the B200 path and this oracle both read the current map).
  Accessor<FloatTree> acc;
  int32_t bmin[3], bmax[3];
  double tmax = 0;
  explicit VolumeRayIntersector(FloatTree& tree) : acc(tree)
  {
    // RootNode / InternalNode / LeafNode::evalActiveBoundingBox(bbox, visitVoxels = false): active tiles with their full
    // extent, leaves that hold at least one active voxel with their node bounding box
    for (int a = 0; a < 3; ++a) { bmin[a] = INT32_MAX; bmax[a] = INT32_MIN; }
    auto expand = [&](const Coord& o, int32_t dim) {
      for (int a = 0; a < 3; ++a)
      {
        bmin[a] = std::min(bmin[a], o[a]);
        bmax[a] = std::max(bmax[a], o[a] + dim - 1);
      }
    };
    for (auto& kv : tree.table)
    {
      const FloatI2* n2 = kv.second;
      for (uint32_t i = n2->valueMask.findNextOn(0); i < FloatI2::NUM; i = n2->valueMask.findNextOn(i + 1))
        expand(Coord(n2->origin[0] + int32_t(i >> 10) * 128, n2->origin[1] + int32_t((i >> 5) & 31) * 128, n2->origin[2] + int32_t(i & 31) * 128), 128);
      for (uint32_t i = n2->childMask.findNextOn(0); i < FloatI2::NUM; i = n2->childMask.findNextOn(i + 1))
      {
        const FloatI1* n1 = n2->nodes[i].child;
        for (uint32_t k = n1->valueMask.findNextOn(0); k < FloatI1::NUM; k = n1->valueMask.findNextOn(k + 1))
          expand(Coord(n1->origin[0] + int32_t(k >> 8) * 8, n1->origin[1] + int32_t((k >> 4) & 15) * 8, n1->origin[2] + int32_t(k & 15) * 8), 8);
        for (uint32_t k = n1->childMask.findNextOn(0); k < FloatI1::NUM; k = n1->childMask.findNextOn(k + 1))
        {
          const FloatLeaf* l = n1->nodes[k].child;
          if (!l->vmask.isOff()) expand(l->origin, 8);
        }
      }
    }
    for (int a = 0; a < 3; ++a) bmax[a] = int32_t(uint32_t(bmax[a]) + 1u); // mBBox.max().offset(1)
  }
  // setIndexRay: the clipped ray is what hits() / march() traverse
  bool setIndexRay(Ray& ray)
  {
    const bool hit = ray.clip(bmin, bmax);
    if (hit) tmax = ray.span.t1;
    return hit;
  }
  void hits(Ray& ray, std::vector<TimeSpan>& list)
  {
    TimeSpan t{-1, -1};
    list.clear();
    auto keep = [&](const TimeSpan& s) { list.push_back(s); return false; };
    VolumeHDDA::level2(ray, acc, t, keep);
    if (t.valid()) list.push_back(t);
  }
  bool march(Ray& ray, double& t0, double& t1)
  {
    TimeSpan t{-1, -1};
    auto stop = [&](const TimeSpan&) { return true; };
    if (ray.valid()) VolumeHDDA::level2(ray, acc, t, stop);
    t0 = t.t0;
    t1 = t.t1;
    return t.valid();
  }
};

// --------------------------------------------------------------------------------------------
// The mapping classes (restated reference logic).
// --------------------------------------------------------------------------------------------
struct Stats
{
  uint64_t rays        = 0; // points seen by raycastPointCloud (incl. NaN + clipped)
  uint64_t nan_skipped = 0;
  uint64_t clipped     = 0;
  uint64_t visits      = 0; // setActiveState calls made by castRayIntoGrid
  uint64_t voxel_updates = 0; // active update voxels consumed by updateMap
  uint64_t state_changes = 0; // voxels written to the change grid
};

struct Source
{
  std::string id;
  double max_range;
  std::unique_ptr<BoolTree> update_grid;
  std::unique_ptr<BoolTree> last_change; // what updateMap returned for this source last time
  // "reduced" update of the last accumulateUpdate call (level 2): the end voxel of every ray, active,
  // value = the ray delivered a hit (was not clipped), plus the scan origin. castRayIntoGrid depends only on the two
  // voxel indices, so re-raycasting this set from the origin reproduces that call's update grid exactly.
  std::unique_ptr<BoolTree> reduced;
  double last_origin[3] = {0, 0, 0};
  Stats stats; // raycast counters of this source (own block: sources may accumulate concurrently, one thread each, V:1373)
};

// VDBMapping<float,Config> + OccupancyVDBMapping node ops, polymorphic like the reference
// (virtual updateFreeNode/updateOccupiedNode, V:1472-1473) so the per-voxel cost model matches.
struct MappingBase
{
  double m_resolution;
  double m_inv_resolution; // ScaleMap::mScaleValuesInverse = 1.0 / scale
  double m_max_range = 0.0;
  bool m_config_set  = false;
  std::unique_ptr<FloatTree> m_vdb_grid;
  std::map<std::string, std::unique_ptr<Source> > m_input_sources; // std::map order, V:1545
  Stats stats;
  bool replicate_probe_quirk = true; // SURVEY F9; false = report only real flag flips
  bool m_fast_mode           = false; // Config::fast_mode V:1466

  explicit MappingBase(double resolution) : m_resolution(resolution), m_inv_resolution(1.0 / resolution)
  {
    m_vdb_grid.reset(new FloatTree(0.0f)); // createVDBMap V:163-169, background TData()
  }
  virtual ~MappingBase() {}
  virtual bool updateFreeNode(float&, bool&) { return false; }
  virtual bool updateOccupiedNode(float&, bool&) { return false; }
  virtual bool setNodeToFree(float&, bool&) { return false; }     // V:1474
  virtual bool setNodeToOccupied(float&, bool&) { return false; } // V:1475
  virtual bool setNodeState(float&, bool&) { return false; }      // V:1476
  std::unique_ptr<BoolTree> m_artificial_area_grid{new BoolTree(false)}; // V:132,1493

  // V:174-186
  void resetMap()
  {
    m_vdb_grid.reset(new FloatTree(0.0f));
    for (auto& kv : m_input_sources) kv.second->update_grid.reset(new BoolTree(false));
  }

  // V:1352-1375 (threads not restated: the oracle is driven synchronously)
  void addInputSource(const std::string& id, double max_range)
  {
    auto s       = std::make_unique<Source>();
    s->id        = id;
    s->max_range = (max_range == 0) ? m_max_range : max_range;
    s->update_grid.reset(new BoolTree(false));
    m_input_sources[id] = std::move(s);
  }

  // V:612-631; Transform::worldToIndex -> ScaleMap::applyInverseMap (multiply by 1/res);
  // Coord::floor = Int32(std::floor(x)).
  Coord worldToIndex(const double w[3]) const
  {
    Coord r;
    for (int i = 0; i < 3; ++i)
    {
      double c = w[i];
      if (std::fmod(c, m_resolution)) c = c + (m_resolution / 2.0);
      r[i] = int32_t(std::floor(c * m_inv_resolution));
    }
    return r;
  }

  // V:550-566 + openvdb::math::Ray<double>(eye, dir, 0, 1) + DDA<RayT,0>(ray, 0)
  void castRayIntoGrid(const Coord& o, const Coord& e, Accessor<BoolTree>& acc, Stats& st)
  {
    if (e == o) return; // V:559 (ray/dda construction has no side effect)
    double dir[3], inv[3], pos[3], next[3], delta[3];
    int32_t step[3];
    Coord voxel;
    for (int a = 0; a < 3; ++a)
    {
      dir[a] = double(e[a]) - double(o[a]);  // V:554 asVec3d() - Coord
      pos[a] = double(o[a]) + 0.5;           // V:557 eye = origin + 0.5; ray(t0=0) = eye + dir*0
      pos[a] = pos[a] + dir[a] * 0.0;
      inv[a] = 1.0 / dir[a];                 // Ray::mInvDir = 1/mDir (true division; 1/0 = inf)
      voxel[a] = int32_t(std::floor(pos[a])); // Coord::floor(pos) & ~(DIM-1), DIM = 1
    }
    for (int a = 0; a < 3; ++a)
    {
      if (dir[a] == 0.0) // math::isZero(dir[axis])
      {
        step[a]  = 0;
        next[a]  = DBL_MAX;
        delta[a] = DBL_MAX;
      }
      else if (inv[a] > 0)
      {
        step[a]  = 1;
        next[a]  = 0.0 + (double(voxel[a] + 1) - pos[a]) * inv[a];
        delta[a] = double(step[a]) * inv[a];
      }
      else
      {
        step[a]  = -1;
        next[a]  = 0.0 + (double(voxel[a]) - pos[a]) * inv[a];
        delta[a] = double(step[a]) * inv[a];
      }
    }
    static const int kMinIndexTable[8] = {2, 1, 9, 1, 2, 9, 0, 0}; // math::MinIndex
    bool more;
    do
    {
      setActiveStateOn(acc, voxel); // V:563
      ++st.visits;
      // DDA::step()
      const int key  = (int(next[0] < next[1]) << 2) + (int(next[0] < next[2]) << 1) + int(next[1] < next[2]);
      const int axis = kMinIndexTable[key];
      const double t = next[axis];
      next[axis] += delta[axis];
      voxel[axis] += step[axis];
      more = (t <= 1.0);
    } while (more);
  }

  // V:577-602: the ray is intersected with the map's node topology; a voxel DDA runs over every hit span and only voxels
  // that are active in the map are activated in the update grid. st.visits counts those setActiveState calls (V:598).
  void castRayIntoGridFast(const Coord& o, const Coord& e, Accessor<FloatTree>& grid_acc, Accessor<BoolTree>& update_acc,
                           VolumeRayIntersector& intersector, Stats& st)
  {
    double eye[3], dir[3];
    for (int a = 0; a < 3; ++a)
    {
      dir[a] = double(e[a]) - double(o[a]); // V:586
      eye[a] = double(o[a]) + 0.5;          // V:587
    }
    Ray ray(eye, dir, 0, 1);
    intersector.setIndexRay(ray); // V:588, result unused
    std::vector<TimeSpan> hits;
    intersector.hits(ray, hits);  // V:589-590
    for (const TimeSpan& hit : hits)
    {
      Ray fine_ray(eye, dir, hit.t0, hit.t1); // V:593
      DDA<0> dda;
      dda.init(fine_ray);
      do
      {
        if (grid_acc.isValueOn(dda.mVoxel)) // V:597
        {
          setActiveStateOn(update_acc, dda.mVoxel);
          ++st.visits;
        }
      } while (dda.step());
    }
  }

  // V:675-721. Vec3::normalize(): *= 1/length; Transform::worldToIndex / indexToWorld of a uniform scale map: multiply by the
  // stored inverse scale / by the scale. m_volume_ray_intersector is only ever set in fast_mode (V:1438): the reference
  // dereferences a null pointer otherwise; here the intersector is always the current map. An EMPTY map (the reference
  // never builds an intersector for one) reports no hit.
  void raytrace(size_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* successes, double* end_points)
  {
    if (m_vdb_grid->empty())
    {
      for (size_t i = 0; i < n; ++i)
      {
        const double* d = directions + 3 * i;
        const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        const double il  = 1.0 / len;
        for (int a = 0; a < 3; ++a)
        {
          const double dn  = (d[a] * il) * max_lengths[i];
          const double oi  = origins[3 * i + a] * m_inv_resolution;
          const double di  = dn * m_inv_resolution;
          end_points[3 * i + a] = (oi + di) * m_resolution;
        }
        successes[i] = 0;
      }
      return;
    }
    VolumeRayIntersector intersector(*m_vdb_grid);
    Accessor<FloatTree> acc(*m_vdb_grid);
    for (size_t i = 0; i < n; ++i)
    {
      const double* d  = directions + 3 * i;
      const double len = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]); // V:691 normalize
      const double il  = 1.0 / len;
      double oi[3], di[3];
      for (int a = 0; a < 3; ++a)
      {
        const double dn = (d[a] * il) * max_lengths[i];  // V:692
        oi[a] = origins[3 * i + a] * m_inv_resolution;   // V:694
        di[a] = dn * m_inv_resolution;                   // V:695
      }
      Ray ray(oi, di, 0, 1);           // V:698
      intersector.setIndexRay(ray);    // V:700
      double t0, t1;
      if (intersector.march(ray, t0, t1)) // V:704
      {
        Ray fine_ray(oi, di, t0, t1);
        DDA<0> dda;
        dda.init(fine_ray);
        while (dda.step() && !acc.isValueOn(dda.mVoxel)) {} // V:709-713
        for (int a = 0; a < 3; ++a) end_points[3 * i + a] = double(dda.mVoxel[a]) * m_resolution; // V:714
        successes[i] = 1;
      }
      else
      {
        for (int a = 0; a < 3; ++a) end_points[3 * i + a] = (oi[a] + di[a]) * m_resolution; // V:719
        successes[i] = 0;
      }
    }
  }

  // V:466-539
  bool raycastPointCloud(const uint8_t* pts, size_t n, size_t stride, const double origin[3], double raycast_range,
                         Accessor<BoolTree>& update_acc, Stats& st, Accessor<BoolTree>* reduced_acc = nullptr)
  {
    if (!m_config_set) return false; // V:478-482
    const Coord ray_origin_index = worldToIndex(origin);
    const bool grid_empty        = m_vdb_grid->empty(); // V:495
    std::unique_ptr<VolumeRayIntersector> intersector;
    std::unique_ptr<Accessor<FloatTree> > grid_acc;
    if (m_fast_mode && !grid_empty)
    {
      intersector.reset(new VolumeRayIntersector(*m_vdb_grid)); // V:1440 (see VolumeRayIntersector about the snapshot)
      grid_acc.reset(new Accessor<FloatTree>(*m_vdb_grid));     // V:496
    }
    const bool origin_nan        = std::isnan(origin[0]) || std::isnan(origin[1]) || std::isnan(origin[2]);
    for (size_t i = 0; i < n; ++i)
    {
      float p[3];
      std::memcpy(p, pts + i * stride, sizeof(p));
      double end[3]      = {double(p[0]), double(p[1]), double(p[2])}; // V:501
      bool max_range_ray = false;
      ++st.rays;
      if (std::isnan(end[0]) || std::isnan(end[1]) || std::isnan(end[2]) || origin_nan) // V:505-510
      {
        ++st.nan_skipped;
        continue;
      }
      // +-inf is UNDEFINED in the reference (it ends in Coord::floor(NaN) / an int cast of inf). The documented behaviour of
      // the B200 path - and therefore of its checker - is to drop such a point like a NaN.
      if (std::isinf(end[0]) || std::isinf(end[1]) || std::isinf(end[2]) || std::isinf(origin[0]) || std::isinf(origin[1]) || std::isinf(origin[2]))
      {
        ++st.nan_skipped;
        continue;
      }
      if (raycast_range > 0.0)
      {
        // Vec3::length(): sqrt(x*x + y*y + z*z), left-associated
        const double d[3] = {end[0] - origin[0], end[1] - origin[1], end[2] - origin[2]};
        const double len  = std::sqrt(d[0] * d[0] + d[1] * d[1] + d[2] * d[2]);
        if (len > raycast_range) // V:512
        {
          // origin + (d.unit() * range): unit() = d / len (true division), V:514-515
          for (int a = 0; a < 3; ++a) end[a] = origin[a] + (d[a] / len) * raycast_range;
          max_range_ray = true;
          ++st.clipped;
        }
      }
      const Coord ray_end_index = worldToIndex(end); // V:519
      if (m_fast_mode) // V:520-527
      {
        if (!grid_empty) castRayIntoGridFast(ray_origin_index, ray_end_index, *grid_acc, update_acc, *intersector, st);
      }
      else castRayIntoGrid(ray_origin_index, ray_end_index, update_acc, st); // V:530
      if (!max_range_ray) setValueOnTrue(update_acc, ray_end_index); // V:533-536
      if (reduced_acc)
      {
        if (max_range_ray) setActiveStateOn(*reduced_acc, ray_end_index);
        else setValueOnTrue(*reduced_acc, ray_end_index);
      }
    }
    return true;
  }

  // V:316-346
  int accumulateUpdate(const uint8_t* pts, size_t n, size_t stride, const double origin[3], const std::string& id)
  {
    auto it = m_input_sources.find(id);
    if (it == m_input_sources.end()) return 1; // V:321-326: print + return
    Source& s = *it->second;
    Accessor<BoolTree> acc(*s.update_grid);
    if (s.max_range > 0) // V:331
    {
      s.reduced.reset(new BoolTree(false));
      for (int a = 0; a < 3; ++a) s.last_origin[a] = origin[a];
      Accessor<BoolTree> racc(*s.reduced);
      if (!raycastPointCloud(pts, n, stride, origin, s.max_range, acc, s.stats, &racc)) return 2;
    }
    return 0;
  }

  // Receiver side of a level-2 (reduced) update: re-raycast every end voxel from the transmitted origin
  // (castRayIntoGrid V:550-566 + the end-point rule V:533-536) into `acc`.
  void raycastReducedGrid(const BoolTree& reduced, const double origin[3], Accessor<BoolTree>& acc)
  {
    const Coord o = worldToIndex(origin);
    reduced.forEachLeaf([&](const BoolLeaf& rl) {
      for (uint32_t n = rl.vmask.findNextOn(0); n < 512; n = rl.vmask.findNextOn(n + 1))
      {
        Coord l = leafOffsetToLocal(n);
        Coord c(rl.origin[0] + l[0], rl.origin[1] + l[1], rl.origin[2] + l[2]);
        ++stats.rays;
        castRayIntoGrid(o, c, acc, stats);
        if (rl.buf.isOn(n)) setValueOnTrue(acc, c);
      }
    });
  }

  // Level-1 apply ("overwrite" grid = the change grid updateMap returned): every active voxel is forced to the
  // occupied (value true) or free (value false) state with the reference's own node ops, V:1474-1475 / O:118-129.
  void overwriteMap(const BoolTree& grid)
  {
    Accessor<FloatTree> acc(*m_vdb_grid);
    auto occ = [&](float& v, bool& a) { setNodeToOccupied(v, a); };
    auto fre = [&](float& v, bool& a) { setNodeToFree(v, a); };
    grid.forEachLeaf([&](const BoolLeaf& gl) {
      for (uint32_t n = gl.vmask.findNextOn(0); n < 512; n = gl.vmask.findNextOn(n + 1))
      {
        Coord l = leafOffsetToLocal(n);
        Coord c(gl.origin[0] + l[0], gl.origin[1] + l[1], gl.origin[2] + l[2]);
        if (gl.buf.isOn(n)) modifyValueAndActiveState(acc, c, occ);
        else modifyValueAndActiveState(acc, c, fre);
      }
    });
  }

  // addPointsToGrid V:431-447 / removePointsFromGrid V:413-429: Coord::floor(grid->worldToIndex(pt)) (the plain
  // transform, NOT the fmod variant of V:612-631), then setNodeToOccupied / setNodeToFree.
  void setPoints(const uint8_t* pts, size_t n, size_t stride, bool occupied)
  {
    Accessor<FloatTree> acc(*m_vdb_grid);
    auto occ = [&](float& v, bool& a) { setNodeToOccupied(v, a); };
    auto fre = [&](float& v, bool& a) { setNodeToFree(v, a); };
    for (size_t i = 0; i < n; ++i)
    {
      float p[3];
      std::memcpy(p, pts + i * stride, sizeof(p));
      if (!std::isfinite(p[0]) || !std::isfinite(p[1]) || !std::isfinite(p[2])) continue; // undefined in the reference
      Coord c;
      for (int a = 0; a < 3; ++a) c[a] = int32_t(std::floor(double(p[a]) * m_inv_resolution));
      if (occupied) modifyValueAndActiveState(acc, c, occ);
      else modifyValueAndActiveState(acc, c, fre);
    }
  }

  // V:1152-1166
  void restoreMapIntegrity()
  {
    Accessor<FloatTree> acc(*m_vdb_grid);
    auto restore = [&](float& v, bool& a) { setNodeState(v, a); };
    m_artificial_area_grid->forEachLeaf([&](const BoolLeaf& gl) {
      for (uint32_t n = gl.vmask.findNextOn(0); n < 512; n = gl.vmask.findNextOn(n + 1))
      {
        Coord l = leafOffsetToLocal(n);
        modifyValueAndActiveState(acc, Coord(gl.origin[0] + l[0], gl.origin[1] + l[1], gl.origin[2] + l[2]), restore);
      }
    });
    m_artificial_area_grid.reset(new BoolTree(false));
  }
  // V:1219-1236; start/end are world points
  void addArtificialWall(const double start[3], const double end[3], double negative_height, double positive_height)
  {
    Accessor<BoolTree> acc(*m_artificial_area_grid);
    const Coord s = worldToIndex(start), e = worldToIndex(end);
    const int negative_index = (int)(negative_height / m_resolution);
    const int positive_index = (int)(positive_height / m_resolution);
    Stats walls; // walls are not scan rays: keep them out of the counters
    for (int i = negative_index; i < positive_index; ++i)
      castRayIntoGrid(Coord(s[0], s[1], s[2] + i), Coord(e[0], e[1], e[2] + i), acc, walls);
  }
  // V:1175-1206; polygons: n_poly polygons, counts[p] points each, xyz triples (the 4th homogeneous component is unused)
  void addArtificialAreas(size_t n_poly, const uint32_t* counts, const double* xyz, double negative_height, double positive_height)
  {
    restoreMapIntegrity();
    size_t base = 0;
    for (size_t p = 0; p < n_poly; ++p)
    {
      for (uint32_t i = 0; i < counts[p]; ++i)
        addArtificialWall(xyz + 3 * (base + i), xyz + 3 * (base + (i + 1) % counts[p]), negative_height, positive_height);
      base += counts[p];
    }
  }

  // V:731-792
  std::unique_ptr<BoolTree> updateMap(const BoolTree& temp_grid)
  {
    auto change = std::make_unique<BoolTree>(false);
    Accessor<BoolTree> change_acc(*change);
    if (temp_grid.empty()) return change; // V:735-738

    bool state_changed = false;
    Accessor<FloatTree> acc(*m_vdb_grid);
    const bool quirk = replicate_probe_quirk;
    bool in_probe_guard = false; // used only when the quirk is switched off
    (void)in_probe_guard;
    auto miss = [&](float& voxel_value, bool& active) {
      bool last_state = active;
      updateFreeNode(voxel_value, active);
      if (last_state != active) state_changed = true;
    };
    auto hit = [&](float& voxel_value, bool& active) {
      bool last_state = active;
      updateOccupiedNode(voxel_value, active);
      if (last_state != active) state_changed = true;
    };

    // cbeginValueOn(): root key order, then child offset order per level, then leaf offset order
    temp_grid.forEachLeaf([&](const BoolLeaf& ul) {
      for (uint32_t n = ul.vmask.findNextOn(0); n < 512; n = ul.vmask.findNextOn(n + 1))
      {
        Coord l = leafOffsetToLocal(n);
        Coord c(ul.origin[0] + l[0], ul.origin[1] + l[1], ul.origin[2] + l[2]);
        ++stats.voxel_updates;
        state_changed = false;
        if (!quirk)
        {
          // "clean" semantics: report only a real flip of the stored voxel flag
          const FloatLeaf* fl = acc.probeLeaf(c);
          const bool before   = fl ? fl->vmask.isOn(n) : false;
          if (ul.buf.isOn(n)) modifyValueAndActiveState(acc, c, hit);
          else modifyValueAndActiveState(acc, c, miss);
          const FloatLeaf* fl2 = acc.probeLeaf(c);
          const bool after     = fl2 ? fl2->vmask.isOn(n) : false;
          state_changed        = (before != after);
        }
        else
        {
          if (ul.buf.isOn(n)) modifyValueAndActiveState(acc, c, hit);
          else modifyValueAndActiveState(acc, c, miss);
        }
        if (state_changed)
        {
          ++stats.state_changes;
          if (ul.buf.isOn(n)) setValueOnTrue(change_acc, c);      // V:772
          else setActiveStateOn(change_acc, c);                   // V:780
        }
      }
    });
    // V:785-789: every artificial-area voxel is forced active (a missing leaf is created with background values)
    m_artificial_area_grid->forEachLeaf([&](const BoolLeaf& gl) {
      for (uint32_t n = gl.vmask.findNextOn(0); n < 512; n = gl.vmask.findNextOn(n + 1))
      {
        Coord l = leafOffsetToLocal(n);
        setActiveStateOn(acc, Coord(gl.origin[0] + l[0], gl.origin[1] + l[1], gl.origin[2] + l[2]));
      }
    });
    return change;
  }

  // V:375-387
  void integrateUpdate()
  {
    for (auto& kv : m_input_sources)
    {
      Source& s     = *kv.second;
      s.last_change = updateMap(*s.update_grid);
      s.update_grid.reset(new BoolTree(false));
    }
  }
};

struct OccupancyMapping : MappingBase
{
  float m_logodds_hit = 0, m_logodds_miss = 0, m_logodds_thres_min = 0, m_logodds_thres_max = 0, m_max_logodds = 0,
        m_min_logodds = 0;
  explicit OccupancyMapping(double res) : MappingBase(res) {}

  // base V:1456-1469 then O:59-89. Returns 0 ok, 1 base rejected, 2 derived rejected (the base
  // has already flipped m_config_set by then, exactly like the reference).
  int setConfig(double max_range, double prob_hit, double prob_miss, double prob_thres_min, double prob_thres_max)
  {
    if (max_range < 0.0) return 1;
    m_max_range  = max_range;
    m_config_set = true;
    if (prob_miss > 0.5) return 2;
    if (prob_hit < 0.5) return 2;
    m_logodds_miss      = static_cast<float>(std::log(prob_miss) - std::log(1 - prob_miss));
    m_logodds_hit       = static_cast<float>(std::log(prob_hit) - std::log(1 - prob_hit));
    m_logodds_thres_min = static_cast<float>(std::log(prob_thres_min) - std::log(1 - prob_thres_min));
    m_logodds_thres_max = static_cast<float>(std::log(prob_thres_max) - std::log(1 - prob_thres_max));
    m_max_logodds       = static_cast<float>(std::log(0.99) - std::log(0.01));
    m_min_logodds       = static_cast<float>(std::log(0.01) - std::log(0.99));
    m_config_set        = true;
    return 0;
  }
  // O:92-104
  bool updateFreeNode(float& voxel_value, bool& active) override
  {
    voxel_value += m_logodds_miss;
    if (voxel_value < m_logodds_thres_min)
    {
      active = false;
      if (voxel_value < m_min_logodds) voxel_value = m_min_logodds;
    }
    return true;
  }
  // O:105-117
  bool updateOccupiedNode(float& voxel_value, bool& active) override
  {
    voxel_value += m_logodds_hit;
    if (voxel_value > m_logodds_thres_max)
    {
      active = true;
      if (voxel_value > m_max_logodds) voxel_value = m_max_logodds;
    }
    return true;
  }
  // O:118-135
  bool setNodeToFree(float& voxel_value, bool& active) override
  {
    voxel_value = m_min_logodds;
    active      = false;
    return true;
  }
  bool setNodeToOccupied(float& voxel_value, bool& active) override
  {
    voxel_value = m_max_logodds;
    active      = true;
    return true;
  }
  bool setNodeState(float& voxel_value, bool& active) override
  {
    active = voxel_value > m_logodds_thres_max;
    return true;
  }
};

// --------------------------------------------------------------------------------------------
// Canonical leaf-set export (leaves sorted by origin x,y,z == RootNode/offset order already,
// because every level is visited in ascending x-major offset order... which is NOT globally
// lexicographic across nodes, so sort explicitly).
// --------------------------------------------------------------------------------------------
struct LeafSet
{
  std::vector<int32_t> origins;  // 3 per leaf
  std::vector<uint64_t> active;  // 8 per leaf
  std::vector<uint64_t> valmask; // 8 per leaf (bool grids) or empty
  std::vector<float> values;     // 512 per leaf (float grids) or empty
  size_t size() const { return origins.size() / 3; }
};

template <typename T>
static void permuteBlocks(std::vector<T>& v, const std::vector<size_t>& order, size_t block)
{
  if (v.empty()) return;
  std::vector<T> out(v.size());
  for (size_t i = 0; i < order.size(); ++i) std::memcpy(&out[i * block], &v[order[i] * block], block * sizeof(T));
  v.swap(out);
}

static void sortLeafSet(LeafSet& s)
{
  size_t n = s.size();
  std::vector<size_t> order(n);
  for (size_t i = 0; i < n; ++i) order[i] = i;
  std::sort(order.begin(), order.end(), [&](size_t a, size_t b) {
    return Coord(s.origins[3 * a], s.origins[3 * a + 1], s.origins[3 * a + 2]) <
           Coord(s.origins[3 * b], s.origins[3 * b + 1], s.origins[3 * b + 2]);
  });
  permuteBlocks(s.origins, order, 3);
  permuteBlocks(s.active, order, 8);
  permuteBlocks(s.valmask, order, 8);
  permuteBlocks(s.values, order, 512);
}

static LeafSet exportBool(const BoolTree& t)
{
  LeafSet s;
  t.forEachLeaf([&](const BoolLeaf& l) {
    for (int a = 0; a < 3; ++a) s.origins.push_back(l.origin[a]);
    for (int w = 0; w < 8; ++w) s.active.push_back(l.vmask.w[w]);
    for (int w = 0; w < 8; ++w) s.valmask.push_back(l.buf.w[w]);
  });
  sortLeafSet(s);
  return s;
}
static LeafSet exportFloat(const FloatTree& t)
{
  LeafSet s;
  t.forEachLeaf([&](const FloatLeaf& l) {
    for (int a = 0; a < 3; ++a) s.origins.push_back(l.origin[a]);
    for (int w = 0; w < 8; ++w) s.active.push_back(l.vmask.w[w]);
    s.values.insert(s.values.end(), l.buf, l.buf + 512);
  });
  sortLeafSet(s);
  return s;
}

// --------------------------------------------------------------------------------------------
// getMapSection V:921-960 with extractSparseLeaf V:999-1011 / extractFullLeaf V:970-989 on an
// inclusive index bounding box (createIndexBoundingBox V:857-871 stays with the caller: it is
// host-side PCL/Eigen maths on 8 corners).
//   result_float = false -> TResultGrid = UpdateGridT (bool, background false)
//   result_float = true  -> TResultGrid = GridT       (float, background 0)
// setValueOff(coord, v) on an (inactive, background) tile only creates nodes if v != background
// (InternalNode::setValueOffAndCache), which is why "existence" is content-driven below.
// --------------------------------------------------------------------------------------------
static inline bool insideBox(const Coord& c, const Coord& mn, const Coord& mx)
{
  return c[0] >= mn[0] && c[0] <= mx[0] && c[1] >= mn[1] && c[1] <= mx[1] && c[2] >= mn[2] && c[2] <= mx[2];
}

static LeafSet mapSection(const FloatTree& map, const Coord& mn, const Coord& mx, bool full, bool result_float)
{
  LeafSet s;
  map.forEachLeaf([&](const FloatLeaf& l) {
    // leaf bbox = [origin, origin+7]; CoordBBox::hasOverlap
    for (int a = 0; a < 3; ++a)
      if (l.origin[a] + 7 < mn[a] || l.origin[a] > mx[a]) return;
    uint64_t act[8] = {0}, vm[8] = {0};
    float vals[512];
    for (int i = 0; i < 512; ++i) vals[i] = 0.0f;
    bool exists = false;
    for (uint32_t n = 0; n < 512; ++n)
    {
      Coord loc = leafOffsetToLocal(n);
      Coord c(l.origin[0] + loc[0], l.origin[1] + loc[1], l.origin[2] + loc[2]);
      if (!insideBox(c, mn, mx)) continue;
      const bool on = l.vmask.isOn(n);
      if (!full)
      {
        if (!on) continue;
        act[n >> 6] |= uint64_t(1) << (n & 63); // setValueOn(coord, true)
        vm[n >> 6] |= uint64_t(1) << (n & 63);
        vals[n] = 1.0f;
        exists  = true;
      }
      else
      {
        const float v = l.buf[n];
        if (on)
        {
          act[n >> 6] |= uint64_t(1) << (n & 63);
          exists = true;
        }
        if (result_float)
        {
          vals[n] = v;
          if (v != 0.0f) exists = true;
        }
        else if (v != 0.0f) // float -> bool
        {
          vm[n >> 6] |= uint64_t(1) << (n & 63);
          exists = true;
        }
      }
    }
    if (!exists) return;
    for (int a = 0; a < 3; ++a) s.origins.push_back(l.origin[a]);
    for (int w = 0; w < 8; ++w) s.active.push_back(act[w]);
    if (result_float) s.values.insert(s.values.end(), vals, vals + 512);
    else
      for (int w = 0; w < 8; ++w) s.valmask.push_back(vm[w]);
  });
  sortLeafSet(s);
  return s;
}

// --------------------------------------------------------------------------------------------
// Receiver side of remote mapping.
// applyMapSectionUpdateGrid V:1058-1085 (smoothing = OpenVDB morphology, out of scope): every ACTIVE map voxel inside
// the inclusive index box [floor(bb_min), floor(bb_max)] is deactivated (value kept), then every active voxel of the
// section is activated in the map (accessor setActiveState(true): a missing leaf is created with background values).
// --------------------------------------------------------------------------------------------
static void applySectionUpdate(FloatTree& map, const Coord& mn, const Coord& mx, size_t n, const int32_t* origins, const uint64_t* active)
{
  // step 1: V:1073-1079
  std::vector<FloatLeaf*> leaves;
  map.forEachLeaf([&](const FloatLeaf& l) { leaves.push_back(const_cast<FloatLeaf*>(&l)); });
  for (FloatLeaf* l : leaves)
  {
    bool overlap = true;
    for (int a = 0; a < 3; ++a)
      if (l->origin[a] + 7 < mn[a] || l->origin[a] > mx[a]) overlap = false;
    if (!overlap) continue;
    for (uint32_t k = l->vmask.findNextOn(0); k < 512; k = l->vmask.findNextOn(k + 1))
    {
      Coord loc = leafOffsetToLocal(k);
      if (insideBox(Coord(l->origin[0] + loc[0], l->origin[1] + loc[1], l->origin[2] + loc[2]), mn, mx)) l->vmask.setOff(k);
    }
  }
  // step 2: V:1080-1083
  Accessor<FloatTree> acc(map);
  struct TouchOp { void operator()(float&, bool& a) const { a = true; } };
  for (size_t i = 0; i < n; ++i)
    for (uint32_t k = 0; k < 512; ++k)
    {
      if (!((active[8 * i + (k >> 6)] >> (k & 63)) & 1u)) continue;
      Coord loc = leafOffsetToLocal(k);
      Coord c(origins[3 * i] + loc[0], origins[3 * i + 1] + loc[1], origins[3 * i + 2] + loc[2]);
      // setActiveState(xyz, true) == InternalNode::setActiveStateAndCache: a child is created when the tile state differs
      // (inactive background tile, on = true), then the leaf bit is set; the value is left alone.
      FloatLeaf* l = const_cast<FloatLeaf*>(acc.probeLeaf(c));
      if (!l)
      {
        // create the path down to the leaf with background tiles (same node construction as touchLeaf on the bool tree)
        Coord key = FloatTree::rootKey(c);
        auto it   = map.table.find(key);
        FloatI2* b;
        if (it == map.table.end()) { b = new FloatI2(c, map.background, false); map.table[key] = b; }
        else b = it->second;
        uint32_t nb = FloatI2::coordToOffset(c);
        if (!b->childMask.isOn(nb)) { b->nodes[nb].child = new FloatI1(c, b->nodes[nb].tile, false); b->childMask.setOn(nb); }
        FloatI1* a1 = b->nodes[nb].child;
        uint32_t na = FloatI1::coordToOffset(c);
        if (!a1->childMask.isOn(na)) { a1->nodes[na].child = new FloatLeaf(c, a1->nodes[na].tile, false); a1->childMask.setOn(na); }
        l = a1->nodes[na].child;
        acc.insert(c, l);
      }
      l->vmask.setOn(leafOffset(c));
    }
}

// applyMapSectionGrid V:1022-1047: for (iter = section->cbeginValueAll()): on -> map.setValueOn(coord, v), off ->
// map.setValueOff(coord, v). cbeginValueAll visits EVERY voxel of every section leaf (so a section leaf overwrites the
// map leaf completely, also outside the box it was cut from) AND every tile of the section tree's internal nodes: an
// inactive background tile yields map.setValueOff(tile origin, background), which zeroes + deactivates that one voxel
// wherever the map already has a leaf (setValueOffAndCache only descends into existing children for a background
// value). `tile_quirk` = false skips those tile visits.
static void applySectionGrid(FloatTree& map, size_t n, const int32_t* origins, const uint64_t* active, const float* values, bool tile_quirk)
{
  Accessor<FloatTree> acc(map);
  auto setOffExisting = [&](const Coord& c) {
    FloatLeaf* l = const_cast<FloatLeaf*>(acc.probeLeaf(c));
    if (!l) return;
    const uint32_t k = leafOffset(c);
    l->buf[k] = 0.0f;
    l->vmask.setOff(k);
  };
  if (tile_quirk && n)
  {
    // the section tree as the receiver sees it: exactly the exported leaves
    FloatTree sec(0.0f);
    Accessor<FloatTree> sacc(sec);
    struct Nop { void operator()(float&, bool&) const {} };
    for (size_t i = 0; i < n; ++i)
    {
      Coord c(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
      Coord key = FloatTree::rootKey(c);
      auto it   = sec.table.find(key);
      FloatI2* b;
      if (it == sec.table.end()) { b = new FloatI2(c, 0.0f, false); sec.table[key] = b; }
      else b = it->second;
      uint32_t nb = FloatI2::coordToOffset(c);
      if (!b->childMask.isOn(nb)) { b->nodes[nb].child = new FloatI1(c, 0.0f, false); b->childMask.setOn(nb); }
      FloatI1* a1 = b->nodes[nb].child;
      uint32_t na = FloatI1::coordToOffset(c);
      if (!a1->childMask.isOn(na)) { a1->nodes[na].child = new FloatLeaf(c, 0.0f, false); a1->childMask.setOn(na); }
    }
    for (auto& kv : sec.table)
    {
      const FloatI2* b = kv.second;
      for (uint32_t nb = 0; nb < FloatI2::NUM; ++nb)
      {
        // offset -> coordinate of the slot's origin (InternalNode::offsetToGlobalCoord)
        const Coord o2(b->origin[0] + int32_t((nb >> 10) << 7), b->origin[1] + int32_t(((nb >> 5) & 31) << 7), b->origin[2] + int32_t((nb & 31) << 7));
        if (!b->childMask.isOn(nb)) { setOffExisting(o2); continue; }
        const FloatI1* a1 = b->nodes[nb].child;
        for (uint32_t na = 0; na < FloatI1::NUM; ++na)
        {
          if (a1->childMask.isOn(na)) continue;
          setOffExisting(Coord(a1->origin[0] + int32_t((na >> 8) << 3), a1->origin[1] + int32_t(((na >> 4) & 15) << 3), a1->origin[2] + int32_t((na & 15) << 3)));
        }
      }
    }
  }
  for (size_t i = 0; i < n; ++i)
  {
    const Coord o(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]);
    FloatLeaf* l = const_cast<FloatLeaf*>(acc.probeLeaf(o));
    if (!l)
    {
      // a leaf is created by the first voxel that is active or differs from the background
      bool content = false;
      for (int w = 0; w < 8 && !content; ++w) content = active[8 * i + w] != 0;
      for (int k = 0; k < 512 && !content; ++k) content = values[512 * i + k] != 0.0f;
      if (!content) continue;
      struct Nop { void operator()(float&, bool&) const {} };
      // create the path (values background, inactive); the loop below then writes every voxel
      Coord key = FloatTree::rootKey(o);
      auto it   = map.table.find(key);
      FloatI2* b;
      if (it == map.table.end()) { b = new FloatI2(o, map.background, false); map.table[key] = b; }
      else b = it->second;
      uint32_t nb = FloatI2::coordToOffset(o);
      if (!b->childMask.isOn(nb)) { b->nodes[nb].child = new FloatI1(o, b->nodes[nb].tile, false); b->childMask.setOn(nb); }
      FloatI1* a1 = b->nodes[nb].child;
      uint32_t na = FloatI1::coordToOffset(o);
      if (!a1->childMask.isOn(na)) { a1->nodes[na].child = new FloatLeaf(o, a1->nodes[na].tile, false); a1->childMask.setOn(na); }
      l = a1->nodes[na].child;
      acc.insert(o, l);
    }
    for (uint32_t k = 0; k < 512; ++k)
    {
      l->buf[k] = values[512 * i + k];
      l->vmask.set(k, (active[8 * i + (k >> 6)] >> (k & 63)) & 1u);
    }
  }
}

struct Handle
{
  OccupancyMapping map;
  LeafSet last_export;
  explicit Handle(double res) : map(res) {}
};

} // namespace vo

// ================================================================================================
// C interface (ctypes / dlopen). All arrays are caller-allocated after a count query.
// ================================================================================================
extern "C" {

void* vdbo_create(double resolution) { return new vo::Handle(resolution); }
void vdbo_destroy(void* h) { delete static_cast<vo::Handle*>(h); }

int vdbo_set_config(void* h, double max_range, double prob_hit, double prob_miss, double thres_min, double thres_max)
{
  return static_cast<vo::Handle*>(h)->map.setConfig(max_range, prob_hit, prob_miss, thres_min, thres_max);
}
void vdbo_set_probe_quirk(void* h, int on) { static_cast<vo::Handle*>(h)->map.replicate_probe_quirk = (on != 0); }

// out[6] = hit, miss, thres_min, thres_max, max_logodds, min_logodds
void vdbo_get_logodds(void* h, float* out)
{
  auto& m = static_cast<vo::Handle*>(h)->map;
  out[0]  = m.m_logodds_hit;
  out[1]  = m.m_logodds_miss;
  out[2]  = m.m_logodds_thres_min;
  out[3]  = m.m_logodds_thres_max;
  out[4]  = m.m_max_logodds;
  out[5]  = m.m_min_logodds;
}

void vdbo_add_source(void* h, const char* id, double max_range) { static_cast<vo::Handle*>(h)->map.addInputSource(id, max_range); }
void vdbo_reset(void* h) { static_cast<vo::Handle*>(h)->map.resetMap(); }

void vdbo_world_to_index(void* h, const double* w, int32_t* out)
{
  vo::Coord c = static_cast<vo::Handle*>(h)->map.worldToIndex(w);
  out[0] = c[0]; out[1] = c[1]; out[2] = c[2];
}

// accumulateUpdate: 0 ok, 1 unknown source (no-op), 2 not configured (no-op)
int vdbo_accumulate(void* h, const char* id, const void* pts, uint64_t n, uint64_t stride, const double* origin)
{
  return static_cast<vo::Handle*>(h)->map.accumulateUpdate(static_cast<const uint8_t*>(pts), n, stride, origin, id);
}
void vdbo_integrate(void* h) { static_cast<vo::Handle*>(h)->map.integrateUpdate(); }
// insertPointCloud V:399-406
int vdbo_insert(void* h, const char* id, const void* pts, uint64_t n, uint64_t stride, const double* origin)
{
  auto& m = static_cast<vo::Handle*>(h)->map;
  int rc  = m.accumulateUpdate(static_cast<const uint8_t*>(pts), n, stride, origin, id);
  m.integrateUpdate();
  return rc;
}

// out[6] = rays, nan_skipped, clipped, visits, voxel_updates, state_changes (cumulative)
void vdbo_stats(void* h, uint64_t* out)
{
  auto& m = static_cast<vo::Handle*>(h)->map;
  vo::Stats s = m.stats;
  for (auto& kv : m.m_input_sources)
  {
    s.rays += kv.second->stats.rays; s.nan_skipped += kv.second->stats.nan_skipped;
    s.clipped += kv.second->stats.clipped; s.visits += kv.second->stats.visits;
  }
  out[0] = s.rays; out[1] = s.nan_skipped; out[2] = s.clipped; out[3] = s.visits; out[4] = s.voxel_updates;
  out[5] = s.state_changes;
}

// ---- exports: "prepare" snapshots a canonical (origin-sorted) leaf set and returns its size;
//      "fetch" copies it out. kind: 0 = map (float), 1 = update grid of source, 2 = last change grid
//      of source, 3 = section as UpdateGridT, 4 = section as GridT.
int64_t vdbo_export_prepare(void* hh, int kind, const char* source, const int32_t* bbmin, const int32_t* bbmax, int full)
{
  auto* h = static_cast<vo::Handle*>(hh);
  auto& m = h->map;
  switch (kind)
  {
    case 0: h->last_export = vo::exportFloat(*m.m_vdb_grid); break;
    case 1:
    case 2: {
      auto it = m.m_input_sources.find(source ? source : "");
      if (it == m.m_input_sources.end()) return -1;
      const vo::BoolTree* t = (kind == 1) ? it->second->update_grid.get() : it->second->last_change.get();
      if (kind == 1 && full == 2) t = it->second->reduced.get(); // level-2 reduced update of the last accumulate
      if (!t) { h->last_export = vo::LeafSet(); break; }
      h->last_export = vo::exportBool(*t);
      break;
    }
    case 5: h->last_export = vo::exportBool(*m.m_artificial_area_grid); break;
    case 3:
    case 4:
      h->last_export = vo::mapSection(*m.m_vdb_grid, vo::Coord(bbmin[0], bbmin[1], bbmin[2]),
                                      vo::Coord(bbmax[0], bbmax[1], bbmax[2]), full != 0, kind == 4);
      break;
    default: return -2;
  }
  return int64_t(h->last_export.size());
}

// any pointer may be null. values: 512 f32 per leaf (kinds 0,4); valmask: 8 u64 per leaf (kinds 1,2,3)
void vdbo_export_fetch(void* hh, int32_t* origins, uint64_t* active, uint64_t* valmask, float* values)
{
  auto& s = static_cast<vo::Handle*>(hh)->last_export;
  if (origins && !s.origins.empty()) std::memcpy(origins, s.origins.data(), s.origins.size() * sizeof(int32_t));
  if (active && !s.active.empty()) std::memcpy(active, s.active.data(), s.active.size() * sizeof(uint64_t));
  if (valmask && !s.valmask.empty()) std::memcpy(valmask, s.valmask.data(), s.valmask.size() * sizeof(uint64_t));
  if (values && !s.values.empty()) std::memcpy(values, s.values.data(), s.values.size() * sizeof(float));
}

// point query on the map (GridT::Accessor::getValue / isValueOn); returns active flag
int vdbo_probe(void* hh, const int32_t* c, float* value)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  vo::Accessor<vo::FloatTree> acc(*m.m_vdb_grid);
  vo::Coord cc(c[0], c[1], c[2]);
  const vo::FloatLeaf* l = acc.probeLeaf(cc);
  if (!l) { *value = 0.0f; return 0; }
  uint32_t n = vo::leafOffset(cc);
  *value     = l->buf[n];
  return l->vmask.isOn(n) ? 1 : 0;
}

uint64_t vdbo_map_leaf_count(void* hh) { return static_cast<vo::Handle*>(hh)->map.m_vdb_grid->leafCount(); }

// import an update grid (leaf records) into a source's update grid: used to model
// updateMap(UpdateGridT::Ptr) with a caller-provided grid and the multi-GPU exchange tests.
int vdbo_update_import(void* hh, const char* source, uint64_t n, const int32_t* origins, const uint64_t* active,
                       const uint64_t* valmask)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  auto it = m.m_input_sources.find(source);
  if (it == m.m_input_sources.end()) return 1;
  vo::Accessor<vo::BoolTree> acc(*it->second->update_grid);
  for (uint64_t i = 0; i < n; ++i)
  {
    bool any = false;
    for (int w = 0; w < 8; ++w) any |= (active[8 * i + w] != 0);
    if (!any) continue;
    vo::BoolLeaf* l = acc.touchLeaf(vo::Coord(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]));
    for (int w = 0; w < 8; ++w)
    {
      l->vmask.w[w] |= active[8 * i + w];
      l->buf.w[w] |= valmask[8 * i + w];
    }
  }
  return 0;
}

int vdbo_apply_section_update(void* hh, const int32_t* bbmin, const int32_t* bbmax, uint64_t n, const int32_t* origins, const uint64_t* active)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  vo::applySectionUpdate(*m.m_vdb_grid, vo::Coord(bbmin[0], bbmin[1], bbmin[2]), vo::Coord(bbmax[0], bbmax[1], bbmax[2]), n, origins, active);
  return 0;
}
int vdbo_apply_section_grid(void* hh, uint64_t n, const int32_t* origins, const uint64_t* active, const float* values, int tile_quirk)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  vo::applySectionGrid(*m.m_vdb_grid, n, origins, active, values, tile_quirk != 0);
  return 0;
}

static void leafsetToTree(vo::BoolTree& t, uint64_t n, const int32_t* origins, const uint64_t* active, const uint64_t* valmask)
{
  vo::Accessor<vo::BoolTree> acc(t);
  for (uint64_t i = 0; i < n; ++i)
  {
    bool any = false;
    for (int w = 0; w < 8; ++w) any |= (active[8 * i + w] != 0);
    if (!any) continue;
    vo::BoolLeaf* l = acc.touchLeaf(vo::Coord(origins[3 * i], origins[3 * i + 1], origins[3 * i + 2]));
    for (int w = 0; w < 8; ++w)
    {
      l->vmask.w[w] |= active[8 * i + w];
      l->buf.w[w] |= valmask[8 * i + w];
    }
  }
}

// createUpdate's counterpart on the receiving map. level 0: the grid is a raw update grid -> updateMap;
// level 1: it is an overwrite (change) grid -> overwriteMap; level 2: it is a reduced grid -> re-raycast from `origin`,
// then updateMap. The change grid of levels 0/2 is kept as the named source's last change grid.
int vdbo_update_apply(void* hh, const char* source, int level, uint64_t n, const int32_t* origins, const uint64_t* active,
                      const uint64_t* valmask, const double* origin)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  auto it = m.m_input_sources.find(source ? source : "");
  if (it == m.m_input_sources.end()) return 1;
  vo::BoolTree in(false);
  leafsetToTree(in, n, origins, active, valmask);
  if (level == 0) it->second->last_change = m.updateMap(in);
  else if (level == 1) m.overwriteMap(in);
  else if (level == 2)
  {
    vo::BoolTree full(false);
    vo::Accessor<vo::BoolTree> acc(full);
    m.raycastReducedGrid(in, origin, acc);
    it->second->last_change = m.updateMap(full);
  }
  else return 2;
  return 0;
}
void vdbo_last_origin(void* hh, const char* source, double* out)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  auto it = m.m_input_sources.find(source ? source : "");
  if (it == m.m_input_sources.end()) return;
  for (int a = 0; a < 3; ++a) out[a] = it->second->last_origin[a];
}
// addPointsToGrid (occupied != 0) / removePointsFromGrid
void vdbo_points_set(void* hh, const void* pts, uint64_t n, uint64_t stride, int occupied)
{
  static_cast<vo::Handle*>(hh)->map.setPoints(static_cast<const uint8_t*>(pts), n, stride, occupied != 0);
}
void vdbo_add_artificial_areas(void* hh, uint64_t n_poly, const uint32_t* counts, const double* xyz, double negative_height, double positive_height)
{
  static_cast<vo::Handle*>(hh)->map.addArtificialAreas(n_poly, counts, xyz, negative_height, positive_height);
}
void vdbo_set_fast_mode(void* hh, int on) { static_cast<vo::Handle*>(hh)->map.m_fast_mode = (on != 0); }
void vdbo_raytrace(void* hh, uint64_t n, const double* origins, const double* directions, const double* max_lengths, int32_t* successes,
                   double* end_points)
{
  static_cast<vo::Handle*>(hh)->map.raytrace(n, origins, directions, max_lengths, successes, end_points);
}
// addArtificialWall V:1217-1236 on its own (no restoreMapIntegrity)
void vdbo_add_artificial_wall(void* hh, const double* start, const double* end, double negative_height, double positive_height)
{
  static_cast<vo::Handle*>(hh)->map.addArtificialWall(start, end, negative_height, positive_height);
}
// castRayIntoGrid V:550-566 for explicit voxel pairs into a source's update grid
int vdbo_cast_index_rays(void* hh, const char* source, uint64_t n, const int32_t* rays6)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  auto it = m.m_input_sources.find(source);
  if (it == m.m_input_sources.end()) return 1;
  vo::Accessor<vo::BoolTree> acc(*it->second->update_grid);
  vo::Stats st;
  for (uint64_t i = 0; i < n; ++i)
    m.castRayIntoGrid(vo::Coord(rays6[6 * i], rays6[6 * i + 1], rays6[6 * i + 2]), vo::Coord(rays6[6 * i + 3], rays6[6 * i + 4], rays6[6 * i + 5]), acc, st);
  return 0;
}
void vdbo_restore_map_integrity(void* hh) { static_cast<vo::Handle*>(hh)->map.restoreMapIntegrity(); }

// empty one source's update grid without applying it (models vdbm_update_partition, which hands the leaves over)
int vdbo_update_clear(void* hh, const char* source)
{
  auto& m = static_cast<vo::Handle*>(hh)->map;
  auto it = m.m_input_sources.find(source);
  if (it == m.m_input_sources.end()) return 1;
  it->second->update_grid.reset(new vo::BoolTree(false));
  return 0;
}

} // extern "C"
