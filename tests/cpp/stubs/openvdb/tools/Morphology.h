// Stand-in with the API shape of <openvdb/tools/Morphology.h> (tests/cpp/stubs/README.md): dilation / erosion of the ACTIVE
// state over the 6 / 18 / 26 neighbourhood; no tiles exist in the stand-in tree, so the tile policy has nothing to do.
#pragma once
#include <set>
#include <vector>
#include <openvdb/openvdb.h>
namespace openvdb {
namespace tools {
enum NearestNeighbors
{
  NN_FACE             = 6,
  NN_FACE_EDGE        = 18,
  NN_FACE_EDGE_VERTEX = 26
};
enum TilePolicy
{
  IGNORE_TILES,
  EXPAND_TILES,
  PRESERVE_TILES
};
namespace stub_morph {
inline bool inStencil(int dx, int dy, int dz, NearestNeighbors nn)
{
  const int k = (dx != 0) + (dy != 0) + (dz != 0);
  return k != 0 && (nn == NN_FACE_EDGE_VERTEX || (nn == NN_FACE_EDGE && k <= 2) || (nn == NN_FACE && k == 1));
}
template <typename TreeT>
std::vector<Coord> activeVoxels(const TreeT& tree)
{
  std::vector<Coord> out;
  for (auto& kv : tree.stubLeaves())
    for (Index n = 0; n < TreeT::LeafNodeType::NUM_VALUES; ++n)
      if (kv.second->isValueOn(n)) out.push_back(kv.second->offsetToGlobalCoord(n));
  return out;
}
} // namespace stub_morph

template <typename TreeT>
void dilateActiveValues(TreeT& tree, int iterations = 1, NearestNeighbors nn = NN_FACE, TilePolicy = PRESERVE_TILES, bool /*threaded*/ = true)
{
  for (int it = 0; it < iterations; ++it)
  {
    const std::vector<Coord> on = stub_morph::activeVoxels(tree);
    for (const Coord& c : on)
      for (int dx = -1; dx <= 1; ++dx)
        for (int dy = -1; dy <= 1; ++dy)
          for (int dz = -1; dz <= 1; ++dz)
            if (stub_morph::inStencil(dx, dy, dz, nn)) tree.touchLeaf(c.offsetBy(dx, dy, dz))->setActiveState(c.offsetBy(dx, dy, dz), true);
  }
}

template <typename TreeT>
void erodeActiveValues(TreeT& tree, int iterations = 1, NearestNeighbors nn = NN_FACE, TilePolicy = PRESERVE_TILES, bool /*threaded*/ = true)
{
  for (int it = 0; it < iterations; ++it)
  {
    const std::vector<Coord> on = stub_morph::activeVoxels(tree);
    std::vector<Coord> off;
    for (const Coord& c : on)
    {
      bool keep = true;
      for (int dx = -1; dx <= 1 && keep; ++dx)
        for (int dy = -1; dy <= 1 && keep; ++dy)
          for (int dz = -1; dz <= 1 && keep; ++dz)
            if (stub_morph::inStencil(dx, dy, dz, nn) && !tree.isValueOn(c.offsetBy(dx, dy, dz))) keep = false;
      if (!keep) off.push_back(c);
    }
    for (const Coord& c : off) tree.touchLeaf(c)->setActiveState(c, false);
  }
}
} // namespace tools
} // namespace openvdb
