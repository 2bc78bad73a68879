// Stand-in with the API shape of <openvdb/io/Stream.h> (tests/cpp/stubs/README.md).
#pragma once
#include <iostream>
#include <openvdb/openvdb.h>
namespace openvdb {
namespace io {
class Stream
{
public:
  explicit Stream(std::istream& is) : grids_(stub_io::readGrids(is)) {} // reads all grids at construction, like OpenVDB
  explicit Stream(std::ostream& os) : os_(&os), grids_(new GridPtrVec) {}
  GridPtrVecPtr getGrids() { return grids_; }
  template <typename GridPtrContainerT>
  void write(const GridPtrContainerT& grids) const
  {
    if (os_) stub_io::writeGrids(*os_, GridPtrVec(grids.begin(), grids.end()));
  }

private:
  std::ostream* os_ = nullptr;
  GridPtrVecPtr grids_;
};
} // namespace io
} // namespace openvdb
