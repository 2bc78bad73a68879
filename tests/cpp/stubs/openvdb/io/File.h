// Stand-in with the API shape of <openvdb/io/File.h> (tests/cpp/stubs/README.md): a flat byte stream, not the .vdb format.
#pragma once
#include <fstream>
#include <string>
#include <openvdb/openvdb.h>
namespace openvdb {
namespace io {
class File
{
public:
  class NameIterator
  {
  public:
    NameIterator(const GridPtrVec* grids, std::size_t i) : grids_(grids), i_(i) {}
    bool operator==(const NameIterator& o) const { return i_ == o.i_; }
    bool operator!=(const NameIterator& o) const { return i_ != o.i_; }
    NameIterator& operator++() { ++i_; return *this; }
    std::string gridName() const { return (*grids_)[i_]->getName(); }
    std::size_t stubIndex() const { return i_; }

  private:
    const GridPtrVec* grids_;
    std::size_t i_;
  };
  explicit File(const std::string& filename) : filename_(filename), grids_(new GridPtrVec) {}
  const std::string& filename() const { return filename_; }
  bool open()
  {
    std::ifstream f(filename_, std::ios::binary);
    if (!f) throw std::runtime_error("openvdb stub: could not open " + filename_); // OpenVDB throws IoError
    grids_ = stub_io::readGrids(f);
    open_  = true;
    return true;
  }
  bool isOpen() const { return open_; }
  void close() { open_ = false; }
  NameIterator beginName() const { return NameIterator(grids_.get(), 0); }
  NameIterator endName() const { return NameIterator(grids_.get(), grids_->size()); }
  GridBase::Ptr readGrid(const std::string& name)
  {
    GridBase::Ptr found;
    for (auto& g : *grids_)
      if (g->getName() == name) found = g; // unnamed grids all match "": the last one wins, like repeated reads of ""
    return found;
  }
  GridPtrVecPtr getGrids() const { return grids_; }
  template <typename GridPtrContainerT>
  void write(const GridPtrContainerT& grids) const
  {
    std::ofstream f(filename_, std::ios::binary);
    if (!f) throw std::runtime_error("openvdb stub: could not write " + filename_);
    stub_io::writeGrids(f, GridPtrVec(grids.begin(), grids.end()));
  }

private:
  std::string filename_;
  GridPtrVecPtr grids_;
  bool open_ = false;
};
} // namespace io
} // namespace openvdb
