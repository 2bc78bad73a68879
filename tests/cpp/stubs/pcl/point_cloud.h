// Stand-in with the API shape of <pcl/point_cloud.h> (tests/cpp/stubs/README.md).
#pragma once
#include <cstdint>
#include <memory>
#include <vector>
namespace pcl {
template <typename PointT>
class PointCloud
{
public:
  using Ptr      = std::shared_ptr<PointCloud<PointT> >; // boost::shared_ptr before PCL 1.11
  using ConstPtr = std::shared_ptr<const PointCloud<PointT> >;
  std::vector<PointT> points; // PCL uses an Eigen::aligned_allocator; the element layout is the same
  std::uint32_t width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void push_back(const PointT& p) { points.push_back(p); }
  typename std::vector<PointT>::const_iterator begin() const { return points.begin(); }
  typename std::vector<PointT>::const_iterator end() const { return points.end(); }
};
} // namespace pcl
