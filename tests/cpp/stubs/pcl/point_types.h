// Stand-in with the API shape of <pcl/point_types.h> (tests/cpp/stubs/README.md).
#pragma once
namespace pcl {
struct alignas(16) PointXYZ
{
  float x, y, z;
  float data_c_pad = 1.0f; // PCL's PointXYZ is a 16-byte record: {x, y, z, 1.0f}
  PointXYZ() : x(0), y(0), z(0) {}
  PointXYZ(float x_, float y_, float z_) : x(x_), y(y_), z(z_) {}
};
static_assert(sizeof(PointXYZ) == 16, "pcl::PointXYZ is a 16-byte record");
} // namespace pcl
