// pcl::PointXYZI / pcl::PointXYZINormal stand-ins with PCL's memory layout (tests/pcl_stub/README.md)
#pragma once
namespace pcl {
struct alignas(16) PointXYZI {
  union { float data[4]; struct { float x, y, z; }; };
  union { struct { float intensity; }; float data_c[4]; };
};
struct alignas(16) PointXYZINormal {          // 48 bytes: data[4], data_n[4], intensity, curvature, padding
  union { float data[4]; struct { float x, y, z; }; };
  union { float data_n[4]; float normal[3]; struct { float normal_x, normal_y, normal_z; }; };
  union { struct { float intensity; float curvature; }; float data_c[4]; };
};
static_assert(sizeof(PointXYZINormal) == 48, "PointXYZINormal layout");
}  // namespace pcl
