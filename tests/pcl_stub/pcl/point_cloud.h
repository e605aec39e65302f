#pragma once
#include <cstddef>
#include <string>
#include <vector>
#include <boost/shared_ptr.hpp>
namespace pcl {
struct PCLHeader { unsigned seq = 0; unsigned long long stamp = 0; std::string frame_id; };
template <typename PointT>
class PointCloud {
 public:
  using Ptr = boost::shared_ptr<PointCloud<PointT>>;
  using ConstPtr = boost::shared_ptr<const PointCloud<PointT>>;
  PCLHeader header;
  std::vector<PointT> points;
  unsigned width = 0, height = 1;
  bool is_dense = true;
  std::size_t size() const { return points.size(); }
  bool empty() const { return points.empty(); }
  void resize(std::size_t n) { points.resize(n); width = (unsigned)n; height = 1; }
  void push_back(const PointT& p) { points.push_back(p); width = (unsigned)points.size(); }
  PointT& operator[](std::size_t i) { return points[i]; }
  const PointT& operator[](std::size_t i) const { return points[i]; }
};
}  // namespace pcl
