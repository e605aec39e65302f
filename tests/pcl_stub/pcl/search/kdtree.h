// pcl::search::KdTree stand-in: same virtual interface as PCL 1.10's, exhaustive search inside (tests/pcl_stub/README.md)
#pragma once
#include <algorithm>
#include <utility>
#include <vector>
#include <pcl/point_cloud.h>
namespace pcl {
namespace search {
template <typename PointT>
class KdTree {
 public:
  using Ptr = boost::shared_ptr<KdTree<PointT>>;
  using ConstPtr = boost::shared_ptr<const KdTree<PointT>>;
  using PointCloudConstPtr = typename PointCloud<PointT>::ConstPtr;
  using IndicesConstPtr = boost::shared_ptr<const std::vector<int>>;
  virtual ~KdTree() {}
  virtual void setInputCloud(const PointCloudConstPtr& cloud, const IndicesConstPtr& = IndicesConstPtr()) { input_ = cloud; builds_++; }
  virtual PointCloudConstPtr getInputCloud() const { return input_; }
  virtual int nearestKSearch(const PointT& p, int k, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances) const {
    k_indices.clear(); k_sqr_distances.clear();
    if (!input_) return 0;
    std::vector<std::pair<float, int>> d(input_->size());
    for (std::size_t i = 0; i < input_->size(); i++) {
      const PointT& q = (*input_)[i];
      float dx = p.x - q.x, dy = p.y - q.y, dz = p.z - q.z;
      d[i] = {(dx * dx + dy * dy) + dz * dz, (int)i};
    }
    std::size_t kk = std::min<std::size_t>((std::size_t)k, d.size());
    std::partial_sort(d.begin(), d.begin() + kk, d.end());
    for (std::size_t i = 0; i < kk; i++) { k_indices.push_back(d[i].second); k_sqr_distances.push_back(d[i].first); }
    return (int)kk;
  }
  virtual int radiusSearch(const PointT&, double, std::vector<int>& k_indices, std::vector<float>& k_sqr_distances, unsigned int = 0) const {
    k_indices.clear(); k_sqr_distances.clear();
    return 0;
  }
  int builds() const { return builds_; }       // mock-only: how often an index was built
 protected:
  PointCloudConstPtr input_;
  int builds_ = 0;
};
}  // namespace search
}  // namespace pcl
