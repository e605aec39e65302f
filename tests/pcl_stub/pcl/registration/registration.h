// pcl::Registration stand-in (tests/pcl_stub/README.md): the members, setters and the align() / initCompute() control
// flow of PCL 1.10's pcl/registration/registration.h + impl/registration.hpp that a subclass relies on.
#pragma once
#include <limits>
#include <string>
#include <vector>
#include <Eigen/Core>
#include <pcl/console/print.h>
#include <pcl/point_cloud.h>
#include <pcl/search/kdtree.h>
namespace pcl {
template <typename PointSource, typename PointTarget, typename Scalar = float>
class Registration {
 public:
  using Matrix4 = Eigen::Matrix4f;
  using Ptr = boost::shared_ptr<Registration<PointSource, PointTarget, Scalar>>;
  using KdTree = pcl::search::KdTree<PointTarget>;
  using KdTreePtr = typename KdTree::Ptr;
  using PointCloudSource = pcl::PointCloud<PointSource>;
  using PointCloudSourcePtr = typename PointCloudSource::Ptr;
  using PointCloudSourceConstPtr = typename PointCloudSource::ConstPtr;
  using PointCloudTarget = pcl::PointCloud<PointTarget>;
  using PointCloudTargetPtr = typename PointCloudTarget::Ptr;
  using PointCloudTargetConstPtr = typename PointCloudTarget::ConstPtr;

  Registration() : tree_(new KdTree) {
    final_transformation_.setIdentity(); transformation_.setIdentity(); previous_transformation_.setIdentity();
  }
  virtual ~Registration() {}

  virtual void setInputSource(const PointCloudSourceConstPtr& cloud) { input_ = cloud; source_cloud_updated_ = true; }
  virtual void setInputTarget(const PointCloudTargetConstPtr& cloud) {
    if (cloud->points.empty()) { PCL_ERROR("[pcl::%s::setInputTarget] Invalid or empty point cloud dataset given!\n", getClassName().c_str()); return; }
    target_ = cloud; target_cloud_updated_ = true;
  }
  PointCloudSourceConstPtr const getInputSource() { return input_; }
  PointCloudTargetConstPtr const getInputTarget() { return target_; }
  void setSearchMethodTarget(const KdTreePtr& tree, bool force_no_recompute = false) {
    tree_ = tree;
    if (force_no_recompute) force_no_recompute_ = true;
    target_cloud_updated_ = true;
  }
  KdTreePtr getSearchMethodTarget() const { return tree_; }
  Matrix4 getFinalTransformation() { return final_transformation_; }
  Matrix4 getLastIncrementalTransformation() { return transformation_; }
  void setMaximumIterations(int n) { max_iterations_ = n; }
  int getMaximumIterations() { return max_iterations_; }
  void setRANSACIterations(int n) { ransac_iterations_ = n; }
  void setMaxCorrespondenceDistance(double d) { corr_dist_threshold_ = d; }
  double getMaxCorrespondenceDistance() { return corr_dist_threshold_; }
  void setTransformationEpsilon(double e) { transformation_epsilon_ = e; }
  double getTransformationEpsilon() { return transformation_epsilon_; }
  void setEuclideanFitnessEpsilon(double e) { euclidean_fitness_epsilon_ = e; }
  bool hasConverged() const { return converged_; }
  const std::string& getClassName() const { return reg_name_; }

  double getFitnessScore(double max_range = std::numeric_limits<double>::max()) {
    double sum = 0.0; int nr = 0;
    std::vector<int> idx(1); std::vector<float> d2(1);
    for (std::size_t i = 0; i < input_->size(); i++) {
      PointSource p = (*input_)[i];
      const float x = p.x, y = p.y, z = p.z;
      p.x = final_transformation_(0, 0) * x + final_transformation_(0, 1) * y + final_transformation_(0, 2) * z + final_transformation_(0, 3);
      p.y = final_transformation_(1, 0) * x + final_transformation_(1, 1) * y + final_transformation_(1, 2) * z + final_transformation_(1, 3);
      p.z = final_transformation_(2, 0) * x + final_transformation_(2, 1) * y + final_transformation_(2, 2) * z + final_transformation_(2, 3);
      tree_->nearestKSearch(p, 1, idx, d2);
      if (!idx.empty() && d2[0] <= max_range) { sum += d2[0]; nr++; }
    }
    return nr > 0 ? sum / nr : std::numeric_limits<double>::max();
  }

  void align(PointCloudSource& output) { align(output, Matrix4::Identity()); }
  void align(PointCloudSource& output, const Matrix4& guess) {
    if (!initCompute()) return;
    output.points.resize(input_->size());
    output.header = input_->header; output.width = (unsigned)input_->size(); output.height = 1; output.is_dense = input_->is_dense;
    for (std::size_t i = 0; i < input_->size(); i++) output.points[i] = input_->points[i];
    converged_ = false;
    final_transformation_ = transformation_ = previous_transformation_ = Matrix4::Identity();
    for (std::size_t i = 0; i < output.size(); i++) output.points[i].data[3] = 1.0f;
    computeTransformation(output, guess);
  }

 protected:
  bool initCompute() {
    if (!target_) { PCL_ERROR("[pcl::registration::%s::compute] No input target dataset was given!\n", getClassName().c_str()); return false; }
    if (!input_) { PCL_ERROR("[pcl::registration::%s::compute] No input source dataset was given!\n", getClassName().c_str()); return false; }
    if (target_cloud_updated_ && !force_no_recompute_) {      // the per-scan FLANN build a GPU subclass must avoid
      tree_->setInputCloud(target_);
      target_cloud_updated_ = false;
    }
    return true;
  }
  virtual void computeTransformation(PointCloudSource& output, const Matrix4& guess) = 0;

  std::string reg_name_;
  KdTreePtr tree_;
  int nr_iterations_ = 0, max_iterations_ = 10, ransac_iterations_ = 0;
  PointCloudSourceConstPtr input_;
  PointCloudTargetConstPtr target_;
  Matrix4 final_transformation_, transformation_, previous_transformation_;
  double transformation_epsilon_ = 0.0, transformation_rotation_epsilon_ = 0.0, euclidean_fitness_epsilon_ = 0.0;
  double corr_dist_threshold_ = 1e30;
  bool converged_ = false;
  bool target_cloud_updated_ = true, source_cloud_updated_ = true, force_no_recompute_ = false;
};
}  // namespace pcl
