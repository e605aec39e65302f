// b200_gicp_pcl.hpp -- pcl::Registration subclass backed by liblocus_b200.so.
// Compiled ONLY inside a LOCUS catkin workspace (needs PCL + frontend_utils); not built in
// this repository's image (no PCL/ROS here).  See INTEGRATION.md.
//
// Drop-in for pcl::MultithreadedGeneralizedIterativeClosestPoint<PointF, PointF>
// (multithreaded_gicp/include/multithreaded_gicp/gicp.h:64-426): same setters, and
// computeTransformation() -- the one pure virtual of pcl::Registration -- forwards to the C ABI.
#pragma once

#include <frontend_utils/CommonStructs.h>   // PointF = pcl::PointXYZINormal
#include <pcl/registration/registration.h>

#include "locus_b200.h"

namespace pcl {

class B200GeneralizedIterativeClosestPoint : public Registration<PointF, PointF> {
 public:
  using Base = Registration<PointF, PointF>;
  using Ptr = boost::shared_ptr<B200GeneralizedIterativeClosestPoint>;

  B200GeneralizedIterativeClosestPoint(int device = 0) {
    reg_name_ = "B200GeneralizedIterativeClosestPoint";
    if (lb_gicp_create(device, &h_) != LB_OK) throw std::runtime_error(lb_last_error_string());
    lb_gicp_default_params(&p_);
    // same defaults as gicp.h:111-127
    max_iterations_ = 200; transformation_epsilon_ = 5e-4; corr_dist_threshold_ = 5.0;
    p_.recompute_source_covariance = 0; p_.recompute_target_covariance = 0;   // reference default (gicp.h:115-116)
  }
  ~B200GeneralizedIterativeClosestPoint() override { lb_gicp_destroy(h_); }

  // extra setters of the reference class (gicp.h:134-143,264-298)
  void setNumThreads(int n) { p_.num_threads = n; }
  void enableTimingOutput(bool e) { p_.enable_timing_output = e; }
  void RecomputeTargetCovariance(bool r) { p_.recompute_target_covariance = r; }
  void RecomputeSourceCovariance(bool r) { p_.recompute_source_covariance = r; }
  void setMaximumOptimizerIterations(int n) { p_.max_optimizer_iterations = n; }
  int getMaximumOptimizerIterations() { return p_.max_optimizer_iterations; }
  void setRotationEpsilon(double e) { p_.rotation_epsilon = e; }
  double getRotationEpsilon() { return p_.rotation_epsilon; }
  void setCorrespondenceRandomness(int k) { p_.k_correspondences = k; }
  int getCorrespondenceRandomness() { return p_.k_correspondences; }

  // gicp.h:162-179 / 196-200
  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (cloud->points.empty()) {
      PCL_ERROR("[pcl::%s::setInputSource] Invalid or empty point cloud dataset given!\n", getClassName().c_str());
      return;
    }
    Base::setInputSource(cloud);
    lb_gicp_set_source(h_, cloud->points.data(), cloud->size(), sizeof(PointF), offsetof(PointF, x),
                       offsetof(PointF, normal_x), LB_MEM_HOST);
  }
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    Base::setInputTarget(cloud);
    lb_gicp_set_target(h_, cloud->points.data(), cloud->size(), sizeof(PointF), offsetof(PointF, x),
                       offsetof(PointF, normal_x), LB_MEM_HOST, nullptr);
  }

 protected:
  // gicp.hpp:405-617
  void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) override {
    p_.transformation_epsilon = transformation_epsilon_;
    p_.max_correspondence_distance = corr_dist_threshold_;
    p_.max_iterations = max_iterations_;
    lb_gicp_set_params(h_, &p_);
    float g[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) g[r * 4 + c] = guess(r, c);
    lb_gicp_result res;
    if (lb_gicp_align(h_, g, &res) != LB_OK) {
      PCL_ERROR("[pcl::%s::computeTransformation] %s\n", getClassName().c_str(), lb_last_error_string());
      return;   // converged_ stays false, final_transformation_ stays the last good one
    }
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) final_transformation_(r, c) = res.final_transformation[r * 4 + c];
    previous_transformation_ = transformation_ = final_transformation_;
    nr_iterations_ = res.iterations;
    converged_ = res.converged != 0;
    // output = final_transformation_ * input (gicp.hpp:586); non-geometric fields were copied by align()
    lb_gicp_transform_source(h_, nullptr, output.points.data(), sizeof(PointF), offsetof(PointF, x),
                             LB_NO_NORMALS, LB_MEM_HOST);
  }

 private:
  lb_gicp* h_ = nullptr;
  lb_gicp_params p_;
};

}  // namespace pcl
