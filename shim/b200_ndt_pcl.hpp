// b200_ndt_pcl.hpp -- pcl::Registration subclass for LOCUS's `registration_method: ndt`, backed by liblocus_b200.so.
// Like b200_gicp_pcl.hpp it is meant for a LOCUS catkin workspace and is compiled here against the PCL mock under
// tests/pcl_stub/ (tests/test_cabi_cpu.py) and run on the GPU through pcl::Registration::align() by tests/shim_harness.cpp
// (tests/test_ndt_gpu.py::test_pcl_ndt_shim_runs).  See INTEGRATION.md 1b.
//
// Drop-in for pclomp::NormalDistributionsTransform<PointF, PointF>
// (multithreaded_gicp/include/multithreaded_ndt/ndt_omp.h:80-537) as PointCloudOdometry.cc:182-195 and
// PointCloudLocalization.cc:267-280 set it up: same setters; computeTransformation() forwards to lb_ndt_align.
//  * setInputTarget uploads the cloud (the voxel Gaussians are built on the GPU at the next align) and, like the GICP shim,
//    installs a lazy search tree so that pcl::Registration::initCompute() builds no FLANN index per scan.
//  * a refused cloud (empty, non-finite source point, voxel index overflow) is reported with PCL_ERROR and leaves both the
//    PCL base and the device handle on their previous input.
//  * `output` = final_transformation_ * input, as the reference leaves it (its last trans_cloud, ndt_omp_impl.hpp:980-986).
//  * transformation_ / previous_transformation_ (the reference's last Newton step as a matrix, ndt_omp_impl.hpp:177-186) are
//    not reported by the C ABI and stay at the identity; LOCUS reads getFinalTransformation() only.
#pragma once

#include <stdexcept>

#include "b200_gicp_pcl.hpp"     // PointF, B200LazyKdTree, locus_b200.h

namespace pcl {

class B200NormalDistributionsTransform : public Registration<PointF, PointF> {
 public:
  using Base = Registration<PointF, PointF>;
  using Ptr = boost::shared_ptr<B200NormalDistributionsTransform>;

  explicit B200NormalDistributionsTransform(int device = 0) {
    reg_name_ = "B200NormalDistributionsTransform";
    if (lb_ndt_create(device, &h_) != LB_OK) throw std::runtime_error(lb_last_error_string());
    lb_ndt_default_params(&p_);
    transformation_epsilon_ = 0.1; max_iterations_ = 35;                  // ndt_omp_impl.hpp:93-94
    lazy_tree_.reset(new B200LazyKdTree<PointF>());
    setSearchMethodTarget(lazy_tree_, /*force_no_recompute=*/true);
  }
  ~B200NormalDistributionsTransform() override { lb_ndt_destroy(h_); }

  // ndt_omp.h:124-196
  void setResolution(float r) { p_.resolution = r; }                      // takes effect at the next align (the voxels are rebuilt)
  float getResolution() const { return p_.resolution; }
  void setStepSize(double s) { p_.step_size = s; }
  double getStepSize() const { return p_.step_size; }
  void setOulierRatio(double r) { p_.outlier_ratio = r; }
  double getOulierRatio() const { return p_.outlier_ratio; }
  void setNeighborhoodSearchMethod(int m) { p_.search_method = m; }       // pclomp::KDTREE 0, DIRECT26 1, DIRECT7 2, DIRECT1 3
  void setNumThreads(int n) { p_.num_threads = n; }
  void enableTimingOutput(bool e) { p_.enable_timing_output = e; }
  double getTransformationProbability() const { return res_.trans_probability; }
  int getFinalNumIteration() const { return res_.nr_iterations; }

  void setInputSource(const PointCloudSourceConstPtr& cloud) override {
    if (!cloud || cloud->points.empty()) {
      PCL_ERROR("[pcl::%s::setInputSource] Invalid or empty point cloud dataset given!\n", getClassName().c_str());
      return;
    }
    if (lb_ndt_set_source(h_, cloud->points.data(), cloud->size(), sizeof(PointF), offsetof(PointF, x), LB_MEM_HOST) != LB_OK) {
      PCL_ERROR("[pcl::%s::setInputSource] %s -- previous input kept\n", getClassName().c_str(), lb_last_error_string());
      return;
    }
    Base::setInputSource(cloud);
  }
  // ndt_omp.h:116-119: setInputTarget + init()
  void setInputTarget(const PointCloudTargetConstPtr& cloud) override {
    if (!cloud || cloud->points.empty()) {
      PCL_ERROR("[pcl::%s::setInputTarget] Invalid or empty point cloud dataset given!\n", getClassName().c_str());
      return;
    }
    if (!apply()) return;                                                  // the lattice check uses the current resolution
    if (lb_ndt_set_target(h_, cloud->points.data(), cloud->size(), sizeof(PointF), offsetof(PointF, x), LB_MEM_HOST) != LB_OK) {
      PCL_ERROR("[pcl::%s::setInputTarget] %s -- previous input kept\n", getClassName().c_str(), lb_last_error_string());
      return;
    }
    Base::setInputTarget(cloud);
    lazy_tree_->setInputCloud(cloud);
  }

  const lb_ndt_result& getLastResult() const { return res_; }
  lb_ndt* handle() { return h_; }

 protected:
  // ndt_omp_impl.hpp:100-208
  void computeTransformation(PointCloudSource& output, const Eigen::Matrix4f& guess) override {
    if (!apply()) return;
    float g[16];
    for (int r = 0; r < 4; r++) for (int c = 0; c < 4; c++) g[r * 4 + c] = guess(r, c);
    if (lb_ndt_align(h_, g, &res_) != LB_OK) {
      PCL_ERROR("[pcl::%s::computeTransformation] %s\n", getClassName().c_str(), lb_last_error_string());
      return;
    }
    for (int r = 0; r < 4; r++)
      for (int c = 0; c < 4; c++) final_transformation_(r, c) = res_.final_transformation[r * 4 + c];
    nr_iterations_ = res_.nr_iterations;
    converged_ = res_.converged != 0;
    const float* T = res_.final_transformation;                            // pcl::transformPointCloud's association
    for (std::size_t i = 0; i < output.size() && i < input_->size(); i++) {
      const PointF& p = input_->points[i];
      output.points[i].x = T[0] * p.x + (T[1] * p.y + (T[2] * p.z + T[3]));
      output.points[i].y = T[4] * p.x + (T[5] * p.y + (T[6] * p.z + T[7]));
      output.points[i].z = T[8] * p.x + (T[9] * p.y + (T[10] * p.z + T[11]));
    }
  }

 private:
  bool apply() {
    p_.transformation_epsilon = transformation_epsilon_;
    p_.max_iterations = max_iterations_;
    p_.max_correspondence_distance = corr_dist_threshold_;
    p_.ransac_iterations = ransac_iterations_;
    if (lb_ndt_set_params(h_, &p_) != LB_OK) {
      PCL_ERROR("[pcl::%s] %s\n", getClassName().c_str(), lb_last_error_string());
      return false;
    }
    return true;
  }

  lb_ndt* h_ = nullptr;
  lb_ndt_params p_;
  lb_ndt_result res_{};
  boost::shared_ptr<B200LazyKdTree<PointF>> lazy_tree_;
};

}  // namespace pcl
