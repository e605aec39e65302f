// b200_normal_computation_nodelet.cc -- nodelet point_cloud_filter/B200NormalComputation.
// Compiled ONLY inside a LOCUS catkin workspace (needs ROS + pcl_ros); see INTEGRATION.md.
// Same virtuals and dynamic-reconfigure fields as point_cloud_filter::NormalComputation
// (point_cloud_filter/src/normal_computation.cc:12-85); norm_est_ (pcl::NormalEstimationOMP<PointXYZI, Normal>)
// is replaced by lb_gicp_compute_normals (search_knn) / lb_gicp_compute_normals_radius (search_radius, including the
// removeNaNNormalsFromPointCloud that follows it, normal_computation.cc:53-57) on the cloud's voxel-hash index.
#include <pluginlib/class_list_macros.h>
#include <pcl/point_types.h>
#include <pcl_conversions/pcl_conversions.h>
#include <pcl_ros/filters/filter.h>
#include <point_cloud_filter/NormalComputationConfig.h>

#include "locus_b200.h"

namespace point_cloud_filter {

class B200NormalComputation : public pcl_ros::Filter {
 protected:
  boost::shared_ptr<dynamic_reconfigure::Server<NormalComputationConfig>> srv_;
  lb_gicp* impl_ = nullptr;      // used as a cloud index + k-NN engine only; align() is never called on it
  int k_ = 20;
  double radius_ = 0.3;
  bool radius_mode_ = false;
  std::vector<float> normals_;
  std::vector<int32_t> valid_;

  bool child_init(ros::NodeHandle& nh, bool& has_service) override {
    has_service = true;
    if (lb_gicp_create(0, &impl_) != LB_OK) { NODELET_FATAL("%s", lb_last_error_string()); return false; }
    srv_ = boost::make_shared<dynamic_reconfigure::Server<NormalComputationConfig>>(nh);
    srv_->setCallback(boost::bind(&B200NormalComputation::config_callback, this, _1, _2));
    return true;
  }

  void filter(const PointCloud2::ConstPtr& input, const IndicesPtr&, PointCloud2& output) override {
    pcl::PointCloud<pcl::PointXYZI> in;
    pcl::fromROSMsg(*input, in);                                   // normal_computation.cc:29-31
    const size_t n = in.points.size();
    normals_.resize(4 * n);
    pcl::PointCloud<pcl::PointXYZINormal> out;
    out.header = in.header;
    if (n > 0) {
      int s = lb_gicp_set_source(impl_, in.points.data(), n, sizeof(pcl::PointXYZI), 0, LB_NO_NORMALS, LB_MEM_HOST);
      size_t n_keep = n;
      valid_.resize(n);
      if (s == LB_OK)
        s = radius_mode_ ? lb_gicp_compute_normals_radius(impl_, 0, radius_, nullptr, normals_.data(), valid_.data(), &n_keep, LB_MEM_HOST)
                         : lb_gicp_compute_normals(impl_, 0, k_, nullptr, normals_.data(), LB_MEM_HOST);
      if (s != LB_OK) { NODELET_WARN("[B200NormalComputation] %s", lb_last_error_string()); return; }
      out.points.resize(n_keep);
      for (size_t o = 0; o < n_keep; o++) {                        // normal_computation.cc:39-49 (+ :53-57 in radius mode)
        const size_t i = radius_mode_ ? (size_t)valid_[o] : o;
        pcl::PointXYZINormal& p = out.points[o];
        p.x = in.points[i].x; p.y = in.points[i].y; p.z = in.points[i].z; p.intensity = in.points[i].intensity;
        p.normal_x = normals_[4 * i]; p.normal_y = normals_[4 * i + 1]; p.normal_z = normals_[4 * i + 2];
      }
    }
    pcl::toROSMsg(out, output);
  }

  void config_callback(NormalComputationConfig& config, uint32_t) {
    if (config.normal_search_method == "search_knn") {
      k_ = config.normal_search_knn; radius_mode_ = false;
    } else if (config.normal_search_method == "search_radius") {
      radius_ = config.normal_search_radius; radius_mode_ = true;
    } else {
      NODELET_ERROR("Wrong normal search method in point_cloud_filter/B200NormalComputation ('%s')",
                    config.normal_search_method.c_str());
    }
  }

 public:
  ~B200NormalComputation() override { lb_gicp_destroy(impl_); }
};

}  // namespace point_cloud_filter

PLUGINLIB_EXPORT_CLASS(point_cloud_filter::B200NormalComputation, nodelet::Nodelet)
