// b200_voxel_grid_nodelet.cc -- nodelet point_cloud_filter/B200VoxelGrid.
// Compiled ONLY inside a LOCUS catkin workspace (needs ROS + pcl_ros); see INTEGRATION.md.
// Same three virtuals, dynamic-reconfigure fields and change_leaf_size topic as
// point_cloud_filter::CustomVoxelGrid (point_cloud_filter/src/custom_voxel_grid.cc:46-151);
// only impl_ (pcl::VoxelGrid<pcl::PCLPointCloud2>) is replaced by the lb_voxel handle.
#include <pluginlib/class_list_macros.h>
#include <pcl_ros/filters/filter.h>
#include <point_cloud_filter/CustomVoxelGridConfig.h>
#include <std_msgs/Float64.h>

#include "locus_b200.h"

namespace point_cloud_filter {

class B200VoxelGrid : public pcl_ros::Filter {
 protected:
  boost::shared_ptr<dynamic_reconfigure::Server<CustomVoxelGridConfig>> srv_;
  lb_voxel* impl_ = nullptr;
  float leaf_ = 0.25f;
  std::string field_ = "z";
  double lim_min_ = -100, lim_max_ = 100;
  bool negative_ = false;
  ros::Subscriber change_leaf_size_sub;

  bool child_init(ros::NodeHandle& nh, bool& has_service) override {
    has_service = true;
    if (lb_voxel_create(0, &impl_) != LB_OK) { NODELET_FATAL("%s", lb_last_error_string()); return false; }
    srv_ = boost::make_shared<dynamic_reconfigure::Server<CustomVoxelGridConfig>>(nh);
    srv_->setCallback(boost::bind(&B200VoxelGrid::config_callback, this, _1, _2));
    change_leaf_size_sub = nh.subscribe("change_leaf_size", 10, &B200VoxelGrid::ChangeLeafSizeRostopic, this);
    return true;
  }

  void ChangeLeafSizeRostopic(const std_msgs::Float64::ConstPtr leaf_size_in) {
    boost::mutex::scoped_lock lock(mutex_);
    if (leaf_ != leaf_size_in->data) { leaf_ = leaf_size_in->data; lb_voxel_set_leaf_size(impl_, leaf_, leaf_, leaf_); }
  }

  void filter(const PointCloud2::ConstPtr& input, const IndicesPtr& indices, PointCloud2& output) override {
    boost::mutex::scoped_lock lock(mutex_);
    std::vector<lb_field> fields(input->fields.size());
    for (size_t i = 0; i < fields.size(); i++) {
      std::strncpy(fields[i].name, input->fields[i].name.c_str(), sizeof(fields[i].name) - 1);
      fields[i].offset = input->fields[i].offset; fields[i].datatype = input->fields[i].datatype; fields[i].count = input->fields[i].count;
    }
    const size_t n = size_t(input->width) * input->height;
    output.header = input->header; output.fields = input->fields; output.point_step = input->point_step;
    output.is_bigendian = input->is_bigendian; output.height = 1; output.is_dense = true;
    output.data.resize(n * input->point_step);
    size_t n_out = 0;
    int s = lb_voxel_filter(impl_, input->data.data(), n, input->point_step, fields.data(), (int)fields.size(),
                            indices ? indices->data() : nullptr, indices ? indices->size() : 0, output.data.data(), n,
                            &n_out, nullptr, LB_MEM_HOST, LB_MEM_HOST);
    if (s != LB_OK) { NODELET_WARN("[B200VoxelGrid] %s", lb_last_error_string()); n_out = 0; }
    output.width = n_out; output.row_step = output.point_step * output.width;
    output.data.resize(n_out * output.point_step);
  }

  void config_callback(CustomVoxelGridConfig& config, uint32_t) {
    boost::mutex::scoped_lock lock(mutex_);
    if (leaf_ != config.leaf_size) { leaf_ = config.leaf_size; lb_voxel_set_leaf_size(impl_, leaf_, leaf_, leaf_); }
    lim_min_ = config.filter_limit_min; lim_max_ = config.filter_limit_max;
    negative_ = config.filter_limit_negative; field_ = config.filter_field_name;
    lb_voxel_set_filter_limits(impl_, field_.c_str(), lim_min_, lim_max_, negative_);
    tf_input_frame_ = config.input_frame; tf_output_frame_ = config.output_frame;
  }

 public:
  ~B200VoxelGrid() override { lb_voxel_destroy(impl_); }
};

}  // namespace point_cloud_filter
PLUGINLIB_EXPORT_CLASS(point_cloud_filter::B200VoxelGrid, nodelet::Nodelet);
