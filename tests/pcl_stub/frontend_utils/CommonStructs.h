// stand-in for the un-vendored frontend_utils/CommonStructs.h (gicp.h:46): the two typedefs the hot path uses
#pragma once
#include <pcl/point_cloud.h>
#include <pcl/point_types.h>
typedef pcl::PointXYZINormal PointF;
typedef pcl::PointCloud<PointF> PointCloudF;
