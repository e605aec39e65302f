// example_odometry.cpp -- compile/link check of the C++ host mirror, and a tiny end-to-end use
// that reads like PointCloudOdometry::UpdateICP (PointCloudOdometry.cc:249-269).
// Built by __graft_entry__.build(); needs a GPU to run.
#include <cmath>
#include <cstdio>

#include "b200_gicp.hpp"

using namespace locus_b200;

int main() {
  int ndev = 0;
  lb_device_count(&ndev);
  if (ndev <= 0) { std::printf("no CUDA device: nothing to run (the library has no CPU path)\n"); return 0; }
  // hollow cube of test_point_cloud_odometry.cpp:60-97, shifted by (+0.05,+0.05,0)
  std::vector<PointF> reference, query;
  for (int ix = 0; ix < 10; ix++)
    for (int iy = 0; iy < 10; iy++)
      for (int iz = 0; iz < 10; iz++)
        if (ix == 0 || iy == 0 || ix == 9 || iy == 9) {
          PointF p{};
          p.x = ix * 0.1f; p.y = iy * 0.1f; p.z = iz * 0.1f; p.data3 = 1.f;
          reference.push_back(p);
          p.x += 0.05f; p.y += 0.05f;
          query.push_back(p);
        }
  B200Gicp icp;
  icp.setTransformationEpsilon(1e-3);
  icp.setMaxCorrespondenceDistance(1.0);
  icp.setMaximumIterations(20);
  icp.setRANSACIterations(0);
  icp.RecomputeSourceCovariance(true);
  icp.RecomputeTargetCovariance(true);
  icp.setInputSource(query.data(), query.size(), false);
  icp.setInputTarget(reference.data(), reference.size(), false);
  std::vector<PointF> aligned;
  if (!icp.align(&aligned)) { std::printf("align failed: %s\n", lb_last_error_string()); return 1; }
  Matrix4f T = icp.getFinalTransformation();
  std::printf("converged=%d t=(%.4f %.4f %.4f) fitness=%.3g\n", (int)icp.hasConverged(), T[3], T[7], T[11], icp.getFitnessScore());
  return (icp.hasConverged() && std::fabs(T[3] + 0.05f) < 1e-2f && std::fabs(T[7] + 0.05f) < 1e-2f) ? 0 : 2;
}
