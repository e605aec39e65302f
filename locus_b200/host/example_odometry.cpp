// example_odometry.cpp -- compile/link check of the C++ host mirror, and a tiny end-to-end use
// that reads like PointCloudOdometry::UpdateICP (PointCloudOdometry.cc:249-269).
// Built by __graft_entry__.build(); needs a GPU to run.
#include <cmath>
#include <cstdio>
#include <vector>

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
  if (!(icp.hasConverged() && std::fabs(T[3] + 0.05f) < 1e-2f && std::fabs(T[7] + 0.05f) < 1e-2f)) return 2;

  // NormalComputation (row f2): the wall x = 0 of the cube has normals (+-1, 0, 0), flipped towards the origin
  B200Gicp ne;
  if (!ne.computeNormals(reference.data(), reference.size(), 12)) { std::printf("normals failed: %s\n", lb_last_error_string()); return 3; }
  int bad = 0;
  for (const PointF& p : reference)
    if (p.x > 0.85f && p.y > 0.25f && p.y < 0.65f && p.z > 0.25f && p.z < 0.65f && !(p.normal_x < -0.99f)) bad++;
  std::printf("normals: %d interior points of the wall x = 0.9 not facing the origin\n", bad);
  if (bad) return 4;

  // `registration_method: ndt` (row f4): the same pair through the NDT mirror (0.5 m voxels on the 0.9 m cube)
  B200Ndt ndt;
  ndt.setTransformationEpsilon(1e-3);
  ndt.setMaximumIterations(20);
  ndt.setResolution(0.5f);
  if (!ndt.setInputSource(query.data(), query.size()) || !ndt.setInputTarget(reference.data(), reference.size()) || !ndt.align()) {
    std::printf("ndt failed: %s\n", lb_last_error_string());
    return 8;
  }
  Matrix4f Tn = ndt.getFinalTransformation();
  std::printf("ndt: converged=%d t=(%.4f %.4f %.4f) after %d Newton steps\n", (int)ndt.hasConverged(), Tn[3], Tn[7], Tn[11], ndt.getFinalNumIteration());
  if (!(ndt.hasConverged() && std::fabs(Tn[3] + 0.05f) < 1e-2f && std::fabs(Tn[7] + 0.05f) < 1e-2f)) return 9;

  // the pipelined chain: the same two clouds as PointCloud2 blobs of PointF records
  B200Odometry odo(0, 2, reference.size(), sizeof(PointF));
  odo.setLeafSize(0.01f);                                     // finer than the 0.1 m lattice: every point survives
  odo.params().transformation_epsilon = 1e-3; odo.params().max_correspondence_distance = 1.0; odo.params().max_iterations = 20;
  if (!odo.applyParams()) return 5;
  std::vector<lb_field> fields(3);
  const char* names[3] = {"x", "y", "z"};
  for (int i = 0; i < 3; i++) {
    std::snprintf(fields[i].name, sizeof(fields[i].name), "%s", names[i]);
    fields[i].offset = 4u * i; fields[i].datatype = LB_FLOAT32; fields[i].count = 1;
  }
  odo.submit(reinterpret_cast<const uint8_t*>(reference.data()), reference.size(), sizeof(PointF), fields);
  odo.submit(reinterpret_cast<const uint8_t*>(query.data()), query.size(), sizeof(PointF), fields);
  lb_odometry_result r0, r1;
  if (!odo.next(&r0) || !odo.next(&r1) || r0.has_pose || !r1.has_pose || r1.status != LB_OK) {
    std::printf("pipeline failed: %s\n", r1.error);
    return 6;
  }
  std::printf("pipeline: n_filtered=%zu t=(%.4f %.4f)\n", r1.n_filtered, r1.gicp.final_transformation[3], r1.gicp.final_transformation[7]);
  return (std::fabs(r1.gicp.final_transformation[3] + 0.05f) < 1e-2f && std::fabs(r1.gicp.final_transformation[7] + 0.05f) < 1e-2f) ? 0 : 7;
}
