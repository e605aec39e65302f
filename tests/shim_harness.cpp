// shim_harness.cpp -- TEST-ONLY driver of shim/b200_gicp_pcl.hpp compiled against the PCL mock (tests/pcl_stub).
// What LOCUS's callers do with `icp_` (PointCloudOdometry.cc:147-155,265-269; PointCloudLocalization.cc:306-336), on
// the reference's hollow-cube fixture (test_point_cloud_odometry.cpp:60-97,280-305).  Prints key=value lines that
// tests/test_gicp_gpu.py::test_pcl_shim_runs checks; exit code 3 = no CUDA device (the shim has no CPU path).
#include <cmath>
#include <cstdio>
#include <limits>

#include "b200_gicp_pcl.hpp"
#include "b200_ndt_pcl.hpp"

static PointCloudF::Ptr hollow_cube(float ox, float oy) {
  PointCloudF::Ptr c(new PointCloudF);
  for (int ix = 0; ix < 10; ix++)
    for (int iy = 0; iy < 10; iy++)
      for (int iz = 0; iz < 10; iz++)
        if (ix == 0 || iy == 0 || ix == 9 || iy == 9) {
          PointF p{};
          p.x = ix * 0.1f + ox; p.y = iy * 0.1f + oy; p.z = iz * 0.1f;
          c->push_back(p);
        }
  return c;
}

int main() {
  boost::shared_ptr<pcl::B200GeneralizedIterativeClosestPoint> gicp;
  try {
    gicp = boost::make_shared<pcl::B200GeneralizedIterativeClosestPoint>(0);
  } catch (const std::exception& e) {
    std::printf("no_device=%s\n", e.what());
    return 3;
  }
  pcl::Registration<PointF, PointF>::Ptr icp = gicp;     // what the callers hold (PointCloudOdometry.h:154)
  gicp->setTransformationEpsilon(1e-3);                    // SetupICP(), PointCloudOdometry.cc:147-155
  gicp->setMaxCorrespondenceDistance(1.0);
  gicp->setMaximumIterations(20);
  gicp->setRANSACIterations(0);
  gicp->setNumThreads(4);
  gicp->enableTimingOutput(false);
  gicp->RecomputeTargetCovariance(true);
  gicp->RecomputeSourceCovariance(true);
  gicp->setEuclideanFitnessEpsilon(0.005);

  PointCloudF::Ptr box = hollow_cube(0.f, 0.f), moved = hollow_cube(0.05f, 0.05f);
  icp->setInputSource(moved);
  icp->setInputTarget(box);
  PointCloudF out;
  icp->align(out);
  Eigen::Matrix4f T = icp->getFinalTransformation();
  std::printf("converged=%d\n", icp->hasConverged() ? 1 : 0);
  std::printf("T=");
  for (int i = 0; i < 16; i++) std::printf("%.9g%s", T.m[i], i == 15 ? "\n" : ",");
  std::printf("iterations=%d\n", gicp->getLastResult().iterations);
  // output cloud = final * input
  double err = 0;
  for (size_t i = 0; i < out.size(); i++) {
    const PointF& p = (*moved)[i];
    float x = T(0, 0) * p.x + (T(0, 1) * p.y + (T(0, 2) * p.z + T(0, 3)));
    err = std::fmax(err, std::fabs(x - out[i].x));
  }
  std::printf("output_err=%g\n", err);
  // no FLANN build by initCompute(); the lazy tree is built by the first search through getSearchMethodTarget()
  auto lazy = boost::static_pointer_cast<pcl::B200LazyKdTree<PointF>>(icp->getSearchMethodTarget());
  std::printf("tree_built_by_align=%d\n", lazy->built() ? 1 : 0);
  std::vector<int> idx; std::vector<float> d2;
  icp->getSearchMethodTarget()->nearestKSearch(out[7], 1, idx, d2);      // PointCloudLocalization.cc:331-335
  std::printf("tree_built_by_search=%d\n", lazy->built() ? 1 : 0);
  std::vector<int> bi; std::vector<float> bd;
  gicp->nearestTarget(out, bi, bd);
  std::printf("batched_nn_agrees=%d\n", (bi[7] == idx[0] && bd[7] == d2[0]) ? 1 : 0);
  std::printf("fitness=%.9g\n", icp->getFitnessScore());
  // a refused cloud keeps the previous input on both sides
  PointCloudF::Ptr bad = hollow_cube(0.f, 0.f);
  (*bad)[3].y = std::numeric_limits<float>::quiet_NaN();
  icp->setInputTarget(bad);
  std::printf("target_kept=%d\n", icp->getInputTarget() == box ? 1 : 0);
  icp->align(out);
  Eigen::Matrix4f T2 = icp->getFinalTransformation();
  bool same = true;
  for (int i = 0; i < 16; i++) same = same && (T.m[i] == T2.m[i]);
  std::printf("same_pose_after_refused_target=%d\n", same ? 1 : 0);
  // from-normals mode (the reference default): zero normals on the source -> identity covariances, like the reference test
  gicp->RecomputeTargetCovariance(false);
  gicp->RecomputeSourceCovariance(false);
  icp->align(out);
  std::printf("from_normals_converged=%d\n", icp->hasConverged() ? 1 : 0);

  // `registration_method: ndt` (SetupICP(), PointCloudOdometry.cc:182-195) through shim/b200_ndt_pcl.hpp
  boost::shared_ptr<pcl::B200NormalDistributionsTransform> ndt = boost::make_shared<pcl::B200NormalDistributionsTransform>(0);
  pcl::Registration<PointF, PointF>::Ptr icp2 = ndt;
  ndt->setTransformationEpsilon(1e-3);
  ndt->setMaxCorrespondenceDistance(1.0);
  ndt->setMaximumIterations(20);
  ndt->setRANSACIterations(0);
  ndt->setNumThreads(4);
  ndt->enableTimingOutput(false);
  ndt->setResolution(0.5f);                                   // the cube is 0.9 m wide
  icp2->setInputSource(moved);
  icp2->setInputTarget(box);
  PointCloudF out2;
  icp2->align(out2);
  Eigen::Matrix4f Tn = icp2->getFinalTransformation();
  std::printf("ndt_converged=%d\n", icp2->hasConverged() ? 1 : 0);
  std::printf("ndt_T=");
  for (int i = 0; i < 16; i++) std::printf("%.9g%s", Tn.m[i], i == 15 ? "\n" : ",");
  std::printf("ndt_iterations=%d\n", ndt->getFinalNumIteration());
  double err2 = 0;
  for (size_t i = 0; i < out2.size(); i++) {
    const PointF& p = (*moved)[i];
    float x = Tn(0, 0) * p.x + (Tn(0, 1) * p.y + (Tn(0, 2) * p.z + Tn(0, 3)));
    err2 = std::fmax(err2, std::fabs(x - out2[i].x));
  }
  std::printf("ndt_output_err=%g\n", err2);
  auto lazy2 = boost::static_pointer_cast<pcl::B200LazyKdTree<PointF>>(icp2->getSearchMethodTarget());
  std::printf("ndt_tree_built_by_align=%d\n", lazy2->built() ? 1 : 0);
  icp2->setInputSource(bad);                                  // a source with a NaN point is refused, the previous one kept
  std::printf("ndt_source_kept=%d\n", icp2->getInputSource() == moved ? 1 : 0);
  return 0;
}
