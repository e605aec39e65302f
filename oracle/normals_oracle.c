/* normals_oracle.c -- TEST INFRASTRUCTURE ONLY (see lb_oracle.h): CPU restatement of what
 * point_cloud_filter::NormalComputation::filter computes in its k-NN mode
 * (point_cloud_filter/src/normal_computation.cc:26-59, norm_est_ = pcl::NormalEstimationOMP<PointXYZI, Normal>,
 * normal_computation.h:38; k = normal_search_knn, cfg/NormalComputation.cfg:15).
 *
 * The arithmetic lives in PCL, which is NOT vendored in the reference (find_package(PCL 1.7 REQUIRED),
 * point_cloud_filter/CMakeLists.txt:34; de-facto ROS Noetic version PCL 1.10.0).  This file restates PCL 1.10's
 * published algorithm, single precision like PCL:
 *   NormalEstimationOMP::computeFeature    features/impl/normal_3d_omp.hpp : per point, nearestKSearch(k) (the
 *                                          point itself included), computePointNormal, flipNormalTowardsViewpoint
 *                                          with the viewpoint (0,0,0) that fromROSMsg leaves in sensor_origin_
 *   computeMeanAndCovarianceMatrix         common/impl/centroid.hpp (dense branch, 9 float accumulators)
 *   solvePlaneParameters                   features/impl/feature.hpp
 *   eigen33 / computeRoots / computeRoots2 common/impl/eigen.hpp
 * PARITY UNPINNED: no golden normals exist anywhere in the reference and PCL cannot be built here.  The neighbour
 * order among equidistant points (FLANN) is taken as ascending index, like the rest of this oracle.
 */
#include <math.h>
#include <stdlib.h>

#include "lb_oracle.h"

static void roots2(float b, float c, float* r) {
  r[0] = 0.f;
  float d = b * b - 4.0f * c;
  if (d < 0.0f) d = 0.0f;
  float sd = sqrtf(d);
  r[2] = 0.5f * (b + sd);
  r[1] = 0.5f * (b - sd);
}

static void roots3(float m[3][3], float* r) {
  float c0 = m[0][0] * m[1][1] * m[2][2] + 2.0f * m[0][1] * m[0][2] * m[1][2] - m[0][0] * m[1][2] * m[1][2] -
             m[1][1] * m[0][2] * m[0][2] - m[2][2] * m[0][1] * m[0][1];
  float c1 = m[0][0] * m[1][1] - m[0][1] * m[0][1] + m[0][0] * m[2][2] - m[0][2] * m[0][2] + m[1][1] * m[2][2] -
             m[1][2] * m[1][2];
  float c2 = m[0][0] + m[1][1] + m[2][2];
  if (fabsf(c0) < 1.1920929e-07f) { roots2(c2, c1, r); return; }
  const float inv3 = 1.0f / 3.0f, sqrt3 = sqrtf(3.0f);
  float c2_3 = c2 * inv3;
  float a_3 = (c1 - c2 * c2_3) * inv3;
  if (a_3 > 0.0f) a_3 = 0.0f;
  float half_b = 0.5f * (c0 + c2_3 * (2.0f * c2_3 * c2_3 - c1));
  float q = half_b * half_b + a_3 * a_3 * a_3;
  if (q > 0.0f) q = 0.0f;
  float rho = sqrtf(-a_3);
  float theta = atan2f(sqrtf(-q), half_b) * inv3;
  float ct = cosf(theta), st = sinf(theta);
  r[0] = c2_3 + 2.0f * rho * ct;
  r[1] = c2_3 - rho * (ct + sqrt3 * st);
  r[2] = c2_3 - rho * (ct - sqrt3 * st);
  float t;
  if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  if (r[1] >= r[2]) {
    t = r[1]; r[1] = r[2]; r[2] = t;
    if (r[0] >= r[1]) { t = r[0]; r[0] = r[1]; r[1] = t; }
  }
  if (r[0] <= 0.0f) roots2(c2, c1, r);
}

/* PCL computePointNormal + flipNormalTowardsViewpoint for one point from its neighbour list (search order) */
static void normal_from_neighbours(const float* pts, int stride_f, const float* p, const int* idx, int found, const float vp[3],
                                   float* o) {
  if (found < 3) {          /* computePointNormal: indices.size() < 3 -> normal and curvature are NaN */
    o[0] = o[1] = o[2] = o[3] = NAN;
    return;
  }
  float a[9] = {0, 0, 0, 0, 0, 0, 0, 0, 0};
  for (int j = 0; j < found; j++) {
    const float* q = pts + (size_t)idx[j] * stride_f;
    a[0] += q[0] * q[0]; a[1] += q[0] * q[1]; a[2] += q[0] * q[2];
    a[3] += q[1] * q[1]; a[4] += q[1] * q[2]; a[5] += q[2] * q[2];
    a[6] += q[0]; a[7] += q[1]; a[8] += q[2];
  }
  float cnt = (float)found;
  for (int j = 0; j < 9; j++) a[j] /= cnt;
  float C[3][3];
  C[0][0] = a[0] - a[6] * a[6]; C[0][1] = a[1] - a[6] * a[7]; C[0][2] = a[2] - a[6] * a[8];
  C[1][1] = a[3] - a[7] * a[7]; C[1][2] = a[4] - a[7] * a[8]; C[2][2] = a[5] - a[8] * a[8];
  C[1][0] = C[0][1]; C[2][0] = C[0][2]; C[2][1] = C[1][2];
  float scale = 0.f;
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) if (fabsf(C[r][c]) > scale) scale = fabsf(C[r][c]);
  if (scale <= 1.17549435e-38f) scale = 1.0f;
  float S[3][3];
  for (int r = 0; r < 3; r++)
    for (int c = 0; c < 3; c++) S[r][c] = C[r][c] / scale;
  float ev[3];
  roots3(S, ev);
  float lambda = ev[0] * scale;
  for (int r = 0; r < 3; r++) S[r][r] -= ev[0];
  float v[3][3];
  const int ra[3] = {0, 0, 1}, rb[3] = {1, 2, 2};
  float len[3];
  for (int t = 0; t < 3; t++) {
    const float* x = S[ra[t]]; const float* y = S[rb[t]];
    v[t][0] = x[1] * y[2] - x[2] * y[1];
    v[t][1] = x[2] * y[0] - x[0] * y[2];
    v[t][2] = x[0] * y[1] - x[1] * y[0];
    len[t] = (v[t][0] * v[t][0] + v[t][1] * v[t][1]) + v[t][2] * v[t][2];
  }
  int best = (len[0] >= len[1] && len[0] >= len[2]) ? 0 : ((len[1] >= len[0] && len[1] >= len[2]) ? 1 : 2);
  float s = sqrtf(len[best]);
  float nrm[3] = {v[best][0] / s, v[best][1] / s, v[best][2] / s};
  float tr = (C[0][0] + C[1][1]) + C[2][2];
  float curv = (tr != 0.f) ? fabsf(lambda / tr) : 0.f;
  float dx = vp[0] - p[0], dy = vp[1] - p[1], dz = vp[2] - p[2];
  if ((dx * nrm[0] + dy * nrm[1]) + dz * nrm[2] < 0.f) { nrm[0] = -nrm[0]; nrm[1] = -nrm[1]; nrm[2] = -nrm[2]; }
  o[0] = nrm[0]; o[1] = nrm[1]; o[2] = nrm[2]; o[3] = curv;
}

/* pts: n points, stride_f floats apart (x, y, z first).  out4: n x (nx, ny, nz, curvature).  Returns 0, or -1 when
 * k < 3 or k > n (PCL then writes NaN normals / FLANN returns fewer neighbours; callers never do that). */
int og_normals_knn(const float* pts, int n, int stride_f, int k, const float vp[3], float* out4, int num_threads) {
  if (k < 3 || k > n) return -1;
  og_kdtree* tree = og_kdtree_build(pts, n, stride_f);
  if (!tree) return -1;
  if (num_threads < 1) num_threads = 1;
#pragma omp parallel num_threads(num_threads)
  {
    int* idx = (int*)malloc(sizeof(int) * (size_t)k);
    float* d2 = (float*)malloc(sizeof(float) * (size_t)k);
#pragma omp for schedule(dynamic, 256)
    for (int i = 0; i < n; i++) {
      const float* p = pts + (size_t)i * stride_f;
      int found = og_kdtree_knn(tree, p, k, idx, d2);
      normal_from_neighbours(pts, stride_f, p, idx, found, vp, out4 + 4 * (size_t)i);
    }
    free(idx); free(d2);
  }
  og_kdtree_free(tree);
  return 0;
}

/* The nodelet's radius mode (normal_computation.cc:73-77 setRadiusSearch; :53-57 removeNaNNormalsFromPointCloud): the
 * neighbours of a point are ALL points closer than `radius` (d2 < float(radius * radius), FLANN's strict compare), taken
 * in ascending distance order; fewer than 3 -> NaN normal, which the nodelet then drops from its output cloud.
 * out4: n x 4 (NaN rows where fewer than 3 neighbours); valid_idx (nullable, capacity n): indices of the points the
 * nodelet keeps, ascending.  Returns their number. */
int og_normals_radius(const float* pts, int n, int stride_f, double radius, const float vp[3], float* out4, int* valid_idx,
                      int num_threads) {
  og_kdtree* tree = og_kdtree_build(pts, n, stride_f);
  if (!tree) return -1;
  if (num_threads < 1) num_threads = 1;
  const float r2 = (float)(radius * radius);
#pragma omp parallel num_threads(num_threads)
  {
    int* idx = (int*)malloc(sizeof(int) * (size_t)(n > 0 ? n : 1));
    float* d2 = (float*)malloc(sizeof(float) * (size_t)(n > 0 ? n : 1));
#pragma omp for schedule(dynamic, 64)
    for (int i = 0; i < n; i++) {
      const float* p = pts + (size_t)i * stride_f;
      int found = og_kdtree_radius(tree, p, r2, idx, d2, n);
      normal_from_neighbours(pts, stride_f, p, idx, found, vp, out4 + 4 * (size_t)i);
    }
    free(idx); free(d2);
  }
  og_kdtree_free(tree);
  int m = 0;
  for (int i = 0; i < n; i++) {
    const float* o = out4 + 4 * (size_t)i;
    if (isfinite(o[0]) && isfinite(o[1]) && isfinite(o[2])) { if (valid_idx) valid_idx[m] = i; m++; }
  }
  return m;
}
