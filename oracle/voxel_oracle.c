/*
 * voxel_oracle.c -- CPU restatement of
 *   pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter   (PCL 1.10 filters/src/voxel_grid.cpp)
 * as driven by point_cloud_filter/src/custom_voxel_grid.cc:76-87.
 * TEST INFRASTRUCTURE ONLY.
 *
 * PCL is not in the reference tree; the in-tree near copy of the index
 * arithmetic is multithreaded_ndt/voxel_grid_covariance_omp_impl.hpp:67-164.
 * No reference test touches VoxelGrid (point_cloud_filter/test is an empty
 * fixture) -> PARITY UNPINNED; documented choices:
 *   - PCL sorts (idx, point#) pairs with unstable std::sort on idx only, so the
 *     float32 summation order inside a voxel is unspecified in PCL itself.
 *     The oracle DEFINES it as ascending input index (stable sort).
 *   - centroid = (first + p2 + p3 ...) / float(count) in float32, true division
 *     (Eigen 3.3 operator/=).
 *   - only FLOAT32 fields are averaged; bytes of other fields are taken from
 *     the voxel's first point (PCL memcpy's them into float slots, which is
 *     meaningless for non-float fields).
 *   - leaf-too-small (int32 overflow of dx*dy*dz): returns -2 (PCL warns).
 * Compile with -ffp-contract=off.
 */
#include "lb_oracle.h"

#include <float.h>
#include <math.h>
#include <stdlib.h>
#include <string.h>

typedef struct { int32_t idx; uint32_t pt; } vox_pair;

static int cmp_pair(const void* a, const void* b) {
  const vox_pair* x = (const vox_pair*)a; const vox_pair* y = (const vox_pair*)b;
  if (x->idx != y->idx) return (x->idx < y->idx) ? -1 : 1;
  return (x->pt < y->pt) ? -1 : (x->pt > y->pt);
}

static inline float ldf(const uint8_t* p) { float f; memcpy(&f, p, 4); return f; }

/* predicate "point is dropped by the filter-field limits", binning pass (double compare) */
static inline int dropped_by_limits_d(float v, const og_voxel_params* P) {
  if (P->filter_limit_negative) return (v < P->filter_limit_max && v > P->filter_limit_min);
  return (v > P->filter_limit_max || v < P->filter_limit_min);
}
/* same for getMinMax3D, which casts the limits to float first */
static inline int dropped_by_limits_f(float v, const og_voxel_params* P) {
  float mn = (float)P->filter_limit_min, mx = (float)P->filter_limit_max;
  if (P->filter_limit_negative) return (v < mx && v > mn);
  return (v > mx || v < mn);
}

/* BodyFilter (body_filter.cc:28-56): pcl::CropBox<PointXYZI>, setNegative(true), setRotation((0, 0, rz)).
 * PCL 1.10 filters/impl/crop_box.hpp: local = inverse(getTransformation(0,0,0, 0,0,rz)) * p in float32; the point
 * is INSIDE unless a local coordinate lies below min or above max; negative -> inside points are removed.
 * Parity unpinned (PCL absent): the last bit of the 3x3 inverse only matters for points within 1e-7 of a face. */
static inline int dropped_by_body(float x, float y, float z, const og_voxel_params* P) {
  if (!P->body_enabled) return 0;
  float A = cosf(P->body_rotation), B = sinf(P->body_rotation);
  float det = A * A + B * B;
  float ia = A / det, ib = B / det;
  float lx = ia * x + ib * y, ly = ia * y - ib * x, lz = z;
  int outside = (lx < P->body_min[0] || ly < P->body_min[1] || lz < P->body_min[2]) ||
                (lx > P->body_max[0] || ly > P->body_max[1] || lz > P->body_max[2]);
  return !outside;
}

int og_voxel_filter(const uint8_t* data, size_t n, uint32_t point_step,
                    uint32_t x_off, uint32_t y_off, uint32_t z_off,
                    const uint32_t* ffo, int n_ff, const og_voxel_params* P,
                    uint8_t* out, size_t* n_out,
                    int32_t* out_voxel_idx, int32_t* out_first_pt, int32_t* out_count,
                    int32_t min_b_out[3], int32_t div_b_out[3]) {
  *n_out = 0;
  if (!data || !out || point_step < 12 || n > 0x7fffffffu) return -1;
  /* setLeafSize: inverse_leaf_size_ = 1 / leaf (float) */
  float inv[3] = {1.0f / P->leaf[0], 1.0f / P->leaf[1], 1.0f / P->leaf[2]};
  const int has_ff = P->filter_field_offset >= 0;

  /* getMinMax3D over finite points passing the limits */
  float min_p[3] = {FLT_MAX, FLT_MAX, FLT_MAX}, max_p[3] = {-FLT_MAX, -FLT_MAX, -FLT_MAX};
  size_t n_valid = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t* p = data + i * point_step;
    if (has_ff) {
      float dv = ldf(p + P->filter_field_offset);
      if (dropped_by_limits_f(dv, P)) continue;
    }
    float x = ldf(p + x_off), y = ldf(p + y_off), z = ldf(p + z_off);
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (dropped_by_body(x, y, z, P)) continue;
    if (x < min_p[0]) min_p[0] = x; if (y < min_p[1]) min_p[1] = y; if (z < min_p[2]) min_p[2] = z;
    if (x > max_p[0]) max_p[0] = x; if (y > max_p[1]) max_p[1] = y; if (z > max_p[2]) max_p[2] = z;
    n_valid++;
  }
  if (n_valid == 0) return 0; /* nothing survives the finite / limits checks: empty output */

  int64_t dx = (int64_t)((max_p[0] - min_p[0]) * inv[0]) + 1;
  int64_t dy = (int64_t)((max_p[1] - min_p[1]) * inv[1]) + 1;
  int64_t dz = (int64_t)((max_p[2] - min_p[2]) * inv[2]) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) return -2;

  int32_t min_b[3], max_b[3], div_b[3], mul[3];
  for (int d = 0; d < 3; d++) {
    min_b[d] = (int32_t)floorf(min_p[d] * inv[d]);
    max_b[d] = (int32_t)floorf(max_p[d] * inv[d]);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  mul[0] = 1; mul[1] = div_b[0]; mul[2] = div_b[0] * div_b[1];
  if (min_b_out) { min_b_out[0] = min_b[0]; min_b_out[1] = min_b[1]; min_b_out[2] = min_b[2]; }
  if (div_b_out) { div_b_out[0] = div_b[0]; div_b_out[1] = div_b[1]; div_b_out[2] = div_b[2]; }

  vox_pair* iv = (vox_pair*)malloc(sizeof(vox_pair) * (n ? n : 1));
  size_t cnt = 0;
  for (size_t i = 0; i < n; i++) {
    const uint8_t* p = data + i * point_step;
    if (has_ff) {
      float dv = ldf(p + P->filter_field_offset);
      if (dropped_by_limits_d(dv, P)) continue;
    }
    float x = ldf(p + x_off), y = ldf(p + y_off), z = ldf(p + z_off);
    if (!isfinite(x) || !isfinite(y) || !isfinite(z)) continue;
    if (dropped_by_body(x, y, z, P)) continue;
    int ijk0 = (int)(floorf(x * inv[0]) - (float)min_b[0]);
    int ijk1 = (int)(floorf(y * inv[1]) - (float)min_b[1]);
    int ijk2 = (int)(floorf(z * inv[2]) - (float)min_b[2]);
    iv[cnt].idx = ijk0 * mul[0] + ijk1 * mul[1] + ijk2 * mul[2];
    iv[cnt].pt = (uint32_t)i;
    cnt++;
  }
  qsort(iv, cnt, sizeof(vox_pair), cmp_pair);

  size_t o = 0;
  for (size_t cp = 0; cp < cnt;) {
    size_t e = cp + 1;
    while (e < cnt && iv[e].idx == iv[cp].idx) e++;
    if ((int)(e - cp) >= P->min_points_per_voxel) {
      uint8_t* dst = out + o * point_step;
      const uint8_t* first = data + (size_t)iv[cp].pt * point_step;
      memcpy(dst, first, point_step);
      if (P->downsample_all_data) {
        for (int f = 0; f < n_ff; f++) {
          float c = ldf(first + ffo[f]);
          for (size_t j = cp + 1; j < e; j++) c = c + ldf(data + (size_t)iv[j].pt * point_step + ffo[f]);
          c = c / (float)(e - cp);
          memcpy(dst + ffo[f], &c, 4);
        }
      } else {
        uint32_t xyz[3] = {x_off, y_off, z_off};
        for (int f = 0; f < 3; f++) {
          float c = ldf(first + xyz[f]);
          for (size_t j = cp + 1; j < e; j++) c = c + ldf(data + (size_t)iv[j].pt * point_step + xyz[f]);
          c = c / (float)(e - cp);
          memcpy(dst + xyz[f], &c, 4);
        }
      }
      if (out_voxel_idx) out_voxel_idx[o] = iv[cp].idx;
      if (out_first_pt) out_first_pt[o] = (int32_t)iv[cp].pt;
      if (out_count) out_count[o] = (int32_t)(e - cp);
      o++;
    }
    cp = e;
  }
  free(iv);
  *n_out = o;
  return 0;
}

/* ------------------------------------------------------------------ row f1 */
/* utils.cc:106-128 normalizePCloud: centroid via pcl::compute3DCentroid (float32
 * accumulation in point order), dist accumulated in float, factor = n/dist,
 * transform = [factor*I | -factor*centroid] applied by pcl::transformPointCloud. */
void og_normalize_pcloud(const float* xyz, int n, float* out) {
  /* pcl::compute3DCentroid(cloud, Eigen::Vector4f&) (utils.cc:108): Scalar = float, dense cloud: the three sums are
   * accumulated in float32 in point order, then divided by the count */
  float acc[3] = {0.f, 0.f, 0.f};
  for (int i = 0; i < n; i++) { acc[0] = acc[0] + xyz[3 * i]; acc[1] = acc[1] + xyz[3 * i + 1]; acc[2] = acc[2] + xyz[3 * i + 2]; }
  float c[3] = {acc[0] / (float)n, acc[1] / (float)n, acc[2] / (float)n};
  float dist = 0;
  for (int i = 0; i < n; i++) {
    float dx = xyz[3 * i] - c[0], dy = xyz[3 * i + 1] - c[1], dz = xyz[3 * i + 2] - c[2];
    dist = dist + sqrtf((dx * dx + dy * dy) + dz * dz);
  }
  float factor = (float)n / dist;
  float t[3] = {-factor * c[0], -factor * c[1], -factor * c[2]};
  for (int i = 0; i < n; i++)
    for (int d = 0; d < 3; d++) {
      /* se3 form: p0 + (p1 + (p2 + c3)) with a diagonal matrix */
      float v = factor * xyz[3 * i + d] + t[d];
      out[3 * i + d] = v;
    }
}

/* PointCloudLocalization.cc:723-750 */
void og_compute_ap(const float* q, int n, const float* nrm, const int64_t* corr, double* Ap) {
  for (int i = 0; i < 36; i++) Ap[i] = 0.0;
  for (int i = 0; i < n; i++) {
    double a[3] = {q[3 * i], q[3 * i + 1], q[3 * i + 2]};
    const float* nn = &nrm[3 * (size_t)corr[i]];
    double nv[3] = {nn[0], nn[1], nn[2]};
    if (isnan(a[0]) || isnan(a[1]) || isnan(a[2]) || isnan(nv[0]) || isnan(nv[1]) || isnan(nv[2])) continue;
    double H[6];
    H[0] = a[1] * nv[2] - a[2] * nv[1];
    H[1] = a[2] * nv[0] - a[0] * nv[2];
    H[2] = a[0] * nv[1] - a[1] * nv[0];
    H[3] = nv[0]; H[4] = nv[1]; H[5] = nv[2];
    for (int r = 0; r < 6; r++)
      for (int c = 0; c < 6; c++) Ap[r * 6 + c] += H[r] * H[c];
  }
}
