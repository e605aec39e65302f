/*
 * lb_oracle.h -- CPU oracle for the LOCUS GICP + VoxelGrid hot path.
 *
 * TEST INFRASTRUCTURE ONLY.  Nothing under oracle/ is part of the product:
 * only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this library, and only as the checker or as
 * the timed CPU baseline.  The product (locus_b200/csrc) never links it.
 *
 * It is a dependency-free C restatement of the reference algorithm
 * (reference = NeBula-Autonomy/LOCUS @ 84c0fed):
 *   - GICP:      multithreaded_gicp/include/multithreaded_gicp/gicp.hpp:64-634
 *                and gicp.h:111-132 (defaults), gicp.h:361-391 (helpers)
 *   - BFGS:      pcl/registration/bfgs.h (PCL 1.10, NOT in the reference tree;
 *                restated from the published algorithm = GSL vector_bfgs2 +
 *                Fletcher line search; call site gicp.hpp:250-271)
 *   - VoxelGrid: pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter (PCL 1.10,
 *                NOT in the tree; call site point_cloud_filter/src/
 *                custom_voxel_grid.cc:76-87; index arithmetic cross-checked
 *                with multithreaded_ndt/voxel_grid_covariance_omp_impl.hpp:67-164)
 *   - Ap 6x6:    point_cloud_localization/src/PointCloudLocalization.cc:694-750
 *                and src/utils.cc:106-128
 *
 * PARITY PINNING STATUS (see DESIGN.md "Oracle"):
 *   The reference cannot be compiled here (needs PCL/Eigen/FLANN/ROS, none
 *   present, no network).  The oracle is pinned against every fixture the
 *   reference's own tests hold for this path:
 *     - hollow-cube shift (test_point_cloud_odometry.cpp:280-305): converged,
 *       fitness < 0.1, inverse translation within 1e-2        -> pinned
 *     - Ap known answers 56.7753 / 56.7753 / 100
 *       (test_point_cloud_localization.cpp:337-339)            -> pinned
 *     - garage PCDs (test_same_output_different_num_threads.cpp): result is
 *       bit-identical for 1..8 threads                         -> pinned
 *       (the test's other half, equality with stock PCL GICP, needs PCL)
 *   There is NO stored golden pose / voxel output anywhere in the reference,
 *   so last-bit details that live in PCL/Eigen/FLANN (summation association,
 *   tie order of equidistant neighbours, SVD basis in degenerate
 *   neighbourhoods, sinf/cosf) are "parity unpinned": documented choices.
 */
#ifndef LB_ORACLE_H_
#define LB_ORACLE_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* ------------------------------------------------------------------ kd-tree */
typedef struct og_kdtree og_kdtree;
/* pts: n points, xyz at pts[i*stride_f + 0..2] (float32). */
og_kdtree* og_kdtree_build(const float* pts, int n, int stride_f);
void og_kdtree_free(og_kdtree* t);
/* exact k-NN; results ascending by (d2, index); returns number found (min(k,n)).
 * d2 is float32: ((dx*dx)+(dy*dy))+(dz*dz), no FMA (FLANN L2_Simple<float>). */
int og_kdtree_knn(const og_kdtree* t, const float q[3], int k, int* idx, float* d2);
/* radius search: every point with d2 < radius2 (strict, FLANN RadiusResultSet), ascending by (d2, index); returns the
 * number found and writes min(found, cap) of them */
int og_kdtree_radius(const og_kdtree* t, const float q[3], float radius2, int* idx, float* d2, int cap);
/* batch 1-NN helper for tests (OpenMP over queries) */
void og_kdtree_nn_batch(const og_kdtree* t, const float* q, int nq, int stride_f,
                        int* idx, float* d2, int num_threads);

/* --------------------------------------------------------------------- GICP */
typedef struct {
  int k_correspondences;          /* gicp.h:112  default 20   */
  double gicp_epsilon;            /* gicp.h:118  default 1e-3 */
  double rotation_epsilon;        /* gicp.h:119  default 2e-3 */
  double transformation_epsilon;  /* gicp.h:126  default 5e-4 */
  double corr_dist_threshold;     /* gicp.h:127  default 5.0  */
  int max_iterations;             /* gicp.h:125  default 200  */
  int max_inner_iterations;       /* gicp.h:121  default 20   */
  int num_threads;                /* gicp.h:117  default 1    */
  int source_cov_from_normals;    /* 1: gicp.hpp:81-82 branch; 0: k-NN branch gicp.hpp:85-154 */
  int target_cov_from_normals;
  int optimizer;                  /* 0 = BFGS (reference), 1 = Gauss-Newton (not in the reference) */
} og_gicp_params;

typedef struct {
  float final_transformation[16]; /* row-major 4x4 float, = previous * guess (gicp.hpp:583) */
  int nr_iterations;
  int converged;
  int n_correspondences;          /* of the last outer iteration */
  double delta;                   /* last convergence ratio (gicp.hpp:526-541) */
  long n_fdf_evals;               /* objective evaluations (operator()/df/fdf calls) */
  long n_inner_iterations;
  double t_covariances_s, t_iterations_s, t_total_s;   /* same spans as gicp.hpp:588-616 */
  double t_lookups_s, t_optimization_s;                 /* gicp.hpp:549-560 */
  int status;                     /* 0 ok, <0 error (k > cloud size, empty source ...) */
} og_gicp_result;

void og_gicp_default_params(og_gicp_params* p);

/* One full align(): setInputSource + setInputTarget + align(output, guess).
 * src/tgt: float32 arrays, xyz at [i*stride_f+0..2]; normals at
 * [i*stride_f+normal_off_f .. +2] (normal_off_f < 0: no normals).
 * guess: row-major 4x4 float (NULL = identity).
 * src_cov_out / tgt_cov_out (nullable): n x 9 doubles row-major, the
 * covariances used.  aligned_out (nullable): n_src x 3 float = final * input. */
int og_gicp_align(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                  const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                  const og_gicp_params* params, const float* guess,
                  og_gicp_result* result,
                  double* src_cov_out, double* tgt_cov_out, float* aligned_out);

/* Scan-to-submap with an UNCHANGED target (BASELINE configs[2]): the reference keeps the target's kd-tree and
 * covariances until setInputTarget is called again (gicp.h:196-200 clears them only there; gicp.hpp:422-426 computes
 * them "if unset").  prepare = setInputTarget + the target half of the first align; align_prepared = setInputSource +
 * align against it.  The target array must outlive the handle. */
typedef struct og_gicp_target og_gicp_target;
og_gicp_target* og_gicp_target_prepare(const float* tgt, int n_tgt, int tgt_stride_f, int tgt_normal_off_f,
                                       const og_gicp_params* params);
void og_gicp_target_set_covariances(og_gicp_target* t, const double* cov9 /* n x 9 */);
void og_gicp_target_free(og_gicp_target* t);
int og_gicp_align_prepared(const float* src, int n_src, int src_stride_f, int src_normal_off_f,
                           const og_gicp_target* target, const og_gicp_params* params, const float* guess,
                           og_gicp_result* result);

/* Summation-order probe for studies: 0 (default) = the reference's single serial loop; c > 0 = partial sums over blocks of
 * c correspondences, added in block order (same terms, other association).  Process-global; see gicp_oracle.c. */
void og_set_sum_chunk(int c);

/* k-NN covariances only (gicp.hpp:85-154). cov_out: n x 9 doubles. */
int og_gicp_covariances(const float* pts, int n, int stride_f, int k,
                        double gicp_epsilon, int num_threads, double* cov_out);

/* getFitnessScore(max_range) of pcl::Registration: mean squared 1-NN distance of
 * T*source in target, counting d2 <= max_range. */
double og_gicp_fitness(const float* src, int n_src, int src_stride_f,
                       const float* tgt, int n_tgt, int tgt_stride_f,
                       const float* T, double max_range, int num_threads);

/* Objective used by the inner solver, exposed for unit tests:
 * f, g[6] at state x[6] for explicit correspondences and Mahalanobis matrices.
 * src4/tgt4: m x 4 float (x,y,z,1); M: m x 9 doubles. */
void og_gicp_fdf(const float* src4, const float* tgt4, const double* M, int m,
                 const double x[6], double* f, double g[6]);
/* applyState (gicp.hpp:619-634): T(x) applied on identity, row-major float 4x4 */
void og_gicp_apply_state(const double x[6], float T[16]);

/* --------------------------------------------------------------------- BFGS */
typedef struct {
  double (*f)(void* ctx, const double* x);
  void (*df)(void* ctx, const double* x, double* g);
  void (*fdf)(void* ctx, const double* x, double* f, double* g);
  void* ctx;
  int n;
} og_functor;

enum { OG_BFGS_NEG_GRAD_EPS = -3, OG_BFGS_NOT_STARTED = -2, OG_BFGS_RUNNING = -1,
       OG_BFGS_SUCCESS = 0, OG_BFGS_NO_PROGRESS = 1 };

#define OG_BFGS_MAXN 8
typedef struct {
  /* parameters */
  int bracket_iters, section_iters, order;
  double rho, sigma, tau1, tau2, tau3, step_size;
  /* state */
  og_functor fn;
  int n;
  double f, gradient[OG_BFGS_MAXN];
  double delta_f, fp0;
  double x0[OG_BFGS_MAXN], dx0[OG_BFGS_MAXN], dg0[OG_BFGS_MAXN], g0[OG_BFGS_MAXN],
      dx[OG_BFGS_MAXN], p[OG_BFGS_MAXN];
  double pnorm, g0norm;
  double f_alpha, df_alpha, x_alpha[OG_BFGS_MAXN], g_alpha[OG_BFGS_MAXN];
  double f_cache_key, df_cache_key, x_cache_key, g_cache_key;
  long n_f, n_df, n_fdf;
} og_bfgs;

void og_bfgs_init_params(og_bfgs* b, og_functor fn);
int og_bfgs_minimize_init(og_bfgs* b, double* x);
int og_bfgs_minimize_one_step(og_bfgs* b, double* x);
int og_bfgs_test_gradient(const og_bfgs* b, double eps);
/* Rosenbrock-style self test driver used by tests: minimise a quadratic
 * f = 0.5 x'Ax - b'x  (A n x n row-major SPD). returns iterations. */
int og_bfgs_minimize_quadratic(const double* A, const double* bvec, int n, double* x,
                               int max_iters, double grad_tol);

/* ---------------------------------------------------------------- VoxelGrid */
typedef struct {
  float leaf[3];
  int filter_field_offset;   /* byte offset of the FLOAT32 filter field; <0: no filter field */
  double filter_limit_min, filter_limit_max;
  int filter_limit_negative;
  int min_points_per_voxel;
  int downsample_all_data;   /* 1 (PCL default): average every FLOAT32 field */
  /* the BodyFilter nodelet ahead of the voxel grid (point_cloud_filter/src/body_filter.cc:28-56 =
   * pcl::CropBox, negative): points inside the box (rotated by body_rotation about z) are removed first */
  int body_enabled;
  float body_min[3], body_max[3];
  float body_rotation;
} og_voxel_params;

/* data: n points of point_step bytes; x/y/z FLOAT32 at the given byte offsets.
 * float_field_offsets: byte offsets of the FLOAT32 fields that get averaged
 * when downsample_all_data (must include x,y,z).  Bytes not covered by an
 * averaged field are copied from the voxel's first point (lowest input index).
 * out: capacity n*point_step.  out_voxel_idx (nullable): the int32 PCL leaf
 * index of every output point (ascending).  out_first_pt (nullable): lowest
 * input index in each voxel; out_count (nullable): points per voxel.
 * returns 0 ok, -1 bad args, -2 leaf too small (int32 index overflow). */
int og_voxel_filter(const uint8_t* data, size_t n, uint32_t point_step,
                    uint32_t x_off, uint32_t y_off, uint32_t z_off,
                    const uint32_t* float_field_offsets, int n_float_fields,
                    const og_voxel_params* params,
                    uint8_t* out, size_t* n_out,
                    int32_t* out_voxel_idx, int32_t* out_first_pt, int32_t* out_count,
                    int32_t min_b_out[3], int32_t div_b_out[3]);

/* ------------------------------------------------------------- Ap (row f1) */
/* normalizePCloud (utils.cc:106-128): out = factor*(p - centroid). xyz n x 3. */
void og_normalize_pcloud(const float* xyz, int n, float* out_xyz);
/* ComputeAp_ForPoint2PlaneICP (PointCloudLocalization.cc:723-750):
 * Ap = sum H'H, H = [a x n, n]; NaN rows skipped.  Ap: 36 doubles row-major. */
void og_compute_ap(const float* query_xyz, int n, const float* ref_normals_xyz,
                   const int64_t* correspondences, double* Ap);

/* point_cloud_filter::NormalComputation (k-NN mode) = pcl::NormalEstimationOMP, restated (normals_oracle.c).
 * out4: n x (nx, ny, nz, curvature). */
int og_normals_knn(const float* pts, int n, int stride_f, int k, const float vp[3], float* out4, int num_threads);
/* the nodelet's radius mode (normal_computation.cc:73-77) + removeNaNNormalsFromPointCloud (:53-57): neighbours = all
 * points with d2 < float(radius^2) in ascending distance order; fewer than 3 -> NaN row.  valid_idx (nullable, capacity
 * n): the points the nodelet keeps.  Returns their number. */
int og_normals_radius(const float* pts, int n, int stride_f, double radius, const float vp[3], float* out4, int* valid_idx,
                      int num_threads);

/* ------------------------------------------------- NDT (ndt_oracle.c; PARITY UNPINNED, see there) */
typedef struct {
  float resolution;                /* ndt_omp_impl.hpp:50  default 1.0: voxel side AND neighbour radius */
  double step_size;                /* :51  0.1  More-Thuente maximum step */
  double outlier_ratio;            /* :52  0.55 */
  double transformation_epsilon;   /* :93  0.1; LOCUS passes its icp_tf_epsilon (PointCloudOdometry.cc:188) */
  int max_iterations;              /* :94  35;  LOCUS passes its icp_iterations (PointCloudOdometry.cc:190) */
  int min_points_per_voxel;        /* voxel_grid_covariance_omp.h:186  6 */
  double min_covar_eigvalue_mult;  /* voxel_grid_covariance_omp.h:187  0.01 */
  int search_method;               /* 0 KDTREE (default, ndt_omp_impl.hpp:96), 1 DIRECT26, 2 DIRECT7, 3 DIRECT1 */
  int num_threads;
} og_ndt_params;

typedef struct {
  float final_transformation[16];  /* row-major 4x4 */
  int converged, nr_iterations;
  double trans_probability;        /* score / number of source points (ndt_omp_impl.hpp:207) */
  long n_evaluations;              /* computeDerivatives calls */
  double pose[6];                  /* x y z roll pitch yaw of the last iterate */
  int status;
} og_ndt_result;

typedef struct og_ndt_target og_ndt_target;
void og_ndt_default_params(og_ndt_params* p);
/* setInputTarget -> init() -> VoxelGridCovariance::filter(true) */
og_ndt_target* og_ndt_target_build(const float* pts, int n, int stride_f, const og_ndt_params* p);
void og_ndt_target_free(og_ndt_target* t);
/* 0 ok, -1 no finite point, -2 voxel index would overflow int32 */
int og_ndt_target_info(const og_ndt_target* t, int* n_valid, int* n_all, int min_b[3], int div_b[3]);
/* the voxel_centroids_ list (ascending voxel index): n_valid entries each */
void og_ndt_target_leaves(const og_ndt_target* t, int* leaf_idx, int* nr_points, double* mean3, double* icov9, float* centroid3);
/* computeDerivatives on T16 * src with the angle derivatives of p (float path; compute_hessian as the reference's flag) */
int og_ndt_derivatives(const og_ndt_target* t, const float* src, int n, int stride_f, const float* T16, const double p[6],
                       int compute_hessian, double* score, double g[6], double H[36]);
/* computeHessian (double path, what the line search calls after its last trial) */
int og_ndt_hessian(const og_ndt_target* t, const float* src, int n, int stride_f, const float* T16, const double p[6], double H[36]);
/* align(output, guess): guess NULL = identity */
int og_ndt_align(const og_ndt_target* t, const float* src, int n, int stride_f, const float* guess, og_ndt_result* result);
void og_ndt_pose_to_matrix(const double p[6], float T[16]);
void og_ndt_euler_xyz(const float T[16], float out[3]);

#ifdef __cplusplus
}
#endif
#endif
