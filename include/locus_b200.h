/*
 * locus_b200.h -- C ABI of the B200-native GICP scan matcher and VoxelGrid
 * front-end that drops in behind LOCUS's registration and filter seams.
 *
 * Plain C, plain pointers and sizes; no C++/torch/PCL types cross this
 * boundary and no exception ever does: every function returns an lb_status
 * (0 = OK, < 0 = error; lb_last_error_string() gives the text).  On any error
 * the handle keeps its previous state (the last good pose stays valid), which
 * is how the reference behaves on solver failure (gicp.hpp:542-547).
 *
 * Reference interfaces each entry point replaces (paths relative to the
 * LOCUS repository):
 *
 *   lb_gicp_create/destroy        make_shared<MultithreadedGeneralizedIterativeClosestPoint>
 *                                 point_cloud_odometry/src/PointCloudOdometry.cc:142-146
 *                                 point_cloud_localization/src/PointCloudLocalization.cc:229-232
 *   lb_gicp_set_params            setTransformationEpsilon / setMaxCorrespondenceDistance /
 *                                 setMaximumIterations / setRANSACIterations /
 *                                 setMaximumOptimizerIterations / setNumThreads /
 *                                 enableTimingOutput / RecomputeTarget|SourceCovariance /
 *                                 setEuclideanFitnessEpsilon
 *                                 PointCloudOdometry.cc:147-155, PointCloudLocalization.cc:234-245,
 *                                 multithreaded_gicp/include/multithreaded_gicp/gicp.h:111-143,264-298
 *   lb_gicp_set_source            icp_->setInputSource()   PointCloudOdometry.cc:265,
 *                                 PointCloudLocalization.cc:306 (gicp.h:162-179)
 *   lb_gicp_set_target            icp_->setInputTarget()   PointCloudOdometry.cc:266,
 *                                 PointCloudLocalization.cc:307 (gicp.h:196-200)
 *   lb_gicp_align                 icp_->align() + getFinalTransformation() + hasConverged()
 *                                 PointCloudOdometry.cc:267-269, PointCloudLocalization.cc:309-313
 *                                 (computeTransformation, gicp.hpp:405-617)
 *   lb_gicp_transform_source      the `output` cloud of align() (gicp.hpp:586) and
 *                                 pcl::transformPointCloudWithNormals, PointCloudLocalization.cc:325
 *   lb_gicp_nn_target             icp_->getSearchMethodTarget()->nearestKSearch(pt, 1, ..)
 *                                 PointCloudLocalization.cc:327-336
 *   lb_gicp_fitness               icp_->getFitnessScore()
 *                                 point_cloud_odometry/test/test_point_cloud_odometry.cpp:298
 *   lb_gicp_point2plane_information   normalizePCloud + ComputeAp_ForPoint2PlaneICP
 *                                 point_cloud_localization/src/utils.cc:106-128, PointCloudLocalization.cc:694-750
 *   lb_gicp_compute_normals       NormalComputation::filter, k-NN mode (pcl::NormalEstimationOMP)
 *                                 point_cloud_filter/src/normal_computation.cc:26-59
 *   lb_gicp_compute_normals_radius   the same nodelet in its radius mode + removeNaNNormalsFromPointCloud
 *                                 normal_computation.cc:53-57,73-77
 *   lb_voxel_create/destroy       pcl::VoxelGrid<pcl::PCLPointCloud2> impl_
 *                                 point_cloud_filter/include/point_cloud_filter/custom_voxel_grid.h:25
 *   lb_voxel_set_leaf_size        impl_.setLeafSize()      custom_voxel_grid.cc:62-73, 97-101
 *   lb_voxel_set_filter_limits    impl_.setFilterFieldName/Limits/LimitsNegative  custom_voxel_grid.cc:103-134
 *   lb_voxel_filter               impl_.setInputCloud + setIndices + filter()     custom_voxel_grid.cc:76-87
 *   lb_ndt_*                      pclomp::NormalDistributionsTransform behind `registration_method: ndt`
 *                                 PointCloudOdometry.cc:182-195, PointCloudLocalization.cc:267-280 (section NDT below)
 *   lb_odometry_submit/next       the per-scan chain of the lidar callback: filtered scan -> odometry_.SetLidar()
 *                                 -> odometry_.UpdateEstimate()   locus/src/Locus.cc:451-453,
 *                                 PointCloudOdometry.cc:221-230 (SetLidar), 237-247 (UpdateEstimate: first scan
 *                                 only stored; afterwards reference_ = previous query_), 249-274 (UpdateICP)
 *
 * Threading: a handle is used from one thread at a time (the reference calls
 * the seam from the dedicated lidar spinner thread, locus/src/Locus.cc:63-69,
 * and CustomVoxelGrid::filter holds mutex_, custom_voxel_grid.cc:79).  One
 * handle owns one CUDA stream; independent scan streams use independent
 * handles (one per GPU in the multi-GPU mode).
 *
 * Ownership: input buffers are borrowed for the duration of the call only;
 * the library owns all device memory.  `mem` says where a buffer lives.
 */
#ifndef LOCUS_B200_H_
#define LOCUS_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define LB_VERSION 100

typedef enum lb_status {
  LB_OK = 0,
  LB_ERR_INVALID_ARG = -1,
  LB_ERR_CUDA = -2,
  LB_ERR_NO_DEVICE = -3,
  LB_ERR_EMPTY_SOURCE = -4,       /* gicp.h:164-171: empty source cloud -> error, state untouched */
  LB_ERR_NO_TARGET = -5,
  LB_ERR_TOO_FEW_POINTS = -6,     /* gicp.hpp:72-79: k_correspondences > cloud size */
  LB_ERR_VOXEL_OVERFLOW = -7,     /* leaf size too small: int32 voxel index would overflow */
  LB_ERR_CAPACITY = -8,           /* output buffer too small */
  LB_ERR_UNSUPPORTED = -9,
  LB_ERR_NO_ALIGN = -10           /* result requested before any successful align() */
} lb_status;

typedef enum lb_mem { LB_MEM_HOST = 0, LB_MEM_DEVICE = 1 } lb_mem;
typedef enum lb_optimizer { LB_OPT_BFGS = 0, LB_OPT_GAUSS_NEWTON = 1 } lb_optimizer;
/* STREAM_ORDERED (default): one outer iteration = three launches on the handle's stream -- the correspondence search as
 * full-occupancy grids (candidates staged through shared memory with TMA bulk copies, the undecided queries of the
 * whole cloud queued and finished by a second grid), then the inner solve as a cooperative grid (leader warp + worker
 * warps per CTA, grid-wide all-reduce through L2) -- with the loop state resident in device memory and a few
 * iterations enqueued ahead of the host (kernels after convergence return at once).
 * PERSISTENT: the whole of align() as ONE cooperative kernel (the default of round 1; 10 % slower per align and 16 %
 * slower in the pipeline on B200).  HOST_DRIVEN sequences one launch per objective evaluation from the host.
 * PERSISTENT_CLUSTER runs the inner solve inside one 16-CTA thread-block cluster (all-reduce pushed through
 * distributed shared memory; sources <= 45056 points, falls back to PERSISTENT otherwise).
 * All four run the same device functions with the same reduction shape: identical bits. */
typedef enum lb_execution { LB_EXEC_PERSISTENT = 0, LB_EXEC_HOST_DRIVEN = 1, LB_EXEC_PERSISTENT_CLUSTER = 2,
                            LB_EXEC_STREAM_ORDERED = 3 } lb_execution;

#define LB_NO_NORMALS ((ptrdiff_t)-1)

int lb_version(void);
const char* lb_last_error_string(void);
const char* lb_status_string(int status);
int lb_device_count(int* n);

/* ------------------------------------------------------------------ GICP */
typedef struct lb_gicp lb_gicp;

typedef struct lb_gicp_params {
  /* values and defaults: gicp.h:111-127 */
  int k_correspondences;                /* 20   setCorrespondenceRandomness */
  double gicp_epsilon;                  /* 1e-3 */
  double rotation_epsilon;              /* 2e-3 setRotationEpsilon */
  double transformation_epsilon;        /* 5e-4 setTransformationEpsilon */
  double max_correspondence_distance;   /* 5.0  setMaxCorrespondenceDistance */
  int max_iterations;                   /* 200  setMaximumIterations */
  int max_optimizer_iterations;         /* 20   setMaximumOptimizerIterations */
  /* 1: k-NN + SVD covariances (gicp.hpp:85-154).  0: from normals
   * (gicp.hpp:81-82) -- needs normals in the cloud, else falls back to k-NN. */
  int recompute_source_covariance;      /* reference default 0; here 1 (normals are optional) */
  int recompute_target_covariance;
  int optimizer;                        /* lb_optimizer; LB_OPT_BFGS = reference-exact */
  int execution;                        /* lb_execution; LB_EXEC_STREAM_ORDERED by default */
  /* accepted for interface compatibility, no effect on the GPU path */
  double euclidean_fitness_epsilon;     /* setEuclideanFitnessEpsilon (unused by gicp.hpp too) */
  int ransac_iterations;                /* setRANSACIterations(0) */
  int num_threads;                      /* setNumThreads */
  int enable_timing_output;             /* enableTimingOutput: spans are always in lb_gicp_result */
  float index_cell_size;                /* voxel-hash cell size in metres; 0 = automatic (a pure function of the cloud) */
  int align_points_per_cta;             /* source points per CTA of the align solve kernel (any execution); 0 = 512 (lowest latency
                                           of one align: ~60 SMs for a 30k-point scan).  Larger values (1024, 2048) use
                                           fewer SMs per align for longer: more aligns fit on the GPU at once, which is
                                           what lb_odometry's workers want.  Changes the shape of the reduction, i.e. the
                                           last bits of the pose, not the algorithm. */
} lb_gicp_params;

typedef struct lb_gicp_result {
  float final_transformation[16];  /* row-major 4x4 = previous_transformation_ * guess (gicp.hpp:583) */
  int converged;                   /* hasConverged(): delta < 1 OR iteration cap (gicp.hpp:566-568) */
  int iterations;                  /* nr_iterations_ */
  int n_correspondences;           /* of the last outer iteration */
  double delta;                    /* last convergence ratio (gicp.hpp:526-541) */
  int n_objective_evals;           /* fused f+gradient passes */
  int n_inner_iterations;
  /* CUDA-event spans, same three spans the reference times (gicp.hpp:588-616) */
  float t_covariances_ms;
  float t_iterations_ms;
  float t_total_ms;
  int status;
  /* previous_transformation_ / transformation_ of pcl::Registration after align(): the guess-free increment the loop
   * converged to (final_transformation = transformation * guess, gicp.hpp:583); getLastIncrementalTransformation() */
  float transformation[16];
} lb_gicp_result;

int lb_gicp_default_params(lb_gicp_params* p);
int lb_gicp_create(int device, lb_gicp** h);
/* same, but work is enqueued on the caller's cudaStream_t (so the caller can
 * bracket it with its own CUDA events) */
int lb_gicp_create_on_stream(int device, void* cuda_stream, lb_gicp** h);
int lb_gicp_destroy(lb_gicp* h);
int lb_gicp_set_params(lb_gicp* h, const lb_gicp_params* p);
int lb_gicp_get_params(lb_gicp* h, lb_gicp_params* p);

/* pts: n points, `stride` bytes apart; float32 x,y,z at byte offset xyz_off;
 * float32 normal_x,y,z at normal_off (LB_NO_NORMALS if absent).
 * Builds the voxel-hash index of the cloud and (lazily, in align) its
 * covariances.  Invalidates previous covariances like gicp.h:177-178,196-200. */
int lb_gicp_set_source(lb_gicp* h, const void* pts, size_t n, size_t stride, size_t xyz_off,
                       ptrdiff_t normal_off, int mem);
int lb_gicp_set_target(lb_gicp* h, const void* pts, size_t n, size_t stride, size_t xyz_off,
                       ptrdiff_t normal_off, int mem, uint64_t* generation /* nullable */);
/* Optional: pre-size the handle for clouds of up to max_points points (its current clouds, its scratch, and
 * `spare_clouds` spare cloud objects, which set_source / set_target rotate through), so that a stream in steady state
 * performs no device allocation -- an allocation synchronises the whole device and stalls every other handle working on
 * it.  lb_odometry does this for its workers. */
int lb_gicp_reserve(lb_gicp* h, size_t max_points, int spare_clouds);
/* scan-to-scan odometry (PointCloudOdometry.cc:243-244 copies the last scan into
 * reference_): re-use the current source's index and covariances as the next
 * target instead of rebuilding them. */
int lb_gicp_promote_source_to_target(lb_gicp* h);

/* Prepared clouds shared between handles.  In scan-to-scan odometry every filtered scan is the source of one
 * registration and the target of the next (PointCloudOdometry.cc:243-244: copyPointCloud(*query_, *reference_)).  The
 * reference recomputes the target's kd-tree and covariances anyway (setInputTarget clears them, gicp.h:196-200);
 * these entry points let a caller that runs registrations on several handles (lb_odometry's workers) compute each
 * scan's index + covariances once: the handle that has the scan as its SOURCE prepares it and shares it, the handle
 * that needs it as TARGET adopts it.  Same bits as rebuilding it (the index is a pure function of the cloud).
 *   lb_gicp_prepare_source   enqueue index + covariances of the current source now (align would do it lazily)
 *   lb_gicp_share_source     a new reference to the prepared source (release it with lb_cloud_release)
 *   lb_gicp_set_target_cloud adopt a shared cloud as target (no copy; this handle's stream waits for its preparation)
 * A shared cloud is immutable: a handle that still holds it and gets new data switches to another object.  The
 * handles must sit on the same device and use the same covariance parameters. */
typedef struct lb_cloud lb_cloud;
int lb_gicp_prepare_source(lb_gicp* h);
int lb_gicp_share_source(lb_gicp* h, lb_cloud** out);
int lb_gicp_set_target_cloud(lb_gicp* h, lb_cloud* cloud);
int lb_cloud_release(lb_cloud* cloud);

/* guess: row-major 4x4 float, NULL = identity (the callers pass none). */
int lb_gicp_align(lb_gicp* h, const float* guess, lb_gicp_result* out);

/* out_pts[i] = T * source[i] (and rotated normals if normal_off >= 0), in the
 * source's original order and layout.  T NULL = final transformation. */
int lb_gicp_transform_source(lb_gicp* h, const float* T, void* out_pts, size_t stride, size_t xyz_off,
                             ptrdiff_t normal_off, int mem);
/* exact 1-NN of n query points in the current target: original target index
 * and float32 squared distance. */
int lb_gicp_nn_target(lb_gicp* h, const void* xyz, size_t n, size_t stride, int32_t* idx, float* d2, int mem);
/* pcl::Registration::getFitnessScore(max_range) for transform T (NULL = final). */
int lb_gicp_fitness(lb_gicp* h, const float* T, double max_range, double* score);
/* SURVEY 8f row f1 -- the step right after align() in PointCloudLocalization::MeasurementUpdate:
 * normalizePCloud (point_cloud_localization/src/utils.cc:106-128) followed by ComputeAp_ForPoint2PlaneICP
 * (PointCloudLocalization.cc:694-750): Ap = sum_i H_i' H_i, H_i = [a_i x n_i, n_i], a_i = (normalised) query point,
 * n_i = normal of reference point correspondences[i] (rotated by T's rotation when T != NULL), NaN rows skipped.
 * query: n points (xyz at q_xyz_off); reference: n_ref points (normal at r_normal_off); correspondences: n int32
 * indices into reference.  Ap36: 36 doubles row-major (host).  All buffers host or all device (`mem`). */
int lb_gicp_point2plane_information(lb_gicp* h, const void* query, size_t n, size_t q_stride, size_t q_xyz_off,
                                    const void* reference, size_t n_ref, size_t r_stride, size_t r_normal_off,
                                    const int32_t* correspondences, const float* T, int normalize, double* Ap36, int mem);
/* SURVEY 8f row f2 -- the step right before the path: point_cloud_filter::NormalComputation::filter in its k-NN
 * mode (point_cloud_filter/src/normal_computation.cc:26-59; norm_est_ = pcl::NormalEstimationOMP<PointXYZI, Normal>,
 * normal_computation.h:38; k = normal_search_knn, cfg/NormalComputation.cfg:15).  Per point of the cloud currently set
 * as source (which = 0) or target (which = 1): the k nearest neighbours (itself included) -> PCL's float32
 * mean/covariance -> eigenvector of the smallest eigenvalue (PCL eigen33) -> flipped towards `viewpoint` (NULL = the
 * origin, what fromROSMsg leaves in sensor_origin_).  out4: n x (normal_x, normal_y, normal_z, curvature) float32 in
 * the caller's point order, host or device.  3 <= k <= 20.  (Radius mode: lb_gicp_compute_normals_radius.) */
int lb_gicp_compute_normals(lb_gicp* h, int which, int k, const float* viewpoint, float* out4, int mem);
/* The nodelet's radius mode (normal_computation.cc:73-77, norm_est_.setRadiusSearch(normal_search_radius)) together
 * with the NaN-normal removal that follows it (:53-57, pcl::removeNaNNormalsFromPointCloud).  Neighbourhood of a point =
 * every point with squared distance < float(radius^2) (FLANN's strict compare), itself included, taken in ascending
 * distance order; fewer than three neighbours -> NaN normal and curvature (pcl::computePointNormal).  out4: n x 4 as
 * above (NaN rows included); valid_idx (nullable, capacity n): ascending indices of the points whose normal is finite,
 * i.e. the points of the nodelet's output cloud; n_valid (nullable): their number.  LB_ERR_CAPACITY when a neighbourhood
 * exceeds 2048 points. */
int lb_gicp_compute_normals_radius(lb_gicp* h, int which, double radius, const float* viewpoint, float* out4,
                                   int32_t* valid_idx, size_t* n_valid, int mem);
/* covariances used by the last align, n x 9 doubles row-major, original point order. which: 0 source, 1 target */
int lb_gicp_get_covariances(lb_gicp* h, int which, double* out9, size_t capacity_points);
/* number of points currently indexed. which: 0 source, 1 target */
int lb_gicp_cloud_size(lb_gicp* h, int which, size_t* n);
/* kernels launched by this handle since creation (bench.py reports it) */
int lb_gicp_launch_count(lb_gicp* h, uint64_t* n);
/* CUDA-event duration (ms, averaged per launch) of a named kernel class since the last reset:
 * "nn_corr", "objective", "knn_cov", "align_persistent" (the iteration part of align() in any execution mode),
 * "loop_nn" (search grids), "loop_solve", "index_build", "nn_query". */
int lb_gicp_kernel_time(lb_gicp* h, const char* name, float* ms_avg, uint64_t* launches);
/* enable_timing: 0 off; 1 event timers of every kernel class plus the align kernel's cycle counters;
 * 2 only the event pair around the align kernel (cheap enough for a throughput run) */
int lb_gicp_reset_kernel_times(lb_gicp* h, int enable_timing);

/* ------------------------------------------------------------- VoxelGrid */
typedef struct lb_voxel lb_voxel;

/* sensor_msgs/PointField datatypes */
enum { LB_INT8 = 1, LB_UINT8 = 2, LB_INT16 = 3, LB_UINT16 = 4, LB_INT32 = 5, LB_UINT32 = 6,
       LB_FLOAT32 = 7, LB_FLOAT64 = 8 };

typedef struct lb_field {
  char name[32];
  uint32_t offset;
  uint8_t datatype;
  uint32_t count;
} lb_field;

int lb_voxel_create(int device, lb_voxel** h);
int lb_voxel_create_on_stream(int device, void* cuda_stream, lb_voxel** h);
int lb_voxel_destroy(lb_voxel* h);
int lb_voxel_set_leaf_size(lb_voxel* h, float lx, float ly, float lz);
int lb_voxel_get_leaf_size(lb_voxel* h, float* leaf3);
/* field_name NULL or "" = no filter field.  Defaults: none, [-FLT_MAX, FLT_MAX], not negative. */
int lb_voxel_set_filter_limits(lb_voxel* h, const char* field_name, double limit_min, double limit_max, int negative);
/* SURVEY 8f row f4 -- the BodyFilter nodelet that runs right before the voxel grid
 * (point_cloud_filter/src/body_filter.cc:28-56: pcl::CropBox<PointXYZI>, setMin/setMax from cfg/BodyFilter.cfg,
 * setRotation((0, 0, rotation)), setNegative(true)) folded into the filter's load predicate: points inside the box
 * (rotated by rotation_z about z) are removed before voxelisation.  enabled = 0 restores the plain VoxelGrid. */
int lb_voxel_set_body_filter(lb_voxel* h, int enabled, const float* min3, const float* max3, float rotation_z);
int lb_voxel_set_min_points_per_voxel(lb_voxel* h, int min_points);
int lb_voxel_set_downsample_all_data(lb_voxel* h, int all);

/* data: n_pts points of point_step bytes described by `fields` (must contain
 * FLOAT32 x, y, z).  indices: accepted for signature compatibility with
 * pcl_ros::Filter::filter(); pcl::VoxelGrid<PCLPointCloud2> ignores them and
 * so does this call.  out: capacity out_capacity_pts points of point_step
 * bytes; *n_out = number of voxels written (ascending voxel index).
 * out_voxel_idx (nullable, host or device like out): int32 PCL leaf index of
 * each output point. */
int lb_voxel_filter(lb_voxel* h, const uint8_t* data, size_t n_pts, uint32_t point_step,
                    const lb_field* fields, int n_fields, const int32_t* indices, size_t n_indices,
                    uint8_t* out, size_t out_capacity_pts, size_t* n_out, int32_t* out_voxel_idx,
                    int mem_in, int mem_out);
/* SURVEY 8f row f4, the rest of the front end folded into the filter's loads.
 * lb_voxel_set_input_passthrough: what each lidar's pcl/PassThrough nodelet does before everything else
 *   (locus/launch/locus.launch:90-133: filter_field_name z, limits [-100, 100], output_frame base_link): points that
 *   are not finite or whose field lies outside [min, max] (inside, when negative) are dropped -- tested on the RAW
 *   value, in the sensor frame, before the input's transform is applied.  field_name NULL or "" = off (default).
 * lb_voxel_filter_merged: 1..3 input clouds treated as one, in the order a, b, c -- the concatenation
 *   point_cloud_merger performs (point_cloud_merger/src/PointCloudMerger.cc:150-178: a + (b + c)); `transform`
 *   (row-major 4x4 float, NULL = identity) is the sensor -> base_link transform the PassThrough nodelet applies to its
 *   output (pcl_ros::transformPointCloud: Eigen Matrix4f * Vector4f per point).  Then, on the transformed
 *   coordinates: the BodyFilter box, the VoxelGrid limits and the grid itself, exactly as lb_voxel_filter.  Needs the
 *   x, y, z, intensity layout (four averaged FLOAT32 fields) when more than one input or a transform is given.
 *   The merger's optional random / radius-outlier filters (off in the shipped config) are not part of this call. */
typedef struct lb_voxel_input {
  const uint8_t* data;        /* n_pts points of point_step bytes (host or device: mem_in) */
  size_t n_pts;
  const float* transform;     /* 16 floats, row-major; NULL = identity */
} lb_voxel_input;
int lb_voxel_set_input_passthrough(lb_voxel* h, const char* field_name, double limit_min, double limit_max, int negative);
int lb_voxel_filter_merged(lb_voxel* h, const lb_voxel_input* inputs, int n_inputs, uint32_t point_step,
                           const lb_field* fields, int n_fields, uint8_t* out, size_t out_capacity_pts, size_t* n_out,
                           int32_t* out_voxel_idx, int mem_in, int mem_out);
int lb_voxel_launch_count(lb_voxel* h, uint64_t* n);
int lb_voxel_kernel_time(lb_voxel* h, float* ms_total_last_call);
/* CUDA-event duration per call (first launch .. result available), averaged over the calls since the last reset */
int lb_voxel_kernel_time_avg(lb_voxel* h, float* ms_avg, uint64_t* calls, int reset);

/* ------------------------------------------------- resident rolling submap
 * The local map LOCUS keeps in its external point_cloud_mapper object, resident in HBM, in the fixed frame:
 *   lb_submap_insert      mapper_->InsertPoints(cloud, incremental)   locus/src/Locus.cc:464-465,531-532: a point enters
 *                         the map iff no map point occupies its voxel (edge = `resolution`) yet; points are taken in
 *                         input order, so within one call the first point of a voxel wins
 *   lb_submap_neighbors   mapper_->ApproxNearestNeighbors(scan, neighbors)   Locus.cc:479-483: for every query point
 *                         the nearest map point (EXACT here; ties -> lowest map index)
 *   lb_submap_crop_box    mapper_->Refresh(current_pose)   Locus.cc:536-537, lo_settings.yaml:58 box_filter_size: keeps
 *                         the points with |p - centre| <= half_size on every axis (pcl::CropBox bounds are inclusive)
 *   lb_gicp_set_target_submap   the map itself as the registration target (BASELINE configs[2]): its voxel-hash index
 *                         is rebuilt only after the map changed, and the k-NN covariance of a map point is computed
 *                         once -- by the first registration after its insertion, from the map as it is then -- and
 *                         cached until the point leaves the window.  (The reference recomputes every target
 *                         covariance for every scan: setInputTarget clears them, gicp.h:196-200.)
 * The mapper package is not vendored in the reference tree; its octree anchoring and the "approximate" of its
 * nearest-neighbour search are therefore unpinned: voxels are world-anchored here (floor(p / resolution)).
 * oracle/submap_oracle.py restates these semantics for the tests.  Map indices are insertion-order positions; they
 * shift when a crop removes points.  One handle is used from one thread at a time. */
typedef struct lb_submap lb_submap;
int lb_submap_create(int device, float resolution, lb_submap** out);
int lb_submap_destroy(lb_submap* m);
int lb_submap_clear(lb_submap* m);
/* pts: n points (float32 x,y,z at xyz_off, `stride` bytes apart), fixed frame.  n_inserted (nullable): points added.
 * inserted_xyz (nullable): the added points, n_inserted x 3 float32, input order (capacity n; host or device like pts):
 * the `incremental_points` output of InsertPoints.  Non-finite points are skipped. */
int lb_submap_insert(lb_submap* m, const void* pts, size_t n, size_t stride, size_t xyz_off, int mem, size_t* n_inserted,
                     float* inserted_xyz);
int lb_submap_crop_box(lb_submap* m, const float* center3, float half_size, size_t* n_removed /* nullable */);
int lb_submap_size(lb_submap* m, size_t* n);
/* changes whenever the point set changed (insert that added points, crop that removed points, clear) */
int lb_submap_generation(lb_submap* m, uint64_t* generation);
/* the map's points, insertion order, n x 3 float32 */
int lb_submap_points(lb_submap* m, float* xyz_out, size_t capacity_points, int mem);
/* neighbors_xyz (nullable): n x 3 float32, the nearest map point of every query; idx / d2 (nullable): its map index and
 * float32 squared distance.  All buffers host or all device (`mem`). */
int lb_submap_neighbors(lb_submap* m, const void* query, size_t n, size_t stride, size_t xyz_off, float* neighbors_xyz,
                        int32_t* idx, float* d2, int mem);
int lb_submap_launch_count(lb_submap* m, uint64_t* n);
/* the submap as target of `h` (same device; k_correspondences <= 20).  Call again after the map changed; while it is
 * unchanged the call only checks the generation.  Uses h's k_correspondences / gicp_epsilon for the covariances. */
int lb_gicp_set_target_submap(lb_gicp* h, lb_submap* m);

/* ------------------------------------------------- scan-to-scan odometry pipeline
 * One robot's lidar stream, scans submitted in order: VoxelGrid(scan k) -> GICP(source = filtered k, target =
 * filtered k-1), i.e. what Locus.cc:451-453 + PointCloudOdometry::UpdateEstimate do per scan.  Results are
 * identical to calling lb_voxel_filter / lb_gicp_set_source / lb_gicp_set_target / lb_gicp_align per scan; what
 * the pipeline adds is overlap: a voxel stage (one host thread, one CUDA stream) and `depth` registration workers
 * (one host thread, one lb_gicp handle and stream each).  Scan k+1 is filtered and indexed while scan k is still in
 * its align kernels -- latency-bound cooperative grids that occupy 30-60 of the 148 SMs -- and up to `depth`
 * aligns are in flight at once.  Registration k does not depend on the pose of registration k-1 (the caller's
 * prior, if any, comes from IMU/odometry: PointCloudOdometry.cc:252-262), so the overlap changes no result.
 * In the reference the voxel filter already runs in its own nodelet thread ahead of the odometry thread.
 *
 * Threading: submit/next/drain are called from one thread (the lidar callback).  A submitted scan buffer (and the
 * optional filtered_out buffer) must stay valid until its result has been returned by lb_odometry_next. */
typedef struct lb_odometry lb_odometry;

typedef struct lb_odometry_result {
  uint64_t ticket;          /* 0-based submission number */
  int status;               /* lb_status of the stages of this scan */
  int has_pose;             /* 0 for the first scan (UpdateEstimate only stores it) or when a stage failed */
  size_t n_filtered;        /* points after VoxelGrid */
  lb_gicp_result gicp;      /* as lb_gicp_align (valid when has_pose) */
  char error[160];          /* message when status != LB_OK */
} lb_odometry_result;

/* depth: registration workers (1..16).  max_points / max_point_step size the ring of filtered clouds. */
int lb_odometry_create(int device, int depth, size_t max_points, uint32_t max_point_step, lb_odometry** out);
int lb_odometry_destroy(lb_odometry* h);
/* the pipeline's VoxelGrid handle: configure leaf / limits on it before the first submit */
lb_voxel* lb_odometry_voxel(lb_odometry* h);
/* registration worker i's handle (0 <= i < depth): read-only use for counters / kernel times while idle */
lb_gicp* lb_odometry_gicp(lb_odometry* h, int i);
int lb_odometry_depth(lb_odometry* h);
/* applied to every worker's lb_gicp handle (call while the pipeline is idle) */
int lb_odometry_set_gicp_params(lb_odometry* h, const lb_gicp_params* p);
/* Cloud sharing (ON by default).  on != 0: every filtered scan's index + covariances are computed once, by the worker
 * that registers it as source, and adopted as target by the worker of the next scan (lb_gicp_share_source /
 * lb_gicp_set_target_cloud) instead of being rebuilt there.  Same poses, bit for bit (the index is a pure function of
 * the cloud).  on == 0: both clouds are rebuilt for every registration, which is what the reference's callers make
 * the reference do (setInputTarget clears the target's covariances, gicp.h:196-200).  Call while the pipeline is idle,
 * before the first scan. */
int lb_odometry_set_cloud_sharing(lb_odometry* h, int on);
/* scan: n_pts points of point_step bytes described by fields (as lb_voxel_filter); xyz must be FLOAT32 fields.
 * guess: row-major 4x4 prior handed to align() (NULL = identity).  filtered_out (nullable): host (LB_MEM_HOST) or
 * device buffer of at least n_pts * point_step bytes that receives the filtered cloud.  Blocks while 2*depth+2 scans
 * are already in flight.  *ticket (nullable) = the submission number. */
int lb_odometry_submit(lb_odometry* h, const uint8_t* scan, size_t n_pts, uint32_t point_step, const lb_field* fields,
                       int n_fields, int mem, const float* guess, uint8_t* filtered_out, int mem_filtered,
                       uint64_t* ticket);
/* next result in submission order.  block != 0: wait for it; block == 0: returns 1 when it is not ready yet.
 * Returns LB_ERR_NO_ALIGN when every submitted scan has already been returned. */
int lb_odometry_next(lb_odometry* h, lb_odometry_result* r, int block);
/* scans submitted but not yet returned by lb_odometry_next */
int lb_odometry_pending(lb_odometry* h, size_t* n);
/* kernels launched by all stages since creation */
int lb_odometry_launch_count(lb_odometry* h, uint64_t* n);
/* host wall-clock accounting since creation, to see which stage bounds the throughput:
 * out6 = { scans filtered, VoxelGrid stage busy s, its wait s (for a scan or a free ring slot),
 *          scans registered, registration workers busy s (summed over workers), their wait s (for a filtered scan) } */
int lb_odometry_stage_times(lb_odometry* h, double* out6);

/* ------------------------------------------------------------------ NDT
 * LOCUS's alternative registration method (`registration_method: ndt`, registration_settings.h:3-12): the OpenMP NDT
 * fork multithreaded_gicp/include/multithreaded_ndt/ndt_omp{.h,_impl.hpp} + voxel_grid_covariance_omp{.h,_impl.hpp},
 * set up at PointCloudOdometry.cc:182-195 and PointCloudLocalization.cc:267-280 and then driven through the same
 * icp_->setInputSource / setInputTarget / align / getFinalTransformation calls as GICP.
 *
 *   lb_ndt_create/destroy     make_shared<pclomp::NormalDistributionsTransform<PointF, PointF>>
 *   lb_ndt_set_params         setTransformationEpsilon / setMaximumIterations / setResolution / setStepSize /
 *                             setOulierRatio / setNeighborhoodSearchMethod (ndt_omp.h:124-196) and the voxel grid's
 *                             setMinPointPerVoxel / setCovEigValueInflationRatio (voxel_grid_covariance_omp.h:203-236)
 *   lb_ndt_set_source         icp_->setInputSource()
 *   lb_ndt_set_target         icp_->setInputTarget() -> init() (ndt_omp.h:116-119,257-262): the target's voxel Gaussians
 *   lb_ndt_align              icp_->align() + getFinalTransformation() + hasConverged() + getTransformationProbability()
 *                             (computeTransformation, ndt_omp_impl.hpp:100-208)
 *   lb_ndt_target_voxels      getTargetCells().getLeaves() restricted to the searchable voxels (voxel_centroids_ order)
 *   lb_ndt_derivatives        computeDerivatives / computeHessian at a given pose (ndt_omp_impl.hpp:221-343,641-718)
 *
 * Semantics kept: non-finite target points are skipped; a source with a non-finite point is refused (previous source
 * kept); a target whose int32 voxel index would overflow is refused (previous target kept); voxels with fewer than
 * min_points_per_voxel points are not searchable; a voxel whose covariance fails the eigenvalue / inverse check stays in
 * the radius search with nr_points = -1 exactly as in the reference (voxel_grid_covariance_omp_impl.hpp:326-355). */
typedef struct lb_ndt lb_ndt;

typedef struct lb_ndt_params {
  float resolution;                 /* 1.0   voxel side and neighbour radius (ndt_omp_impl.hpp:50) */
  double step_size;                 /* 0.1   More-Thuente maximum step (:51) */
  double outlier_ratio;             /* 0.55  (:52) */
  double transformation_epsilon;    /* 0.1   (:93); LOCUS passes icp_tf_epsilon / tf_epsilon */
  int max_iterations;               /* 35    (:94); LOCUS passes icp_iterations / iterations */
  int min_points_per_voxel;         /* 6     voxel_grid_covariance_omp.h:186 */
  double min_covar_eigvalue_mult;   /* 0.01  voxel_grid_covariance_omp.h:187 */
  int search_method;                /* 0 KDTREE (default, :96), 1 DIRECT26, 2 DIRECT7, 3 DIRECT1 (pclomp::NeighborSearchMethod values) */
  /* accepted for interface compatibility (PointCloudOdometry.cc:189,191-193), no effect -- as in the reference's NDT */
  double max_correspondence_distance;
  int ransac_iterations;
  int num_threads;
  int enable_timing_output;
} lb_ndt_params;

typedef struct lb_ndt_result {
  float final_transformation[16];   /* row-major 4x4 */
  int converged;
  int nr_iterations;
  int n_evaluations;                /* computeDerivatives passes */
  int n_target_voxels;              /* searchable voxels of the target */
  double trans_probability;         /* score / source points (ndt_omp_impl.hpp:207) */
  double pose[6];                   /* x y z roll pitch yaw of the last iterate */
  double t_total_s;
} lb_ndt_result;

void lb_ndt_default_params(lb_ndt_params* p);
int lb_ndt_create(int device, lb_ndt** h);
int lb_ndt_create_on_stream(int device, void* cuda_stream, lb_ndt** h);
int lb_ndt_destroy(lb_ndt* h);
int lb_ndt_set_params(lb_ndt* h, const lb_ndt_params* p);
int lb_ndt_set_source(lb_ndt* h, const void* pts, size_t n, size_t stride_bytes, size_t xyz_offset_bytes, int mem);
int lb_ndt_set_target(lb_ndt* h, const void* pts, size_t n, size_t stride_bytes, size_t xyz_offset_bytes, int mem);
/* guess16: row-major 4x4 float, NULL = identity (what LOCUS passes).  Blocking. */
int lb_ndt_align(lb_ndt* h, const float* guess16, lb_ndt_result* result);
/* host output buffers of `capacity` voxels each (any may be NULL); *n_voxels = searchable voxels, ascending voxel index */
int lb_ndt_target_voxels(lb_ndt* h, size_t capacity, size_t* n_voxels, int32_t* leaf_idx, int32_t* nr_points, double* mean3,
                         double* icov9, float* centroid3);
/* sums over the source transformed by T16 (row-major 4x4), angle derivatives taken at pose6.  compute_hessian: 0 / 1 = the
 * float path without / with the Hessian, 2 = the double-precision Hessian pass the line search ends with (score and
 * gradient come back as 0). */
int lb_ndt_derivatives(lb_ndt* h, const float* T16, const double* pose6, int compute_hessian, double* score, double* gradient6,
                       double* hessian36);
int lb_ndt_launch_count(lb_ndt* h, uint64_t* n);

#ifdef __cplusplus
}
#endif
#endif /* LOCUS_B200_H_ */
