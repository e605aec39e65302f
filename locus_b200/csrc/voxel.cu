// voxel.cu -- VoxelGrid front-end (kernel K1): hash-and-centroid reduce.
//
// Replaces pcl::VoxelGrid<pcl::PCLPointCloud2>::applyFilter as driven by
// point_cloud_filter/src/custom_voxel_grid.cc:76-87 (index arithmetic cross-
// checked with multithreaded_ndt/voxel_grid_covariance_omp_impl.hpp:67-164).
//
// Pipeline (all on the handle's stream, data resident in HBM):
//   bbox_kernel        finite + filter-limit predicate, min/max reduce
//   vg_keys_kernel     int32 PCL leaf index per point (bit-exact float32 math)
//   radix_sort_pairs   stable LSD sort of (leaf index, point#)
//   vg_heads_kernel + exclusive_scan + vg_segstart_kernel    voxel segments
//   vg_centroid_kernel one thread per voxel: float32 sum in ascending point
//                      order, divide by count, write the output point
// Algorithmic bytes: N_in*point_step + N_out*point_step.
#include <float.h>
#include <math.h>

#include "prims.cuh"

namespace lb {
const char* last_error();

constexpr int VG_MAX_FIELDS = 16;

struct VoxelFieldsDev {
  int n_ff;                          // averaged FLOAT32 fields
  uint32_t ff_off[VG_MAX_FIELDS];
};

__global__ void __launch_bounds__(256)
vg_keys_kernel(const uint8_t* __restrict__ base, uint32_t n, uint32_t stride, uint32_t x_off, uint32_t y_off,
               uint32_t z_off, int ff_off, double lim_min, double lim_max, int negative, float inv0, float inv1,
               float inv2, int min_b0, int min_b1, int min_b2, int mul1, int mul2, uint32_t sentinel,
               BodyBox body, uint32_t* __restrict__ keys, BBoxAcc* acc_to_reset) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i == 0) bbox_reset(acc_to_reset);   // the host has consumed the bounding box: ready for the next call
  if (i >= n) return;
  const uint8_t* p = base + (size_t)i * stride;
  bool ok = true;
  if (ff_off >= 0) {
    double v = (double)*reinterpret_cast<const float*>(p + ff_off);
    bool drop = negative ? (v < lim_max && v > lim_min) : (v > lim_max || v < lim_min);
    ok = !drop;
  }
  float x = *reinterpret_cast<const float*>(p + x_off);
  float y = *reinterpret_cast<const float*>(p + y_off);
  float z = *reinterpret_cast<const float*>(p + z_off);
  ok = ok && isfinite(x) && isfinite(y) && isfinite(z);
  ok = ok && !body_box_drops(body, x, y, z);       // BodyFilter folded into the load predicate (row f4)
  uint32_t key = sentinel;
  if (ok) {
    // static_cast<int>(floor(x * inv_leaf) - float(min_b))   (voxel_grid_covariance_omp_impl.hpp:159-161)
    int ijk0 = (int)(floorf(x * inv0) - (float)min_b0);
    int ijk1 = (int)(floorf(y * inv1) - (float)min_b1);
    int ijk2 = (int)(floorf(z * inv2) - (float)min_b2);
    int idx = ijk0 + ijk1 * mul1 + ijk2 * mul2;
    key = ((uint32_t)idx < sentinel) ? (uint32_t)idx : sentinel;
  }
  keys[i] = key;
}

__global__ void __launch_bounds__(256)
vg_heads_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t sentinel, uint32_t* __restrict__ flags) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  flags[i] = (k < sentinel && (i == 0 || keys[i - 1] != k)) ? 1u : 0u;
}

__global__ void __launch_bounds__(256)
vg_segstart_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ excl, uint32_t n, uint32_t sentinel,
                   uint32_t* __restrict__ seg_start) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  uint32_t k = keys[i];
  if (k < sentinel && (i == 0 || keys[i - 1] != k)) seg_start[excl[i]] = i;
}

// keep[s] = 1 iff voxel s holds >= min_points points
__global__ void __launch_bounds__(256)
vg_keep_kernel(const uint32_t* __restrict__ keys, const uint32_t* __restrict__ seg_start, const uint32_t* n_seg_dev,
               uint32_t n, int min_points, uint32_t* __restrict__ keep) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t n_seg = *n_seg_dev;
  if (s >= n_seg) { if (s < n) keep[s] = 0; return; }
  uint32_t a = seg_start[s];
  uint32_t k = keys[a];
  uint32_t e = a + 1;
  while (e < n && keys[e] == k) e++;
  keep[s] = ((int)(e - a) >= min_points) ? 1u : 0u;
}

// Gather the averaged FLOAT32 fields of every point into voxel-sorted order (one thread per sorted
// position: scattered 4-byte reads with full memory-level parallelism, coalesced writes), so that the
// per-voxel sequential sums below stream over contiguous memory.
__global__ void __launch_bounds__(256)
vg_gather_kernel(const uint8_t* __restrict__ in, uint32_t stride, const uint32_t* __restrict__ vals, uint32_t n,
                 VoxelFieldsDev F, float* __restrict__ sorted_f) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  const uint8_t* p = in + (size_t)vals[s] * stride;
  float* d = sorted_f + (size_t)s * F.n_ff;
  for (int f = 0; f < F.n_ff; f++) d[f] = *reinterpret_cast<const float*>(p + F.ff_off[f]);
}

// The common record layout (x, y, z, intensity averaged: 16-byte records).  A warp owns 32 consecutive voxels, i.e.
// ONE contiguous range of the voxel-sorted records: the warp copies that range through shared memory with
// coalesced 16-byte loads, and every lane then adds up its own voxel's records from shared memory in ascending
// input order (the order defines the float32 rounding) -- no per-lane chains of dependent global loads, which is
// what made the thread-per-voxel kernel below take 50 us for 30 k voxels.
constexpr int VG4_TILE = 256;       // records per warp tile (4 KB)
__global__ void __launch_bounds__(128)
vg_centroid4_kernel(const uint8_t* __restrict__ in, uint32_t stride, const uint32_t* __restrict__ keys,
                    const uint32_t* __restrict__ vals, const float4* __restrict__ rec,
                    const uint32_t* __restrict__ seg_start, const uint32_t* n_seg_dev,
                    const uint32_t* __restrict__ slot /*nullable*/, const uint32_t* __restrict__ keep /*nullable*/,
                    uint32_t n, VoxelFieldsDev F, uint32_t capacity, uint8_t* __restrict__ out,
                    int32_t* __restrict__ out_voxel_idx) {
  __shared__ float4 tile[4][VG4_TILE];
  const int lane = threadIdx.x & 31, wib = threadIdx.x >> 5;
  const uint32_t n_seg = *n_seg_dev;
  const uint32_t s0 = (blockIdx.x * 4 + wib) * 32;
  if (s0 >= n_seg) return;                                  // whole warp
  const uint32_t s = s0 + lane;
  const bool seg = s < n_seg;
  uint32_t a = 0, e = 0, k = 0;
  if (seg) {
    a = seg_start[s];
    k = keys[a];
    // segment end: next segment's start, or the first sentinel key after the last voxel
    if (s + 1 < n_seg) e = seg_start[s + 1];
    else { e = a + 1; while (e < n && keys[e] == k) e++; }
  }
  const uint32_t n_act = min(32u, n_seg - s0);
  const uint32_t r0 = __shfl_sync(0xffffffffu, a, 0);
  const uint32_t r1 = __shfl_sync(0xffffffffu, e, (int)n_act - 1);
  float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
  for (uint32_t t0 = r0; t0 < r1; t0 += VG4_TILE) {
    const uint32_t t1 = min(r1, t0 + (uint32_t)VG4_TILE);
    for (uint32_t i = t0 + lane; i < t1; i += 32) tile[wib][i - t0] = rec[i];
    __syncwarp();
    if (seg) {
      uint32_t lo = max(a, t0), hi = min(e, t1);
      if (lo < hi && lo == a) { acc = tile[wib][lo - t0]; lo++; }      // the sum starts from the first record itself
      for (; lo + 4 <= hi; lo += 4) {
        const float4 v0 = tile[wib][lo - t0], v1 = tile[wib][lo + 1 - t0], v2 = tile[wib][lo + 2 - t0], v3 = tile[wib][lo + 3 - t0];
        acc.x = acc.x + v0.x; acc.y = acc.y + v0.y; acc.z = acc.z + v0.z; acc.w = acc.w + v0.w;
        acc.x = acc.x + v1.x; acc.y = acc.y + v1.y; acc.z = acc.z + v1.z; acc.w = acc.w + v1.w;
        acc.x = acc.x + v2.x; acc.y = acc.y + v2.y; acc.z = acc.z + v2.z; acc.w = acc.w + v2.w;
        acc.x = acc.x + v3.x; acc.y = acc.y + v3.y; acc.z = acc.z + v3.z; acc.w = acc.w + v3.w;
      }
      for (; lo < hi; lo++) {
        const float4 v = tile[wib][lo - t0];
        acc.x = acc.x + v.x; acc.y = acc.y + v.y; acc.z = acc.z + v.z; acc.w = acc.w + v.w;
      }
    }
    __syncwarp();
  }
  if (!seg) return;
  if (keep && !keep[s]) return;
  const uint32_t o = slot ? slot[s] : s;
  if (o >= capacity) return;
  const float cnt = (float)(e - a);
  const uint8_t* first = in + (size_t)vals[a] * stride;
  uint8_t* dst = out + (size_t)o * stride;
  // bytes not covered by an averaged field come from the voxel's first point
  if (stride == 32 && ((reinterpret_cast<uintptr_t>(in) | reinterpret_cast<uintptr_t>(out)) & 15u) == 0) {
    const uint4 w0 = reinterpret_cast<const uint4*>(first)[0], w1 = reinterpret_cast<const uint4*>(first)[1];
    reinterpret_cast<uint4*>(dst)[0] = w0; reinterpret_cast<uint4*>(dst)[1] = w1;
  } else {
    for (uint32_t w = 0; w < stride / 4; w++)
      reinterpret_cast<uint32_t*>(dst)[w] = reinterpret_cast<const uint32_t*>(first)[w];
  }
  *reinterpret_cast<float*>(dst + F.ff_off[0]) = acc.x / cnt;
  *reinterpret_cast<float*>(dst + F.ff_off[1]) = acc.y / cnt;
  *reinterpret_cast<float*>(dst + F.ff_off[2]) = acc.z / cnt;
  *reinterpret_cast<float*>(dst + F.ff_off[3]) = acc.w / cnt;
  if (out_voxel_idx) out_voxel_idx[o] = (int32_t)k;
}

// One thread per voxel: float32 sum of its points in ascending input order (the order the stable sort
// produced), divide by the count (centroid /= float(n), PCL), write the output point.
__global__ void __launch_bounds__(128)
vg_centroid_kernel(const uint8_t* __restrict__ in, uint32_t stride, const uint32_t* __restrict__ keys,
                   const uint32_t* __restrict__ vals, const float* __restrict__ sorted_f,
                   const uint32_t* __restrict__ seg_start, const uint32_t* n_seg_dev,
                   const uint32_t* __restrict__ slot /*nullable*/, const uint32_t* __restrict__ keep /*nullable*/,
                   uint32_t n, VoxelFieldsDev F, uint32_t capacity, uint8_t* __restrict__ out,
                   int32_t* __restrict__ out_voxel_idx) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  const uint32_t n_seg = *n_seg_dev;
  if (s >= n_seg) return;
  if (keep && !keep[s]) return;
  uint32_t o = slot ? slot[s] : s;
  if (o >= capacity) return;
  uint32_t a = seg_start[s];
  uint32_t k = keys[a];
  // segment end: next segment's start, or the first sentinel key after the last voxel
  uint32_t e;
  if (s + 1 < n_seg) e = seg_start[s + 1];
  else { e = a + 1; while (e < n && keys[e] == k) e++; }
  float acc[VG_MAX_FIELDS];
  const int nf = F.n_ff;
  const float* src = sorted_f + (size_t)a * nf;
#pragma unroll
  for (int f = 0; f < VG_MAX_FIELDS; f++) acc[f] = (f < nf) ? src[f] : 0.f;
  if (nf == 4) {
    // the common case (x, y, z, intensity): 16-byte records, batches of 8 loads in flight, then the adds in
    // ascending input order (the order defines the float32 rounding; only the LOADS are hoisted)
    const float4* rec = reinterpret_cast<const float4*>(sorted_f) + a;
    uint32_t j = 1, len = e - a;
    for (; j + 8 <= len; j += 8) {
      float4 v[8];
#pragma unroll
      for (int u = 0; u < 8; u++) v[u] = rec[j + u];
#pragma unroll
      for (int u = 0; u < 8; u++) { acc[0] = acc[0] + v[u].x; acc[1] = acc[1] + v[u].y; acc[2] = acc[2] + v[u].z; acc[3] = acc[3] + v[u].w; }
    }
    for (; j < len; j++) { float4 v = rec[j]; acc[0] = acc[0] + v.x; acc[1] = acc[1] + v.y; acc[2] = acc[2] + v.z; acc[3] = acc[3] + v.w; }
  } else {
    for (uint32_t j = a + 1; j < e; j++) {
      src += nf;
#pragma unroll
      for (int f = 0; f < VG_MAX_FIELDS; f++)
        if (f < nf) acc[f] = acc[f] + src[f];
    }
  }
  float cnt = (float)(e - a);
  const uint8_t* first = in + (size_t)vals[a] * stride;
  uint8_t* dst = out + (size_t)o * stride;
  // bytes not covered by an averaged field come from the voxel's first point
  for (uint32_t w = 0; w < stride / 4; w++)
    reinterpret_cast<uint32_t*>(dst)[w] = reinterpret_cast<const uint32_t*>(first)[w];
#pragma unroll
  for (int f = 0; f < VG_MAX_FIELDS; f++)
    if (f < nf) *reinterpret_cast<float*>(dst + F.ff_off[f]) = acc[f] / cnt;
  if (out_voxel_idx) out_voxel_idx[o] = (int32_t)k;
}

}  // namespace lb

using namespace lb;

struct lb_voxel {
  Ctx c;
  float leaf[3] = {0.f, 0.f, 0.f};
  char filter_field[32] = "";
  double lim_min = -FLT_MAX, lim_max = FLT_MAX;
  int negative = 0;
  int min_points = 0;
  int downsample_all = 1;
  BodyBox body{0, 1.f, 0.f, {0.f, 0.f, 0.f}, {0.f, 0.f, 0.f}};   // BodyFilter nodelet folded in (lb_voxel_set_body_filter)
  DBuf<uint8_t> d_in, d_out;
  DBuf<int32_t> d_vidx;
  DBuf<uint32_t> keys, flags, seg_start, keep, slot;
  DBuf<float> sorted_f;
  BBoxAcc* d_acc = nullptr;     // device
  uint32_t* d_tot = nullptr;    // device [2]
  BBoxAcc* h_acc = nullptr;     // pinned
  uint32_t* h_tot = nullptr;    // pinned [2]
  SortWork sort;
  ScanWork scan;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  float t_last_ms = 0.f;
};

static int voxel_create_impl(int device, void* stream, bool ext, lb_voxel** out) {
  if (!out) { set_error("lb_voxel_create: null handle pointer"); return LB_ERR_INVALID_ARG; }
  lb_voxel* h = new lb_voxel;
  int s = ctx_init(h->c, device, stream, ext);
  if (s != LB_OK) { delete h; return s; }
  if (cudaMalloc((void**)&h->d_acc, sizeof(BBoxAcc)) != cudaSuccess ||
      cudaMalloc((void**)&h->d_tot, 2 * sizeof(uint32_t)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_acc, sizeof(BBoxAcc)) != cudaSuccess ||
      cudaMallocHost((void**)&h->h_tot, 2 * sizeof(uint32_t)) != cudaSuccess ||
      cudaEventCreate(&h->ev0) != cudaSuccess || cudaEventCreate(&h->ev1) != cudaSuccess) {
    set_error("lb_voxel_create: allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
    delete h;
    return LB_ERR_CUDA;
  }
  bbox_init_kernel<<<1, 32, 0, h->c.stream>>>(h->d_acc);
  cudaStreamSynchronize(h->c.stream);
  *out = h;
  return LB_OK;
}

extern "C" {

int lb_voxel_create(int device, lb_voxel** h) { return voxel_create_impl(device, nullptr, false, h); }
int lb_voxel_create_on_stream(int device, void* stream, lb_voxel** h) { return voxel_create_impl(device, stream, true, h); }

int lb_voxel_destroy(lb_voxel* h) {
  if (!h) return LB_OK;
  cudaSetDevice(h->c.device);
  cudaStreamSynchronize(h->c.stream);
  h->d_in.release(); h->d_out.release(); h->d_vidx.release(); h->keys.release(); h->flags.release();
  h->seg_start.release(); h->keep.release(); h->slot.release(); h->sorted_f.release();
  h->sort.ka.release(); h->sort.kb.release(); h->sort.va.release(); h->sort.vb.release(); h->sort.hist.release();
  h->sort.scan.sums.release(); h->scan.sums.release();
  if (h->d_acc) cudaFree(h->d_acc);
  if (h->d_tot) cudaFree(h->d_tot);
  if (h->h_acc) cudaFreeHost(h->h_acc);
  if (h->h_tot) cudaFreeHost(h->h_tot);
  if (h->ev0) cudaEventDestroy(h->ev0);
  if (h->ev1) cudaEventDestroy(h->ev1);
  ctx_destroy(h->c);
  delete h;
  return LB_OK;
}

int lb_voxel_set_leaf_size(lb_voxel* h, float lx, float ly, float lz) {
  if (!h || !(lx > 0.f) || !(ly > 0.f) || !(lz > 0.f)) { set_error("lb_voxel_set_leaf_size: leaf must be > 0"); return LB_ERR_INVALID_ARG; }
  h->leaf[0] = lx; h->leaf[1] = ly; h->leaf[2] = lz;
  return LB_OK;
}
int lb_voxel_get_leaf_size(lb_voxel* h, float* leaf3) {
  if (!h || !leaf3) return LB_ERR_INVALID_ARG;
  leaf3[0] = h->leaf[0]; leaf3[1] = h->leaf[1]; leaf3[2] = h->leaf[2];
  return LB_OK;
}
int lb_voxel_set_filter_limits(lb_voxel* h, const char* field_name, double limit_min, double limit_max, int negative) {
  if (!h) return LB_ERR_INVALID_ARG;
  h->filter_field[0] = 0;
  if (field_name) { strncpy(h->filter_field, field_name, sizeof(h->filter_field) - 1); h->filter_field[sizeof(h->filter_field) - 1] = 0; }
  h->lim_min = limit_min; h->lim_max = limit_max; h->negative = negative ? 1 : 0;
  return LB_OK;
}
int lb_voxel_set_body_filter(lb_voxel* h, int enabled, const float* min3, const float* max3, float rotation_z) {
  if (!h || (enabled && (!min3 || !max3))) { set_error("lb_voxel_set_body_filter: null argument"); return LB_ERR_INVALID_ARG; }
  h->body.enabled = enabled ? 1 : 0;
  if (enabled) {
    // inverse of PCL's getTransformation(0,0,0, 0,0,rz) taken the way Eigen inverts a 3x3: adjugate / determinant
    float A = cosf(rotation_z), B = sinf(rotation_z);
    float det = A * A + B * B;
    h->body.ia = A / det; h->body.ib = B / det;
    for (int d = 0; d < 3; d++) { h->body.mn[d] = min3[d]; h->body.mx[d] = max3[d]; }
  }
  return LB_OK;
}
int lb_voxel_set_min_points_per_voxel(lb_voxel* h, int m) { if (!h) return LB_ERR_INVALID_ARG; h->min_points = m; return LB_OK; }
int lb_voxel_set_downsample_all_data(lb_voxel* h, int all) { if (!h) return LB_ERR_INVALID_ARG; h->downsample_all = all ? 1 : 0; return LB_OK; }
int lb_voxel_launch_count(lb_voxel* h, uint64_t* n) { if (!h || !n) return LB_ERR_INVALID_ARG; *n = h->c.launches; return LB_OK; }
int lb_voxel_kernel_time(lb_voxel* h, float* ms) { if (!h || !ms) return LB_ERR_INVALID_ARG; *ms = h->t_last_ms; return LB_OK; }

int lb_voxel_filter(lb_voxel* h, const uint8_t* data, size_t n_pts, uint32_t point_step, const lb_field* fields,
                    int n_fields, const int32_t* indices, size_t n_indices, uint8_t* out, size_t out_capacity_pts,
                    size_t* n_out, int32_t* out_voxel_idx, int mem_in, int mem_out) {
  (void)indices; (void)n_indices;  // pcl::VoxelGrid<PCLPointCloud2> ignores indices_ too
  if (!h || !n_out) { set_error("lb_voxel_filter: null handle / n_out"); return LB_ERR_INVALID_ARG; }
  *n_out = 0;
  if (n_pts == 0) return LB_OK;
  if (!data || !out || !fields || n_fields <= 0) { set_error("lb_voxel_filter: null data/out/fields"); return LB_ERR_INVALID_ARG; }
  if (point_step < 12 || (point_step & 3u)) { set_error("lb_voxel_filter: point_step must be a multiple of 4 and >= 12"); return LB_ERR_INVALID_ARG; }
  if (n_pts > 0x7ffffff0ull) { set_error("lb_voxel_filter: too many points"); return LB_ERR_INVALID_ARG; }
  if (!(h->leaf[0] > 0.f)) { set_error("lb_voxel_filter: leaf size not set"); return LB_ERR_INVALID_ARG; }
  int xo = -1, yo = -1, zo = -1, ffo = -1;
  VoxelFieldsDev F; F.n_ff = 0;
  for (int f = 0; f < n_fields; f++) {
    const lb_field& fd = fields[f];
    bool is_f32 = fd.datatype == LB_FLOAT32;
    if (!strcmp(fd.name, "x") && is_f32) xo = (int)fd.offset;
    if (!strcmp(fd.name, "y") && is_f32) yo = (int)fd.offset;
    if (!strcmp(fd.name, "z") && is_f32) zo = (int)fd.offset;
    if (h->filter_field[0] && !strcmp(fd.name, h->filter_field)) {
      if (!is_f32) { set_error("lb_voxel_filter: distance filtering requires a FLOAT32 field"); return LB_ERR_UNSUPPORTED; }
      ffo = (int)fd.offset;
    }
    if (fd.offset + 4 > point_step || (fd.offset & 3u)) {
      if (is_f32) { set_error("lb_voxel_filter: field '%s' misaligned / out of range", fd.name); return LB_ERR_INVALID_ARG; }
    }
  }
  if (xo < 0 || yo < 0 || zo < 0) { set_error("lb_voxel_filter: x/y/z FLOAT32 fields required"); return LB_ERR_INVALID_ARG; }
  if (h->filter_field[0] && ffo < 0) { set_error("lb_voxel_filter: filter field '%s' not found", h->filter_field); return LB_ERR_INVALID_ARG; }
  if (h->downsample_all) {
    for (int f = 0; f < n_fields; f++) {
      if (fields[f].datatype != LB_FLOAT32) continue;
      uint32_t cnt = fields[f].count ? fields[f].count : 1;
      for (uint32_t c = 0; c < cnt; c++) {
        if (F.n_ff >= VG_MAX_FIELDS) { set_error("lb_voxel_filter: more than %d float fields", VG_MAX_FIELDS); return LB_ERR_UNSUPPORTED; }
        F.ff_off[F.n_ff++] = fields[f].offset + 4 * c;
      }
    }
  } else {
    F.n_ff = 3; F.ff_off[0] = xo; F.ff_off[1] = yo; F.ff_off[2] = zo;
  }

  // x,y,z must be contiguous for bbox_kernel's 3-float read (true for every PCL point type)
  if (yo != xo + 4 || zo != xo + 8) { set_error("lb_voxel_filter: x,y,z must be consecutive FLOAT32 fields"); return LB_ERR_UNSUPPORTED; }
  Ctx& c = h->c;
  LB_CUDA(cudaSetDevice(c.device));
  const uint32_t n = (uint32_t)n_pts;
  const size_t bytes = (size_t)n * point_step;
  LB_CUDA(cudaEventRecord(h->ev0, c.stream));
  const uint8_t* d_in = data;
  if (mem_in == LB_MEM_HOST) {
    LB_TRY(h->d_in.ensure(bytes));
    LB_CUDA(cudaMemcpyAsync(h->d_in.p, data, bytes, cudaMemcpyHostToDevice, c.stream));
    d_in = h->d_in.p;
  }
  // ---- bounding box of the surviving points (pcl::getMinMax3D with float limits)
  int bb_blocks = min(cdiv(n, 256), c.sm_count * 2);
  bbox_kernel<<<bb_blocks, 256, 0, c.stream>>>(d_in, n, point_step, (uint32_t)xo, ffo, (float)h->lim_min,
                                                (float)h->lim_max, h->negative, h->body, h->d_acc);
  c.launches += 1;
  LB_CUDA(cudaMemcpyAsync(h->h_acc, h->d_acc, sizeof(BBoxAcc), cudaMemcpyDeviceToHost, c.stream));
  LB_CUDA(cudaStreamSynchronize(c.stream));
  if (h->h_acc->count == 0) {
    bbox_init_kernel<<<1, 32, 0, c.stream>>>(h->d_acc);
    *n_out = 0;
    return LB_OK;
  }
  float min_p[3], max_p[3], inv[3];
  for (int d = 0; d < 3; d++) { min_p[d] = ord2f(h->h_acc->mn[d]); max_p[d] = ord2f(h->h_acc->mx[d]); inv[d] = 1.0f / h->leaf[d]; }
  int64_t dx = (int64_t)((max_p[0] - min_p[0]) * inv[0]) + 1;
  int64_t dy = (int64_t)((max_p[1] - min_p[1]) * inv[1]) + 1;
  int64_t dz = (int64_t)((max_p[2] - min_p[2]) * inv[2]) + 1;
  if (dx * dy * dz > (int64_t)INT32_MAX) {
    bbox_init_kernel<<<1, 32, 0, c.stream>>>(h->d_acc);
    set_error("lb_voxel_filter: leaf size too small for the input dataset, integer indices would overflow");
    return LB_ERR_VOXEL_OVERFLOW;
  }
  int min_b[3], max_b[3], div_b[3];
  for (int d = 0; d < 3; d++) {
    min_b[d] = (int)floorf(min_p[d] * inv[d]);
    max_b[d] = (int)floorf(max_p[d] * inv[d]);
    div_b[d] = max_b[d] - min_b[d] + 1;
  }
  int64_t ncells64 = (int64_t)div_b[0] * div_b[1] * div_b[2];
  if (ncells64 > (int64_t)INT32_MAX) { bbox_init_kernel<<<1, 32, 0, c.stream>>>(h->d_acc); set_error("lb_voxel_filter: voxel index overflow"); return LB_ERR_VOXEL_OVERFLOW; }
  uint32_t sentinel = (uint32_t)ncells64;
  int key_bits = 1;
  while (key_bits < 32 && (1ull << key_bits) <= (uint64_t)sentinel) key_bits++;

  LB_TRY(h->keys.ensure(n)); LB_TRY(h->flags.ensure(n)); LB_TRY(h->seg_start.ensure(n));
  vg_keys_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(d_in, n, point_step, xo, yo, zo, ffo, h->lim_min, h->lim_max,
                                                     h->negative, inv[0], inv[1], inv[2], min_b[0], min_b[1], min_b[2],
                                                     div_b[0], div_b[0] * div_b[1], sentinel, h->body, h->keys.p, h->d_acc);
  c.launches++;
  uint32_t *sk = nullptr, *sv = nullptr;
  LB_TRY(radix_sort_pairs(c, h->sort, h->keys.p, nullptr, n, key_bits, &sk, &sv));
  vg_heads_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(sk, n, sentinel, h->flags.p);
  c.launches++;
  LB_TRY(exclusive_scan_u32(c, h->scan, h->flags.p, h->flags.p, n, &h->d_tot[0]));
  vg_segstart_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(sk, h->flags.p, n, sentinel, h->seg_start.p);
  c.launches++;
  const uint32_t* slot = nullptr; const uint32_t* keep = nullptr;
  const uint32_t* n_final_dev = &h->d_tot[0];
  if (h->min_points > 1) {
    LB_TRY(h->keep.ensure(n)); LB_TRY(h->slot.ensure(n));
    vg_keep_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(sk, h->seg_start.p, &h->d_tot[0], n, h->min_points, h->keep.p);
    c.launches++;
    LB_TRY(exclusive_scan_u32(c, h->scan, h->keep.p, h->slot.p, n, &h->d_tot[1]));
    slot = h->slot.p; keep = h->keep.p; n_final_dev = &h->d_tot[1];
  }
  uint8_t* d_out = out; int32_t* d_vidx = out_voxel_idx;
  uint32_t capacity = (uint32_t)(out_capacity_pts < n ? out_capacity_pts : n);
  if (mem_out == LB_MEM_HOST) {
    LB_TRY(h->d_out.ensure(bytes));
    d_out = h->d_out.p; capacity = n;
    if (out_voxel_idx) { LB_TRY(h->d_vidx.ensure(n)); d_vidx = h->d_vidx.p; }
  }
  LB_TRY(h->sorted_f.ensure((size_t)n * F.n_ff));
  vg_gather_kernel<<<cdiv(n, 256), 256, 0, c.stream>>>(d_in, point_step, sv, n, F, h->sorted_f.p);
  if (F.n_ff == 4)
    vg_centroid4_kernel<<<cdiv(n, 128), 128, 0, c.stream>>>(d_in, point_step, sk, sv, reinterpret_cast<const float4*>(h->sorted_f.p),
                                                             h->seg_start.p, &h->d_tot[0], slot, keep, n, F, capacity, d_out, d_vidx);
  else
    vg_centroid_kernel<<<cdiv(n, 128), 128, 0, c.stream>>>(d_in, point_step, sk, sv, h->sorted_f.p, h->seg_start.p,
                                                            &h->d_tot[0], slot, keep, n, F, capacity, d_out, d_vidx);
  c.launches += 2;
  LB_CUDA(cudaGetLastError());
  LB_CUDA(cudaMemcpyAsync(h->h_tot, n_final_dev, sizeof(uint32_t), cudaMemcpyDeviceToHost, c.stream));
  LB_CUDA(cudaStreamSynchronize(c.stream));
  size_t m = h->h_tot[0];
  if (m > out_capacity_pts) {
    set_error("lb_voxel_filter: output capacity %zu < %zu voxels", out_capacity_pts, m);
    return LB_ERR_CAPACITY;
  }
  if (mem_out == LB_MEM_HOST && m > 0) {
    LB_CUDA(cudaMemcpyAsync(out, d_out, m * point_step, cudaMemcpyDeviceToHost, c.stream));
    if (out_voxel_idx) LB_CUDA(cudaMemcpyAsync(out_voxel_idx, d_vidx, m * sizeof(int32_t), cudaMemcpyDeviceToHost, c.stream));
  }
  LB_CUDA(cudaEventRecord(h->ev1, c.stream));
  LB_CUDA(cudaStreamSynchronize(c.stream));
  cudaEventElapsedTime(&h->t_last_ms, h->ev0, h->ev1);
  *n_out = m;
  return LB_OK;
}

}  // extern "C"
