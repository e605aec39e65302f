// gicp_kernels.cuh -- CUDA kernels of the GICP hot path (K2-K6 of SURVEY.md 2.3).
//
//   gather_cloud_kernel      strided caller cloud -> packed float4 (+ normals)
//   grid_keys_kernel         K2: cell key per point
//   grid_reorder_kernel      K2: cell-contiguous float4 copy + cell histogram
//   knn_cov_kernel<K>        K3: k-NN(20) -> moments -> Jacobi -> regularised covariance
//   normal_cov_kernel        K3': covariance from normals (reference default mode)
//   prep_source_kernel       K6: output = guess * input (gicp.hpp:440)
//   nn_corr_kernel           K4: transform, exact 1-NN in the voxel hash, gate, Mahalanobis
//   objective_kernel<NV>     K5: 13 (BFGS) / 28 (GN) double sums, warp-shuffle + fixed-shape tree
//   align_persistent_kernel  K4+K5+BFGS+outer loop resident on the device (cooperative launch)
//   transform_kernel, nn_query_kernel, fitness_kernel   accessor surface (a8/a9)
#pragma once

#include "bfgs.h"
#include "grid.h"
#include "prims.cuh"

namespace lb {

constexpr int AL_THREADS = 512;      // threads per CTA of the align kernels
constexpr int AL_MAXV = 28;          // widest reduction (Gauss-Newton)
constexpr int AL_PSTRIDE = 32;       // doubles per CTA slot in the partials buffer

// ------------------------------------------------------------------ cloud upload
__global__ void __launch_bounds__(256)
gather_cloud_kernel(const uint8_t* __restrict__ base, uint32_t n, uint32_t stride, uint32_t xyz_off, int normal_off,
                    f4* __restrict__ raw, f4* __restrict__ nrm) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const uint8_t* p = base + (size_t)i * stride;
  const float* q = reinterpret_cast<const float*>(p + xyz_off);
  raw[i] = f4{q[0], q[1], q[2], 1.0f};
  if (normal_off >= 0) {
    const float* m = reinterpret_cast<const float*>(p + normal_off);
    nrm[i] = f4{m[0], m[1], m[2], 0.0f};
  }
}

// ------------------------------------------------------------------ K2 index build
struct GridGeom {
  float ox, oy, oz, inv_h, h;
  int nx, ny, nz;
};

__device__ __forceinline__ uint32_t cell_of(const GridGeom& g, float x, float y, float z) {
  int cx = (int)floorf((x - g.ox) * g.inv_h);
  int cy = (int)floorf((y - g.oy) * g.inv_h);
  int cz = (int)floorf((z - g.oz) * g.inv_h);
  cx = min(max(cx, 0), g.nx - 1); cy = min(max(cy, 0), g.ny - 1); cz = min(max(cz, 0), g.nz - 1);
  return (uint32_t)((cz * g.ny + cy) * g.nx + cx);
}

__global__ void __launch_bounds__(256)
grid_keys_kernel(const f4* __restrict__ raw, uint32_t n, GridGeom g, uint32_t* __restrict__ keys) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f4 p = raw[i];
  keys[i] = cell_of(g, p.x, p.y, p.z);
}

// occupancy probe for the automatic cell size: marks cells, counts first touches
__global__ void __launch_bounds__(256)
grid_occupancy_kernel(const uint32_t* __restrict__ keys, uint32_t n, uint32_t* __restrict__ cell_cnt, uint32_t* n_occ) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  uint32_t first = 0;
  if (i < n) first = (atomicAdd(&cell_cnt[keys[i]], 1u) == 0u) ? 1u : 0u;
  uint32_t b = __ballot_sync(0xffffffffu, first);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(n_occ, (uint32_t)__popc(b));
}

__global__ void __launch_bounds__(256)
grid_reorder_kernel(const f4* __restrict__ raw, const uint32_t* __restrict__ sorted_keys,
                    const uint32_t* __restrict__ sorted_vals, uint32_t n, f4* __restrict__ pts,
                    uint32_t* __restrict__ cell_cnt) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  uint32_t i = sorted_vals[s];
  f4 p = raw[i];
  pts[s] = f4{p.x, p.y, p.z, bits_to_float((int32_t)i)};
  atomicAdd(&cell_cnt[sorted_keys[s]], 1u);
}

// ------------------------------------------------------------------ K3 covariances
template <int KMAX>
__global__ void __launch_bounds__(128)
knn_cov_kernel(GridView g, int k, double eps, double* __restrict__ cov) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= (uint32_t)g.n) return;
  f4 q = g.pts[s];
  KnnList<KMAX> L;
  knn<KMAX>(g, q.x, q.y, q.z, k, L);
  double sum[3] = {0., 0., 0.}, m2[6] = {0., 0., 0., 0., 0., 0.};
#pragma unroll
  for (int j = 0; j < KMAX; j++) {
    if (j < L.cnt) {
      f4 p = g.pts[L.si[j]];
      // pt.x * pt.x is a float product accumulated into double (gicp.hpp:115-126)
      sum[0] += p.x; sum[1] += p.y; sum[2] += p.z;
      m2[0] += p.x * p.x; m2[1] += p.y * p.x; m2[2] += p.y * p.y;
      m2[3] += p.z * p.x; m2[4] += p.z * p.y; m2[5] += p.z * p.z;
    }
  }
  double out[6];
  cov_from_moments(sum, m2, k, eps, out);
  double* d = cov + 6 * (size_t)s;
#pragma unroll
  for (int e = 0; e < 6; e++) d[e] = out[e];
}

__global__ void __launch_bounds__(256)
normal_cov_kernel(const f4* __restrict__ pts, const f4* __restrict__ nrm, uint32_t n, double eps, double* __restrict__ cov) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  int i = float_to_bits(pts[s].w);
  f4 m = nrm[i];
  double out[6];
  cov_from_normal(m.x, m.y, m.z, eps, out);
  double* d = cov + 6 * (size_t)s;
#pragma unroll
  for (int e = 0; e < 6; e++) d[e] = out[e];
}

// ------------------------------------------------------------------ K6 transforms
struct Mat34 { float m[12]; };
struct Mat33d { double m[9]; };

__global__ void __launch_bounds__(256)
prep_source_kernel(const f4* __restrict__ pts, uint32_t n, Mat34 G, f4* __restrict__ work) {
  uint32_t s = blockIdx.x * blockDim.x + threadIdx.x;
  if (s >= n) return;
  f4 p = pts[s];
  float x, y, z;
  xform_pcl(G.m, p.x, p.y, p.z, x, y, z);
  work[s] = f4{x, y, z, p.w};
}

// out[i] = T * raw[i]  (original order), written into a strided layout; normals rotated if present
__global__ void __launch_bounds__(256)
transform_kernel(const f4* __restrict__ raw, const f4* __restrict__ nrm, uint32_t n, Mat34 T, uint8_t* __restrict__ out,
                 uint32_t stride, uint32_t xyz_off, int normal_off) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  f4 p = raw[i];
  float x, y, z;
  xform_pcl(T.m, p.x, p.y, p.z, x, y, z);
  float* q = reinterpret_cast<float*>(out + (size_t)i * stride + xyz_off);
  q[0] = x; q[1] = y; q[2] = z;
  if (normal_off >= 0 && nrm) {
    f4 m = nrm[i];
    // pcl::transformPointCloudWithNormals: rotation only (so3), p0 + (p1 + p2)
    float* r = reinterpret_cast<float*>(out + (size_t)i * stride + normal_off);
    r[0] = T.m[0] * m.x + (T.m[1] * m.y + T.m[2] * m.z);
    r[1] = T.m[4] * m.x + (T.m[5] * m.y + T.m[6] * m.z);
    r[2] = T.m[8] * m.x + (T.m[9] * m.y + T.m[10] * m.z);
  }
}

// exact 1-NN of arbitrary queries in a grid (PointCloudLocalization.cc:327-336)
__global__ void __launch_bounds__(128)
nn_query_kernel(GridView g, const uint8_t* __restrict__ q, uint32_t n, uint32_t stride, int32_t* __restrict__ idx,
                float* __restrict__ d2) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* p = reinterpret_cast<const float*>(q + (size_t)i * stride);
  int bo; float bd;
  int s = nn1(g, p[0], p[1], p[2], 3.0e38f, bo, bd);
  idx[i] = (s >= 0) ? bo : -1;
  d2[i] = bd;
}

// ------------------------------------------------------------------ block / grid reductions
// Fixed-shape, bitwise run-to-run deterministic reduction of NV doubles per thread:
// warp shuffle tree -> per-warp slots in shared memory -> warp 0 tree.  Result valid in
// thread 0 (red[0..NV-1] in shared memory after the trailing barrier).
template <int NV, int THREADS>
__device__ __forceinline__ void block_reduce(double* v, double* red /*[THREADS/32][NV]*/) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
#pragma unroll
  for (int e = 0; e < NV; e++) {
    double x = v[e];
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
    if (lane == 0) red[warp * NV + e] = x;
  }
  __syncthreads();
  if (warp == 0) {
    constexpr int NW = THREADS / 32;
#pragma unroll
    for (int e = 0; e < NV; e++) {
      double x = (lane < NW) ? red[lane * NV + e] : 0.0;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
      if (lane == 0) v[e] = x;
    }
  }
  __syncthreads();
}

// Sum NV values over `nb` CTA slots of the partials buffer in a fixed order.  Called by one full warp
// per value group; result returned in lane 0.
__device__ __forceinline__ double reduce_slots(const double* __restrict__ partials, int nb, int e, int lane) {
  double x = 0.0;
  for (int b = lane; b < nb; b += 32) x += __ldcg(&partials[(size_t)b * AL_PSTRIDE + e]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) x += __shfl_down_sync(0xffffffffu, x, o);
  return x;
}

// ------------------------------------------------------------------ K4 correspondences
struct CorrArgs {
  GridView tgt;
  const double* tgt_cov;   // sorted-target order, 6 per point
  const f4* src;           // guess-transformed source, sorted order
  const double* src_cov;   // 6 per point
  int n_src;
  float max_d2;            // gate, float-exact equivalent of corr_dist^2 (gicp.hpp:438,483)
  f4* corr;                // matched target point, w = sorted target index or -1
  double* M;               // 6 per source point
};

__device__ __forceinline__ int correspond_point(const CorrArgs& a, const float* T, const double* R, int s) {
  f4 p = a.src[s];
  float qx, qy, qz;
  xform(T, p.x, p.y, p.z, qx, qy, qz);
  int bo; float bd;
  int j = nn1(a.tgt, qx, qy, qz, a.max_d2, bo, bd);
  double M[6] = {0., 0., 0., 0., 0., 0.};
  f4 c = f4{0.f, 0.f, 0.f, bits_to_float(-1)};
  if (j >= 0) {
    f4 t = a.tgt.pts[j];
    c = f4{t.x, t.y, t.z, bits_to_float(j)};
    double C1[6], C2[6];
    const double* c1 = a.src_cov + 6 * (size_t)s;
    const double* c2 = a.tgt_cov + 6 * (size_t)j;
#pragma unroll
    for (int e = 0; e < 6; e++) { C1[e] = c1[e]; C2[e] = c2[e]; }
    mahalanobis(R, C1, C2, M);
  }
  a.corr[s] = c;
  double* m = a.M + 6 * (size_t)s;
#pragma unroll
  for (int e = 0; e < 6; e++) m[e] = M[e];
  return j >= 0 ? 1 : 0;
}

__global__ void __launch_bounds__(128)
nn_corr_kernel(CorrArgs a, Mat34 T, Mat33d R, int* __restrict__ m_count) {
  int s = blockIdx.x * blockDim.x + threadIdx.x;
  int hit = 0;
  if (s < a.n_src) hit = correspond_point(a, T.m, R.m, s);
  uint32_t b = __ballot_sync(0xffffffffu, hit);
  if ((threadIdx.x & 31) == 0 && b) atomicAdd(m_count, __popc(b));
}

// ------------------------------------------------------------------ K5 objective
struct ObjArgs {
  const f4* src;
  const f4* corr;
  const double* M;
  int n_src;
};

// accumulate the NV sums of this thread's points.  NV = 13: BFGS objective; 28: Gauss-Newton.
template <int NV>
__device__ __forceinline__ void objective_accumulate(const ObjArgs& a, const float* T, const double* dP, const double* dT,
                                                     const double* dS, int first, int step, double* acc) {
  for (int s = first; s < a.n_src; s += step) {
    f4 c = a.corr[s];
    if (float_to_bits(c.w) < 0) continue;
    f4 p = a.src[s];
    double M[6];
    const double* m = a.M + 6 * (size_t)s;
#pragma unroll
    for (int e = 0; e < 6; e++) M[e] = m[e];
    if constexpr (NV == 13) objective_terms(T, p.x, p.y, p.z, c.x, c.y, c.z, M, acc);
    else gn_terms(T, dP, dT, dS, p.x, p.y, p.z, c.x, c.y, c.z, M, acc);
  }
}

struct Vec6d { double v[6]; };

// Host-driven objective: every CTA reduces its slice into partials[cta]; the last CTA to
// finish sums the slots in fixed order and writes the NV totals to `out` (mapped pinned memory).
template <int NV>
__global__ void __launch_bounds__(AL_THREADS)
objective_kernel(ObjArgs a, Vec6d x, double* __restrict__ partials, unsigned* __restrict__ ticket, double* __restrict__ out) {
  __shared__ double red[(AL_THREADS / 32) * NV];
  __shared__ float sT[12];
  __shared__ double sD[27];
  __shared__ bool last;
  if (threadIdx.x == 0) {
    apply_state(x.v, sT);
    if (NV != 13) r_derivatives(x.v, sD, sD + 9, sD + 18);
  }
  __syncthreads();
  float T[12];
#pragma unroll
  for (int i = 0; i < 12; i++) T[i] = sT[i];
  double acc[NV];
#pragma unroll
  for (int e = 0; e < NV; e++) acc[e] = 0.0;
  objective_accumulate<NV>(a, T, sD, sD + 9, sD + 18, blockIdx.x * AL_THREADS + threadIdx.x, gridDim.x * AL_THREADS, acc);
  block_reduce<NV, AL_THREADS>(acc, red);
  if (threadIdx.x == 0) {
#pragma unroll
    for (int e = 0; e < NV; e++) __stcg(&partials[(size_t)blockIdx.x * AL_PSTRIDE + e], acc[e]);
    __threadfence();
    unsigned t = atomicAdd(ticket, 1u);
    last = (t == gridDim.x - 1);
  }
  __syncthreads();
  if (last) {
    __threadfence();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = warp; e < NV; e += AL_THREADS / 32) {
      double x = reduce_slots(partials, gridDim.x, e, lane);
      if (lane == 0) out[e] = x;
    }
    if (threadIdx.x == 0) *ticket = 0;
  }
}

// ------------------------------------------------------------------ persistent align
struct AlignArgs {
  CorrArgs c;
  double* partials;       // 2 buffers x gridDim.x x AL_PSTRIDE doubles
  unsigned* barrier;      // zeroed before launch
  OuterParams P;
  float guess[16];
  OuterResult* result;
};

struct DeviceBackend {
  const AlignArgs& a;
  double* red;            // shared: [AL_THREADS/32][AL_MAXV]
  double* bc;             // shared: broadcast slots [AL_MAXV + 8]
  unsigned epoch;
  int flip;
  int m;

  __device__ DeviceBackend(const AlignArgs& a_, double* red_, double* bc_) : a(a_), red(red_), bc(bc_), epoch(0), flip(0), m(0) {}

  __device__ __forceinline__ void grid_barrier() {
    __syncthreads();
    if (threadIdx.x == 0) {
      epoch++;
      unsigned target = epoch * gridDim.x;
      __threadfence();
      atomicAdd(a.barrier, 1u);
      while (true) {
        unsigned v;
        asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(a.barrier) : "memory");
        if (v >= target) break;
      }
      __threadfence();
    }
    __syncthreads();
  }

  // grid-wide deterministic sum of NV doubles per thread; totals land in bc[0..NV-1] for every thread
  template <int NV>
  __device__ __forceinline__ void grid_reduce(double* acc) {
    block_reduce<NV, AL_THREADS>(acc, red);
    double* buf = a.partials + (size_t)flip * gridDim.x * AL_PSTRIDE;
    if (threadIdx.x == 0) {
#pragma unroll
      for (int e = 0; e < NV; e++) __stcg(&buf[(size_t)blockIdx.x * AL_PSTRIDE + e], acc[e]);
    }
    grid_barrier();
    int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    for (int e = warp; e < NV; e += AL_THREADS / 32) {
      double x = reduce_slots(buf, gridDim.x, e, lane);
      if (lane == 0) bc[e] = x;
    }
    flip ^= 1;
    __syncthreads();
  }

  __device__ int correspond(const float* T, const double* R) {
    double cnt[1] = {0.0};
    for (int s = blockIdx.x * AL_THREADS + threadIdx.x; s < a.c.n_src; s += gridDim.x * AL_THREADS)
      cnt[0] += (double)correspond_point(a.c, T, R, s);
    grid_reduce<1>(cnt);
    m = (int)bc[0];
    __syncthreads();
    return m;
  }

  __device__ void fdf(const double* x, double* f, double* g) {
    float T[12];
    if (threadIdx.x == 0) {
      float t[12];
      apply_state(x, t);
      float* sT = reinterpret_cast<float*>(bc + AL_MAXV);
#pragma unroll
      for (int i = 0; i < 12; i++) sT[i] = t[i];
    }
    __syncthreads();
    {
      const float* sT = reinterpret_cast<const float*>(bc + AL_MAXV);
#pragma unroll
      for (int i = 0; i < 12; i++) T[i] = sT[i];
    }
    double acc[13];
#pragma unroll
    for (int e = 0; e < 13; e++) acc[e] = 0.0;
    ObjArgs oa{a.c.src, a.c.corr, a.c.M, a.c.n_src};
    objective_accumulate<13>(oa, T, nullptr, nullptr, nullptr, blockIdx.x * AL_THREADS + threadIdx.x,
                             gridDim.x * AL_THREADS, acc);
    grid_reduce<13>(acc);
    if (threadIdx.x == 0) {
      double sums[13], ff, gg[6];
#pragma unroll
      for (int e = 0; e < 13; e++) sums[e] = bc[e];
      objective_finish(sums, m, x, &ff, gg);
      bc[16] = ff;
#pragma unroll
      for (int e = 0; e < 6; e++) bc[17 + e] = gg[e];
    }
    __syncthreads();
    *f = bc[16];
#pragma unroll
    for (int e = 0; e < 6; e++) g[e] = bc[17 + e];
    __syncthreads();
  }

  __device__ int gn(const double* x, double* f, double* b, double* H) {
    // derivative matrices and T are cheap: every thread builds them (identical inputs -> identical bits)
    float T[12];
    double dP[9], dT[9], dS[9];
    apply_state(x, T);
    r_derivatives(x, dP, dT, dS);
    double acc[28];
#pragma unroll
    for (int e = 0; e < 28; e++) acc[e] = 0.0;
    ObjArgs oa{a.c.src, a.c.corr, a.c.M, a.c.n_src};
    objective_accumulate<28>(oa, T, dP, dT, dS, blockIdx.x * AL_THREADS + threadIdx.x, gridDim.x * AL_THREADS, acc);
    grid_reduce<28>(acc);
    *f = bc[0] / (double)m;
#pragma unroll
    for (int e = 0; e < 6; e++) b[e] = bc[1 + e];
#pragma unroll
    for (int e = 0; e < 21; e++) H[e] = bc[7 + e];
    __syncthreads();
    return 0;
  }
};

__global__ void __launch_bounds__(AL_THREADS, 1)
align_persistent_kernel(const __grid_constant__ AlignArgs a) {
  __shared__ double red[(AL_THREADS / 32) * AL_MAXV];
  __shared__ double bc[AL_MAXV + 8];
  DeviceBackend be(a, red, bc);
  OuterResult r;
  gicp_outer_loop(be, a.P, a.guess, r);
  if (blockIdx.x == 0 && threadIdx.x == 0) *a.result = r;
}

// ------------------------------------------------------------------ fitness (a9)
// pcl::Registration::getFitnessScore: mean of squared 1-NN distances <= max_range.
__global__ void __launch_bounds__(128)
fitness_kernel(GridView g, const f4* __restrict__ raw, uint32_t n, Mat34 T, double max_range,
               double* __restrict__ partials /*[gridDim.x][2]*/) {
  __shared__ double red[(128 / 32) * 2];
  double acc[2] = {0.0, 0.0};
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) {
    f4 p = raw[i];
    float x, y, z;
    xform_pcl(T.m, p.x, p.y, p.z, x, y, z);
    int bo; float bd;
    int s = nn1(g, x, y, z, 3.0e38f, bo, bd);
    if (s >= 0 && (double)bd <= max_range) { acc[0] = (double)bd; acc[1] = 1.0; }
  }
  block_reduce<2, 128>(acc, red);
  if (threadIdx.x == 0) { partials[2 * (size_t)blockIdx.x] = acc[0]; partials[2 * (size_t)blockIdx.x + 1] = acc[1]; }
}

}  // namespace lb
