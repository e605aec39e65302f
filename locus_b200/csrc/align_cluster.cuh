// align_cluster.cuh -- persistent align() kernel, thread-block-cluster variant (LB_EXEC_PERSISTENT_CLUSTER).
//
// The all-SM persistent kernel (gicp_kernels.cuh) pays a grid-wide exchange through L2 per objective evaluation
// (publish -> visible -> poll, ~8 k cycles including the skew of ~60 CTAs), and the BFGS inner solve is a chain of
// ~150 dependent evaluations.  Here the inner solve runs inside ONE thread-block cluster of 16 CTAs (one die):
//   * each solver CTA keeps its 1/16 of the correspondences (source point, matched target point, Mahalanobis
//     matrix: 72 B per point, <= CL_CAP points) in shared memory for the whole inner solve;
//   * the all-reduce of the 13 (BFGS) / 28 (Gauss-Newton) partial sums is a PUSH through distributed shared memory:
//     the leader warp of every CTA sends its partials to all 16 CTAs with st.async (data + transaction count on the
//     receiver's mbarrier in one instruction), then sleeps on its own mbarrier until the 16 x NV words have landed and
//     adds them in rank order -- bitwise identical totals in all 16 leaders, no cluster barrier (which would flush
//     L1, where the leader keeps its BFGS state), no polling, and the worker warps do not wait for the exchange at all;
//   * the other clusters of the grid only take part in the correspondence step (K4), which wants every SM:
//     they wait for a command word published through L2 by the rank-0 leader, search their slice, and report
//     their hit count; that grid-wide exchange happens once per OUTER iteration, not per evaluation.
// Control flow is the same bfgs.h code (leader warp per solver CTA), so results equal the other execution
// modes up to the summation order of the 16 partials.
#pragma once

#include <cooperative_groups.h>

#include "gicp_kernels.cuh"

namespace lb {
namespace cg = cooperative_groups;

constexpr int CL_SIZE = 16;            // CTAs per cluster (non-portable size; checked at handle creation)
constexpr int CL_CAP = 2816;           // correspondences per solver CTA held in shared memory (16 x 2816 = 45056 source points)
constexpr int CL_ACC_WARPS = 7;        // warps 1..7 accumulate (warp 0 is the leader)
constexpr int CL_ACC = CL_ACC_WARPS * 32;
constexpr int CL_CMD_WORDS = 24;       // op, T[12], R[9] (+pad)
constexpr int CL_MAX_CTAS = 160;

struct ClusterCache {                  // dynamic shared memory of a solver CTA: 202,752 bytes
  float px[CL_CAP], py[CL_CAP], pz[CL_CAP], qx[CL_CAP], qy[CL_CAP], qz[CL_CAP];
  double M[6][CL_CAP];
};

struct ClusterShared {                 // static shared memory
  int op;
  int m;
  int hits;
  int cnt;                             // correspondences cached in this CTA
  float T[12];
  double R[9];
  double D[27];
  double red[CL_ACC_WARPS * AL_MAXV];
  double bc[AL_MAXV + 4];
  double rx[2][CL_SIZE][AL_MAXV];      // partials pushed into this CTA by every rank (double-buffered)
  unsigned long long mbar[2];          // one mbarrier per buffer: 1 arrival (own leader) + CL_SIZE * NV * 8 bytes
  double cmd[CL_CMD_WORDS];
  double counts[CL_MAX_CTAS];
  long long t_acc, t_sync, t_gather, n_coll, t_corr, t_scalar, t_mark;   // CTA 0 / thread 0 cycle counters (profiling aid)
};

struct ClusterArgs {
  CorrArgs c;
  SlotWord* gslots;                    // [gridDim.x] hit-count words (tag = command epoch)
  SlotWord* gcmd;                      // [CL_CMD_WORDS] command words published by the rank-0 leader
  unsigned long long epoch_base;       // unique per launch
  OuterParams P;
  float guess[16];
  OuterResult* result;
  long long* debug;
};

// ---- correspondence step, executed by every CTA of the grid --------------------------------------------
__device__ __forceinline__ int cl_nn_slice(const ClusterArgs& a, const float* T, const double* R) {
  int chunk = (a.c.n_src + (int)gridDim.x - 1) / (int)gridDim.x;
  int begin = min(a.c.n_src, (int)blockIdx.x * chunk);
  int end = min(a.c.n_src, begin + chunk);
  return correspond_slice(a.c, T, R, begin, end);
}

// solver CTA: all threads.  On return sh.m holds the global number of correspondences and the CTA's chunk of
// (source, target, M) triples is in the shared-memory cache.
__device__ __forceinline__ void cl_do_correspond(const ClusterArgs& a, ClusterShared& sh, ClusterCache& cache,
                                                 unsigned long long cmd_epoch, int rank) {
  float T[12]; double R[9];
#pragma unroll
  for (int i = 0; i < 12; i++) T[i] = sh.T[i];
#pragma unroll
  for (int i = 0; i < 9; i++) R[i] = sh.R[i];
  if (threadIdx.x == 0) sh.hits = 0;
  __syncthreads();
  int hits = cl_nn_slice(a, T, R);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
  if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&sh.hits, hits);
  __threadfence();                               // release this CTA's corr / M writes (once per outer iteration)
  __syncthreads();
  if (threadIdx.x == 0) slot_store(&a.gslots[blockIdx.x], (double)sh.hits, cmd_epoch);
  // gather the hit counts of every CTA of the grid (this is also the grid-wide "correspondences are written" barrier)
  for (int b = threadIdx.x; b < (int)gridDim.x; b += blockDim.x) {
    double v;
    while (!slot_try(&a.gslots[b], cmd_epoch, v)) {}
    sh.counts[b] = v;
  }
  __syncthreads();
  if (threadIdx.x == 0) {
    double m = 0.0;
    for (int b = 0; b < (int)gridDim.x; b++) m += sh.counts[b];
    sh.m = (int)m;
  }
  __threadfence();                               // acquire: the other CTAs' corr / M writes
  // this solver CTA's chunk of the correspondences -> shared memory (L2 reads: the lines may be stale in L1)
  const int cs = (a.c.n_src + CL_SIZE - 1) / CL_SIZE;
  const int begin = min(a.c.n_src, rank * cs), end = min(a.c.n_src, begin + cs);
  for (int i = threadIdx.x; i < end - begin; i += blockDim.x) {
    const int s = begin + i;
    const float4 c4 = __ldcg(reinterpret_cast<const float4*>(a.c.corr) + s);
    const bool ok = __float_as_int(c4.w) >= 0;
    const float4 p4 = __ldcg(reinterpret_cast<const float4*>(a.c.src) + s);
    cache.px[i] = ok ? p4.x : 0.f; cache.py[i] = ok ? p4.y : 0.f; cache.pz[i] = ok ? p4.z : 0.f;
    cache.qx[i] = ok ? c4.x : 0.f; cache.qy[i] = ok ? c4.y : 0.f; cache.qz[i] = ok ? c4.z : 0.f;
#pragma unroll
    for (int e = 0; e < 6; e++) cache.M[e][i] = ok ? __ldcg(a.c.M + 6 * (size_t)s + e) : 0.0;   // M = 0: exact zero terms
  }
  if (threadIdx.x == 0) sh.cnt = end - begin;
  __syncthreads();
}

// ---- distributed-shared-memory push primitives (sm_90+ PTX)
__device__ __forceinline__ uint32_t cl_smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }
__device__ __forceinline__ uint32_t cl_map_rank(uint32_t local_addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(rank));
  return r;
}
__device__ __forceinline__ void cl_mbar_init(unsigned long long* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(cl_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void cl_mbar_arrive_expect_tx(unsigned long long* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(cl_smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ bool cl_mbar_try_wait(unsigned long long* bar, uint32_t parity) {
  uint32_t ok;
  asm volatile("{\n\t.reg .pred p;\n\tmbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\tselp.u32 %0, 1, 0, p;\n\t}"
               : "=r"(ok) : "r"(cl_smem_u32(bar)), "r"(parity) : "memory");
  return ok != 0;
}
// 8 bytes into another CTA's shared memory + 8 bytes of transaction count on that CTA's mbarrier, one instruction
__device__ __forceinline__ void cl_push_f64(uint32_t remote_data, double v, uint32_t remote_bar) {
  asm volatile("st.async.weak.shared::cluster.mbarrier::complete_tx::bytes.b64 [%0], %1, [%2];"
               ::"r"(remote_data), "l"(__double_as_longlong(v)), "r"(remote_bar) : "memory");
}

// ---- objective evaluation inside the solver cluster ------------------------------------------------------
// All threads: accumulate + CTA reduce.  Then only the LEADER warp exchanges (push + mbarrier wait); the worker warps
// return at once and park on the CTA barrier for the next command.  phase: per-buffer mbarrier parity (leader warp).
template <int NV>
__device__ __forceinline__ void cl_do_objective(ClusterShared& sh, const ClusterCache& cache, int rank, int& flip, uint32_t (&phase)[2]) {
  float T[12];
#pragma unroll
  for (int i = 0; i < 12; i++) T[i] = sh.T[i];
  double acc[NV];
#pragma unroll
  for (int e = 0; e < NV; e++) acc[e] = 0.0;
  const bool prof = (blockIdx.x == 0 && threadIdx.x == 0);
  const long long t0 = prof ? clock64() : 0;
  const int t = (int)threadIdx.x - 32;
  if (t >= 0 && t < CL_ACC) {
    const int cnt = sh.cnt;
    for (int i = t; i < cnt; i += CL_ACC) {
      double M[6];
#pragma unroll
      for (int e = 0; e < 6; e++) M[e] = cache.M[e][i];
      if constexpr (NV == 13) objective_terms(T, cache.px[i], cache.py[i], cache.pz[i], cache.qx[i], cache.qy[i], cache.qz[i], M, acc);
      else gn_terms(T, sh.D, sh.D + 9, sh.D + 18, cache.px[i], cache.py[i], cache.pz[i], cache.qx[i], cache.qy[i], cache.qz[i], M, acc);
    }
  }
  const double tot = block_reduce<NV, CL_ACC_WARPS, 1>(acc, sh.red);     // lanes < NV of warp 0 hold the CTA's partials
  if (threadIdx.x >= 32) return;                                         // workers: back to the command loop
  const long long t1 = prof ? clock64() : 0;
  const int lane = threadIdx.x;
  unsigned long long* bar = &sh.mbar[flip];
  if (lane == 0) cl_mbar_arrive_expect_tx(bar, (uint32_t)(CL_SIZE * NV * sizeof(double)));
  if (lane < NV) {
    const uint32_t l_data = cl_smem_u32(&sh.rx[flip][rank][lane]);
    const uint32_t l_bar = cl_smem_u32(bar);
#pragma unroll
    for (int r = 0; r < CL_SIZE; r++) cl_push_f64(cl_map_rank(l_data, (uint32_t)r), tot, cl_map_rank(l_bar, (uint32_t)r));
  }
  while (!cl_mbar_try_wait(bar, phase[flip])) {}
  phase[flip] ^= 1u;
  const long long t2 = prof ? clock64() : 0;
  if (lane < NV) {
    double x = 0.0;
#pragma unroll
    for (int r = 0; r < CL_SIZE; r++) x += sh.rx[flip][r][lane];         // rank order: identical in all CTAs
    sh.bc[lane] = x;
  }
  flip ^= 1;
  __syncwarp();
  if (prof) { const long long t3 = clock64(); sh.t_acc += t1 - t0; sh.t_sync += t2 - t1; sh.t_gather += t3 - t2; sh.n_coll++; sh.t_mark = t3; }
}

// Backend of bfgs.h for the leader warp of a solver CTA (all 32 lanes call every method together).
struct ClusterBackend {
  const ClusterArgs& a;
  ClusterShared& sh;
  ClusterCache& cache;
  unsigned long long& cmd_epoch;
  int& flip;
  uint32_t (&phase)[2];
  int rank;
  int m;

  __device__ ClusterBackend(const ClusterArgs& a_, ClusterShared& sh_, ClusterCache& cache_,
                            unsigned long long& ce_, int& flip_, uint32_t (&phase_)[2], int rank_)
      : a(a_), sh(sh_), cache(cache_), cmd_epoch(ce_), flip(flip_), phase(phase_), rank(rank_), m(0) {}

  __device__ __forceinline__ void warp_trig(const double* x, Trig& t) {
    const int lane = threadIdx.x & 31;
    const int k = lane % 3;
    double sv = 0.0, cv = 0.0;
    if (lane < 6) {
      double ang = (lane < 3) ? (double)half_angle(x, k) : x[3 + k];
      sv = sin(ang);
      cv = cos(ang);
    }
#pragma unroll
    for (int i = 0; i < 3; i++) {
      t.ch[i] = (float)__shfl_sync(0xffffffffu, cv, i);
      t.sh[i] = (float)__shfl_sync(0xffffffffu, sv, i);
      t.c[i] = __shfl_sync(0xffffffffu, cv, 3 + i);
      t.s[i] = __shfl_sync(0xffffffffu, sv, 3 + i);
    }
  }

  // rank-0 leader: publish a command for the CTAs outside the solver cluster
  __device__ __forceinline__ void publish_cmd(int op, const float* T, const double* R) {
    const int lane = threadIdx.x & 31;
    if (rank == 0 && blockIdx.x == 0 && lane < CL_CMD_WORDS) {
      double v = 0.0;
      if (lane == 0) v = (double)op;
      else if (lane <= 12) v = T ? (double)T[lane - 1] : 0.0;
      else if (lane <= 21) v = R ? R[lane - 13] : 0.0;
      slot_store(&a.gcmd[lane], v, cmd_epoch);
    }
  }

  __device__ int correspond(const float* T, const double* R) {
    const int lane = threadIdx.x & 31;
    cmd_epoch++;
    if (lane < 12) sh.T[lane] = T[lane];
    if (lane < 9) sh.R[lane] = R[lane];
    if (lane == 0) sh.op = OP_CORR;
    publish_cmd(OP_CORR, T, R);
    const long long tc0 = clock64();
    __syncthreads();
    cl_do_correspond(a, sh, cache, cmd_epoch, rank);
    m = sh.m;
    if (blockIdx.x == 0 && threadIdx.x == 0) { const long long tc1 = clock64(); sh.t_corr += tc1 - tc0; sh.t_mark = tc1; }
    return m;
  }

  __device__ void fdf(const double* x, double* f, double* g) {
    const int lane = threadIdx.x & 31;
    Trig t;
    warp_trig(x, t);
    float T[12];
    apply_state_trig(x, t, T);
    if (lane < 12) sh.T[lane] = T[lane];
    if (lane == 0) sh.op = OP_FDF;
    if (blockIdx.x == 0 && threadIdx.x == 0) sh.t_scalar += clock64() - sh.t_mark;   // leader time since the last collective
    __syncthreads();
    cl_do_objective<13>(sh, cache, rank, flip, phase);
    double sums[13];
#pragma unroll
    for (int e = 0; e < 13; e++) sums[e] = sh.bc[e];
    objective_finish_trig(sums, m, t, f, g);
  }

  __device__ int gn(const double* x, double* f, double* b, double* H) {
    const int lane = threadIdx.x & 31;
    Trig t;
    warp_trig(x, t);
    float T[12];
    double D[27];
    apply_state_trig(x, t, T);
    r_derivatives_trig(t, D, D + 9, D + 18);
    if (lane < 12) sh.T[lane] = T[lane];
    if (lane < 27) sh.D[lane] = D[lane];
    if (lane == 0) sh.op = OP_GN;
    __syncthreads();
    cl_do_objective<28>(sh, cache, rank, flip, phase);
    *f = sh.bc[0] / (double)m;
#pragma unroll
    for (int e = 0; e < 6; e++) b[e] = sh.bc[1 + e];
#pragma unroll
    for (int e = 0; e < 21; e++) H[e] = sh.bc[7 + e];
    return 0;
  }
};

__global__ void __launch_bounds__(AL_THREADS, 1)
align_cluster_kernel(const __grid_constant__ ClusterArgs a) {
  extern __shared__ __align__(16) unsigned char cl_smem_raw[];
  ClusterCache& cache = *reinterpret_cast<ClusterCache*>(cl_smem_raw);
  __shared__ ClusterShared sh;
  cg::cluster_group cluster = cg::this_cluster();
  const int rank = (int)cluster.block_rank();
  const bool solver = (blockIdx.x / CL_SIZE) == 0;          // cluster 0 runs the inner solves
  unsigned long long cmd_epoch = a.epoch_base;
  int flip = 0;
  uint32_t phase[2] = {0u, 0u};
  const long long t_begin = clock64();
  if (threadIdx.x == 0) {
    cl_mbar_init(&sh.mbar[0], 1); cl_mbar_init(&sh.mbar[1], 1);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  cluster.sync();        // every CTA's mbarriers exist before anybody pushes into them

  if (!solver) {
    // ---- helper CTA: correspondence slices only
    for (;;) {
      cmd_epoch++;
      if (threadIdx.x < CL_CMD_WORDS) {
        double v;
        while (!slot_try(&a.gcmd[threadIdx.x], cmd_epoch, v)) { __nanosleep(100); }
        sh.cmd[threadIdx.x] = v;
      }
      __syncthreads();
      const int op = (int)sh.cmd[0];
      if (op == OP_EXIT) break;
      float T[12]; double R[9];
#pragma unroll
      for (int i = 0; i < 12; i++) T[i] = (float)sh.cmd[1 + i];
#pragma unroll
      for (int i = 0; i < 9; i++) R[i] = sh.cmd[13 + i];
      if (threadIdx.x == 0) sh.hits = 0;
      __syncthreads();
      int hits = cl_nn_slice(a, T, R);
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) hits += __shfl_xor_sync(0xffffffffu, hits, o);
      if ((threadIdx.x & 31) == 0 && hits) atomicAdd(&sh.hits, hits);
      __threadfence();                             // release this CTA's corr / M writes
      __syncthreads();
      if (threadIdx.x == 0) slot_store(&a.gslots[blockIdx.x], (double)sh.hits, cmd_epoch);
      __syncthreads();
    }
    return;
  }

  // ---- solver CTA
  if (threadIdx.x == 0) { sh.t_acc = 0; sh.t_sync = 0; sh.t_gather = 0; sh.n_coll = 0; sh.t_corr = 0; sh.t_scalar = 0; sh.t_mark = clock64(); }
  if (threadIdx.x < 32) {
    ClusterBackend be(a, sh, cache, cmd_epoch, flip, phase, rank);
    OuterResult r;
    gicp_outer_loop(be, a.P, a.guess, r);
    cmd_epoch++;
    be.publish_cmd(OP_EXIT, nullptr, nullptr);
    if (threadIdx.x == 0) sh.op = OP_EXIT;
    __syncthreads();
    if (blockIdx.x == 0 && threadIdx.x == 0) {
      *a.result = r;
      if (a.debug) {
        a.debug[0] = clock64() - t_begin; a.debug[1] = sh.t_acc; a.debug[2] = sh.t_sync; a.debug[3] = sh.n_coll;
        a.debug[6] = sh.t_scalar; a.debug[7] = sh.t_gather; a.debug[8] = sh.t_corr;
      }
    }
  } else {
    for (;;) {
      __syncthreads();
      const int op = sh.op;
      if (op == OP_EXIT) break;
      if (op == OP_CORR) { cmd_epoch++; cl_do_correspond(a, sh, cache, cmd_epoch, rank); }
      else if (op == OP_FDF) cl_do_objective<13>(sh, cache, rank, flip, phase);
      else cl_do_objective<28>(sh, cache, rank, flip, phase);
    }
  }
  cluster.sync();     // nobody leaves while its shared memory may still be read by a peer
}

}  // namespace lb
