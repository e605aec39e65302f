// prims.cuh -- device-wide primitives shared by the VoxelGrid and GICP paths:
// stream context + growable device buffers, bounding-box reduction, stable LSD
// radix sort of (key, value) pairs, exclusive scan.  All hand-written for
// sm_100a; no CUB/Thrust on the product path.
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../include/locus_b200.h"
#include "hd.h"

namespace lb {

// ------------------------------------------------------------------ errors
void set_error(const char* fmt, ...);

#define LB_CUDA(call)                                                                     \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess) {                                                             \
      lb::set_error("%s:%d: %s -> %s", __FILE__, __LINE__, #call, cudaGetErrorString(e__)); \
      return LB_ERR_CUDA;                                                                 \
    }                                                                                     \
  } while (0)

#define LB_TRY(call)            \
  do {                          \
    int s__ = (call);           \
    if (s__ != LB_OK) return s__; \
  } while (0)

struct Ctx {
  int device = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  uint64_t launches = 0;
  int sm_count = 148;
};

int ctx_init(Ctx& c, int device, void* external_stream, bool use_external);
void ctx_destroy(Ctx& c);

// device allocations made by DBuf::ensure since the library was loaded (a steady-state call path must not add any)
unsigned long long& dbuf_alloc_count();

template <class T>
struct DBuf {
  T* p = nullptr;
  size_t cap = 0;
  int ensure(size_t n) {
    if (n <= cap) return LB_OK;
    dbuf_alloc_count()++;
    static const bool trace = getenv("LB_ALLOC_TRACE") != nullptr;      // tuning aid: which buffer still grows in steady state
    if (trace) fprintf(stderr, "[lb alloc] %zu x %zu bytes requested (capacity was %zu)\n", n, sizeof(T), cap);
    size_t nc = cap ? cap : 1024;
    while (nc < n) nc = nc + nc / 2 + 1024;
    if (p) LB_CUDA(cudaFree(p));
    p = nullptr; cap = 0;
    LB_CUDA(cudaMalloc((void**)&p, nc * sizeof(T)));
    cap = nc;
    return LB_OK;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

static inline int cdiv(long long a, long long b) { return (int)((a + b - 1) / b); }

// ------------------------------------------------------------------ bbox
// ordered-uint encoding of float so that unsigned min/max == float min/max
__device__ __forceinline__ uint32_t f2ord(float f) {
  uint32_t b = __float_as_uint(f);
  return (b & 0x80000000u) ? ~b : (b | 0x80000000u);
}
__host__ __device__ __forceinline__ float ord2f(uint32_t o) {
  uint32_t b = (o & 0x80000000u) ? (o & 0x7fffffffu) : ~o;
#if defined(__CUDA_ARCH__)
  return __uint_as_float(b);
#else
  float f; memcpy(&f, &b, 4); return f;
#endif
}

struct BBoxAcc {          // device-resident accumulator, 8 words
  uint32_t mn[3], mx[3];
  uint32_t count, pad;
};

__global__ void bbox_init_kernel(BBoxAcc* acc);

// accumulate one point into warp-level then global bounding box (call with all 32 lanes; valid = lane has a point)
__device__ __forceinline__ void bbox_warp_accumulate(bool valid, float x, float y, float z, BBoxAcc* acc) {
  uint32_t mn0 = 0xffffffffu, mn1 = 0xffffffffu, mn2 = 0xffffffffu, mx0 = 0, mx1 = 0, mx2 = 0, cnt = 0;
  if (valid) { mn0 = mx0 = f2ord(x); mn1 = mx1 = f2ord(y); mn2 = mx2 = f2ord(z); cnt = 1; }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    mn0 = min(mn0, __shfl_xor_sync(0xffffffffu, mn0, o));
    mn1 = min(mn1, __shfl_xor_sync(0xffffffffu, mn1, o));
    mn2 = min(mn2, __shfl_xor_sync(0xffffffffu, mn2, o));
    mx0 = max(mx0, __shfl_xor_sync(0xffffffffu, mx0, o));
    mx1 = max(mx1, __shfl_xor_sync(0xffffffffu, mx1, o));
    mx2 = max(mx2, __shfl_xor_sync(0xffffffffu, mx2, o));
    cnt += __shfl_xor_sync(0xffffffffu, cnt, o);
  }
  if ((threadIdx.x & 31) == 0 && cnt > 0) {
    atomicMin(&acc->mn[0], mn0); atomicMin(&acc->mn[1], mn1); atomicMin(&acc->mn[2], mn2);
    atomicMax(&acc->mx[0], mx0); atomicMax(&acc->mx[1], mx1); atomicMax(&acc->mx[2], mx2);
    atomicAdd(&acc->count, cnt);
  }
}
__device__ __forceinline__ void bbox_reset(BBoxAcc* acc) {
  acc->mn[0] = acc->mn[1] = acc->mn[2] = 0xffffffffu;
  acc->mx[0] = acc->mx[1] = acc->mx[2] = 0u;
  acc->count = 0; acc->pad = 0;
}

// Finite-point bounding box of a strided cloud (x,y,z float32 at xyz_off).
// Optional limit filter (VoxelGrid getMinMax3D): field value v at ff_off is
// dropped if  negative ? (v < fmax && v > fmin) : (v > fmax || v < fmin), float compare.
__global__ void bbox_kernel(const uint8_t* __restrict__ base, uint32_t n, uint32_t stride, uint32_t xyz_off,
                            int ff_off, float fmin, float fmax, int negative, BodyBox body, BBoxAcc* acc);

// ------------------------------------------------------------------ scan
__device__ __forceinline__ uint32_t warp_incl_scan(uint32_t v) {
#pragma unroll
  for (int o = 1; o < 32; o <<= 1) {
    uint32_t t = __shfl_up_sync(0xffffffffu, v, o);
    if ((threadIdx.x & 31) >= o) v += t;
  }
  return v;
}

// exclusive scan of one value per thread across the block; returns exclusive prefix, total in *total
template <int T>
__device__ __forceinline__ uint32_t block_excl_scan(uint32_t v, uint32_t* total, uint32_t* smem /*T/32+1*/) {
  uint32_t incl = warp_incl_scan(v);
  int w = threadIdx.x >> 5, l = threadIdx.x & 31;
  if (l == 31) smem[w] = incl;
  __syncthreads();
  if (w == 0) {
    uint32_t s = (l < T / 32) ? smem[l] : 0;
    uint32_t si = warp_incl_scan(s);
    if (l < T / 32) smem[l] = si - s;
    if (l == T / 32 - 1) smem[T / 32] = si;
  }
  __syncthreads();
  uint32_t r = smem[w] + incl - v;
  *total = smem[T / 32];
  __syncthreads();
  return r;
}

// Exclusive scan of n uint32 (in may alias out).  total (nullable, device) receives the grand total.
struct ScanWork { DBuf<uint32_t> sums; };
int exclusive_scan_u32(Ctx& c, ScanWork& w, const uint32_t* in, uint32_t* out, size_t n, uint32_t* total_dev);

// ------------------------------------------------------------------ sort
// Stable LSD radix sort of (key, val) pairs on the low `key_bits` bits.
// vals_in == nullptr means val = element index.  Result ends in *keys_out/*vals_out
// (pointers into the ping-pong buffers a/b).
struct SortWork {
  DBuf<uint32_t> ka, kb, va, vb, hist;
  ScanWork scan;
};
int radix_sort_pairs(Ctx& c, SortWork& w, const uint32_t* keys_in, const uint32_t* vals_in, size_t n, int key_bits,
                     uint32_t** keys_out, uint32_t** vals_out);
// Same sort (val = element index) when the key width is only known on the device: passes for max_key_bits are launched,
// those beyond *key_bits_dev return at once.  Result: buffers w.ka / w.va after an odd number of executed passes
// ((*key_bits_dev + 7) / 8), w.kb / w.vb after an even number; consumers pick on the device.
int radix_sort_pairs_devbits(Ctx& c, SortWork& w, const uint32_t* keys_in, size_t n, int max_key_bits, const int* key_bits_dev);

}  // namespace lb
