// odometry_stub.h -- TEST-ONLY stand-ins so that locus_b200/csrc/odometry.cu (the host-side pipeline: threads,
// ring of filtered clouds, ordering, back-pressure, error propagation) can be compiled and exercised on a machine
// without CUDA.  The fake stages below do no geometry: a "scan" is a buffer whose first 8 bytes hold its id, the
// fake VoxelGrid copies that id into its output, and the fake GICP reports which ids it was given -- enough to
// detect a ring slot that was overwritten too early, a wrong source/target pairing or an out-of-order result.
// Never linked into liblocus_b200.so.
#pragma once
#include <atomic>
#include <chrono>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <random>
#include <thread>

#include "../include/locus_b200.h"

// ---- the few CUDA runtime calls odometry.cu makes
typedef int cudaError_t;
typedef void* cudaStream_t;
enum { cudaSuccess = 0, cudaMemcpyDeviceToHost = 2, cudaMemcpyDeviceToDevice = 3, cudaStreamNonBlocking = 1 };
inline cudaError_t cudaSetDevice(int) { return cudaSuccess; }
inline cudaError_t cudaMalloc(void** p, size_t n) { *p = malloc(n); return *p ? cudaSuccess : 2; }
inline cudaError_t cudaFree(void* p) { free(p); return cudaSuccess; }
inline cudaError_t cudaMemcpyAsync(void* d, const void* s, size_t n, int, cudaStream_t) { memcpy(d, s, n); return cudaSuccess; }
inline cudaError_t cudaStreamSynchronize(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithFlags(cudaStream_t* s, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaStreamCreateWithPriority(cudaStream_t* s, int, int) { *s = nullptr; return cudaSuccess; }
inline cudaError_t cudaDeviceGetStreamPriorityRange(int* lo, int* hi) { *lo = 0; *hi = -1; return cudaSuccess; }
inline cudaError_t cudaStreamDestroy(cudaStream_t) { return cudaSuccess; }
inline cudaError_t cudaGetLastError() { return cudaSuccess; }
inline const char* cudaGetErrorString(cudaError_t) { return "stub"; }

namespace lb {
inline char* err_buf() { static thread_local char b[512] = ""; return b; }
inline void set_error(const char* fmt, ...) {
  va_list ap; va_start(ap, fmt); vsnprintf(err_buf(), 512, fmt, ap); va_end(ap);
}
}  // namespace lb
#define LB_TRY(call) do { int s__ = (call); if (s__ != LB_OK) return s__; } while (0)

// ---- fake stages
struct lb_voxel { float leaf = 0.f; uint64_t launches = 0; };
struct lb_gicp { uint64_t src = ~0ull, tgt = ~0ull; size_t n_src = 0, n_tgt = 0; uint64_t launches = 0; lb_gicp_params P; };

namespace stub {
inline std::atomic<int>& aligns_running() { static std::atomic<int> v{0}; return v; }
inline std::atomic<int>& aligns_peak() { static std::atomic<int> v{0}; return v; }
inline std::atomic<int>& max_sleep_us() { static std::atomic<int> v{300}; return v; }
inline void nap(uint64_t seed) {
  int m = max_sleep_us().load();
  if (m > 0) std::this_thread::sleep_for(std::chrono::microseconds((seed * 2654435761u >> 7) % (uint64_t)m));
}
}  // namespace stub

extern "C" {
inline const char* lb_last_error_string(void) { return lb::err_buf(); }
inline int lb_voxel_create(int, lb_voxel** h) { *h = new lb_voxel; return LB_OK; }
inline int lb_voxel_create_on_stream(int, void*, lb_voxel** h) { *h = new lb_voxel; return LB_OK; }
inline int lb_voxel_destroy(lb_voxel* h) { delete h; return LB_OK; }
inline int lb_voxel_launch_count(lb_voxel* h, uint64_t* n) { *n = h->launches; return LB_OK; }
// scan layout of the fakes: [0..7] id, [8..11] flags (bit 0: make the filter fail)
inline int lb_voxel_filter(lb_voxel* h, const uint8_t* data, size_t n_pts, uint32_t point_step, const lb_field*, int,
                           const int32_t*, size_t, uint8_t* out, size_t cap, size_t* n_out, int32_t*, int, int) {
  uint64_t id; uint32_t flags;
  memcpy(&id, data, 8); memcpy(&flags, data + 8, 4);
  stub::nap(id * 3 + 1);
  if (flags & 1u) { lb::set_error("stub voxel failure for scan %llu", (unsigned long long)id); return LB_ERR_INVALID_ARG; }
  size_t m = 3 + (size_t)(id % 5);
  if (m > n_pts) m = n_pts;
  if (m > cap) return LB_ERR_CAPACITY;
  for (size_t i = 0; i < m; i++) memcpy(out + i * point_step, &id, 8);     // every output point carries the scan id
  *n_out = m;
  h->launches += 17;
  return LB_OK;
}
inline int lb_gicp_create(int, lb_gicp** h) { *h = new lb_gicp; return LB_OK; }
inline int lb_gicp_destroy(lb_gicp* h) { delete h; return LB_OK; }
inline int lb_gicp_reserve(lb_gicp*, size_t, int) { return LB_OK; }
inline int lb_gicp_set_params(lb_gicp* h, const lb_gicp_params* p) { h->P = *p; return LB_OK; }
inline int lb_gicp_launch_count(lb_gicp* h, uint64_t* n) { *n = h->launches; return LB_OK; }
inline int lb_gicp_set_source(lb_gicp* h, const void* pts, size_t n, size_t stride, size_t, ptrdiff_t, int) {
  memcpy(&h->src, pts, 8); h->n_src = n;
  uint64_t last; memcpy(&last, (const uint8_t*)pts + (n - 1) * stride, 8);
  if (last != h->src) { lb::set_error("stub: torn source cloud"); return LB_ERR_INVALID_ARG; }
  return LB_OK;
}
inline int lb_gicp_set_target(lb_gicp* h, const void* pts, size_t n, size_t stride, size_t, ptrdiff_t, int, uint64_t*) {
  memcpy(&h->tgt, pts, 8); h->n_tgt = n;
  uint64_t last; memcpy(&last, (const uint8_t*)pts + (n - 1) * stride, 8);
  if (last != h->tgt) { lb::set_error("stub: torn target cloud"); return LB_ERR_INVALID_ARG; }
  return LB_OK;
}
struct lb_cloud { uint64_t id; size_t n; };
inline std::atomic<int>& clouds_alive() { static std::atomic<int> v{0}; return v; }
inline int lb_gicp_prepare_source(lb_gicp* h) { stub::nap(h->src * 11 + 3); h->launches += 12; return LB_OK; }
inline std::atomic<int>& clouds_peak() { static std::atomic<int> v{0}; return v; }
inline int lb_gicp_share_source(lb_gicp* h, lb_cloud** out) {
  *out = new lb_cloud{h->src, h->n_src};
  int now = ++clouds_alive();
  int pk = clouds_peak().load();
  while (now > pk && !clouds_peak().compare_exchange_weak(pk, now)) {}
  return LB_OK;
}
inline int lb_cloud_release(lb_cloud* c) { delete c; --clouds_alive(); return LB_OK; }
inline int lb_gicp_set_target_cloud(lb_gicp* h, lb_cloud* c) { h->tgt = c->id; h->n_tgt = c->n; return LB_OK; }
inline int lb_gicp_align(lb_gicp* h, const float* guess, lb_gicp_result* out) {
  int now = ++stub::aligns_running();
  int pk = stub::aligns_peak().load();
  while (now > pk && !stub::aligns_peak().compare_exchange_weak(pk, now)) {}
  stub::nap(h->src * 7 + 5);
  memset(out, 0, sizeof(*out));
  out->final_transformation[0] = (float)h->src; out->final_transformation[1] = (float)h->tgt;
  out->final_transformation[2] = (float)h->n_src; out->final_transformation[3] = (float)h->n_tgt;
  out->final_transformation[4] = guess ? guess[3] : -1.0f;
  out->iterations = h->P.max_iterations; out->converged = 1;
  h->launches += 40;
  --stub::aligns_running();
  return LB_OK;
}
}
