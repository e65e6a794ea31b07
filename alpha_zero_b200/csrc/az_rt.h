// az_rt.h — thin runtime layer under the host code of the engine.
// CUDA build: cudaMalloc / cudaMemcpyAsync / kernel launches on the engine's stream.
// -DAZ_EMU build (tests/emu only): the "device" is host memory and a "kernel" is a loop over warps,
// so the host-side logic and the per-game device routines can be exercised without a GPU.
#pragma once
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>

extern thread_local std::string g_az_error;
static inline int az_fail(int code, const std::string& msg) {
  g_az_error = msg;
  return code;
}

#ifdef AZ_EMU
typedef void* az_stream_t;
struct AzRt {
  az_stream_t stream = nullptr;
  unsigned long long launches = 0;
};
static inline int rt_init(AzRt&, int) { return 0; }
static inline void rt_destroy(AzRt&) {}
static inline void* rt_alloc(size_t bytes) { return calloc(bytes ? bytes : 1, 1); }
static inline void rt_free(void* p) { free(p); }
static inline void rt_h2d(AzRt&, void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
static inline void rt_d2h(AzRt&, void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
static inline void rt_d2h_async(AzRt&, void* dst, const void* src, size_t n) { memcpy(dst, src, n); }
static inline void rt_d2d(AzRt&, void* dst, const void* src, size_t n) { memmove(dst, src, n); }
static inline void rt_zero(AzRt&, void* dst, size_t n) { memset(dst, 0, n); }
static inline int rt_sync(AzRt&) { return 0; }
#else
#include <cuda_runtime.h>
typedef cudaStream_t az_stream_t;
struct AzRt {
  az_stream_t stream = nullptr;
  int device = 0;
  unsigned long long launches = 0;
};
#define AZ_CUDA_OK(expr)                                                                       \
  do {                                                                                         \
    cudaError_t _e = (expr);                                                                   \
    if (_e != cudaSuccess)                                                                     \
      return az_fail(-6, std::string("CUDA error: ") + cudaGetErrorString(_e) + " at " #expr); \
  } while (0)
static inline int rt_init(AzRt& rt, int device) {
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n <= 0)
    return az_fail(-6, std::string("no CUDA device available: the engine has no CPU fallback (") + cudaGetErrorString(e) + ")");
  AZ_CUDA_OK(cudaSetDevice(device));
  rt.device = device;
  AZ_CUDA_OK(cudaStreamCreateWithFlags(&rt.stream, cudaStreamNonBlocking));
  return 0;
}
static inline void rt_destroy(AzRt& rt) {
  if (rt.stream) cudaStreamDestroy(rt.stream);
  rt.stream = nullptr;
}
static inline void* rt_alloc(size_t bytes) {
  void* p = nullptr;
  if (cudaMalloc(&p, bytes ? bytes : 1) != cudaSuccess) {
    cudaGetLastError();  // an allocation failure is reported by the caller, it must not surface later as a launch error
    return nullptr;
  }
  // cudaMemset on device memory is asynchronous and runs on the legacy default stream, which the engine's non-blocking
  // stream does not wait for: without this synchronisation the zero fill can land AFTER the first copy / kernel that
  // uses the buffer (seen when several processes share the GPU: freshly uploaded weights came back as zeros)
  cudaMemset(p, 0, bytes ? bytes : 1);
  cudaStreamSynchronize(0);
  return p;
}
static inline void rt_free(void* p) { if (p) cudaFree(p); }
static inline void rt_h2d(AzRt& rt, void* dst, const void* src, size_t n) {
  cudaMemcpyAsync(dst, src, n, cudaMemcpyHostToDevice, rt.stream);
  cudaStreamSynchronize(rt.stream);
}
static inline void rt_d2h(AzRt& rt, void* dst, const void* src, size_t n) {
  cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, rt.stream);
  cudaStreamSynchronize(rt.stream);
}
// queued only: the caller synchronises the stream (rt_sync) before reading dst
static inline void rt_d2h_async(AzRt& rt, void* dst, const void* src, size_t n) {
  cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToHost, rt.stream);
}
static inline void rt_d2d(AzRt& rt, void* dst, const void* src, size_t n) {
  cudaMemcpyAsync(dst, src, n, cudaMemcpyDeviceToDevice, rt.stream);
}
static inline void rt_zero(AzRt& rt, void* dst, size_t n) { cudaMemsetAsync(dst, 0, n, rt.stream); }
static inline int rt_sync(AzRt& rt) {
  cudaError_t e = cudaStreamSynchronize(rt.stream);
  if (e == cudaSuccess) e = cudaGetLastError();
  if (e != cudaSuccess) return az_fail(-6, std::string("CUDA error: ") + cudaGetErrorString(e));
  return 0;
}
#endif
