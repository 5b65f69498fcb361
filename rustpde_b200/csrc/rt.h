// rt.h -- thin runtime layer: device memory, streams, launches.
//
// Product build (nvcc): plain CUDA runtime.  The only other mode, RP_EMU, is
// the CPU *emulation of the CUDA kernels* used by the GPU-less test-suite
// (tests/cuemu); it is never compiled into librustpde_b200.so.
#pragma once
#include <cstddef>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>

#include "../../include/rustpde_b200.h"

#ifdef RP_EMU
#include "cuemu.h"
typedef int cudaStream_t;
#define RP_HD
#else
#include <cuda_runtime.h>
#define RP_HD __host__ __device__
#endif

namespace rp {

struct Error : std::runtime_error {
  int code;
  Error(int c, const std::string& m) : std::runtime_error(m), code(c) {}
};

// status codes: RP_OK, RP_ERR_* of the C ABI (include/rustpde_b200.h)

#ifndef RP_EMU
#define RP_CUDA_CHECK(x)                                                                      \
  do {                                                                                        \
    cudaError_t e_ = (x);                                                                     \
    if (e_ != cudaSuccess)                                                                    \
      throw rp::Error(RP_ERR_CUDA, std::string(#x) + ": " + cudaGetErrorString(e_));      \
  } while (0)
#endif

namespace rt {

inline void* dmalloc(size_t bytes) {
  if (bytes == 0) bytes = 16;
#ifdef RP_EMU
  void* p = aligned_alloc(256, (bytes + 255) & ~(size_t)255);
  if (!p) throw Error(RP_ERR_CUDA, "emu alloc failed");
  memset(p, 0, bytes);
  return p;
#else
  void* p = nullptr;
  RP_CUDA_CHECK(cudaMalloc(&p, bytes));
  RP_CUDA_CHECK(cudaMemset(p, 0, bytes));
  return p;
#endif
}
inline void dfree(void* p) {
  if (!p) return;
#ifdef RP_EMU
  free(p);
#else
  cudaFree(p);
#endif
}
inline void h2d(void* d, const void* h, size_t bytes, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  memcpy(d, h, bytes);
#else
  RP_CUDA_CHECK(cudaMemcpyAsync(d, h, bytes, cudaMemcpyHostToDevice, s));
#endif
}
inline void d2h(void* h, const void* d, size_t bytes, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  memcpy(h, d, bytes);
#else
  RP_CUDA_CHECK(cudaMemcpyAsync(h, d, bytes, cudaMemcpyDeviceToHost, s));
#endif
}
inline void d2d(void* dst, const void* src, size_t bytes, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  memmove(dst, src, bytes);
#else
  RP_CUDA_CHECK(cudaMemcpyAsync(dst, src, bytes, cudaMemcpyDeviceToDevice, s));
#endif
}
// pitched copies (row-major, widths in bytes)
inline void h2d_2d(void* d, size_t dpitch, const void* h, size_t hpitch, size_t width, size_t rows, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  for (size_t r = 0; r < rows; ++r) memcpy((char*)d + r * dpitch, (const char*)h + r * hpitch, width);
#else
  RP_CUDA_CHECK(cudaMemcpy2DAsync(d, dpitch, h, hpitch, width, rows, cudaMemcpyHostToDevice, s));
#endif
}
inline void d2h_2d(void* h, size_t hpitch, const void* d, size_t dpitch, size_t width, size_t rows, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  for (size_t r = 0; r < rows; ++r) memcpy((char*)h + r * hpitch, (const char*)d + r * dpitch, width);
#else
  RP_CUDA_CHECK(cudaMemcpy2DAsync(h, hpitch, d, dpitch, width, rows, cudaMemcpyDeviceToHost, s));
#endif
}
inline void dzero(void* d, size_t bytes, cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
  memset(d, 0, bytes);
#else
  RP_CUDA_CHECK(cudaMemsetAsync(d, 0, bytes, s));
#endif
}
inline void sync(cudaStream_t s) {
#ifdef RP_EMU
  (void)s;
#else
  RP_CUDA_CHECK(cudaStreamSynchronize(s));
#endif
}
inline void check_launch(const char* what) {
#ifdef RP_EMU
  (void)what;
#else
  cudaError_t e = cudaGetLastError();
  if (e != cudaSuccess) throw Error(RP_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
#endif
}

}  // namespace rt
}  // namespace rp

#ifdef RP_EMU
#define RP_LAUNCH(kern, grid, block, smem, stream, ...) \
  cuemu::launch((grid), (block), (smem), [&]() { kern(__VA_ARGS__); })
#define RP_DYN_SMEM(type, name) type* name = (type*)cuemu::dyn_smem()
#else
#define RP_LAUNCH(kern, grid, block, smem, stream, ...)       \
  do {                                                        \
    kern<<<(grid), (block), (smem), (stream)>>>(__VA_ARGS__); \
    rp::rt::check_launch(#kern);                              \
  } while (0)
#define RP_DYN_SMEM(type, name) \
  extern __shared__ __align__(16) unsigned char rp_dyn_smem_raw_[]; \
  type* name = (type*)rp_dyn_smem_raw_
#endif
