// cuemu.h -- TEST INFRASTRUCTURE ONLY.
//
// A tiny single-process emulator of the CUDA execution model (grid of blocks,
// block of threads as cooperative fibers, __syncthreads, warp shuffles,
// dynamic shared memory, mma.sync m8n8k4 f64).  It exists because the build
// container has no GPU: compiling the *same* kernel sources with -DRP_EMU
// against this header lets the CPU test-suite check every kernel's index
// arithmetic and numerics against the oracle before GPU minutes are spent.
//
// It is never linked into librustpde_b200.so (the product): the emulated build
// is a separate shared object under tests/cuemu/_build/ that only tests load.
#pragma once
#include <algorithm>
#include <cmath>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <functional>
#include <vector>

#define __global__
#define __device__
#define __host__
#define __forceinline__ inline __attribute__((always_inline))
#define __launch_bounds__(...)
#define __noinline__ __attribute__((noinline))
#define __shared__ static
#define __align__(n) __attribute__((aligned(n)))

struct double2 {
  double x, y;
};
static inline double2 make_double2(double x, double y) { return double2{x, y}; }
struct uint3 {
  unsigned x, y, z;
};
struct dim3 {
  unsigned x, y, z;
  dim3(unsigned x_ = 1, unsigned y_ = 1, unsigned z_ = 1) : x(x_), y(y_), z(z_) {}
};

namespace cuemu {

extern "C" void cuemu_switch(void** old_sp, void* new_sp);

struct Fiber {
  void* sp = nullptr;
  char* stack = nullptr;
  bool done = true;
};

struct BlockState {
  std::vector<Fiber> fibers;
  void* sched_sp = nullptr;
  int cur = 0;
  int nthreads = 0;
  int alive = 0;
  int bar_count = 0;
  unsigned bar_gen = 0;
  std::vector<char> smem;
  // warp exchange buffers: [warp][lane][8 doubles]
  std::vector<double> xchg;
  const std::function<void()>* body = nullptr;
};

inline BlockState& bs() {
  static BlockState s;
  return s;
}
inline uint3& tidx() {
  static uint3 v;
  return v;
}
inline uint3& bidx() {
  static uint3 v;
  return v;
}
inline dim3& bdim() {
  static dim3 v;
  return v;
}
inline dim3& gdim() {
  static dim3 v;
  return v;
}

static const size_t kStack = 64 * 1024;

inline void yield() {
  BlockState& b = bs();
  cuemu_switch(&b.fibers[b.cur].sp, b.sched_sp);
}

inline void fiber_entry() {
  BlockState& b = bs();
  (*b.body)();
  b.fibers[b.cur].done = true;
  b.alive--;
  // a thread that exits no longer participates in barriers
  if (b.alive > 0 && b.bar_count == b.alive) {
    b.bar_count = 0;
    b.bar_gen++;
  }
  yield();
  abort();  // never resumed
}

inline void run_block(int nthreads, const std::function<void()>& body) {
  BlockState& b = bs();
  if ((int)b.fibers.size() < nthreads) {
    size_t old = b.fibers.size();
    b.fibers.resize(nthreads);
    for (size_t i = old; i < b.fibers.size(); ++i) b.fibers[i].stack = (char*)aligned_alloc(64, kStack);
  }
  b.nthreads = nthreads;
  b.alive = nthreads;
  b.bar_count = 0;
  b.bar_gen = 0;
  b.body = &body;
  b.xchg.assign((size_t)((nthreads + 31) / 32) * 32 * 8, 0.0);
  for (int t = 0; t < nthreads; ++t) {
    Fiber& f = b.fibers[t];
    f.done = false;
    uintptr_t top = ((uintptr_t)(f.stack + kStack)) & ~(uintptr_t)15;
    void** sp = (void**)(top - 8);  // fake return slot -> entry sees rsp % 16 == 8
    *sp = nullptr;
    *(--sp) = (void*)&fiber_entry;  // popped by ret
    for (int i = 0; i < 6; ++i) *(--sp) = nullptr;  // rbp rbx r12 r13 r14 r15
    f.sp = (void*)sp;
  }
  dim3 bd = bdim();
  while (b.alive > 0) {
    for (int t = 0; t < nthreads; ++t) {
      if (b.fibers[t].done) continue;
      b.cur = t;
      uint3& ti = tidx();
      ti.x = t % bd.x;
      ti.y = (t / bd.x) % bd.y;
      ti.z = t / (bd.x * bd.y);
      cuemu_switch(&b.sched_sp, b.fibers[t].sp);
    }
  }
}

inline void launch(dim3 grid, dim3 block, size_t smem, const std::function<void()>& body) {
  BlockState& b = bs();
  gdim() = grid;
  bdim() = block;
  int nthreads = (int)(block.x * block.y * block.z);
  if (b.smem.size() < smem + 64) b.smem.resize(smem + 64);
  for (unsigned z = 0; z < grid.z; ++z)
    for (unsigned y = 0; y < grid.y; ++y)
      for (unsigned x = 0; x < grid.x; ++x) {
        bidx() = uint3{x, y, z};
        memset(b.smem.data(), 0xff, smem);  // poison (NaN) so uninitialised reads show up
        run_block(nthreads, body);
      }
}

inline void* dyn_smem() {
  uintptr_t p = (uintptr_t)bs().smem.data();
  return (void*)((p + 63) & ~(uintptr_t)63);
}

inline void syncthreads() {
  BlockState& b = bs();
  unsigned gen = b.bar_gen;
  b.bar_count++;
  if (b.bar_count == b.alive) {
    b.bar_count = 0;
    b.bar_gen++;
    yield();  // keep round-robin fairness
    return;
  }
  while (b.bar_gen == gen) yield();
}

inline int lane_id() { return bs().cur & 31; }
inline int warp_id() { return bs().cur >> 5; }
inline double* xchg_slot(int lane) { return &bs().xchg[((size_t)warp_id() * 32 + lane) * 8]; }

// All lanes of a warp must execute the same shuffle sequence (true for
// well-formed *_sync code with a full mask).
template <typename T>
inline T shfl_idx(T v, int src) {
  static_assert(sizeof(T) <= 8, "shfl of <= 8 bytes");
  memcpy(xchg_slot(lane_id()), &v, sizeof(T));
  yield();
  T r;
  int wbase = warp_id() * 32;
  if (src < 0 || src > 31 || wbase + src >= bs().nthreads) src = lane_id();
  memcpy(&r, xchg_slot(src), sizeof(T));
  yield();
  return r;
}

// mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 : D = A(8x4) * B(4x8) + C
// fragment layout (PTX ISA): a: row = lane/4, col = lane%4
//                            b: row(k) = lane%4, col(n) = lane/4
//                            c/d: row = lane/4, cols = 2*(lane%4) + {0,1}
inline void mma_m8n8k4(double& d0, double& d1, double a, double b, double c0, double c1) {
  int lane = lane_id();
  double* s = xchg_slot(lane);
  s[0] = a;
  s[1] = b;
  yield();
  int row = lane >> 2, cq = (lane & 3) * 2;
  double acc0 = c0, acc1 = c1;
  for (int k = 0; k < 4; ++k) {
    double av = xchg_slot(row * 4 + k)[0];
    double b0 = xchg_slot((cq + 0) * 4 + k)[1];
    double b1 = xchg_slot((cq + 1) * 4 + k)[1];
    acc0 = std::fma(av, b0, acc0);
    acc1 = std::fma(av, b1, acc1);
  }
  yield();
  d0 = acc0;
  d1 = acc1;
}

}  // namespace cuemu

using std::max;
using std::min;
#define threadIdx (cuemu::tidx())
#define blockIdx (cuemu::bidx())
#define blockDim (cuemu::bdim())
#define gridDim (cuemu::gdim())

static inline void __syncthreads() { cuemu::syncthreads(); }
// warp barrier: in the round-robin fiber model one yield lets every other thread reach the same point
static inline void __syncwarp(unsigned = 0xffffffffu) { cuemu::yield(); }
template <typename T>
static inline T __shfl_sync(unsigned, T v, int src, int = 32) { return cuemu::shfl_idx(v, src); }
template <typename T>
static inline T __shfl_xor_sync(unsigned, T v, int m, int = 32) { return cuemu::shfl_idx(v, cuemu::lane_id() ^ m); }
template <typename T>
static inline T __shfl_down_sync(unsigned, T v, unsigned d, int = 32) {
  int s = cuemu::lane_id() + (int)d;
  return cuemu::shfl_idx(v, s > 31 ? cuemu::lane_id() : s);
}
template <typename T>
static inline T __shfl_up_sync(unsigned, T v, unsigned d, int = 32) {
  int s = cuemu::lane_id() - (int)d;
  return cuemu::shfl_idx(v, s < 0 ? cuemu::lane_id() : s);
}
template <typename T>
static inline T __ldg(const T* p) { return *p; }
static inline double atomicAdd(double* p, double v) {
  double o = *p;
  *p = o + v;
  return o;
}
static inline int atomicAdd(int* p, int v) {
  int o = *p;
  *p = o + v;
  return o;
}
static inline int __clz(int v) { return v ? __builtin_clz((unsigned)v) : 32; }
static inline double __dmul_rn(double a, double b) { return a * b; }
static inline double __dadd_rn(double a, double b) { return a + b; }
static inline double __fma_rn(double a, double b, double c) { return std::fma(a, b, c); }
