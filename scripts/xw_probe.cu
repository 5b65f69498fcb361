// xw_probe.cu -- feasibility probe: warp-serial column sweeps (one thread = one parity chain of one column, no
// chunking, no shared memory) for the x-direction banded solves.  Build: nvcc -O3 -gencode arch=compute_100a,code=sm_100a
// Timing only; the production kernels live in rustpde_b200/csrc/fast_xw.cu.
#include <cstdio>
#include <cstdlib>
#include <cmath>
#include <vector>
#include <cuda_runtime.h>

#define CK(x)                                                                 \
  do {                                                                        \
    cudaError_t e_ = (x);                                                     \
    if (e_ != cudaSuccess) {                                                  \
      printf("CUDA error %s at %d\n", cudaGetErrorString(e_), __LINE__);      \
      exit(1);                                                                \
    }                                                                         \
  } while (0)

struct Args {
  const double* in0[3];
  const double* in1[3];
  const double* in2[3];
  double* tmp[3];
  double* out[3];
  const double4* cf;   // per row: lo, di, up, fp
  const double4* cb;   // per row: bs, bp1, bp2, -
  int n, ncols, ld;
};

// MAP 0: warp = 16 columns x 2 parities; MAP 1: warp = 32 columns, thread = both parities (ILP 2)
template <int U, int NIN>
__global__ void __launch_bounds__(128) sweep16(Args a) {
  const int f = blockIdx.y;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int col = w * 16 + (lane & 15), p = lane >> 4;
  if (w * 16 >= a.ncols) return;
  const bool ok = col < a.ncols;
  const int cc = ok ? col : a.ncols - 1;
  const int n = a.n, ld = a.ld;
  const int M = (n - p + 1) >> 1;
  const double* __restrict__ i0 = a.in0[f] + cc;
  const double* __restrict__ i1 = a.in1[f] + cc;
  const double* __restrict__ i2 = a.in2[f] + cc;
  double* __restrict__ tmp = a.tmp[f] + cc;
  double* __restrict__ out = a.out[f] + cc;
  double y = 0.0;
  for (int t0 = 0; t0 < M; t0 += U) {
    double q[U], k[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = 2 * min(t0 + u, M - 1) + p;
      const double2 ca = __ldg((const double2*)&a.cf[i]), cb_ = __ldg((const double2*)&a.cf[i] + 1);
      const double4 c = make_double4(ca.x, ca.y, cb_.x, cb_.y);
      double v = c.x * i0[(size_t)i * ld];
      if (NIN > 1) v = fma(c.y, i1[(size_t)i * ld], v);
      if (NIN > 2) v = fma(c.z, i2[(size_t)i * ld], v);
      q[u] = v;
      k[u] = c.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (t0 + u < M) {
        y = fma(k[u], y, q[u]);
        if (ok) tmp[(size_t)(2 * (t0 + u) + p) * ld] = y;
      }
  }
  double z1 = 0.0, z2 = 0.0;
  for (int t0 = 0; t0 < M; t0 += U) {
    double q[U], k1[U], k2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = 2 * (M - 1 - min(t0 + u, M - 1)) + p;
      const double2 ca = __ldg((const double2*)&a.cb[i]), cb_ = __ldg((const double2*)&a.cb[i] + 1);
      const double4 c = make_double4(ca.x, ca.y, cb_.x, cb_.y);
      q[u] = c.x * tmp[(size_t)i * ld];
      k1[u] = c.y;
      k2[u] = c.z;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (t0 + u < M) {
        const double z = fma(k1[u], z1, fma(k2[u], z2, q[u]));
        z2 = z1, z1 = z;
        if (ok) out[(size_t)(2 * (M - 1 - (t0 + u)) + p) * ld] = z;
      }
  }
}


// y direction: chain along the contiguous axis; warp = 16 rows x 2 parities (adjacent lanes = the two parities of a row)
template <int U, int NIN>
__global__ void __launch_bounds__(128) sweepy(Args a, int nrows, int ny) {
  const int f = blockIdx.y;
  const int w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, lane = threadIdx.x & 31;
  const int row = w * 16 + (lane >> 1), p = lane & 1;
  if (w * 16 >= nrows) return;
  const bool ok = row < nrows;
  const int rr = ok ? row : nrows - 1;
  const int n = ny, ld = a.ld;
  const int M = (n - p + 1) >> 1;
  const double* __restrict__ i0 = a.in0[f] + (size_t)rr * ld;
  const double* __restrict__ i1 = a.in1[f] + (size_t)rr * ld;
  const double* __restrict__ i2 = a.in2[f] + (size_t)rr * ld;
  double* __restrict__ tmp = a.tmp[f] + (size_t)rr * ld;
  double* __restrict__ out = a.out[f] + (size_t)rr * ld;
  double y = 0.0;
  for (int t0 = 0; t0 < M; t0 += U) {
    double q[U], k[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = 2 * min(t0 + u, M - 1) + p;
      const double2 ca = __ldg((const double2*)&a.cf[i]), cb_ = __ldg((const double2*)&a.cf[i] + 1);
      const double4 c = make_double4(ca.x, ca.y, cb_.x, cb_.y);
      double v = c.x * i0[i];
      if (NIN > 1) v = fma(c.y, i1[i], v);
      if (NIN > 2) v = fma(c.z, i2[i], v);
      q[u] = v;
      k[u] = c.w;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (t0 + u < M) {
        y = fma(k[u], y, q[u]);
        if (ok) tmp[2 * (t0 + u) + p] = y;
      }
  }
  double z1 = 0.0, z2 = 0.0;
  for (int t0 = 0; t0 < M; t0 += U) {
    double q[U], k1[U], k2[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int i = 2 * (M - 1 - min(t0 + u, M - 1)) + p;
      const double2 ca = __ldg((const double2*)&a.cb[i]), cb_ = __ldg((const double2*)&a.cb[i] + 1);
      const double4 c = make_double4(ca.x, ca.y, cb_.x, cb_.y);
      q[u] = c.x * tmp[i];
      k1[u] = c.y;
      k2[u] = c.z;
    }
#pragma unroll
    for (int u = 0; u < U; ++u)
      if (t0 + u < M) {
        const double z = fma(k1[u], z1, fma(k2[u], z2, q[u]));
        z2 = z1, z1 = z;
        if (ok) out[2 * (M - 1 - (t0 + u)) + p] = z;
      }
  }
}

// ---- warp-serial B2 + Fdma sweeps with a cp.async ring (the production design) ----
__device__ __forceinline__ void cp16(void* dst, const double* src, int bytes) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(src), "r"(bytes) : "memory");
}
template <int N>
__device__ __forceinline__ void cpwait() { asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory"); }
__device__ __forceinline__ void cpcommit() { asm volatile("cp.async.commit_group;" ::: "memory"); }

template <int RB, int K, int WPB>
__global__ void __launch_bounds__(32 * WPB) adi_ring(Args a) {
  extern __shared__ double smem[];
  constexpr int D = K + 1;
  constexpr int SLOT = RB * 16 + RB * 4;  // data rows + coefficient rows (4 doubles per row)
  const int wib = threadIdx.x >> 5;
  double* ring = smem + wib * (D * SLOT);
  const int f = blockIdx.y, lane = threadIdx.x & 31, lc = lane & 15, p = lane >> 4;
  const int c0 = (blockIdx.x * WPB + wib) * 16;
  if (c0 >= a.ncols) return;
  const int col = c0 + lc;
  const bool ok = col < a.ncols;
  const int n = a.n, m = n - 2, ld = a.ld;
  const double* __restrict__ in = a.in0[f];
  double* __restrict__ tmp = a.tmp[f];
  double* __restrict__ out = a.out[f];
  const int crow = lane >> 3, ccol = c0 + 2 * (lane & 7);
  const int cbytes = max(0, min(16, (a.ncols - ccol) * 8));
  // coef: table of 4 doubles per row, rows [0, ncoef); coefficient row of data row r is r + cshift
  auto issue = [&](const double* src, int nrows, const double4* coef, int ncoef, int cshift, int b, int nbat) {
    if (b >= 0 && b < nbat) {
      double* slot = ring + (b % D) * SLOT;
#pragma unroll
      for (int k = 0; k < RB / 4; ++k) {
        const int r = b * RB + 4 * k + crow;
        const bool v = r < nrows && cbytes > 0;
        cp16(&slot[(4 * k + crow) * 16 + 2 * (lane & 7)], src + (v ? (size_t)r * ld + ccol : 0), v ? cbytes : 0);
      }
      // RB rows x 32 bytes of coefficients = RB * 2 chunks of 16 bytes
      if (lane < RB * 2) {
        const int r = b * RB + (lane >> 1) + cshift;
        const bool v = r >= 0 && r < ncoef;
        cp16(&slot[RB * 16 + lane * 2], (const double*)(coef + (v ? r : 0)) + 2 * (lane & 1), v ? 16 : 0);
      }
    }
    cpcommit();
  };
  {
    const int nbat = (n + 4 + RB - 1) / RB;  // two extra chain steps drain the window
    for (int b = 0; b < K; ++b) issue(in, n, a.cf, m, -4, b, nbat);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, y = 0.0;
    for (int b = 0; b < nbat; ++b) {
      issue(in, n, a.cf, m, -4, b + K, nbat);
      cpwait<K>();
      __syncwarp();
      const double* slot = ring + (b % D) * SLOT;
      double rin[RB / 2];
      double4 c[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        rin[u] = slot[(2 * u + p) * 16 + lc];
        c[u] = *(const double4*)&slot[RB * 16 + (2 * u + p) * 4];  // coefficients of row i = (this row) - 4
      }
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        r0 = r1, r1 = r2, r2 = rin[u];
        const int i = b * RB + 2 * u + p - 4;
        const double q = fma(c[u].x, r0, fma(c[u].y, r1, (i + 4 < n) ? c[u].z * r2 : 0.0));
        const bool v = i >= 0 && i < m;
        y = v ? fma(c[u].w, y, q) : y;
        if (v && ok) tmp[(size_t)i * ld + col] = y;
      }
      __syncwarp();
    }
  }
  __threadfence_block();
  __syncwarp();
  {
    const int nbat = (m + RB - 1) / RB;
    for (int b = 0; b < K; ++b) issue(tmp, m, a.cb, m, 0, nbat - 1 - b, nbat);
    double z1 = 0.0, z2 = 0.0;
    for (int bb = 0; bb < nbat; ++bb) {
      const int b = nbat - 1 - bb;
      issue(tmp, m, a.cb, m, 0, b - K, nbat);
      cpwait<K>();
      __syncwarp();
      const double* slot = ring + (b % D) * SLOT;
      double rin[RB / 2];
      double4 c[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        rin[u] = slot[(2 * u + p) * 16 + lc];
        c[u] = *(const double4*)&slot[RB * 16 + (2 * u + p) * 4];
      }
#pragma unroll
      for (int u = RB / 2 - 1; u >= 0; --u) {
        const int i = b * RB + 2 * u + p;
        const bool v = i < m;
        const double z = fma(c[u].y, z1, fma(c[u].z, z2, c[u].x * rin[u]));
        if (v) z2 = z1, z1 = z;
        if (v && ok) out[(size_t)i * ld + col] = z;
      }
      __syncwarp();
    }
  }
}
// ---- v2: pointer-increment staging, one warp barrier per batch (D = K + 2), LDS of batch b+1 under the chain of batch b
template <int RB, int K, int WPB>
__global__ void __launch_bounds__(32 * WPB) adi_ring2(Args a) {
  extern __shared__ double smem[];
  constexpr int D = K + 2;
  constexpr int SLOT = RB * 16 + RB * 4;
  constexpr int U = RB / 2;
  const int wib = threadIdx.x >> 5;
  double* ring = smem + wib * (D * SLOT);
  const int f = blockIdx.y, lane = threadIdx.x & 31, lc = lane & 15, p = lane >> 4;
  const int c0 = (blockIdx.x * WPB + wib) * 16;
  if (c0 >= a.ncols) return;
  const int col = c0 + lc;
  const bool ok = col < a.ncols;
  const int n = a.n, m = n - 2, ld = a.ld;
  double* __restrict__ tmp = a.tmp[f];
  double* __restrict__ out = a.out[f];
  const int crow = lane >> 3, ccol = c0 + 2 * (lane & 7);
  const int cbytes = max(0, min(16, (a.ncols - ccol) * 8));
  const size_t rstep = (size_t)4 * ld;  // 4 rows
  const unsigned sdst0 = (unsigned)__cvta_generic_to_shared(ring) + (crow * 16 + 2 * (lane & 7)) * 8;
  const unsigned sdstc = (unsigned)__cvta_generic_to_shared(ring) + (RB * 16 + lane * 2) * 8;
  // stage batch b (rows [b RB, b RB + RB) of src; coefficient rows shifted by cshift)
  auto issue = [&](const double* src, int nrows, const double* coef, int ncoef, int cshift, int b, int nbat) {
    if (b >= 0 && b < nbat) {
      const unsigned so = (unsigned)((b % D) * SLOT * 8);
      const int r0 = b * RB + crow;
      const double* g = src + (size_t)r0 * ld + ccol;
      if (r0 + RB - 4 < nrows && cbytes == 16) {
#pragma unroll
        for (int k = 0; k < RB / 4; ++k)
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(sdst0 + so + k * 4 * 16 * 8), "l"(g + k * rstep) : "memory");
      } else {
#pragma unroll
        for (int k = 0; k < RB / 4; ++k) {
          const bool v = r0 + 4 * k < nrows && cbytes > 0;
          asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdst0 + so + k * 4 * 16 * 8), "l"(v ? g + k * rstep : src), "r"(v ? cbytes : 0) : "memory");
        }
      }
      const int rc = b * RB + (lane >> 1) + cshift;
      const bool v = rc >= 0 && rc < ncoef;
      asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(sdstc + so), "l"(coef + (v ? (size_t)rc * 4 + 2 * (lane & 1) : 0)), "r"(v ? 16 : 0) : "memory");
    }
    cpcommit();
  };
  double rin[U];
  double4 c[U];
  auto lds = [&](int b) {
    const double* slot = ring + (b % D) * SLOT;
#pragma unroll
    for (int u = 0; u < U; ++u) {
      rin[u] = slot[(2 * u + p) * 16 + lc];
      c[u] = *(const double4*)&slot[RB * 16 + (2 * u + p) * 4];
    }
  };
  {
    const double* in = a.in0[f];
    const int nbat = (n + 4 + RB - 1) / RB;
    for (int b = 0; b < K; ++b) issue(in, n, (const double*)a.cf, m, -4, b, nbat);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, y = 0.0;
    double* tp = tmp + (ptrdiff_t)(p - 4) * ld + col;
    const size_t st2 = (size_t)2 * ld;
    issue(in, n, (const double*)a.cf, m, -4, K, nbat);
    cpwait<K>();
    __syncwarp();
    lds(0);
    for (int b = 0; b < nbat; ++b) {
      double q[U], kk[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {  // window products that do not depend on the chain
        const double ra = u == 0 ? r1 : (u == 1 ? r2 : rin[u - 2]);
        const double rb_ = u == 0 ? r2 : rin[u - 1];
        q[u] = fma(c[u].x, ra, fma(c[u].y, rb_, c[u].z * rin[u]));
        kk[u] = c[u].w;
      }
      r1 = rin[U - 2], r2 = rin[U - 1];
      // next batch: stage b + K + 1, wait for b + 1, read it while the chain below runs
      issue(in, n, (const double*)a.cf, m, -4, b + K + 1, nbat);
      cpwait<K>();
      __syncwarp();
      if (b + 1 < nbat) lds(b + 1);
      const int i0 = b * RB + p - 4;
      if (ok && i0 >= 0 && i0 + RB - 2 < m) {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          y = fma(kk[u], y, q[u]);
          tp[u * st2] = y;
        }
      } else {
#pragma unroll
        for (int u = 0; u < U; ++u) {
          y = fma(kk[u], y, q[u]);
          const int i = i0 + 2 * u;
          if (ok && i >= 0 && i < m) tp[u * st2] = y;
        }
      }
      tp += (size_t)RB * ld;
    }
  }
  __threadfence_block();
  __syncwarp();
  {
    const int nbat = (m + RB - 1) / RB;
    for (int b = 0; b < K; ++b) issue(tmp, m, (const double*)a.cb, m, 0, nbat - 1 - b, nbat);
    double z1 = 0.0, z2 = 0.0;
    double* op = out + (ptrdiff_t)((nbat - 1) * RB + p) * ld + col;
    const size_t st2 = (size_t)2 * ld;
    issue(tmp, m, (const double*)a.cb, m, 0, nbat - 1 - K, nbat);
    cpwait<K>();
    __syncwarp();
    lds(nbat - 1);
    for (int b = nbat - 1; b >= 0; --b) {
      double q[U], k1[U], k2[U];
#pragma unroll
      for (int u = 0; u < U; ++u) q[u] = c[u].x * rin[u], k1[u] = c[u].y, k2[u] = c[u].z;
      issue(tmp, m, (const double*)a.cb, m, 0, b - K - 1, nbat);
      cpwait<K>();
      __syncwarp();
      if (b >= 1) lds(b - 1);
      const int i0 = b * RB + p;
      if (ok && i0 + RB - 2 < m) {
#pragma unroll
        for (int u = U - 1; u >= 0; --u) {
          const double z = fma(k1[u], z1, fma(k2[u], z2, q[u]));
          z2 = z1, z1 = z;
          op[u * st2] = z;
        }
      } else {
#pragma unroll
        for (int u = U - 1; u >= 0; --u) {
          const double z = fma(k1[u], z1, fma(k2[u], z2, q[u]));
          z2 = z1, z1 = z;
          if (ok && i0 + 2 * u < m) op[u * st2] = z;
        }
      }
      op -= (size_t)RB * ld;
    }
  }
}
// the same arithmetic, plain loads (checker)
__global__ void adi_plain(Args a, double* out2) {
  const int col = blockIdx.x * blockDim.x + threadIdx.x, f = blockIdx.y;
  if (col >= a.ncols) return;
  const int n = a.n, m = n - 2, ld = a.ld;
  const double* in = a.in0[f] + col;
  double* tmp = a.tmp[f] + col;
  double* out = out2 + (size_t)f * n * ld + col;
  for (int p = 0; p < 2; ++p) {
    double y = 0.0;
    for (int i = p; i < m; i += 2) {
      const double4 c = a.cf[i];
      const double q = fma(c.x, in[(size_t)i * ld], fma(c.y, in[(size_t)(i + 2) * ld], (i + 4 < n) ? c.z * in[(size_t)(i + 4) * ld] : 0.0));
      y = fma(c.w, y, q);
      tmp[(size_t)i * ld] = y;
    }
    double z1 = 0.0, z2 = 0.0;
    const int last = ((m - 1 - p) / 2) * 2 + p;
    for (int i = last; i >= 0; i -= 2) {
      const double4 c = a.cb[i];
      const double z = fma(c.y, z1, fma(c.z, z2, c.x * tmp[(size_t)i * ld]));
      z2 = z1, z1 = z;
      out[(size_t)i * ld] = z;
    }
  }
}

__global__ void copyk(const double2* a, double2* b, size_t n) {
  for (size_t i = blockIdx.x * (size_t)blockDim.x + threadIdx.x; i < n; i += (size_t)gridDim.x * blockDim.x) b[i] = a[i];
}

template <class F>
static float timeit(F f, int reps = 20) {
  cudaEvent_t e0, e1;
  CK(cudaEventCreate(&e0));
  CK(cudaEventCreate(&e1));
  for (int i = 0; i < 3; ++i) f();
  CK(cudaDeviceSynchronize());
  CK(cudaEventRecord(e0));
  for (int i = 0; i < reps; ++i) f();
  CK(cudaEventRecord(e1));
  CK(cudaEventSynchronize(e1));
  float ms;
  CK(cudaEventElapsedTime(&ms, e0, e1));
  return ms / reps;
}

int main() {
  const int n = 2048, ncols = 2049, ld = 2056;
  const size_t elems = (size_t)n * ld;
  Args a;
  std::vector<double> h(elems);
  for (size_t i = 0; i < elems; ++i) h[i] = (double)((i * 2654435761u) % 1000) / 1000.0 - 0.5;
  double* buf[15];
  for (int k = 0; k < 15; ++k) {
    CK(cudaMalloc(&buf[k], elems * 8));
    CK(cudaMemcpy(buf[k], h.data(), elems * 8, cudaMemcpyHostToDevice));
  }
  for (int f = 0; f < 3; ++f) {
    a.in0[f] = buf[f];
    a.in1[f] = buf[3 + f];
    a.in2[f] = buf[6 + f];
    a.tmp[f] = buf[9 + f];
    a.out[f] = buf[12 + f];
  }
  std::vector<double4> cf(n), cb(n);
  for (int i = 0; i < n; ++i) {
    cf[i] = make_double4(0.3, 0.5, 0.2, 0.4);
    cb[i] = make_double4(0.7, 0.3, -0.1, 0.0);
  }
  double4 *dcf, *dcb;
  CK(cudaMalloc(&dcf, n * 32));
  CK(cudaMalloc(&dcb, n * 32));
  CK(cudaMemcpy(dcf, cf.data(), n * 32, cudaMemcpyHostToDevice));
  CK(cudaMemcpy(dcb, cb.data(), n * 32, cudaMemcpyHostToDevice));
  a.cf = dcf;
  a.cb = dcb;
  a.n = n;
  a.ncols = ncols;
  a.ld = ld;
  {
    float ms = timeit([&] { copyk<<<148 * 8, 512>>>((const double2*)buf[0], (double2*)buf[12], elems / 2); });
    printf("copy 1 field: %.4f ms  (%.1f GB/s)\n", ms, 2.0 * elems * 8 / ms * 1e-6);
  }
  const int nw = (ncols + 15) / 16;
#define RUN(U, NIN, BT)                                                                                   \
  {                                                                                                       \
    const int wpb = (BT) / 32;                                                                            \
    dim3 g((nw + wpb - 1) / wpb, 3);                                                                      \
    float ms = timeit([&] { sweep16<U, NIN><<<g, BT>>>(a); });                                            \
    const double bytes = 3.0 * elems * 8 * ((NIN) + 3);                                                   \
    printf("sweep16 U=%d NIN=%d block=%d: %.4f ms  (%.1f GB/s algorithmic)\n", U, NIN, BT, ms, bytes / ms * 1e-6); \
  }
  RUN(4, 3, 32)
  RUN(8, 3, 32)
  RUN(16, 3, 32)
  RUN(8, 1, 32)
  RUN(16, 1, 32)
  RUN(32, 1, 32)
  RUN(8, 3, 64)
  RUN(16, 3, 64)
  RUN(8, 3, 128)

#define RUNY(U, NIN, BT)                                                                                  \
  {                                                                                                       \
    const int nwy = (n + 15) / 16, wpb = (BT) / 32;                                                       \
    dim3 g((nwy + wpb - 1) / wpb, 3);                                                                     \
    float ms = timeit([&] { sweepy<U, NIN><<<g, BT>>>(a, n, ncols); });                                   \
    const double bytes = 3.0 * elems * 8 * ((NIN) + 3);                                                   \
    printf("sweepy  U=%d NIN=%d block=%d: %.4f ms  (%.1f GB/s algorithmic)\n", U, NIN, BT, ms, bytes / ms * 1e-6); \
  }
  RUNY(8, 3, 32)
  RUNY(16, 3, 32)
  RUNY(16, 1, 32)
  RUNY(32, 1, 32)
  CK(cudaDeviceSynchronize());

  double* out2;
  CK(cudaMalloc(&out2, 3 * elems * 8));
  CK(cudaMemset(out2, 0, 3 * elems * 8));
  adi_plain<<<dim3((ncols + 63) / 64, 3), 64>>>(a, out2);
  CK(cudaDeviceSynchronize());
  std::vector<double> ref(elems), got(elems);
#define RUNR(RB, K, WPB)                                                                                  \
  {                                                                                                       \
    const int smem = (K + 1) * (RB * 20) * 8 * WPB;                                                         \
    CK(cudaFuncSetAttribute(adi_ring<RB, K, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));    \
    dim3 g((nw + WPB - 1) / WPB, 3);                                                                      \
    for (int f = 0; f < 3; ++f) CK(cudaMemset(a.out[f], 0, elems * 8));                                   \
    float ms = timeit([&] { adi_ring<RB, K, WPB><<<g, 32 * WPB, smem>>>(a); });                           \
    double err = 0.0;                                                                                     \
    for (int f = 0; f < 3; ++f) {                                                                         \
      CK(cudaMemcpy(ref.data(), out2 + (size_t)f * elems, elems * 8, cudaMemcpyDeviceToHost));            \
      CK(cudaMemcpy(got.data(), a.out[f], elems * 8, cudaMemcpyDeviceToHost));                            \
      for (int i = 0; i < n - 2; ++i)                                                                     \
        for (int c = 0; c < ncols; ++c) err = fmax(err, fabs(ref[(size_t)i * ld + c] - got[(size_t)i * ld + c])); \
    }                                                                                                     \
    const double bytes = 3.0 * elems * 8 * 4;                                                             \
    printf("adi_ring RB=%d K=%d WPB=%d smem=%d: %.4f ms  (%.1f GB/s algorithmic)  maxdiff %.3e\n", RB, K, WPB, smem, ms, bytes / ms * 1e-6, err); \
  }
  RUNR(16, 2, 1)
  RUNR(16, 4, 1)
  RUNR(16, 8, 1)
  RUNR(16, 12, 1)
  RUNR(32, 4, 1)
  RUNR(32, 6, 1)
  RUNR(8, 8, 1)
  RUNR(16, 8, 2)
  RUNR(16, 8, 4)

#define RUNR2(RB, K, WPB)                                                                                 \
  {                                                                                                       \
    const int smem = (K + 2) * (RB * 20) * 8 * WPB;                                                       \
    CK(cudaFuncSetAttribute(adi_ring2<RB, K, WPB>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));   \
    dim3 g((nw + WPB - 1) / WPB, 3);                                                                      \
    for (int f = 0; f < 3; ++f) CK(cudaMemset(a.out[f], 0, elems * 8));                                   \
    float ms = timeit([&] { adi_ring2<RB, K, WPB><<<g, 32 * WPB, smem>>>(a); });                          \
    double err = 0.0;                                                                                     \
    for (int f = 0; f < 3; ++f) {                                                                         \
      CK(cudaMemcpy(ref.data(), out2 + (size_t)f * elems, elems * 8, cudaMemcpyDeviceToHost));            \
      CK(cudaMemcpy(got.data(), a.out[f], elems * 8, cudaMemcpyDeviceToHost));                            \
      for (int i = 0; i < n - 2; ++i)                                                                     \
        for (int c = 0; c < ncols; ++c) err = fmax(err, fabs(ref[(size_t)i * ld + c] - got[(size_t)i * ld + c])); \
    }                                                                                                     \
    const double bytes = 3.0 * elems * 8 * 4;                                                             \
    printf("adi_ring2 RB=%d K=%d WPB=%d smem=%d: %.4f ms  (%.1f GB/s algorithmic)  maxdiff %.3e\n", RB, K, WPB, smem, ms, bytes / ms * 1e-6, err); \
  }
  RUNR2(16, 2, 1)
  RUNR2(16, 4, 1)
  RUNR2(16, 6, 1)
  RUNR2(16, 10, 1)
  RUNR2(16, 6, 2)
  printf("done\n");
  return 0;
}
