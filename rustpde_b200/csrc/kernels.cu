// kernels.cu -- __global__ entry points and their launch wrappers:
//   * lane_vm_kernel : the lane-program interpreter (lane_vm.cuh)
//   * dgemm_dmma     : FP64 tensor-core GEMM (mma.sync m8n8k4) for the
//                      fast-diagonalisation contractions (fdma_tensor.rs:212-233)
//   * small elementwise / reduction kernels for the diagnostics
#include "kernels.h"
#include "lane_vm.cuh"

namespace rp {

// (512, 1): without the explicit min-blocks ptxas squeezed the whole call tree into 32 registers
__global__ void __launch_bounds__(512, 1) lane_vm_kernel(const Program* __restrict__ progs) { lane_vm_body(progs); }

// block tile 128 x BN x 16 (BN = 64: 4 x 2 warps of 32 x 32; BN = 56: 8 x 1 warps of 16 x 56), 3-stage cp.async
// pipeline; shared tiles: A [m][k] (row stride GLDK), B [k][n] (row stride BN + 4), both padded so that the DMMA
// fragment loads are conflict-free
enum { GBM = 128, GBK = 16, GLDK = GBK + 4, GSTAGES = 3 };
enum { GA_STAGE = GBM * GLDK };
template <int WN, int NJ>
struct GemmCfg {
  static constexpr int WM = 8 / WN;          // warps along m
  static constexpr int NI = GBM / (8 * WM);  // m8 fragments per warp
  static constexpr int BN = WN * NJ * 8, LDN = BN + 4;
  static constexpr int B_STAGE = GBK * LDN;
  static constexpr int SMEM = GSTAGES * (GA_STAGE + B_STAGE) * (int)sizeof(double);
};
template <int WN, int NJ>
__global__ void __launch_bounds__(256, 2) dgemm_dmma_kernel(GemmArgs g0, GemmArgs g1);

void init_kernels() {
#ifndef RP_EMU
  static unsigned long long done = 0;  // one bit per device
  int dev = 0;
  cudaGetDevice(&dev);
  if (done >> (dev & 63) & 1) return;
  RP_CUDA_CHECK(cudaFuncSetAttribute(lane_vm_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, RP_MAX_SMEM));
  RP_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<2, 4>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<2, 4>::SMEM));
  RP_CUDA_CHECK(cudaFuncSetAttribute(dgemm_dmma_kernel<1, 7>, cudaFuncAttributeMaxDynamicSharedMemorySize, GemmCfg<1, 7>::SMEM));
  done |= 1ull << (dev & 63);
#endif
}

void launch_lane_programs(const Program* d_progs, int nprogs, int nblocks, int nthreads, int smem_bytes, cudaStream_t s) {
  if (nblocks <= 0 || nprogs <= 0) return;
  init_kernels();
  if (smem_bytes > RP_MAX_SMEM) throw Error(RP_ERR_INTERNAL, "lane program needs too much shared memory");
  RP_LAUNCH(lane_vm_kernel, dim3(nblocks, nprogs), dim3(nthreads), (size_t)smem_bytes, s, d_progs);
}

// ===========================================================================
// FP64 DMMA GEMM:  C[m, n] = sum_k A[m, k] * B[k, n]      (row-major)
//   B element (k, n) at B[(b_r0 + k*b_rs)*ldb + n], C likewise with c_r0/c_rs
//   so the even/odd parity-split solves can address interleaved rows.
// Block tile 128 x 64 x 16, 8 warps (4 x 2), warp tile 32 x 32 (16 m8n8k4 DMMAs per
// k-step of 4), operands staged by cp.async (16-byte chunks, zero-filled at the
// edges) through a 3-stage ring, two blocks per SM.  blockIdx.z selects one of two
// independent products (the even and the odd half of a parity-split solve).
// ===========================================================================
RP_DEV void dmma(double& d0, double& d1, double a, double b) {
#ifdef RP_EMU
  cuemu::mma_m8n8k4(d0, d1, a, b, d0, d1);
#else
  asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};"
               : "+d"(d0), "+d"(d1)
               : "d"(a), "d"(b));
#endif
}
// 16-byte asynchronous copy global -> shared; only the first `bytes` (0, 8 or 16) are read, the rest is zero
RP_DEV void cp_async16(double* smem_dst, const double* gsrc, int bytes) {
#ifdef RP_EMU
  smem_dst[0] = bytes >= 8 ? gsrc[0] : 0.0;
  smem_dst[1] = bytes >= 16 ? gsrc[1] : 0.0;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
#endif
}
RP_DEV void cp_async_commit() {
#ifndef RP_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>
RP_DEV void cp_async_wait() {
#ifndef RP_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

template <int WN, int NJ>
__global__ void __launch_bounds__(256, 2) dgemm_dmma_kernel(GemmArgs g0, GemmArgs g1) {
  typedef GemmCfg<WN, NJ> C;
  constexpr int NI = C::NI, BN = C::BN, LDN = C::LDN;
  const GemmArgs& g = blockIdx.z ? g1 : g0;
  if ((int)(blockIdx.y * GBM) >= g.M) return;
  RP_DYN_SMEM(double, sm);
  double* As = sm;                        // [GSTAGES][GBM][GLDK]
  double* Bs = sm + GSTAGES * GA_STAGE;   // [GSTAGES][GBK][LDN]
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int m0 = blockIdx.y * GBM, n0 = blockIdx.x * BN;
  const int wm = (warp / WN) * (8 * NI), wn = (warp % WN) * (8 * NJ);
  const int lr = lane >> 2, lc = lane & 3;
  // number of 8-column blocks of this warp that hold valid columns (ragged last tile)
  int jmax = (g.N - n0 - wn + 7) >> 3;
  jmax = jmax < 0 ? 0 : (jmax > NJ ? NJ : jmax);
  double acc[NI][NJ][2];
#pragma unroll
  for (int i = 0; i < NI; ++i)
#pragma unroll
    for (int j = 0; j < NJ; ++j) acc[i][j][0] = acc[i][j][1] = 0.0;
  const int nk = (g.K + GBK - 1) / GBK;
  auto issue = [&](int kt) {
    if (kt < nk) {
      const int st = kt % GSTAGES, k0 = kt * GBK;
      double* a = As + st * GA_STAGE;
      double* b = Bs + st * C::B_STAGE;
#pragma unroll
      for (int q = 0; q < 4; ++q) {  // A: 128 rows x 8 chunks
        const int ch = tid + q * 256, row = ch >> 3, kc = (ch & 7) * 2;
        const int gm = m0 + row, gk = k0 + kc;
        int bytes = (gm < g.M) ? (g.K - gk) * 8 : 0;
        bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
        const double* src = bytes ? g.A + (size_t)gm * g.lda + gk : g.A;
        cp_async16(a + row * GLDK + kc, src, bytes);
      }
      constexpr int BCH = BN / 2, BTOT = GBK * BCH;  // B: 16 rows x BN / 2 chunks
#pragma unroll
      for (int q = 0; q < (BTOT + 255) / 256; ++q) {
        const int ch = tid + q * 256;
        if (BTOT % 256 == 0 || ch < BTOT) {
          const int row = ch / BCH, nc = (ch % BCH) * 2;
          const int gk = k0 + row, gn = n0 + nc;
          int bytes = (gk < g.K) ? (g.N - gn) * 8 : 0;
          bytes = bytes < 0 ? 0 : (bytes > 16 ? 16 : bytes);
          const double* src = bytes ? g.B + (size_t)(g.b_r0 + (long long)gk * g.b_rs) * g.ldb + gn : g.B;
          cp_async16(b + row * LDN + nc, src, bytes);
        }
      }
    }
    cp_async_commit();
  };
#pragma unroll
  for (int s = 0; s < GSTAGES - 1; ++s) issue(s);
  for (int kt = 0; kt < nk; ++kt) {
    cp_async_wait<GSTAGES - 2>();
    __syncthreads();
    issue(kt + GSTAGES - 1);
    const double* a = As + (kt % GSTAGES) * GA_STAGE;
    const double* b = Bs + (kt % GSTAGES) * C::B_STAGE;
    if (jmax > 0) {
#pragma unroll
      for (int kk = 0; kk < GBK; kk += 4) {
        double af[NI], bf[NJ];
#pragma unroll
        for (int i = 0; i < NI; ++i) af[i] = a[(wm + 8 * i + lr) * GLDK + kk + lc];
#pragma unroll
        for (int j = 0; j < NJ; ++j) bf[j] = b[(kk + lc) * LDN + wn + 8 * j + lr];
#pragma unroll
        for (int j = 0; j < NJ; ++j)
          if (j < jmax) {
#pragma unroll
            for (int i = 0; i < NI; ++i) dmma(acc[i][j][0], acc[i][j][1], af[i], bf[j]);
          }
      }
    }
  }
#pragma unroll
  for (int i = 0; i < NI; ++i) {
    const int gm = m0 + wm + 8 * i + lr;
    if (gm >= g.M) continue;
    double* crow = g.C + (size_t)(g.c_r0 + (long long)gm * g.c_rs) * g.ldc;
#pragma unroll
    for (int j = 0; j < NJ; ++j) {
      const int gn = n0 + wn + 8 * j + 2 * lc;
      if (gn < g.N) crow[gn] = acc[i][j][0];
      if (gn + 1 < g.N) crow[gn + 1] = acc[i][j][1];
    }
  }
}

// Tile width: 64 columns, or 56 when that needs less time in units of (waves of 2 blocks per SM) x (tile width) --
// e.g. the two 1023 x 1023 x 2047 products of the parity-split solve at 2048 x 2049 are 2 x 8 x 37 = 592 tiles of
// 128 x 56 = exactly two waves on 148 SMs, against 512 tiles = 1.73 -> 2 waves of 128 x 64.
static int gemm_sm_count() {
#ifndef RP_EMU
  static int n = 0;
  if (!n) {
    int dev = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&n, cudaDevAttrMultiProcessorCount, dev);
  }
  return n;
#else
  return 1;
#endif
}
static int gemm_tile_width(int M, int N, int nz) {
  static const char* e = getenv("RUSTPDE_B200_GEMM_BN");
  if (e) return atoi(e) == 56 ? 56 : 64;
  const long long slots = 2LL * gemm_sm_count();
  const long long mt = (M + GBM - 1) / GBM;
  const long long t64 = mt * ((N + 63) / 64) * nz, t56 = mt * ((N + 55) / 56) * nz;
  const long long c64 = ((t64 + slots - 1) / slots) * 64, c56 = ((t56 + slots - 1) / slots) * 56;
  // (a grid that fits one wave keeps the 64-wide tile: its blocks do not all share an SM, and the 4 x 2 warp layout
  // needs fewer fragment loads per DMMA)
  return (t64 > slots && c56 * 21 < c64 * 20) ? 56 : 64;
}

static const int kSmem56 = GemmCfg<1, 7>::SMEM, kSmem64 = GemmCfg<2, 4>::SMEM;
void launch_dgemm(const GemmArgs& g, cudaStream_t s) {
  if (g.M <= 0 || g.N <= 0) return;
  init_kernels();
  const int bn = gemm_tile_width(g.M, g.N, 1);
  dim3 grid((g.N + bn - 1) / bn, (g.M + GBM - 1) / GBM);
  if (bn == 56)
    RP_LAUNCH((dgemm_dmma_kernel<1, 7>), grid, dim3(256), (size_t)kSmem56, s, g, g);
  else
    RP_LAUNCH((dgemm_dmma_kernel<2, 4>), grid, dim3(256), (size_t)kSmem64, s, g, g);
}
void launch_dgemm2(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t s) {
  if (g0.N <= 0 || g0.M <= 0) return;
  init_kernels();
  const int mmax = std::max(g0.M, g1.M);
  const int bn = gemm_tile_width(mmax, g0.N, 2);
  dim3 grid((g0.N + bn - 1) / bn, (mmax + GBM - 1) / GBM, 2);
  if (bn == 56)
    RP_LAUNCH((dgemm_dmma_kernel<1, 7>), grid, dim3(256), (size_t)kSmem56, s, g0, g1);
  else
    RP_LAUNCH((dgemm_dmma_kernel<2, 4>), grid, dim3(256), (size_t)kSmem64, s, g0, g1);
}

// ===========================================================================
// small helpers
// ===========================================================================
__global__ void zero_elems_kernel(double* p, int count) {
  const int i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < count) p[i] = 0.0;
}
void launch_zero_elems(double* p, int count, cudaStream_t s) {
  RP_LAUNCH(zero_elems_kernel, dim3(1), dim3(32), (size_t)0, s, p, count);
}

RP_DEV double block_sum(double v, double* red) {
  for (int o = 16; o > 0; o >>= 1) v += __shfl_down_sync(0xffffffffu, v, o);
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  if (lane == 0) red[w] = v;
  __syncthreads();
  double tot = 0.0;
  if (threadIdx.x == 0)
    for (int i = 0; i < (int)((blockDim.x + 31) >> 5); ++i) tot += red[i];
  __syncthreads();
  return tot;  // valid on thread 0
}

// out[0] = sum_ij wx[i] wy[j] g(a, b)[i, j];  mode selects g
//   0: a    1: sqrt(a^2+b^2)    2: 0.5 (a^2+b^2)    3: a^2 (sum of squares, weights ignored)
// Two stages with a fixed summation order (bitwise reproducible between runs, unlike atomicAdd): every block
// writes its partial sum to scratch[blockIdx.x], a single block adds the partials in index order.
enum { WSUM_MAX_BLOCKS = 592 };
__global__ void __launch_bounds__(256) wsum_kernel(const double* a, const double* b, long long ld, int rows, int cols,
                                                    const double* wx, const double* wy, int mode, double* scratch) {
  __shared__ double red[32];
  double acc = 0.0;
  const long long total = (long long)rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / cols), j = (int)(idx % cols);
    const double va = a[(size_t)i * ld + j];
    double gval;
    if (mode == 0)
      gval = va;
    else if (mode == 3)
      gval = va * va;
    else {
      const double vb = b[(size_t)i * ld + j];
      const double e = va * va + vb * vb;
      gval = (mode == 1) ? sqrt(e) : 0.5 * e;
    }
    const double w = (mode == 3) ? 1.0 : wx[i] * wy[j];
    acc = fma(w, gval, acc);
  }
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) scratch[blockIdx.x] = tot;
}
__global__ void __launch_bounds__(256) wsum_final_kernel(const double* scratch, int n, double* out) {
  __shared__ double red[32];
  double acc = 0.0;
  for (int i = threadIdx.x; i < n; i += blockDim.x) acc += scratch[i];
  const double tot = block_sum(acc, red);
  if (threadIdx.x == 0) out[0] = tot;
}
// `out` must hold 1 + WSUM_MAX_BLOCKS doubles: out[0] receives the sum, out[1..] is scratch
void launch_wsum(const double* a, const double* b, long long ld, int rows, int cols, const double* wx, const double* wy,
                 int mode, double* out, cudaStream_t s) {
  long long total = (long long)rows * cols;
  int blocks = (int)std::min<long long>((total + 255) / 256, (long long)WSUM_MAX_BLOCKS);
  if (blocks < 1) blocks = 1;
  RP_LAUNCH(wsum_kernel, dim3(blocks), dim3(256), (size_t)0, s, a, b, ld, rows, cols, wx, wy, mode, out + 1);
  RP_LAUNCH(wsum_final_kernel, dim3(1), dim3(256), (size_t)0, s, (const double*)(out + 1), blocks, out);
}

// out[j] = sum_i w[i] a[i][j]  (average_axis along axis 0, average.rs:25-33).  Stage 1: block (bx, by) sums the rows
// i = by, by + gridDim.y, ... of 256 columns into scratch[by][j]; stage 2 adds the gridDim.y partials in order.
enum { AVG0_ROWPARTS = 32 };
__global__ void __launch_bounds__(256) avg_axis0_kernel(const double* a, long long ld, int rows, int cols, const double* w,
                                                         double* scratch) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double acc = 0.0;
  for (int i = blockIdx.y; i < rows; i += gridDim.y) acc = fma(w[i], a[(size_t)i * ld + j], acc);
  scratch[(size_t)blockIdx.y * cols + j] = acc;
}
__global__ void __launch_bounds__(256) avg_axis0_final_kernel(const double* scratch, int parts, int cols, double* out) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  if (j >= cols) return;
  double acc = 0.0;
  for (int p = 0; p < parts; ++p) acc += scratch[(size_t)p * cols + j];
  out[j] = acc;
}
// scratch: AVG0_ROWPARTS * cols doubles
void launch_avg_axis0(const double* a, long long ld, int rows, int cols, const double* w, double* scratch, double* out,
                      cudaStream_t s) {
  const int parts = std::min<int>(AVG0_ROWPARTS, rows);
  dim3 grid((cols + 255) / 256, parts);
  RP_LAUNCH(avg_axis0_kernel, grid, dim3(256), (size_t)0, s, a, ld, rows, cols, w, scratch);
  RP_LAUNCH(avg_axis0_final_kernel, dim3((cols + 255) / 256), dim3(256), (size_t)0, s, (const double*)scratch, parts, cols, out);
}

// out = (a + b*c*s1) * s0   elementwise on pitched arrays (eval_nuvol, functions.rs:60-72)
__global__ void combine_kernel(double* out, const double* a, const double* b, const double* c, long long ld, int rows,
                               int cols, double s0, double s1) {
  const long long total = (long long)rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / cols), j = (int)(idx % cols);
    const size_t o = (size_t)i * ld + j;
    double r = a[o];
    if (b) r += (c ? b[o] * c[o] : b[o]) * s1;
    out[o] = r * s0;
  }
}
void launch_combine(double* out, const double* a, const double* b, const double* c, long long ld, int rows, int cols,
                    double s0, double s1, cudaStream_t s) {
  long long total = (long long)rows * cols;
  int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
  if (blocks < 1) blocks = 1;
  RP_LAUNCH(combine_kernel, dim3(blocks), dim3(256), (size_t)0, s, out, a, b, c, ld, rows, cols, s0, s1);
}

// dealias (navier.rs:1022-1032): zero rows >= cut_row and columns >= cut_col of a spectral array (rc doubles per element)
__global__ void dealias_kernel(double* p, long long ld, int rows, int cols, int rc, int cut_row, int cut_col) {
  const long long total = (long long)rows * cols;
  for (long long idx = (long long)blockIdx.x * blockDim.x + threadIdx.x; idx < total; idx += (long long)gridDim.x * blockDim.x) {
    const int i = (int)(idx / cols), j = (int)(idx % cols);
    if (i >= cut_row || j >= cut_col)
      for (int k = 0; k < rc; ++k) p[((size_t)i * ld + j) * rc + k] = 0.0;
  }
}
void launch_dealias(double* p, long long ld, int rows, int cols, int rc, int cut_row, int cut_col, cudaStream_t s) {
  long long total = (long long)rows * cols;
  int blocks = (int)std::min<long long>((total + 255) / 256, 1184);
  if (blocks < 1) blocks = 1;
  RP_LAUNCH(dealias_kernel, dim3(blocks), dim3(256), (size_t)0, s, p, ld, rows, cols, rc, cut_row, cut_col);
}

// out[r][j] = lo[r] in[r][j] + di[r] in[r+2][j] + up[r] in[r+4][j], r < m = n - 2: the B2 preconditioner along x
// (matvec.rs:172-193) of a stand-alone Hholtz / Poisson solve; elementwise in j, so plain coalesced rows
__global__ void __launch_bounds__(256) b2x_kernel(const double* in, long long ldi, double* out, long long ldo, int n, int cols,
                                                   const double* lo, const double* di, const double* up) {
  const int j = blockIdx.x * blockDim.x + threadIdx.x;
  const int m = n - 2;
  if (j >= cols) return;
  for (int r = blockIdx.y; r < m; r += gridDim.y) {
    double v = lo[r] * in[(size_t)r * ldi + j] + di[r] * in[(size_t)(r + 2) * ldi + j];
    if (r + 4 < n) v = fma(up[r], in[(size_t)(r + 4) * ldi + j], v);
    out[(size_t)r * ldo + j] = v;
  }
}
void launch_b2x(const double* in, long long ldi, double* out, long long ldo, int n, int cols, const double* lo, const double* di,
                const double* up, cudaStream_t s) {
  dim3 grid((cols + 255) / 256, std::min(n - 2, 1024));
  RP_LAUNCH(b2x_kernel, grid, dim3(256), (size_t)0, s, in, ldi, out, ldo, n, cols, lo, di, up);
}

}  // namespace rp
