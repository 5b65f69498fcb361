// fast_y.cuh -- shared pieces of the y-pass kernels (fast_y.cu: confined and common passes,
// fast_p.cu: periodic passes on complex rows).
#pragma once
#include "fast.cuh"

namespace rp {
namespace fk {

template <int LOG2L, int LC_ = 2>
struct YCfg {
  static constexpr int LC = LC_, LR = 2 * LC_;  // complex / real lanes per tile
  static constexpr int N = 1 << LOG2L;
  static constexpr int n = N + 1;
  static constexpr int NTHR = (N * LC / 8) < 64 ? 64 : (N * LC / 8);
  // kernels without an FFT (banded sweeps only): the scans use at most 512 threads anyway, and a 1024-thread block
  // would cap them at 64 registers (the 8193-point instantiations spilled up to 2.4 KB per thread)
  static constexpr int NTHRS = NTHR > 512 ? 512 : NTHR;
  static constexpr int ROWS = N + 8;
  static constexpr int TILE = ROWS * LR;  // doubles per tile
  static constexpr int CL = chunk_len(n, NTHR, LC);
  static constexpr int SMEM1 = TILE * 8 + scan_threads(NTHR) * 56 + 512;      // one tile + scratch
  static constexpr int SMEM2 = (TILE + LR * ROWS) * 8 + scan_threads(NTHR) * 56 + 512;  // tile + pivot rows + scratch
  static constexpr int SMEMC = (TILE + LC * ROWS) * 8 + scan_threads(NTHR) * 56 + 512;  // same on complex rows
};

#define YK_SMEM(td, red)          \
  RP_DYN_SMEM(double, td);        \
  double* red = td + C::TILE

// tile(j, lane) = f(j, lane) for j < nfill.  Loads are issued in batches of FK_FILL_U
// per thread before the first store, so that enough global requests are in flight.
#define FK_FILL_U 8
template <int LC, int NTHR, class F>
FK_DEV void tile_fill(double* td, int nfill, F f) {
  constexpr int LR = 2 * LC;
  const int tot = nfill * LR;
  for (int it0 = threadIdx.x; it0 < tot; it0 += NTHR * FK_FILL_U) {
    double v[FK_FILL_U];
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) {
      const int it = min(it0 + u * NTHR, tot - 1);  // clamped: the loads stay unconditional
      v[u] = f(it / LR, it % LR);
    }
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) {
      const int it = it0 + u * NTHR;
      if (it < tot) td[didx<LC>(it / LR, it % LR)] = v[u];
    }
  }
}
// g(j, lane, value of natural element j) for j < nout
template <int LC, int NTHR, class G>
FK_DEV void tile_drain(const double* td, int sn, int nout, G g) {
  constexpr int LR = 2 * LC;
  for (int it = threadIdx.x; it < nout * LR; it += NTHR) {
    const int lane = it % LR, j = it / LR;
    g(j, lane, td[didx<LC>(rowof(sn, j), lane)]);
  }
}

// *addr(j, lane) -= tile(j, lane) for j < nout (addr returns nullptr for lanes outside the matrix).  The old values
// of a batch are all requested before the first store: a plain `-=` loop exposes one DRAM round trip per element.
template <int LC, int NTHR, class A>
FK_DEV void tile_drain_sub(const double* td, int sn, int nout, A addr) {
  constexpr int LR = 2 * LC, U = FK_FILL_U;
  const int tot = nout * LR;
  for (int it0 = threadIdx.x; it0 < tot; it0 += NTHR * U) {
    double* p[U];
    double old[U];
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * NTHR;
      p[u] = it < tot ? addr(it / LR, it % LR) : nullptr;
      old[u] = p[u] ? *p[u] : 0.0;
    }
#pragma unroll
    for (int u = 0; u < U; ++u) {
      const int it = it0 + u * NTHR;
      if (p[u]) *p[u] = old[u] - td[didx<LC>(rowof(sn, it / LR), it % LR)];
    }
  }
}

// composite -> ortho stencil applied while loading row r of `a` (m = n-2 columns):
// p_j = d_j c_j + l_{j-2} c_{j-2}   (composite_stencil.rs:207-229)
// All loads are unconditional (clamped indices, zero weights) so that the compiler
// can issue a whole batch of them before the first use.
FK_DEV double ld_stencil(const Mat& a, int r, int j, int m, const double* __restrict__ sd, const double* __restrict__ sl) {
  const int rr = min(r, a.rows - 1), j0 = min(j, m - 1), j2 = max(j - 2, 0);
  const double* row = a.p + (size_t)rr * a.ld;
  const double v0 = row[j0], v2 = row[j2];
  const bool ok = r < a.rows;
  const double d = (ok && j < m) ? __ldg(&sd[j0]) : 0.0;
  const double l = (ok && j >= 2) ? __ldg(&sl[j2]) : 0.0;
  return fma(l, v2, d * v0);
}
// plain element (r, j) of a, zero for rows outside
FK_DEV double ld_row(const Mat& a, int r, int j) {
  const double v = a.p[(size_t)min(r, a.rows - 1) * a.ld + j];
  return r < a.rows ? v : 0.0;
}


// Swept pivot reciprocals of NR rows -> ti[j * NR + l] (row fastest), as asynchronous 8-byte copies: every element is
// in flight at once and no register is held, so the DRAM round trip overlaps the tile fill that follows instead of
// costing one exposed latency per loop iteration (this loop and the rhs fill were 56 % of pk_hholtz's stall samples,
// profiles/r2_ncu_pk_hholtz_periodic2048.txt).  One commit group; the caller waits (cp_async_wait<0>) before its barrier.
template <int NR, int NTHR>
FK_DEV void stage_pivots(double* ti, const double* __restrict__ inv, long long inv_ld, int r0, int nrows, int m) {
  for (int it = threadIdx.x; it < m * NR; it += NTHR) {
    const int l = it / m, j = it - l * m;
    cp_async8(&ti[j * NR + l], inv + (size_t)min(r0 + l, nrows - 1) * inv_ld + j);
  }
  cp_async_commit();
}

// Per-mode banded solve (A + mu C) x = B2 g along the tile (fdma_tensor.rs:219-227, hholtz.rs:182-190):
// tile td holds g (n entries, natural layout), tile ti the swept pivot reciprocals 1/dia'_i of the lane
// (set-up data); everything else of the sweep is recomputed from the raw bands.  `mu` = lam + alpha of
// the calling thread's lane (threadIdx.x & 3).  Result: m = n - 2 entries in td.
// `ti` holds the reciprocals of the NR = (2 LC) >> CSHIFT rows of the tile, row fastest: ti[i * NR + row];
// lane l reads row l >> CSHIFT (CSHIFT = 0: 2 LC real rows; CSHIFT = 1: re/im lanes of LC complex rows share
// their row's values).
template <int LC, int NTHR, int CL, int ROWS, int CSHIFT>
FK_DEV void mode_solve(double* td, const double* ti, int n, const B2Tabs& B, const ModeTabs& M, double mu, double* red) {
  const int m = n - 2;
  constexpr int NR = (2 * LC) >> CSHIFT;
  (void)B;
  auto inv = [&](int i, int l) { return ti[i * NR + (l >> CSHIFT)]; };
  // the raw bands come from the chunk-major packed tables (one contiguous run per warp and step)
  const double2* PF = (const double2*)M.pf;  // slot: lo, di | up, a_low[i-2] | c_low[i-2], -
  const double2* PB = (const double2*)M.pb;  // slot: a_up1, c_up1 | a_up2, c_up2 | a_up2[i-2], c_up2[i-2] | a_low[i-2], c_low[i-2]
  // forward: x_i -= l_{i-2} x_{i-2},  l_j = low_j / dia'_j   (fdma.rs:104-107 on the swept system)
  scan1<LC, NTHR, CL, true>(
      m, red,
      [&](int i, int l, int s) {
        const double2 a = __ldg(&PF[3 * s]), b = __ldg(&PF[3 * s + 1]);
        return fma(a.x, td[didx<LC>(i, l)], fma(a.y, td[didx<LC>(i + 2, l)], (i + 4 < n) ? b.x * td[didx<LC>(i + 4, l)] : 0.0));
      },
      [&](int i, int l, int s) {
        if (i < 2) return 0.0;
        const double alow = __ldg(&PF[3 * s + 1]).y, clow = __ldg(&PF[3 * s + 2]).x;
        return -fma(mu, clow, alow) * inv(i - 2, l);
      },
      [&](int i, int l, double y) { td[didx<LC>(i, l)] = y; });
  // backward: x_i = (x_i - up1'_i x_{i+2} - up2_i x_{i+4}) / dia'_i   (fdma.rs:108-117)
  scan2<LC, NTHR, CL, false>(
      m, red, [&](int i, int l) { return inv(i, l) * td[didx<LC>(i, l)]; },
      [&](int i, int l, int s) {
        if (i >= m - 2) return 0.0;
        const double2 u1p = __ldg(&PB[4 * s]);
        double u1 = fma(mu, u1p.y, u1p.x);
        if (i >= 2) {
          const double2 u2m = __ldg(&PB[4 * s + 2]), lo = __ldg(&PB[4 * s + 3]);
          const double lw = fma(mu, lo.y, lo.x) * inv(i - 2, l);
          u1 = fma(-lw, fma(mu, u2m.y, u2m.x), u1);
        }
        return -u1 * inv(i, l);
      },
      [&](int i, int l, int s) {
        if (i >= m - 4) return 0.0;
        const double2 u2 = __ldg(&PB[4 * s + 1]);
        return -fma(mu, u2.y, u2.x) * inv(i, l);
      },
      [&](int i, int l, double y) { td[didx<LC>(i, l)] = y; });
}

// ---- launch helpers -----------------------------------------------------------------
static int log2_of(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

// (log2 lane length, complex lanes per tile): long lanes use the 2-real-lane tile
#define YK_SIZES(X) X(5, 2) X(6, 2) X(7, 1) X(8, 2) X(9, 2) X(10, 2) X(11, 2) X(12, 1) X(13, 1)

template <class K>
static void set_smem(K kern, int bytes) {
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
#else
  (void)kern;
  (void)bytes;
#endif
}

#define YK_LAUNCH(kern, two_tiles, nrows, ny, args, nby)                                              \
  do {                                                                                                \
    const int nby_ = (nby);                                                                           \
    const int l_ = log2_of((ny)-1);                                                                   \
    const int nrows_ = (nrows);                                                                       \
    bool ok_ = false;                                                                                 \
    YK_SIZES(YK_CASE_##kern)                                                                          \
    if (!ok_) throw Error(RP_ERR_INTERNAL, #kern ": unsupported lane length");                        \
  } while (0)

#define YK_CASE_BODY(kern, L, LCV, smem_sel, args) YK_CASE_BODY_T(kern, L, LCV, smem_sel, args, NTHR)
#define YK_CASE_BODY_T(kern, L, LCV, smem_sel, args, NT)                                     \
  if (l_ == L) {                                                                              \
    typedef YCfg<L, LCV> C;                                                                   \
    const int nb_ = ((nrows_) + C::LR - 1) / C::LR;                                           \
    const int sm_ = (smem_sel) == 2 ? C::SMEMC : ((smem_sel) == 1 ? C::SMEM2 : C::SMEM1);     \
    auto kp_ = kern<L, LCV>;                                                                  \
    static unsigned long long init_ = 0; /* one bit per device */                                                                \
    if (first_use_on_device(init_)) {                                                                             \
      set_smem(kp_, sm_);                                                                     \
    }                                                                                         \
    RP_LAUNCH(kp_, dim3(nb_, nby_), dim3(C::NT), (size_t)sm_, s, args);                       \
    ok_ = true;                                                                               \
  }


}  // namespace fk
}  // namespace rp
