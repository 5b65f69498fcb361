// fast_y.cu -- specialised y-pass kernels of Navier2D::update (lanes contiguous
// in memory; a block owns 4 adjacent rows).  See fast.cuh for the tile model.
//
//   yk_backward : B_y S_y (composite.rs:480-506) of the x-backward results, value
//                 and d/dy (ortho.rs:107-125) -> physical space
//   yk_conv     : u . grad(f) products (conv_term.rs:41) + forward DCT-y + dealias-y
//   yk_adi      : y half of HholtzAdi (hholtz_adi.rs:113,129)
//   yk_mode     : per-mode banded solve of FdmaTensor (fdma_tensor.rs:219-227)
//   yk_project  : u -= from_ortho(grad phi), y part (navier.rs:683-695)
//   yk_pres     : pressure update (navier.rs:717-721) + d/dy p for the next step
#include "fast.cuh"

namespace rp {
namespace fk {

template <int LOG2L>
struct YCfg {
  static constexpr int N = 1 << LOG2L;
  static constexpr int n = N + 1;
  static constexpr int NTHR = (N / 4) < 64 ? 64 : (N / 4);
  static constexpr int ROWS = N + 4;
  static constexpr int CL = chunk_len(n, NTHR);
  static constexpr int SMEM1 = ROWS * 32 + scan_threads(NTHR) * 56 + 512;      // one tile + scratch
  static constexpr int SMEM2 = 2 * ROWS * 32 + scan_threads(NTHR) * 56 + 512;  // two tiles + scratch
};

#define YK_SMEM(td, red)          \
  RP_DYN_SMEM(double, td);        \
  double* red = td + C::ROWS * 4

// tile(j, lane) = f(j, lane) for j < nfill.  Loads are issued in batches of FK_FILL_U
// per thread before the first store, so that enough global requests are in flight.
#define FK_FILL_U 8
template <int NTHR, class F>
FK_DEV void tile_fill(double* td, int nfill, F f) {
  const int tot = nfill * 4;
  for (int it0 = threadIdx.x; it0 < tot; it0 += NTHR * FK_FILL_U) {
    double v[FK_FILL_U];
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) {
      const int it = min(it0 + u * NTHR, tot - 1);  // clamped: the loads stay unconditional
      v[u] = f(it >> 2, it & 3);
    }
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) {
      const int it = it0 + u * NTHR;
      if (it < tot) td[didx(it >> 2, it & 3)] = v[u];
    }
  }
}
// g(j, lane, value of natural element j) for j < nout
template <int NTHR, class G>
FK_DEV void tile_drain(const double* td, int sn, int nout, G g) {
  for (int it = threadIdx.x; it < nout * 4; it += NTHR) {
    const int lane = it & 3, j = it >> 2;
    g(j, lane, td[didx(rowof(sn, j), lane)]);
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

// ---------------------------------------------------------------------------------
template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 2) yk_backward(YBackwardArgs3 a3) {
  typedef YCfg<LOG2L> C;
  const YBackwardArgs& a = a3.a[blockIdx.y];
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, m = n - 2, N = C::N;
  auto fill = [&](const Mat& s) {
    tile_fill<C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(s, r0 + l, j, m, a.sd, a.sl); });
    __syncthreads();
  };
  auto drain = [&](const Mat& o) {
    tile_drain<C::NTHR>(td, N, n, [&](int j, int l, double v) {
      if (r0 + l < o.rows) o.p[(size_t)(r0 + l) * o.ld + j] = v;
    });
  };
  fill(a.a);
  if (a.val.p) {
    dct_pow2<LOG2L, C::NTHR, true>(td, a.t, red);
    drain(a.val);
    __syncthreads();
    fill(a.a);
  }
  cheb_diff<C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  dct_pow2<LOG2L, C::NTHR, true>(td, a.t, red);
  drain(a.dy);
  __syncthreads();
  fill(a.adx);
  dct_pow2<LOG2L, C::NTHR, true>(td, a.t, red);
  drain(a.dx);
}

template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 2) yk_conv(YConvArgs3 a3) {
  typedef YCfg<LOG2L> C;
  const YConvArgs& a = a3.a[blockIdx.y];
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, N = C::N;
  const bool has_bc = a.bcx.p != nullptr;
  tile_fill<C::NTHR>(td, n, [&](int j, int l) {
    const int r = min(r0 + l, a.u.rows - 1);
    double gx = a.du.p[(size_t)r * a.du.ld + j], gy = a.dv.p[(size_t)r * a.dv.ld + j];
    const double u = a.u.p[(size_t)r * a.u.ld + j], v = a.v.p[(size_t)r * a.v.ld + j];
    if (has_bc) {
      gx += a.bcx.p[(size_t)r * a.bcx.ld + j];
      gy += a.bcy.p[(size_t)r * a.bcy.ld + j];
    }
    return (r0 + l < a.u.rows) ? fma(u, gx, v * gy) : 0.0;
  });
  __syncthreads();
  dct_pow2<LOG2L, C::NTHR, false>(td, a.t, red);
  tile_drain<C::NTHR>(td, N, n, [&](int j, int l, double v) {
    if (r0 + l < a.out.rows) a.out.p[(size_t)(r0 + l) * a.out.ld + j] = (j < a.cut) ? v : 0.0;
  });
}

template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 2) yk_adi(YAdiArgs3 a3) {
  typedef YCfg<LOG2L> C;
  const YAdiArgs& a = a3.a[blockIdx.y];
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, m = n - 2;
  tile_fill<C::NTHR>(td, n, [&](int j, int l) { return ld_row(a.w, r0 + l, j); });
  __syncthreads();
  b2_fdma<C::NTHR, C::CL>(td, -1, n, a.b2, a.f, red);
  tile_drain<C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.out.rows) a.out.p[(size_t)(r0 + l) * a.out.ld + j] = v;
  });
  if (a.mode == 0) return;
  // ortho coefficients p_j = d_j x_j + l_{j-2} x_{j-2} of the new velocity component
  auto pj = [&](int j, int l) {
    double v = 0.0;
    if (j < m) v = __ldg(&a.sd[j]) * td[didx(j, l)];
    if (j >= 2) v = fma(__ldg(&a.sl[j - 2]), td[didx(j - 2, l)], v);
    return v;
  };
  if (a.mode == 1) {
    for (int it = threadIdx.x; it < n * 4; it += C::NTHR) {
      const int l = it & 3, j = it >> 2;
      if (r0 + l < a.aux.rows) a.aux.p[(size_t)(r0 + l) * a.aux.ld + j] = pj(j, l);
    }
    return;
  }
  __syncthreads();
  scan1<C::NTHR, C::CL, false>(
      n, red, [&](int i, int l) { return (2.0 * (double)i * a.isy) * pj(i, l); }, [](int, int) { return 1.0; },
      [&](int i, int l, double y) {
        if (i >= 1) td[didx(i - 1, l)] = (i == 1) ? 0.5 * y : y;
        if (i == n - 1) td[didx(n - 1, l)] = 0.0;
      });
  tile_drain<C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.aux.rows) a.aux.p[(size_t)(r0 + l) * a.aux.ld + j] = v;
  });
}

template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 1) yk_mode(YModeArgs a) {
  typedef YCfg<LOG2L> C;
  YK_SMEM(td, red0);
  double* ti = td + C::ROWS * 4;
  double* red = ti + C::ROWS * 4;
  (void)red0;
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, m = n - 2;
  tile_fill<C::NTHR>(td, n, [&](int j, int l) { return ld_row(a.g, r0 + l, j); });
  tile_fill<C::NTHR>(ti, m, [&](int j, int l) { return a.m.inv[(size_t)min(r0 + l, a.g.rows - 1) * a.m.inv_ld + j]; });
  __syncthreads();
  const ModeTabs& M = a.m;
  const B2Tabs& B = a.b2;
  const int myl = threadIdx.x & 3;
  const double mu = __ldg(&M.lam[min(r0 + myl, a.g.rows - 1)]) + M.alpha;
  // forward: x_i -= l_{i-2} x_{i-2},  l_j = low_j / dia'_j   (fdma.rs:104-107 on the swept system)
  auto lw = [&](int i, int l) {  // low'_{i-2}
    return fma(mu, __ldg(&M.c_low[i - 2]), __ldg(&M.a_low[i - 2])) * ti[didx(i - 2, l)];
  };
  scan1<C::NTHR, C::CL, true>(
      m, red,
      [&](int i, int l) {
        return fma(__ldg(&B.lo[i]), td[didx(i, l)],
                   fma(__ldg(&B.di[i]), td[didx(i + 2, l)], (i + 4 < n) ? __ldg(&B.up[i]) * td[didx(i + 4, l)] : 0.0));
      },
      [&](int i, int l) { return i >= 2 ? -lw(i, l) : 0.0; }, [&](int i, int l, double y) { td[didx(i, l)] = y; });
  // backward: x_i = (x_i - up1'_i x_{i+2} - up2_i x_{i+4}) / dia'_i   (fdma.rs:108-117)
  scan2<C::NTHR, C::CL, false>(
      m, red, [&](int i, int l) { return ti[didx(i, l)] * td[didx(i, l)]; },
      [&](int i, int l) {
        if (i >= m - 2) return 0.0;
        double u1 = fma(mu, __ldg(&M.c_up1[i]), __ldg(&M.a_up1[i]));
        if (i >= 2) u1 = fma(-lw(i, l), fma(mu, __ldg(&M.c_up2[i - 2]), __ldg(&M.a_up2[i - 2])), u1);
        return -u1 * ti[didx(i, l)];
      },
      [&](int i, int l) {
        if (i >= m - 4) return 0.0;
        return -fma(mu, __ldg(&M.c_up2[i]), __ldg(&M.a_up2[i])) * ti[didx(i, l)];
      },
      [&](int i, int l, double y) { td[didx(i, l)] = y; });
  tile_drain<C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.h.rows) a.h.p[(size_t)(r0 + l) * a.h.ld + j] = v;
  });
}

template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 2) yk_project(YProjectArgs a) {
  typedef YCfg<LOG2L> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, m = n - 2;
  // ux -= from_ortho_y(S_y a1)
  tile_fill<C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(a.a1, r0 + l, j, m, a.nsd, a.nsl); });
  __syncthreads();
  from_ortho<C::NTHR, C::CL>(td, -1, n, a.t, red);
  tile_drain<C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.ux.rows) a.ux.p[(size_t)(r0 + l) * a.ux.ld + j] -= v;
  });
  __syncthreads();
  // uy -= from_ortho_y(D_y S_y a2 / sy)
  tile_fill<C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(a.a2, r0 + l, j, m, a.nsd, a.nsl); });
  __syncthreads();
  cheb_diff<C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  from_ortho<C::NTHR, C::CL>(td, -1, n, a.t, red);
  tile_drain<C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.uy.rows) a.uy.p[(size_t)(r0 + l) * a.uy.ld + j] -= v;
  });
}

template <int LOG2L>
__global__ void __launch_bounds__(YCfg<LOG2L>::NTHR, 2) yk_pres(YPresArgs a) {
  typedef YCfg<LOG2L> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * 4;
  constexpr int n = C::n, m = n - 2;
  const int mx = a.phi.rows;
  // to_ortho(phi): S_x across lanes (rows i, i-2 of phi), S_y along the lane
  tile_fill<C::NTHR>(td, n, [&](int j, int l) {
    const int i = min(r0 + l, a.pres.rows - 1);
    const double s0 = ld_stencil(a.phi, min(i, mx - 1), j, m, a.ysd, a.ysl);
    const double s2 = ld_stencil(a.phi, max(i - 2, 0), j, m, a.ysd, a.ysl);
    const double xd = (i < mx) ? __ldg(&a.xsd[min(i, mx - 1)]) : 0.0, xl = (i >= 2) ? __ldg(&a.xsl[max(i - 2, 0)]) : 0.0;
    const double v = fma(xl, s2, xd * s0);
    const double p = fma(-a.nu, a.div.p[(size_t)i * a.div.ld + j], a.pres.p[(size_t)i * a.pres.ld + j]) + v * a.inv_dt;
    return (r0 + l < a.pres.rows) ? p : 0.0;
  });
  __syncthreads();
  tile_drain<C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.pres.rows) a.pres.p[(size_t)(r0 + l) * a.pres.ld + j] = v;
  });
  __syncthreads();
  cheb_diff<C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  tile_drain<C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.dyp.rows) a.dyp.p[(size_t)(r0 + l) * a.dyp.ld + j] = v;
  });
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
static int log2_of(int v) {
  int l = 0;
  while ((1 << l) < v) ++l;
  return ((1 << l) == v) ? l : -1;
}

#define YK_SIZES(X) X(5) X(6) X(9) X(10) X(11)

bool y_supported(int n1) {
  const int l = log2_of(n1 - 1);
#define X(L) \
  if (l == L) return true;
  YK_SIZES(X)
#undef X
  return false;
}

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
    const int nb_ = ((nrows) + 3) / 4;                                                                \
    bool ok_ = false;                                                                                 \
    YK_SIZES(YK_CASE_##kern)                                                                          \
    if (!ok_) throw Error(RP_ERR_INTERNAL, #kern ": unsupported lane length");                        \
  } while (0)

#define YK_CASE_BODY(kern, L, two_tiles, args)                                                \
  if (l_ == L) {                                                                              \
    typedef YCfg<L> C;                                                                        \
    const int sm_ = (two_tiles) ? C::SMEM2 : C::SMEM1;                                        \
    static bool init_ = false;                                                                \
    if (!init_) {                                                                             \
      set_smem(kern<L>, sm_);                                                                 \
      init_ = true;                                                                           \
    }                                                                                         \
    RP_LAUNCH(kern<L>, dim3(nb_, nby_), dim3(C::NTHR), (size_t)sm_, s, args);                       \
    ok_ = true;                                                                               \
  }

#define YK_CASE_yk_backward(L) YK_CASE_BODY(yk_backward, L, false, a)
#define YK_CASE_yk_conv(L) YK_CASE_BODY(yk_conv, L, false, a)
#define YK_CASE_yk_adi(L) YK_CASE_BODY(yk_adi, L, false, a)
#define YK_CASE_yk_mode(L) YK_CASE_BODY(yk_mode, L, true, a)
#define YK_CASE_yk_project(L) YK_CASE_BODY(yk_project, L, false, a)
#define YK_CASE_yk_pres(L) YK_CASE_BODY(yk_pres, L, false, a)

void launch_y_backward(const YBackwardArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_backward, false, a.a[0].a.rows, a.a[0].t.n, a, nb); }
void launch_y_conv(const YConvArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_conv, false, a.a[0].u.rows, a.a[0].t.n, a, nb); }
void launch_y_adi(const YAdiArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_adi, false, a.a[0].w.rows, a.a[0].ny, a, nb); }
void launch_y_mode(const YModeArgs& a, cudaStream_t s) { YK_LAUNCH(yk_mode, true, a.g.rows, a.ny, a, 1); }
void launch_y_project(const YProjectArgs& a, cudaStream_t s) { YK_LAUNCH(yk_project, false, a.a1.rows, a.ny, a, 1); }
void launch_y_pres(const YPresArgs& a, cudaStream_t s) { YK_LAUNCH(yk_pres, false, a.pres.rows, a.ny, a, 1); }

}  // namespace fk
}  // namespace rp
