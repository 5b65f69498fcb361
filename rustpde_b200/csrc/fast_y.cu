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
#include "fast_y.cuh"

namespace rp {
namespace fk {

// ---------------------------------------------------------------------------------
template <int LOG2L, int LC>
FK_DEV void yk_backward_body(const YBackwardArgs& a, const YBackwardArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2, N = C::N;
  auto fill = [&](const Mat& s) {
    tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(s, r0 + l, j, m, a.sd, a.sl); });
    __syncthreads();
  };
  auto drain = [&](const Mat& o) {
    tile_drain<LC, C::NTHR>(td, N, n, [&](int j, int l, double v) {
      if (r0 + l < o.rows) o.p[(size_t)(r0 + l) * o.ld + j] = v;
    });
  };
  fill(a.a);
  if (a.val.p) {
    dct_pow2<LC, LOG2L, C::NTHR, true>(td, a.t, red);
    drain(a.val);
    __syncthreads();
    fill(a.a);
  }
  cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  dct_pow2<LC, LOG2L, C::NTHR, true>(td, a.t, red);
  drain(a.dy);
  __syncthreads();
  fill(a.adx);
  dct_pow2<LC, LOG2L, C::NTHR, true>(td, a.t, red);
  drain(a.dx);
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_backward(YBackwardArgs3 a3) {
  if (blockIdx.y == 0)
    yk_backward_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    yk_backward_body<LOG2L, LC>(a3.a[1], a3);
  else
    yk_backward_body<LOG2L, LC>(a3.a[2], a3);
}

template <int LOG2L, int LC>
FK_DEV void yk_conv_body(const YConvArgs& a, const YConvArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, N = C::N;
  const bool has_bcx = a.bcx.p != nullptr, has_bcy = a.bcy.p != nullptr, has_solid = a.mask.p != nullptr;
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) {
    const int r = min(r0 + l, a.u.rows - 1);
    double gx = a.du.p[(size_t)r * a.du.ld + j], gy = a.dv.p[(size_t)r * a.dv.ld + j];
    const double u = a.u.p[(size_t)r * a.u.ld + j], v = a.v.p[(size_t)r * a.v.ld + j];
    if (has_bcx) gx += a.bcx.p[(size_t)r * a.bcx.ld + j];  // (either may be absent: identically zero, navier.cu rebuild_bc)
    if (has_bcy) gy += a.bcy.p[(size_t)r * a.bcy.ld + j];
    double c = fma(u, gx, v * gy);
    if (has_solid) {  // conv -= -1/eta * mask * (w [+ wbc] - value)
      double w = a.w.p[(size_t)r * a.w.ld + j];
      if (a.wbc.p) w += a.wbc.p[(size_t)r * a.wbc.ld + j];
      c = fma(a.ieta * a.mask.p[(size_t)r * a.mask.ld + j], w - a.sval.p[(size_t)r * a.sval.ld + j], c);
    }
    return (r0 + l < a.u.rows) ? c : 0.0;
  });
  __syncthreads();
  dct_pow2<LC, LOG2L, C::NTHR, false>(td, a.t, red);
  tile_drain<LC, C::NTHR>(td, N, n, [&](int j, int l, double v) {
    if (r0 + l < a.out.rows) a.out.p[(size_t)(r0 + l) * a.out.ld + j] = (j < a.cut) ? v : 0.0;
  });
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_conv(YConvArgs3 a3) {
  if (blockIdx.y == 0)
    yk_conv_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    yk_conv_body<LOG2L, LC>(a3.a[1], a3);
  else
    yk_conv_body<LOG2L, LC>(a3.a[2], a3);
}

template <int LOG2L, int LC>
FK_DEV void yk_adi_body(const YAdiArgs& a, const YAdiArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2;
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_row(a.w, r0 + l, j); });
  __syncthreads();
  b2_fdma_perm<LC, C::NTHR, C::CL>(td, -1, n, a.pt1, a.pt2, red);
  tile_drain<LC, C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.out.rows) a.out.p[(size_t)(r0 + l) * a.out.ld + j] = v;
  });
  if (a.mode == 0) return;
  // ortho coefficients p_j = d_j x_j + l_{j-2} x_{j-2} of the new velocity component
  auto pj = [&](int j, int l) {
    double v = 0.0;
    if (j < m) v = __ldg(&a.sd[j]) * td[didx<LC>(j, l)];
    if (j >= 2) v = fma(__ldg(&a.sl[j - 2]), td[didx<LC>(j - 2, l)], v);
    return v;
  };
  if (a.mode == 1) {
    // batches of 8: the stencil coefficients come from L2, one exposed round trip per element otherwise
    constexpr int U = 8, TOT = n * C::LR;
    for (int it0 = threadIdx.x; it0 < TOT; it0 += C::NTHR * U) {
      double v[U];
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int it = min(it0 + u * C::NTHR, TOT - 1);
        v[u] = pj(it / C::LR, it % C::LR);
      }
#pragma unroll
      for (int u = 0; u < U; ++u) {
        const int it = it0 + u * C::NTHR;
        const int l = it % C::LR, j = it / C::LR;
        if (it < TOT && r0 + l < a.aux.rows) a.aux.p[(size_t)(r0 + l) * a.aux.ld + j] = v[u];
      }
    }
    return;
  }
  __syncthreads();
  scan1<LC, C::NTHR, C::CL, false>(
      n, red, [&](int i, int l) { return (2.0 * (double)i * a.isy) * pj(i, l); }, [](int, int) { return 1.0; },
      [&](int i, int l, double y) {
        if (i >= 1) td[didx<LC>(i - 1, l)] = (i == 1) ? 0.5 * y : y;
        if (i == n - 1) td[didx<LC>(n - 1, l)] = 0.0;
      });
  tile_drain<LC, C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.aux.rows) a.aux.p[(size_t)(r0 + l) * a.aux.ld + j] = v;
  });
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_adi(YAdiArgs3 a3) {
  if (blockIdx.y == 0)
    yk_adi_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    yk_adi_body<LOG2L, LC>(a3.a[1], a3);
  else
    yk_adi_body<LOG2L, LC>(a3.a[2], a3);
}

template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, 1) yk_mode(YModeArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red0);
  double* ti = td + C::TILE;
  double* red = ti + C::LR * C::ROWS;
  (void)red0;
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2;
  stage_pivots<C::LR, C::NTHR>(ti, a.m.inv, a.m.inv_ld, r0, a.g.rows, m);  // pivot reciprocals of the LR rows, asynchronous
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_row(a.g, r0 + l, j); });
  cp_async_wait<0>();
  __syncthreads();
  const double mu = __ldg(&a.m.lam[min(r0 + (int)(threadIdx.x % C::LR), a.g.rows - 1)]) + a.m.alpha;
  mode_solve<LC, C::NTHR, C::CL, C::ROWS, 0>(td, ti, n, a.b2, a.m, mu, red);
  tile_drain<LC, C::NTHR>(td, -1, m, [&](int j, int l, double v) {
    if (r0 + l < a.h.rows) a.h.p[(size_t)(r0 + l) * a.h.ld + j] = v;
  });
}

template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_project(YProjectArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2;
  // ux -= from_ortho_y(S_y a1)
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(a.a1, r0 + l, j, m, a.nsd, a.nsl); });
  __syncthreads();
  from_ortho<LC, C::NTHR, C::CL>(td, -1, n, a.t, red);
  tile_drain_sub<LC, C::NTHR>(td, -1, m, [&](int j, int l) -> double* {
    return (r0 + l < a.ux.rows) ? a.ux.p + (size_t)(r0 + l) * a.ux.ld + j : nullptr;
  });
  __syncthreads();
  // uy -= from_ortho_y(D_y S_y a2 / sy)
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(a.a2, r0 + l, j, m, a.nsd, a.nsl); });
  __syncthreads();
  cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  from_ortho<LC, C::NTHR, C::CL>(td, -1, n, a.t, red);
  tile_drain_sub<LC, C::NTHR>(td, -1, m, [&](int j, int l) -> double* {
    return (r0 + l < a.uy.rows) ? a.uy.p + (size_t)(r0 + l) * a.uy.ld + j : nullptr;
  });
}

template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_pres(YPresArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2;
  const int mx = a.phi.rows;
  // to_ortho(phi): S_x across lanes (rows i, i-2 of phi), S_y along the lane
  if (a.only_dyp) {  // block-uniform
    tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_row(a.pres, r0 + l, j); });
    __syncthreads();
    cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
    tile_drain<LC, C::NTHR>(td, -1, n, [&](int j, int l, double v) {
      if (r0 + l < a.dyp.rows) a.dyp.p[(size_t)(r0 + l) * a.dyp.ld + j] = v;
    });
    return;
  }
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) {
    const int i = min(r0 + l, a.pres.rows - 1);
    const double s0 = ld_stencil(a.phi, min(i, mx - 1), j, m, a.ysd, a.ysl);
    const double s2 = ld_stencil(a.phi, max(i - 2, 0), j, m, a.ysd, a.ysl);
    const double xd = (i < mx) ? __ldg(&a.xsd[min(i, mx - 1)]) : 0.0, xl = (i >= 2) ? __ldg(&a.xsl[max(i - 2, 0)]) : 0.0;
    const double v = fma(xl, s2, xd * s0);
    const double p = fma(-a.nu, a.div.p[(size_t)i * a.div.ld + j], a.pres.p[(size_t)i * a.pres.ld + j]) + v * a.inv_dt;
    return (r0 + l < a.pres.rows) ? p : 0.0;
  });
  __syncthreads();
  tile_drain<LC, C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.pres.rows) a.pres.p[(size_t)(r0 + l) * a.pres.ld + j] = v;
  });
  __syncthreads();
  cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  tile_drain<LC, C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < a.dyp.rows) a.dyp.p[(size_t)(r0 + l) * a.dyp.ld + j] = v;
  });
}

// blockIdx.y = 0: vx = S_y ux;  1: ey = D_y S_y uy / sy   (the y parts of the divergence, for the diagnostics)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) yk_divprep(YDivPrepArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * C::LR;
  constexpr int n = C::n, m = n - 2;
  const bool second = blockIdx.y != 0;
  const Mat& src = second ? a.uy : a.ux;
  const Mat& dst = second ? a.ey : a.vx;
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ld_stencil(src, r0 + l, j, m, a.sd, a.sl); });
  __syncthreads();
  if (second) cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  tile_drain<LC, C::NTHR>(td, -1, n, [&](int j, int l, double v) {
    if (r0 + l < dst.rows) dst.p[(size_t)(r0 + l) * dst.ld + j] = v;
  });
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
ScanShape y_scan_shape(int n1) {
  const int l = log2_of(n1 - 1);
#define X(L, LCV) \
  if (l == L) return ScanShape{scan_threads(YCfg<L, LCV>::NTHR) / (4 * LCV), YCfg<L, LCV>::CL};
  YK_SIZES(X)
#undef X
  return ScanShape{0, 0};
}

std::vector<double> perm_table(int m, bool fwd, ScanShape sh, int W, const std::vector<std::vector<double>>& src,
                               const std::vector<int>& shift) {
  const int ng = sh.ng, cl = (((m + 1) >> 1) + ng - 1) / ng;
  if (cl > sh.rows) throw Error(RP_ERR_INTERNAL, "perm_table: chunk longer than the kernel's bound");
  std::vector<double> out((size_t)sh.rows * ng * 2 * W, 0.0);
  for (int u = 0; u < cl; ++u)
    for (int g = 0; g < ng; ++g)
      for (int p = 0; p < 2; ++p) {
        const int Mp = (m - p + 1) >> 1, t = g * cl + u;
        if (t >= Mp) continue;
        const int i = 2 * (fwd ? t : Mp - 1 - t) + p;
        const size_t slot = ((size_t)u * ng + g) * 2 + p;
        for (int k = 0; k < (int)src.size(); ++k) {
          const int j = i + shift[k];
          if (j >= 0 && j < (int)src[k].size()) out[slot * W + k] = src[k][j];
        }
      }
  return out;
}

bool y_supported(int n1) {
  const int l = log2_of(n1 - 1);
#define X(L, LCV) \
  if (l == L) return true;
  YK_SIZES(X)
#undef X
  return false;
}

#define YK_CASE_yk_backward(L, LCV) YK_CASE_BODY(yk_backward, L, LCV, 0, a)
#define YK_CASE_yk_conv(L, LCV) YK_CASE_BODY(yk_conv, L, LCV, 0, a)
#define YK_CASE_yk_adi(L, LCV) YK_CASE_BODY(yk_adi, L, LCV, 0, a)
#define YK_CASE_yk_mode(L, LCV) YK_CASE_BODY(yk_mode, L, LCV, 1, a)
#define YK_CASE_yk_project(L, LCV) YK_CASE_BODY(yk_project, L, LCV, 0, a)
#define YK_CASE_yk_pres(L, LCV) YK_CASE_BODY(yk_pres, L, LCV, 0, a)
#define YK_CASE_yk_divprep(L, LCV) YK_CASE_BODY(yk_divprep, L, LCV, 0, a)

void launch_y_backward(const YBackwardArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_backward, false, a.a[0].a.rows, a.a[0].t.n, a, nb); }
void launch_y_conv(const YConvArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_conv, false, a.a[0].u.rows, a.a[0].t.n, a, nb); }
void launch_y_adi(const YAdiArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(yk_adi, false, a.a[0].w.rows, a.a[0].ny, a, nb); }
void launch_y_mode(const YModeArgs& a, cudaStream_t s) { YK_LAUNCH(yk_mode, true, a.g.rows, a.ny, a, 1); }
void launch_y_project(const YProjectArgs& a, cudaStream_t s) { YK_LAUNCH(yk_project, false, a.a1.rows, a.ny, a, 1); }
void launch_y_pres(const YPresArgs& a, cudaStream_t s) { YK_LAUNCH(yk_pres, false, a.pres.rows, a.ny, a, 1); }
void launch_y_divprep(const YDivPrepArgs& a, cudaStream_t s) { YK_LAUNCH(yk_divprep, false, a.ux.rows, a.ny, a, 2); }

}  // namespace fk
}  // namespace rp
