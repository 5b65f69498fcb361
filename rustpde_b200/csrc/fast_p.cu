// fast_p.cu -- specialised kernels of the periodic Navier2D::update (Fourier x Chebyshev,
// navier.rs:384-467 + 737-765): real FFT along x on column strips, per-mode passes on complex rows.
//
//   pk_c2r      : x-backward, c2r (r2c.rs:283-303) of a spectral array and of its (ik/sx) derivative
//   pk_r2c      : x-forward, r2c (r2c.rs:250-272) + dealias cut in kx (navier.rs:1022-1032)
//   pk_hholtz   : rhs assembly (navier.rs:622-674) + per-mode Helmholtz solve (hholtz.rs:156-197)
//   pk_divpois  : divergence (navier.rs:698-703) + per-mode Poisson solve (poisson.rs:131-149)
//   pk_project  : projection and pressure update (navier.rs:683-721)
// The y passes in physical space (B_y S_y, products, forward DCT-y) are the kernels of fast_y.cu.
//
// x kernels: a block owns 4 real columns = two packed complex FFT lanes (columns a, b of a lane travel
// as a + i b through one complex FFT of length n).  y kernels: a block owns 2 complex rows = 4 real
// lanes (re, im of row r0, re, im of row r0 + 1); both parts of a row share the row's eigenvalue.
#include "fast_y.cuh"

namespace rp {
namespace fk {

// rank that owns global index i of a scattered axis
FK_DEV int owner_of(const Scatter& sc, int i) {
  int q = 0;
  while (q + 1 < sc.nparts && i >= sc.beg[q + 1]) ++q;
  return q;
}

// ---- x kernels ---------------------------------------------------------------------------------
template <int LOG2N, int LC_>
struct PXCfg {
  static constexpr int LC = LC_;
  static constexpr int N = 1 << LOG2N;
  static constexpr int NTHR = (N * LC / 16) < 64 ? 64 : (N * LC / 16);
  static constexpr int SMEM = N * LC * 16;
};

template <int LOG2N, int LC>
FK_DEV void pk_c2r_body(const PC2rArgs& a, const PC2rArgs3& a3) {
  typedef PXCfg<LOG2N, LC> C;
  RP_DYN_SMEM(double, td);
  cplx* tc = (cplx*)td;
  constexpr int n = C::N, mkr = n / 2 + 1;
  const int c0 = blockIdx.x * 2 * LC;
  const int ncols = a.src.cols;
  const cplx* src = (const cplx*)a.src.p;
  const double inv_n = 1.0 / (double)n;
  for (int pass = 0; pass < 2; ++pass) {
    const Mat& o = pass ? a.dx : a.val;
    if (o.p == nullptr) continue;  // block-uniform
    for (int it = threadIdx.x; it < mkr * LC; it += C::NTHR) {
      const int k = it / LC, c = it % LC;
      const int ca = c0 + 2 * c, cb = ca + 1;
      cplx xa = mk(0.0, 0.0), xb = mk(0.0, 0.0);
      if (ca < ncols) xa = src[(size_t)k * a.src.ld + ca];
      if (cb < ncols) xb = src[(size_t)k * a.src.ld + cb];
      if (pass) {  // times i k / sx   (r2c.rs:88-99, space2.rs:247-264)
        const double kk = (double)k * a.isx;
        xa = mk(-kk * xa.y, kk * xa.x);
        xb = mk(-kk * xb.y, kk * xb.x);
      }
      if (k == 0 || 2 * k == n) {  // c2r ignores Im of the DC and Nyquist bins
        xa.y = 0.0;
        xb.y = 0.0;
      }
      tc[cidx<LC>(k, c)] = mk(xa.x - xb.y, xa.y + xb.x);                                    // X_a + i X_b
      if (k > 0 && 2 * k != n) tc[cidx<LC>(n - k, c)] = mk(xa.x + xb.y, -xa.y + xb.x);     // conj(X_a) + i conj(X_b)
    }
    __syncthreads();
    fft<LC, LOG2N, C::NTHR, true, false>(tc, a.tw, nullptr);
    for (int it = threadIdx.x; it < n * LC; it += C::NTHR) {
      const int i = it / LC, c = it % LC;
      const int ca = c0 + 2 * c;
      const cplx z = tc[cidx<LC>(i, c)];
      double* row = o.p + (size_t)i * o.ld;
      if (ca < o.cols) row[ca] = z.x * inv_n;
      if (ca + 1 < o.cols) row[ca + 1] = z.y * inv_n;
    }
    __syncthreads();
  }
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2N, int LC>
__global__ void __launch_bounds__(PXCfg<LOG2N, LC>::NTHR, PXCfg<LOG2N, LC>::SMEM > 110 * 1024 ? 1 : 2) pk_c2r(PC2rArgs3 a3) {
  if (blockIdx.y == 0)
    pk_c2r_body<LOG2N, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    pk_c2r_body<LOG2N, LC>(a3.a[1], a3);
  else
    pk_c2r_body<LOG2N, LC>(a3.a[2], a3);
}

template <int LOG2N, int LC>
FK_DEV void pk_r2c_body(const PR2cArgs& a, const PR2cArgs3& a3) {
  typedef PXCfg<LOG2N, LC> C;
  RP_DYN_SMEM(double, td);
  cplx* tc = (cplx*)td;
  constexpr int n = C::N, mkr = n / 2 + 1;
  const int c0 = blockIdx.x * 2 * LC;
  const int ncols = a.dst.cols;
  for (int it = threadIdx.x; it < n * LC; it += C::NTHR) {
    const int i = it / LC, c = it % LC;
    const int ca = c0 + 2 * c;
    double va = 0.0, vb = 0.0;
    if (a.u.p) {  // physical-space products (conv_term.rs:41) fused into the load
      auto prod = [&](int cc) {
        const size_t o = (size_t)i;
        double gx = a.du.p[o * a.du.ld + cc], gy = a.dv.p[o * a.dv.ld + cc];
        if (a.bcx.p) gx += a.bcx.p[o * a.bcx.ld + cc];
        if (a.bcy.p) gy += a.bcy.p[o * a.bcy.ld + cc];
        return fma(a.u.p[o * a.u.ld + cc], gx, a.v.p[o * a.v.ld + cc] * gy);
      };
      if (ca < ncols) va = prod(ca);
      if (ca + 1 < ncols) vb = prod(ca + 1);
    } else {
      const double* row = a.src.p + (size_t)i * a.src.ld;
      if (ca < ncols) va = row[ca];
      if (ca + 1 < ncols) vb = row[ca + 1];
    }
    tc[cidx<LC>(i, c)] = mk(va, vb);
  }
  __syncthreads();
  fft<LC, LOG2N, C::NTHR, false, false>(tc, a.tw, nullptr);
  cplx* dst = (cplx*)a.dst.p;
  for (int it = threadIdx.x; it < mkr * LC; it += C::NTHR) {
    const int k = it / LC, c = it % LC;
    const int ca = c0 + 2 * c, cb = ca + 1;
    const cplx zk = tc[cidx<LC>(k, c)], zc = tc[cidx<LC>(k == 0 ? 0 : n - k, c)];
    const cplx zm = mk(zc.x, -zc.y);
    const cplx s = cadd(zk, zm), d = csub(zk, zm);
    const double h = (k < a.cut) ? 0.5 : 0.0;  // dealias: modes kx >= cut are zeroed
    cplx* row;
    if (a.sdst.nparts > 0) {  // fused transpose: row k lives on the rank that owns mode k
      const int q = owner_of(a.sdst, k);
      row = (cplx*)a.sdst.ptr[q] + (size_t)(k - a.sdst.beg[q]) * a.ny + a.j0;
    } else {
      row = dst + (size_t)k * a.dst.ld;
    }
    if (ca < ncols) row[ca] = mk(h * s.x, h * s.y);
    if (cb < ncols) row[cb] = mk(h * d.y, -h * d.x);
  }
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2N, int LC>
__global__ void __launch_bounds__(PXCfg<LOG2N, LC>::NTHR, PXCfg<LOG2N, LC>::SMEM > 110 * 1024 ? 1 : 2) pk_r2c(PR2cArgs3 a3) {
  if (blockIdx.y == 0)
    pk_r2c_body<LOG2N, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    pk_r2c_body<LOG2N, LC>(a3.a[1], a3);
  else
    pk_r2c_body<LOG2N, LC>(a3.a[2], a3);
}

// ---- y kernels on complex rows -------------------------------------------------------------------
// lane l of the tile: row r0 + (l >> 1), part l & 1 (re / im)
FK_DEV int prow_of(int r0, int l) { return r0 + (l >> 1); }  // lanes (2r, 2r+1) = (re, im) of row r0 + r
// element (row, j).part of a complex array, zero outside [0, rows) x [0, cols)
FK_DEV double ldc(const Mat& a, int r, int j, int part) {
  const double v = a.p[((size_t)min(r, a.rows - 1) * a.ld + min(max(j, 0), a.cols - 1)) * 2 + part];
  return (r < a.rows && j >= 0 && j < a.cols) ? v : 0.0;
}
// composite -> ortho along y while loading: p_j = d_j c_j + l_{j-2} c_{j-2}
FK_DEV double ldc_stencil(const Mat& a, int r, int j, int part, const double* __restrict__ sd, const double* __restrict__ sl) {
  const int m = a.cols;
  const double v0 = ldc(a, r, min(j, m - 1), part), v2 = ldc(a, r, max(j - 2, 0), part);
  const double d = (j < m) ? __ldg(&sd[min(j, m - 1)]) : 0.0;
  const double l = (j >= 2) ? __ldg(&sl[max(j - 2, 0)]) : 0.0;
  return fma(l, v2, d * v0);
}
// (i k s z).part for z = (re, im): re' = -k s im, im' = k s re
FK_DEV double ik_part(double ks, double re, double im, int part) { return part ? ks * re : -ks * im; }

#ifndef PK_FILL_U
#define PK_FILL_U 8  // loads of the rhs assembly kept in flight per thread (elements per batch)
#endif
template <int LOG2L, int LC>
FK_DEV void pk_hholtz_body(const PHholtzArgs& a, const PHholtzArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red0);
  double* ti = td + C::TILE;
  double* red = ti + LC * C::ROWS;
  (void)red0;
  const int r0 = blockIdx.x * LC;
  constexpr int n = C::n, m = n - 2;
  const int nrows = a.chat.rows;
  stage_pivots<LC, C::NTHRS>(ti, a.m.inv, a.m.inv_ld, r0, nrows, m);  // pivot reciprocals of the LC complex rows, asynchronous
  if (a.mode == 1) {  // - dt/sy d/dy pres   (navier.rs:646)
    tile_fill<LC, C::NTHRS>(td, n, [&](int j, int l) { return ldc(a.pres, prow_of(r0, l), j, l & 1); });
    __syncthreads();
    cheb_diff<LC, C::NTHRS, C::CL>(td, -1, td, -1, n, -a.dt * a.isy, red);
  }
  {
    const int tot = n * C::LR;
    constexpr int FU = PK_FILL_U;
    for (int it0 = threadIdx.x; it0 < tot; it0 += C::NTHRS * FU) {
      double v[FU];
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        const int it = min(it0 + u * C::NTHRS, tot - 1);
        const int j = it / C::LR, l = it % C::LR, r = prow_of(r0, l), part = l & 1;
        double x = -a.dt * ldc(a.chat, r, j, part);                       // - dt * conv          (630, 651, 671)
        x += ldc_stencil(a.fld, r, j, part, a.sd, a.sl);                  // + to_ortho(field)    (625, 644, 663)
        if (a.mode == 0) {                                                // - dt/sx d/dx pres    (627)
          const double ks = -a.dt * a.isx * (double)(a.k0 + min(r, nrows - 1));
          x += ik_part(ks, ldc(a.pres, r, j, 0), ldc(a.pres, r, j, 1), part);
        } else if (a.mode == 1) {                                         // + dt * (that + tbc)  (648)
          x = fma(a.dt, ldc_stencil(a.tmp, r, j, part, a.tsd, a.tsl) + ldc(a.tbc, r, j, part), x);
        } else {                                                          // + dt ka lap(fieldbc) (665-668)
          x += ldc(a.bcdiff, r, j, part);
        }
        v[u] = x;
      }
#pragma unroll
      for (int u = 0; u < FU; ++u) {
        const int it = it0 + u * C::NTHRS;
        if (it < tot) {
          double* w = &td[didx<LC>(it / C::LR, it % C::LR)];
          *w = (a.mode == 1) ? *w + v[u] : v[u];
        }
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  const double mu = __ldg(&a.m.lam[min(prow_of(r0, threadIdx.x % C::LR), nrows - 1)]) + a.m.alpha;
  mode_solve<LC, C::NTHRS, C::CL, C::ROWS, 1>(td, ti, n, a.b2, a.m, mu, red);
  tile_drain<LC, C::NTHRS>(td, -1, m, [&](int j, int l, double v) {
    const int r = prow_of(r0, l);
    if (r < a.out.rows) a.out.p[((size_t)r * a.out.ld + j) * 2 + (l & 1)] = v;
  });
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHRS, 1) pk_hholtz(PHholtzArgs3 a3) {
  if (blockIdx.y == 0)
    pk_hholtz_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    pk_hholtz_body<LOG2L, LC>(a3.a[1], a3);
  else
    pk_hholtz_body<LOG2L, LC>(a3.a[2], a3);
}

template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHRS, 1) pk_divpois(PDivPoisArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red0);
  double* ti = td + C::TILE;
  double* red = ti + LC * C::ROWS;
  (void)red0;
  const int r0 = blockIdx.x * LC;
  constexpr int n = C::n, m = n - 2;
  const int nrows = a.ux.rows;
  // div = i k / sx S_y ux + D_y S_y uy / sy   (navier.rs:698-703)
  stage_pivots<LC, C::NTHRS>(ti, a.m.inv, a.m.inv_ld, r0, nrows, m);  // pivot reciprocals of the LC complex rows, asynchronous
  tile_fill<LC, C::NTHRS>(td, n, [&](int j, int l) { return ldc_stencil(a.uy, prow_of(r0, l), j, l & 1, a.sd, a.sl); });
  cp_async_wait<0>();
  __syncthreads();
  cheb_diff<LC, C::NTHRS, C::CL>(td, -1, td, -1, n, a.isy, red);
  for (int it = threadIdx.x; it < n * C::LR; it += C::NTHRS) {
    const int j = it / C::LR, l = it % C::LR, r = prow_of(r0, l), part = l & 1;
    const double ks = a.isx * (double)(a.k0 + min(r, nrows - 1));
    const double re = ldc_stencil(a.ux, r, j, 0, a.sd, a.sl), im = ldc_stencil(a.ux, r, j, 1, a.sd, a.sl);
    const double v = td[didx<LC>(j, l)] + ik_part(ks, re, im, part);
    td[didx<LC>(j, l)] = v;
    if (r < a.div.rows) a.div.p[((size_t)r * a.div.ld + j) * 2 + part] = v;
  }
  __syncthreads();
  const double mu = __ldg(&a.m.lam[min(prow_of(r0, threadIdx.x % C::LR), nrows - 1)]) + a.m.alpha;
  mode_solve<LC, C::NTHRS, C::CL, C::ROWS, 1>(td, ti, n, a.b2, a.m, mu, red);
  tile_drain<LC, C::NTHRS>(td, -1, m, [&](int j, int l, double v) {
    const int r = prow_of(r0, l);
    if (a.k0 + r == 0 && j == 0) v = 0.0;  // pres[1].vhat[[0,0]] = 0   (navier.rs:714)
    if (r < a.phi.rows) a.phi.p[((size_t)r * a.phi.ld + j) * 2 + (l & 1)] = v;
  });
}

template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHRS, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) pk_project(PProjectArgs a) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * LC;
  constexpr int n = C::n, m = n - 2;
  const int nrows = a.phi.rows;
  // ux -= from_ortho_y(i k / sx S_y phi)   (navier.rs:683-695)
  tile_fill<LC, C::NTHRS>(td, n, [&](int j, int l) {
    const int r = prow_of(r0, l);
    const double ks = a.isx * (double)(a.k0 + min(r, nrows - 1));
    return ik_part(ks, ldc_stencil(a.phi, r, j, 0, a.nsd, a.nsl), ldc_stencil(a.phi, r, j, 1, a.nsd, a.nsl), l & 1);
  });
  __syncthreads();
  from_ortho<LC, C::NTHRS, C::CL>(td, -1, n, a.t, red);
  tile_drain_sub<LC, C::NTHRS>(td, -1, m, [&](int j, int l) -> double* {
    const int r = prow_of(r0, l);
    return (r < a.ux.rows) ? a.ux.p + ((size_t)r * a.ux.ld + j) * 2 + (l & 1) : nullptr;
  });
  __syncthreads();
  // to_ortho(phi): pressure update p += -nu div + to_ortho(phi) / dt   (navier.rs:717-721)
  tile_fill<LC, C::NTHRS>(td, n, [&](int j, int l) { return ldc_stencil(a.phi, prow_of(r0, l), j, l & 1, a.nsd, a.nsl); });
  __syncthreads();
  for (int it0 = threadIdx.x; it0 < n * C::LR; it0 += C::NTHRS * 4) {
    double dv[4], pv[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = min(it0 + u * C::NTHRS, n * C::LR - 1);
      dv[u] = ldc(a.div, prow_of(r0, it % C::LR), it / C::LR, it & 1);
      pv[u] = ldc(a.pres, prow_of(r0, it % C::LR), it / C::LR, it & 1);
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = it0 + u * C::NTHRS;
      if (it < n * C::LR) {
        const int j = it / C::LR, l = it % C::LR, r = prow_of(r0, l);
        if (r < a.pres.rows) a.pres.p[((size_t)r * a.pres.ld + j) * 2 + (l & 1)] = fma(-a.nu, dv[u], pv[u]) + td[didx<LC>(j, l)] * a.inv_dt;
      }
    }
  }
  __syncthreads();
  // uy -= from_ortho_y(D_y S_y phi / sy)
  cheb_diff<LC, C::NTHRS, C::CL>(td, -1, td, -1, n, a.isy, red);
  from_ortho<LC, C::NTHRS, C::CL>(td, -1, n, a.t, red);
  tile_drain_sub<LC, C::NTHRS>(td, -1, m, [&](int j, int l) -> double* {
    const int r = prow_of(r0, l);
    return (r < a.uy.rows) ? a.uy.p + ((size_t)r * a.uy.ld + j) * 2 + (l & 1) : nullptr;
  });
}

// ---- y transforms on complex rows (slab decomposition over kx) ------------------------------------------
template <int LOG2L, int LC>
FK_DEV void pk_ybackward_body(const PYBackArgs& a, const PYBackArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * LC;
  constexpr int n = C::n, N = C::N;
  auto fill = [&]() {
    tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ldc_stencil(a.src, prow_of(r0, l), j, l & 1, a.sd, a.sl); });
    __syncthreads();
  };
  auto drain = [&](const Mat& o, const Scatter& sc) {
    tile_drain<LC, C::NTHR>(td, N, n, [&](int j, int l, double v) {
      const int r = prow_of(r0, l);
      if (r >= o.rows) return;
      if (sc.nparts > 0) {  // fused transpose: column j lives on the rank that owns it
        const int q = owner_of(sc, j);
        sc.ptr[q][((size_t)(a.k0 + r) * (sc.beg[q + 1] - sc.beg[q]) + (j - sc.beg[q])) * 2 + (l & 1)] = v;
      } else {
        o.p[((size_t)r * o.ld + j) * 2 + (l & 1)] = v;
      }
    });
  };
  fill();
  if (a.val.p) {
    dct_pow2<LC, LOG2L, C::NTHR, true>(td, a.t, red);
    drain(a.val, a.sval);
    __syncthreads();
    fill();
  }
  cheb_diff<LC, C::NTHR, C::CL>(td, -1, td, -1, n, a.isy, red);
  dct_pow2<LC, LOG2L, C::NTHR, true>(td, a.t, red);
  drain(a.dy, a.sdy);
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) pk_ybackward(PYBackArgs3 a3) {
  if (blockIdx.y == 0)
    pk_ybackward_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    pk_ybackward_body<LOG2L, LC>(a3.a[1], a3);
  else
    pk_ybackward_body<LOG2L, LC>(a3.a[2], a3);
}

template <int LOG2L, int LC>
FK_DEV void pk_yforward_body(const PYFwdArgs& a, const PYFwdArgs3& a3) {
  typedef YCfg<LOG2L, LC> C;
  YK_SMEM(td, red);
  const int r0 = blockIdx.x * LC;
  constexpr int n = C::n, N = C::N;
  tile_fill<LC, C::NTHR>(td, n, [&](int j, int l) { return ldc(a.src, prow_of(r0, l), j, l & 1); });
  __syncthreads();
  dct_pow2<LC, LOG2L, C::NTHR, false>(td, a.t, red);
  tile_drain<LC, C::NTHR>(td, N, n, [&](int j, int l, double v) {
    const int r = prow_of(r0, l);
    if (r < a.dst.rows) a.dst.p[((size_t)r * a.dst.ld + j) * 2 + (l & 1)] = (j < a.cut) ? v : 0.0;
  });
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2L, int LC>
__global__ void __launch_bounds__(YCfg<LOG2L, LC>::NTHR, YCfg<LOG2L, LC>::SMEM1 > 110 * 1024 ? 1 : 2) pk_yforward(PYFwdArgs3 a3) {
  if (blockIdx.y == 0)
    pk_yforward_body<LOG2L, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    pk_yforward_body<LOG2L, LC>(a3.a[1], a3);
  else
    pk_yforward_body<LOG2L, LC>(a3.a[2], a3);
}

// ---- launchers -----------------------------------------------------------------------------------
#define PX_SIZES(X) X(5, 2) X(6, 2) X(7, 1) X(8, 2) X(9, 2) X(10, 2) X(11, 2) X(12, 1) X(13, 1)

bool px_supported(int n0) {
  const int l = log2_of(n0);
#define X(L, LCV) \
  if (l == L) return true;
  PX_SIZES(X)
#undef X
  return false;
}

#define PX_CASE(kern, L, LCV)                                                         \
  if (l_ == L) {                                                                      \
    typedef PXCfg<L, LCV> C;                                                          \
    auto kp_ = kern<L, LCV>;                                                          \
    const int nb_ = ((ncols_) + 2 * LCV - 1) / (2 * LCV);                             \
    static unsigned long long init_ = 0; /* one bit per device */                                                        \
    if (first_use_on_device(init_)) {                                                                     \
      set_smem(kp_, C::SMEM);                                                         \
    }                                                                                 \
    RP_LAUNCH(kp_, dim3(nb_, nby_), dim3(C::NTHR), (size_t)C::SMEM, s, a);            \
    ok_ = true;                                                                       \
  }
#define PX_CASE_pk_c2r(L, LCV) PX_CASE(pk_c2r, L, LCV)
#define PX_CASE_pk_r2c(L, LCV) PX_CASE(pk_r2c, L, LCV)
#define PX_LAUNCH(kern, ncols, nx, nby)                                            \
  do {                                                                             \
    const int l_ = log2_of(nx);                                                    \
    const int ncols_ = (ncols), nby_ = (nby);                                      \
    bool ok_ = false;                                                              \
    PX_SIZES(PX_CASE_##kern)                                                       \
    if (!ok_) throw Error(RP_ERR_INTERNAL, #kern ": unsupported lane length");     \
  } while (0)

void launch_p_c2r(const PC2rArgs3& a, int nb, cudaStream_t s) { PX_LAUNCH(pk_c2r, a.a[0].src.cols, a.a[0].n, nb); }
void launch_p_r2c(const PR2cArgs3& a, int nb, cudaStream_t s) { PX_LAUNCH(pk_r2c, a.a[0].dst.cols, a.a[0].n, nb); }

#define YK_CASE_pk_hholtz(L, LCV) YK_CASE_BODY_T(pk_hholtz, L, LCV, 2, a, NTHRS)
#define YK_CASE_pk_divpois(L, LCV) YK_CASE_BODY_T(pk_divpois, L, LCV, 2, a, NTHRS)
#define YK_CASE_pk_project(L, LCV) YK_CASE_BODY_T(pk_project, L, LCV, 0, a, NTHRS)

// complex rows: a block owns 2 rows -> the launch helper's "rows / 4" becomes "2 * rows / 4"
void launch_p_hholtz(const PHholtzArgs3& a, int nb, cudaStream_t s) {
  if (pw_enabled(a.a[0].chat.rows, false) && a.a[0].m.rf && a.a[0].dyp.p && a.a[0].rs) return launch_pw_hholtz(a, nb, s);
  YK_LAUNCH(pk_hholtz, true, 2 * a.a[0].chat.rows, a.a[0].ny, a, nb);
}
void launch_p_divpois(const PDivPoisArgs& a, cudaStream_t s) {
  if (pw_enabled(a.ux.rows, true) && a.m.rf && a.rs) return launch_pw_divpois(a, s);
  YK_LAUNCH(pk_divpois, true, 2 * a.ux.rows, a.ny, a, 1);
}
#define YK_CASE_pk_ybackward(L, LCV) YK_CASE_BODY(pk_ybackward, L, LCV, 0, a)
#define YK_CASE_pk_yforward(L, LCV) YK_CASE_BODY(pk_yforward, L, LCV, 0, a)
void launch_p_ybackward(const PYBackArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(pk_ybackward, false, 2 * a.a[0].src.rows, a.a[0].t.n, a, nb); }
void launch_p_yforward(const PYFwdArgs3& a, int nb, cudaStream_t s) { YK_LAUNCH(pk_yforward, false, 2 * a.a[0].src.rows, a.a[0].t.n, a, nb); }
void launch_p_project(const PProjectArgs& a, cudaStream_t s) {
  if (pw_project_enabled(a.phi.rows) && a.w1 && a.w2 && a.z1.p && a.z2.p) return launch_pw_project(a, s);
  YK_LAUNCH(pk_project, false, 2 * a.phi.rows, a.ny, a, 1);
}

}  // namespace fk
}  // namespace rp
