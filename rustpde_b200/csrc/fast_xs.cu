// fast_xs.cu -- x-direction banded sweeps of the confined Navier2D::update as STREAMING column scans.
//
// Along x the lanes are strided in memory, but 8 adjacent columns of a row are one 64-byte run.  The recurrences of
// the reference along x (Fdma forward / backward sweep fdma.rs:101-118, the TDMA of from_ortho linalg.rs:14-57, the
// Chebyshev derivative ortho.rs:107-125) need no shared-memory tile for that: a block owns a strip of 8 columns, a
// thread owns one chunk of one parity chain of one column (fast.cuh scan1n / scan2n: chunk maps chained through a
// small shared array, two walks), reads its chunk straight from global memory -- every load of a warp is 4 runs of
// 64 contiguous bytes, whole sectors -- and keeps it in registers.  The strip of a block (8 x 2048 x 8 B = 128 KB
// per array) lives in the L1 that the missing tile leaves free, so the stages of a kernel hand their intermediates
// to each other through it.  Compared with the tile kernels of fast_x.cu (one 512-thread block per SM beside a 197 KB
// Bluestein tile) these passes run 1024 threads per SM with ~16 independent loads in flight per thread.
//
//   xs_rhs_adi : rhs assembly (navier.rs:622-674) + x half of HholtzAdi (hholtz_adi.rs:108,128)
//   xs_div     : divergence (navier.rs:698-703) + B2_x of the Poisson rhs
//   xs_project : x part of u -= from_ortho(grad phi) (navier.rs:683-695)
//   xs_dxp     : -dt/sx D_x pres for the next step's ux rhs (navier.rs:627)
#include "fast.cuh"

namespace rp {
namespace fk {

namespace {
constexpr int XS_NMAX = 2049;  // longest x lane (XK_SIZES: Bluestein length <= 4096)
// A block owns LR = 2 LC adjacent columns and runs NTHR threads: 2 LR parity chains x NG = 64 chunks of <= 17 elements.
//   XsCfg<4, 1024>: 8 columns (64-byte runs), one block per SM;  XsCfg<2, 512>: 4 columns (one 32-byte sector), two blocks per SM
template <int LC_, int NTHR_>
struct XsCfg {
  static constexpr int LC = LC_, LR = 2 * LC_, NTHR = NTHR_, NCH = 2 * LR;
  static constexpr int NG = NTHR / NCH;
  static constexpr int CL = (((XS_NMAX + 1) / 2) + NG - 1) / NG;
  static constexpr int RED = (NG * NCH + ((NG + 7) / 8) * NCH) * 6 * 8;  // scan2n scratch (bytes)
  static constexpr int MINB = NTHR > 512 ? 1 : 2;
};

struct Strip {  // the 8-column strip of a block
  int c0;
  FK_DEV int col(int l) const { return c0 + l; }
};
FK_DEV double gld(const Mat& a, int i, int c) {  // a[i][c], zero outside
  return (i >= 0 && i < a.rows && c >= 0 && c < a.cols) ? a.p[(size_t)i * a.ld + c] : 0.0;
}
// elementwise loop over rows [0, nrows) x the 8 columns of the strip, f(i, col); consecutive threads = consecutive columns
template <class C, class F>
FK_DEV void strip_rows(int nrows, int c0, F f) {
  const int l = threadIdx.x % C::LR;
  for (int i = threadIdx.x / C::LR; i < nrows; i += C::NTHR / C::LR) f(i, c0 + l);
}
// S_x S_y f at (i, c): composite -> ortho in both directions (composite_stencil.rs:207-229), f is [mx, my]
FK_DEV double sxsy(const Mat& f, int i, int c, const double* __restrict__ xsd, const double* __restrict__ xsl,
                   const double* __restrict__ ysd, const double* __restrict__ ysl) {
  const int mx = f.rows, my = f.cols;
  const double yd = (c < my) ? __ldg(&ysd[c]) : 0.0, yl = (c >= 2 && c - 2 < my) ? __ldg(&ysl[c - 2]) : 0.0;
  auto g = [&](int r) { return fma(yl, gld(f, r, c - 2), yd * gld(f, r, c)); };
  double v = 0.0;
  if (i < mx) v = __ldg(&xsd[i]) * g(i);
  if (i >= 2 && i - 2 < mx) v = fma(__ldg(&xsl[i - 2]), g(i - 2), v);
  return v;
}
// Chebyshev derivative along x of src (n rows), times sc, into dst (ortho.rs:107-125); src_at(i, c) gives the coefficient
template <class C, class Src>
FK_DEV void xs_cheb_diff(int n, int c0, const Mat& dst, double sc, double* red, Src src_at) {
  scan1n<C::LC, C::NTHR, C::CL, false, C::NTHR>(
      n, red, [&](int i, int l) { return (2.0 * (double)i * sc) * src_at(i, c0 + l); }, [](int, int) { return 1.0; },
      [&](int i, int l, double y) {
        const int c = c0 + l;
        if (c >= dst.cols) return;
        if (i >= 1) dst.p[(size_t)(i - 1) * dst.ld + c] = (i == 1) ? 0.5 * y : y;
        if (i == n - 1) dst.p[(size_t)(n - 1) * dst.ld + c] = 0.0;
      });
}
// from_ortho along x (composite_stencil.rs:250-276): src (n rows, ortho) -> dst (m = n - 2 rows, composite), in place on dst
template <class C>
FK_DEV void xs_from_ortho(int n, int c0, const Mat& src, const Mat& dst, const TdmaTabs& T, double* red) {
  const int m = n - 2;
  const double2* PF = (const double2*)T.pf;  // slot: sd, sl | fs, fp
  const double* PB = T.pb;
  scan1n<C::LC, C::NTHR, C::CL, true, C::NTHR>(
      m, red,
      [&](int i, int l, int s) {
        const double2 a = __ldg(&PF[2 * s]);
        const double c = fma(a.x, gld(src, i, c0 + l), a.y * gld(src, i + 2, c0 + l));
        return __ldg(&PF[2 * s + 1]).x * c;
      },
      [&](int, int, int s) { return __ldg(&PF[2 * s + 1]).y; },
      [&](int i, int l, double y) {
        if (c0 + l < dst.cols) dst.p[(size_t)i * dst.ld + c0 + l] = y;
      });
  scan1n<C::LC, C::NTHR, C::CL, false, C::NTHR>(
      m, red, [&](int i, int l) { return gld(dst, i, c0 + l); }, [&](int, int, int s) { return __ldg(&PB[s]); },
      [&](int i, int l, double y) {
        if (c0 + l < dst.cols) dst.p[(size_t)i * dst.ld + c0 + l] = y;
      });
}
}  // namespace

// ---------------------------------------------------------------------------------
template <class C>
FK_DEV void xs_rhs_adi_body(const XsRhsAdiArgs& a) {
  RP_DYN_SMEM(double, red);
  const int c0 = blockIdx.x * C::LR;
  const int n = a.nx, m = n - 2;
  // rhs = -dt conv (already cut and scaled by the forward DCT kernel) + to_ortho(field) + explicit terms
  strip_rows<C>(n, c0, [&](int i, int c) {
    if (c >= a.rhs.cols) return;
    double v = gld(a.chat, i, c) + sxsy(a.fld, i, c, a.fxsd, a.fxsl, a.fysd, a.fysl);
    if (a.mode == 0) {
      v += gld(a.dxp, i, c);  // - dt/sx d/dx pres (navier.rs:627), prepared by xs_dxp
    } else if (a.mode == 1) {  // - dt/sy d/dy pres + dt (that + tbc)   (navier.rs:646-648)
      const double that = sxsy(a.tmp, i, c, a.txsd, a.txsl, a.tysd, a.tysl) + gld(a.tbc, i, c);
      v = fma(a.dt, that, fma(-a.dt, gld(a.dyp, i, c), v));
    } else {
      v += gld(a.bcdiff, i, c);  // + dt ka (dxx + dyy) fieldbc (navier.rs:665-668)
    }
    a.rhs.p[(size_t)i * a.rhs.ld + c] = v;
  });
  __syncthreads();
  // B2_x matvec fused into the forward sweep, then the backward sweep (fdma.rs:101-118), in place on out
  const double2* P1 = (const double2*)a.pt1;  // slot: lo, di | up, fp
  const double2* P2 = (const double2*)a.pt2;  // slot: bs, bp1 | bp2, -
  scan1n<C::LC, C::NTHR, C::CL, true, C::NTHR>(
      m, red,
      [&](int i, int l, int s) {
        const double2 p = __ldg(&P1[2 * s]), q = __ldg(&P1[2 * s + 1]);
        const int c = c0 + l;
        return fma(p.x, gld(a.rhs, i, c), fma(p.y, gld(a.rhs, i + 2, c), (i + 4 < n) ? q.x * gld(a.rhs, i + 4, c) : 0.0));
      },
      [&](int, int, int s) { return __ldg(&P1[2 * s + 1]).y; },
      [&](int i, int l, double y) {
        if (c0 + l < a.out.cols) a.out.p[(size_t)i * a.out.ld + c0 + l] = y;
      });
  scan2n<C::LC, C::NTHR, C::CL, false, C::NTHR>(
      m, red, [&](int i, int l, int s) { return __ldg(&P2[2 * s]).x * gld(a.out, i, c0 + l); },
      [&](int, int, int s) { return __ldg(&P2[2 * s]).y; }, [&](int, int, int s) { return __ldg(&P2[2 * s + 1]).x; },
      [&](int i, int l, double y) {
        if (c0 + l < a.out.cols) a.out.p[(size_t)i * a.out.ld + c0 + l] = y;
      });
}
// blockIdx.y selects the field (block-uniform branch keeps the arguments direct constant-bank operands)
template <class C>
__global__ void __launch_bounds__(C::NTHR, C::MINB) xs_rhs_adi(XsRhsAdiArgs3 a3) {
  if (blockIdx.y == 0)
    xs_rhs_adi_body<C>(a3.a[0]);
  else if (blockIdx.y == 1)
    xs_rhs_adi_body<C>(a3.a[1]);
  else
    xs_rhs_adi_body<C>(a3.a[2]);
}

// div = D_x S_x vx / sx + S_x ey ; r1 = B2_x div
template <class C>
__global__ void __launch_bounds__(C::NTHR, C::MINB) xs_div(XDivArgs a) {
  RP_DYN_SMEM(double, red);
  const int c0 = blockIdx.x * C::LR;
  const int n = a.nx, m = n - 2;
  auto sx = [&](const Mat& f, int i, int c) {
    double v = 0.0;
    if (i < m) v = __ldg(&a.sd[i]) * gld(f, i, c);
    if (i >= 2) v = fma(__ldg(&a.sl[i - 2]), gld(f, i - 2, c), v);
    return v;
  };
  xs_cheb_diff<C>(n, c0, a.div, a.isx, red, [&](int i, int c) { return sx(a.vx, i, c); });
  strip_rows<C>(n, c0, [&](int i, int c) {
    if (c < a.div.cols) a.div.p[(size_t)i * a.div.ld + c] += sx(a.ey, i, c);
  });
  __syncthreads();
  strip_rows<C>(m, c0, [&](int i, int c) {
    if (c >= a.r1.cols) return;
    const double up = (i + 4 < n) ? __ldg(&a.b2.up[i]) * gld(a.div, i + 4, c) : 0.0;
    a.r1.p[(size_t)i * a.r1.ld + c] = fma(__ldg(&a.b2.lo[i]), gld(a.div, i, c), fma(__ldg(&a.b2.di[i]), gld(a.div, i + 2, c), up));
  });
}

// a1 = from_ortho_x(D_x S_x phi)/sx, a2 = from_ortho_x(S_x phi); p / d are [nx, cols] scratch arrays
template <class C>
__global__ void __launch_bounds__(C::NTHR, C::MINB) xs_project(XsProjectArgs a) {
  RP_DYN_SMEM(double, red);
  const int c0 = blockIdx.x * C::LR;
  const int n = a.nx, m = n - 2;
  strip_rows<C>(n, c0, [&](int i, int c) {
    if (c >= a.p.cols) return;
    double v = 0.0;
    if (i < m) v = __ldg(&a.nsd[i]) * gld(a.phi, i, c);
    if (i >= 2) v = fma(__ldg(&a.nsl[i - 2]), gld(a.phi, i - 2, c), v);
    a.p.p[(size_t)i * a.p.ld + c] = v;
  });
  __syncthreads();
  xs_cheb_diff<C>(n, c0, a.d, a.isx, red, [&](int i, int c) { return gld(a.p, i, c); });
  xs_from_ortho<C>(n, c0, a.d, a.a1, a.t, red);
  xs_from_ortho<C>(n, c0, a.p, a.a2, a.t, red);
}

// dst = sc * D_x src along x (both [nx, cols]): -dt/sx d/dx pres of navier.rs:627, once per step
template <class C>
__global__ void __launch_bounds__(C::NTHR, C::MINB) xs_dxp(XsDiffArgs a) {
  RP_DYN_SMEM(double, red);
  const int c0 = blockIdx.x * C::LR;
  xs_cheb_diff<C>(a.nx, c0, a.dst, a.sc, red, [&](int i, int c) { return gld(a.src, i, c); });
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
bool xs_supported(int n0) { return n0 >= 8 && n0 <= XS_NMAX; }
// RUSTPDE_B200_XS_CFG=1: 4 columns x 512 threads, two blocks per SM (default: 8 columns x 1024 threads)
static int xs_cfg() {
  static const char* e = getenv("RUSTPDE_B200_XS_CFG");
  return e ? atoi(e) : 0;
}
ScanShape xs_scan_shape() {
  if (xs_cfg() == 1) return ScanShape{XsCfg<2, 512>::NG, XsCfg<2, 512>::CL};
  return ScanShape{XsCfg<4, 1024>::NG, XsCfg<4, 1024>::CL};
}

template <class K>
static void xs_prepare(K kern, int bytes) {
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
#else
  (void)kern;
  (void)bytes;
#endif
}
#define XS_LAUNCH_C(C, kern, ncols, nby, args)                                                        \
  do {                                                                                                \
    static unsigned long long init_ = 0; /* one bit per device */                                     \
    if (first_use_on_device(init_)) xs_prepare(kern<C>, C::RED);                                      \
    RP_LAUNCH(kern<C>, dim3(((ncols) + C::LR - 1) / C::LR, (nby)), dim3(C::NTHR), (size_t)C::RED, s, args); \
  } while (0)
#define XS_LAUNCH(kern, ncols, nby, args)                          \
  do {                                                             \
    typedef XsCfg<4, 1024> C0_;                                    \
    typedef XsCfg<2, 512> C1_;                                     \
    if (xs_cfg() == 1)                                             \
      XS_LAUNCH_C(C1_, kern, ncols, nby, args);                    \
    else                                                           \
      XS_LAUNCH_C(C0_, kern, ncols, nby, args);                    \
  } while (0)

void launch_xs_rhs_adi(const XsRhsAdiArgs3& a, int nb, cudaStream_t s) { XS_LAUNCH(xs_rhs_adi, a.a[0].rhs.cols, nb, a); }
void launch_xs_div(const XDivArgs& a, cudaStream_t s) { XS_LAUNCH(xs_div, a.div.cols, 1, a); }
void launch_xs_project(const XsProjectArgs& a, cudaStream_t s) { XS_LAUNCH(xs_project, a.phi.cols, 1, a); }
void launch_xs_dxp(const XsDiffArgs& a, cudaStream_t s) { XS_LAUNCH(xs_dxp, a.src.cols, 1, a); }

}  // namespace fk
}  // namespace rp
