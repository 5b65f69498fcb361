// fast_x.cu -- specialised x-pass kernels of the confined Navier2D::update (lanes
// strided in memory; a block owns 4 adjacent columns).  The x axis of the
// confined configurations has n0 = 2^k points, i.e. an odd DCT-I period, so the
// DCT goes through Bluestein with a power-of-two FFT (fast.cuh dct_bluestein).
//
//   xk_backward : B_x S_x u^ and B_x D_x S_x u^ / sx       (composite.rs:480-506, ortho.rs:107-125)
//   xk_forward  : forward DCT-x + dealias (navier.rs:1022-1032) + rhs assembly
//                 (navier.rs:622-674) + x half of HholtzAdi (hholtz_adi.rs:108,128)
//   xk_div      : divergence (navier.rs:698-703) + B2_x of the Poisson rhs
//   xk_project  : x part of u -= from_ortho(grad phi) (navier.rs:683-695)
#include "fast.cuh"

namespace rp {
namespace fk {


// LC = 2: 4 columns per block (one block per SM with the 128 KB Bluestein tile at LB = 4096);
// LC = 1: 2 columns per block, half the shared memory, two blocks per SM.
template <int LOG2LB, int LC_>
struct XCfg {
  static constexpr int LC = LC_, LR = 2 * LC_;
  static constexpr int LB = 1 << LOG2LB;
  static constexpr int NMAX = LB / 2 + 1;  // 2 (n - 1) - 1 <= LB
  static constexpr int NTHR = (LB * LC / 16) < 64 ? 64 : (LB * LC / 16);
  static constexpr int AROWS = LB / 2 + 8;
  static constexpr int CL = chunk_len(NMAX, NTHR, LC);
  static constexpr int RED = scan_threads(NTHR) * 56 + 512;
  static constexpr int SMEM_A = AROWS * LR * 8 + RED;
  static constexpr int SMEM_AA = 2 * AROWS * LR * 8 + RED;
  static constexpr int SMEM_AW = (AROWS + LB) * LR * 8 + RED;
  static constexpr int MINB = SMEM_AW > 110 * 1024 ? 1 : 2;
};

#define FK_FILL_U 8
template <int LC, int NTHR, class F>
FK_DEV void xtile_fill(double* td, int nfill, F f) {  // batched like tile_fill (fast_y.cu)
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

// composite -> ortho stencil along x applied while loading column c of `a`
// (m = n-2 rows): p_i = d_i c_i + l_{i-2} c_{i-2}   (composite_stencil.rs:207-229)
// All loads are unconditional (clamped indices, zero weights) so that the compiler
// can issue a whole batch of them before the first use.
FK_DEV double ld_stencil_x(const Mat& a, int i, int c, const double* __restrict__ sd, const double* __restrict__ sl) {
  const int cc = min(c, a.cols - 1), i0 = min(i, a.rows - 1), i2 = max(i - 2, 0);
  const double v0 = a.p[(size_t)i0 * a.ld + cc], v2 = a.p[(size_t)i2 * a.ld + cc];
  const bool ok = c < a.cols;
  const double d = (ok && i < a.rows) ? __ldg(&sd[i0]) : 0.0;
  const double l = (ok && i >= 2) ? __ldg(&sl[i2]) : 0.0;
  return fma(l, v2, d * v0);
}
// plain element (i, c) of a, zero outside
FK_DEV double ld_plain(const Mat& a, int i, int c) {
  const double v = a.p[(size_t)min(i, a.rows - 1) * a.ld + min(c, a.cols - 1)];
  return (i < a.rows && c < a.cols) ? v : 0.0;
}
// S_x S_y f at ortho index (i, j); f is [mx, my]
FK_DEV double ld_stencil_xy(const Mat& f, int i, int j, const double* __restrict__ xsd, const double* __restrict__ xsl,
                            const double* __restrict__ ysd, const double* __restrict__ ysl) {
  const int i0 = min(i, f.rows - 1), i2 = max(i - 2, 0), j0 = min(j, f.cols - 1), j2 = max(j - 2, 0);
  const double* r0 = f.p + (size_t)i0 * f.ld;
  const double* r2 = f.p + (size_t)i2 * f.ld;
  const double v00 = r0[j0], v02 = r0[j2], v20 = r2[j0], v22 = r2[j2];
  const double yd = (j < f.cols) ? __ldg(&ysd[j0]) : 0.0, yl = (j >= 2) ? __ldg(&ysl[j2]) : 0.0;
  const double xd = (i < f.rows) ? __ldg(&xsd[i0]) : 0.0, xl = (i >= 2) ? __ldg(&xsl[i2]) : 0.0;
  const double t0 = fma(yl, v02, yd * v00), t2 = fma(yl, v22, yd * v20);
  return fma(xl, t2, xd * t0);
}

// dst(i) = d_i src(i) + l_{i-2} src(i-2), i < n: composite (m = n-2 rows of src) -> ortho along the tile axis
// (composite_stencil.rs:207-229).  A 4-column strip costs one L1 line per row and array, so every
// array is read from global memory once and the stencil runs on the shared-memory copy.
template <int LC, int NTHR>
FK_DEV void xstencil_tile(double* dst, const double* src, int n, const double* __restrict__ sd, const double* __restrict__ sl) {
  constexpr int LR = 2 * LC;
  const int m = n - 2;
  for (int it = threadIdx.x; it < n * LR; it += NTHR) {
    const int l = it % LR, i = it / LR;
    double v = 0.0;
    if (i < m) v = __ldg(&sd[i]) * src[didx<LC>(i, l)];
    if (i >= 2) v = fma(__ldg(&sl[i - 2]), src[didx<LC>(i - 2, l)], v);
    dst[didx<LC>(i, l)] = v;
  }
}

// ---------------------------------------------------------------------------------
template <int LOG2LB, int LC>
FK_DEV void xk_backward_body(const XBackwardArgs& a, const XBackwardArgs3& a3) {
  typedef XCfg<LOG2LB, LC> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta);
  double* tw = ta + C::AROWS * LR;
  double* red = tw + C::LB * LR;
  const int c0 = blockIdx.x * LR;
  const int n = a.t.n, N = n - 1;
  xtile_fill<LC, C::NTHR>(tw, n - 2, [&](int i, int l) { return ld_plain(a.src, i, c0 + l); });
  {  // first strip of the block that runs on this SM next
    const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a3.next_wave;
    if (nxt < (int)(gridDim.x * gridDim.y)) prefetch_strip<C::NTHR, false>(a3.a[nxt / gridDim.x].src, (nxt % gridDim.x) * LR, n - 2);
  }
  __syncthreads();
  xstencil_tile<LC, C::NTHR>(ta, tw, n, a.sd, a.sl);
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    const Mat& o = pass ? a.dx : a.val;
    if (pass) cheb_diff<LC, C::NTHR, C::CL>(ta, -1, ta, -1, n, a.isx, red);
    dct_bluestein<LC, LOG2LB, C::NTHR, true>(ta, tw, a.t, red);
    for (int it = threadIdx.x; it < n * LR; it += C::NTHR) {
      const int l = it % LR, i = it / LR;
      if (c0 + l < o.cols) o.p[(size_t)i * o.ld + c0 + l] = tw[didx<LC>(rowof(N, i), l)];
    }
    __syncthreads();
  }
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2LB, int LC>
__global__ void __launch_bounds__(XCfg<LOG2LB, LC>::NTHR, XCfg<LOG2LB, LC>::MINB) xk_backward(XBackwardArgs3 a3) {
  if (blockIdx.y == 0)
    xk_backward_body<LOG2LB, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    xk_backward_body<LOG2LB, LC>(a3.a[1], a3);
  else
    xk_backward_body<LOG2LB, LC>(a3.a[2], a3);
}

template <int LOG2LB, int LC>
FK_DEV void xk_forward_body(const XForwardArgs& a, const XForwardArgs3& a3) {
  typedef XCfg<LOG2LB, LC> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta);
  double* tw = ta + C::AROWS * LR;
  double* red = tw + C::LB * LR;
  const int c0 = blockIdx.x * LR;
  const int n = a.t.n, N = n - 1;
  const int ncols = a.conv.cols;
  xtile_fill<LC, C::NTHR>(ta, n, [&](int i, int l) { return ld_plain(a.conv, i, c0 + l); });
  {  // first strip of the block that runs on this SM next -> L2
    const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a3.next_wave;
    if (nxt < (int)(gridDim.x * gridDim.y)) prefetch_strip<C::NTHR, false>(a3.a[nxt / gridDim.x].conv, (nxt % gridDim.x) * LR, n);
  }
  __syncthreads();
  dct_bluestein<LC, LOG2LB, C::NTHR, false>(ta, tw, a.t, red);
  // rhs assembly in the split(N) layout of W.  Every global array is read once per element:
  // S_y is applied while loading (columns j, j-2), S_x on the shared-memory copy in A.
  const int mxr = n - 2;
  auto ld_sy = [&](const Mat& f, int i, int l, const double* __restrict__ ysd, const double* __restrict__ ysl) {
    const int j = c0 + l, j0 = min(j, f.cols - 1), j2 = min(max(j - 2, 0), f.cols - 1);
    const double* row = f.p + (size_t)i * f.ld;
    const double v0 = row[j0], v2 = row[j2];
    const double yd = (j < f.cols) ? __ldg(&ysd[j0]) : 0.0, yl = (j >= 2 && j - 2 < f.cols) ? __ldg(&ysl[j2]) : 0.0;
    return fma(yl, v2, yd * v0);
  };
  auto sx_at = [&](int i, int l, const double* __restrict__ xsd, const double* __restrict__ xsl) {
    double v = 0.0;
    if (i < mxr) v = __ldg(&xsd[i]) * ta[didx<LC>(i, l)];
    if (i >= 2) v = fma(__ldg(&xsl[i - 2]), ta[didx<LC>(i - 2, l)], v);
    return v;
  };
  // - dt * dealiased conv + to_ortho(field)   (navier.rs:625, 630, 651, 671)
  xtile_fill<LC, C::NTHR>(ta, mxr, [&](int i, int l) { return ld_sy(a.fld, i, l, a.fysd, a.fysl); });
  __syncthreads();
  for (int it0 = threadIdx.x; it0 < n * LR; it0 += C::NTHR * 4) {
    double add[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = min(it0 + u * C::NTHR, n * LR - 1);
      add[u] = (a.mode == 2) ? ld_plain(a.bcdiff, it / LR, c0 + (it % LR)) : 0.0;  // + dt ka (dxx + dyy) fieldbc (665-668)
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int it = it0 + u * C::NTHR;
      if (it < n * LR) {
        const int l = it % LR, i = it / LR;
        double* w = &tw[didx<LC>(rowof(N, i), l)];
        const double v = (i < a.cut) ? -a.dt * (*w) : 0.0;
        *w = v + sx_at(i, l, a.fxsd, a.fxsl) + add[u];
      }
    }
  }
  __syncthreads();
  if (a.mode == 0) {  // - dt/sx d/dx pres   (navier.rs:627)
    xtile_fill<LC, C::NTHR>(ta, n, [&](int i, int l) { return ld_plain(a.pres, i, c0 + l); });
    __syncthreads();
    cheb_diff<LC, C::NTHR, C::CL>(ta, -1, ta, -1, n, -a.dt * a.isx, red);
    for (int it = threadIdx.x; it < n * LR; it += C::NTHR) {
      const int l = it % LR, i = it / LR;
      tw[didx<LC>(rowof(N, i), l)] += ta[didx<LC>(i, l)];
    }
  } else if (a.mode == 1) {  // - dt/sy d/dy pres + dt * (that + tbc)   (navier.rs:646-648)
    xtile_fill<LC, C::NTHR>(ta, mxr, [&](int i, int l) { return ld_sy(a.tmp, i, l, a.tysd, a.tysl); });
    __syncthreads();
    for (int it0 = threadIdx.x; it0 < n * LR; it0 += C::NTHR * 4) {
      double g1[4], g2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = min(it0 + u * C::NTHR, n * LR - 1);
        g1[u] = ld_plain(a.dyp, it / LR, c0 + (it % LR));
        g2[u] = ld_plain(a.tbc, it / LR, c0 + (it % LR));
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int it = it0 + u * C::NTHR;
        if (it < n * LR) {
          const int l = it % LR, i = it / LR;
          const double that = sx_at(i, l, a.txsd, a.txsl) + g2[u];
          double* w = &tw[didx<LC>(rowof(N, i), l)];
          *w = fma(a.dt, that, fma(-a.dt, g1[u], *w));
        }
      }
    }
  }
  __syncthreads();
  b2_fdma<LC, C::NTHR, C::CL>(tw, N, n, a.b2, a.f, red);
  const int m = n - 2;
  for (int it = threadIdx.x; it < m * LR; it += C::NTHR) {
    const int l = it % LR, i = it / LR;
    if (c0 + l < ncols) a.out.p[(size_t)i * a.out.ld + c0 + l] = tw[didx<LC>(rowof(N, i), l)];
  }
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2LB, int LC>
__global__ void __launch_bounds__(XCfg<LOG2LB, LC>::NTHR, XCfg<LOG2LB, LC>::MINB) xk_forward(XForwardArgs3 a3) {
  if (blockIdx.y == 0)
    xk_forward_body<LOG2LB, LC>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    xk_forward_body<LOG2LB, LC>(a3.a[1], a3);
  else
    xk_forward_body<LOG2LB, LC>(a3.a[2], a3);
}

template <int LOG2LB, int LC>
__global__ void __launch_bounds__(XCfg<LOG2LB, LC>::NTHR, XCfg<LOG2LB, LC>::MINB) xk_div(XDivArgs a) {
  typedef XCfg<LOG2LB, LC> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta);
  double* red = ta + C::AROWS * LR;
  const int c0 = blockIdx.x * LR;
  const int n = a.nx, m = n - 2;
  const int ncols = a.vx.cols;
  double* tb = red + C::RED / 8;  // second tile (after the scan scratch)
  xtile_fill<LC, C::NTHR>(tb, m, [&](int i, int l) { return ld_plain(a.vx, i, c0 + l); });
  __syncthreads();
  xstencil_tile<LC, C::NTHR>(ta, tb, n, a.sd, a.sl);
  __syncthreads();
  xtile_fill<LC, C::NTHR>(tb, m, [&](int i, int l) { return ld_plain(a.ey, i, c0 + l); });
  cheb_diff<LC, C::NTHR, C::CL>(ta, -1, ta, -1, n, a.isx, red);
  for (int it = threadIdx.x; it < n * LR; it += C::NTHR) {
    const int l = it % LR, i = it / LR;
    double e = 0.0;
    if (i < m) e = __ldg(&a.sd[i]) * tb[didx<LC>(i, l)];
    if (i >= 2) e = fma(__ldg(&a.sl[i - 2]), tb[didx<LC>(i - 2, l)], e);
    const double v = ta[didx<LC>(i, l)] + e;
    ta[didx<LC>(i, l)] = v;
    if (c0 + l < ncols) a.div.p[(size_t)i * a.div.ld + c0 + l] = v;
  }
  __syncthreads();
  for (int it = threadIdx.x; it < m * LR; it += C::NTHR) {
    const int l = it % LR, i = it / LR;
    const double v = fma(__ldg(&a.b2.lo[i]), ta[didx<LC>(i, l)],
                         fma(__ldg(&a.b2.di[i]), ta[didx<LC>(i + 2, l)], (i + 4 < n) ? __ldg(&a.b2.up[i]) * ta[didx<LC>(i + 4, l)] : 0.0));
    if (c0 + l < ncols) a.r1.p[(size_t)i * a.r1.ld + c0 + l] = v;
  }
}

template <int LOG2LB, int LC>
__global__ void __launch_bounds__(XCfg<LOG2LB, LC>::NTHR, XCfg<LOG2LB, LC>::MINB) xk_project(XProjectArgs a) {
  typedef XCfg<LOG2LB, LC> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta);
  double* tb = ta + C::AROWS * LR;
  double* red = tb + C::AROWS * LR;
  const int c0 = blockIdx.x * LR;
  const int n = a.nx, m = n - 2;
  const int ncols = a.phi.cols;
  xtile_fill<LC, C::NTHR>(tb, m, [&](int i, int l) { return ld_plain(a.phi, i, c0 + l); });
  __syncthreads();
  xstencil_tile<LC, C::NTHR>(ta, tb, n, a.nsd, a.nsl);
  __syncthreads();
  cheb_diff<LC, C::NTHR, C::CL>(ta, -1, tb, -1, n, a.isx, red);
  from_ortho<LC, C::NTHR, C::CL>(tb, -1, n, a.t, red);
  from_ortho<LC, C::NTHR, C::CL>(ta, -1, n, a.t, red);
  for (int it = threadIdx.x; it < m * LR; it += C::NTHR) {
    const int l = it % LR, i = it / LR;
    if (c0 + l < ncols) {
      a.a1.p[(size_t)i * a.a1.ld + c0 + l] = tb[didx<LC>(i, l)];
      a.a2.p[(size_t)i * a.a2.ld + c0 + l] = ta[didx<LC>(i, l)];
    }
  }
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
// (log2 Bluestein length, complex lanes per tile); the 2-lane variants of the large sizes are selected with
// RUSTPDE_B200_XLC=1 (two blocks per SM instead of one)
#define XK_SIZES(X) X(6, 2) X(7, 2) X(11, 2) X(12, 2) X(11, 1) X(12, 1)

static int bluestein_log2(int n0) {  // tables.cu: Lb = next_pow2(2N - 1)
  const int need = 2 * (n0 - 1) - 1;
  int l = 0;
  while ((1 << l) < need) ++l;
  return l;
}

bool x_supported(int n0) {
  if (n0 < 8) return false;
  const int N = n0 - 1;
  if ((N & (N - 1)) == 0) return false;  // power-of-two period: the tables hold no chirp
  const int l = bluestein_log2(n0);
#define X(L, LCV) \
  if (l == L) return true;
  XK_SIZES(X)
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

#define XK_CASE_BODY(kern, L, LCV, smem_expr)                                    \
  if (!ok_ && l_ == L && (LCV == lc_ || L < 11)) {                               \
    typedef XCfg<L, LCV> C;                                                      \
    auto kp_ = kern<L, LCV>;                                                     \
    const int sm_ = (smem_expr);                                                 \
    const int nb_ = ((ncols_) + C::LR - 1) / C::LR;                              \
    static bool init_ = false;                                                   \
    if (!init_) {                                                                \
      set_smem(kp_, sm_);                                                        \
      init_ = true;                                                              \
    }                                                                            \
    RP_LAUNCH(kp_, dim3(nb_, nby_), dim3(C::NTHR), (size_t)sm_, s, a);           \
    ok_ = true;                                                                  \
  }
#define XK_CASE_xk_backward(L, LCV) XK_CASE_BODY(xk_backward, L, LCV, C::SMEM_AW)
#define XK_CASE_xk_forward(L, LCV) XK_CASE_BODY(xk_forward, L, LCV, C::SMEM_AW)
#define XK_CASE_xk_div(L, LCV) XK_CASE_BODY(xk_div, L, LCV, C::SMEM_AA)
#define XK_CASE_xk_project(L, LCV) XK_CASE_BODY(xk_project, L, LCV, C::SMEM_AA)

#define XK_LAUNCH(kern, ncols, nx, nby)                                            \
  do {                                                                             \
    const int nby_ = (nby);                                                        \
    const int l_ = bluestein_log2(nx);                                             \
    const int ncols_ = (ncols);                                                    \
    const int lc_ = x_lanes();                                                     \
    bool ok_ = false;                                                              \
    XK_SIZES(XK_CASE_##kern)                                                       \
    if (!ok_) throw Error(RP_ERR_INTERNAL, #kern ": unsupported lane length");     \
  } while (0)

static int x_lanes() {  // complex lanes per tile of the large Bluestein kernels (RUSTPDE_B200_XLC=1|2)
  static int lc = 0;
  if (!lc) {
    const char* e = getenv("RUSTPDE_B200_XLC");
    lc = (e && e[0] == '1') ? 1 : 2;
  }
  return lc;
}
static int sm_count() {
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
void launch_x_backward(const XBackwardArgs3& a_, int nb, cudaStream_t s) {
  XBackwardArgs3 a = a_;
  a.next_wave = sm_count() * (x_lanes() == 1 ? 2 : 1);  // the block that follows on the same SM is about one wave ahead
  XK_LAUNCH(xk_backward, a.a[0].src.cols, a.a[0].t.n, nb);
}
void launch_x_forward(const XForwardArgs3& a_, int nb, cudaStream_t s) {
  XForwardArgs3 a = a_;
  a.next_wave = sm_count() * (x_lanes() == 1 ? 2 : 1);
  XK_LAUNCH(xk_forward, a.a[0].conv.cols, a.a[0].t.n, nb);
}
void launch_x_div(const XDivArgs& a, cudaStream_t s) { XK_LAUNCH(xk_div, a.vx.cols, a.nx, 1); }
void launch_x_project(const XProjectArgs& a, cudaStream_t s) { XK_LAUNCH(xk_project, a.phi.cols, a.nx, 1); }

}  // namespace fk
}  // namespace rp
