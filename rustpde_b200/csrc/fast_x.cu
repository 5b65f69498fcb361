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
//
// Column strips are strided in memory (one 32-byte sector per row), so every array is read exactly once per
// element, 16 bytes per thread; xk_forward stages the strips of its later phases asynchronously (cp.async) into
// tiles that are idle at the time (xstage / xstencil_cols).
#include "fast.cuh"

namespace rp {
namespace fk {


// A block owns 4 adjacent columns = 2 packed complex lanes (LC = 2); with the 128 KB Bluestein tile at
// LB = 4096 that is one block per SM.  All tile traffic is in 16-byte units (one complex lane = two columns).
template <int LOG2LB>
struct XCfg {
  static constexpr int LC = 2, LR = 4;
  static constexpr int LB = 1 << LOG2LB;
  static constexpr int NMAX = LB / 2 + 1;  // 2 (n - 1) - 1 <= LB
  static constexpr int NTHR = (LB * LC / 16) < 64 ? 64 : (LB * LC / 16);
  static constexpr int AROWS = LB / 2 + 8;
  static constexpr int NSC = scan_threads(NTHR);
  static constexpr int RED = NSC * 56 + 1024;  // >= scan1v_bytes(NSC), scan2v_bytes(NSC / 2)
  static constexpr int ABYTES = AROWS * LR * 8;
  // the second-order sweep of xk_forward borrows the (then idle) A tile as scratch when it is large enough
  static constexpr int NSC2 = scan2v_bytes(NSC) <= ABYTES ? NSC : NSC / 2;
  static constexpr int SMEM_AA = 2 * ABYTES + RED;
  // W tile: Lb rows of the Bluestein FFT, plus room for a strip of n <= Lb / 2 + 1 rows staged behind the result of a
  // power-of-two DCT, which occupies rows [0, Lb / 2 + 4) (xk_forward)
  static constexpr int WROWS = LB + 16;
  static constexpr int SMEM_AW = (AROWS + WROWS) * LR * 8 + RED;
  static_assert(SMEM_AW <= 227 * 1024, "shared memory of the x kernels");
  static constexpr int MINB = SMEM_AW > 110 * 1024 ? 1 : 2;
  static_assert(scan1v_bytes(NSC) <= RED && scan2v_bytes(NSC / 2) <= RED, "scan scratch");
};

#define FK_FILL_U 8
// tile rows [0, nfill): tc[i][c] = f(i), c = the calling thread's complex lane (threadIdx.x % 2).  The loads of a
// batch are all issued before the first store.
template <int NTHR, class F>
FK_DEV void xfill(cplx* tc, int nfill, F f) {
  const int c = threadIdx.x & 1;
  for (int i0 = threadIdx.x >> 1; i0 < nfill; i0 += (NTHR / 2) * FK_FILL_U) {
    cplx v[FK_FILL_U];
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) v[u] = f(min(i0 + u * (NTHR / 2), nfill - 1));  // clamped: unconditional loads
#pragma unroll
    for (int u = 0; u < FK_FILL_U; ++u) {
      const int i = i0 + u * (NTHR / 2);
      if (i < nfill) tc[cidx<2>(i, c)] = v[u];
    }
  }
}
// L2 prefetch of rows [0, nrows) of the 4-column strip at c0 (plus the two columns before it with HALO)
template <int NTHR, bool HALO>
FK_DEV void xprefetch(const Mat& a, int c0, int nrows) {
  if (a.p == nullptr || c0 >= a.cols) return;
  nrows = min(nrows, a.rows);
  for (int i = threadIdx.x; i < nrows; i += NTHR) {
    const double* row = a.p + (size_t)i * a.ld;
    prefetch_l2(row + c0);
    if (HALO && c0 >= 2) prefetch_l2(row + c0 - 2);
  }
}

// dst(i) = d_i src(i) + l_{i-2} src(i-2), i < n: composite (m = n-2 rows of src) -> ortho along the tile axis
// (composite_stencil.rs:207-229)
template <int NTHR>
FK_DEV void xstencil_tile(cplx* dst, const cplx* src, int n, const double* __restrict__ sd, const double* __restrict__ sl) {
  const int m = n - 2, c = threadIdx.x & 1;
  for (int i = threadIdx.x >> 1; i < n; i += NTHR / 2) {
    cplx v = mk(0.0, 0.0);
    if (i < m) v = cscale(src[cidx<2>(i, c)], __ldg(&sd[i]));
    if (i >= 2) v = sfma(__ldg(&sl[i - 2]), src[cidx<2>(i - 2, c)], v);
    dst[cidx<2>(i, c)] = v;
  }
}

// DCT-I along the x tile: tile A (natural layout, n rows) -> tile W (split(N) layout); A is left untouched and idle
// from `after_pre()` on.  Odd periods N = n - 1 (n0 = 2^k, every BASELINE configuration) go through Bluestein; a
// power-of-two period (n0 = 2^k + 1: the tables hold no chirp) is copied to W and transformed in place by the y
// kernels' dct_pow2 -- N = Lb / 2, the same thread count and tile rows (P2 instantiations of the kernels, chosen at launch).
template <int LOG2LB, int NTHR, bool BWD, bool P2, class Hook>
FK_DEV void dct_x(const cplx* ta, cplx* tw, const DctTab& T, double* red, Hook after_pre) {
  if constexpr (!P2) {
    dct_bluestein<2, LOG2LB, NTHR, BWD>((const double*)ta, (double*)tw, T, red, after_pre);
  } else {
    const int c = threadIdx.x & 1;
    for (int i = threadIdx.x >> 1; i < T.n; i += NTHR / 2) tw[cidx<2>(i, c)] = ta[cidx<2>(i, c)];
    __syncthreads();
    after_pre();
    dct_pow2<2, LOG2LB - 1, NTHR, BWD>((double*)tw, T, red);
  }
}
template <int LOG2LB, int NTHR, bool BWD, bool P2>
FK_DEV void dct_x(const cplx* ta, cplx* tw, const DctTab& T, double* red) {
  dct_x<LOG2LB, NTHR, BWD, P2>(ta, tw, T, red, [] {});
}

// ---------------------------------------------------------------------------------
template <int LOG2LB, bool P2>
FK_DEV void xk_backward_body(const XBackwardArgs& a, const XBackwardArgs3& a3) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  cplx* tw = ta + C::AROWS * 2;
  double* red = (double*)(tw + C::WROWS * 2);
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.t.n, N = n - 1;
  xfill<C::NTHR>(tw, n - 2, [&](int i) { return ld2(a.src, i, col); });
  {  // first strip of the block that runs on this SM next
    const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a3.next_wave;
    if (nxt < (int)(gridDim.x * gridDim.y)) xprefetch<C::NTHR, false>(a3.a[nxt / gridDim.x].src, (nxt % gridDim.x) * LR, n - 2);
  }
  __syncthreads();
  xstencil_tile<C::NTHR>(ta, tw, n, a.sd, a.sl);
  __syncthreads();
  for (int pass = 0; pass < 2; ++pass) {
    const Mat& o = pass ? a.dx : a.val;
    if (pass) cheb_diff_v<2, C::NTHR, C::NMAX>(ta, -1, ta, -1, n, a.isx, red);
    dct_x<LOG2LB, C::NTHR, true, P2>(ta, tw, a.t, red);
    for (int i = threadIdx.x >> 1; i < n; i += C::NTHR / 2) st2(o, i, col, tw[cidx<2>(rowof(N, i), c)]);
    __syncthreads();
  }
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2LB, bool P2>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_backward(XBackwardArgs3 a3) {
  if (blockIdx.y == 0)
    xk_backward_body<LOG2LB, P2>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    xk_backward_body<LOG2LB, P2>(a3.a[1], a3);
  else
    xk_backward_body<LOG2LB, P2>(a3.a[2], a3);
}

// Asynchronous staging of a 4-column strip (cp.async, 16 bytes per thread and row): dst[i][c] = (f[i][col], f[i][col+1])
// for i < nrows, zero outside the matrix.  One commit group per call.
template <int NTHR>
FK_DEV void xstage(cplx* dst, const Mat& f, int nrows, int col) {
  const int c = threadIdx.x & 1;
  const int nb = (col + 1 < f.cols) ? 16 : (col < f.cols ? 8 : 0);
  const double* base = f.p + (nb ? col : 0);
  for (int i = threadIdx.x >> 1; i < nrows; i += NTHR / 2) {
    const bool r = i < f.rows;
    cp_async16(&dst[cidx<2>(i, c)], base + (size_t)(r ? i : 0) * f.ld, r ? nb : 0);
  }
  cp_async_commit();
}
// composite -> ortho stencil across the columns of a staged strip, in place: t(col) = d(col) t(col) + l(col-2) t(col-2).
// Columns c0-2, c0-1 belong to the neighbouring strip and come from global memory (an L2 hit: that strip is being
// worked on by the neighbouring block); columns c0, c0+1 for the second lane pair come from the first one by shuffle.
template <int NTHR>
FK_DEV void xstencil_cols(cplx* t, const Mat& f, int nrows, int col, const StencilPair& sp) {
  const int c = threadIdx.x & 1;
  const bool left = c == 0 && col >= 2 && col - 2 < f.cols;
  for (int i0 = 0; i0 < nrows; i0 += NTHR / 2) {
    const int i = i0 + (threadIdx.x >> 1);
    const bool ok = i < nrows && i < f.rows;
    const cplx v0 = ok ? t[cidx<2>(i, c)] : mk(0.0, 0.0);
    cplx v2 = mk(__shfl_up_sync(0xffffffffu, v0.x, 1), __shfl_up_sync(0xffffffffu, v0.y, 1));
    if (c == 0) v2 = (left && ok) ? *(const cplx*)(f.p + (size_t)i * f.ld + col - 2) : mk(0.0, 0.0);
    if (left && col - 1 >= f.cols) v2.y = 0.0;
    if (i < nrows) t[cidx<2>(i, c)] = mk(fma(sp.l.x, v2.x, sp.d.x * v0.x), fma(sp.l.y, v2.y, sp.d.y * v0.y));
  }
}

template <int LOG2LB, bool P2>
FK_DEV void xk_forward_body(const XForwardArgs& a, const XForwardArgs3& a3) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  cplx* tw = ta + C::AROWS * 2;
  cplx* tu = tw + C::AROWS * 2;  // rows [LB/2 + 8, ..) of W: idle once the inverse transform is done
  double* red = (double*)(tw + C::WROWS * 2);
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.t.n, N = n - 1;
  const int mxr = n - 2;
  xfill<C::NTHR>(ta, n, [&](int i) { return ld2(a.conv, i, col); });
  {  // first strip of the block that runs on this SM next -> L2
    const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a3.next_wave;
    if (nxt < (int)(gridDim.x * gridDim.y)) xprefetch<C::NTHR, false>(a3.a[nxt / gridDim.x].conv, (nxt % gridDim.x) * LR, n);
  }
  __syncthreads();
  // The strips of the later phases are staged asynchronously (cp.async) into tiles that are idle at the time: the
  // old field into A under the FFT stages, pres / temp into the upper half of W under the rhs assembly.
  dct_x<LOG2LB, C::NTHR, false, P2>(ta, tw, a.t, red, [&] { xstage<C::NTHR>(ta, a.fld, mxr, col); });
  if (a.mode == 0)
    xstage<C::NTHR>(tu, a.pres, n, col);
  else if (a.mode == 1)
    xstage<C::NTHR>(tu, a.tmp, mxr, col);
  else
    cp_async_commit();  // (keeps the group count uniform)
  cp_async_wait<1>();   // the old field has landed
  __syncthreads();
  // rhs assembly in the split(N) layout of W.  Every global array is read once per element: S_y across the
  // columns of the staged strip, S_x along it.
  auto sx_at = [&](const cplx* t, int i, const double* __restrict__ xsd, const double* __restrict__ xsl) {
    cplx v = mk(0.0, 0.0);
    if (i < mxr) v = cscale(t[cidx<2>(i, c)], __ldg(&xsd[i]));
    if (i >= 2) v = sfma(__ldg(&xsl[i - 2]), t[cidx<2>(i - 2, c)], v);
    return v;
  };
  // - dt * dealiased conv + to_ortho(field)   (navier.rs:625, 630, 651, 671)
  xstencil_cols<C::NTHR>(ta, a.fld, mxr, col, stencil_pair(a.fld.cols, col, a.fysd, a.fysl));
  __syncthreads();
  for (int i0 = threadIdx.x >> 1; i0 < n; i0 += (C::NTHR / 2) * 4) {
    cplx add[4];
#pragma unroll
    for (int u = 0; u < 4; ++u)  // + dt ka (dxx + dyy) fieldbc (665-668)
      add[u] = (a.mode == 2 && i0 + u * (C::NTHR / 2) < a.bcdiff_rows) ? ld2(a.bcdiff, min(i0 + u * (C::NTHR / 2), n - 1), col) : mk(0.0, 0.0);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      const int i = i0 + u * (C::NTHR / 2);
      if (i < n) {
        cplx* w = &tw[cidx<2>(rowof(N, i), c)];
        const cplx v = (i < a.cut) ? cscale(*w, -a.dt) : mk(0.0, 0.0);
        *w = cadd(cadd(v, sx_at(ta, i, a.fxsd, a.fxsl)), add[u]);
      }
    }
  }
  cp_async_wait<0>();
  __syncthreads();
  if (a.mode == 0) {  // - dt/sx d/dx pres   (navier.rs:627)
    cheb_diff_v<2, C::NTHR, C::NMAX>(tu, -1, tu, -1, n, -a.dt * a.isx, red);
    for (int i = threadIdx.x >> 1; i < n; i += C::NTHR / 2) {
      cplx* w = &tw[cidx<2>(rowof(N, i), c)];
      *w = cadd(*w, tu[cidx<2>(i, c)]);
    }
  } else if (a.mode == 1) {  // - dt/sy d/dy pres + dt * (that + tbc)   (navier.rs:646-648)
    xstencil_cols<C::NTHR>(tu, a.tmp, mxr, col, stencil_pair(a.tmp.cols, col, a.tysd, a.tysl));
    __syncthreads();
    for (int i0 = threadIdx.x >> 1; i0 < n; i0 += (C::NTHR / 2) * 4) {
      cplx g1[4], g2[4];
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = min(i0 + u * (C::NTHR / 2), n - 1);
        g1[u] = ld2(a.dyp, i, col);
        g2[u] = (i < a.tbc_rows) ? ld2(a.tbc, i, col) : mk(0.0, 0.0);
      }
#pragma unroll
      for (int u = 0; u < 4; ++u) {
        const int i = i0 + u * (C::NTHR / 2);
        if (i < n) {
          const cplx that = cadd(sx_at(tu, i, a.txsd, a.txsl), g2[u]);
          cplx* w = &tw[cidx<2>(rowof(N, i), c)];
          *w = sfma(a.dt, that, sfma(-a.dt, g1[u], *w));
        }
      }
    }
  }
  __syncthreads();
  if (a.rhs.p != nullptr) {  // the sweeps run as warp-serial column sweeps (fast_xw.cu)
    for (int i = threadIdx.x >> 1; i < n; i += C::NTHR / 2) st2(a.rhs, i, col, tw[cidx<2>(rowof(N, i), c)]);
    return;
  }
  // the A tile is idle from here on: scratch of the second-order sweep when it is large enough
  b2_fdma_v<2, C::NTHR, C::NMAX, C::NSC2>(tw, N, n, a.pt1, a.pt2, red, C::NSC2 == C::NSC ? (double*)ta : red);
  for (int i = threadIdx.x >> 1; i < mxr; i += C::NTHR / 2) st2(a.out, i, col, tw[cidx<2>(rowof(N, i), c)]);
}
// blockIdx.y selects the field; the branch is block-uniform and keeps every argument a direct constant-bank
// operand (a dynamically indexed `a3.a[blockIdx.y]` costs an LDC per access)
template <int LOG2LB, bool P2>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_forward(XForwardArgs3 a3) {
  if (blockIdx.y == 0)
    xk_forward_body<LOG2LB, P2>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    xk_forward_body<LOG2LB, P2>(a3.a[1], a3);
  else
    xk_forward_body<LOG2LB, P2>(a3.a[2], a3);
}

// forward DCT-x + dealias cut + scale only: the rhs assembly and the x sweeps run as streaming column scans (fast_xs.cu)
template <int LOG2LB, bool P2>
FK_DEV void xk_fdct_body(const XFdctArgs& a, const XFdctArgs3& a3) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  cplx* tw = ta + C::AROWS * 2;
  double* red = (double*)(tw + C::WROWS * 2);
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.t.n, N = n - 1;
  xfill<C::NTHR>(ta, n, [&](int i) { return ld2(a.conv, i, col); });
  {  // first strip of the block that runs on this SM next -> L2
    const int nxt = blockIdx.y * gridDim.x + blockIdx.x + a3.next_wave;
    if (nxt < (int)(gridDim.x * gridDim.y)) xprefetch<C::NTHR, false>(a3.a[nxt / gridDim.x].conv, (nxt % gridDim.x) * LR, n);
  }
  __syncthreads();
  dct_x<LOG2LB, C::NTHR, false, P2>(ta, tw, a.t, red);
  for (int i = threadIdx.x >> 1; i < n; i += C::NTHR / 2)
    st2(a.out, i, col, (i < a.cut) ? cscale(tw[cidx<2>(rowof(N, i), c)], a.scale) : mk(0.0, 0.0));
}
template <int LOG2LB, bool P2>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_fdct(XFdctArgs3 a3) {
  if (blockIdx.y == 0)
    xk_fdct_body<LOG2LB, P2>(a3.a[0], a3);
  else if (blockIdx.y == 1)
    xk_fdct_body<LOG2LB, P2>(a3.a[1], a3);
  else
    xk_fdct_body<LOG2LB, P2>(a3.a[2], a3);
}

template <int LOG2LB>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_div(XDivArgs a) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  double* red = (double*)(ta + C::AROWS * 2);
  cplx* tb = (cplx*)(red + C::RED / 8);  // second tile (after the scan scratch)
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.nx, m = n - 2;
  xprefetch<C::NTHR, false>(a.ey, c0, m);
  xfill<C::NTHR>(tb, m, [&](int i) { return ld2(a.vx, i, col); });
  __syncthreads();
  xstencil_tile<C::NTHR>(ta, tb, n, a.sd, a.sl);
  __syncthreads();
  xfill<C::NTHR>(tb, m, [&](int i) { return ld2(a.ey, i, col); });
  cheb_diff_v<2, C::NTHR, C::NMAX>(ta, -1, ta, -1, n, a.isx, red);
  for (int i = threadIdx.x >> 1; i < n; i += C::NTHR / 2) {
    cplx e = mk(0.0, 0.0);
    if (i < m) e = cscale(tb[cidx<2>(i, c)], __ldg(&a.sd[i]));
    if (i >= 2) e = sfma(__ldg(&a.sl[i - 2]), tb[cidx<2>(i - 2, c)], e);
    const cplx v = cadd(ta[cidx<2>(i, c)], e);
    ta[cidx<2>(i, c)] = v;
    st2(a.div, i, col, v);
  }
  __syncthreads();
  for (int i = threadIdx.x >> 1; i < m; i += C::NTHR / 2) {
    const cplx up = (i + 4 < n) ? cscale(ta[cidx<2>(i + 4, c)], __ldg(&a.b2.up[i])) : mk(0.0, 0.0);
    st2(a.r1, i, col, sfma(__ldg(&a.b2.lo[i]), ta[cidx<2>(i, c)], sfma(__ldg(&a.b2.di[i]), ta[cidx<2>(i + 2, c)], up)));
  }
}

template <int LOG2LB>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_project(XProjectArgs a) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  cplx* tb = ta + C::AROWS * 2;
  double* red = (double*)(tb + C::AROWS * 2);
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.nx, m = n - 2;
  xfill<C::NTHR>(tb, m, [&](int i) { return ld2(a.phi, i, col); });
  __syncthreads();
  xstencil_tile<C::NTHR>(ta, tb, n, a.nsd, a.nsl);
  __syncthreads();
  cheb_diff_v<2, C::NTHR, C::NMAX>(ta, -1, tb, -1, n, a.isx, red);
  from_ortho_v<2, C::NTHR, C::NMAX>(tb, -1, n, a.t, red);
  from_ortho_v<2, C::NTHR, C::NMAX>(ta, -1, n, a.t, red);
  for (int i = threadIdx.x >> 1; i < m; i += C::NTHR / 2) {
    st2(a.a1, i, col, tb[cidx<2>(i, c)]);
    st2(a.a2, i, col, ta[cidx<2>(i, c)]);
  }
}

// x half of a stand-alone HholtzAdi::solve (hholtz_adi.rs:108,128): B2_x matvec + Fdma_x along the strided lanes
template <int LOG2LB>
__global__ void __launch_bounds__(XCfg<LOG2LB>::NTHR, XCfg<LOG2LB>::MINB) xk_adi(XAdiArgs a) {
  typedef XCfg<LOG2LB> C;
  constexpr int LR = C::LR;
  RP_DYN_SMEM(double, ta_);
  cplx* ta = (cplx*)ta_;
  cplx* tb = ta + C::AROWS * 2;  // scratch of the second-order sweep (idle second tile)
  double* red = (double*)(tb + C::AROWS * 2);
  const int c0 = blockIdx.x * LR, col = c0 + 2 * (threadIdx.x & 1), c = threadIdx.x & 1;
  const int n = a.nx, m = n - 2;
  xfill<C::NTHR>(ta, n, [&](int i) { return ld2(a.in, i, col); });
  __syncthreads();
  b2_fdma_v<2, C::NTHR, C::NMAX, C::NSC2>(ta, -1, n, a.pt1, a.pt2, red, C::NSC2 == C::NSC ? (double*)tb : red);
  for (int i = threadIdx.x >> 1; i < m; i += C::NTHR / 2) st2(a.out, i, col, ta[cidx<2>(i, c)]);
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
// log2 Bluestein length (complex lanes per tile: always 2)
#define XK_SIZES(X) X(6, 2) X(7, 2) X(8, 2) X(9, 2) X(10, 2) X(11, 2) X(12, 2)

static int bluestein_log2(int n0) {  // tables.cu: Lb = next_pow2(2N - 1)
  const int need = 2 * (n0 - 1) - 1;
  int l = 0;
  while ((1 << l) < need) ++l;
  return l;
}

ScanShape x_scan_shape(int n0) {
  const int l = bluestein_log2(n0);
#define X(L, LCV) \
  if (l == L) return ScanShape{XCfg<L>::NSC / 4, chunk_len_v(XCfg<L>::NMAX, XCfg<L>::NSC, 2)};
  XK_SIZES(X)
#undef X
  return ScanShape{0, 0};
}
ScanShape x_scan2_shape(int n0) {
  const int l = bluestein_log2(n0);
#define X(L, LCV) \
  if (l == L) return ScanShape{XCfg<L>::NSC2 / 4, chunk_len_v(XCfg<L>::NMAX, XCfg<L>::NSC2, 2)};
  XK_SIZES(X)
#undef X
  return ScanShape{0, 0};
}

bool x_supported(int n0) {
  if (n0 < 8) return false;
  const int l = bluestein_log2(n0);  // (a power-of-two period N runs dct_pow2 inside the tile of Lb = 2 N: dct_x)
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
  if (!ok_ && l_ == L) {                                                         \
    typedef XCfg<L> C;                                                           \
    auto kp_ = kern<L>;                                                          \
    const int sm_ = (smem_expr);                                                 \
    const int nb_ = ((ncols_) + C::LR - 1) / C::LR;                              \
    static unsigned long long init_ = 0; /* one bit per device */                                                   \
    if (first_use_on_device(init_)) {                                                                \
      set_smem(kp_, sm_);                                                        \
    }                                                                            \
    RP_LAUNCH(kp_, dim3(nb_, nby_), dim3(C::NTHR), (size_t)sm_, s, a);           \
    ok_ = true;                                                                  \
  }
// kernels with a DCT: the Bluestein instantiation, or the power-of-two one when the tables hold no chirp (p2_)
#define XK_CASE_BODY_DCT(kern, L, LCV, smem_expr)                                \
  if (!ok_ && l_ == L) {                                                         \
    typedef XCfg<L> C;                                                           \
    auto kp_ = p2_ ? kern<L, true> : kern<L, false>;                             \
    const int sm_ = (smem_expr);                                                 \
    const int nb_ = ((ncols_) + C::LR - 1) / C::LR;                              \
    static unsigned long long init_[2] = {0, 0}; /* one bit per device */       \
    if (first_use_on_device(init_[p2_ ? 1 : 0])) {                               \
      set_smem(kp_, sm_);                                                        \
    }                                                                            \
    RP_LAUNCH(kp_, dim3(nb_, nby_), dim3(C::NTHR), (size_t)sm_, s, a);           \
    ok_ = true;                                                                  \
  }
#define XK_CASE_xk_backward(L, LCV) XK_CASE_BODY_DCT(xk_backward, L, LCV, C::SMEM_AW)
#define XK_CASE_xk_forward(L, LCV) XK_CASE_BODY_DCT(xk_forward, L, LCV, C::SMEM_AW)
#define XK_CASE_xk_div(L, LCV) XK_CASE_BODY(xk_div, L, LCV, C::SMEM_AA)
#define XK_CASE_xk_project(L, LCV) XK_CASE_BODY(xk_project, L, LCV, C::SMEM_AA)
#define XK_CASE_xk_adi(L, LCV) XK_CASE_BODY(xk_adi, L, LCV, C::SMEM_AA)
#define XK_CASE_xk_fdct(L, LCV) XK_CASE_BODY_DCT(xk_fdct, L, LCV, C::SMEM_AW)

#define XK_LAUNCH(kern, ncols, nx, nby)                                            \
  do {                                                                             \
    const int nby_ = (nby);                                                        \
    const int l_ = bluestein_log2(nx);                                             \
    const int ncols_ = (ncols);                                                    \
    bool ok_ = false;                                                              \
    XK_SIZES(XK_CASE_##kern)                                                       \
    if (!ok_) throw Error(RP_ERR_INTERNAL, #kern ": unsupported lane length");     \
  } while (0)

// RUSTPDE_B200_KFLAGS -> the development switches of this translation unit's kernels (fast.cuh fk_kflags_c)
void apply_kflags() {
#ifndef RP_EMU
  static unsigned long long done = 0;
  if (!first_use_on_device(done)) return;
  const char* e = getenv("RUSTPDE_B200_KFLAGS");
  const int v = e ? atoi(e) : 0;
  RP_CUDA_CHECK(cudaMemcpyToSymbol(fk_kflags_c, &v, sizeof(int)));
#endif
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
  a.next_wave = sm_count();  // the block that follows on the same SM is about one wave ahead
  const bool p2_ = a.a[0].t.chirp == nullptr;
  XK_LAUNCH(xk_backward, a.a[0].src.cols, a.a[0].t.n, nb);
}
void launch_x_forward(const XForwardArgs3& a_, int nb, cudaStream_t s) {
  XForwardArgs3 a = a_;
  a.next_wave = sm_count();
  const bool p2_ = a.a[0].t.chirp == nullptr;
  XK_LAUNCH(xk_forward, a.a[0].conv.cols, a.a[0].t.n, nb);
}
void launch_x_fdct(const XFdctArgs3& a_, int nb, cudaStream_t s) {
  XFdctArgs3 a = a_;
  a.next_wave = sm_count();
  const bool p2_ = a.a[0].t.chirp == nullptr;
  XK_LAUNCH(xk_fdct, a.a[0].conv.cols, a.a[0].t.n, nb);
}
void launch_x_div(const XDivArgs& a, cudaStream_t s) { XK_LAUNCH(xk_div, a.vx.cols, a.nx, 1); }
void launch_x_project(const XProjectArgs& a, cudaStream_t s) { XK_LAUNCH(xk_project, a.phi.cols, a.nx, 1); }
void launch_x_adi(const XAdiArgs& a, cudaStream_t s) { XK_LAUNCH(xk_adi, a.in.cols, a.nx, 1); }

}  // namespace fk
}  // namespace rp
