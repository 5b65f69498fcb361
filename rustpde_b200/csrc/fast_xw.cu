// fast_xw.cu -- x-direction banded sweeps of the confined Navier2D::update as WARP-SERIAL column sweeps.
//
// The recurrences of the reference along x (Fdma forward / backward sweep fdma.rs:101-118, the TDMA of from_ortho
// linalg.rs:14-57, the Chebyshev derivative ortho.rs:107-125) are sequential along a lane, and the lanes of an x pass
// are the columns -- contiguous in memory.  So a warp owns a strip of 16 adjacent columns (one 128-byte line per row)
// and each of its threads owns ONE parity chain of ONE column, which it walks from end to end in the reference's own
// order: no chunking, no carry exchange, no second walk, no block barrier.  What makes that fast is the memory side:
// the rows ahead of the chain are staged into a per-warp shared-memory ring by cp.async (LDGSTS, 16 bytes per lane,
// XW_K batches of XW_RB rows in flight per array), together with the per-row coefficients, so the chain itself only
// sees shared-memory loads, a handful of FP64 instructions and one coalesced store per element.  A sweep that needs
// the result of the previous one in the opposite direction (forward / backward substitution) hands it over through a
// global scratch array that the same warp reads back last-in-first-out (mostly L2 hits).
//
//   xw_adi     : x half of HholtzAdi (hholtz_adi.rs:108,128): B2_x matvec + Fdma_x forward + backward sweep
//   xw_div     : divergence (navier.rs:698-703) + B2_x of the Poisson rhs
//   xw_project : x part of u -= from_ortho(grad phi) (navier.rs:683-695)
#include "fast.cuh"

namespace rp {
namespace fk {

namespace {
constexpr int XW_RB = 16;  // rows per batch (8 steps of each parity chain)
constexpr int XW_WPB = 2;  // warps (strips) per block

FK_DEV void xw_syncwarp() { __syncwarp(); }

// Stage rows [r0, r0 + XW_RB) x columns [c0, c0 + 16) of f into dst[XW_RB][16]; rows outside [0, f.rows) and columns
// >= f.cols are zero-filled.  f.p 16-byte aligned, f.ld even, c0 even.
FK_DEV void xw_stage(double* dst, const Mat& f, int r0, int c0, int lane) {
  const int crow = lane >> 3, cc = c0 + 2 * (lane & 7);
  const int nb = max(0, min(16, (f.cols - cc) * 8));
#pragma unroll
  for (int k = 0; k < XW_RB / 4; ++k) {
    const int r = r0 + 4 * k + crow;
    const bool v = r >= 0 && r < f.rows && nb > 0;
    cp_async16(&dst[(4 * k + crow) * 16 + 2 * (lane & 7)], f.p + (v ? (size_t)r * f.ld + cc : 0), v ? nb : 0);
  }
}
// Stage XW_RB rows of a table of 4 doubles per row: dst[j][0..3] = tab[r0 + j][0..3], zero outside [0, nt)
FK_DEV void xw_stage_tab4(double* dst, const double* __restrict__ tab, int nt, int r0, int lane) {
  static_assert(XW_RB * 2 == 32, "one 16-byte chunk per lane");
  const int r = r0 + (lane >> 1);
  const bool v = r >= 0 && r < nt;
  cp_async16(&dst[lane * 2], tab + (v ? (size_t)r * 4 + 2 * (lane & 1) : 0), v ? 16 : 0);
}
// Stage XW_RB rows of a table of W doubles per row (W even): dst[j][0..W) = tab[r0 + j][0..W), zero outside [0, nt)
template <int W>
FK_DEV void xw_stage_tab(double* dst, const double* __restrict__ tab, int nt, int r0, int lane) {
  constexpr int H = W / 2, TOT = XW_RB * H;  // 16-byte chunks per row / per batch
#pragma unroll
  for (int id0 = 0; id0 < TOT; id0 += 32) {
    const int id = id0 + lane;
    if (TOT % 32 == 0 || id < TOT) {
      const int j = id / H, h = id % H, r = r0 + j;
      const bool v = r >= 0 && r < nt;
      cp_async16(&dst[j * W + 2 * h], tab + (v ? (size_t)r * W + 2 * h : 0), v ? 16 : 0);
    }
  }
}
struct Q4 {
  double x, y, z, w;
};
FK_DEV Q4 xw_ld4(const double* p) {
  const double2 a = *(const double2*)p, b = *(const double2*)(p + 2);
  return Q4{a.x, a.y, b.x, b.y};
}
}  // namespace

// ---------------------------------------------------------------------------------
// out = Fdma_x(B2_x in):  forward sweep  y_i = fp_i y_{i-2} + lo_i r_i + di_i r_{i+2} + up_i r_{i+4}   (i < m = n - 2)
//                         backward sweep z_i = bs_i y_i + bp1_i z_{i+2} + bp2_i z_{i+4}
// cf[i] = {lo, di, up, fp}, cb[i] = {bs, bp1, bp2, 0} (pack_rows); y goes through `tmp`.
template <int K>
FK_DEV void xw_adi_body(const XwAdiArgs& a, double* ring, int c0, int lane) {
  constexpr int D = K + 1, RB = XW_RB;
  constexpr int SLOT = RB * 16 + RB * 4;
  const int lc = lane & 15, p = lane >> 4, col = c0 + lc;
  const int n = a.nx, m = n - 2;
  const bool ok = col < a.out.cols;
  {
    // the chain runs 2 steps (4 rows) behind the staged rows: r_{i+4} is the newest element of the window
    const int nbat = (n + 4 + RB - 1) / RB;
    auto issue = [&](int b) {
      if (b < nbat) {
        double* slot = ring + (b % D) * SLOT;
        xw_stage(slot, a.in, b * RB, c0, lane);
        xw_stage_tab4(slot + RB * 16, a.cf, m, b * RB - 4, lane);
      }
      cp_async_commit();
    };
    for (int b = 0; b < K; ++b) issue(b);
    double r0 = 0.0, r1 = 0.0, r2 = 0.0, y = 0.0;
    double* tp = a.tmp.p + (ptrdiff_t)(p - 4) * a.tmp.ld + col;
    for (int b = 0; b < nbat; ++b) {
      issue(b + K);
      cp_async_wait<K>();
      xw_syncwarp();
      const double* slot = ring + (b % D) * SLOT;
      double rin[RB / 2];
      Q4 c[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        rin[u] = slot[(2 * u + p) * 16 + lc];
        c[u] = xw_ld4(&slot[RB * 16 + (2 * u + p) * 4]);  // coefficients of row i = (staged row) - 4, zero outside [0, m)
      }
      const int i0 = b * RB + p - 4;
      const bool inner = ok && i0 >= 0 && i0 + RB - 2 < m;
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        r0 = r1, r1 = r2, r2 = rin[u];
        y = fma(c[u].w, y, fma(c[u].x, r0, fma(c[u].y, r1, c[u].z * r2)));
        const int i = i0 + 2 * u;
        if (inner || (ok && i >= 0 && i < m)) tp[(ptrdiff_t)(2 * u) * a.tmp.ld] = y;
      }
      tp += (ptrdiff_t)RB * a.tmp.ld;
      xw_syncwarp();
    }
  }
#ifndef RP_EMU
  __threadfence_block();
#endif
  xw_syncwarp();
  {
    const int nbat = (m + RB - 1) / RB;
    auto issue = [&](int b) {
      if (b >= 0) {
        double* slot = ring + (b % D) * SLOT;
        xw_stage(slot, a.tmp, b * RB, c0, lane);
        xw_stage_tab4(slot + RB * 16, a.cb, m, b * RB, lane);
      }
      cp_async_commit();
    };
    for (int b = 0; b < K; ++b) issue(nbat - 1 - b);
    double z1 = 0.0, z2 = 0.0;
    double* op = a.out.p + (ptrdiff_t)((nbat - 1) * RB + p) * a.out.ld + col;
    for (int b = nbat - 1; b >= 0; --b) {
      issue(b - K);
      cp_async_wait<K>();
      xw_syncwarp();
      const double* slot = ring + (b % D) * SLOT;
      double rin[RB / 2];
      Q4 c[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        rin[u] = slot[(2 * u + p) * 16 + lc];
        c[u] = xw_ld4(&slot[RB * 16 + (2 * u + p) * 4]);
      }
      const int i0 = b * RB + p;
      const bool inner = ok && i0 + RB - 2 < m;
#pragma unroll
      for (int u = RB / 2 - 1; u >= 0; --u) {
        // rows >= m are zero-filled together with their coefficients: z stays 0 until the first real row
        const double z = fma(c[u].y, z1, fma(c[u].z, z2, c[u].x * rin[u]));
        z2 = z1, z1 = z;
        if (inner || (ok && i0 + 2 * u < m)) op[(ptrdiff_t)(2 * u) * a.out.ld] = z;
      }
      op -= (ptrdiff_t)RB * a.out.ld;
      xw_syncwarp();
    }
  }
}
constexpr int XW_ADI_K = 6;
constexpr int XW_ADI_SMEM = XW_WPB * (XW_ADI_K + 1) * (XW_RB * 20) * 8;
__global__ void __launch_bounds__(32 * XW_WPB) xw_adi(XwAdiArgs3 a3) {
  RP_DYN_SMEM(double, smem);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int c0 = (blockIdx.x * XW_WPB + wib) * 16;
  double* ring = smem + wib * ((XW_ADI_K + 1) * (XW_RB * 20));
  // blockIdx.y selects the field; the branch is block-uniform and keeps the arguments direct constant-bank operands
  if (blockIdx.y == 0) {
    if (c0 < a3.a[0].out.cols) xw_adi_body<XW_ADI_K>(a3.a[0], ring, c0, lane);
  } else if (blockIdx.y == 1) {
    if (c0 < a3.a[1].out.cols) xw_adi_body<XW_ADI_K>(a3.a[1], ring, c0, lane);
  } else {
    if (c0 < a3.a[2].out.cols) xw_adi_body<XW_ADI_K>(a3.a[2], ring, c0, lane);
  }
}


// ---------------------------------------------------------------------------------
// div = D_x S_x vx / sx + S_x ey,  r1 = B2_x div   (navier.rs:698-703, poisson.rs:131-137) in ONE sweep from the last row
// to the first: the Chebyshev derivative (ortho.rs:107-125) is a running sum over the rows above of the other parity,
//   div_j = (j == 0 ? 1/2 : 1) sum_{k = j+1, j+3, ..} (2 k / sx) V_k + E_j,   V_k = sd_k vx_k + sl_{k-2} vx_{k-2},
//   E_j = sd_j ey_j + sl_{j-2} ey_{j-2},   r1_i = lo_i div_i + di_i div_{i+2} + up_i div_{i+4}   (i < m = n - 2)
// tab[j] = {sd_{j+1}, sl_{j-1}, 2 (j+1) / sx, sd_j, sl_{j-2}, lo_j, di_j, up_j} (xw_div_table), zero outside the bands.
// The chain of parity p works on row j = (staged row of its own parity) + 2: the newest elements it needs are ey_{j-2}
// and vx_{j-1}; the other-parity value of the odd chain arrives one step early and waits in a register.
template <int K>
FK_DEV void xw_div_body(const XwDivArgs& a, double* ring, int c0, int lane) {
  constexpr int D = K + 1, RB = XW_RB;
  constexpr int SLOT = 2 * RB * 16 + RB * 8;
  const int lc = lane & 15, p = lane >> 4, col = c0 + lc;
  const int n = a.nx, m = n - 2;
  const bool ok = col < a.div.cols;
  const int nbat = (n + RB - 1) / RB;
  auto issue = [&](int b) {  // batches nbat - 1 .. -1 (rows below 0 are staged as zeros)
    if (b >= -1) {
      double* slot = ring + ((b + 1) % D) * SLOT;
      xw_stage(slot, a.vx, b * RB, c0, lane);
      xw_stage(slot + RB * 16, a.ey, b * RB, c0, lane);
      xw_stage_tab<8>(slot + 2 * RB * 16, a.tab, n, b * RB + 2, lane);
    }
    cp_async_commit();
  };
  for (int b = 0; b < K; ++b) issue(nbat - 1 - b);
  double vq0 = 0.0, vq1 = 0.0, vq2 = 0.0, eold = 0.0, acc = 0.0, dv2 = 0.0, dv4 = 0.0;
  for (int b = nbat - 1; b >= -1; --b) {
    issue(b - K);
    cp_async_wait<K>();
    xw_syncwarp();
    const double* slot = ring + ((b + 1) % D) * SLOT;
    double vn[RB / 2], en[RB / 2];
    Q4 ca[RB / 2], cb[RB / 2];
#pragma unroll
    for (int u = 0; u < RB / 2; ++u) {
      vn[u] = slot[(2 * u + 1 - p) * 16 + lc];
      en[u] = slot[RB * 16 + (2 * u + p) * 16 + lc];
      ca[u] = xw_ld4(&slot[2 * RB * 16 + (2 * u + p) * 8]);
      cb[u] = xw_ld4(&slot[2 * RB * 16 + (2 * u + p) * 8 + 4]);
    }
    const int j0 = b * RB + p + 2;
    const bool inner = ok && j0 >= 0 && j0 + RB - 2 < m;
    double* dp = a.div.p + (ptrdiff_t)j0 * a.div.ld + col;
    double* rp_ = a.r1.p + (ptrdiff_t)j0 * a.r1.ld + col;
#pragma unroll
    for (int u = RB / 2 - 1; u >= 0; --u) {
      vq2 = vq1, vq1 = vq0, vq0 = vn[u];
      const double va = p ? vq1 : vq0, vb = p ? vq2 : vq1;  // vx_{j-1}, vx_{j+1}
      const double V = fma(ca[u].y, va, ca[u].x * vb);
      acc = acc + __dmul_rn(V, ca[u].z);
      const int j = j0 + 2 * u;
      const double E = fma(cb[u].x, en[u], ca[u].w * eold);
      eold = en[u];
      const double dv = ((j == 0) ? 0.5 * acc : acc) + E;
      const double r = fma(cb[u].y, dv, fma(cb[u].z, dv2, cb[u].w * dv4));
      dv4 = dv2, dv2 = dv;
      if (inner || (ok && j >= 0 && j < n)) dp[(ptrdiff_t)(2 * u) * a.div.ld] = dv;
      if (inner || (ok && j >= 0 && j < m)) rp_[(ptrdiff_t)(2 * u) * a.r1.ld] = r;
    }
    xw_syncwarp();
  }
}
constexpr int XW_DIV_K = 4;
constexpr int XW_DIV_SMEM = (XW_DIV_K + 1) * (2 * XW_RB * 16 + XW_RB * 8) * 8;
__global__ void __launch_bounds__(32) xw_div(XwDivArgs a) {
  RP_DYN_SMEM(double, smem);
  const int c0 = blockIdx.x * 16;
  if (c0 < a.div.cols) xw_div_body<XW_DIV_K>(a, smem, c0, threadIdx.x & 31);
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// a1 = from_ortho_x(D_x S_x phi / sx),  a2 = from_ortho_x(S_x phi)   (navier.rs:683-695; composite_stencil.rs:250-276).
// The derivative needs a sweep from the last row to the first, so the (S^T S) solve is eliminated in the same direction
// (UL instead of the reference's LU order, linalg.rs:14-57: the same tridiagonal solve, pivots q_i = a_i - b_i^2 / q_{i+2}):
//   pass 1 (descending)  o_j = nsd_j phi_j + nsl_{j-2} phi_{j-2},  d_j = (j == 0 ? 1/2 : 1) sum_{k = j+1, j+3, ..} (2 k / sx) o_k,
//                        c_i = sd_i t_i + sl_i t_{i+2}  (t = d, o),   z_i = gs_i c_i + gp_i z_{i+2}        -> a1, a2
//   pass 2 (ascending)   x_i = z_i + hp_i x_{i-2}                                                         in place
// t1[j] = {nsd_j, nsl_{j-2}, w_j nsd_{j+1}, w_j nsl_{j-1}, sd_j, sl_j, gs_j, gp_j}, w_j = 2 (j+1) / sx;  t2[i] = {hp_i, 0}.
template <int K>
FK_DEV void xw_project_body(const XwProjectArgs& a, double* ring, int c0, int lane) {
  constexpr int D = K + 1, RB = XW_RB;
  const int lc = lane & 15, p = lane >> 4, col = c0 + lc;
  const int n = a.nx, m = n - 2;
  const bool ok = col < a.phi.cols;
  {
    constexpr int SLOT = RB * 16 + RB * 8;
    const int nbat = (n + RB - 1) / RB;
    auto issue = [&](int b) {
      if (b >= -1) {
        double* slot = ring + ((b + 1) % D) * SLOT;
        xw_stage(slot, a.phi, b * RB, c0, lane);
        xw_stage_tab<8>(slot + RB * 16, a.t1, n, b * RB + 2, lane);
      }
      cp_async_commit();
    };
    for (int b = 0; b < K; ++b) issue(nbat - 1 - b);
    double q0 = 0.0, q1 = 0.0, q2 = 0.0, fold = 0.0, acc = 0.0, oprev = 0.0, dprev = 0.0, z1 = 0.0, z2 = 0.0;
    for (int b = nbat - 1; b >= -1; --b) {
      issue(b - K);
      cp_async_wait<K>();
      xw_syncwarp();
      const double* slot = ring + ((b + 1) % D) * SLOT;
      double own[RB / 2], oth[RB / 2];
      Q4 ca[RB / 2], cb[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        own[u] = slot[(2 * u + p) * 16 + lc];
        oth[u] = slot[(2 * u + 1 - p) * 16 + lc];
        ca[u] = xw_ld4(&slot[RB * 16 + (2 * u + p) * 8]);
        cb[u] = xw_ld4(&slot[RB * 16 + (2 * u + p) * 8 + 4]);
      }
      const int j0 = b * RB + p + 2;
      const bool inner = ok && j0 >= 0 && j0 + RB - 2 < m;
      double* p1 = a.a1.p + (ptrdiff_t)j0 * a.a1.ld + col;
      double* p2 = a.a2.p + (ptrdiff_t)j0 * a.a2.ld + col;
#pragma unroll
      for (int u = RB / 2 - 1; u >= 0; --u) {
        q2 = q1, q1 = q0, q0 = oth[u];
        const double fa = p ? q1 : q0, fb = p ? q2 : q1;  // phi_{j-1}, phi_{j+1}
        const double o = fma(ca[u].y, own[u], ca[u].x * fold);  // o_j
        fold = own[u];
        acc = acc + fma(ca[u].w, fa, ca[u].z * fb);  // + w_j o_{j+1}
        const int j = j0 + 2 * u;
        const double d = (j == 0) ? 0.5 * acc : acc;
        const double c1 = fma(cb[u].y, dprev, cb[u].x * d), c2 = fma(cb[u].y, oprev, cb[u].x * o);
        oprev = o, dprev = d;
        z1 = fma(cb[u].w, z1, cb[u].z * c1);
        z2 = fma(cb[u].w, z2, cb[u].z * c2);
        if (inner || (ok && j >= 0 && j < m)) {
          p1[(ptrdiff_t)(2 * u) * a.a1.ld] = z1;
          p2[(ptrdiff_t)(2 * u) * a.a2.ld] = z2;
        }
      }
      xw_syncwarp();
    }
  }
  cp_async_wait<0>();
#ifndef RP_EMU
  __threadfence_block();
#endif
  xw_syncwarp();
  {
    constexpr int SLOT = 2 * RB * 16 + RB * 2;
    const int nbat = (m + RB - 1) / RB;
    auto issue = [&](int b) {
      if (b < nbat) {
        double* slot = ring + (b % D) * SLOT;
        xw_stage(slot, a.a1, b * RB, c0, lane);
        xw_stage(slot + RB * 16, a.a2, b * RB, c0, lane);
        xw_stage_tab<2>(slot + 2 * RB * 16, a.t2, m, b * RB, lane);
      }
      cp_async_commit();
    };
    for (int b = 0; b < K; ++b) issue(b);
    double x1 = 0.0, x2 = 0.0;
    for (int b = 0; b < nbat; ++b) {
      issue(b + K);
      cp_async_wait<K>();
      xw_syncwarp();
      const double* slot = ring + (b % D) * SLOT;
      double za[RB / 2], zb[RB / 2], hp[RB / 2];
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        za[u] = slot[(2 * u + p) * 16 + lc];
        zb[u] = slot[RB * 16 + (2 * u + p) * 16 + lc];
        hp[u] = slot[2 * RB * 16 + (2 * u + p) * 2];
      }
      const int i0 = b * RB + p;
      const bool inner = ok && i0 + RB - 2 < m;
      double* p1 = a.a1.p + (ptrdiff_t)i0 * a.a1.ld + col;
      double* p2 = a.a2.p + (ptrdiff_t)i0 * a.a2.ld + col;
#pragma unroll
      for (int u = 0; u < RB / 2; ++u) {
        x1 = fma(hp[u], x1, za[u]);
        x2 = fma(hp[u], x2, zb[u]);
        if (inner || (ok && i0 + 2 * u < m)) {
          p1[(ptrdiff_t)(2 * u) * a.a1.ld] = x1;
          p2[(ptrdiff_t)(2 * u) * a.a2.ld] = x2;
        }
      }
      xw_syncwarp();
    }
  }
}
constexpr int XW_PRJ_K = 4;
constexpr int XW_PRJ_SMEM = (XW_PRJ_K + 1) * (2 * XW_RB * 16 + XW_RB * 2) * 8;  // (pass 2 has the larger slot)
static_assert(2 * XW_RB * 16 + XW_RB * 2 >= XW_RB * 16 + XW_RB * 8, "slot");
__global__ void __launch_bounds__(32) xw_project(XwProjectArgs a) {
  RP_DYN_SMEM(double, smem);
  const int c0 = blockIdx.x * 16;
  if (c0 < a.phi.cols) xw_project_body<XW_PRJ_K>(a, smem, c0, threadIdx.x & 31);
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// launchers
// ---------------------------------------------------------------------------------
bool xw_supported(int n0) { return n0 >= 8; }

std::vector<double> pack_rows(int rows, int W, const std::vector<std::vector<double>>& src, const std::vector<int>& shift) {
  std::vector<double> out((size_t)rows * W, 0.0);
  for (int i = 0; i < rows; ++i)
    for (int k = 0; k < (int)src.size(); ++k) {
      const int j = i + shift[k];
      if (j >= 0 && j < (int)src[k].size()) out[(size_t)i * W + k] = src[k][j];
    }
  return out;
}


// Coefficient rows of xw_div (see the kernel): sd / sl = Dirichlet stencil along x (m entries), lo / di / up = B2 rows (m)
std::vector<double> xw_div_table(int n, double isx, const std::vector<double>& sd, const std::vector<double>& sl,
                                 const std::vector<double>& lo, const std::vector<double>& di, const std::vector<double>& up) {
  const int m = n - 2;
  auto at = [](const std::vector<double>& v, int i) { return (i >= 0 && i < (int)v.size()) ? v[i] : 0.0; };
  std::vector<double> t((size_t)n * 8, 0.0);
  for (int j = 0; j < n; ++j) {
    double* r = &t[(size_t)j * 8];
    if (j + 1 < m) r[0] = at(sd, j + 1);
    if (j + 1 <= n - 1) r[1] = at(sl, j - 1);
    if (j + 1 <= n - 1) r[2] = 2.0 * (double)(j + 1) * isx;
    if (j < m) r[3] = at(sd, j);
    r[4] = at(sl, j - 2);
    if (j < m) r[5] = at(lo, j), r[6] = at(di, j), r[7] = at(up, j);
  }
  return t;
}
// Coefficient rows of xw_project: nsd / nsl = Neumann stencil of phi along x, sd / sl = Dirichlet stencil of the velocity
void xw_project_tables(int n, double isx, const std::vector<double>& nsd, const std::vector<double>& nsl, const std::vector<double>& sd,
                       const std::vector<double>& sl, std::vector<double>& t1, std::vector<double>& t2) {
  const int m = n - 2;
  auto at = [](const std::vector<double>& v, int i) { return (i >= 0 && i < (int)v.size()) ? v[i] : 0.0; };
  // (S^T S) of the Dirichlet stencil (composite_stencil.rs:160-171), eliminated from the last row upwards
  std::vector<double> av(m), bv(m, 0.0), q(m);
  for (int i = 0; i < m; ++i) av[i] = sd[i] * sd[i] + sl[i] * sl[i];
  for (int i = 0; i + 2 < m; ++i) bv[i] = sd[i + 2] * sl[i];
  for (int i = m - 1; i >= 0; --i) q[i] = (i + 2 < m) ? av[i] - bv[i] * bv[i] / q[i + 2] : av[i];
  t1.assign((size_t)n * 8, 0.0);
  t2.assign((size_t)m * 2, 0.0);
  for (int j = 0; j < n; ++j) {
    double* r = &t1[(size_t)j * 8];
    const double w = (j + 1 <= n - 1) ? 2.0 * (double)(j + 1) * isx : 0.0;
    if (j < m) r[0] = at(nsd, j);
    r[1] = at(nsl, j - 2);
    if (j + 1 < m) r[2] = w * at(nsd, j + 1);
    r[3] = w * at(nsl, j - 1);
    if (j < m) {
      r[4] = sd[j], r[5] = sl[j];
      r[6] = 1.0 / q[j];
      if (j + 2 < m) r[7] = -bv[j] / q[j];
    }
  }
  for (int i = 2; i < m; ++i) t2[(size_t)i * 2] = -bv[i - 2] / q[i];
}

template <class K>
static void xw_prepare(K kern, int bytes) {
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
#else
  (void)kern;
  (void)bytes;
#endif
}
#define XW_LAUNCH(kern, smem, ncols, nby, args)                                                                        \
  do {                                                                                                                  \
    static unsigned long long init_ = 0; /* one bit per device */                                                       \
    if (first_use_on_device(init_)) xw_prepare(kern, (smem));                                                           \
    const int nstrips_ = ((ncols) + 15) / 16;                                                                           \
    RP_LAUNCH(kern, dim3((nstrips_ + XW_WPB - 1) / XW_WPB, (nby)), dim3(32 * XW_WPB), (size_t)(smem), s, args);         \
  } while (0)

void launch_xw_adi(const XwAdiArgs3& a, int nb, cudaStream_t s) { XW_LAUNCH(xw_adi, XW_ADI_SMEM, a.a[0].out.cols, nb, a); }

// single-field sweeps: one warp (strip) per block, so that every strip gets an SM of its own
#define XW_LAUNCH1(kern, smem, ncols, args)                                                          \
  do {                                                                                                \
    static unsigned long long init_ = 0; /* one bit per device */                                     \
    if (first_use_on_device(init_)) xw_prepare(kern, (smem));                                         \
    RP_LAUNCH(kern, dim3(((ncols) + 15) / 16, 1), dim3(32), (size_t)(smem), s, args);                 \
  } while (0)
void launch_xw_div(const XwDivArgs& a, cudaStream_t s) { XW_LAUNCH1(xw_div, XW_DIV_SMEM, a.div.cols, a); }
void launch_xw_project(const XwProjectArgs& a, cudaStream_t s) { XW_LAUNCH1(xw_project, XW_PRJ_SMEM, a.phi.cols, a); }

}  // namespace fk
}  // namespace rp
