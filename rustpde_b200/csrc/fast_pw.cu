// fast_pw.cu -- per-mode passes of the periodic Navier2D::update on complex rows as WARP-SERIAL row sweeps.
//
//   pw_hholtz  : rhs assembly (navier.rs:622-674) + per-mode Helmholtz solve (hholtz.rs:156-197, fdma.rs:101-118)
//   pw_divpois : divergence (navier.rs:698-703) + per-mode Poisson solve (poisson.rs:131-149)
//   pw_project : projection + pressure update (navier.rs:683-721)
//
// The tile kernels of fast_p.cu (pk_hholtz, pk_divpois) hold a whole lane in shared memory and cut every recurrence
// into 64 .. 128 chunks (two walks and a carry exchange per scan).  For long lanes that leaves one small block per SM
// (an 8193-point lane pair is 128 KB), and the kernels are bound by latency at 11 % of the HBM roofline.  Here a
// thread owns ONE parity chain of ONE real lane (re or im of one Fourier mode kx) and walks it from end to end in the
// reference's own order; a warp owns 8 complex rows = 32 chains whose elements of one step are one 32-byte sector per
// row.  The columns ahead of the chain are staged into a per-warp shared-memory ring by cp.async (16 columns per batch,
// PW_K batches in flight per array) together with the per-column band tables, so nothing of a lane has to fit on chip:
// the sweeps stream at any lane length, with thousands of chains in flight at the large grids (8192 x 8193: 49 164).
// Forward and backward substitution hand their intermediate over through the output array itself (the same thread
// re-reads what it wrote, last in first out); d/dy pres of the uy rhs is a sweep of its own into a scratch array.
#include "fast.cuh"

namespace rp {
namespace fk {

namespace {
constexpr int PW_CB = 16;  // staged columns per batch (8 steps of each parity chain)
constexpr int PW_WPB = 2;  // warps per block
constexpr int PW_K = 2;    // batches in flight
constexpr int PW_D = PW_K + 1;
constexpr int PW_PC = 18;  // pitch of a staged complex row in complex elements (288 B: the 4 rows of a half-warp hit distinct banks)
constexpr int PW_PR = 22;  // pitch of a staged real row (pivot reciprocals) in doubles (176 B)
constexpr int PW_ARR = 8 * PW_PC * 2;  // doubles of one staged complex array

// Stage columns [j0, j0 + NC) of the 8 complex rows from r0 of f into dst[(row * PW_PC + c) * 2 + part]; zero outside f.
// Lanes 4 r .. 4 r + 3 copy row r, 16 bytes (one complex element) each, chunks (lane & 3) + 4 k: per array and batch that
// is one address and NC / 4 LDGSTS with immediate offsets; whole batches inside the matrix need no column predicate.
template <int NC>
FK_DEV void pw_stage_c(double* dst, const Mat& f, int r0, int j0, int lane) {
  static_assert(NC <= PW_PC, "pitch");
  const int row = lane >> 2, q = lane & 3, r = r0 + row;
  const bool rv = r < f.rows;
  const double* g = f.p + ((ptrdiff_t)(rv ? r : 0) * f.ld + j0 + q) * 2;
  double* d = dst + (row * PW_PC + q) * 2;
  if (j0 >= 0 && j0 + NC <= f.cols) {
#pragma unroll
    for (int k = 0; k < (NC + 3) / 4; ++k)
      if (4 * k + 3 < NC || q + 4 * k < NC) cp_async16(d + 8 * k, g + 8 * k, rv ? 16 : 0);
  } else {
#pragma unroll
    for (int k = 0; k < (NC + 3) / 4; ++k)
      if (4 * k + 3 < NC || q + 4 * k < NC) {
        const int j = j0 + q + 4 * k;
        const bool v = rv && j >= 0 && j < f.cols;
        cp_async16(d + 8 * k, v ? g + 8 * k : f.p, v ? 16 : 0);
      }
  }
}
// Stage columns [j0, j0 + NC) (j0, NC even) of the 8 rows from r0 of a real row-major table (ld even) into dst[row * PW_PR + c]
template <int NC>
FK_DEV void pw_stage_r(double* dst, const double* __restrict__ p, long long ld, int nrows, int ncols, int r0, int j0, int lane) {
  static_assert(NC % 2 == 0 && NC <= PW_PR, "pitch");
  constexpr int H = NC / 2;  // 16-byte chunks per row
  const int row = lane >> 2, q = lane & 3, r = r0 + row;
  const bool rv = r < nrows;
  const double* g = p + (ptrdiff_t)(rv ? r : 0) * ld + j0 + 2 * q;
  double* d = dst + row * PW_PR + 2 * q;
#pragma unroll
  for (int k = 0; k < (H + 3) / 4; ++k)
    if (4 * k + 3 < H || q + 4 * k < H) {
      const int j = j0 + 2 * (q + 4 * k);
      const bool v = rv && j >= 0 && j < ncols;
      cp_async16(d + 8 * k, v ? g + 8 * k : p, v ? min(16, (ncols - j) * 8) : 0);
    }
}
// Stage NCOL rows of a table of W doubles per column index: dst[c * W + k] = tab[(j0 + c) * W + k], zero outside [0, nt)
template <int W, int NCOL>
FK_DEV void pw_stage_tab(double* dst, const double* __restrict__ tab, int nt, int j0, int lane) {
  constexpr int H = W / 2, TOT = NCOL * H;
#pragma unroll
  for (int id0 = 0; id0 < TOT; id0 += 32) {
    const int id = id0 + lane;
    if (TOT % 32 == 0 || id < TOT) {
      const int c = id / H, h = id % H, j = j0 + c;
      const bool v = j >= 0 && j < nt;
      cp_async16(&dst[c * W + 2 * h], tab + (v ? (size_t)j * W + 2 * h : 0), v ? 16 : 0);
    }
  }
}
FK_DEV void pw_pass_fence() {  // what a pass stored is read back by other lanes of the warp (cp.async) in the next one
  cp_async_wait<0>();
#ifndef RP_EMU
  __threadfence_block();
#endif
  __syncwarp();
}

struct PwLane {
  int rl, p, part, r;  // row inside the warp, parity of the chain, re / im, matrix row
  bool rok;
};
FK_DEV PwLane pw_lane(int r0, int lane, int nrows) {
  PwLane L;
  L.rl = lane >> 2, L.p = (lane >> 1) & 1, L.part = lane & 1;
  L.r = r0 + L.rl;
  L.rok = L.r < nrows;
  return L;
}

// dst_j = (j == 0 ? 1/2 : 1) * sum_{k > j, k - j odd} (2 k sc) src_k,  j < n   (Chebyshev derivative ortho.rs:107-125
// along y, times sc).  A sweep from the last column to the first; the thread of output parity q reads the other parity.
FK_DEV void pw_diff_pass(const Mat& src, const Mat& dst, double sc, double* ring, int r0, int lane) {
  constexpr int NC = PW_CB + 2, SLOT = PW_ARR;
  const int n = src.cols;
  const PwLane L = pw_lane(r0, lane, src.rows);
  const int nbat = (n + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b >= 0) pw_stage_c<NC>(ring + (b % PW_D) * SLOT, src, r0, b * PW_CB, lane);  // columns [J0, J0 + 18)
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(nbat - 1 - b);
  double acc = 0.0;
  for (int b = nbat - 1; b >= 0; --b) {
    issue(b - PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT + (L.rl * PW_PC) * 2 + L.part;
    double v[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) v[u] = s[(2 * u + L.p + 1) * 2];  // column j + 1
#pragma unroll
    for (int u = PW_CB / 2 - 1; u >= 0; --u) {
      const int j = b * PW_CB + 2 * u + L.p;
      acc = acc + __dmul_rn(2.0 * (double)(j + 1) * sc, v[u]);  // (zero beyond the last column: staged zeros)
      if (L.rok && j < n) dst.p[((size_t)L.r * dst.ld + j) * 2 + L.part] = (j == 0) ? 0.5 * acc : acc;
    }
    __syncwarp();
  }
}

// Forward half of the per-mode banded solve (A + mu C) x = B2 rhs along y (fdma_tensor.rs:219-227 on the swept system):
//   x_i = g_i - l_{i-2} x_{i-2},  g_i = lo_i rhs_i + di_i rhs_{i+2} + up_i rhs_{i+4},  l_j = (a_low_j + mu c_low_j) / dia'_j
// with the rhs assembled on the fly (MODE 0 ux, 1 uy, 2 temperature: navier.rs:622-674; MODE 3: rhs = a.chat as it is).
// rf[i] = {lo, di, up, a_low[i-2], c_low[i-2], 0} (pack_rows); the chain runs 2 steps behind the staged columns.
template <int MODE>
FK_DEV void pw_forward_pass(const PHholtzArgs& a, double* ring, int r0, int lane, double mu) {
  constexpr int NA = MODE == 1 ? 5 : (MODE == 3 ? 1 : 3);
  constexpr int O_INV = NA * PW_ARR, O_RF = O_INV + 8 * PW_PR, O_RS = O_RF + PW_CB * 6, SLOT = O_RS + PW_CB * 4;
  const int n = a.ny, m = n - 2;
  const PwLane L = pw_lane(r0, lane, a.chat.rows);
  const int nbat = (n + 4 + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b < nbat) {
      double* s = ring + (b % PW_D) * SLOT;
      const int j0 = b * PW_CB;
      pw_stage_c<PW_CB>(s, a.chat, r0, j0, lane);
      if (MODE != 3) pw_stage_c<PW_CB>(s + PW_ARR, a.fld, r0, j0, lane);
      if (MODE == 0) pw_stage_c<PW_CB>(s + 2 * PW_ARR, a.pres, r0, j0, lane);
      if (MODE == 1) {
        pw_stage_c<PW_CB>(s + 2 * PW_ARR, a.tmp, r0, j0, lane);
        pw_stage_c<PW_CB>(s + 3 * PW_ARR, a.tbc, r0, j0, lane);
        pw_stage_c<PW_CB>(s + 4 * PW_ARR, a.dyp, r0, j0, lane);
      }
      if (MODE == 2) pw_stage_c<PW_CB>(s + 2 * PW_ARR, a.bcdiff, r0, j0, lane);
      pw_stage_r<PW_CB>(s + O_INV, a.m.inv, a.m.inv_ld, a.chat.rows, m, r0, j0 - 6, lane);  // 1 / dia'_{i-2}, i = j - 4
      pw_stage_tab<6, PW_CB>(s + O_RF, a.m.rf, m, j0 - 4, lane);
      if (MODE != 3) pw_stage_tab<4, PW_CB>(s + O_RS, a.rs, n, j0, lane);  // {sd_j, sl_{j-2}, tsd_j, tsl_{j-2}}, zero outside the bands
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(b);
  const double ks = -a.dt * a.isx * (double)(a.k0 + min(L.r, a.chat.rows - 1));
  double fprev = 0.0, tprev = 0.0;  // field / temperature coefficient of column j - 2
  double w1 = 0.0, w2 = 0.0, x = 0.0;
  double* op = a.out.p + ((size_t)L.r * a.out.ld) * 2 + L.part;
  for (int b = 0; b < nbat; ++b) {
    issue(b + PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    const int so = (L.rl * PW_PC) * 2 + L.part;
    double rhs[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      double v = s[so + c * 2];
      if (MODE != 3) {
        v = -a.dt * v;  // - dt * conv   (630, 651, 671)
        const double fv = s[PW_ARR + so + c * 2];
        const double2 st = *(const double2*)&s[O_RS + c * 4];
        v += fma(st.y, fprev, st.x * fv);  // + to_ortho(field)   (625, 644, 663)
        fprev = fv;
      }
      if (MODE == 0) {  // - dt/sx d/dx pres   (627)
        const double re = s[2 * PW_ARR + (L.rl * PW_PC + c) * 2], im = s[2 * PW_ARR + (L.rl * PW_PC + c) * 2 + 1];
        v += L.part ? ks * re : -ks * im;
      } else if (MODE == 1) {  // + dt * (that + tbc) - dt/sy d/dy pres   (646-648)
        const double tv = s[2 * PW_ARR + so + c * 2];
        const double2 tt = *(const double2*)&s[O_RS + c * 4 + 2];
        const double that = fma(tt.y, tprev, tt.x * tv) + s[3 * PW_ARR + so + c * 2];
        tprev = tv;
        v = s[4 * PW_ARR + so + c * 2] + fma(a.dt, that, v);
      } else if (MODE == 2) {  // + dt ka lap(fieldbc)   (665-668)
        v += s[2 * PW_ARR + so + c * 2];
      }
      rhs[u] = v;
    }
    double g[PW_CB / 2], c1[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      const double2 t0 = *(const double2*)&s[O_RF + c * 6], t1 = *(const double2*)&s[O_RF + c * 6 + 2],
                    t2 = *(const double2*)&s[O_RF + c * 6 + 4];  // lo, di | up, a_low | c_low, -
      const double ra = u == 0 ? w1 : (u == 1 ? w2 : rhs[u - 2]), rb = u == 0 ? w2 : rhs[u - 1];
      g[u] = fma(t0.x, ra, fma(t0.y, rb, t1.x * rhs[u]));
      c1[u] = -fma(mu, t2.x, t1.y) * s[O_INV + L.rl * PW_PR + c];
    }
    w1 = rhs[PW_CB / 2 - 2], w2 = rhs[PW_CB / 2 - 1];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      x = fma(c1[u], x, g[u]);
      const int i = b * PW_CB + 2 * u + L.p - 4;
      if (L.rok && i >= 0 && i < m) op[(size_t)i * 2] = x;
    }
    __syncwarp();
  }
}

// Backward half: x_i = (x_i - up1'_i x_{i+2} - up2_i x_{i+4}) / dia'_i   (fdma.rs:108-117), in place on xo ([rows, m] complex).
// rb[i] = {a_up1, c_up1, a_up2, c_up2, a_up2[i-2], c_up2[i-2], a_low[i-2], c_low[i-2]} (pack_rows, zero outside the bands)
FK_DEV void pw_backward_pass(const Mat& xo, const ModeTabs& M, int k0, bool zero00, double* ring, int r0, int lane, double mu) {
  constexpr int O_INV = PW_ARR, O_RB = O_INV + 8 * PW_PR, SLOT = O_RB + PW_CB * 8;
  const int m = xo.cols;
  const PwLane L = pw_lane(r0, lane, xo.rows);
  const int nbat = (m + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b >= 0) {
      double* s = ring + (b % PW_D) * SLOT;
      const int j0 = b * PW_CB;
      pw_stage_c<PW_CB>(s, xo, r0, j0, lane);
      pw_stage_r<PW_CB + 2>(s + O_INV, M.inv, M.inv_ld, xo.rows, m, r0, j0 - 2, lane);  // columns i - 2 .. i
      pw_stage_tab<8, PW_CB>(s + O_RB, M.rb, m, j0, lane);
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(nbat - 1 - b);
  double z1 = 0.0, z2 = 0.0;
  double* op = xo.p + ((size_t)L.r * xo.ld) * 2 + L.part;
  for (int b = nbat - 1; b >= 0; --b) {
    issue(b - PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    double q[PW_CB / 2], k1[PW_CB / 2], k2[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      const double xv = s[(L.rl * PW_PC + c) * 2 + L.part];
      const double ivm = s[O_INV + L.rl * PW_PR + c], iv = s[O_INV + L.rl * PW_PR + c + 2];
      const double2 t0 = *(const double2*)&s[O_RB + c * 8], t1 = *(const double2*)&s[O_RB + c * 8 + 2],
                    t2 = *(const double2*)&s[O_RB + c * 8 + 4], t3 = *(const double2*)&s[O_RB + c * 8 + 6];
      q[u] = iv * xv;
      double u1 = fma(mu, t0.y, t0.x);
      const double lw = fma(mu, t3.y, t3.x) * ivm;  // zero for i < 2 (empty band rows)
      u1 = fma(-lw, fma(mu, t2.y, t2.x), u1);
      k1[u] = -u1 * iv;
      k2[u] = -fma(mu, t1.y, t1.x) * iv;
    }
#pragma unroll
    for (int u = PW_CB / 2 - 1; u >= 0; --u) {
      // columns >= m are staged as zeros together with their band rows: z stays 0 until the first real column
      const double z = fma(k1[u], z1, fma(k2[u], z2, q[u]));
      z2 = z1, z1 = z;
      const int i = b * PW_CB + 2 * u + L.p;
      if (L.rok && i < m) op[(size_t)i * 2] = (zero00 && i == 0 && k0 + L.r == 0) ? 0.0 : z;
    }
    __syncwarp();
  }
}

// div = i k / sx S_y ux + D_y S_y uy / sy   (navier.rs:698-703): one sweep from the last column to the first
FK_DEV void pw_div_pass(const PDivPoisArgs& a, double* ring, int r0, int lane) {
  constexpr int NC = PW_CB + 2, O_RS = 2 * PW_ARR, SLOT = O_RS + NC * 4;
  const int n = a.ny;
  const PwLane L = pw_lane(r0, lane, a.ux.rows);
  const int nbat = (n + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b >= 0) {
      double* s = ring + (b % PW_D) * SLOT;
      pw_stage_c<NC>(s, a.uy, r0, b * PW_CB - 1, lane);           // columns j - 1 .. j + 1
      pw_stage_c<NC>(s + PW_ARR, a.ux, r0, b * PW_CB - 2, lane);  // columns j - 2 .. j
      pw_stage_tab<4, NC>(s + O_RS, a.rs, n, b * PW_CB, lane);    // {sd_j, sl_{j-2}, 2 j / sy, 0} of columns j, j + 1
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(nbat - 1 - b);
  const double ks = a.isx * (double)(a.k0 + min(L.r, a.ux.rows - 1));
  double acc = 0.0;
  for (int b = nbat - 1; b >= 0; --b) {
    issue(b - PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    const double* sy = s + (L.rl * PW_PC) * 2 + L.part;
    const double* sx = s + PW_ARR + (L.rl * PW_PC) * 2;
    double ta[PW_CB / 2], ik[PW_CB / 2], wk[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      const double2 cj = *(const double2*)&s[O_RS + c * 4], ck = *(const double2*)&s[O_RS + (c + 1) * 4];
      // S_y uy at column k = j + 1 (staged index c + 2), from columns k and k - 2, times 2 k / sy
      ta[u] = fma(ck.y, sy[c * 2], ck.x * sy[(c + 2) * 2]);
      wk[u] = s[O_RS + (c + 1) * 4 + 2];
      // S_y ux at column j (staged index c + 2), both parts
      const double re = fma(cj.y, sx[c * 2], cj.x * sx[(c + 2) * 2]), im = fma(cj.y, sx[c * 2 + 1], cj.x * sx[(c + 2) * 2 + 1]);
      ik[u] = L.part ? ks * re : -ks * im;
    }
#pragma unroll
    for (int u = PW_CB / 2 - 1; u >= 0; --u) {
      const int j = b * PW_CB + 2 * u + L.p;
      acc = acc + __dmul_rn(wk[u], ta[u]);
      const double d = (j == 0) ? 0.5 * acc : acc;
      if (L.rok && j < n) a.div.p[((size_t)L.r * a.div.ld + j) * 2 + L.part] = d + ik[u];
    }
    __syncwarp();
  }
}

// Projection + pressure update (navier.rs:683-721) as three row sweeps, the from_ortho solves in the reference's own
// order (S^T, then the pre-factored (S^T S) forward and backward substitution of linalg.rs:14-57):
//   pass 1 (last column to first)  o_j = nsd_j phi_j + nsl_{j-2} phi_{j-2}   (to_ortho of the pseudo-pressure, Neumann stencil)
//                                  p_j += -nu div_j + o_j / dt                                              (717-721)
//                                  d_j = (j == 0 ? 1/2 : 1) sum_{k = j+1, j+3, ..} (2 k / sy) o_k           (ortho.rs:107-125)
//                                  c_i = sd_i t_i + sl_i t_{i+2},  t = i k / sx o  |  d      -> scratch z1 (ux), z2 (uy)
//   pass 2 (first column to last)  y_i = fs_i c_i + fp_i y_{i-2}                             in place on the scratch
//   pass 3 (last column to first)  x_i = y_i + bp_i x_{i+2};   ux_i -= x1_i,  uy_i -= x2_i                  (683-695)
// w1[j] = {nsd_j, nsl_{j-2}, w_j nsd_{j+1}, w_j nsl_{j-1}, sd_j, sl_j, 0, 0}, w_j = 2 (j+1) / sy;  w2[i] = {fs_i, fp_i, bp_i, 0}
// (pw_project_tables).
FK_DEV void pw_project_pass1(const PProjectArgs& a, double* ring, int r0, int lane) {
  constexpr int NC = PW_CB + 2, O_T = 3 * PW_ARR, SLOT = O_T + PW_CB * 8;
  const int n = a.ny, m = n - 2;
  const PwLane L = pw_lane(r0, lane, a.phi.rows);
  const int nbat = (n + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b >= 0) {
      double* s = ring + (b % PW_D) * SLOT;
      pw_stage_c<NC>(s, a.phi, r0, b * PW_CB - 2, lane);  // columns j - 2 .. j - 1 of every j of the batch
      pw_stage_c<PW_CB>(s + PW_ARR, a.div, r0, b * PW_CB, lane);
      pw_stage_c<PW_CB>(s + 2 * PW_ARR, a.pres, r0, b * PW_CB, lane);
      pw_stage_tab<8, PW_CB>(s + O_T, a.w1, n, b * PW_CB, lane);
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(nbat - 1 - b);
  const double ks = a.isx * (double)(a.k0 + min(L.r, a.phi.rows - 1));
  double fown = 0.0, foth = 0.0;  // phi_j, phi_{j+1}: what the previous step read as phi_{j-2}, phi_{j-1}
  double acc = 0.0, tprev = 0.0, dprev = 0.0;
  for (int b = nbat - 1; b >= 0; --b) {
    issue(b - PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    const int so = (L.rl * PW_PC) * 2 + L.part;
    double f2[PW_CB / 2], f1[PW_CB / 2], dv[PW_CB / 2], pv[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      f2[u] = s[so + c * 2];        // phi_{j-2}
      f1[u] = s[so + (c + 1) * 2];  // phi_{j-1}
      dv[u] = s[PW_ARR + so + c * 2];
      pv[u] = s[2 * PW_ARR + so + c * 2];
    }
#pragma unroll
    for (int u = PW_CB / 2 - 1; u >= 0; --u) {
      const int c = 2 * u + L.p, j = b * PW_CB + c;
      const double2 t0 = *(const double2*)&s[O_T + c * 8], t1 = *(const double2*)&s[O_T + c * 8 + 2],
                    t2 = *(const double2*)&s[O_T + c * 8 + 4];
      const double o = fma(t0.y, f2[u], t0.x * fown);  // o_j
      acc = acc + fma(t1.y, f1[u], t1.x * foth);       // + (2 (j+1) / sy) o_{j+1}
      fown = f2[u], foth = f1[u];
      if (L.rok && j < n) a.pres.p[((size_t)L.r * a.pres.ld + j) * 2 + L.part] = fma(-a.nu, dv[u], pv[u]) + o * a.inv_dt;
      const double oo = __shfl_xor_sync(0xffffffffu, o, 1);  // the other part of o_j
      const double t = L.part ? ks * oo : -ks * oo;           // (i k / sx o).part: re' = -ks im, im' = ks re
      const double d = (j == 0) ? 0.5 * acc : acc;
      const double c1 = fma(t2.y, tprev, t2.x * t), c2 = fma(t2.y, dprev, t2.x * d);
      tprev = t, dprev = d;
      if (L.rok && j < m) {
        a.z1.p[((size_t)L.r * a.z1.ld + j) * 2 + L.part] = c1;
        a.z2.p[((size_t)L.r * a.z2.ld + j) * 2 + L.part] = c2;
      }
    }
    __syncwarp();
  }
}
FK_DEV void pw_project_pass2(const PProjectArgs& a, double* ring, int r0, int lane) {
  constexpr int O_T = 2 * PW_ARR, SLOT = O_T + PW_CB * 4;
  const int m = a.ny - 2;
  const PwLane L = pw_lane(r0, lane, a.phi.rows);
  const int nbat = (m + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b < nbat) {
      double* s = ring + (b % PW_D) * SLOT;
      pw_stage_c<PW_CB>(s, a.z1, r0, b * PW_CB, lane);
      pw_stage_c<PW_CB>(s + PW_ARR, a.z2, r0, b * PW_CB, lane);
      pw_stage_tab<4, PW_CB>(s + O_T, a.w2, m, b * PW_CB, lane);
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(b);
  double y1 = 0.0, y2 = 0.0;
  for (int b = 0; b < nbat; ++b) {
    issue(b + PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    const int so = (L.rl * PW_PC) * 2 + L.part;
    double ca[PW_CB / 2], cb[PW_CB / 2];
    double2 ff[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      ca[u] = s[so + c * 2], cb[u] = s[PW_ARR + so + c * 2];
      ff[u] = *(const double2*)&s[O_T + c * 4];  // fs, fp
    }
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int i = b * PW_CB + 2 * u + L.p;
      y1 = fma(ff[u].y, y1, ff[u].x * ca[u]);
      y2 = fma(ff[u].y, y2, ff[u].x * cb[u]);
      if (L.rok && i < m) {
        a.z1.p[((size_t)L.r * a.z1.ld + i) * 2 + L.part] = y1;
        a.z2.p[((size_t)L.r * a.z2.ld + i) * 2 + L.part] = y2;
      }
    }
    __syncwarp();
  }
}
FK_DEV void pw_project_pass3(const PProjectArgs& a, double* ring, int r0, int lane) {
  constexpr int O_T = 4 * PW_ARR, SLOT = O_T + PW_CB * 4;
  const int m = a.ny - 2;
  const PwLane L = pw_lane(r0, lane, a.phi.rows);
  const int nbat = (m + PW_CB - 1) / PW_CB;
  auto issue = [&](int b) {
    if (b >= 0) {
      double* s = ring + (b % PW_D) * SLOT;
      pw_stage_c<PW_CB>(s, a.z1, r0, b * PW_CB, lane);
      pw_stage_c<PW_CB>(s + PW_ARR, a.z2, r0, b * PW_CB, lane);
      pw_stage_c<PW_CB>(s + 2 * PW_ARR, a.ux, r0, b * PW_CB, lane);
      pw_stage_c<PW_CB>(s + 3 * PW_ARR, a.uy, r0, b * PW_CB, lane);
      pw_stage_tab<4, PW_CB>(s + O_T, a.w2, m, b * PW_CB, lane);
    }
    cp_async_commit();
  };
  for (int b = 0; b < PW_K; ++b) issue(nbat - 1 - b);
  double x1 = 0.0, x2 = 0.0;
  for (int b = nbat - 1; b >= 0; --b) {
    issue(b - PW_K);
    cp_async_wait<PW_K>();
    __syncwarp();
    const double* s = ring + (b % PW_D) * SLOT;
    const int so = (L.rl * PW_PC) * 2 + L.part;
    double ya[PW_CB / 2], yb[PW_CB / 2], ua[PW_CB / 2], ub[PW_CB / 2], bp[PW_CB / 2];
#pragma unroll
    for (int u = 0; u < PW_CB / 2; ++u) {
      const int c = 2 * u + L.p;
      ya[u] = s[so + c * 2], yb[u] = s[PW_ARR + so + c * 2];
      ua[u] = s[2 * PW_ARR + so + c * 2], ub[u] = s[3 * PW_ARR + so + c * 2];
      bp[u] = s[O_T + c * 4 + 2];  // zero beyond the last column (staged zeros): x stays 0 until the first real one
    }
#pragma unroll
    for (int u = PW_CB / 2 - 1; u >= 0; --u) {
      const int i = b * PW_CB + 2 * u + L.p;
      x1 = fma(bp[u], x1, ya[u]);
      x2 = fma(bp[u], x2, yb[u]);
      if (L.rok && i < m) {
        a.ux.p[((size_t)L.r * a.ux.ld + i) * 2 + L.part] = ua[u] - x1;
        a.uy.p[((size_t)L.r * a.uy.ld + i) * 2 + L.part] = ub[u] - x2;
      }
    }
    __syncwarp();
  }
}
constexpr int PW_PRJ_SLOT = 4 * PW_ARR + PW_CB * 4 > 3 * PW_ARR + PW_CB * 8 ? 4 * PW_ARR + PW_CB * 4 : 3 * PW_ARR + PW_CB * 8;
constexpr int PW_PRJ_WARP = PW_D * PW_PRJ_SLOT;
constexpr int pw_fwd_slot(int na) { return na * PW_ARR + 8 * PW_PR + PW_CB * 10 + 8; }
constexpr int PW_BWD_SLOT = PW_ARR + 8 * PW_PR + PW_CB * 8;
constexpr int pw_warp_doubles(int na) { return PW_D * (pw_fwd_slot(na) > PW_BWD_SLOT ? pw_fwd_slot(na) : PW_BWD_SLOT); }
}  // namespace

// (one warp per block here: the ring of the five-array rhs assembly is 43 KB per warp, and five one-warp blocks fit an SM
// where two two-warp blocks do -- 2.96 -> 2.73 ms at 8192 x 8193)
constexpr int PW_HH_WPB = 1;
__global__ void __launch_bounds__(32 * PW_HH_WPB) pw_hholtz(PHholtzArgs3 a3, int warp_doubles) {
  RP_DYN_SMEM(double, smem);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (blockIdx.x * PW_HH_WPB + wib) * 8;
  double* ring = smem + wib * warp_doubles;
  // blockIdx.y selects the field; the branch is block-uniform and keeps the arguments direct constant-bank operands
  const PHholtzArgs& a = blockIdx.y == 0 ? a3.a[0] : (blockIdx.y == 1 ? a3.a[1] : a3.a[2]);
  if (r0 >= a.chat.rows) return;
  const double mu = __ldg(&a.m.lam[min(r0 + (lane >> 2), a.chat.rows - 1)]) + a.m.alpha;
  if (a.mode == 0) {
    pw_forward_pass<0>(a, ring, r0, lane, mu);
  } else if (a.mode == 1) {
    pw_diff_pass(a.pres, a.dyp, -a.dt * a.isy, ring, r0, lane);  // - dt/sy d/dy pres   (navier.rs:646)
    pw_pass_fence();
    pw_forward_pass<1>(a, ring, r0, lane, mu);
  } else {
    pw_forward_pass<2>(a, ring, r0, lane, mu);
  }
  pw_pass_fence();
  pw_backward_pass(a.out, a.m, a.k0, false, ring, r0, lane, mu);
  cp_async_wait<0>();
}

__global__ void __launch_bounds__(32 * PW_WPB) pw_divpois(PDivPoisArgs a, int warp_doubles) {
  RP_DYN_SMEM(double, smem);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (blockIdx.x * PW_WPB + wib) * 8;
  double* ring = smem + wib * warp_doubles;
  if (r0 >= a.ux.rows) return;
  const double mu = __ldg(&a.m.lam[min(r0 + (lane >> 2), a.ux.rows - 1)]) + a.m.alpha;
  pw_div_pass(a, ring, r0, lane);
  pw_pass_fence();
  PHholtzArgs h;  // forward sweep of the Poisson solve on the stored divergence
  h.chat = a.div, h.out = a.phi, h.m = a.m, h.ny = a.ny, h.k0 = a.k0;
  h.dt = 0.0, h.isx = 0.0, h.rs = nullptr;
  pw_forward_pass<3>(h, ring, r0, lane, mu);
  pw_pass_fence();
  pw_backward_pass(a.phi, a.m, a.k0, true, ring, r0, lane, mu);  // pres[1].vhat[[0,0]] = 0   (navier.rs:714)
  cp_async_wait<0>();
}

__global__ void __launch_bounds__(32 * PW_WPB) pw_project(PProjectArgs a) {
  RP_DYN_SMEM(double, smem);
  const int wib = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int r0 = (blockIdx.x * PW_WPB + wib) * 8;
  double* ring = smem + wib * PW_PRJ_WARP;
  if (r0 >= a.phi.rows) return;
  pw_project_pass1(a, ring, r0, lane);
  pw_pass_fence();
  pw_project_pass2(a, ring, r0, lane);
  pw_pass_fence();
  pw_project_pass3(a, ring, r0, lane);
  cp_async_wait<0>();
}

// ---------------------------------------------------------------------------------
// Which kernel runs a per-mode pass of `rows` complex rows.  The row sweeps are bound by the latency of one chain
// (a fixed ~0.12 us per column pair and pass, whatever the row count) until there are enough rows to fill the machine,
// the tile kernels by throughput at every size: measured cross-over on a B200 at ~420 rows for the Helmholtz pass and
// ~1200 rows for the divergence + Poisson pass (profiles/r3_pw_crossover.txt).  RUSTPDE_B200_PW=0 / 1 forces one of them.
bool pw_enabled(int rows, bool divpois) {
  const char* e = getenv("RUSTPDE_B200_PW");  // (read at every launch: the step is a captured graph, and tests toggle it)
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  return rows >= (divpois ? 1280 : 448);  // (three fields in one launch: navier.cu)
}
bool pw_project_enabled(int rows) {  // projection + pressure update: one field, two passes
  const char* e = getenv("RUSTPDE_B200_PW");
  if (e && e[0] == '0') return false;
  if (e && e[0] == '1') return true;
  // three latency-bound passes (~0.29 us per column pair in all, whatever the row count) against ~0.15 .. 0.8 us per row
  // of the tile kernel: measured on a B200 at 8192 x 8193 (4097 rows) 1.58 ms against 3.22 ms, at 2048 x 2049 (1025 rows)
  // 0.298 ms against 0.152 ms
  return rows >= 1536;
}

template <class K>
static void pw_prepare(K kern, int bytes) {
#ifndef RP_EMU
  RP_CUDA_CHECK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, bytes));
#else
  (void)kern;
  (void)bytes;
#endif
}
void launch_pw_hholtz(const PHholtzArgs3& a, int nb, cudaStream_t s) {
  constexpr int WD = pw_warp_doubles(5), SMEM = PW_HH_WPB * WD * 8;
  static unsigned long long init_ = 0;  // one bit per device
  if (first_use_on_device(init_)) pw_prepare(pw_hholtz, SMEM);
  const int nwarps = (a.a[0].chat.rows + 7) / 8;
  RP_LAUNCH(pw_hholtz, dim3((nwarps + PW_HH_WPB - 1) / PW_HH_WPB, nb), dim3(32 * PW_HH_WPB), (size_t)SMEM, s, a, WD);
}
// Coefficient rows of pw_project: nsd / nsl = Neumann stencil of phi along y (m entries), sd / sl = Dirichlet stencil of the
// velocity, fs / fp / bp = its pre-factored (S^T S) solve (tables.cu)
void pw_project_tables(int n, double isy, const std::vector<double>& nsd, const std::vector<double>& nsl, const std::vector<double>& sd,
                       const std::vector<double>& sl, const std::vector<double>& fs, const std::vector<double>& fp,
                       const std::vector<double>& bp, std::vector<double>& w1, std::vector<double>& w2) {
  const int m = n - 2;
  auto at = [](const std::vector<double>& v, int i) { return (i >= 0 && i < (int)v.size()) ? v[i] : 0.0; };
  w1.assign((size_t)n * 8, 0.0);
  w2.assign((size_t)m * 4, 0.0);
  for (int j = 0; j < n; ++j) {
    double* r = &w1[(size_t)j * 8];
    const double w = (j + 1 <= n - 1) ? 2.0 * (double)(j + 1) * isy : 0.0;
    if (j < m) r[0] = at(nsd, j);
    r[1] = at(nsl, j - 2);
    if (j + 1 < m) r[2] = w * at(nsd, j + 1);
    r[3] = w * at(nsl, j - 1);
    if (j < m) r[4] = at(sd, j), r[5] = at(sl, j);
  }
  for (int i = 0; i < m; ++i) {
    double* r = &w2[(size_t)i * 4];
    r[0] = at(fs, i), r[1] = at(fp, i), r[2] = at(bp, i);
  }
}
void launch_pw_project(const PProjectArgs& a, cudaStream_t s) {
  constexpr int SMEM = PW_WPB * PW_PRJ_WARP * 8;
  static unsigned long long init_ = 0;
  if (first_use_on_device(init_)) pw_prepare(pw_project, SMEM);
  const int nwarps = (a.phi.rows + 7) / 8;
  RP_LAUNCH(pw_project, dim3((nwarps + PW_WPB - 1) / PW_WPB, 1), dim3(32 * PW_WPB), (size_t)SMEM, s, a);
}
void launch_pw_divpois(const PDivPoisArgs& a, cudaStream_t s) {
  constexpr int WD = pw_warp_doubles(2), SMEM = PW_WPB * WD * 8;
  static unsigned long long init_ = 0;
  if (first_use_on_device(init_)) pw_prepare(pw_divpois, SMEM);
  const int nwarps = (a.ux.rows + 7) / 8;
  RP_LAUNCH(pw_divpois, dim3((nwarps + PW_WPB - 1) / PW_WPB, 1), dim3(32 * PW_WPB), (size_t)SMEM, s, a, WD);
}

}  // namespace fk
}  // namespace rp
