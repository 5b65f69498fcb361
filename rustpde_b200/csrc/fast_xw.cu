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

}  // namespace fk
}  // namespace rp
