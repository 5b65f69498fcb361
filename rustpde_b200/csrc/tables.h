// tables.h -- host-side construction of the device-resident plan tables:
// FFT twiddles / Bluestein chirps, DCT-I pre/post factors, Galerkin stencils,
// the pre-factored (S^T S) tridiagonal of from_ortho, the B2 preconditioner
// bands and the banded Helmholtz/Poisson operator diagonals (closed forms of
// SURVEY.md 8(a''), derived from funspace/src/chebyshev/ortho.rs:147-180 and
// composite_stencil.rs:117-171).
#pragma once
#include <map>
#include <memory>
#include <vector>

#include "lane_prog.h"

namespace rp {

enum BaseKind : int {
  BASE_CHEBYSHEV = 0,
  BASE_CHEB_DIRICHLET = 1,
  BASE_CHEB_NEUMANN = 2,
  BASE_CHEB_DIRICHLET_BC = 3,
  BASE_CHEB_NEUMANN_BC = 4,
  BASE_FOURIER_R2C = 5,
};

// owning device buffer
struct DevBuf {
  void* p = nullptr;
  size_t bytes = 0;
  DevBuf() {}
  explicit DevBuf(size_t b) : p(rt::dmalloc(b)), bytes(b) {}
  DevBuf(const DevBuf&) = delete;
  DevBuf& operator=(const DevBuf&) = delete;
  DevBuf(DevBuf&& o) noexcept : p(o.p), bytes(o.bytes) { o.p = nullptr; }
  DevBuf& operator=(DevBuf&& o) noexcept {
    if (this != &o) {
      rt::dfree(p);
      p = o.p;
      bytes = o.bytes;
      o.p = nullptr;
    }
    return *this;
  }
  ~DevBuf() { rt::dfree(p); }
  template <class T>
  T* as() const { return (T*)p; }
};

template <class T>
DevBuf upload(const std::vector<T>& v) {
  DevBuf b(v.size() * sizeof(T));
  if (!v.empty()) {
    rt::h2d(b.p, v.data(), v.size() * sizeof(T), 0);
    rt::sync(0);
  }
  return b;
}
template <class T>
DevBuf upload_struct(const T& v) {
  DevBuf b(sizeof(T));
  rt::h2d(b.p, &v, sizeof(T), 0);
  rt::sync(0);
  return b;
}

// banded diagonals of an (m x m) operator, offsets -2, 0, +2, +4
struct Diags {
  std::vector<double> low, dia, up1, up2;  // sizes m-2, m, m-2, m-4
  void resize(int m) {
    low.assign(m > 2 ? m - 2 : 0, 0.0);
    dia.assign(m, 0.0);
    up1.assign(m > 2 ? m - 2 : 0, 0.0);
    up2.assign(m > 4 ? m - 4 : 0, 0.0);
  }
};

struct FftTables {
  FftPlan plan;  // pointers into the buffers below
  DevBuf tw, chirp, bhat;
  DevBuf twc;      // per-span compact twiddles (tables.cu build_fft_tables)
  DevBuf bhat_dr;  // bhat in the digit-reversed order of the in-place DIF FFT, transposed [8][Lb/8] (fast.cuh dif_dit_mid)
  int fft_len() const { return plan.pow2 ? plan.L : plan.Lb; }
};
void build_fft_tables(int L, FftTables& out);

// One function-space base (funspace constructors, lib.rs:230-345)
struct Base {
  int kind, n, m;
  std::vector<double> x;  // coords()
  // composite stencil (d, l) (m entries)
  std::vector<double> sd, sl;
  // boundary stencils (kinds *_BC)
  double t0[2], t1[2];
  // closed-form operator bands (Chebyshev family)
  Diags A, C;                        // A = I2.S (mat_b), C = B2.S (mat_a)
  std::vector<double> b2lo, b2di, b2up;  // preconditioner rows r=0..n-3
  // device tables
  FftTables fft;       // complex DFT of length N = n-1 (Chebyshev) or n (Fourier)
  DevBuf d_dct;        // DctPlan
  DevBuf d_fft;        // FftPlan (Fourier)
  DevBuf d_sc;         // (sin, cos) table
  DevBuf d_sd, d_sl;   // stencil
  DevBuf d_tdma, d_tfs, d_tfp, d_tbp;
  DevBuf d_b2lo, d_b2di, d_b2up;
  bool is_cheb() const { return kind != BASE_FOURIER_R2C; }
  bool is_composite() const { return kind == BASE_CHEB_DIRICHLET || kind == BASE_CHEB_NEUMANN; }
  bool is_bc() const { return kind == BASE_CHEB_DIRICHLET_BC || kind == BASE_CHEB_NEUMANN_BC; }
  int fft_len() const { return fft.fft_len(); }
  int dft_len() const { return fft.plan.L; }
};
std::shared_ptr<Base> get_base(int kind, int n);

// Sweep of fdma.rs:73-82 on raw diagonals (in place).
void fdma_sweep(Diags& d);

// Device tables of a pre-swept Fdma (ADI flavour)
struct FdmaDev {
  DevBuf tab, fp, bs, bp1, bp2;
  int n = 0;
};
void build_fdma_dev(const Diags& swept, FdmaDev& out);

// Per-lane (A + (lam+alpha) C): raw bands + lam + table of swept pivot reciprocals
struct FdmaModeDev {
  DevBuf tab, a_low, a_dia, a_up1, a_up2, c_low, c_dia, c_up1, c_up2, lam, inv;
  int n = 0, nlanes = 0;
  long long inv_ld = 0;
  double alpha = 0.0;
};
void build_fdma_mode_dev(const Diags& A, const Diags& C, const std::vector<double>& lam, double alpha, FdmaModeDev& out);

}  // namespace rp
