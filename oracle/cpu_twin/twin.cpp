// twin.cpp -- C++ lane-parallel CPU restatement of rustpde's Navier2D::update().
//
// TEST / BASELINE INFRASTRUCTURE ONLY (see oracle/__init__.py): it is the CPU baseline that bench.py times
// (`cpu_baseline`, `--impl reference`) and is itself checked against the numpy oracle (tests/test_cpu_twin.py).
// Nothing under rustpde_b200/ may use it.
//
// The reference is pure Rust and cannot be built in this image.  This file restates the same algorithm in the same
// op order and with the same parallel structure: every 1-D operator runs lane by lane along an axis of a row-major
// [n0, n1] array, the lanes of one operator call are distributed over the host threads (OpenMP here, rayon's
// par_for_each over lanes there: funspace/src/chebyshev/ortho.rs:352,406, composite.rs:255-316, src/solver/fdma.rs:161,
// matvec.rs:212), every stage allocates its output like the reference's `to_owned()` / fresh Array2, and the dense
// contractions of the fast diagonalisation go through BLAS dgemm on ONE thread (README.md:10-16: OPENBLAS_NUM_THREADS=1).
//   navier.rs:737-765 update, 538-616 conv_*, 622-674 solve_*, 683-721 projection; conv_term.rs:22-42;
//   space2.rs:182-356; ortho.rs:107-125, 337-407; composite_stencil.rs:207-276; linalg.rs:14-57; r2c.rs:88-99, 250-303;
//   fdma.rs:73-118; matvec.rs:172-193; hholtz_adi.rs:98-130; hholtz.rs:156-197; poisson.rs:131-149; fdma_tensor.rs:195-234.
// Third-party arithmetic (ndrustfft / rustfft / rustdct / realfft) is replaced by an own FFT: radix-2 for powers of
// two, Bluestein otherwise; DCT-I through a complex FFT of length 2(n-1) exactly like rustdct's Dct1ConvertToFft.
#include <dlfcn.h>
#include <omp.h>

#include <algorithm>
#include <cmath>
#include <complex>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <memory>
#include <stdexcept>
#include <vector>

typedef std::complex<double> cd;

// ---------------------------------------------------------------------------------------------------
// FFT
// ---------------------------------------------------------------------------------------------------
struct Fft {
  int n = 0;
  bool pow2 = false;
  std::vector<cd> tw;       // pow2: exp(-2 pi i k / n), k < n/2
  std::vector<int> rev;     // pow2: bit reversal
  int L = 0;                // Bluestein length
  std::vector<cd> chirp, bhat;
  std::unique_ptr<Fft> inner;
  explicit Fft(int n_) : n(n_) {
    pow2 = (n & (n - 1)) == 0;
    if (pow2) {
      tw.resize(std::max(1, n / 2));
      for (int k = 0; k < n / 2; ++k) tw[k] = std::polar(1.0, -2.0 * M_PI * k / n);
      rev.resize(n);
      int lg = 0;
      while ((1 << lg) < n) ++lg;
      for (int i = 0; i < n; ++i) {
        int r = 0;
        for (int b = 0; b < lg; ++b)
          if (i >> b & 1) r |= 1 << (lg - 1 - b);
        rev[i] = r;
      }
    } else {
      L = 1;
      while (L < 2 * n - 1) L <<= 1;
      inner.reset(new Fft(L));
      chirp.resize(n);
      for (int j = 0; j < n; ++j) {
        const long long jj = ((long long)j * j) % (2LL * n);
        chirp[j] = std::polar(1.0, -M_PI * (double)jj / n);
      }
      bhat.assign(L, cd(0, 0));
      bhat[0] = std::conj(chirp[0]);
      for (int j = 1; j < n; ++j) bhat[j] = bhat[L - j] = std::conj(chirp[j]);
      std::vector<cd> w(L);
      inner->forward(bhat.data(), w.data());
    }
  }
  int work_len() const { return pow2 ? 0 : L; }
  // in place, forward (exp(-i...)); work: work_len() elements
  void forward(cd* x, cd* work) const {
    if (pow2) {
      for (int i = 0; i < n; ++i)
        if (i < rev[i]) std::swap(x[i], x[rev[i]]);
      for (int len = 2; len <= n; len <<= 1) {
        const int half = len / 2, step = n / len;
        for (int s = 0; s < n; s += len)
          for (int k = 0; k < half; ++k) {
            const cd t = x[s + k + half] * tw[k * step];
            x[s + k + half] = x[s + k] - t;
            x[s + k] += t;
          }
      }
      return;
    }
    for (int j = 0; j < n; ++j) work[j] = x[j] * chirp[j];
    for (int j = n; j < L; ++j) work[j] = cd(0, 0);
    inner->forward(work, nullptr);
    for (int j = 0; j < L; ++j) work[j] = std::conj(work[j] * bhat[j]);
    inner->forward(work, nullptr);  // inverse through conjugation
    const double s = 1.0 / L;
    for (int j = 0; j < n; ++j) x[j] = std::conj(work[j]) * s * chirp[j];
  }
  void inverse(cd* x, cd* work) const {  // unnormalised
    for (int j = 0; j < n; ++j) x[j] = std::conj(x[j]);
    forward(x, work);
    for (int j = 0; j < n; ++j) x[j] = std::conj(x[j]);
  }
};

struct Work {  // per-thread scratch
  std::vector<cd> a, b;
  std::vector<double> r;
  std::vector<cd> lane_in, lane_out;
};

// ---------------------------------------------------------------------------------------------------
// bases
// ---------------------------------------------------------------------------------------------------
enum { CHEB = 0, CD = 1, CN = 2, FOURIER = 5 };
struct Diags {
  std::vector<double> low, dia, up1, up2;
  void resize(int m) {
    low.assign(std::max(0, m - 2), 0.0), dia.assign(m, 0.0), up1.assign(std::max(0, m - 2), 0.0), up2.assign(std::max(0, m - 4), 0.0);
  }
};
struct Base {
  int kind, n, m;
  std::vector<double> sd, sl, off, mainv;  // stencil and S^T S (composite_stencil.rs:117-171)
  std::vector<double> fwd_sc, bwd_sc;      // (-1)^k/(n-1), (-1)^k/2 (ortho.rs:51-58)
  std::unique_ptr<Fft> fft;                // length 2(n-1) (Chebyshev family) or n (Fourier)
  Diags A, C;                              // A = I2 S, C = B2 S (SURVEY 8a'')
  Diags pre;                               // (n-2) x n preconditioner rows as MatVecFdma bands
  bool cheb() const { return kind != FOURIER; }
  bool composite() const { return kind == CD || kind == CN; }
  Base(int kind_, int n_) : kind(kind_), n(n_) {
    if (kind == FOURIER) {
      m = n / 2 + 1;
      fft.reset(new Fft(n));
      return;
    }
    m = composite() ? n - 2 : n;
    fft.reset(new Fft(2 * (n - 1)));
    fwd_sc.resize(n), bwd_sc.resize(n);
    for (int k = 0; k < n; ++k) {
      const double sg = (k & 1) ? -1.0 : 1.0;
      fwd_sc[k] = sg * (1.0 / (n - 1));
      bwd_sc[k] = sg / 2.0;
    }
    if (composite()) {
      sd.assign(m, 1.0), sl.assign(m, -1.0);
      if (kind == CN)
        for (int k = 0; k < m; ++k) sl[k] = -1.0 * ((double)k * k) / (((double)k + 2.0) * ((double)k + 2.0));
      mainv.resize(m), off.resize(std::max(0, m - 2));
      for (int i = 0; i < m; ++i) mainv[i] = sd[i] * sd[i] + sl[i] * sl[i];
      for (int i = 0; i + 2 < m; ++i) off[i] = sd[i + 2] * sl[i];
    }
    // B2 = pinv(n, 2) rows (ortho.rs:160-171)
    std::vector<double> lo(n, 0.0), di(n, 0.0), up(n, 0.0);
    lo[2] = 0.25;
    for (int i = 3; i < n; ++i) lo[i] = 1.0 / (4.0 * i * (i - 1.0));
    for (int i = 2; i < n - 2; ++i) di[i] = -1.0 / (2.0 * ((double)i * i - 1.0));
    for (int i = 2; i < n - 4; ++i) up[i] = 1.0 / (4.0 * i * (i + 1.0));
    const int mm = n - 2;
    pre.low.assign(mm, 0.0), pre.dia.assign(mm, 0.0), pre.up1.assign(mm, 0.0), pre.up2.assign(mm, 0.0);
    for (int r = 0; r < mm; ++r) {
      const int i = r + 2;
      pre.dia[r] = lo[i];
      if (r < mm - 2) pre.up1[r] = di[i];
      if (r < mm - 4) pre.up2[r] = up[i];
    }
    auto s_entry = [&](int row, int col) -> double {
      if (col < 0 || col >= mm) return 0.0;
      if (kind == CHEB) return row == col + 2 ? 1.0 : 0.0;
      if (row == col) return sd[col];
      if (row == col + 2) return sl[col];
      return 0.0;
    };
    auto c_entry = [&](int r, int c) {
      const int i = r + 2;
      double tot = 0.0;
      const int ks[3] = {i - 2, i, i + 2};
      const double bs[3] = {lo[i], di[i], (i + 2 < n) ? up[i] : 0.0};
      for (int q = 0; q < 3; ++q)
        if (ks[q] >= 0 && ks[q] < n && bs[q] != 0.0) tot += bs[q] * s_entry(ks[q], c);
      return tot;
    };
    A.resize(mm), C.resize(mm);
    for (int r = 0; r < mm; ++r) {
      A.dia[r] = s_entry(r + 2, r), C.dia[r] = c_entry(r, r);
      if (r + 2 < mm) {
        A.up1[r] = s_entry(r + 2, r + 2), C.up1[r] = c_entry(r, r + 2);
        A.low[r] = s_entry(r + 4, r), C.low[r] = c_entry(r + 2, r);
      }
      if (r + 4 < mm) A.up2[r] = s_entry(r + 2, r + 4), C.up2[r] = c_entry(r, r + 4);
    }
  }
};

// ---------------------------------------------------------------------------------------------------
// lane drivers: f(const T* x, T* y, Work&) maps a lane of len_in to a lane of len_out along `axis`
// ---------------------------------------------------------------------------------------------------
template <class T>
struct Arr {
  int r = 0, c = 0;
  std::vector<T> d;
  Arr() {}
  Arr(int r_, int c_) : r(r_), c(c_), d((size_t)r_ * c_, T(0)) {}
  T& at(int i, int j) { return d[(size_t)i * c + j]; }
  const T& at(int i, int j) const { return d[(size_t)i * c + j]; }
};
static std::vector<Work> g_work;
static Work& my_work() { return g_work[omp_get_thread_num()]; }

static double g_t_axis[2] = {0, 0};
struct AxisTimer {
  int a;
  double t0;
  explicit AxisTimer(int a_) : a(a_), t0(omp_get_wtime()) {}
  ~AxisTimer() { g_t_axis[a] += omp_get_wtime() - t0; }
};
template <class Ti, class To, class F>
Arr<To> lanes(const Arr<Ti>& in, int axis, int len_out, F f) {
  AxisTimer tm(axis);
  Arr<To> out(axis == 0 ? len_out : in.r, axis == 1 ? len_out : in.c);
  if (axis == 1) {
#pragma omp parallel for schedule(static)
    for (int i = 0; i < in.r; ++i) f(&in.d[(size_t)i * in.c], &out.d[(size_t)i * out.c], my_work());
  } else {
    // lanes along the strided axis are copied to contiguous buffers (like ndrustfft); 8 adjacent lanes at a time so that
    // every cache line of the array is touched once
    constexpr int B = 8;
#pragma omp parallel for schedule(static)
    for (int j0 = 0; j0 < in.c; j0 += B) {
      const int nb = std::min(B, in.c - j0);
      std::vector<Ti> x((size_t)B * in.r);
      std::vector<To> y((size_t)B * len_out);
      for (int i = 0; i < in.r; ++i)
        for (int q = 0; q < nb; ++q) x[(size_t)q * in.r + i] = in.d[(size_t)i * in.c + j0 + q];
      for (int q = 0; q < nb; ++q) f(&x[(size_t)q * in.r], &y[(size_t)q * len_out], my_work());
      for (int i = 0; i < len_out; ++i)
        for (int q = 0; q < nb; ++q) out.d[(size_t)i * out.c + j0 + q] = y[(size_t)q * len_out + i];
    }
  }
  return out;
}

// DCT-I (unnormalised, scipy type 1) of a lane through a complex FFT of length 2(n-1)
template <class T>
static void dct1(const Base& b, const T* x, T* y, Work& w) {
  const int n = b.n, L = 2 * (n - 1);
  w.a.resize(L);
  w.b.resize(b.fft->work_len());
  for (int j = 0; j < n; ++j) w.a[j] = cd(x[j]);
  for (int j = 1; j < n - 1; ++j) w.a[L - j] = cd(x[j]);
  b.fft->forward(w.a.data(), w.b.data());
  for (int k = 0; k < n; ++k) {
    if constexpr (std::is_same<T, double>::value)
      y[k] = w.a[k].real();
    else
      y[k] = w.a[k];
  }
}
template <class T>
static void cheb_forward(const Base& b, const T* x, T* y, Work& w) {  // ortho.rs:337-360
  dct1(b, x, y, w);
  for (int k = 0; k < b.n; ++k) y[k] *= b.fwd_sc[k];
  y[0] *= 0.5, y[b.n - 1] *= 0.5;
}
template <class T>
static void cheb_backward(const Base& b, const T* x, T* y, Work& w) {  // ortho.rs:383-407
  std::vector<T> t(x, x + b.n);
  for (int k = 0; k < b.n; ++k) t[k] *= b.bwd_sc[k];
  t[0] *= 2.0, t[b.n - 1] *= 2.0;
  dct1(b, t.data(), y, w);
}
template <class T>
static void stencil_mul(const Base& b, const T* c, T* p) {  // composite_stencil.rs:207-229
  const int n = b.n;
  p[0] = c[0] * b.sd[0], p[1] = c[1] * b.sd[1];
  for (int i = 2; i < n - 2; ++i) p[i] = c[i] * b.sd[i] + c[i - 2] * b.sl[i - 2];
  p[n - 2] = c[n - 4] * b.sl[n - 4], p[n - 1] = c[n - 3] * b.sl[n - 3];
}
template <class T>
static void stencil_solve(const Base& b, const T* p, T* x) {  // composite_stencil.rs:250-276 + linalg.rs:14-57
  const int m = b.m;
  std::vector<T> d(m), g(m);
  std::vector<double> w(std::max(1, m - 2));
  for (int i = 0; i < m; ++i) d[i] = p[i] * b.sd[i] + p[i + 2] * b.sl[i];
  const std::vector<double>&a = b.off, &bb = b.mainv, &c = b.off;
  w[0] = c[0] / bb[0];
  g[0] = d[0] / bb[0];
  if (c.size() > 1) w[1] = c[1] / bb[1];
  g[1] = d[1] / bb[1];
  for (int i = 2; i < m - 2; ++i) w[i] = c[i] / (bb[i] - a[i - 2] * w[i - 2]);
  for (int i = 2; i < m; ++i) g[i] = (d[i] - g[i - 2] * a[i - 2]) / (bb[i] - a[i - 2] * w[i - 2]);
  x[m - 1] = g[m - 1], x[m - 2] = g[m - 2];
  for (int i = m - 2; i > 0; --i) x[i - 1] = g[i - 1] - x[i + 1] * w[i - 1];
}
template <class T>
static void cheb_diff(int n, T* d, int times) {  // ortho.rs:107-125
  for (int t = 0; t < times; ++t) {
    d[0] = d[1];
    for (int i = 1; i < n - 1; ++i) d[i] = 2.0 * (double)(i + 1) * d[i + 1];
    d[n - 1] = T(0);
    for (int i = n - 3; i > 0; --i) d[i] = d[i] + d[i + 2];
    d[0] = d[0] + d[2] / 2.0;
  }
}

// ---- axis operators on arrays (each allocates its result, like the reference) ----
template <class T>
Arr<T> to_ortho_axis(const Base& b, const Arr<T>& a, int axis) {
  if (!b.composite()) return a;
  return lanes<T, T>(a, axis, b.n, [&](const T* x, T* y, Work&) { stencil_mul(b, x, y); });
}
template <class T>
Arr<T> from_ortho_axis(const Base& b, const Arr<T>& a, int axis) {
  if (!b.composite()) return a;
  return lanes<T, T>(a, axis, b.m, [&](const T* x, T* y, Work&) { stencil_solve(b, x, y); });
}
template <class T>
Arr<T> diff_axis(const Base& b, const Arr<T>& a, int axis, int times) {  // composite.rs:556-570 / r2c.rs:88-99
  if (b.kind == FOURIER) {
    Arr<T> out = a;
    if constexpr (!std::is_same<T, double>::value) {
      for (int t = 0; t < times; ++t)
#pragma omp parallel for schedule(static)
        for (int i = 0; i < out.r; ++i)
          for (int j = 0; j < out.c; ++j) out.at(i, j) *= cd(0.0, (double)(axis == 0 ? i : j));
    }
    return out;
  }
  Arr<T> o = to_ortho_axis(b, a, axis);
  if (times == 0) return o;
  return lanes<T, T>(o, axis, b.n, [&](const T* x, T* y, Work&) {
    std::copy(x, x + b.n, y);
    cheb_diff(b.n, y, times);
  });
}

// ---------------------------------------------------------------------------------------------------
// Space2 / Field2 (space2.rs:182-356, field.rs:103-129).  P = double, S = double or complex
// ---------------------------------------------------------------------------------------------------
template <class S>
struct Field {
  const Base *b0, *b1;
  Arr<double> v;
  Arr<S> vhat;
  Field(const Base* x, const Base* y) : b0(x), b1(y), v(x->n, y->n), vhat(x->m, y->m) {}
  void forward() {  // y first, then x
    Arr<double> buf = lanes<double, double>(v, 1, b1->m, [&](const double* x, double* y, Work& w) {
      std::vector<double> t(b1->n);
      cheb_forward(*b1, x, t.data(), w);
      if (b1->composite())
        stencil_solve(*b1, t.data(), y);
      else
        std::copy(t.begin(), t.end(), y);
    });
    if (b0->kind == FOURIER) {
      if constexpr (!std::is_same<S, double>::value)
        vhat = lanes<double, cd>(buf, 0, b0->m, [&](const double* x, cd* y, Work& w) {  // r2c.rs:250-264
          const int n = b0->n;
          w.a.resize(n), w.b.resize(b0->fft->work_len());
          for (int j = 0; j < n; ++j) w.a[j] = cd(x[j], 0.0);
          b0->fft->forward(w.a.data(), w.b.data());
          for (int k = 0; k < b0->m; ++k) y[k] = w.a[k];
        });
    } else {
      if constexpr (std::is_same<S, double>::value)
        vhat = lanes<double, double>(buf, 0, b0->m, [&](const double* x, double* y, Work& w) {
          std::vector<double> t(b0->n);
          cheb_forward(*b0, x, t.data(), w);
          if (b0->composite())
            stencil_solve(*b0, t.data(), y);
          else
            std::copy(t.begin(), t.end(), y);
        });
    }
  }
  void backward() {  // x first, then y
    Arr<double> buf;
    if (b0->kind == FOURIER) {
      if constexpr (!std::is_same<S, double>::value)
        buf = lanes<cd, double>(vhat, 0, b0->n, [&](const cd* x, double* y, Work& w) {  // r2c.rs:289-303
          const int n = b0->n;
          w.a.resize(n), w.b.resize(b0->fft->work_len());
          for (int k = 0; k < b0->m; ++k) w.a[k] = x[k];
          for (int k = 1; k < n - b0->m + 1; ++k) w.a[n - k] = std::conj(x[k]);
          b0->fft->inverse(w.a.data(), w.b.data());
          for (int j = 0; j < n; ++j) y[j] = w.a[j].real() / n;
        });
    } else {
      if constexpr (std::is_same<S, double>::value)
        buf = lanes<double, double>(vhat, 0, b0->n, [&](const double* x, double* y, Work& w) {
          std::vector<double> t(b0->n);
          if (b0->composite())
            stencil_mul(*b0, x, t.data());
          else
            std::copy(x, x + b0->n, t.begin());
          cheb_backward(*b0, t.data(), y, w);
        });
    }
    v = lanes<double, double>(buf, 1, b1->n, [&](const double* x, double* y, Work& w) {
      std::vector<double> t(b1->n);
      if (b1->composite())
        stencil_mul(*b1, x, t.data());
      else
        std::copy(x, x + b1->n, t.begin());
      cheb_backward(*b1, t.data(), y, w);
    });
  }
  Arr<S> to_ortho() const { return to_ortho_axis(*b1, to_ortho_axis(*b0, vhat, 0), 1); }
  void from_ortho(const Arr<S>& a) { vhat = from_ortho_axis(*b1, from_ortho_axis(*b0, a, 0), 1); }
  Arr<S> gradient(int d0, int d1, const double* scale) const {  // space2.rs:247-264
    Arr<S> out = diff_axis(*b1, diff_axis(*b0, vhat, 0, d0), 1, d1);
    if (scale) {
      const double sc = std::pow(scale[0], d0) * std::pow(scale[1], d1);
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < out.d.size(); ++k) out.d[k] = out.d[k] / sc;
    }
    return out;
  }
};

// ---------------------------------------------------------------------------------------------------
// solvers
// ---------------------------------------------------------------------------------------------------
static void fdma_sweep(Diags& d) {  // fdma.rs:73-82
  const int n = (int)d.dia.size();
  for (int i = 2; i < n; ++i) {
    d.low[i - 2] /= d.dia[i - 2];
    d.dia[i] -= d.low[i - 2] * d.up1[i - 2];
    if (i < n - 2) d.up1[i] -= d.low[i - 2] * d.up2[i - 2];
  }
}
template <class T>
static void fdma_lane(const Diags& d, T* x) {  // fdma.rs:101-118
  const int n = (int)d.dia.size();
  for (int i = 2; i < n; ++i) x[i] = x[i] - x[i - 2] * d.low[i - 2];
  x[n - 1] = x[n - 1] / d.dia[n - 1];
  x[n - 2] = x[n - 2] / d.dia[n - 2];
  x[n - 3] = (x[n - 3] - x[n - 1] * d.up1[n - 3]) / d.dia[n - 3];
  x[n - 4] = (x[n - 4] - x[n - 2] * d.up1[n - 4]) / d.dia[n - 4];
  for (int i = n - 5; i >= 0; --i) x[i] = (x[i] - x[i + 2] * d.up1[i] - x[i + 4] * d.up2[i]) / d.dia[i];
}
template <class T>
Arr<T> matvec_axis(const Diags& p, const Arr<T>& a, int axis) {  // matvec.rs:172-230
  const int n = (int)p.dia.size();
  return lanes<T, T>(a, axis, n, [&](const T* x, T* y, Work&) {
    for (int i = 0; i < n; ++i) {
      T o = x[i] * p.dia[i];
      if (i > 1) o += x[i - 2] * p.low[i];
      if (i < n - 2) o += x[i + 2] * p.up1[i];
      if (i < n - 4) o += x[i + 4] * p.up2[i];
      y[i] = o;
    }
  });
}
template <class T>
Arr<T> fdma_axis(const Diags& d, const Arr<T>& a, int axis) {  // fdma.rs:161-175
  const int n = (int)d.dia.size();
  return lanes<T, T>(a, axis, n, [&](const T* x, T* y, Work&) {
    std::copy(x, x + n, y);
    fdma_lane(d, y);
  });
}
static Diags combine(const Diags& a, double sa, const Diags& b, double sb) {
  Diags r;
  r.resize((int)a.dia.size());
  for (size_t i = 0; i < r.low.size(); ++i) r.low[i] = a.low[i] * sa + b.low[i] * sb;
  for (size_t i = 0; i < r.dia.size(); ++i) r.dia[i] = a.dia[i] * sa + b.dia[i] * sb;
  for (size_t i = 0; i < r.up1.size(); ++i) r.up1[i] = a.up1[i] * sa + b.up1[i] * sb;
  for (size_t i = 0; i < r.up2.size(); ++i) r.up2[i] = a.up2[i] * sa + b.up2[i] * sb;
  return r;
}

typedef void (*dgemm_t)(const char*, const char*, const int64_t*, const int64_t*, const int64_t*, const double*, const double*,
                        const int64_t*, const double*, const int64_t*, const double*, double*, const int64_t*);
typedef void (*dgemm32_t)(const char*, const char*, const int*, const int*, const int*, const double*, const double*, const int*,
                          const double*, const int*, const double*, double*, const int*);
static void* g_dgemm = nullptr;
static bool g_dgemm64 = false;
// C[M x N] = A[M x K] . B[K x N], row-major, through Fortran dgemm (C^T = B^T A^T); own loop when no BLAS was given
static void gemm(int M, int N, int K, const double* A, const double* B, double* C) {
  if (g_dgemm) {
    const double one = 1.0, zero = 0.0;
    if (g_dgemm64) {
      const int64_t m = N, n = M, k = K;
      ((dgemm_t)g_dgemm)("N", "N", &m, &n, &k, &one, B, &m, A, &k, &zero, C, &m);
    } else {
      const int m = N, n = M, k = K;
      ((dgemm32_t)g_dgemm)("N", "N", &m, &n, &k, &one, B, &m, A, &k, &zero, C, &m);
    }
    return;
  }
#pragma omp parallel for schedule(static)
  for (int i = 0; i < M; ++i) {
    double* c = C + (size_t)i * N;
    std::fill(c, c + N, 0.0);
    for (int k = 0; k < K; ++k) {
      const double a = A[(size_t)i * K + k];
      const double* b = B + (size_t)k * N;
      for (int j = 0; j < N; ++j) c[j] += a * b[j];
    }
  }
}

template <class S>
struct Solver {
  // kind 0: HholtzAdi, 1: FdmaTensor (Hholtz / Poisson)
  int kind = 0;
  const Base *b0 = nullptr, *b1 = nullptr;
  Diags fx, fy;             // ADI: swept (C - c A) per axis
  Diags ay, cy;             // tensor: raw A_y (scaled), C_y
  std::vector<double> lam;  // tensor: eigenvalues (x axis)
  double alpha = 0.0;
  std::vector<double> P, Q;  // tensor, Chebyshev x: fwd = Q^-1 Cx^-1, bwd = Q  (m0 x m0)
  Arr<S> solve(const Arr<S>& input) const {
    Arr<S> rhs = b0->cheb() ? matvec_axis(b0->pre, input, 0) : input;  // hholtz_adi.rs:108-113 / hholtz.rs:166-175
    rhs = matvec_axis(b1->pre, rhs, 1);
    if (kind == 0) return fdma_axis(fy, fdma_axis(fx, rhs, 0), 1);  // hholtz_adi.rs:128-129
    // fdma_tensor.rs:195-234
    Arr<S> out = rhs;
    const int m0 = rhs.r, n1 = rhs.c;
    if (!P.empty()) {
      if constexpr (std::is_same<S, double>::value) gemm(m0, n1, m0, P.data(), rhs.d.data(), out.d.data());
    }
#pragma omp parallel for schedule(static)
    for (int i = 0; i < m0; ++i) {  // fresh Fdma = A + C * (lam_i + alpha), sweep, solve (hholtz.rs:182-190)
      Diags f = combine(ay, 1.0, cy, lam[i] + alpha);
      fdma_sweep(f);
      fdma_lane(f, &out.d[(size_t)i * n1]);
    }
    if (!Q.empty()) {
      if constexpr (std::is_same<S, double>::value) {
        Arr<S> o2(m0, n1);
        gemm(m0, n1, m0, Q.data(), out.d.data(), o2.d.data());
        return o2;
      }
    }
    return out;
  }
};

// ---------------------------------------------------------------------------------------------------
// Navier2D
// ---------------------------------------------------------------------------------------------------
template <class T>
static void axpy(Arr<T>& y, double a, const Arr<T>& x) {
#pragma omp parallel for schedule(static)
  for (size_t k = 0; k < y.d.size(); ++k) y.d[k] += x.d[k] * a;
}

template <class S>
struct Navier {
  int nx, ny;
  bool periodic;
  double dt, nu, ka, scale[2], time = 0.0;
  bool dealias = true;
  std::vector<std::unique_ptr<Base>> bases;
  std::unique_ptr<Field<S>> ux, uy, temp, pres0, pres1, field, fieldbc;  // fieldbc: ortho x ortho holder of the BC coefficients
  Solver<S> sol[4];
  const Base* base(int kind, int n) {
    for (auto& b : bases)
      if (b->kind == kind && b->n == n) return b.get();
    bases.emplace_back(new Base(kind, n));
    return bases.back().get();
  }
  Navier(int nx_, int ny_, double ra, double pr, double dt_, double aspect, bool adiabatic, bool periodic_, const double* lam,
         const double* q, const double* p)
      : nx(nx_), ny(ny_), periodic(periodic_), dt(dt_) {
    scale[0] = aspect, scale[1] = 1.0;
    const double h = scale[1] * 2.0;
    nu = std::sqrt(pr / (ra / std::pow(h, 3.0)));
    ka = std::sqrt(1.0 / ((ra / std::pow(h, 3.0)) * pr));
    const int kxu = periodic ? FOURIER : CD, kxt = periodic ? FOURIER : (adiabatic ? CN : CD);
    const int kxo = periodic ? FOURIER : CHEB, kxn = periodic ? FOURIER : CN;
    ux.reset(new Field<S>(base(kxu, nx), base(CD, ny)));
    uy.reset(new Field<S>(base(kxu, nx), base(CD, ny)));
    temp.reset(new Field<S>(base(kxt, nx), base(CD, ny)));
    pres0.reset(new Field<S>(base(kxo, nx), base(CHEB, ny)));
    pres1.reset(new Field<S>(base(kxn, nx), base(CN, ny)));
    field.reset(new Field<S>(base(kxo, nx), base(CHEB, ny)));
    fieldbc.reset(new Field<S>(base(kxo, nx), base(CHEB, ny)));
    // bc_rbc (navier.rs:314-332, 474-492): only T_1(y) is present in the ortho basis
    fieldbc->vhat.at(0, 1) = periodic ? S(-0.5 * nx) : S(-0.5);
    const double sx2 = scale[0] * scale[0], sy2 = scale[1] * scale[1];
    Field<S>* fl[3] = {ux.get(), uy.get(), temp.get()};
    const double cc[3] = {nu, nu, ka};
    for (int f = 0; f < 3; ++f) {
      Solver<S>& s = sol[f];
      s.b0 = fl[f]->b0, s.b1 = fl[f]->b1;
      const double cx = dt * cc[f] / sx2, cy = dt * cc[f] / sy2;
      if (!periodic) {  // HholtzAdi: mat = C - c A, pre-swept (hholtz_adi.rs:54-55)
        s.kind = 0;
        s.fx = combine(s.b0->C, 1.0, s.b0->A, -cx);
        s.fy = combine(s.b1->C, 1.0, s.b1->A, -cy);
        fdma_sweep(s.fx), fdma_sweep(s.fy);
      } else {  // Hholtz: a = -c A, c = C, alpha = 1; lam_k = -(-k^2) c
        s.kind = 1;
        s.ay = combine(s.b1->A, -cy, s.b1->A, 0.0);
        s.cy = s.b1->C;
        s.alpha = 1.0;
        s.lam.resize(s.b0->m);
        for (int k = 0; k < s.b0->m; ++k) s.lam[k] = -1.0 * (-(double)k * k) * cx;
      }
    }
    {  // Poisson (poisson.rs:50-91)
      Solver<S>& s = sol[3];
      s.kind = 1;
      s.b0 = pres1->b0, s.b1 = pres1->b1;
      s.ay = combine(s.b1->A, 1.0 / sy2, s.b1->A, 0.0);
      s.cy = s.b1->C;
      s.alpha = 0.0;
      const int m0 = s.b0->m;
      s.lam.resize(m0);
      if (periodic) {
        for (int k = 0; k < m0; ++k) s.lam[k] = (-(double)k * k) / sx2;
      } else {
        if (!lam || !q || !p) throw std::runtime_error("confined twin needs the eigen set-up data (lam, Q, P)");
        s.lam.assign(lam, lam + m0);
        s.Q.assign(q, q + (size_t)m0 * m0);
        s.P.assign(p, p + (size_t)m0 * m0);
      }
      if (std::fabs(s.lam[0]) < 1e-10)  // poisson.rs:80-83
        for (auto& l : s.lam) l -= 1e-10;
    }
  }
  // navier.rs:1035-1076
  void apply_ic(Field<S>& f, double amp, double m, double n, bool sin_cos) {
    const Base &bx = *f.b0, &by = *f.b1;
    std::vector<double> x(bx.n), y(by.n);
    for (int i = 0; i < bx.n; ++i) x[i] = bx.cheb() ? -std::sin(M_PI * ((bx.n - 1) - 2.0 * i) / (2.0 * (bx.n - 1))) : 2.0 * M_PI / bx.n * i;
    for (int j = 0; j < by.n; ++j) y[j] = -std::sin(M_PI * ((by.n - 1) - 2.0 * j) / (2.0 * (by.n - 1)));
    for (int i = 0; i < bx.n; ++i)
      for (int j = 0; j < by.n; ++j) {
        const double xs = (x[i] - x[0]) / (x[bx.n - 1] - x[0]), ys = (y[j] - y[0]) / (y[by.n - 1] - y[0]);
        f.v.at(i, j) = sin_cos ? amp * std::sin(M_PI * m * xs) * std::cos(M_PI * n * ys) : amp * std::cos(M_PI * m * xs) * std::sin(M_PI * n * ys);
      }
    f.forward();
  }
  Arr<double> conv_term(const Field<S>& f, const Arr<double>& u, int d0, int d1) {  // conv_term.rs:22-42
    field->vhat = f.gradient(d0, d1, scale);
    field->backward();
    Arr<double> out(u.r, u.c);
#pragma omp parallel for schedule(static)
    for (size_t k = 0; k < out.d.size(); ++k) out.d[k] = u.d[k] * field->v.d[k];
    return out;
  }
  Arr<S> finish_conv(Arr<double>& conv) {  // navier.rs:562-568
    field->v = conv;
    field->forward();
    if (dealias) {  // navier.rs:1022-1032
      const int cx = field->vhat.r * 2 / 3, cy = field->vhat.c * 2 / 3;
      for (int i = 0; i < field->vhat.r; ++i)
        for (int j = 0; j < field->vhat.c; ++j)
          if (i >= cx || j >= cy) field->vhat.at(i, j) = S(0);
    }
    return field->vhat;
  }
  Arr<S> conv_of(const Field<S>& f, const Arr<double>& uxp, const Arr<double>& uyp, bool with_bc) {
    Arr<double> conv = conv_term(f, uxp, 1, 0);
    axpy(conv, 1.0, conv_term(f, uyp, 0, 1));
    if (with_bc) {
      axpy(conv, 1.0, conv_term(*fieldbc, uxp, 1, 0));
      axpy(conv, 1.0, conv_term(*fieldbc, uyp, 0, 1));
    }
    return finish_conv(conv);
  }
  void update() {  // navier.rs:737-765
    Arr<S> that = temp->to_ortho();
    axpy(that, 1.0, fieldbc->to_ortho());
    ux->backward();
    uy->backward();
    const Arr<double> uxp = ux->v, uyp = uy->v;
    {  // solve_ux (622-633)
      Arr<S> rhs = ux->to_ortho();
      axpy(rhs, -dt, pres0->gradient(1, 0, scale));
      axpy(rhs, -dt, conv_of(*ux, uxp, uyp, false));
      ux->vhat = sol[0].solve(rhs);
    }
    {  // solve_uy (636-654)
      Arr<S> rhs = uy->to_ortho();
      axpy(rhs, -dt, pres0->gradient(0, 1, scale));
      axpy(rhs, dt, that);
      axpy(rhs, -dt, conv_of(*uy, uxp, uyp, false));
      uy->vhat = sol[1].solve(rhs);
    }
    Arr<S> div = ux->gradient(1, 0, scale);  // 698-703
    axpy(div, 1.0, uy->gradient(0, 1, scale));
    pres1->vhat = sol[3].solve(div);  // 710-715
    pres1->vhat.at(0, 0) = S(0);
    {  // project_velocity(1.0) (683-695)
      const Arr<S> dpdx = pres1->gradient(1, 0, scale), dpdy = pres1->gradient(0, 1, scale);
      const Arr<S> uxo = ux->vhat, uyo = uy->vhat;
      ux->from_ortho(dpdx);
      uy->from_ortho(dpdy);
#pragma omp parallel for schedule(static)
      for (size_t k = 0; k < uxo.d.size(); ++k) {
        ux->vhat.d[k] = ux->vhat.d[k] * -1.0 + uxo.d[k];
        uy->vhat.d[k] = uy->vhat.d[k] * -1.0 + uyo.d[k];
      }
    }
    axpy(pres0->vhat, -nu, div);  // update_pres (717-721)
    axpy(pres0->vhat, 1.0 / dt, pres1->to_ortho());
    {  // solve_temp (660-674)
      Arr<S> rhs = temp->to_ortho();
      axpy(rhs, dt * ka, fieldbc->gradient(2, 0, scale));
      axpy(rhs, dt * ka, fieldbc->gradient(0, 2, scale));
      axpy(rhs, -dt, conv_of(*temp, uxp, uyp, true));
      temp->vhat = sol[2].solve(rhs);
    }
    time += dt;
  }
  Field<S>* by_index(int w) {
    Field<S>* f[5] = {temp.get(), ux.get(), uy.get(), pres0.get(), pres1.get()};
    return f[w];
  }
};

// ---------------------------------------------------------------------------------------------------
// C interface (ctypes: oracle/cpu_twin/__init__.py)
// ---------------------------------------------------------------------------------------------------
struct Handle {
  bool periodic;
  Navier<double>* r = nullptr;
  Navier<cd>* c = nullptr;
};
extern "C" {
int tw_set_threads(int n) {
  if (n > 0) omp_set_num_threads(n);
  const int t = omp_get_max_threads();
  g_work.resize(t);
  return t;
}
int tw_set_blas(const char* path) {  // Fortran dgemm of an OpenBLAS / LAPACK shared library (single-threaded by the caller's env)
  void* h = dlopen(path, RTLD_NOW | RTLD_GLOBAL);
  if (!h) return 1;
  const char* names[] = {"scipy_dgemm_64_", "dgemm_64_", "scipy_dgemm_", "dgemm_"};
  for (int i = 0; i < 4; ++i)
    if (void* f = dlsym(h, names[i])) {
      g_dgemm = f;
      g_dgemm64 = i < 2;
      return 0;
    }
  return 2;
}
void* tw_create(int nx, int ny, double ra, double pr, double dt, double aspect, int adiabatic, int periodic, const double* lam,
                const double* q, const double* p) {
  if (g_work.empty()) tw_set_threads(0);
  try {
    Handle* h = new Handle;
    h->periodic = periodic != 0;
    if (h->periodic)
      h->c = new Navier<cd>(nx, ny, ra, pr, dt, aspect, adiabatic != 0, true, lam, q, p);
    else
      h->r = new Navier<double>(nx, ny, ra, pr, dt, aspect, adiabatic != 0, false, lam, q, p);
    return h;
  } catch (const std::exception& e) {
    fprintf(stderr, "cpu twin: %s\n", e.what());
    return nullptr;
  }
}
void tw_destroy(void* hv) {
  Handle* h = (Handle*)hv;
  delete h->r;
  delete h->c;
  delete h;
}
void tw_set_ics(void* hv, double amp_v, double amp_t, double m, double n) {  // set_velocity + set_temperature (927-936)
  Handle* h = (Handle*)hv;
  if (h->periodic) {
    h->c->apply_ic(*h->c->ux, amp_v, m, n, true), h->c->apply_ic(*h->c->uy, -amp_v, m, n, false), h->c->apply_ic(*h->c->temp, -amp_t, m, n, false);
  } else {
    h->r->apply_ic(*h->r->ux, amp_v, m, n, true), h->r->apply_ic(*h->r->uy, -amp_v, m, n, false), h->r->apply_ic(*h->r->temp, -amp_t, m, n, false);
  }
}
void tw_update(void* hv, int nsteps) {
  Handle* h = (Handle*)hv;
  for (int i = 0; i < nsteps; ++i) {
    if (h->periodic)
      h->c->update();
    else
      h->r->update();
  }
}
void tw_axis_times(double* t) { t[0] = g_t_axis[0], t[1] = g_t_axis[1]; }
double tw_time(void* hv) {
  Handle* h = (Handle*)hv;
  return h->periodic ? h->c->time : h->r->time;
}
// vhat of field `which` (0 temp, 1 ux, 2 uy, 3 pres, 4 pseudo pressure) as doubles (complex interleaved); returns its length
long long tw_get_vhat(void* hv, int which, double* out, long long cap) {
  Handle* h = (Handle*)hv;
  if (h->periodic) {
    const auto& a = h->c->by_index(which)->vhat;
    const long long len = 2LL * a.d.size();
    if (out && cap >= len) memcpy(out, a.d.data(), len * sizeof(double));
    return len;
  }
  const auto& a = h->r->by_index(which)->vhat;
  const long long len = (long long)a.d.size();
  if (out && cap >= len) memcpy(out, a.d.data(), len * sizeof(double));
  return len;
}
}
