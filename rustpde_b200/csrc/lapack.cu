// lapack.cu -- set-up time access to LAPACK (dgeev, dgetrf/dgetri, dgemm),
// the same routines the reference reaches through ndarray-linalg / OpenBLAS
// (src/solver/utils.rs:66-106, fdma_tensor.rs:117-128).  The library is
// dlopen'ed: $RUSTPDE_B200_LAPACK, rp_set_lapack_library(), or the usual
// sonames.  Only used when a solver is constructed (never in the time loop).
#include <dlfcn.h>

#include <algorithm>
#include <cstdint>
#include <mutex>
#include <numeric>

#include "model.h"

namespace rp {

namespace {
std::mutex g_mu;
std::string g_path;
void* g_handle = nullptr;
bool g_ilp64 = false;
void *g_dgeev = nullptr, *g_dgetrf = nullptr, *g_dgetri = nullptr, *g_dgemm = nullptr;
std::string g_err;

void* find_sym(void* h, const char* base, bool* ilp64) {
  const char* pre[2] = {"", "scipy_"};
  const char* suf[3] = {"_", "_64_", "64_"};
  for (auto p : pre)
    for (int s = 0; s < 3; ++s) {
      std::string name = std::string(p) + base + suf[s];
      void* f = dlsym(h, name.c_str());
      if (f) {
        if (ilp64) *ilp64 = (s != 0);
        return f;
      }
    }
  return nullptr;
}

bool try_open(const std::string& path) {
  void* h = dlopen(path.c_str(), RTLD_NOW | RTLD_LOCAL);
  if (!h) {
    g_err = std::string("dlopen(") + path + "): " + dlerror();
    return false;
  }
  bool i1 = false, i2 = false, i3 = false, i4 = false;
  void* a = find_sym(h, "dgeev", &i1);
  void* b = find_sym(h, "dgetrf", &i2);
  void* c = find_sym(h, "dgetri", &i3);
  void* d = find_sym(h, "dgemm", &i4);
  if (!a || !b || !c || !d) {
    g_err = path + ": dgeev/dgetrf/dgetri/dgemm not all found";
    dlclose(h);
    return false;
  }
  g_handle = h;
  g_ilp64 = i1;
  g_dgeev = a, g_dgetrf = b, g_dgetri = c, g_dgemm = d;
  return true;
}

bool ensure_loaded() {
  if (g_handle) return true;
  std::vector<std::string> cands;
  if (!g_path.empty()) cands.push_back(g_path);
  if (const char* e = getenv("RUSTPDE_B200_LAPACK")) cands.push_back(e);
  for (const char* n : {"libopenblas.so.0", "libopenblas.so", "liblapack.so.3", "liblapack.so", "libmkl_rt.so"}) cands.push_back(n);
  for (auto& c : cands)
    if (try_open(c)) return true;
  return false;
}

template <class I>
void do_eig_setup(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                  std::vector<double>& Q, std::vector<double>& P) {
  typedef void (*dgetrf_t)(I*, I*, double*, I*, I*, I*);
  typedef void (*dgetri_t)(I*, double*, I*, I*, double*, I*, I*);
  typedef void (*dgeev_t)(const char*, const char*, I*, double*, I*, double*, double*, double*, I*, double*, I*, double*, I*, I*);
  typedef void (*dgemm_t)(const char*, const char*, I*, I*, I*, double*, const double*, I*, const double*, I*, double*, double*, I*);
  auto f_getrf = (dgetrf_t)g_dgetrf;
  auto f_getri = (dgetri_t)g_dgetri;
  auto f_geev = (dgeev_t)g_dgeev;
  auto f_gemm = (dgemm_t)g_dgemm;
  I n = m, info = 0;
  // row-major product C = A * B through column-major dgemm: C^T = B^T A^T
  auto matmul = [&](const std::vector<double>& A, const std::vector<double>& B, std::vector<double>& C) {
    C.assign((size_t)m * m, 0.0);
    double one = 1.0, zero = 0.0;
    f_gemm("N", "N", &n, &n, &n, &one, B.data(), &n, A.data(), &n, &zero, C.data(), &n);
  };
  // inverse: inv(X^T) = inv(X)^T, so LAPACK on the row-major bytes gives the row-major inverse
  auto inverse = [&](std::vector<double> X) {
    std::vector<I> ipiv(m);
    f_getrf(&n, &n, X.data(), &n, ipiv.data(), &info);
    if (info != 0) throw Error(RP_ERR_LAPACK, "dgetrf failed");
    I lwork = (I)m * 64;
    std::vector<double> work((size_t)lwork);
    f_getri(&n, X.data(), &n, ipiv.data(), work.data(), &lwork, &info);
    if (info != 0) throw Error(RP_ERR_LAPACK, "dgetri failed");
    return X;
  };
  std::vector<double> Cinv = inverse(Cx);
  std::vector<double> X;
  matmul(Cinv, Ax, X);  // xmat = inv(Cx) . Ax      (fdma_tensor.rs:122)
  // dgeev wants column-major: hand it X^T-as-bytes == column-major X
  std::vector<double> Xc((size_t)m * m);
  for (int i = 0; i < m; ++i)
    for (int j = 0; j < m; ++j) Xc[(size_t)j * m + i] = X[(size_t)i * m + j];
  std::vector<double> wr(m), wi(m), vr((size_t)m * m);
  double vl_dummy = 0.0;
  I one_i = 1, lwork = -1;
  double wq = 0.0;
  f_geev("N", "V", &n, Xc.data(), &n, wr.data(), wi.data(), &vl_dummy, &one_i, vr.data(), &n, &wq, &lwork, &info);
  lwork = (I)wq + 1;
  std::vector<double> work((size_t)lwork);
  f_geev("N", "V", &n, Xc.data(), &n, wr.data(), wi.data(), &vl_dummy, &one_i, vr.data(), &n, work.data(), &lwork, &info);
  if (info != 0) throw Error(RP_ERR_LAPACK, "dgeev failed");
  // utils.rs:80-94: keep real parts (a complex pair shares the real part vector), sort descending
  std::vector<int> src(m);
  for (int j = 0; j < m; ++j) src[j] = j;
  for (int j = 0; j + 1 < m; ++j)
    if (wi[j] > 0.0) src[j + 1] = j, ++j;
  std::vector<int> perm(m);
  std::iota(perm.begin(), perm.end(), 0);
  std::stable_sort(perm.begin(), perm.end(), [&](int a, int b) { return wr[a] < wr[b]; });
  std::reverse(perm.begin(), perm.end());
  lam.resize(m);
  Q.assign((size_t)m * m, 0.0);
  for (int j = 0; j < m; ++j) {
    lam[j] = wr[perm[j]];
    const double* col = &vr[(size_t)src[perm[j]] * m];
    for (int i = 0; i < m; ++i) Q[(size_t)i * m + j] = col[i];
  }
  std::vector<double> Qi = inverse(Q);
  matmul(Qi, Cinv, P);  // fwd = Q^-1 . inv(Cx)    (fdma_tensor.rs:125)
}
}  // namespace

void lapack_set_library(const char* path) {
  std::lock_guard<std::mutex> g(g_mu);
  g_path = path ? path : "";
  if (g_handle) {
    dlclose(g_handle);
    g_handle = nullptr;
  }
}

bool lapack_available(std::string* why) {
  std::lock_guard<std::mutex> g(g_mu);
  bool ok = ensure_loaded();
  if (!ok && why) *why = g_err;
  return ok;
}

static void eig_setup_locked(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                             std::vector<double>& Q, std::vector<double>& P) {
  if (!ensure_loaded())
    throw Error(RP_ERR_LAPACK,
                "no LAPACK library found for the fast-diagonalisation set-up (set RUSTPDE_B200_LAPACK or call "
                "rp_set_lapack_library): " + g_err);
  if (g_ilp64)
    do_eig_setup<int64_t>(m, Cx, Ax, lam, Q, P);
  else
    do_eig_setup<int32_t>(m, Cx, Ax, lam, Q, P);
}

void lapack_eig_setup(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                      std::vector<double>& Q, std::vector<double>& P) {
  std::lock_guard<std::mutex> g(g_mu);
  eig_setup_locked(m, Cx, Ax, lam, Q, P);
}

// The Chebyshev operators only couple modes of equal parity (offsets -2, 0, +2, +4),
// so inv(Cx).Ax is exactly block diagonal after an even/odd permutation.  This
// variant diagonalises the two half-size blocks separately and merges the
// result (eigenvalues sorted descending like utils.rs:80-94): the same
// decomposition as lapack_eig_setup in exact arithmetic, with Q and P exactly
// checkerboard (cross-parity entries are 0.0), which is what lets the solve run
// two half-size GEMM pairs.
void lapack_eig_setup_parity(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                             std::vector<double>& Q, std::vector<double>& P) {
  std::lock_guard<std::mutex> g(g_mu);
  std::vector<double> l2[2], q2[2], p2[2];
  int mp[2];
  for (int par = 0; par < 2; ++par) {
    const int mm = (m - par + 1) / 2;
    mp[par] = mm;
    std::vector<double> c((size_t)mm * mm), a((size_t)mm * mm);
    for (int i = 0; i < mm; ++i)
      for (int j = 0; j < mm; ++j) {
        c[(size_t)i * mm + j] = Cx[(size_t)(2 * i + par) * m + 2 * j + par];
        a[(size_t)i * mm + j] = Ax[(size_t)(2 * i + par) * m + 2 * j + par];
      }
    eig_setup_locked(mm, c, a, l2[par], q2[par], p2[par]);
  }
  // merge, descending (both halves are already sorted descending)
  lam.assign(m, 0.0);
  Q.assign((size_t)m * m, 0.0);
  P.assign((size_t)m * m, 0.0);
  int i0 = 0, i1 = 0;
  for (int k = 0; k < m; ++k) {
    int par;
    if (i0 >= mp[0]) par = 1;
    else if (i1 >= mp[1]) par = 0;
    else par = (l2[0][i0] >= l2[1][i1]) ? 0 : 1;
    const int kk = par ? i1++ : i0++;
    const int mm = mp[par];
    lam[k] = l2[par][kk];
    for (int a = 0; a < mm; ++a) {
      Q[(size_t)(2 * a + par) * m + k] = q2[par][(size_t)a * mm + kk];
      P[(size_t)k * m + 2 * a + par] = p2[par][(size_t)kk * mm + a];
    }
  }
}

}  // namespace rp
