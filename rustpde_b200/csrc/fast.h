// fast.h -- host interface of the specialised Navier2D kernels (fast_y.cu, fast_x.cu).
//
// "y kernels" own 4 adjacent rows of a row-major [n0, n1] array (lanes are
// contiguous in memory), "x kernels" own 4 adjacent columns (lanes are strided).
// Supported sizes: y lanes of n1 = 2^k + 1 points (DCT-I through a power-of-two
// FFT), x lanes of any n0 whose Bluestein length 2^k >= 2 (n0 - 1) - 1 is
// instantiated.  Everything else stays on the generic lane programs.
#pragma once
#include <vector>

#include "lane_prog.h"

namespace rp {
namespace fk {

struct Mat {  // pitched real matrix in global memory
  double* p;
  long long ld;
  int rows, cols;
};

struct DctTab {
  int n;                 // lane length, N = n - 1
  const double2* sc;     // (sin, cos)(pi j / N), j = 0..N/2
  const double2* tw;     // per-span compact twiddles tw[S/2 - 1 + q] = exp(-2 pi i q / S), S <= N (pow2) or Lb (Bluestein)
  const double2* chirp;  // Bluestein: exp(-i pi j^2 / N), j < N
  const double2* bhat;   // Bluestein: FFT_Lb(wrapped conj chirp) / Lb, digit-reversed order of fft_dif, transposed [8][Lb/8] (fast.cuh dif_dit_mid)
};
struct FdmaTabs {  // pre-swept banded solve (tables.h FdmaDev)
  const double *fp, *bs, *bp1, *bp2;
};
struct B2Tabs {  // B2 preconditioner rows (matvec.rs:172-193)
  const double *lo, *di, *up;
};
struct TdmaTabs {  // from_ortho: S^T then the pre-factored (S^T S) solve
  const double *sd, *sl, *fs, *fp, *bp;
  const double *pf, *pb;  // chunk-major packed copies (perm_table): forward {sd, sl, fs, fp} (W = 4), backward {bp} (W = 1)
};
struct ModeTabs {  // per-lane A + (lam + alpha) C (fdma_tensor.rs:219-227)
  const double *a_low, *a_up1, *a_up2, *c_low, *c_up1, *c_up2;
  const double* lam;
  double alpha;
  const double* inv;  // swept pivot reciprocals, [lanes][inv_ld]
  long long inv_ld;
  // chunk-major packed copies (perm_table): forward sweep {b2 lo, di, up, a_low[i-2], c_low[i-2], -} (W = 6),
  // backward sweep {a_up1, c_up1, a_up2, c_up2, a_up2[i-2], c_up2[i-2], a_low[i-2], c_low[i-2]} (W = 8)
  const double *pf, *pb;
  // plain per-column copies for the warp-serial sweeps (fast_pw.cu, pack_rows): rf [m][6] = {b2 lo, di, up, a_low[i-2],
  // c_low[i-2], 0}, rb [m][8] = {a_up1, c_up1, a_up2, c_up2, a_up2[i-2], c_up2[i-2], a_low[i-2], c_low[i-2]}
  const double *rf, *rb;
};

bool y_supported(int n1);
bool x_supported(int n0);
void apply_kflags();  // development switches (RUSTPDE_B200_KFLAGS) -> device constants; call outside stream capture

// true the first time it is called for `mask` on the current CUDA device (kernel attributes such as the dynamic
// shared-memory limit are per device, not per process)
inline bool first_use_on_device(unsigned long long& mask) {
#ifndef RP_EMU
  int dev = 0;
  cudaGetDevice(&dev);
  const unsigned long long bit = 1ull << (dev & 63);
  if (mask & bit) return false;
  mask |= bit;
  return true;
#else
  if (mask) return false;
  mask = 1;
  return true;
#endif
}

// ---- chunk-major coefficient tables of the chunked scans (fast.cuh) ------------------------------
// A scan over m elements is cut into NG chunks per parity chain of cl = ceil(ceil(m / 2) / NG) elements; element
// (group g, step u, parity p) is natural index i = 2 t + p (forward) or 2 (M_p - 1 - t) + p (backward), t = g cl + u,
// M_p = (m - p + 1) / 2.  perm_table() packs W coefficients per element at slot (u NG + g) 2 + p, so that the
// slots a warp reads in one step of its chunks are contiguous in memory:
//   out[slot * W + k] = src[k][i + shift[k]]   (0 where i + shift[k] is outside [0, len[k]) or t >= M_p)
// the kernels fetch a whole compile-time chunk bound `rows` of slots per group unconditionally, so the tables are
// sized by it (rows >= cl)
struct ScanShape {
  int ng, rows;
};
ScanShape y_scan_shape(int n1);   // scalar scans of the y kernels for lanes of n1 points (ng = 0: unsupported)
ScanShape x_scan_shape(int n0);   // two-lane first-order scans of the x kernels for lanes of n0 points
ScanShape x_scan2_shape(int n0);  // two-lane second-order scan of xk_forward
std::vector<double> perm_table(int m, bool fwd, ScanShape sh, int W, const std::vector<std::vector<double>>& src,
                               const std::vector<int>& shift);

// ---- y kernels ------------------------------------------------------------------
struct YBackwardArgs {  // B_y S_y (value), B_y D_y S_y / sy, and B_y S_y of a second array
  Mat a, adx;           // [rows, my]
  Mat val, dy, dx;      // [rows, ny]; val.p may be null
  const double *sd, *sl;
  double isy;
  DctTab t;
};
struct YBackwardArgs3 {
  YBackwardArgs a[3];
};
void launch_y_backward(const YBackwardArgs3& a, int nbatch, cudaStream_t s);  // nbatch independent fields, one launch

struct YConvArgs {  // out = cut_y(F_y(u * (du + bcx) + v * (dv + bcy)))   (conv_term.rs:41)
  Mat u, du, v, dv, bcx, bcy;  // bcx/bcy.p may be null
  // solid-mask volume penalisation (navier.rs:552-560, 580-584, 604-608): + ieta * mask * (w + wbc - sval); mask.p may be null
  Mat mask, sval, w, wbc;      // w: the field's own physical value; wbc: physical boundary field (temperature) or null
  double ieta;
  Mat out;
  int cut;  // first zeroed y mode (navier.rs:1029), >= ny: none
  DctTab t;
};
struct YConvArgs3 {
  YConvArgs a[3];
};
void launch_y_conv(const YConvArgs3& a, int nbatch, cudaStream_t s);

struct YAdiArgs {  // y half of HholtzAdi (hholtz_adi.rs:113,129) + pieces of the divergence
  Mat w;           // [mx, ny]
  Mat out;         // [mx, my]
  Mat aux;         // mode 1: S_y out  [mx, ny];  mode 2: D_y S_y out / sy
  int mode;
  const double *sd, *sl;
  double isy;
  B2Tabs b2;
  FdmaTabs f;
  const double *pt1, *pt2;  // chunk-major packed tables: {b2 lo, di, up, f.fp} (forward), {f.bs, bp1, bp2, 0} (backward)
  int ny;
};
struct YAdiArgs3 {
  YAdiArgs a[3];
};
void launch_y_adi(const YAdiArgs3& a, int nbatch, cudaStream_t s);

struct YModeArgs {  // per-mode banded solve of the fast diagonalisation (fdma_tensor.rs:219-227)
  Mat g;            // [mx, ny]
  Mat h;            // [mx, my]
  B2Tabs b2;
  ModeTabs m;
  int ny;
};
void launch_y_mode(const YModeArgs& a, cudaStream_t s);

struct YProjectArgs {  // u -= from_ortho(grad phi), y part (navier.rs:683-695)
  Mat a1, a2;          // [mx, my]: from_ortho_x(D_x S_x phi)/sx, from_ortho_x(S_x phi)
  Mat ux, uy;          // [mx, my], updated in place
  const double *nsd, *nsl;  // Neumann stencil of phi along y
  TdmaTabs t;          // Dirichlet from_ortho along y
  double isy;
  int ny;
};
void launch_y_project(const YProjectArgs& a, cudaStream_t s);

struct YPresArgs {  // p += -nu div + to_ortho(phi)/dt (navier.rs:717-721); dyp = D_y p / sy
  Mat phi;          // [mx, my] Neumann composite
  Mat div, pres, dyp;  // [nx, ny]
  const double *xsd, *xsl;  // Neumann stencil along x (across lanes)
  const double *ysd, *ysl;  // Neumann stencil along y
  double inv_dt, nu, isy;
  int ny;
  int only_dyp;  // 1: leave pres as it is and only refresh dyp = D_y pres / sy (after pres was rewritten from outside)
};
void launch_y_pres(const YPresArgs& a, cudaStream_t s);

struct YDivPrepArgs {  // y parts of the divergence of (ux, uy) (navier.rs:698-703): vx = S_y ux, ey = D_y S_y uy / sy
  Mat ux, uy;          // [mx, my]
  Mat vx, ey;          // [mx, ny]
  const double *sd, *sl;
  double isy;
  int ny;
};
void launch_y_divprep(const YDivPrepArgs& a, cudaStream_t s);

// ---- x kernels ------------------------------------------------------------------
struct XBackwardArgs {  // B_x S_x u^ and B_x D_x S_x u^ / sx
  Mat src;              // [mx, cols]
  Mat val, dx;          // [nx, cols]
  const double *sd, *sl;
  double isx;
  DctTab t;
};
struct XBackwardArgs3 {
  XBackwardArgs a[3];
  int next_wave;  // set by the launcher: linear block distance to prefetch ahead
};
void launch_x_backward(const XBackwardArgs3& a, int nbatch, cudaStream_t s);

struct XForwardArgs {  // forward DCT-x + dealias + rhs assembly + x half of HholtzAdi
  Mat conv;            // [nx, ny] after the y-forward transform
  Mat out;             // [mx, ny]
  int cut;             // first zeroed x mode
  double dt;
  // + S_x S_y u^_f
  Mat fld;             // [mx, my]
  const double *fxsd, *fxsl, *fysd, *fysl;
  // mode 0: - dt/sx D_x pres ; 1: - dt dyp + dt (S_x S_y T^ + tbc) ; 2: + bcdiff
  int mode;
  Mat pres, dyp, tmp, tbc, bcdiff;
  int tbc_rows, bcdiff_rows;  // rows >= these of tbc / bcdiff hold exact zeros and are not read
  const double *txsd, *txsl, *tysd, *tysl;
  double isx;
  B2Tabs b2;
  FdmaTabs f;
  const double *pt1, *pt2;  // chunk-major packed tables: {b2 lo, di, up, f.fp} (forward), {f.bs, bp1, bp2, 0} (backward)
  DctTab t;
  Mat rhs;  // p != null: stop after the rhs assembly and store the [nx, ny] ortho rhs here (the sweeps run in xw_adi)
};
struct XForwardArgs3 {
  XForwardArgs a[3];
  int next_wave;
};
void launch_x_forward(const XForwardArgs3& a, int nbatch, cudaStream_t s);

struct XDivArgs {  // div = D_x S_x vx / sx + S_x ey ; r1 = B2_x div
  Mat vx, ey;      // [mx, ny]
  Mat div;         // [nx, ny]
  Mat r1;          // [mx, ny]
  const double *sd, *sl;
  double isx;
  B2Tabs b2;
  int nx;
};
void launch_x_div(const XDivArgs& a, cudaStream_t s);
void launch_xs_div(const XDivArgs& a, cudaStream_t s);  // the same pass as a streaming column scan (fast_xs.cu)

struct XProjectArgs {  // a1 = from_ortho_x(D_x S_x phi)/sx, a2 = from_ortho_x(S_x phi)
  Mat phi;             // [mx, my]
  Mat a1, a2;          // [mx, my]
  const double *nsd, *nsl;  // Neumann stencil along x
  TdmaTabs t;               // Dirichlet from_ortho along x
  double isx;
  int nx;
};
void launch_x_project(const XProjectArgs& a, cudaStream_t s);

struct XFdctArgs {  // out = scale * cut_x(F_x conv): forward DCT-x + dealias (navier.rs:1028) + the -dt of navier.rs:630
  Mat conv;          // [nx, ny] after the y-forward transform
  Mat out;           // [nx, ny]
  int cut;           // first zeroed x mode
  double scale;
  DctTab t;
};
struct XFdctArgs3 {
  XFdctArgs a[3];
  int next_wave;
};
void launch_x_fdct(const XFdctArgs3& a, int nbatch, cudaStream_t s);

// ---- streaming column scans along x (fast_xs.cu): no tile, a block owns 8 adjacent columns ----
bool xs_supported(int n0);
ScanShape xs_scan_shape();  // scan shape of every chunk-major table these kernels read
struct XsRhsAdiArgs {  // rhs assembly (navier.rs:622-674) + x half of HholtzAdi (hholtz_adi.rs:108,128)
  Mat chat;            // [nx, ny]: -dt * dealiased convection term (XFdctArgs.out)
  Mat rhs;             // [nx, ny] scratch (may alias chat)
  Mat out;             // [mx, ny]
  Mat fld;             // [mx, my] old composite coefficients: + S_x S_y fld
  const double *fxsd, *fxsl, *fysd, *fysl;
  int mode;            // 0: + dxp ; 1: - dt dyp + dt (S_x S_y tmp + tbc) ; 2: + bcdiff
  Mat dxp, dyp, tmp, tbc, bcdiff;
  const double *txsd, *txsl, *tysd, *tysl;
  double dt;
  const double *pt1, *pt2;  // chunk-major packed tables (xs_scan_shape): {b2 lo, di, up, f.fp}, {f.bs, bp1, bp2, 0}
  int nx;
};
struct XsRhsAdiArgs3 {
  XsRhsAdiArgs a[3];
};
void launch_xs_rhs_adi(const XsRhsAdiArgs3& a, int nbatch, cudaStream_t s);
struct XsProjectArgs {  // a1 = from_ortho_x(D_x S_x phi)/sx, a2 = from_ortho_x(S_x phi)
  Mat phi;              // [mx, my]
  Mat p, d;             // [nx, my] scratch
  Mat a1, a2;           // [mx, my]
  const double *nsd, *nsl;
  TdmaTabs t;           // pf / pb in xs_scan_shape
  double isx;
  int nx;
};
void launch_xs_project(const XsProjectArgs& a, cudaStream_t s);
struct XsDiffArgs {  // dst = sc * D_x src along x
  Mat src, dst;      // [nx, cols]
  double sc;
  int nx;
};
void launch_xs_dxp(const XsDiffArgs& a, cudaStream_t s);

struct XAdiArgs {  // x half of a stand-alone HholtzAdi::solve: out = Fdma_x(B2_x in)
  Mat in;          // [nx, cols]
  Mat out;         // [mx, cols]
  const double *pt1, *pt2;  // chunk-major packed tables: {b2 lo, di, up, f.fp} (forward), {f.bs, bp1, bp2, 0} (backward)
  int nx;
};
void launch_x_adi(const XAdiArgs& a, cudaStream_t s);

// ---- warp-serial column sweeps along x (fast_xw.cu): a warp owns 16 adjacent columns, a thread one parity chain ----
bool xw_supported(int n0);
// row-major table of W doubles per row: out[i * W + k] = src[k][i + shift[k]] (0 outside the source)
std::vector<double> pack_rows(int rows, int W, const std::vector<std::vector<double>>& src, const std::vector<int>& shift);
struct XwAdiArgs {  // out = Fdma_x(B2_x in)   (hholtz_adi.rs:108,128)
  Mat in;           // [nx, cols]
  Mat tmp;          // [mx, cols] scratch (forward-sweep result)
  Mat out;          // [mx, cols]
  const double *cf, *cb;  // pack_rows tables [mx][4]: {b2 lo, di, up, f.fp}, {f.bs, bp1, bp2, 0}
  int nx;
};
struct XwAdiArgs3 {
  XwAdiArgs a[3];
};
void launch_xw_adi(const XwAdiArgs3& a, int nbatch, cudaStream_t s);
struct XwDivArgs {  // div = D_x S_x vx / sx + S_x ey ; r1 = B2_x div   (one sweep, last row to first)
  Mat vx, ey;       // [mx, ny]
  Mat div;          // [nx, ny]
  Mat r1;           // [mx, ny]
  const double* tab;  // xw_div_table [nx][8]
  int nx;
};
std::vector<double> xw_div_table(int n, double isx, const std::vector<double>& sd, const std::vector<double>& sl,
                                 const std::vector<double>& lo, const std::vector<double>& di, const std::vector<double>& up);
void launch_xw_div(const XwDivArgs& a, cudaStream_t s);
struct XwProjectArgs {  // a1 = from_ortho_x(D_x S_x phi)/sx, a2 = from_ortho_x(S_x phi)   (two sweeps; a1 / a2 hold the intermediate)
  Mat phi;              // [mx, my]
  Mat a1, a2;           // [mx, my]
  const double *t1, *t2;  // xw_project_tables [nx][8], [mx][2]
  int nx;
};
void xw_project_tables(int n, double isx, const std::vector<double>& nsd, const std::vector<double>& nsl, const std::vector<double>& sd,
                       const std::vector<double>& sl, std::vector<double>& t1, std::vector<double>& t2);
void launch_xw_project(const XwProjectArgs& a, cudaStream_t s);

// Destination of a fused transpose: part q (a peer GPU's buffer, mapped through CUDA IPC, or a local one) owns
// the global indices [beg[q], beg[q+1]) along the scattered axis.  nparts == 0: no scatter (dense output).
struct Scatter {
  double* ptr[8];
  int beg[9];
  int nparts;
};

// ---- periodic path (fast_p.cu): complex arrays are passed as Mat with ld / cols in complex units ----
bool px_supported(int n0);

struct PC2rArgs {  // c2r along x of a spectral array and of its (i k / sx) derivative
  Mat src;         // [mk, cols] complex
  Mat val, dx;     // [nx, cols] real
  double isx;
  const double2* tw;  // per-span compact twiddles of length n (see DctTab)
  int n;
};
struct PC2rArgs3 {
  PC2rArgs a[3];
};
void launch_p_c2r(const PC2rArgs3& a, int nbatch, cudaStream_t s);

struct PR2cArgs {  // r2c along x + dealias cut
  Mat src;         // [nx, cols] real; ignored when u.p is set
  Mat u, du, v, dv, bcx, bcy;  // optional: transform u * (du + bcx) + v * (dv + bcy) instead of src (conv_term.rs:41)
  Mat dst;         // [mk, cols] complex
  int cut;         // first zeroed kx
  const double2* tw;
  int n;
  // fused transpose (slab decomposition): row k goes to the peer that owns it, into its [rows_q, ny] array at
  // columns j0 + c (NVLink stores from inside the kernel instead of a separate all-to-all)
  Scatter sdst;
  int j0, ny;
};
struct PR2cArgs3 {
  PR2cArgs a[3];
};
void launch_p_r2c(const PR2cArgs3& a, int nbatch, cudaStream_t s);

struct PHholtzArgs {  // rhs assembly + per-mode Helmholtz solve on complex rows
  Mat chat;           // [mk, ny] convection term
  Mat fld, out;       // [mk, my] old / new composite coefficients (may alias)
  Mat pres, tmp, tbc, bcdiff;  // mode 0: pres; mode 1: pres, tmp [mk, my], tbc; mode 2: bcdiff
  const double *sd, *sl, *tsd, *tsl;
  int mode;
  double dt, isx, isy;
  B2Tabs b2;
  ModeTabs m;  // lam / inv already offset to the first row of the slab
  int ny;
  int k0;      // global kx of row 0 (slab decomposition over kx; 0 on one GPU)
  Mat dyp;     // [mk, ny] scratch of the warp-serial kernel (mode 1: - dt/sy d/dy pres)
  const double* rs;  // warp-serial kernel: pack_rows table [ny][4] = {sd_j, sl_{j-2}, tsd_j, tsl_{j-2}}
};
struct PHholtzArgs3 {
  PHholtzArgs a[3];
};
void launch_p_hholtz(const PHholtzArgs3& a, int nbatch, cudaStream_t s);

struct PDivPoisArgs {  // divergence + per-mode Poisson solve
  Mat ux, uy;          // [mk, my]
  Mat div;             // [mk, ny]
  Mat phi;             // [mk, my]
  const double *sd, *sl;
  double isx, isy;
  B2Tabs b2;
  ModeTabs m;
  int ny;
  int k0;
  const double* rs;  // warp-serial kernel: pack_rows table [ny][4] = {sd_j, sl_{j-2}, 2 j / sy, 0}
};
void launch_p_divpois(const PDivPoisArgs& a, cudaStream_t s);
// the same two passes as warp-serial row sweeps (fast_pw.cu); launch_p_hholtz / launch_p_divpois pick them by row count
// (RUSTPDE_B200_PW=0 / 1 forces the tile kernels / the row sweeps)
bool pw_enabled(int rows, bool divpois);
void launch_pw_hholtz(const PHholtzArgs3& a, int nbatch, cudaStream_t s);
void launch_pw_divpois(const PDivPoisArgs& a, cudaStream_t s);

struct PProjectArgs {  // projection + pressure update
  Mat phi, ux, uy;     // [mk, my]
  Mat div, pres;       // [mk, ny]
  const double *nsd, *nsl;
  TdmaTabs t;
  double isx, isy, nu, inv_dt;
  int ny;
  int k0;
  // row-sweep form (fast_pw.cu pw_project), optional: pw_project_tables of the y bases and two [mk, my] complex scratch arrays
  const double *w1 = nullptr, *w2 = nullptr;
  Mat z1{nullptr, 0, 0, 0}, z2{nullptr, 0, 0, 0};
};
void launch_p_project(const PProjectArgs& a, cudaStream_t s);  // picks pw_project by row count when the tables are there
void launch_pw_project(const PProjectArgs& a, cudaStream_t s);
void pw_project_tables(int n, double isy, const std::vector<double>& nsd, const std::vector<double>& nsl, const std::vector<double>& sd,
                       const std::vector<double>& sl, const std::vector<double>& fs, const std::vector<double>& fp,
                       const std::vector<double>& bp, std::vector<double>& w1, std::vector<double>& w2);
bool pw_project_enabled(int rows);

// slab decomposition over kx: the y transforms run on complex rows of the kx slab, before (backward) and
// after (forward) the x transform on y slabs -- the operators commute
struct PYBackArgs {  // B_y S_y and B_y D_y S_y / sy of complex rows
  Mat src;           // [rows, my] complex composite coefficients
  Mat val, dy;       // [rows, ny] complex
  const double *sd, *sl;
  double isy;
  DctTab t;
  // fused transpose: column j goes to the peer that owns it, into its [mk, ny_q] array at row k0 + r
  Scatter sval, sdy;
  int k0;
};
struct PYBackArgs3 {
  PYBackArgs a[3];
};
void launch_p_ybackward(const PYBackArgs3& a, int nbatch, cudaStream_t s);

struct PYFwdArgs {  // forward DCT-y of complex rows + dealias cut in y
  Mat src, dst;     // [rows, ny] complex
  int cut;
  DctTab t;
};
struct PYFwdArgs3 {
  PYFwdArgs a[3];
};
void launch_p_yforward(const PYFwdArgs3& a, int nbatch, cudaStream_t s);

}  // namespace fk
}  // namespace rp
