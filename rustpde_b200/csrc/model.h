// model.h -- host-side mirror of the reference's objects on the Navier2D path
// (Space2 / Field2: src/field.rs, funspace/src/space2.rs; HholtzAdi / Hholtz /
// Poisson: src/solver/*.rs; Navier2D: src/navier/navier.rs), each one a thin
// owner of device arrays plus the lane programs / GEMMs that implement it.
#pragma once
#include <array>
#include <functional>
#include <map>
#include <memory>
#include <string>
#include <vector>

#include "fast.h"
#include "kernels.h"
#include "tables.h"

namespace rp {

// pitched device 2-D array, row-major; complex arrays hold interleaved (re, im)
struct Arr {
  DevBuf buf;
  int rows = 0, cols = 0;
  long long ld = 0;  // in elements of its own type (double or double2)
  bool cplx = false;
  void alloc(int r, int c, bool cx);
  double* d() const { return buf.as<double>(); }
  size_t elem_bytes() const { return cplx ? 16 : 8; }
  void upload(const double* host, cudaStream_t s);        // dense row-major host array
  void download(double* host, cudaStream_t s) const;
  void zero(cudaStream_t s);
};

// non-owning view of a pitched device array (what lane-program loads/stores take)
struct ArrRef {
  const void* p;
  long long ld;
  int rows, cols;
  bool cplx;
  ArrRef(const void* p_, long long ld_, int r, int c, bool cx) : p(p_), ld(ld_), rows(r), cols(c), cplx(cx) {}
  ArrRef(const Arr& a) : p(a.buf.p), ld(a.ld), rows(a.rows), cols(a.cols), cplx(a.cplx) {}
};

// A finalised lane program on the device
struct Built {
  DevBuf dprog, dprof;
  int nblocks = 0, nthreads = 0, smem = 0;
  void dump_prof(const char* name) const;  // development aid (RUSTPDE_B200_OPPROF=1)
  double bytes = 0.0;  // algorithmic global-memory bytes of one launch (loads + stores)
  bool valid = false;
  void launch(cudaStream_t s) const;
};

class ProgBuilder {
 public:
  ProgBuilder(int axis, int nunits);
  void ld(int r, ArrRef a, int n, Lay lay, double s = 1.0, int flags = 0, int shift = 0,
          const double* lanecoef = nullptr, int zfill = 0, int lane2 = -1);
  void st(int r, ArrRef a, int n, Lay lay, double s = 1.0, int flags = 0, int cut_i = -1, int cut_lane = -1,
          int lane2 = -1);
  void toortho(int r, const Base& b, Lay lay);
  void fromortho(int r, const Base& b, Lay lay);
  Lay diff(int r, int n, Lay lay, int times, double scale);
  Lay dct(int r, const Base& b, Lay lay, bool backward);
  void bandmv(int r, const Base& b, Lay lay);
  void fdma(int r, const FdmaDev& f, Lay lay);
  void fdmamode(int r, int rinv, const FdmaModeDev& f, Lay lay, bool complex_lanes);
  void copy(int rd, int rs, int n, Lay ld, Lay ls);
  void axpy(int rd, int rs, int n, double a, Lay ld, Lay ls);
  void scale(int r, int n, double a, Lay lay);
  void mulpw(int rd, int ra, int rb, int n, Lay ld, Lay lab, bool acc);
  void cut(int r, int n, int from, Lay lay);
  void mulik(int r, int n, double a, Lay lay, bool elem_k = false);
  void zero(int r, int n, Lay lay);
  void setzero00(int r, Lay lay, bool complex_lanes);
  void rfft_st(int r0, ArrRef dst, const Base& b, double s = 1.0, int cut_k = -1);   // r2c + store
  void irfft_ld(int r0, ArrRef src, const Base& b, double s = 1.0, bool mulik = false);  // load + c2r
  Built build();
  int default_anchor = 0;  // set by callers: SPLIT anchor large enough for the program

 private:
  Instr& add(int op);
  void touch(int r, Lay lay, int n);
  Program p_;
  int fftlen_ = 0, wbcap_ = 0;
  double bytes_unit_ = 0.0;  // bytes moved per slot unit
};

struct Space2 {
  std::shared_ptr<Base> b0, b1;
};

// FieldBase<f64, f64, T2, S, 2>  (src/field.rs:66-129)
class Field2 {
 public:
  explicit Field2(const Space2& sp);
  Space2 sp;
  int n0, n1, m0, m1, o0, o1;  // physical, spectral, ortho-spectral shapes
  bool cplx;
  Arr v, vhat, ortho;  // ortho: staging for to_ortho / from_ortho / gradient results
  std::vector<double> x[2], dx[2];
  cudaStream_t stream = 0;
  unsigned long long vhat_version = 0;  // bumped whenever vhat is rewritten from outside a Navier2D step

  void forward();
  void backward();
  void to_ortho();    // vhat -> ortho
  void from_ortho();  // ortho -> vhat
  void gradient(int dx, int dy, const double* scale);  // vhat -> ortho
  // weighted averages (src/field/average.rs:25-57)
  double average();
  void average_axis0(std::vector<double>& out);
  const double* weights_x() const { return wx_.as<double>(); }
  const double* weights_y() const { return wy_.as<double>(); }

 private:
  Arr ta_, tb_;
  Built fwd_y_, fwd_x_, bwd_x_, bwd_y_, to_x_, to_y_, from_x_, from_y_;
  std::map<std::array<long long, 4>, std::pair<Built, Built>> grad_;
  DevBuf wx_, wy_, red_;
  void build_transforms();
};

// HholtzAdi (src/solver/hholtz_adi.rs:32-130)
struct AxisAdi {
  FdmaDev fdma;
};
// Hholtz / Poisson through FdmaTensor (src/solver/{hholtz,poisson,fdma_tensor}.rs)
struct TensorSolver {
  FdmaModeDev mode;     // per-lane banded solve along y
  Arr P, Q;             // fwd = Q^-1 Cx^-1, bwd = Q  (only when x is Chebyshev; full matrices, strict mode)
  std::vector<double> lam;
  bool x_diag = false;  // Fourier axis: no GEMM
  // parity-split mode (Q, P exactly checkerboard): modes are stored grouped by the
  // parity of their eigenvector (me even ones first), the contractions run on the
  // even and odd sub-blocks only (half the flops)
  bool split = false;
  int me = 0, mo = 0;
  Arr Pe, Po, Qe, Qo;
};

enum SolverKind { SOLVER_HHOLTZ = 0, SOLVER_HHOLTZ_ADI = 1, SOLVER_POISSON = 2 };

struct EigData {  // optional externally supplied set-up data (lam, Q, P = Q^-1 Cx^-1), row-major m x m
  const double* lam = nullptr;
  const double* q = nullptr;
  const double* p = nullptr;
};

class Solver2 {
 public:
  Solver2(int kind, const Space2& sp, double cx, double cy, double alpha, const EigData* eig);
  int kind;
  Space2 sp;
  int n0, n1, m0, m1;  // rhs is (n0 x n1) ortho, solution (m0 x m1) composite
  bool x_fourier;
  AxisAdi adi[2];
  TensorSolver ts;
  cudaStream_t stream = 0;
  // generic entry: in_/out_ are solver-owned staging arrays
  Arr in_r, out_r, in_c, out_c;
  void solve(bool complex_data);
  void solve_dev(const Arr& in, Arr& out, bool complex_data);  // device arrays in the solver's shapes, no host copies
  void export_eig(double* lam, double* q, double* p) const;
  // building blocks used by Navier2D's fused programs
  void emit_x(ProgBuilder& pb, int r, Lay lay) const;        // bandmv_x (+ fdma_x for ADI)
  void emit_y(ProgBuilder& pb, int r, int rinv, Lay lay, bool complex_lanes) const;  // bandmv_y + solve_y
  void gemm_fwd(const Arr& in, Arr& out, int ncols_real) const;
  void gemm_bwd(const Arr& in, Arr& out, int ncols_real) const;
  // per-mode tables of the specialised y kernels (fast.h ModeTabs) incl. the chunk-major packed copies; cached
  fk::ModeTabs mode_tabs();
  // true when solve(false) (real data) runs on the specialised kernels: HholtzAdi = xk_adi + yk_adi; Hholtz / Poisson
  // with Chebyshev x = b2x + DMMA GEMM + yk_mode + DMMA GEMM (else: generic lane programs)
  bool fast_path() const;
  int launches_per_solve(bool complex_data) const;

 private:
  Arr t1_[2], t2_[2], t3_[2];
  Built px_[2], py_[2];
  std::vector<double> hq_, hp_, lam_export_;
  void build_programs(bool complex_data);
  void solve_fast();
  std::vector<DevBuf> perm_;  // chunk-major coefficient tables (fast.h perm_table)
  bool have_mode_tabs_ = false, have_adi_tabs_ = false;
  fk::ModeTabs mode_tabs_;
  const double *adi_pt_[2][2] = {{nullptr, nullptr}, {nullptr, nullptr}};
};

// Navier2D (src/navier/navier.rs:153-195)
class Navier2D {
 public:
  Navier2D(int nx, int ny, double ra, double pr, double dt, double aspect, bool adiabatic, bool periodic,
           const EigData* eig);
  ~Navier2D();
  int nx, ny;
  bool periodic, adiabatic;
  double ra, pr, nu, ka, dt, time, scale[2];
  bool dealias = true;
  std::unique_ptr<Field2> temp, ux, uy, pres0, pres1, field;
  std::unique_ptr<Solver2> solver[4];
  cudaStream_t stream = 0;

  void set_velocity(double amp, double m, double n);
  void set_temperature(double amp, double m, double n);
  void set_tempbc_ortho(const double* that_bc);  // ortho coefficients of the BC field (o0 x o1 [x2])
  void set_solid(const double* mask, const double* value);  // [nx, ny] each; mask == nullptr: none
  void update(int nsteps);
  void eval(double* nu, double* nuvol, double* re, double* div_norm, double* ekin);
  void sync();
  Field2* field_by_index(int which);
  int launches_per_step() const { return launches_per_step_; }
  const Arr& tempbc_ortho() const { return tbc_ortho_; }
  void apply_ic(Field2& f, double amp, double m, double n, bool sin_cos);
  bool uses_specialised_kernels() {
    build_step();
    return !fast_ops_.empty();
  }
  struct OpInfo {
    std::string name;
    double bytes, flops;  // algorithmic bytes / flops of one launch
  };
  const std::vector<OpInfo>& op_info() {
    build_step();
    return opinfo_;
  }
  void profile(int reps, std::vector<double>& ms);
  // slab decomposition over kx (periodic path; rustpde_b200/slab.py drives the phases and owns the exchange buffers)
  // world > 0: fused transposes -- the outputs are scattered straight into the peers' buffers (peers[a * world + q])
  void slab_phase1(int k0, int mkl, double* const out[6], int world = 0, const int* joff = nullptr, double* const* peers = nullptr);
  void slab_phase2(int j0, int nyl, const double* const in[6], double* work, double* const out[3], int world = 0,
                   const int* koff = nullptr, double* const* peers = nullptr);
  void slab_phase3(int k0, int mkl, const double* const in[3]);
  void set_graph(bool on) {
    use_graph_ = on;
    graph_dirty_ = true;
  }
  // Double-buffered state upload (host buffers in the vhat layouts of temp, ux, uy, pres[0]): stage_state() queues
  // the host-to-device copies on a copy stream of its own and returns; commit_staged() makes the compute stream
  // wait for them and moves the staged state into place (device-to-device).  The copies of the next state
  // therefore overlap the update() of the current one.  Host buffers must stay valid (and should be pinned)
  // until the commit.
  void stage_state(const double* t, const double* u, const double* v, const double* p);
  void commit_staged();
  // asynchronous state download on a second copy stream (see navier.cu)
  void fetch_state(double* t, double* u, double* v, double* p);
  void fetch_wait();
  // checkpoint / restart in the reference's group / dataset layout (snapshot.cu)
  void write_snapshot(const char* path);
  void read_snapshot(const char* path);
  // exit() without a per-step host sync
  void div_async();
  bool div_poll(double* out, bool wait);

 private:
  void build_step();
  void build_step_confined();
  void build_step_periodic();
  void build_step_confined_fast();
  void build_step_periodic_fast();
  void add_fast(const char* name, double bytes, std::function<void()> fn);
  std::vector<std::function<void()>> fast_ops_;  // specialised kernels (fast.h), launched on `stream`
  void build_y_phase();
  void add_prog(ProgBuilder& pb, const char* name);
  void rebuild_bc();
  void run_step();
  void prepare_step();
  void copy_bc_to_field();  // field.vhat = ortho coefficients of the boundary-condition field
#ifndef RP_EMU
  void capture_graph();
#endif
  double div_norm();
  void enqueue_div2();
  Arr a1_, a2_;
  Arr snap_[4];
  double div_last_ = 0.0;
  bool div_have_ = false, div_pending_ = false;
  bool graph_dirty_ = true;
  bool use_graph_ = true;
  unsigned long long dyp_version_ = ~0ull;  // pres0 version dyp_ was computed from
  // work arrays
  Arr ax_[3], adx_[3];           // x-backward results: value and d/dx   [nx x my]
  Arr phys_[8];                  // ux, uy, dxu, dyu, dxv, dyv, dxT, dyT   [nx x ny]
  Arr bconv_[3];                 // conv after y-forward                   [nx x ny]
  Arr chat_[3];                  // conv after x-forward (periodic: complex (mk x ny))
  Arr w_[3];                     // after x-part of the implicit solve     [mx x ny]
  Arr vx_, ey_, div_, r1_, g_, h_, dyp_;
  Arr solid_mask_, solid_val_, phys_t_, tbc_phys_, zero_phys_;  // solid-mask penalisation (navier.rs:552-608)
  bool has_solid_ = false;
  void set_solid_args(fk::YConvArgs& a, int f);
  const Arr& zero_phys();
  Arr dxp_, xs_p_, xs_d_;        // -dt/sx d/dx pres; scratch of the column-scan projection  (confined, fast_xs.cu)
  Arr tbc_ortho_, dxtbc_, dytbc_, bcdiff_;
  std::vector<Built> step_;      // programs of one update() in launch order
  struct StepOp {
    int kind;  // 0 lane program, 1 gemm fwd, 2 gemm bwd, 3 zero elem, 4 specialised kernel
    int idx;
  };
  std::vector<StepOp> ops_;
  std::vector<OpInfo> opinfo_;
  void run_op(const StepOp& op);
  int launches_per_step_ = 0;
  DevBuf red_;
  Arr stage_[4];
  bool staged_ = false;
  // time-invariant boundary arrays: rows of tbc_ortho_ / bcdiff_ beyond these hold exact zeros (skipped by the specialised
  // confined kernels); dxtbc_ / dytbc_ identically zero (the Rayleigh-Benard boundary field varies in y only)
  int tbc_rows_ = 1 << 30, bcdiff_rows_ = 1 << 30;
  bool dxtbc_zero_ = false, dytbc_zero_ = false;
  fk::XwDivArgs xw_div_{};  // warp-serial divergence sweep (fast_xw.cu), when the step uses it
  std::function<void()> fast_dyp_, fast_div_;  // specialised refresh of d/dy pres and divergence of (ux, uy), when available
  std::vector<DevBuf> perm_;  // chunk-major coefficient tables of the specialised kernels (fast.h perm_table)
  const double* pw_rs_[4] = {nullptr, nullptr, nullptr, nullptr};  // stencil tables of the warp-serial periodic passes (fast_pw.cu): ux, uy, temp, divergence
  const double* pw_prj_[2] = {nullptr, nullptr};  // pw_project_tables (fast_pw.cu)
  void build_pw_tables();
  std::map<const void*, std::pair<const double*, const double*>> perm_mode_;
  std::map<std::pair<const void*, int>, std::pair<const double*, const double*>> perm_tdma_;
  fk::TdmaTabs tdma_of(const Base& b, int n, fk::ScanShape ng);
#ifndef RP_EMU
  cudaStream_t copy_stream_ = nullptr, fetch_stream_ = nullptr;
  cudaEvent_t ev_staged_ = nullptr, ev_consumed_ = nullptr, ev_fetch_ready_ = nullptr, ev_fetched_ = nullptr, ev_div_ = nullptr;
  double* div_host_ = nullptr;
#endif
#ifndef RP_EMU
  cudaGraphExec_t graph_ = nullptr;
  bool graph_ok_ = false;
#endif
};

// Navier2DAdjoint (src/navier/navier_adjoint.rs:128-176)
class Navier2DAdjoint {
 public:
  Navier2DAdjoint(int nx, int ny, double ra, double pr, double dt, double aspect, bool adiabatic, bool periodic);
  ~Navier2DAdjoint();
  int nx, ny;
  bool periodic;
  double ra, pr, nu, ka, dt, dt_navier, time = 0.0, scale[2], res_tol = 1e-8;
  bool dealias = true;
  std::unique_ptr<Navier2D> navier;
  std::unique_ptr<Field2> temp[2], ux[2], uy[2], pres[2], field;  // [adjoint field, Navier-Stokes residual]
  std::unique_ptr<Solver2> solver_pres, smoother[3];
  cudaStream_t stream = 0;
  void set_velocity(double amp, double m, double n);
  void set_temperature(double amp, double m, double n);
  void update(int nsteps);
  void eval(double* nu, double* nuvol, double* re, double* div_norm);
  void residuals(double smooth[3], double unsmooth[3]);
  bool exit();
  Field2* field_by_index(int which);    // 0 temp, 1 ux, 2 uy, 3 pres, 4 pseudo pressure, 5..7 residual temp / ux / uy
  Solver2* solver_by_index(int which);  // 0 smoother ux|uy, 1 smoother temp, 2 pressure Poisson, 3 inner Navier2D's Poisson

 private:
  void conv_term(Field2& f, const Arr& u, int d0, int d1);
  void finish_conv();
  void conv_u(int comp);
  void solve_u(int comp);
  void solve_temp();
  void divergence_to_rhs();
  void update_residual();
  double norm_l2(const Arr& a);
  Arr rhs_, unsm_[3], phys_[3], conv_, bcv_, old_;
  DevBuf red_;
};

// LAPACK access for the set-up eigendecomposition (src/solver/utils.rs:66-106)
void lapack_set_library(const char* path);
bool lapack_available(std::string* why);
void lapack_eig_setup(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                      std::vector<double>& Q, std::vector<double>& P);
void lapack_eig_setup_parity(int m, const std::vector<double>& Cx, const std::vector<double>& Ax, std::vector<double>& lam,
                             std::vector<double>& Q, std::vector<double>& P);

}  // namespace rp
