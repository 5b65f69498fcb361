// field.cu -- Field2 (src/field.rs:66-129) on the device.
//
// forward  = y then x, backward = x then y, to/from_ortho and gradient =
// axis 0 then axis 1 (funspace/src/space2.rs:182-356).  Each axis pass is one
// lane program; the intermediate lives in a field-owned scratch array.
#include <cmath>

#include "model.h"

namespace rp {

static std::vector<double> grid_dx(const std::vector<double>& x, bool periodic) {  // field.rs:135-163
  const int n = (int)x.size();
  std::vector<double> dx(n);
  if (periodic) {
    for (auto& d : dx) d = x[2] - x[1];
    return dx;
  }
  for (int i = 0; i < n; ++i) {
    const double xl = (i == 0) ? x[0] : (x[i] + x[i - 1]) / 2.0;
    const double xr = (i == n - 1) ? x[n - 1] : (x[i + 1] + x[i]) / 2.0;
    dx[i] = xr - xl;
  }
  return dx;
}

Field2::Field2(const Space2& s) : sp(s) {
  const Base &b0 = *sp.b0, &b1 = *sp.b1;
  if (!b1.is_cheb()) throw Error(RP_ERR_INVALID, "axis 1 must be a Chebyshev-family base (Space2R2r / Space2R2c)");
  if (b0.is_bc() || b1.is_bc())
    throw Error(RP_ERR_INVALID, "boundary-condition bases are set-up helpers; use rp_navier_set_tempbc_ortho");
  n0 = b0.n;
  n1 = b1.n;
  m0 = b0.m;
  m1 = b1.m;
  cplx = !b0.is_cheb();
  o0 = cplx ? m0 : n0;
  o1 = n1;
  v.alloc(n0, n1, false);
  vhat.alloc(m0, m1, cplx);
  ortho.alloc(o0, o1, cplx);
  ta_.alloc(n0, m1, false);
  tb_.alloc(std::max(o0, m0), std::max(o1, m1), cplx);
  x[0] = b0.x;
  x[1] = b1.x;
  dx[0] = grid_dx(x[0], !b0.is_cheb());
  dx[1] = grid_dx(x[1], false);
  build_transforms();
}

static int half_up(int n) { return (n + 1) / 2; }

void Field2::build_transforms() {
  const Base &b0 = *sp.b0, &b1 = *sp.b1;
  const Lay nat = lay_natural();
  {  // forward, y: physical rows -> spectral-y rows
    ProgBuilder pb(AXIS_Y, half_up(n0));
    pb.ld(0, v, n1, nat);
    Lay l = pb.dct(0, b1, nat, false);
    pb.fromortho(0, b1, l);
    pb.st(0, ta_, m1, l);
    fwd_y_ = pb.build();
  }
  {  // forward, x
    ProgBuilder pb(AXIS_X, half_up(m1));
    if (b0.is_cheb()) {
      pb.ld(0, ta_, n0, nat);
      Lay l = pb.dct(0, b0, nat, false);
      pb.fromortho(0, b0, l);
      pb.st(0, vhat, m0, l);
    } else {
      pb.ld(0, ta_, n0, nat);
      pb.rfft_st(0, vhat, b0);
    }
    fwd_x_ = pb.build();
  }
  {  // backward, x
    ProgBuilder pb(AXIS_X, half_up(m1));
    if (b0.is_cheb()) {
      Lay l = lay_split(n0);
      pb.ld(0, vhat, m0, l, 1.0, 0, 0, nullptr, n0);
      pb.toortho(0, b0, l);
      l = pb.dct(0, b0, l, true);
      pb.st(0, ta_, n0, l);
    } else {
      pb.irfft_ld(0, vhat, b0);
      pb.st(0, ta_, n0, nat);
    }
    bwd_x_ = pb.build();
  }
  {  // backward, y
    ProgBuilder pb(AXIS_Y, half_up(n0));
    Lay l = lay_split(n1);
    pb.ld(0, ta_, m1, l, 1.0, 0, 0, nullptr, n1);
    pb.toortho(0, b1, l);
    l = pb.dct(0, b1, l, true);
    pb.st(0, v, n1, l);
    bwd_y_ = pb.build();
  }
  // to_ortho / from_ortho (composite.rs:255-316); x pass only when base 0 is composite
  const bool xcomp = b0.is_composite();
  if (xcomp) {
    ProgBuilder pb(AXIS_X, half_up(m1));
    Lay l = lay_split(n0);
    pb.ld(0, vhat, m0, l, 1.0, 0, 0, nullptr, n0);
    pb.toortho(0, b0, l);
    pb.st(0, tb_, n0, l);
    to_x_ = pb.build();
    ProgBuilder pf(AXIS_X, half_up(o1));
    pf.ld(0, ortho, n0, l);
    pf.fromortho(0, b0, l);
    pf.st(0, tb_, m0, l);
    from_x_ = pf.build();
  }
  {
    const Arr& src = xcomp ? tb_ : vhat;
    ProgBuilder pb(AXIS_Y, cplx ? o0 : half_up(o0));
    Lay l = lay_split(n1);
    pb.ld(0, src, m1, l, 1.0, 0, 0, nullptr, n1);
    pb.toortho(0, b1, l);
    pb.st(0, ortho, o1, l);
    to_y_ = pb.build();
    const Arr& fsrc = xcomp ? tb_ : ortho;
    ProgBuilder pf(AXIS_Y, cplx ? m0 : half_up(m0));
    pf.ld(0, fsrc, o1, l);
    pf.fromortho(0, b1, l);
    pf.st(0, vhat, m1, l);
    from_y_ = pf.build();
  }
  // averaging weights dx/|x_last - x_0| (average.rs:25-57)
  std::vector<double> wx(n0), wy(n1);
  const double lx = std::fabs(x[0][x[0].size() - 1] - x[0][0]), ly = std::fabs(x[1][n1 - 1] - x[1][0]);
  for (int i = 0; i < n0; ++i) wx[i] = dx[0][i] / lx;
  for (int j = 0; j < n1; ++j) wy[j] = dx[1][j] / ly;
  wx_ = upload(wx);
  wy_ = upload(wy);
  red_ = DevBuf(sizeof(double) * (RP_WSUM_DOUBLES + (size_t)(RP_AVG0_ROWPARTS + 1) * n1));
}

void Field2::forward() {
  ++vhat_version;
  fwd_y_.launch(stream);
  fwd_x_.launch(stream);
}
void Field2::backward() {
  bwd_x_.launch(stream);
  bwd_y_.launch(stream);
}
void Field2::to_ortho() {
  if (to_x_.valid) to_x_.launch(stream);
  to_y_.launch(stream);
}
void Field2::from_ortho() {
  ++vhat_version;
  if (from_x_.valid) from_x_.launch(stream);
  from_y_.launch(stream);
}

void Field2::gradient(int ddx, int ddy, const double* scale) {
  const Base &b0 = *sp.b0, &b1 = *sp.b1;
  const double sx = scale ? std::pow(scale[0], ddx) : 1.0, sy = scale ? std::pow(scale[1], ddy) : 1.0;
  std::array<long long, 4> key = {ddx, ddy, 0, 0};
  memcpy(&key[2], &sx, 8);
  memcpy(&key[3], &sy, 8);
  auto it = grad_.find(key);
  if (it == grad_.end()) {
    std::pair<Built, Built> pr;
    const bool xpass = b0.is_cheb() && (b0.is_composite() || ddx > 0);
    if (xpass) {  // differentiate along x (space2.rs:247-264)
      ProgBuilder pb(AXIS_X, half_up(m1));
      Lay l = lay_split(n0);
      pb.ld(0, vhat, m0, l, 1.0, 0, 0, nullptr, n0);
      pb.toortho(0, b0, l);
      l = pb.diff(0, n0, l, ddx, 1.0 / sx);
      pb.st(0, tb_, n0, l);
      pr.first = pb.build();
    }
    {
      const Arr& src = xpass ? tb_ : vhat;
      ProgBuilder pb(AXIS_Y, cplx ? o0 : half_up(o0));
      Lay l = lay_split(n1);
      pb.ld(0, src, m1, l, 1.0, 0, 0, nullptr, n1);
      if (!b0.is_cheb())
        for (int k = 0; k < ddx; ++k) pb.mulik(0, m1, 1.0, l);  // (ik)^ddx, r2c.rs:88-99
      pb.toortho(0, b1, l);
      double sc = 1.0 / sy;
      if (!b0.is_cheb()) sc /= sx;
      l = pb.diff(0, n1, l, ddy, sc);
      pb.st(0, ortho, o1, l);
      pr.second = pb.build();
    }
    it = grad_.emplace(key, std::move(pr)).first;
  }
  if (it->second.first.valid) it->second.first.launch(stream);
  it->second.second.launch(stream);
}

double Field2::average() {
  launch_wsum(v.d(), nullptr, v.ld, n0, n1, wx_.as<double>(), wy_.as<double>(), 0, red_.as<double>(), stream);
  double r = 0.0;
  rt::d2h(&r, red_.p, 8, stream);
  rt::sync(stream);
  return r;
}

// average.rs:25-33 along axis 0, reduced on the device (fixed summation order); only n1 values come back
void Field2::average_axis0(std::vector<double>& out) {
  double* scratch = red_.as<double>() + RP_WSUM_DOUBLES;
  double* res = scratch + (size_t)RP_AVG0_ROWPARTS * n1;
  launch_avg_axis0(v.d(), v.ld, n0, n1, wx_.as<double>(), scratch, res, stream);
  out.assign(n1, 0.0);
  rt::d2h(out.data(), res, sizeof(double) * n1, stream);
  rt::sync(stream);
}

}  // namespace rp
