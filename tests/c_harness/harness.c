/* harness.c -- plain C (gcc) caller of the C ABI, no Python / ctypes in between.
 *
 * Mirrors what the reference's README example and examples/hholtz_2d.rs do through the Rust API:
 *   Navier2D::new(64, 64, 1e5, 1.0, 0.02, 1.0, true); set_velocity(0.2,1,1); set_temperature(0.2,1,1);
 *   5 x update(); eval Nu / Nuvol / Re / |div|            (navier.rs:139-152, README)
 *   Hholtz::new2(&field, [1,1], 10.0).solve(rhs)          (examples/hholtz_2d.rs:7-31, 34 x 33)
 * and prints every number with 17 significant digits; tests/test_c_harness.py compares them with the oracle.
 *
 *   gcc -O2 -I include tests/c_harness/harness.c -o harness -L rustpde_b200 -lrustpde_b200 -lm
 *   ./harness /path/to/liblapack-provider.so
 */
#include <math.h>
#include <stdio.h>
#include <stdlib.h>

#include "rustpde_b200.h"

#define CHECK(call)                                                            \
  do {                                                                         \
    int rc_ = (call);                                                          \
    if (rc_ != RP_OK) {                                                        \
      fprintf(stderr, "%s failed (%d): %s\n", #call, rc_, rp_last_error());    \
      return 1;                                                                \
    }                                                                          \
  } while (0)

int main(int argc, char** argv) {
  if (argc > 1) CHECK(rp_set_lapack_library(argv[1]));
  CHECK(rp_init(0));
  printf("emulated %d version %d\n", rp_is_emulated(), rp_version());

  /* ---- Navier2D, config 1 ---- */
  rp_navier_t* nav = NULL;
  CHECK(rp_navier_create(64, 64, 1e5, 1.0, 0.02, 1.0, 1, 0, &nav));
  CHECK(rp_navier_set_velocity(nav, 0.2, 1.0, 1.0));
  CHECK(rp_navier_set_temperature(nav, 0.2, 1.0, 1.0));
  CHECK(rp_navier_update(nav, 5));
  double nu, nuvol, re, divn, ekin, t;
  CHECK(rp_navier_eval(nav, &nu, &nuvol, &re, &divn, &ekin));
  CHECK(rp_navier_get_time(nav, &t));
  printf("navier time %.17g nu %.17g nuvol %.17g re %.17g div %.17g ekin %.17g\n", t, nu, nuvol, re, divn, ekin);
  rp_field_t* temp = NULL;
  CHECK(rp_navier_field(nav, RP_FIELD_TEMP, &temp));
  int phys[2], spec[2], ortho[2], cx;
  CHECK(rp_field_shape(temp, phys, spec, ortho, &cx));
  size_t len = (size_t)spec[0] * spec[1] * (cx ? 2 : 1);
  double* vhat = (double*)malloc(len * sizeof(double));
  CHECK(rp_field_download_vhat(temp, vhat, len));
  printf("temp_vhat %d %d", spec[0], spec[1]);
  for (int i = 0; i < 4; ++i)
    for (int j = 0; j < 4; ++j) printf(" %.17g", vhat[(size_t)i * spec[1] + j]);
  printf("\n");
  free(vhat);
  /* eigen set-up data of the pressure solver, so that the checker can feed the same (lam, Q, P) to the oracle */
  {
    const int m = 62;
    double* lam = (double*)malloc(sizeof(double) * m);
    double* q = (double*)malloc(sizeof(double) * m * m);
    double* p = (double*)malloc(sizeof(double) * m * m);
    CHECK(rp_navier_export_eig(nav, lam, q, p));
    FILE* f = fopen(argc > 2 ? argv[2] : "harness_eig.bin", "wb");
    if (!f) return 2;
    fwrite(lam, sizeof(double), m, f);
    fwrite(q, sizeof(double), (size_t)m * m, f);
    fwrite(p, sizeof(double), (size_t)m * m, f);
    fclose(f);
    free(lam), free(q), free(p);
  }
  CHECK(rp_navier_destroy(nav));

  /* ---- Hholtz::new2 on cheb_dirichlet x cheb_dirichlet 34 x 33, analytic rhs ---- */
  rp_field_t* f = NULL;
  const int nx = 34, ny = 33;
  CHECK(rp_field_create(RP_BASE_CHEB_DIRICHLET, nx, RP_BASE_CHEB_DIRICHLET, ny, &f));
  double *x = (double*)malloc(sizeof(double) * nx), *y = (double*)malloc(sizeof(double) * ny);
  CHECK(rp_field_coords(f, 0, x, nx));
  CHECK(rp_field_coords(f, 1, y, ny));
  double* v = (double*)malloc(sizeof(double) * nx * ny);
  const double n = acos(-1.0) / 2.0, alpha = 0.1;
  for (int i = 0; i < nx; ++i)
    for (int j = 0; j < ny; ++j) v[i * ny + j] = cos(n * x[i]) * cos(n * y[j]);
  CHECK(rp_field_upload_v(f, v, (size_t)nx * ny));
  CHECK(rp_field_forward(f));
  double* rhs = (double*)malloc(sizeof(double) * nx * ny);
  CHECK(rp_field_to_ortho(f, rhs, (size_t)nx * ny));
  rp_solver_t* h = NULL;
  CHECK(rp_hholtz_create(f, 1.0, 1.0, 1.0 / alpha, &h));
  double* sol = (double*)malloc(sizeof(double) * (nx - 2) * (ny - 2));
  CHECK(rp_solver_solve(h, rhs, (size_t)nx * ny, sol, (size_t)(nx - 2) * (ny - 2), 0));
  /* size mismatch must be reported, not crash (reference: panic!, fdma_tensor.rs:201-209) */
  const int rc = rp_solver_solve(h, rhs, (size_t)nx * ny - 1, sol, (size_t)(nx - 2) * (ny - 2), 0);
  printf("shape_error_code %d\n", rc);
  CHECK(rp_field_upload_vhat(f, sol, (size_t)(nx - 2) * (ny - 2)));
  CHECK(rp_field_backward(f));
  double* back = (double*)malloc(sizeof(double) * nx * ny);
  CHECK(rp_field_download_v(f, back, (size_t)nx * ny));
  double err = 0.0;
  for (int i = 0; i < nx * ny; ++i) {
    const double e = fabs(back[i] - alpha / (1.0 + alpha * n * n * 2.0) * v[i]);
    if (e > err) err = e;
  }
  printf("hholtz analytic_max_abs_err %.17g sol00 %.17g sol11 %.17g\n", err, sol[0], sol[(ny - 2) + 1]);
  CHECK(rp_solver_destroy(h));
  CHECK(rp_field_destroy(f));
  free(x), free(y), free(v), free(rhs), free(sol), free(back);
  printf("HARNESS_OK\n");
  return 0;
}
