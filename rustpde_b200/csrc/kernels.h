// kernels.h -- launch wrappers of the CUDA kernels (kernels.cu).
#pragma once
#include "lane_prog.h"

namespace rp {

enum { RP_MAX_SMEM = 227 * 1024 };

void init_kernels();  // opt the kernels in to > 48 KB dynamic shared memory (idempotent)
void launch_lane_programs(const Program* d_progs, int nprogs, int nblocks, int nthreads, int smem_bytes, cudaStream_t s);

struct GemmArgs {
  const double* A;  // M x K, row-major, lda
  const double* B;  // K x N
  double* C;        // M x N
  int M, N, K;
  long long lda, ldb, ldc;
  int b_r0, b_rs, c_r0, c_rs;  // row offset / row stride multiplier of B and C
};
void launch_dgemm(const GemmArgs& g, cudaStream_t s);
void launch_dgemm2(const GemmArgs& g0, const GemmArgs& g1, cudaStream_t s);  // two independent products, one launch

void launch_zero_elems(double* p, int count, cudaStream_t s);
// weighted sum with a fixed summation order; `out` holds RP_WSUM_DOUBLES doubles (out[0] = result, rest scratch)
enum { RP_WSUM_DOUBLES = 1 + 592 };
void launch_wsum(const double* a, const double* b, long long ld, int rows, int cols, const double* wx, const double* wy,
                 int mode, double* out, cudaStream_t s);
// out[j] = sum_i w[i] a[i][j]; scratch holds RP_AVG0_ROWPARTS * cols doubles
enum { RP_AVG0_ROWPARTS = 32 };
void launch_avg_axis0(const double* a, long long ld, int rows, int cols, const double* w, double* scratch, double* out,
                      cudaStream_t s);
void launch_combine(double* out, const double* a, const double* b, const double* c, long long ld, int rows, int cols,
                    double s0, double s1, cudaStream_t s);

void launch_dealias(double* p, long long ld, int rows, int cols, int rc, int cut_row, int cut_col, cudaStream_t s);
void launch_b2x(const double* in, long long ldi, double* out, long long ldo, int n, int cols, const double* lo, const double* di,
                const double* up, cudaStream_t s);

}  // namespace rp
