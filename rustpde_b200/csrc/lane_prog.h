// lane_prog.h -- "lane programs": the instruction format shared by the host
// program builders (navier.cpp, field.cpp, solver.cpp) and the device
// interpreter (lane_vm.cuh).
//
// A lane program runs on one thread block that owns T "slots".  A slot holds
// one *packed lane* per register: a double2 array whose .x/.y components are
// two independent real lanes (two adjacent rows or columns of a real array) or
// the re/im parts of one complex lane.  Every operator on the Navier2D path is
// linear with real coefficients along the lane, so both components ride
// through the same arithmetic -- this is what lets one complex FFT serve two
// real DCT-I / r2c lanes.
//
// Element i of a lane lives at slot  sigma(i) = (i odd ? o0 + so*(i/2) : e0 + se*(i/2))
// (struct Lay).  NATURAL = {0,2,1,2}; SPLIT(A) = {0,1,A,-1} keeps the even-
// and odd-index chains contiguous, which is what the stride-2 recurrences
// (Chebyshev derivative, TDMA, FDMA) and the DCT-I post-processing want.
#pragma once
#include <cstdint>

#include "rt.h"

namespace rp {

struct Lay {
  int e0, se, o0, so;
};
RP_HD static inline Lay lay_natural() { return Lay{0, 2, 1, 2}; }
RP_HD static inline Lay lay_split(int anchor) { return Lay{0, 1, anchor, -1}; }
// layout after one Chebyshev differentiation (see op DIFF)
RP_HD static inline Lay lay_after_diff(Lay l) { return Lay{l.o0, l.so, l.e0 + l.se, l.se}; }

enum Op : int {
  OP_END = 0,
  OP_LD,         // r0 (+)= coef * src[lane+shift][0..n)         ; zero-fill [n, n2)
  OP_ST,         // dst[lane][0..n) (+)= s0 * r0                  ; cut -> zeros
  OP_ZERO,       // r0[0..n) = 0
  OP_COPY,       // r0 = r1                      (lay, lay2)
  OP_AXPY,       // r0 += s0 * r1                (lay, lay2)
  OP_SCALE,      // r0 *= s0
  OP_MULPW,      // r0 (+)= r1 (.) r2   componentwise
  OP_CUT,        // r0[i0..n) = 0
  OP_MULIK,      // r0 *= i * k_lane * s0        (complex lanes)
  OP_TOORTHO,    // composite (n-2) -> ortho (n), tables p0=d, p1=l
  OP_FROMORTHO,  // ortho (n) -> composite (n-2), tables p0=d, p1=l, p2=tdma
  OP_DIFF,       // Chebyshev derivative, i0 times, scale s0 ; layout changes
  OP_DCT,        // DCT-I with Chebyshev scaling, i0 = 0 fwd / 1 bwd; p0 = DctPlan
  OP_RFFT,       // r0 (real pair, n) -> r1, r2 (complex, n/2+1) ; p0 = FftPlan
  OP_IRFFT,      // r1, r2 -> r0
  OP_BANDMV,     // B2 preconditioner (n) -> (n-2), tables p0,p1,p2
  OP_FDMA,       // banded solve, shared pre-swept factors p0 = FdmaTab
  OP_FDMAMODE,   // banded solve, per-lane (A + (lam+alpha) C); r1 = 1/dia' lane
  OP_SETZERO00,  // element 0 of lane 0 := 0
};

enum LdFlags : int {
  LF_ACC = 1,        // accumulate instead of set
  LF_COMPLEX = 2,    // array holds double2 (complex) lanes; else real pair
  LF_BCAST = 4,      // real array, one lane broadcast to both components
  LF_MULIK = 8,      // multiply by i * k_lane * s0 (complex lanes only)
  LF_LANECOEF = 16,  // multiply by p1[source lane]
  LF_CUT = 32,       // ST: elements >= i0 are written as 0; lanes >= i1 are written as 0
  LF_LANE2 = 64,     // complex array, lane = 2*P + i2 (two complex lanes per slot)
  LF_ELEMK = 128,    // MULIK: wavenumber = element index (x lanes) instead of lane index
};

struct Instr {
  int op;
  int r0, r1, r2;
  int n, n2;
  int flags;
  int shift;
  Lay lay, lay2;
  const void* p0;
  const void* p1;
  const void* p2;
  const void* p3;
  long long ld;  // leading dimension of the array (elements of its own type)
  int nlanes;    // valid lanes of the array along the lane-index axis
  int i0, i1, i2;
  double s0, s1;
};

enum { AXIS_X = 0, AXIS_Y = 1 };
enum { RP_CHUNK_HOST = 32 };  // chain chunk length of the block-parallel recurrences
enum { RP_MAX_INSTR = 40 };

struct Program {
  int ninstr;
  int axis;   // AXIS_Y: lanes are rows (contiguous); AXIS_X: lanes are columns (strided)
  int T;      // slots per block
  int nreg;   // lane registers per slot
  int cap;    // register capacity (double2 elements, un-padded)
  int wb_cap; // Bluestein work-buffer capacity per buffer (0 = none)
  int wb_T;   // number of work buffers
  int nunits; // number of slot units to process (pairs of real lanes / complex lanes)
  int nthreads;
  int smem_bytes;
  long long* prof;  // development aid: per-opcode cycle counters of block 0 (NULL = off)
  Instr ins[RP_MAX_INSTR];
};

// ---- device-resident plan tables ----------------------------------------
struct FftPlan {
  int L;            // transform length
  int pow2;         // 1: radix passes on L; 0: Bluestein with Lb
  int Lb;           // Bluestein FFT length (pow2)
  const double2* tw;     // exp(-2 pi i k / L) (pow2) or exp(-2 pi i k / Lb)
  const double2* chirp;  // conj chirp  exp(-i pi j^2 / L), j < L     (Bluestein)
  const double2* bhat;   // FFT_Lb of the wrapped chirp, times 1/Lb   (Bluestein)
};

struct DctPlan {
  int n;                 // lane length, N = n - 1
  FftPlan fft;           // complex DFT of length N
  const double2* sc;     // (sin, cos)(pi j / N), j = 0..N/2
};

struct TdmaTab {  // (S^T S) solve of from_ortho, pre-factored (composite_stencil.rs:250-276)
  const double* fs;  // g_i = fs_i * d_i + fp_i * g_{i-2}
  const double* fp;
  const double* bp;  // x_i = g_i + bp_i * x_{i+2}
};

struct FdmaTab {  // pre-swept Fdma (fdma.rs:73-118)
  const double* fp;   // x_i += fp_i * x_{i-2}            (= -low_{i-2})
  const double* bs;   // x_i = bs_i * x_i + bp1_i * x_{i+2} + bp2_i * x_{i+4}
  const double* bp1;
  const double* bp2;
};

struct FdmaModeTab {  // per-lane A + (lam + alpha) C, raw diagonals (fdma_tensor.rs:219-227)
  const double *a_low, *a_dia, *a_up1, *a_up2;
  const double *c_low, *c_dia, *c_up1, *c_up2;
  const double* lam;  // per lane
  double alpha;
};

}  // namespace rp
