// fast.cuh -- device building blocks of the specialised Navier2D kernels
// (fast_x.cu, fast_y.cu).
//
// A thread block owns a *tile* in shared memory: `rows` elements along the
// transformed axis times 4 real lanes (4 adjacent columns for an x pass, 4
// adjacent rows for a y pass).  The 4 lanes are the fastest dimension, i.e. the
// tile is a `double[rows][4]` == `double2[rows][2]` array: two packed complex
// lanes whose re/im parts are two independent real lanes.  Every operator on the
// path is real-linear along the lane, so one complex FFT serves two real
// DCT-I lanes (ortho.rs:337-407 via a real DFT of half length).
//
// Because the lanes are the fastest dimension, consecutive threads always touch
// consecutive shared-memory words in every FFT pass; the only pattern that
// would collide (first Stockham pass, output stride 8 rows) is removed by the
// row swizzle prow().
//
// Sequential recurrences of the reference (Chebyshev derivative ortho.rs:107-125,
// TDMA linalg.rs:14-57, FDMA fdma.rs:101-118) run as chunked scans: each thread
// owns a chunk of one parity chain of one lane, caches it in registers, the
// chunk maps are chained through a small shared array, and the chunk is
// re-walked with its carry-in.
#pragma once
#include <utility>

#include "fast.h"

namespace rp {
namespace fk {

typedef double2 cplx;
#define FK_DEV __device__ __forceinline__

FK_DEV cplx mk(double x, double y) { return make_double2(x, y); }
FK_DEV cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
FK_DEV cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
FK_DEV cplx cmul(cplx a, cplx b) { return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
FK_DEV cplx csqr(cplx a) { return mk(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }
FK_DEV cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }

// L2 prefetch of the 32-byte sector that holds *p (no register or shared-memory cost): used to pull
// the column strips a later phase (or the block that will run next on this SM) reads out of DRAM
// while the current phase computes.
FK_DEV void prefetch_l2(const void* p) {
#ifndef RP_EMU
  asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
#else
  (void)p;
#endif
}
// ---- asynchronous global -> shared copies (LDGSTS) ---------------------------------------------
// 16-byte chunk; only the first `bytes` (0, 8 or 16) are read, the rest is zero-filled
FK_DEV void cp_async16(void* smem_dst, const double* gsrc, int bytes) {
#ifdef RP_EMU
  double* d = (double*)smem_dst;
  d[0] = bytes >= 8 ? gsrc[0] : 0.0;
  d[1] = bytes >= 16 ? gsrc[1] : 0.0;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;" ::"r"(d), "l"(gsrc), "r"(bytes) : "memory");
#endif
}
// 8-byte asynchronous copy (one double)
FK_DEV void cp_async8(void* smem_dst, const double* gsrc) {
#ifdef RP_EMU
  *(double*)smem_dst = *gsrc;
#else
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(d), "l"(gsrc) : "memory");
#endif
}
FK_DEV void cp_async_commit() {
#ifndef RP_EMU
  asm volatile("cp.async.commit_group;" ::: "memory");
#endif
}
template <int N>  // wait until at most N of the most recent groups are still in flight
FK_DEV void cp_async_wait() {
#ifndef RP_EMU
  asm volatile("cp.async.wait_group %0;" ::"n"(N) : "memory");
#endif
}

// at most 512 threads of a block take part in the scans (bounds the scratch array); the rest only
// joins the barriers
RP_HD constexpr int scan_threads(int nthr) { return nthr > 512 ? 512 : nthr; }

// ---- tile addressing --------------------------------------------------------
// LC = packed complex lanes per tile row (2: the default 4-real-lane tile; 1: long lanes whose 4-lane tile
// would not fit shared memory).  A row is LC * 16 bytes; the swizzle spreads 8 / LC-row groups.
template <int LC>
FK_DEV int prow(int i) {
  return i ^ ((i >> 3) & (8 / LC - 1));
}
template <int LC>
FK_DEV int didx(int i, int r) {  // as double
  return prow<LC>(i) * (2 * LC) + r;
}
template <int LC>
FK_DEV int cidx(int i, int c) {  // as double2
  return prow<LC>(i) * LC + c;
}
// layout of a lane inside a tile: natural (sn < 0: element i at row i) or
// split(sn) (even i at row i/2, odd i at row sn - i/2), the output layout of the DCT
FK_DEV int rowof(int sn, int i) { return sn < 0 ? i : ((i & 1) ? sn - (i >> 1) : (i >> 1)); }

// ---- radix butterflies (forward, e^{-2 pi i / R}), in-order output ----------
template <int R>
struct Bfly;
template <>
struct Bfly<2> {
  static FK_DEV void run(cplx* v) {
    const cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
FK_DEV void dft4(cplx& x0, cplx& x1, cplx& x2, cplx& x3) {
  const cplx s0 = cadd(x0, x2), d0 = csub(x0, x2), s1 = cadd(x1, x3), d1 = csub(x1, x3);
  x0 = cadd(s0, s1);
  x2 = csub(s0, s1);
  x1 = mk(d0.x + d1.y, d0.y - d1.x);
  x3 = mk(d0.x - d1.y, d0.y + d1.x);
}
template <>
struct Bfly<4> {
  static FK_DEV void run(cplx* v) { dft4(v[0], v[1], v[2], v[3]); }
};
template <>
struct Bfly<8> {
  static FK_DEV void run(cplx* v) {
    dft4(v[0], v[2], v[4], v[6]);
    dft4(v[1], v[3], v[5], v[7]);
    const double h = 0.70710678118654752440;
    const cplx b0 = v[1];
    const cplx b1 = mk(h * (v[3].x + v[3].y), h * (v[3].y - v[3].x));
    const cplx b2 = mk(v[5].y, -v[5].x);
    const cplx b3 = mk(h * (v[7].y - v[7].x), -h * (v[7].x + v[7].y));
    const cplx a0 = v[0], a1 = v[2], a2 = v[4], a3 = v[6];
    v[0] = cadd(a0, b0);
    v[4] = csub(a0, b0);
    v[1] = cadd(a1, b1);
    v[5] = csub(a1, b1);
    v[2] = cadd(a2, b2);
    v[6] = csub(a2, b2);
    v[3] = cadd(a3, b3);
    v[7] = csub(a3, b3);
  }
};

// v[r] *= w^r from the single table value w (log-depth products)
template <int R>
FK_DEV void apply_twiddles(cplx* v, cplx w1) {
  v[1] = cmul(v[1], w1);
  if constexpr (R > 2) {
    const cplx w2 = csqr(w1);
    v[2] = cmul(v[2], w2);
    const cplx w3 = cmul(w2, w1);
    v[3] = cmul(v[3], w3);
    if constexpr (R > 4) {
      const cplx w4 = csqr(w2);
      v[4] = cmul(v[4], w4);
      v[5] = cmul(v[5], cmul(w4, w1));
      v[6] = cmul(v[6], cmul(w4, w2));
      v[7] = cmul(v[7], cmul(w4, w3));
    }
  }
}

// One in-place Stockham pass of radix R over both complex lanes of the tile.
// L = 1 << LOG2L rows; tw = per-span compact twiddles: tw[S/2 - 1 + q] = exp(-2 pi i q / S) (tables.cu).
// MUL: the loaded values are first multiplied by mulv[row] (Bluestein filter).
template <int LC, int LOG2L, int NTHR, int R, int NS, bool CONJ_IN, bool CONJ_OUT, bool MUL>
FK_DEV void fft_pass(cplx* tc, const cplx* __restrict__ tw, const cplx* __restrict__ mulv) {
  constexpr int L = 1 << LOG2L;
  constexpr int NB = L / R;
  constexpr int TOT = NB * LC;
  constexpr int KB = (TOT + NTHR - 1) / NTHR;
  const int tid = threadIdx.x;
  cplx v[KB][R];
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    const int b = tid + kb * NTHR;
    if (TOT % NTHR == 0 || b < TOT) {
      const int q = b / LC, c = b % LC;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        cplx x = tc[cidx<LC>(q + r * NB, c)];
        if (MUL) x = cmul(x, __ldg(&mulv[q + r * NB]));
        if (CONJ_IN) x.y = -x.y;
        v[kb][r] = x;
      }
      if (NS > 1) {
        const int k = q & (NS - 1);
        apply_twiddles<R>(v[kb], __ldg(&tw[NS * R / 2 - 1 + k]));  // compact table: exp(-2 pi i k / (NS R))
      }
      Bfly<R>::run(v[kb]);
    }
  }
  __syncthreads();
#pragma unroll
  for (int kb = 0; kb < KB; ++kb) {
    const int b = tid + kb * NTHR;
    if (TOT % NTHR == 0 || b < TOT) {
      const int q = b / LC, c = b % LC;
      const int k = q & (NS - 1);
      const int o = (q - k) * R + k;
#pragma unroll
      for (int r = 0; r < R; ++r) {
        cplx x = v[kb][r];
        if (CONJ_OUT) x.y = -x.y;
        tc[cidx<LC>(o + r * NS, c)] = x;
      }
    }
  }
  __syncthreads();
}

// Complex FFT of length L = 1 << LOG2L on rows [0, L) of the tile (both complex
// lanes), natural order in and out, radix-8 passes plus one radix-2/4 pass.
// INV: unnormalised inverse (conjugate in / out).  MUL: input multiplied by mulv[row].
template <int LC, int LOG2L, int NTHR, int LOG2NS, bool INV, bool MUL>
FK_DEV void fft_rec(cplx* tc, const cplx* __restrict__ tw, const cplx* __restrict__ mulv) {
  constexpr int REM = LOG2L - LOG2NS;  // log2 of what is left
  if constexpr (REM > 0) {
    constexpr int LR = REM >= 3 ? 3 : REM;
    constexpr bool first = (LOG2NS == 0), last = (REM == LR);
    fft_pass<LC, LOG2L, NTHR, (1 << LR), (1 << LOG2NS), INV && first, INV && last, MUL && first>(tc, tw, mulv);
    fft_rec<LC, LOG2L, NTHR, LOG2NS + LR, INV, MUL>(tc, tw, mulv);
  }
}
template <int LC, int LOG2L, int NTHR, bool INV, bool MUL>
FK_DEV void fft(cplx* tc, const cplx* __restrict__ tw, const cplx* __restrict__ mulv) {
  fft_rec<LC, LOG2L, NTHR, 0, INV, MUL>(tc, tw, mulv);
}

// ---- in-place decimation-in-frequency / decimation-in-time FFT pair ----------------
// Used for the Bluestein convolution: forward DIF (natural in, digit-reversed out),
// pointwise filter stored in the same digit-reversed order, inverse DIT (digit-
// reversed in, natural out).  Every butterfly reads and writes the same rows, so
// a thread can run its butterflies one after the other (8 values live instead of
// all of them) and there is one barrier per stage.  Stage radices: 2^(LOG2L % 3)
// first (span L), then 8 down to span 8 (DIF); mirrored for DIT.
template <int R>
FK_DEV void bfly_zero_half(cplx* v) {  // DFT-R of (v[0..R/2), 0, ..., 0)
  if constexpr (R == 2) {
    v[1] = v[0];
  } else if constexpr (R == 4) {
    const cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[2] = csub(a, b);
    v[1] = mk(a.x + b.y, a.y - b.x);
    v[3] = mk(a.x - b.y, a.y + b.x);
  } else {
    const double h = 0.70710678118654752440;
    // even-indexed inputs (v0, v2, 0, 0) and odd-indexed inputs (v1, v3, 0, 0)
    const cplx e0 = cadd(v[0], v[2]), e2 = csub(v[0], v[2]);
    const cplx e1 = mk(v[0].x + v[2].y, v[0].y - v[2].x), e3 = mk(v[0].x - v[2].y, v[0].y + v[2].x);
    const cplx o0 = cadd(v[1], v[3]), o2 = csub(v[1], v[3]);
    const cplx o1 = mk(v[1].x + v[3].y, v[1].y - v[3].x), o3 = mk(v[1].x - v[3].y, v[1].y + v[3].x);
    const cplx b1 = mk(h * (o1.x + o1.y), h * (o1.y - o1.x));
    const cplx b2 = mk(o2.y, -o2.x);
    const cplx b3 = mk(h * (o3.y - o3.x), -h * (o3.x + o3.y));
    v[0] = cadd(e0, o0);
    v[4] = csub(e0, o0);
    v[1] = cadd(e1, b1);
    v[5] = csub(e1, b1);
    v[2] = cadd(e2, b2);
    v[6] = csub(e2, b2);
    v[3] = cadd(e3, b3);
    v[7] = csub(e3, b3);
  }
}

// ---- group barriers ----------------------------------------------------------------------------------------
// Every stage of span S <= 512 rows of the in-place DIF / DIT pair touches only rows of one 512-row sub-block, and with
// the butterfly mapping b = tid + k NTHR (2 complex lanes per row, radix 8) all butterflies of the sub-blocks
// {tid / 128 + k NTHR / 128} belong to the 128 threads tid / 128.  Those stages therefore only need a barrier over
// that 4-warp group (named barrier 1 + tid / 128): the groups drift apart, so the shared-memory phase of one
// group overlaps the FP64 phase of another instead of the whole block alternating between the two pipes.
// Development switch (RUSTPDE_B200_KFLAGS, read once per process by the launchers): bit 0 = block-wide barriers
// everywhere (the round-1 behaviour), for A/B measurements.
#ifndef RP_EMU
static __constant__ int fk_kflags_c = 0;
#endif
template <int LC, int NTHR, int LOG2S>
FK_DEV void stage_sync() {
#ifndef RP_EMU
  if constexpr (LC == 2 && NTHR > 128 && NTHR % 128 == 0 && LOG2S <= 9) {
    if (!(fk_kflags_c & 1)) {
      asm volatile("bar.sync %0, 128;" ::"r"(1 + (int)(threadIdx.x >> 7)) : "memory");
      return;
    }
  }
#endif
  __syncthreads();
}

// DIF stage of span S = 1 << LOG2S, radix R.  ZERO_HALF: rows >= L/2 hold zeros (not read).
template <int LC, int LOG2L, int NTHR, int LOG2S, int R, bool ZERO_HALF>
FK_DEV void dif_stage(cplx* tc, const cplx* __restrict__ tw) {
  constexpr int L = 1 << LOG2L, S = 1 << LOG2S, SR = S / R;
  constexpr int TOT = (L / R) * LC;
#pragma unroll 2
  for (int b = threadIdx.x; b < TOT; b += NTHR) {
    const int u = b / LC, c = b % LC;
    const int q = u & (SR - 1), base = (u - q) * R + q;
    cplx v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = (ZERO_HALF && r >= R / 2) ? mk(0.0, 0.0) : tc[cidx<LC>(base + r * SR, c)];
    if (ZERO_HALF)
      bfly_zero_half<R>(v);
    else
      Bfly<R>::run(v);
    if (SR > 1) apply_twiddles<R>(v, __ldg(&tw[S / 2 - 1 + q]));
#pragma unroll
    for (int r = 0; r < R; ++r) tc[cidx<LC>(base + r * SR, c)] = v[r];
  }
  stage_sync<LC, NTHR, LOG2S>();  // the next stage (span S / R) stays inside this stage's blocks
}
// inverse DIT stage (unnormalised) of span S >= 64 (the span-8 stage is fused into dif_dit_mid, which also
// conjugates the input); LAST conjugates the results; HALF_OUT: only rows < L/2 are stored.
template <int LC, int LOG2L, int NTHR, int LOG2S, int R, bool LAST, bool HALF_OUT>
FK_DEV void dit_stage(cplx* tc, const cplx* __restrict__ tw) {
  constexpr int L = 1 << LOG2L, S = 1 << LOG2S, SR = S / R;
  constexpr int TOT = (L / R) * LC;
#pragma unroll 2
  for (int b = threadIdx.x; b < TOT; b += NTHR) {
    const int u = b / LC, c = b % LC;
    const int q = u & (SR - 1), base = (u - q) * R + q;
    cplx v[R];
#pragma unroll
    for (int r = 0; r < R; ++r) v[r] = tc[cidx<LC>(base + r * SR, c)];
    if (SR > 1) apply_twiddles<R>(v, __ldg(&tw[S / 2 - 1 + q]));
    Bfly<R>::run(v);
#pragma unroll
    for (int r = 0; r < (HALF_OUT ? R / 2 : R); ++r) {
      cplx x = v[r];
      if (LAST) x.y = -x.y;
      tc[cidx<LC>(base + r * SR, c)] = x;
    }
  }
  // the next DIT stage has span 8 S (or this was the last one): group-level only if that still fits a sub-block
  stage_sync<LC, NTHR, (LAST ? 31 : LOG2S + 3)>();
}
template <int LC, int LOG2L, int NTHR, int LOG2S, bool HALF_OUT>
FK_DEV void fft_dit_rec(cplx* tc, const cplx* __restrict__ tw) {
  static_assert(LOG2S >= 6, "the span-8 stage belongs to dif_dit_mid");
  constexpr int LR = (LOG2S == LOG2L && (LOG2L % 3)) ? (LOG2L % 3) : 3;
  constexpr bool last = (LOG2S == LOG2L);
  dit_stage<LC, LOG2L, NTHR, LOG2S, (1 << LR), last, HALF_OUT && last>(tc, tw);
  if constexpr (!last) {
    constexpr int NEXT = (LOG2S + 3 > LOG2L) ? LOG2L : LOG2S + 3;
    fft_dit_rec<LC, LOG2L, NTHR, NEXT, HALF_OUT>(tc, tw);
  }
}
// Last DIF stage (span 8), pointwise filter, first inverse DIT stage (span 8) in one pass: both stages
// work on the same 8 consecutive rows without twiddles, so the values stay in registers (one barrier and
// one shared-memory round trip less per convolution; same arithmetic as the separate stages).
template <int LC, int LOG2L, int NTHR>
FK_DEV void dif_dit_mid(cplx* tc, const cplx* __restrict__ mulv) {
  constexpr int L = 1 << LOG2L;
  constexpr int TOT = (L / 8) * LC;
#pragma unroll 1
  for (int b = threadIdx.x; b < TOT; b += NTHR) {
    const int u = b / LC, c = b % LC;
    cplx v[8], m[8];
#pragma unroll
    for (int r = 0; r < 8; ++r) m[r] = __ldg(&mulv[r * (L / 8) + u]);  // filter table stored [r][u]
#pragma unroll
    for (int r = 0; r < 8; ++r) v[r] = tc[cidx<LC>(u * 8 + r, c)];
    Bfly<8>::run(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) {
      v[r] = cmul(v[r], m[r]);
      v[r].y = -v[r].y;
    }
    Bfly<8>::run(v);
#pragma unroll
    for (int r = 0; r < 8; ++r) tc[cidx<LC>(u * 8 + r, c)] = v[r];
  }
  stage_sync<LC, NTHR, 6>();  // next: the DIT stage of span 64
}
// DIF stages down to span 64 (the span-8 stage is fused into dif_dit_mid)
template <int LC, int LOG2L, int NTHR, int LOG2S, bool ZERO_HALF>
FK_DEV void fft_dif_rec_hi(cplx* tc, const cplx* __restrict__ tw) {
  if constexpr (LOG2S > 3) {
    constexpr int LR = (LOG2S % 3) ? (LOG2S % 3) : 3;
    dif_stage<LC, LOG2L, NTHR, LOG2S, (1 << LR), ZERO_HALF && LOG2S == LOG2L>(tc, tw);
    fft_dif_rec_hi<LC, LOG2L, NTHR, LOG2S - LR, ZERO_HALF>(tc, tw);
  }
}
// circular convolution with the filter whose digit-reversed spectrum is mulv: rows >= L/2 of the input are
// zero, only rows < L/2 of the result are produced (the Bluestein use)
template <int LC, int LOG2L, int NTHR>
FK_DEV void fft_convolve_half(cplx* tc, const cplx* __restrict__ tw, const cplx* __restrict__ mulv) {
  static_assert(LOG2L >= 6, "fused middle stage needs at least two stages");
  fft_dif_rec_hi<LC, LOG2L, NTHR, LOG2L, true>(tc, tw);
  dif_dit_mid<LC, LOG2L, NTHR>(tc, mulv);
  fft_dit_rec<LC, LOG2L, NTHR, 6, true>(tc, tw);
}
// pre-combine of one pair (j, N-j): Chebyshev scaling of the backward transform
// (ortho.rs:398-404), reduction of the length-2N even DFT to a length-N real DFT.
template <bool BWD>
FK_DEV void precombine_pair(cplx a, cplx c, int j, int N, cplx sc, cplx& za, cplx& zb, cplx& f1) {
  const int jm = N - j;
  if (BWD) {  // c_k * (-1)^k / 2, ends doubled
    double ga = (j & 1) ? -0.5 : 0.5, gc = (jm & 1) ? -0.5 : 0.5;
    if (j == 0) ga *= 2.0;
    if (jm == N) gc *= 2.0;
    a = cscale(a, ga);
    c = cscale(c, gc);
  }
  const cplx sum = cadd(a, c), dif = csub(a, c);
  za = mk(fma(-sc.x, dif.x, 0.5 * sum.x), fma(-sc.x, dif.y, 0.5 * sum.y));
  zb = mk(fma(sc.x, dif.x, 0.5 * sum.x), fma(sc.x, dif.y, 0.5 * sum.y));
  const double w = (j == 0) ? 0.5 * sc.y : sc.y;
  f1.x = fma(w, dif.x, f1.x);
  f1.y = fma(w, dif.y, f1.y);
}

// per-complex-lane block reduction of f1 (threads with equal tid % LC share a lane)
template <int LC, int NTHR>
FK_DEV void reduce_f1(cplx f1, cplx* f1red) {
#pragma unroll
  for (int o = LC; o < 32; o <<= 1) {
    f1.x += __shfl_xor_sync(0xffffffffu, f1.x, o);
    f1.y += __shfl_xor_sync(0xffffffffu, f1.y, o);
  }
  const int tid = threadIdx.x;
  if ((tid & 31) < LC) f1red[(tid >> 5) * LC + (tid % LC)] = f1;
}

// recombine of one pair: X_{2k} = Z_k + Z_{N-k}, D_k = i (Z_k - Z_{N-k}); forward
// scaling (-1)^k/(n-1), ends halved (ortho.rs:355-359)
template <bool BWD>
FK_DEV void recombine_pair(cplx zk, cplx zm, int k, int N, cplx& xe, cplx& dk) {
  double h = BWD ? 1.0 : 1.0 / (double)N;
  if (!BWD && (k == 0 || 2 * k == N)) h *= 0.5;
  xe = cscale(cadd(zk, zm), h);
  dk = mk(-(zk.y - zm.y), zk.x - zm.x);
}

// Odd outputs of the DCT: prefix sum along rows N, N-1, ..., N-Ko of the tile; forward transform: times -1/N,
// last one halved when N is odd.  A thread sums both real lanes of a complex lane (16-byte accesses); carries:
// warp-level inclusive scan, then the warp totals through shared memory
template <int LC, int NTHR, int CLR, bool BWD>
FK_DEV void dct_odd_scan_v(cplx* tc, int N, double* red_) {
  constexpr int NSC = scan_threads(NTHR), NG = NSC / LC;
  static_assert(NSC % 32 == 0 && 32 % LC == 0, "scan threads");
  cplx* red = (cplx*)red_;
  const int tid = threadIdx.x, c = tid % LC, g = tid / LC;
  const bool act = tid < NSC;
  const int M = act ? (N - 1) / 2 + 1 : 0;  // Ko + 1
  const int cl = ((N - 1) / 2 + 1 + NG - 1) / NG;
  const int t0 = g * cl, t1 = min(t0 + cl, M);
  cplx q[CLR];
  cplx s = mk(0.0, 0.0);
#pragma unroll
  for (int u = 0; u < CLR; ++u) {
    const int t = t0 + u;
    if (t < t1) {
      q[u] = tc[cidx<LC>(N - t, c)];
      s = cadd(s, q[u]);
    }
  }
  // inclusive scan of the chunk sums over the threads of the warp that share the lane
  cplx inc = s;
#pragma unroll
  for (int o = LC; o < 32; o <<= 1) {
    const double ux = __shfl_up_sync(0xffffffffu, inc.x, o), uy = __shfl_up_sync(0xffffffffu, inc.y, o);
    if ((tid & 31) >= o) inc = mk(inc.x + ux, inc.y + uy);
  }
  if (act && (tid & 31) >= 32 - LC) red[(tid >> 5) * LC + c] = inc;  // warp totals
  __syncthreads();
  cplx y;  // exclusive prefix inside the warp
  {
    const double ux = __shfl_up_sync(0xffffffffu, inc.x, LC), uy = __shfl_up_sync(0xffffffffu, inc.y, LC);
    y = ((tid & 31) >= LC) ? mk(ux, uy) : mk(0.0, 0.0);
  }
  for (int w = 0; act && w < (tid >> 5); ++w) y = cadd(y, red[w * LC + c]);
  const double ho = BWD ? 1.0 : -1.0 / (double)N;
#pragma unroll
  for (int u = 0; u < CLR; ++u) {
    const int t = t0 + u;
    if (t < t1) {
      y = cadd(y, q[u]);
      cplx o = cscale(y, ho);
      if (!BWD && 2 * t + 1 == N) o = cscale(o, 0.5);
      tc[cidx<LC>(N - t, c)] = o;
    }
  }
  __syncthreads();
}

// DCT-I with the Chebyshev scaling, power-of-two N = 1 << LOG2L, in place on a
// tile of n = N + 1 rows in natural layout.  Result in split(N) layout.
template <int LC, int LOG2L, int NTHR, bool BWD>
FK_DEV void dct_pow2(double* td, const DctTab& T, double* red) {
  constexpr int N = 1 << LOG2L;
  constexpr int NP = N / 2 + 1;
  cplx* tc = (cplx*)td;
  cplx* f1red = (cplx*)red;
  const int tid = threadIdx.x, c = tid % LC;
  cplx f1 = mk(0.0, 0.0);
  for (int it = tid; it < NP * LC; it += NTHR) {
    const int j = it / LC, jm = N - j;
    const cplx a = tc[cidx<LC>(j, c)], cc = tc[cidx<LC>(jm, c)];
    cplx za, zb;
    precombine_pair<BWD>(a, cc, j, N, __ldg(&T.sc[j]), za, zb, f1);
    tc[cidx<LC>(j, c)] = za;
    if (jm != j && j != 0) tc[cidx<LC>(jm, c)] = zb;
  }
  reduce_f1<LC, NTHR>(f1, f1red);
  __syncthreads();
  fft<LC, LOG2L, NTHR, false, false>(tc, T.tw, nullptr);
  constexpr int Ko = (N - 1) / 2;
  for (int it = tid; it < NP * LC; it += NTHR) {
    const int k = it / LC;
    const cplx zk = tc[cidx<LC>(k, c)], zm = tc[cidx<LC>(k == 0 ? 0 : N - k, c)];
    cplx xe, dk;
    recombine_pair<BWD>(zk, zm, k, N, xe, dk);
    tc[cidx<LC>(k, c)] = xe;
    if (k >= 1 && k <= Ko) tc[cidx<LC>(N - k, c)] = dk;
    if (k == 0) {
      cplx s = mk(0.0, 0.0);
      for (int w = 0; w < NTHR / 32; ++w) s = cadd(s, f1red[w * LC + c]);
      tc[cidx<LC>(N, c)] = cscale(s, 2.0);
    }
  }
  __syncthreads();
  dct_odd_scan_v<LC, NTHR, (N / 2 + scan_threads(NTHR) / LC - 1) / (scan_threads(NTHR) / LC), BWD>(tc, N, red);
}

// DCT-I of arbitrary N = n - 1 through Bluestein (FFT length Lb = 1 << LOG2LB >=
// 2N - 1): reads tile A (natural layout, n rows), result in tile W (Lb rows),
// split(N) layout.  A is left untouched.
// `after_pre()` runs (all threads) once tile A has been consumed, i.e. under the FFT stages: the caller may start
// asynchronous copies into A there.
template <int LC, int LOG2LB, int NTHR, bool BWD, class Hook>
FK_DEV void dct_bluestein(const double* ta, double* tw_, const DctTab& T, double* red, Hook after_pre) {
  constexpr int LB = 1 << LOG2LB;
  const int N = T.n - 1;
  const int NP = N / 2 + 1;
  const cplx* A = (const cplx*)ta;
  cplx* W = (cplx*)tw_;
  cplx* f1red = (cplx*)red;
  const int tid = threadIdx.x, c = tid % LC;
  cplx f1 = mk(0.0, 0.0);
  for (int it = tid; it < NP * LC; it += NTHR) {
    const int j = it / LC, jm = N - j;
    cplx za, zb;
    precombine_pair<BWD>(A[cidx<LC>(j, c)], A[cidx<LC>(jm, c)], j, N, __ldg(&T.sc[j]), za, zb, f1);
    W[cidx<LC>(j, c)] = cmul(za, __ldg(&T.chirp[j]));
    if (jm != j && j != 0) W[cidx<LC>(jm, c)] = cmul(zb, __ldg(&T.chirp[jm]));
  }
  for (int it = LC * N + tid; it < (LB / 2) * LC; it += NTHR) W[cidx<LC>(it / LC, c)] = mk(0.0, 0.0);  // rows [N, LB/2)
  reduce_f1<LC, NTHR>(f1, f1red);
  __syncthreads();
  after_pre();
  fft_convolve_half<LC, LOG2LB, NTHR>(W, T.tw, T.bhat);
  const int Ko = (N - 1) / 2;
  for (int it = tid; it < NP * LC; it += NTHR) {
    const int k = it / LC;
    const cplx zk = cmul(W[cidx<LC>(k, c)], __ldg(&T.chirp[k]));
    const cplx zm = (k == 0) ? zk : cmul(W[cidx<LC>(N - k, c)], __ldg(&T.chirp[N - k]));
    cplx xe, dk;
    recombine_pair<BWD>(zk, zm, k, N, xe, dk);
    W[cidx<LC>(k, c)] = xe;
    if (k >= 1 && k <= Ko) W[cidx<LC>(N - k, c)] = dk;
    if (k == 0) {
      cplx s = mk(0.0, 0.0);
      for (int w = 0; w < NTHR / 32; ++w) s = cadd(s, f1red[w * LC + c]);
      W[cidx<LC>(N, c)] = cscale(s, 2.0);
    }
  }
  __syncthreads();
  dct_odd_scan_v<LC, NTHR, (LB / 4 + scan_threads(NTHR) / LC - 1) / (scan_threads(NTHR) / LC), BWD>(W, N, red);
}
template <int LC, int LOG2LB, int NTHR, bool BWD>
FK_DEV void dct_bluestein(const double* ta, double* tw_, const DctTab& T, double* red) {
  dct_bluestein<LC, LOG2LB, NTHR, BWD>(ta, tw_, T, red, [] {});
}

// ---- chunk-major ("permuted") coefficient tables -------------------------------------------------
// In a chunked scan the threads of a warp work on elements that are a whole chunk apart, so a load of
// coefficient[i] touches one 128-byte line per chunk group -- the tables, not the data, then dominate the L1
// traffic of the banded solves.  A permuted table stores the coefficients of scan element (group g, step u,
// parity p) at slot (u * NG + g) * 2 + p: the slots a warp reads in one step are contiguous.  Functors of the
// scans may take the slot as a third argument (host side: fk::perm_table, fast.h).
template <class F, class... A>
struct fk_takes {
  template <class G>
  static auto test(int) -> decltype(std::declval<G>()(std::declval<A>()...), char());
  template <class G>
  static long test(...);
  static constexpr bool value = sizeof(test<F>(0)) == 1;
};
template <class F>
FK_DEV auto fcall(F& f, int i, int l, int slot) {
  if constexpr (fk_takes<F, int, int, int>::value)
    return f(i, l, slot);
  else
    return f(i, l);
}

// ---- chunked scans over the 8 parity chains of a tile ---------------------------
// Dependency order t = 0..M-1 of the chain of parity p; element m = FWD ? t : M-1-t,
// natural index i = 2m + p.
//   y_t = in(i, lane) + c1(i, lane) * y_{t-1}
// `out(i, lane, y)` is called after every input of the block has been read, so
// it may overwrite the tile the inputs came from (any row).
// The inputs and coefficients of a chunk are fetched by branch-free loops (clamped indices) before the serial
// chain starts: all loads of a (sub-)batch are in flight together, instead of one exposed L2 round trip per chunk
// element (the coefficient tables do not stay in the small L1 left beside the tiles).
#define FK_SCAN_SB 8  // sub-batch of the loops that re-fetch coefficients
template <int LC, int NTHR, int CL, bool FWD, int NSC, class In, class C1, class Out>
FK_DEV void scan1n(int n, double* red, In in, C1 c1, Out out) {
  constexpr int LR = 2 * LC, NCH = 2 * LR;  // chains per tile: LR lanes x 2 parities
  constexpr int NG = NSC / NCH;
  const int tid = threadIdx.x, ch = tid % NCH, g = tid / NCH;
  const int lane = ch % LR, p = ch / LR;
  const int M = (tid < NSC) ? ((n - p + 1) >> 1) : 0;  // threads beyond NSC own empty chunks
  const int cl = (((n + 1) >> 1) + NG - 1) / NG;
  const int t0 = g * cl, t1 = min(t0 + cl, M);
  auto idx = [&](int t) { return 2 * (FWD ? t : M - 1 - t) + p; };
  constexpr int SB = FK_SCAN_SB;
  double q[CL];
  double A = 1.0, b = 0.0;
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      double k[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          q[u0 + v] = fcall(in, i, lane, sl);
          k[v] = fcall(c1, i, lane, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          b = fma(k[v], b, q[u0 + v]);
          A *= k[v];
        }
    }
  }
  // two-level carry: maps of groups of 8 chunks, then the chunks inside the group
  double* red2 = red + NG * NCH * 2;
  const bool act = tid < NSC;
  if (act) {
    red[(g * NCH + ch) * 2] = A;
    red[(g * NCH + ch) * 2 + 1] = b;
  }
  __syncthreads();
  if (act && (g & 7) == 0) {
    double Ag = 1.0, bg = 0.0;
#pragma unroll
    for (int kk = 0; kk < 8; ++kk)
      if (g + kk < NG) {
        const double Ak = red[((g + kk) * NCH + ch) * 2], bk = red[((g + kk) * NCH + ch) * 2 + 1];
        bg = fma(Ak, bg, bk);
        Ag *= Ak;
      }
    red2[((g >> 3) * NCH + ch) * 2] = Ag;
    red2[((g >> 3) * NCH + ch) * 2 + 1] = bg;
  }
  __syncthreads();
  double y = 0.0;
  if (act) {
    for (int gg = 0; gg < (g >> 3); ++gg) y = fma(red2[(gg * NCH + ch) * 2], y, red2[(gg * NCH + ch) * 2 + 1]);
    for (int gg = (g & ~7); gg < g; ++gg) y = fma(red[(gg * NCH + ch) * 2], y, red[(gg * NCH + ch) * 2 + 1]);
  }
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      double k[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) k[v] = fcall(c1, idx(min(t0 + u0 + v, M - 1)), lane, ((u0 + v) * NG + g) * 2 + p);
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          y = fma(k[v], y, q[u0 + v]);
          out(idx(t0 + u0 + v), lane, y);
        }
    }
  }
  __syncthreads();
}

template <int LC, int NTHR, int CL, bool FWD, class In, class C1, class Out>
FK_DEV void scan1(int n, double* red, In in, C1 c1, Out out) {
  scan1n<LC, NTHR, CL, FWD, scan_threads(NTHR)>(n, red, in, c1, out);
}

//   y_t = in(i, lane) + c1(i, lane) * y_{t-1} + c2(i, lane) * y_{t-2}
// in(i, lane) is evaluated again in the second walk (no register copy of the chunk): out(i, lane, .) of a
// thread may only overwrite what in(i, lane) of the same thread read -- true for every use (element-wise in place).
template <int LC, int NTHR, int CL, bool FWD, int NSC, class In, class C1, class C2, class Out>
FK_DEV void scan2n(int n, double* red, In in, C1 c1, C2 c2, Out out) {
  constexpr int LR = 2 * LC, NCH = 2 * LR;  // chains per tile: LR lanes x 2 parities
  constexpr int NG = NSC / NCH;
  constexpr int SB = FK_SCAN_SB;
  const int tid = threadIdx.x, ch = tid % NCH, g = tid / NCH;
  const int lane = ch % LR, p = ch / LR;
  const int M = (tid < NSC) ? ((n - p + 1) >> 1) : 0;
  const int cl = (((n + 1) >> 1) + NG - 1) / NG;
  const int t0 = g * cl, t1 = min(t0 + cl, M);
  auto idx = [&](int t) { return 2 * (FWD ? t : M - 1 - t) + p; };
  // particular solution and the two homogeneous ones, states (y_{t-1}, y_{t-2})
  double p1 = 0.0, p2 = 0.0, a1 = 1.0, a2 = 0.0, b1 = 0.0, b2 = 1.0;
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      double q[SB], k1[SB], k2[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          q[v] = fcall(in, i, lane, sl);
          k1[v] = fcall(c1, i, lane, sl);
          k2[v] = fcall(c2, i, lane, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          const double pn = fma(k1[v], p1, fma(k2[v], p2, q[v]));
          const double an = fma(k1[v], a1, k2[v] * a2);
          const double bn = fma(k1[v], b1, k2[v] * b2);
          p2 = p1, p1 = pn;
          a2 = a1, a1 = an;
          b2 = b1, b1 = bn;
        }
    }
  }
  const bool act = tid < NSC;
  if (act) {
    double* r = red + (g * NCH + ch) * 6;
    r[0] = a1, r[1] = b1, r[2] = p1, r[3] = a2, r[4] = b2, r[5] = p2;
  }
  __syncthreads();
  // two-level carry: maps of groups of 8 chunks, then the chunks inside the group
  double* red2 = red + NG * NCH * 6;
  if (act && (g & 7) == 0) {
    // group map  [y1; y2] -> G [y1; y2] + h,  G = [g11 g12; g21 g22]
    double g11 = 1.0, g12 = 0.0, g21 = 0.0, g22 = 1.0, h1 = 0.0, h2 = 0.0;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (g + k < NG) {
        const double* rr = red + ((g + k) * NCH + ch) * 6;
        const double n11 = fma(rr[0], g11, rr[1] * g21), n12 = fma(rr[0], g12, rr[1] * g22);
        const double n21 = fma(rr[3], g11, rr[4] * g21), n22 = fma(rr[3], g12, rr[4] * g22);
        const double m1 = fma(rr[0], h1, fma(rr[1], h2, rr[2])), m2 = fma(rr[3], h1, fma(rr[4], h2, rr[5]));
        g11 = n11, g12 = n12, g21 = n21, g22 = n22, h1 = m1, h2 = m2;
      }
    double* w = red2 + ((g >> 3) * NCH + ch) * 6;
    w[0] = g11, w[1] = g12, w[2] = h1, w[3] = g21, w[4] = g22, w[5] = h2;
  }
  __syncthreads();
  double y1 = 0.0, y2 = 0.0;
  for (int gg = 0; act && gg < (g >> 3); ++gg) {
    const double* rr = red2 + (gg * NCH + ch) * 6;
    const double n1 = fma(rr[0], y1, fma(rr[1], y2, rr[2]));
    const double n2 = fma(rr[3], y1, fma(rr[4], y2, rr[5]));
    y1 = n1, y2 = n2;
  }
  for (int gg = (g & ~7); act && gg < g; ++gg) {
    const double* rr = red + (gg * NCH + ch) * 6;
    const double n1 = fma(rr[0], y1, fma(rr[1], y2, rr[2]));
    const double n2 = fma(rr[3], y1, fma(rr[4], y2, rr[5]));
    y1 = n1, y2 = n2;
  }
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      double q[SB], k1[SB], k2[SB];  // in() again: valid because out(i, .) only overwrites what in(i, .) read
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          q[v] = fcall(in, i, lane, sl);
          k1[v] = fcall(c1, i, lane, sl);
          k2[v] = fcall(c2, i, lane, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          const double yn = fma(k1[v], y1, fma(k2[v], y2, q[v]));
          y2 = y1, y1 = yn;
          out(idx(t0 + u0 + v), lane, yn);
        }
    }
  }
  __syncthreads();
}

template <int LC, int NTHR, int CL, bool FWD, class In, class C1, class C2, class Out>
FK_DEV void scan2(int n, double* red, In in, C1 c1, C2 c2, Out out) {
  scan2n<LC, NTHR, CL, FWD, scan_threads(NTHR)>(n, red, in, c1, c2, out);
}

// chunk length bound for a lane of n elements handled by NTHR threads
RP_HD constexpr int chunk_len(int n, int nthr, int lc = 2) {
  return (((n + 1) / 2) + scan_threads(nthr) / (4 * lc) - 1) / (scan_threads(nthr) / (4 * lc));
}

// Chebyshev derivative (ortho.rs:107-125) of the lane held in tile `src`
// (layout sn_s, n elements), times `sc`, written to tile `dst` (layout sn_d;
// may alias src):  b_k = sum_{p > k, p - k odd} 2 p a_p  (k >= 1), b_0 = half of that.
template <int LC, int NTHR, int CL>
FK_DEV void cheb_diff(const double* src, int sn_s, double* dst, int sn_d, int n, double sc, double* red) {
  scan1<LC, NTHR, CL, false>(
      n, red, [&](int i, int l) { return (2.0 * (double)i * sc) * src[didx<LC>(rowof(sn_s, i), l)]; },
      [](int, int) { return 1.0; },
      [&](int i, int l, double y) {
        if (i >= 1) dst[didx<LC>(rowof(sn_d, i - 1), l)] = (i == 1) ? 0.5 * y : y;
        if (i == n - 1) dst[didx<LC>(rowof(sn_d, n - 1), l)] = 0.0;
      });
}

// HholtzAdi half step along the tile axis (hholtz_adi.rs:108-129): B2 matvec (n -> m = n-2) fused into the forward
// sweep, then the backward sweep (fdma.rs:101-118).  In place on tile t (layout sn); result elements 0..m-1.
// Coefficients from the chunk-major packed tables (fast.h perm_table): pt1[slot] = {lo, di, up, fp} in the order
// of the forward sweep, pt2[slot] = {bs, bp1, bp2, -} in the order of the backward sweep
template <int LC, int NTHR, int CL>
FK_DEV void b2_fdma_perm(double* t, int sn, int n, const double* __restrict__ pt1, const double* __restrict__ pt2, double* red) {
  const int m = n - 2;
  const double2* P1 = (const double2*)pt1;
  const double2* P2 = (const double2*)pt2;
  scan1<LC, NTHR, CL, true>(
      m, red,
      [&](int i, int l, int s) {
        const double2 a = __ldg(&P1[2 * s]), b = __ldg(&P1[2 * s + 1]);  // lo, di | up, fp
        return fma(a.x, t[didx<LC>(rowof(sn, i), l)],
                   fma(a.y, t[didx<LC>(rowof(sn, i + 2), l)], (i + 4 < n) ? b.x * t[didx<LC>(rowof(sn, i + 4), l)] : 0.0));
      },
      [&](int, int, int s) { return __ldg(&P1[2 * s + 1]).y; }, [&](int i, int l, double y) { t[didx<LC>(rowof(sn, i), l)] = y; });
  scan2<LC, NTHR, CL, false>(
      m, red, [&](int i, int l, int s) { return __ldg(&P2[2 * s]).x * t[didx<LC>(rowof(sn, i), l)]; },
      [&](int, int, int s) { return __ldg(&P2[2 * s]).y; }, [&](int, int, int s) { return __ldg(&P2[2 * s + 1]).x; },
      [&](int i, int l, double y) { t[didx<LC>(rowof(sn, i), l)] = y; });
}

// from_ortho (composite_stencil.rs:250-276): c = S^T p, then the (S^T S) solve.
// In place on tile t (layout sn): n ortho coefficients -> m = n-2 composite ones.
template <int LC, int NTHR, int CL>
FK_DEV void from_ortho(double* t, int sn, int n, const TdmaTabs& T, double* red) {
  const int m = n - 2;
  const double2* PF = (const double2*)T.pf;  // slot: sd, sl | fs, fp   (chunk-major packed, fast.h)
  const double* PB = T.pb;
  scan1<LC, NTHR, CL, true>(
      m, red,
      [&](int i, int l, int s) {
        const double2 a = __ldg(&PF[2 * s]);
        const double c = fma(a.x, t[didx<LC>(rowof(sn, i), l)], a.y * t[didx<LC>(rowof(sn, i + 2), l)]);
        return __ldg(&PF[2 * s + 1]).x * c;
      },
      [&](int, int, int s) { return __ldg(&PF[2 * s + 1]).y; }, [&](int i, int l, double y) { t[didx<LC>(rowof(sn, i), l)] = y; });
  scan1<LC, NTHR, CL, false>(
      m, red, [&](int i, int l) { return t[didx<LC>(rowof(sn, i), l)]; }, [&](int, int, int s) { return __ldg(&PB[s]); },
      [&](int i, int l, double y) { t[didx<LC>(rowof(sn, i), l)] = y; });
}

// =====================================================================================
// Two-lane ("v") forms of the building blocks: a thread works on one packed complex
// lane = two real lanes at a time (16-byte shared-memory and global accesses, one
// coefficient load per pair of lanes).  Same arithmetic per lane as the scalar forms.
// =====================================================================================
FK_DEV cplx lmul(cplx a, cplx b) { return mk(a.x * b.x, a.y * b.y); }  // lane-wise product
FK_DEV cplx lfma(cplx a, cplx b, cplx c) { return mk(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }
FK_DEV cplx sfma(double s, cplx b, cplx c) { return mk(fma(s, b.x, c.x), fma(s, b.y, c.y)); }
// recurrence coefficients are either shared by the two lanes (double) or per lane (cplx)
FK_DEV cplx kfma(double k, cplx y, cplx q) { return sfma(k, y, q); }
FK_DEV cplx kfma(cplx k, cplx y, cplx q) { return lfma(k, y, q); }
FK_DEV cplx kmul(double k, cplx a) { return cscale(a, k); }
FK_DEV cplx kmul(cplx k, cplx a) { return lmul(k, a); }

// chunk length bound of a two-lane scan run by `nsc` threads
RP_HD constexpr int chunk_len_v(int n, int nsc, int lc = 2) { return (((n + 1) / 2) + nsc / (2 * lc) - 1) / (nsc / (2 * lc)); }
// bytes of scratch of scan1v / scan2v with nsc threads (two carry levels)
RP_HD constexpr int scan1v_bytes(int nsc) { return nsc * 32 + (nsc / 8 + 8) * 32; }
RP_HD constexpr int scan2v_bytes(int nsc) { return nsc * 96 + (nsc / 8 + 8) * 96; }

//   y_t = in(i, c) + c1(i, c) * y_{t-1}   on both real lanes of complex lane c (see scan1)
template <int LC, int NTHR, int NSC, int CL, bool FWD, class In, class C1, class Out>
FK_DEV void scan1v(int n, double* red_, In in, C1 c1, Out out) {
  constexpr int NCH = 2 * LC;  // chains per tile: LC complex lanes x 2 parities
  constexpr int NG = NSC / NCH;
  static_assert(NSC <= NTHR && NSC % NCH == 0, "scan threads");
  cplx* red = (cplx*)red_;
  const int tid = threadIdx.x, ch = tid % NCH, g = tid / NCH;
  const int c = ch % LC, p = ch / LC;
  const bool act = tid < NSC;
  const int M = act ? ((n - p + 1) >> 1) : 0;  // threads beyond NSC own empty chunks
  const int cl = (((n + 1) >> 1) + NG - 1) / NG;
  const int t0 = g * cl, t1 = min(t0 + cl, M);
  auto idx = [&](int t) { return 2 * (FWD ? t : M - 1 - t) + p; };
  typedef decltype(fcall(c1, 0, 0, 0)) K;
  constexpr int SB = FK_SCAN_SB;
  cplx q[CL];
  cplx A = mk(1.0, 1.0), b = mk(0.0, 0.0);
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {  // branch-free fetch of a sub-batch (see scan1), then its part of the chain
      K k[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          q[u0 + v] = fcall(in, i, c, sl);
          k[v] = fcall(c1, i, c, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          b = kfma(k[v], b, q[u0 + v]);
          A = kmul(k[v], A);
        }
    }
  }
  cplx* red2 = red + NG * NCH * 2;
  if (act) {
    red[(g * NCH + ch) * 2] = A;
    red[(g * NCH + ch) * 2 + 1] = b;
  }
  __syncthreads();
  if (act && (g & 7) == 0) {
    cplx Ag = mk(1.0, 1.0), bg = mk(0.0, 0.0);
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (g + k < NG) {
        const cplx Ak = red[((g + k) * NCH + ch) * 2], bk = red[((g + k) * NCH + ch) * 2 + 1];
        bg = lfma(Ak, bg, bk);
        Ag = lmul(Ag, Ak);
      }
    red2[((g >> 3) * NCH + ch) * 2] = Ag;
    red2[((g >> 3) * NCH + ch) * 2 + 1] = bg;
  }
  __syncthreads();
  cplx y = mk(0.0, 0.0);
  if (act) {
    for (int gg = 0; gg < (g >> 3); ++gg) y = lfma(red2[(gg * NCH + ch) * 2], y, red2[(gg * NCH + ch) * 2 + 1]);
    for (int gg = (g & ~7); gg < g; ++gg) y = lfma(red[(gg * NCH + ch) * 2], y, red[(gg * NCH + ch) * 2 + 1]);
  }
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      K k[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) k[v] = fcall(c1, idx(min(t0 + u0 + v, M - 1)), c, ((u0 + v) * NG + g) * 2 + p);
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          y = kfma(k[v], y, q[u0 + v]);
          out(idx(t0 + u0 + v), c, y);
        }
    }
  }
  __syncthreads();
}

//   y_t = in(i, c) + c1(i, c) * y_{t-1} + c2(i, c) * y_{t-2}   (see scan2)
// CACHE = false: in(i, c) is evaluated again in the second walk instead of being kept in registers --
// only valid when out(i, c, .) of a thread touches nothing but what in(i, c) of the same thread read.
template <int LC, int NTHR, int NSC, int CL, bool FWD, bool CACHE, class In, class C1, class C2, class Out>
FK_DEV void scan2v(int n, double* red_, In in, C1 c1, C2 c2, Out out) {
  constexpr int NCH = 2 * LC;
  constexpr int NG = NSC / NCH;
  static_assert(NSC <= NTHR && NSC % NCH == 0, "scan threads");
  cplx* red = (cplx*)red_;
  const int tid = threadIdx.x, ch = tid % NCH, g = tid / NCH;
  const int c = ch % LC, p = ch / LC;
  const bool act = tid < NSC;
  const int M = act ? ((n - p + 1) >> 1) : 0;
  const int cl = (((n + 1) >> 1) + NG - 1) / NG;
  const int t0 = g * cl, t1 = min(t0 + cl, M);
  auto idx = [&](int t) { return 2 * (FWD ? t : M - 1 - t) + p; };
  typedef decltype(fcall(c1, 0, 0, 0)) K1;
  typedef decltype(fcall(c2, 0, 0, 0)) K2;
  constexpr int SB = FK_SCAN_SB;
  cplx q[CACHE ? CL : 1];
  const cplx one = mk(1.0, 1.0), zero = mk(0.0, 0.0);
  // particular solution and the two homogeneous ones, states (y_{t-1}, y_{t-2})
  cplx p1 = zero, p2 = zero, a1 = one, a2 = zero, b1 = zero, b2 = one;
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {  // branch-free fetch of a sub-batch, then its part of the chain
      cplx qq[SB];
      K1 k1[SB];
      K2 k2[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          qq[v] = fcall(in, i, c, sl);
          if (CACHE) q[u0 + v] = qq[v];
          k1[v] = fcall(c1, i, c, sl);
          k2[v] = fcall(c2, i, c, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          const cplx pn = kfma(k1[v], p1, kfma(k2[v], p2, qq[v]));
          const cplx an = kfma(k1[v], a1, kmul(k2[v], a2));
          const cplx bn = kfma(k1[v], b1, kmul(k2[v], b2));
          p2 = p1, p1 = pn;
          a2 = a1, a1 = an;
          b2 = b1, b1 = bn;
        }
    }
  }
  if (act) {
    cplx* r = red + (g * NCH + ch) * 6;
    r[0] = a1, r[1] = b1, r[2] = p1, r[3] = a2, r[4] = b2, r[5] = p2;
  }
  __syncthreads();
  cplx* red2 = red + NG * NCH * 6;
  if (act && (g & 7) == 0) {
    cplx g11 = one, g12 = zero, g21 = zero, g22 = one, h1 = zero, h2 = zero;
#pragma unroll
    for (int k = 0; k < 8; ++k)
      if (g + k < NG) {
        const cplx* rr = red + ((g + k) * NCH + ch) * 6;
        const cplx n11 = lfma(rr[0], g11, lmul(rr[1], g21)), n12 = lfma(rr[0], g12, lmul(rr[1], g22));
        const cplx n21 = lfma(rr[3], g11, lmul(rr[4], g21)), n22 = lfma(rr[3], g12, lmul(rr[4], g22));
        const cplx m1 = lfma(rr[0], h1, lfma(rr[1], h2, rr[2])), m2 = lfma(rr[3], h1, lfma(rr[4], h2, rr[5]));
        g11 = n11, g12 = n12, g21 = n21, g22 = n22, h1 = m1, h2 = m2;
      }
    cplx* w = red2 + ((g >> 3) * NCH + ch) * 6;
    w[0] = g11, w[1] = g12, w[2] = h1, w[3] = g21, w[4] = g22, w[5] = h2;
  }
  __syncthreads();
  cplx y1 = zero, y2 = zero;
  for (int gg = 0; act && gg < (g >> 3); ++gg) {
    const cplx* rr = red2 + (gg * NCH + ch) * 6;
    const cplx n1 = lfma(rr[0], y1, lfma(rr[1], y2, rr[2]));
    const cplx n2 = lfma(rr[3], y1, lfma(rr[4], y2, rr[5]));
    y1 = n1, y2 = n2;
  }
  for (int gg = (g & ~7); act && gg < g; ++gg) {
    const cplx* rr = red + (gg * NCH + ch) * 6;
    const cplx n1 = lfma(rr[0], y1, lfma(rr[1], y2, rr[2]));
    const cplx n2 = lfma(rr[3], y1, lfma(rr[4], y2, rr[5]));
    y1 = n1, y2 = n2;
  }
  if (M > 0) {
#pragma unroll
    for (int u0 = 0; u0 < CL; u0 += SB) {
      cplx qq[SB];
      K1 k1[SB];
      K2 k2[SB];
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL) {
          const int i = idx(min(t0 + u0 + v, M - 1));
          const int sl = ((u0 + v) * NG + g) * 2 + p;
          // (elements past the end of the chunk belong to the next group, which is rewriting them in place during this
          // walk: their values are never used, so they are not read either -- compute-sanitizer racecheck)
          if (CACHE)
            qq[v] = q[u0 + v];
          else if (t0 + u0 + v < t1)
            qq[v] = fcall(in, i, c, sl);
          else
            qq[v] = zero;
          k1[v] = fcall(c1, i, c, sl);
          k2[v] = fcall(c2, i, c, sl);
        }
#pragma unroll
      for (int v = 0; v < SB; ++v)
        if (u0 + v < CL && t0 + u0 + v < t1) {
          const cplx yn = kfma(k1[v], y1, kfma(k2[v], y2, qq[v]));
          y2 = y1, y1 = yn;
          out(idx(t0 + u0 + v), c, yn);
        }
    }
  }
  __syncthreads();
}

// Chebyshev derivative (see cheb_diff), tiles viewed as cplx[rows][LC]
template <int LC, int NTHR, int NMAX>
FK_DEV void cheb_diff_v(const cplx* src, int sn_s, cplx* dst, int sn_d, int n, double sc, double* red) {
  constexpr int NSC = scan_threads(NTHR);
  scan1v<LC, NTHR, NSC, chunk_len_v(NMAX, NSC, LC), false>(
      n, red, [&](int i, int c) { return cscale(src[cidx<LC>(rowof(sn_s, i), c)], 2.0 * (double)i * sc); },
      [](int, int) { return 1.0; },
      [&](int i, int c, cplx y) {
        if (i >= 1) dst[cidx<LC>(rowof(sn_d, i - 1), c)] = (i == 1) ? cscale(y, 0.5) : y;
        if (i == n - 1) dst[cidx<LC>(rowof(sn_d, n - 1), c)] = mk(0.0, 0.0);
      });
}

// HholtzAdi half step (see b2_fdma).  NSC2 threads run the second-order sweep; its scratch `red2`
// needs scan2v_bytes(NSC2) bytes, `red` scan1v_bytes(scan_threads(NTHR)).
template <int LC, int NTHR, int NMAX, int NSC2>
FK_DEV void b2_fdma_v(cplx* t, int sn, int n, const double* __restrict__ pt1, const double* __restrict__ pt2, double* red,
                      double* red2) {
  const int m = n - 2;
  constexpr int NSC = scan_threads(NTHR);
  const double2* P1 = (const double2*)pt1;  // slot: lo, di | up, fp   (chunk-major packed, fast.h)
  const double2* P2 = (const double2*)pt2;  // slot: bs, bp1 | bp2, -
  scan1v<LC, NTHR, NSC, chunk_len_v(NMAX, NSC, LC), true>(
      m, red,
      [&](int i, int c, int s) {
        const double2 a = __ldg(&P1[2 * s]), b = __ldg(&P1[2 * s + 1]);
        const cplx up = (i + 4 < n) ? cscale(t[cidx<LC>(rowof(sn, i + 4), c)], b.x) : mk(0.0, 0.0);
        return sfma(a.x, t[cidx<LC>(rowof(sn, i), c)], sfma(a.y, t[cidx<LC>(rowof(sn, i + 2), c)], up));
      },
      [&](int, int, int s) { return __ldg(&P1[2 * s + 1]).y; }, [&](int i, int c, cplx y) { t[cidx<LC>(rowof(sn, i), c)] = y; });
  scan2v<LC, NTHR, NSC2, chunk_len_v(NMAX, NSC2, LC), false, false>(
      m, red2, [&](int i, int c, int s) { return cscale(t[cidx<LC>(rowof(sn, i), c)], __ldg(&P2[2 * s]).x); },
      [&](int, int, int s) { return __ldg(&P2[2 * s]).y; }, [&](int, int, int s) { return __ldg(&P2[2 * s + 1]).x; },
      [&](int i, int c, cplx y) { t[cidx<LC>(rowof(sn, i), c)] = y; });
}

// from_ortho (see above), two-lane form
template <int LC, int NTHR, int NMAX>
FK_DEV void from_ortho_v(cplx* t, int sn, int n, const TdmaTabs& T, double* red) {
  const int m = n - 2;
  constexpr int NSC = scan_threads(NTHR);
  constexpr int CL = chunk_len_v(NMAX, NSC, LC);
  const double2* PF = (const double2*)T.pf;  // slot: sd, sl | fs, fp
  const double* PB = T.pb;
  scan1v<LC, NTHR, NSC, CL, true>(
      m, red,
      [&](int i, int c, int s) {
        const double2 a = __ldg(&PF[2 * s]);
        const cplx v = sfma(a.x, t[cidx<LC>(rowof(sn, i), c)], cscale(t[cidx<LC>(rowof(sn, i + 2), c)], a.y));
        return cscale(v, __ldg(&PF[2 * s + 1]).x);
      },
      [&](int, int, int s) { return __ldg(&PF[2 * s + 1]).y; }, [&](int i, int c, cplx y) { t[cidx<LC>(rowof(sn, i), c)] = y; });
  scan1v<LC, NTHR, NSC, CL, false>(
      m, red, [&](int i, int c) { return t[cidx<LC>(rowof(sn, i), c)]; }, [&](int, int, int s) { return __ldg(&PB[s]); },
      [&](int i, int c, cplx y) { t[cidx<LC>(rowof(sn, i), c)] = y; });
}

// ---- 16-byte global accesses of a pitched real matrix (two adjacent columns = one complex lane) ----------
// (a[i][col], a[i][col + 1]), zero outside the matrix; col even, a.p 16-byte aligned, a.ld even
FK_DEV cplx ld2(const Mat& a, int i, int col) {
  const int cc = (col < a.cols) ? col : 0;  // keeps the access inside the row pitch
  const cplx v = *(const cplx*)(a.p + (size_t)min(i, a.rows - 1) * a.ld + cc);
  const bool r = i < a.rows;
  return mk((r && col < a.cols) ? v.x : 0.0, (r && col + 1 < a.cols) ? v.y : 0.0);
}
FK_DEV void st2(const Mat& o, int i, int col, cplx v) {
  double* q = o.p + (size_t)i * o.ld + col;
  if (col + 1 < o.cols)
    *(cplx*)q = v;
  else if (col < o.cols)
    *q = v.x;
}
// composite -> ortho stencil across columns applied while loading: out_j = d_j a[i][j] + l_{j-2} a[i][j-2],
// j = col, col + 1; (yd, yl) are the coefficient pairs prepared by stencil_pair() (zero where out of range)
struct StencilPair {
  cplx d, l;
  int c0, c2;
};
FK_DEV StencilPair stencil_pair(int cols, int col, const double* __restrict__ sd, const double* __restrict__ sl) {
  StencilPair s;
  s.d = mk(col < cols ? __ldg(&sd[col]) : 0.0, col + 1 < cols ? __ldg(&sd[col + 1]) : 0.0);
  s.l = mk((col >= 2 && col - 2 < cols) ? __ldg(&sl[col - 2]) : 0.0, (col >= 2 && col - 1 < cols) ? __ldg(&sl[col - 1]) : 0.0);
  s.c0 = col < cols ? col : 0;
  s.c2 = (col >= 2 && col - 2 < cols) ? col - 2 : 0;
  return s;
}
FK_DEV cplx ld2_stencil(const Mat& f, int i, const StencilPair& s) {  // i < f.rows
  const double* row = f.p + (size_t)i * f.ld;
  const cplx v0 = *(const cplx*)(row + s.c0), v2 = *(const cplx*)(row + s.c2);
  // selects, not products with zero: the padding columns of the pitch are never initialised
  const double a0 = s.d.x != 0.0 ? v0.x : 0.0, a1 = s.d.y != 0.0 ? v0.y : 0.0;
  const double b0 = s.l.x != 0.0 ? v2.x : 0.0, b1 = s.l.y != 0.0 ? v2.y : 0.0;
  return mk(fma(s.l.x, b0, s.d.x * a0), fma(s.l.y, b1, s.d.y * a1));
}

}  // namespace fk
}  // namespace rp
