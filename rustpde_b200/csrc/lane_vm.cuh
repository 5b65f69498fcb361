// lane_vm.cuh -- device side of the lane programs (see lane_prog.h).
//
// One thread block owns T slots of packed lanes in shared memory and runs a
// short program of block-parallel stages over them: strided/contiguous loads
// and stores, Galerkin stencils, Chebyshev derivative, DCT-I / real FFT,
// banded matvec and banded solves.  All sequential recurrences of the
// reference (ortho.rs:107-125, linalg.rs:14-57, fdma.rs:101-118) run as
// block-wide scans: every thread walks a short chunk, the chunk summaries are
// combined with warp shuffles, and the chunk is re-walked with its carry.
//
// Performance rules followed throughout (B200: FP64 and HBM are roughly
// balanced for these transforms, so latency hiding is what matters):
//   * shared-memory pointers are derived from the smem symbol in the function
//     that uses them (LDS/STS, no generic or local traffic);
//   * loops are written in explicit batches (load B values, then compute,
//     then store) so several independent memory operations are in flight;
//   * no integer division in inner loops (all strides are powers of two).
#pragma once
#include "lane_prog.h"

namespace rp {

typedef double2 cplx;
#define RP_DEV __device__ __forceinline__
#ifdef RP_INLINE_OPS
#define RP_DEVNI __device__ __forceinline__
#else
#define RP_DEVNI __device__ __noinline__
#endif
#define RP_DEVCALL __device__ __noinline__

RP_DEV cplx mk(double x, double y) { return make_double2(x, y); }
RP_DEV cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
RP_DEV cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
RP_DEV cplx cmul(cplx a, cplx b) { return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
RP_DEV cplx csqr(cplx a) { return mk(fma(a.x, a.x, -a.y * a.y), 2.0 * a.x * a.y); }
RP_DEV cplx cscale(cplx a, double s) { return mk(a.x * s, a.y * s); }
RP_DEV cplx cconj(cplx a) { return mk(a.x, -a.y); }
// componentwise (the two packed real lanes are independent)
RP_DEV cplx pmul(cplx a, cplx b) { return mk(a.x * b.x, a.y * b.y); }
RP_DEV cplx pfma(cplx a, cplx b, cplx c) { return mk(fma(a.x, b.x, c.x), fma(a.y, b.y, c.y)); }

RP_DEV int padi(int s) { return s + (s >> 5); }
RP_DEV int slot_of(const Lay& L, int i) {
  int h = i >> 1;
  return (i & 1) ? L.o0 + L.so * h : L.e0 + L.se * h;
}
RP_DEV int ilog2(int v) { return 31 - __clz(v); }

// Dynamic shared memory base (see header comment).
#ifdef RP_EMU
#define RP_SMEM ((cplx*)cuemu::dyn_smem())
#else
extern __shared__ __align__(16) unsigned char rp_dyn_smem_raw_[];
#define RP_SMEM ((cplx*)rp_dyn_smem_raw_)
#endif

// development aid: sub-op cycle marks of thread 0 / block 0 (RUSTPDE_B200_OPPROF=1)
#ifndef RP_EMU
#define RP_MARK_INIT long long mark_t_ = clock64()
#define RP_MARK(slot)                                                   \
  do {                                                                  \
    if (blockIdx.x == 0 && threadIdx.x == 0 && pg->prof) {              \
      const long long c_ = clock64();                                   \
      pg->prof[slot] += c_ - mark_t_;                                   \
      mark_t_ = c_;                                                     \
    }                                                                   \
  } while (0)
#else
#define RP_MARK_INIT
#define RP_MARK(slot)
#endif

// Block context.  Built by value inside each (noinline) op from the program
// header; never passed by reference across a call, so it stays in registers.
struct Blk {
  cplx* regs;
  cplx* wb;
  cplx* scr;
  int wb_off;  // offset of wb from RP_SMEM in cplx units
  int T, logT, capP, wbP, wbT;
  int tid, nthr;
  int unit0, nunits, axis;
};
RP_DEV cplx* lane_ptr(const Blk& b, int r, int t) { return b.regs + (r * b.T + t) * b.capP; }
RP_DEV Blk make_blk(const Program* __restrict__ pg) {
  Blk b;
  b.T = pg->T;
  b.logT = ilog2(b.T);
  b.capP = padi(pg->cap) + 1;
  const int wbcap = pg->wb_cap;
  b.wbP = wbcap ? padi(wbcap) + 1 : 0;
  b.wbT = pg->wb_T;
  b.regs = RP_SMEM;
  b.wb_off = pg->nreg * b.T * b.capP;
  b.wb = RP_SMEM + b.wb_off;
  b.scr = b.wb + b.wbT * b.wbP;
  b.tid = threadIdx.x;
  b.nthr = blockDim.x;
  b.unit0 = blockIdx.x * b.T;
  b.nunits = pg->nunits;
  b.axis = pg->axis;
  return b;
}

// ===========================================================================
// Radix-R DFT in registers (forward, e^{-2 pi i / R}); in-order output.
// ===========================================================================
template <int R>
struct Dft;
template <>
struct Dft<2> {
  static RP_DEV void run(cplx* v) {
    cplx a = v[0], b = v[1];
    v[0] = cadd(a, b);
    v[1] = csub(a, b);
  }
};
template <>
struct Dft<4> {
  static RP_DEV void run(cplx* v) {
    cplx a0 = cadd(v[0], v[2]), a1 = csub(v[0], v[2]);
    cplx a2 = cadd(v[1], v[3]), a3 = csub(v[1], v[3]);
    cplx m3 = mk(a3.y, -a3.x);  // -i * a3
    v[0] = cadd(a0, a2);
    v[2] = csub(a0, a2);
    v[1] = cadd(a1, m3);
    v[3] = csub(a1, m3);
  }
};
template <int R>
RP_DEV cplx dft_tw(int k);  // exp(-2 pi i k / R), k < R/2
template <>
RP_DEV cplx dft_tw<8>(int k) {
  const double h = 0.70710678118654752440;
  switch (k) {
    case 0: return mk(1.0, 0.0);
    case 1: return mk(h, -h);
    case 2: return mk(0.0, -1.0);
    default: return mk(-h, -h);
  }
}
template <>
RP_DEV cplx dft_tw<16>(int k) {
  const double h = 0.70710678118654752440, c = 0.92387953251128675613, s = 0.38268343236508977173;
  switch (k) {
    case 0: return mk(1.0, 0.0);
    case 1: return mk(c, -s);
    case 2: return mk(h, -h);
    case 3: return mk(s, -c);
    case 4: return mk(0.0, -1.0);
    case 5: return mk(-s, -c);
    case 6: return mk(-h, -h);
    default: return mk(-c, -s);
  }
}
template <int R>
struct Dft {
  static RP_DEV void run(cplx* v) {
    cplx e[R / 2], o[R / 2];
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      e[k] = v[2 * k];
      o[k] = v[2 * k + 1];
    }
    Dft<R / 2>::run(e);
    Dft<R / 2>::run(o);
#pragma unroll
    for (int k = 0; k < R / 2; ++k) {
      cplx t = cmul(o[k], dft_tw<R>(k));
      v[k] = cadd(e[k], t);
      v[k + R / 2] = csub(e[k], t);
    }
  }
};

// v[r] *= w^r, r = 1..R-1, from the single table value w (log-depth products,
// error ~ 4 ulp), applied as soon as each power is known to bound live registers.
template <int R>
RP_DEV void apply_twiddles(cplx* v, cplx w1) {
  v[1] = cmul(v[1], w1);
  if constexpr (R > 2) {
    const cplx w2 = csqr(w1);
    v[2] = cmul(v[2], w2);
    const cplx w3 = cmul(w2, w1);
    v[3] = cmul(v[3], w3);
    if constexpr (R > 4) {
      const cplx w4 = csqr(w2);
      v[4] = cmul(v[4], w4);
      const cplx w5 = cmul(w4, w1);
      v[5] = cmul(v[5], w5);
      const cplx w6 = cmul(w4, w2);
      v[6] = cmul(v[6], w6);
      const cplx w7 = cmul(w4, w3);
      v[7] = cmul(v[7], w7);
      if constexpr (R > 8) {
        const cplx w8 = csqr(w4);
        v[8] = cmul(v[8], w8);
        v[9] = cmul(v[9], cmul(w8, w1));
        v[10] = cmul(v[10], cmul(w8, w2));
        v[11] = cmul(v[11], cmul(w8, w3));
        v[12] = cmul(v[12], cmul(w8, w4));
        v[13] = cmul(v[13], cmul(w8, w5));
        v[14] = cmul(v[14], cmul(w8, w6));
        v[15] = cmul(v[15], cmul(w8, w7));
      }
    }
  }
}

// One Stockham radix-R pass, in place through registers.  Each thread owns K
// butterflies of one FFT; `per` = (L/R)/K threads serve one FFT.  Arguments
// are scalars (offsets, not pointers).  All of L, R, K, Ns, nthr are powers of 2.
enum { FF_CONJ_IN = 1, FF_CONJ_OUT = 2, FF_LIN = 4 };
template <int R, int K>
RP_DEVCALL void fft_pass(int tid, int nthr, int base_off, int stride, int nslots, int L, int Ns,
                       const cplx* __restrict__ tw, int fl, double oscale, Lay lin) {
  const int nb = L / R;
  const int per = nb / K, lper = ilog2(per);
  const int fpr = nthr >> lper;  // FFTs per round
  const int twstep = L / (Ns * R);
  const bool conj_in = fl & FF_CONJ_IN, conj_out = fl & FF_CONJ_OUT, use_lin = fl & FF_LIN;
  cplx* const base = RP_SMEM + base_off;
  const int q = tid & (per - 1), tq = tid >> lper;
  for (int s0 = 0; s0 < nslots; s0 += fpr) {
    const int t = s0 + tq;
    const bool act = (t < nslots);
    cplx v[K][R];
    if (act) {
      const cplx* x = base + t * stride;
      cplx w1[K];
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const int j = q + per * kk;
        w1[kk] = (Ns > 1) ? __ldg(&tw[(j & (Ns - 1)) * twstep]) : mk(1.0, 0.0);
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int src = j + r * nb;
          cplx c = x[padi(use_lin ? slot_of(lin, src) : src)];
          if (conj_in) c.y = -c.y;
          v[kk][r] = c;
        }
      }
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        if (Ns > 1) apply_twiddles<R>(v[kk], w1[kk]);
        Dft<R>::run(v[kk]);
      }
    }
    __syncthreads();
    if (act) {
      cplx* x = base + t * stride;
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const int j = q + per * kk;
        const int k = j & (Ns - 1);
        const int o = (j - k) * R + k;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          cplx c = v[kk][r];
          if (conj_out) c.y = -c.y;
          c.x *= oscale;
          c.y *= oscale;
          x[padi(o + r * Ns)] = c;
        }
      }
    }
    __syncthreads();
  }
}

template <int R>
RP_DEV void fft_pass_k(int tid, int nthr, int K, int base_off, int stride, int nslots, int L, int Ns, const cplx* tw,
                       int fl, double os, Lay lin) {
  if (K == 1) {
    fft_pass<R, 1>(tid, nthr, base_off, stride, nslots, L, Ns, tw, fl, os, lin);
    return;
  }
  if constexpr (R <= 8) {
    if (K == 2) {
      fft_pass<R, 2>(tid, nthr, base_off, stride, nslots, L, Ns, tw, fl, os, lin);
      return;
    }
  }
  if constexpr (R <= 4) {
    if (K == 4) {
      fft_pass<R, 4>(tid, nthr, base_off, stride, nslots, L, Ns, tw, fl, os, lin);
      return;
    }
  }
  if constexpr (R <= 2) {
    if (K == 8) {
      fft_pass<R, 8>(tid, nthr, base_off, stride, nslots, L, Ns, tw, fl, os, lin);
      return;
    }
  }
}

// Complex FFT of pow2 length L on nslots buffers (RP_SMEM + base_off + t*stride),
// natural order in and out.  inverse: conj-in / conj-out around the forward kernel.
// Requires nthr >= L/16 (host planner guarantees it).  use_lin: the first pass
// reads element j from slot_of(lin, j) (absorbs an input permutation for free).
RP_DEVCALL void fft_run(int tid, int nthr, int base_off, int stride, int nslots, int L, const cplx* tw, bool inverse,
                      double scale, Lay lin, bool use_lin) {
  int rem = L, Ns = 1;
  while (rem > 1) {
    const int R = rem >= 8 ? 8 : rem;  // radix 8: L/8 butterflies keep every thread busy at nthr = L/8
    const bool first = (Ns == 1), last = (rem == R);
    int K = (L / R) / nthr;  // butterflies per thread; R*K = L/nthr <= 16 when K > 1
    if (K < 1) K = 1;
    const int fl = ((inverse && first) ? FF_CONJ_IN : 0) | ((inverse && last) ? FF_CONJ_OUT : 0) |
                   ((use_lin && first) ? FF_LIN : 0);
    const double os = last ? scale : 1.0;
    switch (R) {
      case 16: fft_pass_k<16>(tid, nthr, K, base_off, stride, nslots, L, Ns, tw, fl, os, lin); break;
      case 8: fft_pass_k<8>(tid, nthr, K, base_off, stride, nslots, L, Ns, tw, fl, os, lin); break;
      case 4: fft_pass_k<4>(tid, nthr, K, base_off, stride, nslots, L, Ns, tw, fl, os, lin); break;
      default: fft_pass_k<2>(tid, nthr, K, base_off, stride, nslots, L, Ns, tw, fl, os, lin); break;
    }
    Ns *= R;
    rem /= R;
  }
}
RP_DEV void fft_run(const Blk& b, int base_off, int stride, int nslots, int L, const cplx* tw, bool inverse, double scale) {
  fft_run(b.tid, b.nthr, base_off, stride, nslots, L, tw, inverse, scale, lay_natural(), false);
}

// Batched grid-stride loops are written out explicitly at each site:
//   for (i0 = start; i0 < n; i0 += 4*step) { load 4 -> registers; compute/store 4 }
// (register arrays with compile-time indices only -- lambdas capturing arrays by
// reference ended up in local memory and dominated the run time).
#define RP_B4 _Pragma("unroll") for (int u = 0; u < 4; ++u)

// Complex DFT of arbitrary length on lane slots [0, L) of register buffers
// (natural order), in place.  pow2 -> fft_run; else Bluestein through b.wb.
RP_DEV void dft_any(const Blk& b, int base_off, int stride, int nslots, const FftPlan& P, bool inverse, double scale) {
  if (P.pow2) {
    fft_run(b, base_off, stride, nslots, P.L, P.tw, inverse, scale);
    return;
  }
  const int L = P.L, Lb = P.Lb;
  cplx* const base = RP_SMEM + base_off;
  const cplx* __restrict__ chirp = P.chirp;
  const cplx* __restrict__ bhat = P.bhat;
  for (int g = 0; g < nslots; g += b.wbT) {
    const int ns = min(b.wbT, nslots - g);
    for (int t = 0; t < ns; ++t) {
      const cplx* x = base + (g + t) * stride;
      cplx* w = b.wb + t * b.wbP;
      for (int j0 = b.tid; j0 < Lb; j0 += 4 * b.nthr) {
        cplx v[4], c[4];
        RP_B4 {
          const int j = j0 + u * b.nthr;
          v[u] = mk(0.0, 0.0);
          c[u] = mk(0.0, 0.0);
          if (j < L) {
            v[u] = x[padi(j)];
            c[u] = __ldg(&chirp[j]);
          }
        }
        RP_B4 {
          const int j = j0 + u * b.nthr;
          if (j < Lb) {
            cplx a = v[u];
            if (inverse) a.y = -a.y;
            w[padi(j)] = cmul(a, c[u]);
          }
        }
      }
    }
    __syncthreads();
    fft_run(b, b.wb_off, b.wbP, ns, Lb, P.tw, false, 1.0);
    for (int t = 0; t < ns; ++t) {
      cplx* w = b.wb + t * b.wbP;
      for (int j0 = b.tid; j0 < Lb; j0 += 4 * b.nthr) {
        cplx v[4], c[4];
        RP_B4 {
          const int j = j0 + u * b.nthr;
          if (j < Lb) {
            v[u] = w[padi(j)];
            c[u] = __ldg(&bhat[j]);
          }
        }
        RP_B4 {
          const int j = j0 + u * b.nthr;
          if (j < Lb) w[padi(j)] = cmul(v[u], c[u]);
        }
      }
    }
    __syncthreads();
    fft_run(b, b.wb_off, b.wbP, ns, Lb, P.tw, true, 1.0);
    for (int t = 0; t < ns; ++t) {
      cplx* x = base + (g + t) * stride;
      const cplx* w = b.wb + t * b.wbP;
      for (int j0 = b.tid; j0 < L; j0 += 4 * b.nthr) {
        cplx v[4], c[4];
        RP_B4 {
          const int j = j0 + u * b.nthr;
          if (j < L) {
            v[u] = w[padi(j)];
            c[u] = __ldg(&chirp[j]);
          }
        }
        RP_B4 {
          const int j = j0 + u * b.nthr;
          if (j < L) {
            cplx a = cmul(v[u], c[u]);
            if (inverse) a.y = -a.y;
            x[padi(j)] = cscale(a, scale);
          }
        }
      }
    }
    __syncthreads();
  }
}

// ===========================================================================
// Linear recurrences along parity chains as block-wide scans.
//   y_k = s*q_k + p*y_{k-1} [+ r*y_{k-2}]      (k in dependency order)
// Chain c of slot t: element k at slot c0[c] + cs[c]*m, m = fwd ? k : M-1-k.
// F(t, c, m) -> coefficients (componentwise double2).
// tpc = nthr / (T*nch) threads serve one chain (a multiple of 32); each walks
// a chunk of ceil(M/tpc) elements, the chunk maps  y -> A y + b  are combined
// by an inclusive warp scan (shuffles) and a short cross-warp fix-up through
// shared memory, and the chunk is re-walked with its carry-in.
// ===========================================================================
// coefficient type CT: double (same coefficient for both packed components)
// or cplx (componentwise, e.g. two real rows with different eigenvalues)
RP_DEV cplx kmulv(double k, cplx v) { return mk(k * v.x, k * v.y); }
RP_DEV cplx kmulv(cplx k, cplx v) { return pmul(k, v); }
RP_DEV cplx kfmav(double k, cplx v, cplx c) { return mk(fma(k, v.x, c.x), fma(k, v.y, c.y)); }
RP_DEV cplx kfmav(cplx k, cplx v, cplx c) { return pfma(k, v, c); }
RP_DEV double kmul(double a, double b) { return a * b; }
RP_DEV cplx kmul(cplx a, cplx b) { return pmul(a, b); }
RP_DEV double kfma(double a, double b, double c) { return fma(a, b, c); }
RP_DEV cplx kfma(cplx a, cplx b, cplx c) { return pfma(a, b, c); }
template <class CT>
RP_DEV CT kconst(double v);
template <>
RP_DEV double kconst<double>(double v) {
  return v;
}
template <>
RP_DEV cplx kconst<cplx>(double v) {
  return mk(v, v);
}
RP_DEV double shfl_up_k(double v, int d) { return __shfl_up_sync(0xffffffffu, v, d); }
RP_DEV cplx shfl_up_k(cplx v, int d) {
  return mk(__shfl_up_sync(0xffffffffu, v.x, d), __shfl_up_sync(0xffffffffu, v.y, d));
}

template <class CT>
struct CoefT {
  CT s, p, r;
};
struct Chains {
  int nch;  // chains per slot (1 or 2)
  int c0[2], cs[2], M[2];
};
template <class CT>
struct AffT {  // [y1; y2] -> A [y1; y2] + b
  CT a11, a12, a21, a22;
  cplx b1, b2;
};
template <class CT, bool SECOND>
RP_DEV AffT<CT> aff_combine(const AffT<CT>& p, const AffT<CT>& c) {  // apply p first, then c
  AffT<CT> r;
  if (SECOND) {
    r.a11 = kfma(c.a11, p.a11, kmul(c.a12, p.a21));
    r.a12 = kfma(c.a11, p.a12, kmul(c.a12, p.a22));
    r.a21 = kfma(c.a21, p.a11, kmul(c.a22, p.a21));
    r.a22 = kfma(c.a21, p.a12, kmul(c.a22, p.a22));
    r.b1 = kfmav(c.a11, p.b1, kfmav(c.a12, p.b2, c.b1));
    r.b2 = kfmav(c.a21, p.b1, kfmav(c.a22, p.b2, c.b2));
  } else {
    r.a11 = kmul(c.a11, p.a11);
    r.b1 = kfmav(c.a11, p.b1, c.b1);
    r.a12 = r.a21 = r.a22 = kconst<CT>(0.0);
    r.b2 = mk(0, 0);
  }
  return r;
}
template <class CT, bool SECOND>
RP_DEV AffT<CT> aff_shfl_up(const AffT<CT>& a, int d) {
  AffT<CT> r;
  r.a11 = shfl_up_k(a.a11, d);
  r.b1 = shfl_up_k(a.b1, d);
  if (SECOND) {
    r.a12 = shfl_up_k(a.a12, d);
    r.a21 = shfl_up_k(a.a21, d);
    r.a22 = shfl_up_k(a.a22, d);
    r.b2 = shfl_up_k(a.b2, d);
  } else {
    r.a12 = r.a21 = r.a22 = kconst<CT>(0.0);
    r.b2 = mk(0, 0);
  }
  return r;
}

template <class CT, bool SECOND, class F>
RP_DEV void chain_solve(const Blk& b, int reg, const Chains& ch, bool fwd, F coef) {
  const int nchains = b.T * ch.nch;
  const int ltpc = ilog2(b.nthr) - ilog2(nchains);  // threads per chain (pow2, >= 32)
  const int tpc = 1 << ltpc;
  const int cid = b.tid >> ltpc, j = b.tid & (tpc - 1);
  const int t = (ch.nch == 2) ? (cid >> 1) : cid, c = (ch.nch == 2) ? (cid & 1) : 0;
  int Mmax = ch.M[0];
  if (ch.nch > 1 && ch.M[1] > Mmax) Mmax = ch.M[1];
  const int Cs = (Mmax + tpc - 1) >> ltpc;
  // (ternaries, not ch.X[c]: a dynamically indexed struct array would live in local memory)
  const int M = c ? ch.M[1] : ch.M[0];
  const int c0 = c ? ch.c0[1] : ch.c0[0], cs = c ? ch.cs[1] : ch.cs[0];
  cplx* const x = lane_ptr(b, reg, t);
  const int k0 = min(M, j * Cs), k1 = min(M, k0 + Cs);
  const int lane = b.tid & 31, wic = j >> 5, wpc = tpc >> 5;  // warp in chain, warps per chain
  const CT one = kconst<CT>(1.0), zero = kconst<CT>(0.0);
  // ---- pass A: chunk map (zero-carry result and homogeneous responses) ----
  cplx y1 = mk(0, 0), y2 = mk(0, 0);
  CT u1 = one, u2 = zero, v1 = zero, v2 = one;
  for (int kb = k0; kb < k1; kb += 2) {
    CoefT<CT> cf[2];
    cplx q[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (kb + u < k1) {
        const int m = fwd ? kb + u : M - 1 - (kb + u);
        cf[u] = coef(t, c, m);
        q[u] = x[padi(c0 + cs * m)];
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (kb + u < k1) {
        if (SECOND) {
          const cplx y = kfmav(cf[u].p, y1, kfmav(cf[u].r, y2, kmulv(cf[u].s, q[u])));
          y2 = y1;
          y1 = y;
          const CT uu = kfma(cf[u].p, u1, kmul(cf[u].r, u2));
          u2 = u1;
          u1 = uu;
          const CT vv = kfma(cf[u].p, v1, kmul(cf[u].r, v2));
          v2 = v1;
          v1 = vv;
        } else {
          y1 = kfmav(cf[u].p, y1, kmulv(cf[u].s, q[u]));
          u1 = kmul(cf[u].p, u1);
        }
      }
  }
  AffT<CT> inc;
  inc.a11 = u1;
  inc.b1 = y1;
  if (SECOND) {
    inc.a12 = v1;
    inc.a21 = u2;
    inc.a22 = v2;
    inc.b2 = y2;
  } else {
    inc.a12 = inc.a21 = inc.a22 = zero;
    inc.b2 = mk(0, 0);
  }
  // ---- inclusive warp scan of the chunk maps --------------------------------
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const AffT<CT> prev = aff_shfl_up<CT, SECOND>(inc, d);
    if (lane >= d) inc = aff_combine<CT, SECOND>(prev, inc);
  }
  AffT<CT> exc = aff_shfl_up<CT, SECOND>(inc, 1);  // map of the chunks before mine inside the warp
  if (lane == 0) {
    exc.a11 = one;
    exc.a22 = one;
    exc.a12 = exc.a21 = zero;
    exc.b1 = exc.b2 = mk(0, 0);
  }
  cplx ci1, ci2;  // carry-in: y_{k0-1}, y_{k0-2}
  if (wpc > 1) {
    // warp totals through shared memory: 4 coefficient slots (as cplx) + 2 vectors
    cplx* tot = b.scr + (cid * wpc + wic) * 6;
    if (lane == 31) {
      tot[0] = kmulv(inc.a11, mk(1, 1));
      tot[4] = inc.b1;
      if (SECOND) {
        tot[1] = kmulv(inc.a12, mk(1, 1));
        tot[2] = kmulv(inc.a21, mk(1, 1));
        tot[3] = kmulv(inc.a22, mk(1, 1));
        tot[5] = inc.b2;
      }
    }
    __syncthreads();
    cplx p1 = mk(0, 0), p2 = mk(0, 0);  // state entering my warp
    for (int w = 0; w < wic; ++w) {
      const cplx* s = b.scr + (cid * wpc + w) * 6;
      if (SECOND) {
        const cplx n1 = pfma(s[0], p1, pfma(s[1], p2, s[4]));
        const cplx n2 = pfma(s[2], p1, pfma(s[3], p2, s[5]));
        p1 = n1;
        p2 = n2;
      } else {
        p1 = pfma(s[0], p1, s[4]);
      }
    }
    if (SECOND) {
      ci1 = kfmav(exc.a11, p1, kfmav(exc.a12, p2, exc.b1));
      ci2 = kfmav(exc.a21, p1, kfmav(exc.a22, p2, exc.b2));
    } else {
      ci1 = kfmav(exc.a11, p1, exc.b1);
      ci2 = mk(0, 0);
    }
  } else {
    ci1 = exc.b1;
    ci2 = exc.b2;
  }
  // ---- pass C: re-walk with the carry, in place ------------------------------
  y1 = ci1;
  y2 = ci2;
  for (int kb = k0; kb < k1; kb += 2) {
    CoefT<CT> cf[2];
    cplx q[2];
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (kb + u < k1) {
        const int m = fwd ? kb + u : M - 1 - (kb + u);
        cf[u] = coef(t, c, m);
        q[u] = x[padi(c0 + cs * m)];
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (kb + u < k1) {
        cplx y;
        if (SECOND) {
          y = kfmav(cf[u].p, y1, kfmav(cf[u].r, y2, kmulv(cf[u].s, q[u])));
          y2 = y1;
        } else {
          y = kfmav(cf[u].p, y1, kmulv(cf[u].s, q[u]));
        }
        y1 = y;
        q[u] = y;
      }
#pragma unroll
    for (int u = 0; u < 2; ++u)
      if (kb + u < k1) {
        const int m = fwd ? kb + u : M - 1 - (kb + u);
        x[padi(c0 + cs * m)] = q[u];
      }
  }
  __syncthreads();
}

RP_DEV Chains chains_of(const Lay& L, int n) {
  Chains ch;
  ch.nch = 2;
  ch.c0[0] = L.e0;
  ch.cs[0] = L.se;
  ch.M[0] = (n + 1) >> 1;
  ch.c0[1] = L.o0;
  ch.cs[1] = L.so;
  ch.M[1] = n >> 1;
  return ch;
}

// Stencil along the lane, elementwise and in place:
//   dir = -1: out_i = w(i, in_i, in_{i-2})                 (to_ortho)
//   dir = +1: out_i = w(i, in_i, in_{i+2}, in_{i+4})        (S^T, B2 matvec)
// Elements are processed in rounds of E per thread, walking against `dir`, so
// a round only ever reads slots that no earlier round has written; results are
// held in registers across the round's single barrier.
// nin = number of valid inputs (others read as 0), nout = outputs written.
template <class W>
RP_DEV void lane_stencil(const Blk& b, int reg, const Lay& L, int nin, int nout, int dir, W wfun) {
  enum { E = 4 };
  const int nmax = nin > nout ? nin : nout;
  const int span = E * b.nthr;
  const int nrounds = (nmax + span - 1) / span;
  for (int t = 0; t < b.T; ++t) {
    cplx* const x = lane_ptr(b, reg, t);
    auto in = [&](int i) -> cplx { return (i >= 0 && i < nin) ? x[padi(slot_of(L, i))] : mk(0, 0); };
    for (int r = 0; r < nrounds; ++r) {
      const int base = (dir > 0) ? r * span : (nrounds - 1 - r) * span;
      cplx a0[E], a1[E], a2[E];
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = base + b.tid + u * b.nthr;
        if (i < nmax) {
          a0[u] = in(i);
          a1[u] = in(i + 2 * dir);
          a2[u] = (dir > 0) ? in(i + 4) : mk(0, 0);
        }
      }
      cplx res[E];
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = base + b.tid + u * b.nthr;
        if (i < nmax) res[u] = wfun(i, a0[u], a1[u], a2[u]);
      }
      __syncthreads();
#pragma unroll
      for (int u = 0; u < E; ++u) {
        const int i = base + b.tid + u * b.nthr;
        if (i < nout) x[padi(slot_of(L, i))] = res[u];
      }
    }
  }
  __syncthreads();
}

// ===========================================================================
// Loads / stores between global arrays and lanes
// ===========================================================================
struct LaneSel {
  int la, lb;   // lane indices of the .x / .y component (complex: la == lb)
  bool va, vb;  // valid
};
RP_DEV LaneSel lane_sel(const Blk& b, const Instr& I, int t) {
  const int P = b.unit0 + t;
  LaneSel s;
  if (I.flags & (LF_COMPLEX | LF_BCAST)) {
    int l = ((I.flags & LF_LANE2) ? 2 * P + I.i2 : P) + I.shift;
    s.la = s.lb = l;
    s.va = s.vb = (P < b.nunits) && l >= 0 && l < I.nlanes;
  } else {
    s.la = 2 * P + I.shift;
    s.lb = s.la + 1;
    s.va = (P < b.nunits) && s.la >= 0 && s.la < I.nlanes;
    s.vb = (P < b.nunits) && s.lb >= 0 && s.lb < I.nlanes;
  }
  return s;
}

// Per-thread iteration space of a load/store: AXIS_Y walks one slot at a time
// with i = tid + k*nthr (coalesced along the row); AXIS_X gives every thread a
// fixed slot t = tid % T and i = tid / T + k * nthr / T (the T slots of one row
// are adjacent in memory).
RP_DEVNI void op_ld(const Program* __restrict__ pg, int pc) {
  RP_MARK_INIT;
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, nz = (I.n2 > n && !(I.flags & LF_ACC)) ? I.n2 : n;
  const bool acc = I.flags & LF_ACC, isc = I.flags & LF_COMPLEX, isb = I.flags & LF_BCAST, mik = I.flags & LF_MULIK;
  const double* __restrict__ lc = (const double*)I.p1;
  const double* __restrict__ srd = (const double*)I.p0;
  const cplx* __restrict__ src = (const cplx*)I.p0;
  const long long ld = I.ld;
  const Lay L = I.lay;
  const bool ydir = (b.axis == AXIS_Y);
  RP_MARK(40);
  const int nt = ydir ? b.T : 1;
  for (int tt = 0; tt < nt; ++tt) {
    const int t = ydir ? tt : (b.tid & (b.T - 1));
    const int istart = ydir ? b.tid : (b.tid >> b.logT);
    const int istep = ydir ? b.nthr : (b.nthr >> b.logT);
    const LaneSel s = lane_sel(b, I, t);
    double ca = I.s0, cb = I.s0;
    if (I.flags & LF_LANECOEF) {
      ca *= s.va ? __ldg(&lc[s.la]) : 0.0;
      cb *= s.vb ? __ldg(&lc[s.lb]) : 0.0;
    }
    const double kfac = (double)s.la * ca;
    // element (lane l, index i) lives at l*sl + i*si
    const long long sl = ydir ? ld : 1, si = ydir ? 1 : ld;
    const long long oa = (long long)s.la * sl, ob = (long long)s.lb * sl;
    cplx* const x = lane_ptr(b, I.r0, t);
    for (int i0 = istart; i0 < nz; i0 += 4 * istep) {
      cplx v[4], o[4];
      RP_B4 {
        const int i = i0 + u * istep;
        cplx a = mk(0, 0);
        if (i < n) {
          if (isc) {
            if (s.va) a = src[oa + i * si];
          } else if (isb) {
            if (s.va) {
              const double d = srd[oa + i * si];
              a = mk(d, d);
            }
          } else {
            if (s.va) a.x = srd[oa + i * si];
            if (s.vb) a.y = srd[ob + i * si];
          }
          if (acc) o[u] = x[padi(slot_of(L, i))];
        }
        v[u] = a;
      }
      RP_B4 {
        const int i = i0 + u * istep;
        if (i < nz) {
          cplx a = v[u];
          a = mik ? mk(-kfac * a.y, kfac * a.x) : mk(a.x * ca, a.y * cb);
          cplx* px = x + padi(slot_of(L, i));
          if (acc) {
            if (i < n) *px = cadd(o[u], a);
          } else {
            *px = a;
          }
        }
      }
    }
  }
  RP_MARK(41);
  __syncthreads();
  RP_MARK(42);
}

RP_DEVNI void op_st(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n;
  const bool acc = I.flags & LF_ACC, isc = I.flags & LF_COMPLEX, cut = I.flags & LF_CUT;
  double* __restrict__ drd = (double*)I.p0;
  cplx* __restrict__ dst = (cplx*)I.p0;
  const long long ld = I.ld;
  const Lay L = I.lay;
  const bool ydir = (b.axis == AXIS_Y);
  const int nt = ydir ? b.T : 1;
  for (int tt = 0; tt < nt; ++tt) {
    const int t = ydir ? tt : (b.tid & (b.T - 1));
    const int istart = ydir ? b.tid : (b.tid >> b.logT);
    const int istep = ydir ? b.nthr : (b.nthr >> b.logT);
    const LaneSel s = lane_sel(b, I, t);
    const long long sl = ydir ? ld : 1, si = ydir ? 1 : ld;
    const long long oa = (long long)s.la * sl, ob = (long long)s.lb * sl;
    const bool lcuta = cut && I.i1 >= 0 && s.la >= I.i1, lcutb = cut && I.i1 >= 0 && s.lb >= I.i1;
    const cplx* const x = lane_ptr(b, I.r0, t);
    const double s0 = I.s0;
    const int cut_i = I.i0;
    for (int i0 = istart; i0 < n; i0 += 4 * istep) {
      cplx v[4], old[4];
      RP_B4 {
        const int i = i0 + u * istep;
        if (i < n) {
          v[u] = x[padi(slot_of(L, i))];
          if (acc) {
            old[u] = mk(0, 0);
            if (isc) {
              if (s.va) old[u] = dst[oa + i * si];
            } else {
              if (s.va) old[u].x = drd[oa + i * si];
              if (s.vb) old[u].y = drd[ob + i * si];
            }
          }
        }
      }
      RP_B4 {
        const int i = i0 + u * istep;
        if (i < n) {
          cplx a = cscale(v[u], s0);
          const bool ecut = cut && i >= cut_i;
          if (ecut || lcuta) a.x = 0.0;
          if (ecut || lcutb) a.y = 0.0;
          if (isc && (ecut || lcuta)) a = mk(0, 0);
          if (acc) a = cadd(a, old[u]);
          if (isc) {
            if (s.va) dst[oa + i * si] = a;
          } else {
            if (s.va) drd[oa + i * si] = a.x;
            if (s.vb) drd[ob + i * si] = a.y;
          }
        }
      }
    }
  }
  __syncthreads();
}

// elementwise two/three-register ops; mode: 0 copy, 1 axpy, 2 scale, 3 mulpw, 4 mulpw-acc, 5 zero, 6 cut
RP_DEVNI void op_elem(const Program* __restrict__ pg, int pc, int mode) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n;
  const Lay L0 = I.lay, L1 = I.lay2;
  for (int t = 0; t < b.T; ++t) {
    cplx* const x0 = lane_ptr(b, I.r0, t);
    const cplx* const x1 = lane_ptr(b, I.r1, t);
    const cplx* const x2 = lane_ptr(b, I.r2, t);
    const double s0 = I.s0;
    const int cut_i = I.i0;
    for (int i0 = b.tid; i0 < n; i0 += 4 * b.nthr) {
      cplx va[4], vb[4], vc[4];
      RP_B4 {
        const int i = i0 + u * b.nthr;
        if (i < n) {
          if (mode == 1 || mode == 2 || mode == 4) va[u] = x0[padi(slot_of(L0, i))];
          if (mode == 0 || mode == 1 || mode == 3 || mode == 4) vb[u] = x1[padi(slot_of(L1, i))];
          if (mode == 3 || mode == 4) vc[u] = x2[padi(slot_of(L1, i))];
        }
      }
      RP_B4 {
        const int i = i0 + u * b.nthr;
        if (i < n && !(mode == 6 && i < cut_i)) {
          cplx r;
          switch (mode) {
            case 0: r = vb[u]; break;
            case 1: r = mk(fma(s0, vb[u].x, va[u].x), fma(s0, vb[u].y, va[u].y)); break;
            case 2: r = cscale(va[u], s0); break;
            case 3: r = pmul(vb[u], vc[u]); break;
            case 4: r = pfma(vb[u], vc[u], va[u]); break;
            default: r = mk(0, 0); break;
          }
          x0[padi(slot_of(L0, i))] = r;
        }
      }
    }
  }
  __syncthreads();
}

RP_DEVNI void op_mulik(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n;
  const Lay L = I.lay;
  const bool ek = I.flags & LF_ELEMK;
  for (int t = 0; t < b.T; ++t) {
    cplx* const x = lane_ptr(b, I.r0, t);
    const double kl = (double)(b.unit0 + t) * I.s0;
    for (int i = b.tid; i < n; i += b.nthr) {
      const double k = ek ? (double)i * I.s0 : kl;
      cplx* px = x + padi(slot_of(L, i));
      const cplx v = *px;
      *px = mk(-k * v.y, k * v.x);
    }
  }
  __syncthreads();
}

RP_DEVNI void op_setzero00(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  if (b.unit0 == 0 && b.tid == 0) {
    cplx* x0 = lane_ptr(b, I.r0, 0) + padi(slot_of(I.lay, 0));
    if (I.flags & LF_COMPLEX)
      *x0 = mk(0, 0);
    else
      x0->x = 0.0;
  }
  __syncthreads();
}

// ---- Galerkin stencils ------------------------------------------------------
// to_ortho (composite_stencil.rs:207-229): p_i = d_i c_i + l_{i-2} c_{i-2}
RP_DEVNI void op_toortho(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, m = n - 2;
  const double* __restrict__ d = (const double*)I.p0;
  const double* __restrict__ l = (const double*)I.p1;
  lane_stencil(b, I.r0, I.lay, m, n, -1, [=](int i, cplx a0, cplx a1, cplx) -> cplx {
    const double di = (i < m) ? __ldg(&d[i]) : 0.0;
    const double li = (i >= 2) ? __ldg(&l[i - 2]) : 0.0;
    return mk(fma(di, a0.x, li * a1.x), fma(di, a0.y, li * a1.y));
  });
}

// from_ortho (composite_stencil.rs:250-276): c = S^T p, then (S^T S) solve.
RP_DEVNI void op_fromortho(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, m = n - 2;
  const double* __restrict__ d = (const double*)I.p0;
  const double* __restrict__ l = (const double*)I.p1;
  const TdmaTab tt = *(const TdmaTab*)I.p2;
  lane_stencil(b, I.r0, I.lay, n, m, +1, [=](int i, cplx a0, cplx a1, cplx) -> cplx {
    if (i >= m) return mk(0, 0);
    const double di = __ldg(&d[i]), li = __ldg(&l[i]);
    return mk(fma(di, a0.x, li * a1.x), fma(di, a0.y, li * a1.y));
  });
  const Chains ch = chains_of(I.lay, m);
  const double* __restrict__ fs = tt.fs;
  const double* __restrict__ fp = tt.fp;
  const double* __restrict__ bp = tt.bp;
  chain_solve<double, false>(b, I.r0, ch, true, [=](int, int c, int mm) -> CoefT<double> {
    const int i = 2 * mm + c;
    CoefT<double> cf;
    cf.s = __ldg(&fs[i]);
    cf.p = __ldg(&fp[i]);
    cf.r = 0.0;
    return cf;
  });
  chain_solve<double, false>(b, I.r0, ch, false, [=](int, int c, int mm) -> CoefT<double> {
    const int i = 2 * mm + c;
    CoefT<double> cf;
    cf.s = 1.0;
    cf.p = __ldg(&bp[i]);
    cf.r = 0.0;
    return cf;
  });
}

// B2 preconditioner (matvec.rs:172-193): out_r = lo_r x_r + di_r x_{r+2} + up_r x_{r+4}
RP_DEVNI void op_bandmv(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, m = n - 2;
  const double* __restrict__ lo = (const double*)I.p0;
  const double* __restrict__ di = (const double*)I.p1;
  const double* __restrict__ up = (const double*)I.p2;
  lane_stencil(b, I.r0, I.lay, n, m, +1, [=](int i, cplx a0, cplx a1, cplx a2) -> cplx {
    if (i >= m) return mk(0, 0);
    const double w0 = __ldg(&lo[i]), w1 = __ldg(&di[i]), w2 = __ldg(&up[i]);
    return mk(fma(w0, a0.x, fma(w1, a1.x, w2 * a2.x)), fma(w0, a0.y, fma(w1, a1.y, w2 * a2.y)));
  });
}

// Chebyshev derivative (ortho.rs:107-125) as suffix sums by parity:
//   b_k = sum_{p>k, p-k odd} 2 p a_p  (k>=1),  b_0 = half of that.
// Run in place on the chain of p; b_{p-1} ends up in the slot of a_p, so the
// layout becomes lay_after_diff(lay).
RP_DEVNI void op_diff(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  Lay L = I.lay;
  const int n = I.n;
  for (int rep = 0; rep < I.i0; ++rep) {
    const double sc = (rep == 0) ? I.s0 : 1.0;
    const Chains ch = chains_of(L, n);
    chain_solve<double, false>(b, I.r0, ch, false, [=](int, int c, int mm) -> CoefT<double> {
      CoefT<double> cf;
      cf.s = 2.0 * (double)(2 * mm + c) * sc;
      cf.p = 1.0;
      cf.r = 0.0;
      return cf;
    });
    const Lay Ln = lay_after_diff(L);
    for (int t = b.tid; t < b.T; t += b.nthr) {
      cplx* x = lane_ptr(b, I.r0, t);
      cplx* x0 = &x[padi(slot_of(Ln, 0))];
      *x0 = cscale(*x0, 0.5);
      x[padi(slot_of(Ln, n - 1))] = mk(0, 0);
    }
    __syncthreads();
    L = Ln;
  }
}

// Pre-swept banded solve (fdma.rs:101-118)
RP_DEVNI void op_fdma(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const FdmaTab ft = *(const FdmaTab*)I.p0;
  const Chains ch = chains_of(I.lay, I.n);
  const double* __restrict__ fp = ft.fp;
  const double* __restrict__ bs = ft.bs;
  const double* __restrict__ bp1 = ft.bp1;
  const double* __restrict__ bp2 = ft.bp2;
  chain_solve<double, false>(b, I.r0, ch, true, [=](int, int c, int mm) -> CoefT<double> {
    const int i = 2 * mm + c;
    CoefT<double> cf;
    cf.s = 1.0;
    cf.p = __ldg(&fp[i]);
    cf.r = 0.0;
    return cf;
  });
  chain_solve<double, true>(b, I.r0, ch, false, [=](int, int c, int mm) -> CoefT<double> {
    const int i = 2 * mm + c;
    CoefT<double> cf;
    cf.s = __ldg(&bs[i]);
    cf.p = __ldg(&bp1[i]);
    cf.r = __ldg(&bp2[i]);
    return cf;
  });
}

// Per-lane banded solve (A + (lam+alpha) C) x = b  (fdma_tensor.rs:219-227).
// Register r1 holds the lane of swept-pivot reciprocals 1/dia'_i (set-up data),
// everything else of the sweep is recomputed from the raw diagonals.
RP_DEVNI void op_fdmamode(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const FdmaModeTab mt = *(const FdmaModeTab*)I.p0;
  const int n = I.n;
  const Chains ch = chains_of(I.lay, n);
  const Lay L = I.lay;
  const int r1 = I.r1;
  const bool cplxl = (I.flags & LF_COMPLEX) != 0;
  const int nlanes = I.nlanes;
  const double* __restrict__ a_low = mt.a_low;
  const double* __restrict__ a_up1 = mt.a_up1;
  const double* __restrict__ a_up2 = mt.a_up2;
  const double* __restrict__ c_low = mt.c_low;
  const double* __restrict__ c_up1 = mt.c_up1;
  const double* __restrict__ c_up2 = mt.c_up2;
  const double* __restrict__ lam = mt.lam;
  const double alpha = mt.alpha;
  auto mu_of = [=](int t) -> cplx {
    const int P = b.unit0 + t;
    int la = cplxl ? P : 2 * P, lb = cplxl ? P : 2 * P + 1;
    if (la >= nlanes) la = nlanes - 1;
    if (lb >= nlanes) lb = nlanes - 1;
    return mk(__ldg(&lam[la]) + alpha, __ldg(&lam[lb]) + alpha);
  };
  // forward: x_i -= l_{i-2} x_{i-2},  l_j = low_j / dia'_j
  chain_solve<cplx, false>(b, I.r0, ch, true, [=](int t, int c, int mm) -> CoefT<cplx> {
    const int i = 2 * mm + c;
    CoefT<cplx> cf;
    cf.s = mk(1, 1);
    cf.r = mk(0, 0);
    cf.p = mk(0, 0);
    if (i >= 2) {
      const cplx mu = mu_of(t);
      const double al = __ldg(&a_low[i - 2]), cl = __ldg(&c_low[i - 2]);
      const cplx inv = lane_ptr(b, r1, t)[padi(slot_of(L, i - 2))];
      cf.p = mk(-fma(mu.x, cl, al) * inv.x, -fma(mu.y, cl, al) * inv.y);
    }
    return cf;
  });
  // backward: x_i = (x_i - up1'_i x_{i+2} - up2_i x_{i+4}) / dia'_i
  chain_solve<cplx, true>(b, I.r0, ch, false, [=](int t, int c, int mm) -> CoefT<cplx> {
    const int i = 2 * mm + c;
    const cplx mu = mu_of(t);
    const cplx inv = lane_ptr(b, r1, t)[padi(slot_of(L, i))];
    cplx u1 = mk(0, 0), u2 = mk(0, 0);
    if (i < n - 2) {
      const double a = __ldg(&a_up1[i]), cc = __ldg(&c_up1[i]);
      u1 = mk(fma(mu.x, cc, a), fma(mu.y, cc, a));
      if (i >= 2) {
        const double al = __ldg(&a_low[i - 2]), cl = __ldg(&c_low[i - 2]);
        const double a2 = __ldg(&a_up2[i - 2]), c2 = __ldg(&c_up2[i - 2]);
        const cplx invm = lane_ptr(b, r1, t)[padi(slot_of(L, i - 2))];
        const cplx lw = mk(fma(mu.x, cl, al) * invm.x, fma(mu.y, cl, al) * invm.y);
        u1.x = fma(-lw.x, fma(mu.x, c2, a2), u1.x);
        u1.y = fma(-lw.y, fma(mu.y, c2, a2), u1.y);
      }
    }
    if (i < n - 4) {
      const double a = __ldg(&a_up2[i]), cc = __ldg(&c_up2[i]);
      u2 = mk(fma(mu.x, cc, a), fma(mu.y, cc, a));
    }
    CoefT<cplx> cf;
    cf.s = inv;
    cf.p = mk(-u1.x * inv.x, -u1.y * inv.y);
    cf.r = mk(-u2.x * inv.x, -u2.y * inv.y);
    return cf;
  });
}

// ===========================================================================
// DCT-I with the Chebyshev scaling (ortho.rs:337-360 forward, 383-407 backward)
// on packed lanes: pre-combine to a length-N real-DFT problem per component,
// one complex DFT of length N, recombine; odd outputs by a prefix sum.
// Output layout: SPLIT(N).
// ===========================================================================
RP_DEVNI void op_dct(const Program* __restrict__ pg, int pc) {
  RP_MARK_INIT;
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const DctPlan P = *(const DctPlan*)I.p0;
  const int n = P.n, N = n - 1;
  const bool backward = I.i0 != 0;
  const Lay lin = I.lay;
  const int npairs = N / 2 + 1;
  const int nwarp = (b.nthr + 31) >> 5;
  const bool pow2 = P.fft.pow2;
  const int group = pow2 ? b.T : b.wbT;
  const cplx* __restrict__ sctab = P.sc;
  const cplx* __restrict__ chirp = P.fft.chirp;
  const cplx* __restrict__ bhat = P.fft.bhat;
  const int tid = b.tid, nthr = b.nthr;
  for (int g = 0; g < b.T; g += group) {
    const int ns = min(group, b.T - g);
    // ---- pre-combine, in place at the source slots (the first FFT pass
    //      reads through the layout map; Bluestein writes to the work buffer)
    for (int tt = 0; tt < ns; ++tt) {
      const int t = g + tt;
      cplx* const x = lane_ptr(b, I.r0, t);
      cplx* const wbuf = b.wb + tt * b.wbP;
      cplx f1 = mk(0, 0);
      for (int j0 = tid; j0 < npairs; j0 += 4 * nthr) {
        cplx va[4], vc[4], vs[4], ca[4], cc[4];
        RP_B4 {
          const int j = j0 + u * nthr;
          if (j < npairs) {
            va[u] = x[padi(slot_of(lin, j))];
            vc[u] = x[padi(slot_of(lin, N - j))];
            vs[u] = __ldg(&sctab[j]);
            if (!pow2) {
              ca[u] = __ldg(&chirp[j]);
              cc[u] = __ldg(&chirp[(N - j) % N]);
            }
          }
        }
        RP_B4 {
          const int j = j0 + u * nthr;
          if (j < npairs) {
            const int jm = N - j;
            cplx a = va[u], c = vc[u];
            if (backward) {  // c_k * (-1)^k / 2, ends doubled
              double ga = (j & 1) ? -0.5 : 0.5, gc = (jm & 1) ? -0.5 : 0.5;
              if (j == 0) ga *= 2.0;
              if (jm == N) gc *= 2.0;
              a = cscale(a, ga);
              c = cscale(c, gc);
            }
            const cplx sc = vs[u];
            const cplx sum = cadd(a, c), dif = csub(a, c);
            const cplx za = mk(fma(-sc.x, dif.x, 0.5 * sum.x), fma(-sc.x, dif.y, 0.5 * sum.y));
            const cplx zb = mk(fma(sc.x, dif.x, 0.5 * sum.x), fma(sc.x, dif.y, 0.5 * sum.y));
            const double w = (j == 0) ? 0.5 * sc.y : sc.y;
            f1.x = fma(w, dif.x, f1.x);
            f1.y = fma(w, dif.y, f1.y);
            if (pow2) {
              x[padi(slot_of(lin, j))] = za;
              if (jm != j && j != 0) x[padi(slot_of(lin, jm))] = zb;
            } else {
              wbuf[padi(j)] = cmul(za, ca[u]);
              if (jm != j && j != 0) wbuf[padi(jm)] = cmul(zb, cc[u]);
            }
          }
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        f1.x += __shfl_down_sync(0xffffffffu, f1.x, o);
        f1.y += __shfl_down_sync(0xffffffffu, f1.y, o);
      }
      if ((tid & 31) == 0) b.scr[t * 32 + (tid >> 5)] = f1;
    }
    __syncthreads();
    RP_MARK(32);
    // ---- complex DFT of length N ----------------------------------------
    if (pow2) {
      fft_run(tid, nthr, (I.r0 * b.T + g) * b.capP, b.capP, ns, N, P.fft.tw, false, 1.0, lin, true);
    } else {
      const int Lb = P.fft.Lb;
      for (int t = 0; t < ns; ++t) {
        cplx* w = b.wb + t * b.wbP;
        for (int j = N + tid; j < Lb; j += nthr) w[padi(j)] = mk(0, 0);
      }
      __syncthreads();
      fft_run(b, b.wb_off, b.wbP, ns, Lb, P.fft.tw, false, 1.0);
      for (int t = 0; t < ns; ++t) {
        cplx* w = b.wb + t * b.wbP;
        for (int j0 = tid; j0 < Lb; j0 += 4 * nthr) {
          cplx v[4], c[4];
          RP_B4 {
            const int j = j0 + u * nthr;
            if (j < Lb) {
              v[u] = w[padi(j)];
              c[u] = __ldg(&bhat[j]);
            }
          }
          RP_B4 {
            const int j = j0 + u * nthr;
            if (j < Lb) w[padi(j)] = cmul(v[u], c[u]);
          }
        }
      }
      __syncthreads();
      fft_run(b, b.wb_off, b.wbP, ns, Lb, P.fft.tw, true, 1.0);
      for (int t = 0; t < ns; ++t) {
        const cplx* w = b.wb + t * b.wbP;
        cplx* x = lane_ptr(b, I.r0, g + t);
        for (int j0 = tid; j0 < N; j0 += 4 * nthr) {
          cplx v[4], c[4];
          RP_B4 {
            const int j = j0 + u * nthr;
            if (j < N) {
              v[u] = w[padi(j)];
              c[u] = __ldg(&chirp[j]);
            }
          }
          RP_B4 {
            const int j = j0 + u * nthr;
            if (j < N) x[padi(j)] = cmul(v[u], c[u]);
          }
        }
      }
      __syncthreads();
    }
  }
  RP_MARK(33);
  // ---- recombine: X_{2k} = Z_k + Z_{N-k};  D_k = i (Z_k - Z_{N-k}) -------
  const double he = backward ? 1.0 : 1.0 / (double)N;  // even outputs: (+1)/N
  const int Ko = (N - 1) / 2;                           // odd outputs O_0..O_Ko
  for (int t = 0; t < b.T; ++t) {
    cplx* const x = lane_ptr(b, I.r0, t);
    for (int k0 = tid; k0 < npairs; k0 += 4 * nthr) {
      cplx vk[4], vm[4];
      RP_B4 {
        const int k = k0 + u * nthr;
        if (k < npairs) {
          vk[u] = x[padi(k)];
          vm[u] = x[padi(k == 0 ? 0 : N - k)];
        }
      }
      RP_B4 {
        const int k = k0 + u * nthr;
        if (k < npairs) {
          const cplx zk = vk[u], zm = vm[u];
          double h = he;
          if (!backward && (k == 0 || 2 * k == N)) h *= 0.5;
          x[padi(k)] = cscale(cadd(zk, zm), h);
          if (k >= 1 && k <= Ko) x[padi(N - k)] = mk(-(zk.y - zm.y), zk.x - zm.x);
          if (k == 0) {
            cplx f1 = mk(0, 0);
            for (int w = 0; w < nwarp; ++w) f1 = cadd(f1, b.scr[t * 32 + w]);
            x[padi(N)] = cscale(f1, 2.0);
          }
        }
      }
    }
  }
  __syncthreads();
  RP_MARK(34);
  // ---- odd outputs: prefix sum along slots N, N-1, ... -------------------
  Chains ch;
  ch.nch = 1;
  ch.c0[0] = N;
  ch.cs[0] = -1;
  ch.M[0] = Ko + 1;
  ch.c0[1] = 0;
  ch.cs[1] = 0;
  ch.M[1] = 0;
  chain_solve<double, false>(b, I.r0, ch, true, [=](int, int, int) -> CoefT<double> {
    CoefT<double> cf;
    cf.s = 1.0;
    cf.p = 1.0;
    cf.r = 0.0;
    return cf;
  });
  RP_MARK(35);
  if (!backward) {  // odd outputs: (-1)/N, last one halved when N is odd
    const double ho = -1.0 / (double)N;
    for (int t = 0; t < b.T; ++t) {
      cplx* const x = lane_ptr(b, I.r0, t);
      for (int k = tid; k <= Ko; k += nthr) {
        cplx* px = &x[padi(N - k)];
        *px = cscale(*px, (2 * k + 1 == N) ? 0.5 * ho : ho);
      }
    }
    __syncthreads();
  }
}

// ===========================================================================
// Real FFT along x (r2c.rs:250-303) on a packed pair of real lanes (columns
// 2P, 2P+1), fused with the global-memory side so only one register is needed:
//   RFFT : r0 (natural, n reals per component) -> FFT -> X_a, X_b stored to the
//          complex array's columns 2P, 2P+1 (n/2+1 rows), scale s0, dealias cut i0
//   IRFFT: complex columns 2P, 2P+1 (optionally times i k s0) -> packed spectrum
//          -> inverse FFT -> r0.  Im of the DC and Nyquist bins is ignored (c2r).
// ===========================================================================
RP_DEVNI void op_rfft(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const FftPlan P = *(const FftPlan*)I.p1;
  const int n = P.L, m = n / 2 + 1;
  dft_any(b, I.r0 * b.T * b.capP, b.capP, b.T, P, false, 1.0);
  cplx* __restrict__ dst = (cplx*)I.p0;
  const long long ld = I.ld;
  const int t = b.tid & (b.T - 1);
  const int ca = 2 * (b.unit0 + t), cb = ca + 1;
  const bool va = ca < I.nlanes, vb = cb < I.nlanes;
  const cplx* const z = lane_ptr(b, I.r0, t);
  const bool cut = I.flags & LF_CUT;
  const int cut_k = I.i0;
  const double s0 = I.s0;
  const int kstart = b.tid >> b.logT, kstep = b.nthr >> b.logT;
  for (int k0 = kstart; k0 < m; k0 += 4 * kstep) {
    cplx vk[4], vm[4];
    RP_B4 {
      const int k = k0 + u * kstep;
      if (k < m) {
        vk[u] = z[padi(k)];
        vm[u] = z[padi(k == 0 ? 0 : n - k)];
      }
    }
    RP_B4 {
      const int k = k0 + u * kstep;
      if (k < m) {
        const cplx zk = vk[u], zm = cconj(vm[u]);
        const cplx s = cadd(zk, zm), d = csub(zk, zm);
        double h = 0.5 * s0;
        if (cut && k >= cut_k) h = 0.0;
        if (va) dst[k * ld + ca] = mk(h * s.x, h * s.y);
        if (vb) dst[k * ld + cb] = mk(h * d.y, -h * d.x);
      }
    }
  }
  __syncthreads();
}

RP_DEVNI void op_irfft(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const FftPlan P = *(const FftPlan*)I.p1;
  const int n = P.L, m = n / 2 + 1;
  const cplx* __restrict__ src = (const cplx*)I.p0;
  const long long ld = I.ld;
  const int t = b.tid & (b.T - 1);
  const int ca = 2 * (b.unit0 + t), cb = ca + 1;
  const bool va = ca < I.nlanes, vb = cb < I.nlanes;
  const bool mik = I.flags & LF_MULIK;
  const double s0 = I.s0;
  cplx* const z = lane_ptr(b, I.r0, t);
  const int kstart = b.tid >> b.logT, kstep = b.nthr >> b.logT;
  for (int k0 = kstart; k0 < m; k0 += 4 * kstep) {
    cplx xa4[4], xb4[4];
    RP_B4 {
      const int k = k0 + u * kstep;
      if (k < m) {
        xa4[u] = va ? src[k * ld + ca] : mk(0, 0);
        xb4[u] = vb ? src[k * ld + cb] : mk(0, 0);
      }
    }
    RP_B4 {
      const int k = k0 + u * kstep;
      if (k < m) {
        cplx xa = xa4[u], xb = xb4[u];
        if (mik) {
          const double kk = (double)k * s0;
          xa = mk(-kk * xa.y, kk * xa.x);
          xb = mk(-kk * xb.y, kk * xb.x);
        } else {
          xa = cscale(xa, s0);
          xb = cscale(xb, s0);
        }
        if (k == 0 || 2 * k == n) {
          xa.y = 0.0;
          xb.y = 0.0;
        }
        z[padi(k)] = mk(xa.x - xb.y, xa.y + xb.x);  // X_a + i X_b
        if (k > 0 && 2 * k != n) z[padi(n - k)] = mk(xa.x + xb.y, -xa.y + xb.x);  // conj(X_a) + i conj(X_b)
      }
    }
  }
  __syncthreads();
  dft_any(b, I.r0 * b.T * b.capP, b.capP, b.T, P, true, 1.0 / (double)n);
}

// ===========================================================================
// The interpreter
// ===========================================================================
RP_DEV void lane_vm_body(const Program* __restrict__ progs) {
  const Program* __restrict__ pg = progs + blockIdx.y;
  if ((int)(blockIdx.x * pg->T) >= pg->nunits) return;
  const int ninstr = pg->ninstr;
#ifndef RP_EMU
  long long* const prof = (blockIdx.x == 0 && threadIdx.x == 0) ? pg->prof : nullptr;
#endif
  for (int pc = 0; pc < ninstr; ++pc) {
    const int op = pg->ins[pc].op;
#ifndef RP_EMU
    long long t0 = 0;
    if (prof) t0 = clock64();
#endif
    switch (op) {
      case OP_LD: op_ld(pg, pc); break;
      case OP_ST: op_st(pg, pc); break;
      case OP_ZERO: op_elem(pg, pc, 5); break;
      case OP_COPY: op_elem(pg, pc, 0); break;
      case OP_AXPY: op_elem(pg, pc, 1); break;
      case OP_SCALE: op_elem(pg, pc, 2); break;
      case OP_MULPW: op_elem(pg, pc, (pg->ins[pc].flags & LF_ACC) ? 4 : 3); break;
      case OP_CUT: op_elem(pg, pc, 6); break;
      case OP_MULIK: op_mulik(pg, pc); break;
      case OP_TOORTHO: op_toortho(pg, pc); break;
      case OP_FROMORTHO: op_fromortho(pg, pc); break;
      case OP_DIFF: op_diff(pg, pc); break;
      case OP_DCT: op_dct(pg, pc); break;
      case OP_RFFT: op_rfft(pg, pc); break;
      case OP_IRFFT: op_irfft(pg, pc); break;
      case OP_BANDMV: op_bandmv(pg, pc); break;
      case OP_FDMA: op_fdma(pg, pc); break;
      case OP_FDMAMODE: op_fdmamode(pg, pc); break;
      case OP_SETZERO00: op_setzero00(pg, pc); break;
      default: break;
    }
#ifndef RP_EMU
    if (prof) prof[op] += clock64() - t0;
#endif
  }
}

}  // namespace rp
