// lane_vm.cuh -- device side of the lane programs (see lane_prog.h).
//
// One thread block owns T slots of packed lanes in shared memory and runs a
// short program of block-parallel stages over them: strided/contiguous loads
// and stores, Galerkin stencils, Chebyshev derivative, DCT-I / real FFT,
// banded matvec and banded solves.  All sequential recurrences of the
// reference (ortho.rs:107-125, linalg.rs:14-57, fdma.rs:101-118) are run as
// chunked two-pass recurrences so that the whole block works on them.
#pragma once
#include "lane_prog.h"

namespace rp {

typedef double2 cplx;
#define RP_DEV __device__ __forceinline__
#define RP_DEVNI __device__ __noinline__

RP_DEV cplx mk(double x, double y) { return make_double2(x, y); }
RP_DEV cplx cadd(cplx a, cplx b) { return mk(a.x + b.x, a.y + b.y); }
RP_DEV cplx csub(cplx a, cplx b) { return mk(a.x - b.x, a.y - b.y); }
RP_DEV cplx cmul(cplx a, cplx b) { return mk(fma(a.x, b.x, -a.y * b.y), fma(a.x, b.y, a.y * b.x)); }
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

// Dynamic shared memory base.  Every shared-memory pointer below is derived
// from this symbol inside the function that uses it, so the compiler emits
// LDS/STS (not generic loads) and never has to assume aliasing with global
// or local memory.
#ifdef RP_EMU
#define RP_SMEM ((cplx*)cuemu::dyn_smem())
#else
extern __shared__ __align__(16) unsigned char rp_dyn_smem_raw_[];
#define RP_SMEM ((cplx*)rp_dyn_smem_raw_)
#endif

// Block context.  Built by value inside each (noinline) op from the program
// header; it is never passed by reference across a call, so it stays in registers.
struct Blk {
  cplx* regs;
  cplx* wb;
  cplx* scr;
  int wb_off;  // offset of wb from RP_SMEM in cplx units
  int T, capP, wbP, wbT;
  int tid, nthr;
  int unit0, nunits, axis;
};
RP_DEV cplx* lane_ptr(const Blk& b, int r, int t) { return b.regs + (r * b.T + t) * b.capP; }
RP_DEV Blk make_blk(const Program* __restrict__ pg) {
  Blk b;
  b.T = pg->T;
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

// One Stockham radix-R pass, in place through registers.  Each thread owns K
// butterflies of one FFT; `per` = (L/R)/K threads serve one FFT.
// Arguments are scalars (offsets, not pointers): the buffers are addressed from
// RP_SMEM so the accesses compile to LDS/STS.
enum { FF_CONJ_IN = 1, FF_CONJ_OUT = 2, FF_LIN = 4 };
template <int R, int K>
RP_DEVNI void fft_pass(int tid, int nthr, int base_off, int stride, int nslots, int L, int Ns,
                       const cplx* __restrict__ tw, int fl, double oscale, Lay lin) {
  const int nb = L / R;
  const int per = nb / K;
  const int fpr = nthr / per;  // FFTs per round
  const int twstep = L / (Ns * R);
  const bool conj_in = fl & FF_CONJ_IN, conj_out = fl & FF_CONJ_OUT, use_lin = fl & FF_LIN;
  cplx* const base = RP_SMEM + base_off;
  for (int s0 = 0; s0 < nslots; s0 += fpr) {
    const int t = s0 + tid / per;
    const int q = tid % per;
    const bool act = (t < nslots) && (tid < fpr * per);
    cplx v[K][R];
    if (act) {
      const cplx* x = base + t * stride;
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const int j = q + per * kk;
        const int k = j % Ns;
#pragma unroll
        for (int r = 0; r < R; ++r) {
          const int src = j + r * nb;
          cplx c = x[padi(use_lin ? slot_of(lin, src) : src)];
          if (conj_in) c.y = -c.y;
          if (r > 0 && Ns > 1) c = cmul(c, __ldg(&tw[r * k * twstep]));
          v[kk][r] = c;
        }
        Dft<R>::run(v[kk]);
      }
    }
    __syncthreads();
    if (act) {
      cplx* x = base + t * stride;
#pragma unroll
      for (int kk = 0; kk < K; ++kk) {
        const int j = q + per * kk;
        const int k = j % Ns;
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
RP_DEVNI void fft_run(int tid, int nthr, int base_off, int stride, int nslots, int L, const cplx* tw, bool inverse,
                      double scale, Lay lin, bool use_lin) {
  int rem = L, Ns = 1;
  while (rem > 1) {
    const int R = rem >= 16 ? 16 : rem;
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
    for (int idx = b.tid; idx < ns * Lb; idx += b.nthr) {
      const int t = idx / Lb, j = idx % Lb;
      cplx v = mk(0.0, 0.0);
      if (j < L) {
        v = base[(g + t) * stride + padi(j)];
        if (inverse) v.y = -v.y;
        v = cmul(v, __ldg(&chirp[j]));
      }
      b.wb[t * b.wbP + padi(j)] = v;
    }
    __syncthreads();
    fft_run(b, b.wb_off, b.wbP, ns, Lb, P.tw, false, 1.0);
    for (int idx = b.tid; idx < ns * Lb; idx += b.nthr) {
      const int t = idx / Lb, j = idx % Lb;
      cplx* w = &b.wb[t * b.wbP + padi(j)];
      *w = cmul(*w, __ldg(&bhat[j]));
    }
    __syncthreads();
    fft_run(b, b.wb_off, b.wbP, ns, Lb, P.tw, true, 1.0);
    for (int idx = b.tid; idx < ns * L; idx += b.nthr) {
      const int t = idx / L, j = idx % L;
      cplx v = cmul(b.wb[t * b.wbP + padi(j)], __ldg(&chirp[j]));
      if (inverse) v.y = -v.y;
      base[(g + t) * stride + padi(j)] = cscale(v, scale);
    }
    __syncthreads();
  }
}

// ===========================================================================
// Chunked two-pass linear recurrences along parity chains.
//   y_k = s*q_k + p*y_{k-1} + r*y_{k-2}      (k in dependency order)
// Chain c of slot t: element k at slot c0[c] + cs[c]*m, m = fwd ? k : M-1-k.
// F(t, c, m) -> coefficients (componentwise double2).
// ===========================================================================
struct Coef {
  cplx s, p, r;
};
struct Chains {
  int nch;       // chains per slot (1 or 2)
  int c0[2], cs[2], M[2];
};
enum { RP_CHUNK = RP_CHUNK_HOST };

template <class F>
RP_DEV void chain_solve(const Blk& b, int reg, const Chains& ch, bool fwd, F coef) {
  int Mmax = ch.M[0];
  if (ch.nch > 1 && ch.M[1] > Mmax) Mmax = ch.M[1];
  const int nck = (Mmax + RP_CHUNK - 1) / RP_CHUNK;
  const int nitems = b.T * ch.nch * nck;
  // pass 1: zero-carry chunk summaries + homogeneous responses
  for (int it = b.tid; it < nitems; it += b.nthr) {
    const int ck = it % nck, c = (it / nck) % ch.nch, t = it / (nck * ch.nch);
    const int M = ch.M[c];
    const cplx* x = lane_ptr(b, reg, t);
    cplx y1 = mk(0, 0), y2 = mk(0, 0), u1 = mk(1, 1), u2 = mk(0, 0), v1 = mk(0, 0), v2 = mk(1, 1);
    const int k0 = ck * RP_CHUNK, k1 = min(M, k0 + RP_CHUNK);
    for (int k = k0; k < k1; ++k) {
      const int m = fwd ? k : M - 1 - k;
      const Coef cf = coef(t, c, m);
      const cplx q = x[padi(ch.c0[c] + ch.cs[c] * m)];
      cplx y = pfma(cf.p, y1, pfma(cf.r, y2, pmul(cf.s, q)));
      y2 = y1;
      y1 = y;
      cplx u = pfma(cf.p, u1, pmul(cf.r, u2));
      u2 = u1;
      u1 = u;
      cplx v = pfma(cf.p, v1, pmul(cf.r, v2));
      v2 = v1;
      v1 = v;
    }
    cplx* s = b.scr + (size_t)it * 8;
    s[0] = y1;
    s[1] = y2;
    s[2] = u1;
    s[3] = u2;
    s[4] = v1;
    s[5] = v2;
  }
  __syncthreads();
  // pass 2: carry chain over chunks (one thread per chain)
  for (int cid = b.tid; cid < b.T * ch.nch; cid += b.nthr) {
    cplx c1 = mk(0, 0), c2 = mk(0, 0);
    for (int ck = 0; ck < nck; ++ck) {
      cplx* s = b.scr + (size_t)(cid * nck + ck) * 8;
      s[6] = c1;
      s[7] = c2;
      cplx n1 = pfma(s[2], c1, pfma(s[4], c2, s[0]));
      cplx n2 = pfma(s[3], c1, pfma(s[5], c2, s[1]));
      c1 = n1;
      c2 = n2;
    }
  }
  __syncthreads();
  // pass 3: re-run with the true carries, in place
  for (int it = b.tid; it < nitems; it += b.nthr) {
    const int ck = it % nck, c = (it / nck) % ch.nch, t = it / (nck * ch.nch);
    const int M = ch.M[c];
    cplx* x = lane_ptr(b, reg, t);
    const cplx* s = b.scr + (size_t)it * 8;
    cplx y1 = s[6], y2 = s[7];
    const int k0 = ck * RP_CHUNK, k1 = min(M, k0 + RP_CHUNK);
    for (int k = k0; k < k1; ++k) {
      const int m = fwd ? k : M - 1 - k;
      const Coef cf = coef(t, c, m);
      cplx* px = &x[padi(ch.c0[c] + ch.cs[c] * m)];
      cplx y = pfma(cf.p, y1, pfma(cf.r, y2, pmul(cf.s, *px)));
      y2 = y1;
      y1 = y;
      *px = y;
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

// Chain stencil: out_m = sum_{d=0..2} w_d(i) * in_{m + dir*d} walking the chain
// so that in-place is safe (halo elements are fetched before the barrier).
//   dir = -1: out_i uses in_i, in_{i-2}          (to_ortho)
//   dir = +1: out_i uses in_i, in_{i+2}, in_{i+4} (S^T, B2 matvec)
// nin = number of valid input elements (others read as 0), nout = outputs.
template <class W>
RP_DEV void chain_stencil(const Blk& b, int reg, const Lay& L, int nin, int nout, int dir, W wfun) {
  const int nmax = nin > nout ? nin : nout;
  const Chains ch = chains_of(L, nmax);
  int Mmax = ch.M[0];
  const int nck = (Mmax + RP_CHUNK - 1) / RP_CHUNK;
  const int nitems = b.T * 2 * nck;
  // every thread handles at most RP_SI items so halos can sit in registers
  enum { RP_SI = 4 };
  cplx h1[RP_SI], h2[RP_SI];
  int cnt = 0;
  for (int it = b.tid; it < nitems && cnt < RP_SI; it += b.nthr, ++cnt) {
    const int ck = it % nck, c = (it / nck) % 2, t = it / (nck * 2);
    const cplx* x = lane_ptr(b, reg, t);
    // halo = the two chain elements just beyond the chunk in direction dir
    const int mh = (dir > 0) ? (ck + 1) * RP_CHUNK : ck * RP_CHUNK - 1;
    const int mh2 = mh + dir;
    const int i1 = 2 * mh + c, i2 = 2 * mh2 + c;
    h1[cnt] = (mh >= 0 && i1 < nin) ? x[padi(ch.c0[c] + ch.cs[c] * mh)] : mk(0, 0);
    h2[cnt] = (mh2 >= 0 && i2 < nin) ? x[padi(ch.c0[c] + ch.cs[c] * mh2)] : mk(0, 0);
  }
  __syncthreads();
  cnt = 0;
  for (int it = b.tid; it < nitems && cnt < RP_SI; it += b.nthr, ++cnt) {
    const int ck = it % nck, c = (it / nck) % 2, t = it / (nck * 2);
    cplx* x = lane_ptr(b, reg, t);
    const int M = ch.M[c];
    const int m0 = ck * RP_CHUNK, m1 = min(M, m0 + RP_CHUNK);
    if (m0 >= m1) continue;
    if (dir > 0) {
      // ascending: window (in_m, in_{m+1}, in_{m+2})
      cplx a0, a1, a2;
      auto ld = [&](int m) -> cplx {
        if (m >= m1) {
          return (m == m1) ? h1[cnt] : h2[cnt];
        }
        const int i = 2 * m + c;
        return (i < nin) ? x[padi(ch.c0[c] + ch.cs[c] * m)] : mk(0, 0);
      };
      a0 = ld(m0);
      a1 = ld(m0 + 1);
      a2 = ld(m0 + 2);
      for (int m = m0; m < m1; ++m) {
        const int i = 2 * m + c;
        cplx o = wfun(i, a0, a1, a2);
        if (i < nout) x[padi(ch.c0[c] + ch.cs[c] * m)] = o;
        a0 = a1;
        a1 = a2;
        a2 = ld(m + 3);
      }
    } else {
      // descending: window (in_m, in_{m-1})
      auto ld = [&](int m) -> cplx {
        if (m < m0) return (m == m0 - 1) ? h1[cnt] : h2[cnt];
        const int i = 2 * m + c;
        return (i < nin) ? x[padi(ch.c0[c] + ch.cs[c] * m)] : mk(0, 0);
      };
      cplx a0 = ld(m1 - 1), a1 = ld(m1 - 2);
      for (int m = m1 - 1; m >= m0; --m) {
        const int i = 2 * m + c;
        cplx o = wfun(i, a0, a1, mk(0, 0));
        if (i < nout) x[padi(ch.c0[c] + ch.cs[c] * m)] = o;
        a0 = a1;
        a1 = ld(m - 2);
      }
    }
  }
  __syncthreads();
}

// ===========================================================================
// Lane index helpers
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

RP_DEVNI void op_ld(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, nz = (I.n2 > n && !(I.flags & LF_ACC)) ? I.n2 : n;
  const int total = b.T * nz;
  const double* lc = (const double*)I.p1;
  for (int idx = b.tid; idx < total; idx += b.nthr) {
    int t, i;
    if (b.axis == AXIS_Y) {
      t = idx / nz;
      i = idx % nz;
    } else {
      i = idx / b.T;
      t = idx % b.T;
    }
    const LaneSel s = lane_sel(b, I, t);
    cplx v = mk(0, 0);
    if (i < n) {
      if (I.flags & LF_COMPLEX) {
        if (s.va) {
          const cplx* src = (const cplx*)I.p0;
          v = (b.axis == AXIS_Y) ? src[(size_t)s.la * I.ld + i] : src[(size_t)i * I.ld + s.la];
        }
      } else if (I.flags & LF_BCAST) {
        if (s.va) {
          const double* src = (const double*)I.p0;
          double d = (b.axis == AXIS_Y) ? src[(size_t)s.la * I.ld + i] : src[(size_t)i * I.ld + s.la];
          v = mk(d, d);
        }
      } else {
        const double* src = (const double*)I.p0;
        if (b.axis == AXIS_Y) {
          if (s.va) v.x = src[(size_t)s.la * I.ld + i];
          if (s.vb) v.y = src[(size_t)s.lb * I.ld + i];
        } else {
          if (s.va) v.x = src[(size_t)i * I.ld + s.la];
          if (s.vb) v.y = src[(size_t)i * I.ld + s.lb];
        }
      }
      double ca = I.s0, cb = I.s0;
      if (I.flags & LF_LANECOEF) {
        ca *= s.va ? __ldg(&lc[s.la]) : 0.0;
        cb *= s.vb ? __ldg(&lc[s.lb]) : 0.0;
      }
      if (I.flags & LF_MULIK) {
        const double k = (double)s.la * ca;
        v = mk(-k * v.y, k * v.x);
      } else {
        v = mk(v.x * ca, v.y * cb);
      }
    }
    cplx* x = lane_ptr(b, I.r0, t) + padi(slot_of(I.lay, i));
    if (I.flags & LF_ACC) {
      if (i < n) *x = cadd(*x, v);
    } else {
      *x = v;
    }
  }
  __syncthreads();
}

RP_DEVNI void op_st(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n;
  const int total = b.T * n;
  for (int idx = b.tid; idx < total; idx += b.nthr) {
    int t, i;
    if (b.axis == AXIS_Y) {
      t = idx / n;
      i = idx % n;
    } else {
      i = idx / b.T;
      t = idx % b.T;
    }
    const LaneSel s = lane_sel(b, I, t);
    cplx v = lane_ptr(b, I.r0, t)[padi(slot_of(I.lay, i))];
    v = cscale(v, I.s0);
    bool cuta = false, cutb = false;
    if (I.flags & LF_CUT) {
      cuta = (i >= I.i0) || (I.i1 >= 0 && s.la >= I.i1);
      cutb = (i >= I.i0) || (I.i1 >= 0 && s.lb >= I.i1);
      if (cuta) v.x = 0.0;
      if (cutb) v.y = 0.0;
      if ((I.flags & LF_COMPLEX) && cuta) v = mk(0, 0);
    }
    if (I.flags & LF_COMPLEX) {
      if (s.va) {
        cplx* dst = (cplx*)I.p0;
        cplx* d = (b.axis == AXIS_Y) ? &dst[(size_t)s.la * I.ld + i] : &dst[(size_t)i * I.ld + s.la];
        *d = (I.flags & LF_ACC) ? cadd(*d, v) : v;
      }
    } else {
      double* dst = (double*)I.p0;
      if (s.va) {
        double* d = (b.axis == AXIS_Y) ? &dst[(size_t)s.la * I.ld + i] : &dst[(size_t)i * I.ld + s.la];
        *d = (I.flags & LF_ACC) ? *d + v.x : v.x;
      }
      if (s.vb) {
        double* d = (b.axis == AXIS_Y) ? &dst[(size_t)s.lb * I.ld + i] : &dst[(size_t)i * I.ld + s.lb];
        *d = (I.flags & LF_ACC) ? *d + v.y : v.y;
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
  for (int idx = b.tid; idx < b.T * n; idx += b.nthr) {
    const int t = idx / n, i = idx % n;
    cplx* x0 = lane_ptr(b, I.r0, t) + padi(slot_of(I.lay, i));
    switch (mode) {
      case 0: *x0 = lane_ptr(b, I.r1, t)[padi(slot_of(I.lay2, i))]; break;
      case 1: {
        cplx a = lane_ptr(b, I.r1, t)[padi(slot_of(I.lay2, i))];
        *x0 = mk(fma(I.s0, a.x, x0->x), fma(I.s0, a.y, x0->y));
      } break;
      case 2: *x0 = cscale(*x0, I.s0); break;
      case 3:
      case 4: {
        cplx a = lane_ptr(b, I.r1, t)[padi(slot_of(I.lay2, i))];
        cplx c = lane_ptr(b, I.r2, t)[padi(slot_of(I.lay2, i))];
        *x0 = (mode == 4) ? pfma(a, c, *x0) : pmul(a, c);
      } break;
      case 5: *x0 = mk(0, 0); break;
      case 6:
        if (i >= I.i0) *x0 = mk(0, 0);
        break;
    }
  }
  __syncthreads();
}

RP_DEVNI void op_mulik(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n;
  for (int idx = b.tid; idx < b.T * n; idx += b.nthr) {
    const int t = idx / n, i = idx % n;
    const double k = (double)((I.flags & LF_ELEMK) ? i : (b.unit0 + t)) * I.s0;
    cplx* x0 = lane_ptr(b, I.r0, t) + padi(slot_of(I.lay, i));
    cplx v = *x0;
    *x0 = mk(-k * v.y, k * v.x);
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
  const double* d = (const double*)I.p0;
  const double* l = (const double*)I.p1;
  chain_stencil(b, I.r0, I.lay, m, n, -1, [=](int i, cplx a0, cplx a1, cplx) -> cplx {
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
  const double* d = (const double*)I.p0;
  const double* l = (const double*)I.p1;
  const TdmaTab tt = *(const TdmaTab*)I.p2;
  chain_stencil(b, I.r0, I.lay, n, m, +1, [=](int i, cplx a0, cplx a1, cplx) -> cplx {
    if (i >= m) return mk(0, 0);
    const double di = __ldg(&d[i]), li = __ldg(&l[i]);
    return mk(fma(di, a0.x, li * a1.x), fma(di, a0.y, li * a1.y));
  });
  const Chains ch = chains_of(I.lay, m);
  chain_solve(b, I.r0, ch, true, [=](int, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    const double s = __ldg(&tt.fs[i]), p = __ldg(&tt.fp[i]);
    Coef cf;
    cf.s = mk(s, s);
    cf.p = mk(p, p);
    cf.r = mk(0, 0);
    return cf;
  });
  chain_solve(b, I.r0, ch, false, [=](int, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    const double p = __ldg(&tt.bp[i]);
    Coef cf;
    cf.s = mk(1, 1);
    cf.p = mk(p, p);
    cf.r = mk(0, 0);
    return cf;
  });
}

// B2 preconditioner (matvec.rs:172-193): out_r = lo_r x_r + di_r x_{r+2} + up_r x_{r+4}
RP_DEVNI void op_bandmv(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const int n = I.n, m = n - 2;
  const double* lo = (const double*)I.p0;
  const double* di = (const double*)I.p1;
  const double* up = (const double*)I.p2;
  chain_stencil(b, I.r0, I.lay, n, m, +1, [=](int i, cplx a0, cplx a1, cplx a2) -> cplx {
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
    chain_solve(b, I.r0, ch, false, [=](int, int c, int mm) -> Coef {
      const double s = 2.0 * (double)(2 * mm + c) * sc;
      Coef cf;
      cf.s = mk(s, s);
      cf.p = mk(1, 1);
      cf.r = mk(0, 0);
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
  chain_solve(b, I.r0, ch, true, [=](int, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    const double p = __ldg(&ft.fp[i]);
    Coef cf;
    cf.s = mk(1, 1);
    cf.p = mk(p, p);
    cf.r = mk(0, 0);
    return cf;
  });
  chain_solve(b, I.r0, ch, false, [=](int, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    const double s = __ldg(&ft.bs[i]), p = __ldg(&ft.bp1[i]), r = __ldg(&ft.bp2[i]);
    Coef cf;
    cf.s = mk(s, s);
    cf.p = mk(p, p);
    cf.r = mk(r, r);
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
  auto mu_of = [=](int t) -> cplx {
    const int P = b.unit0 + t;
    int la = cplxl ? P : 2 * P, lb = cplxl ? P : 2 * P + 1;
    if (la >= I.nlanes) la = I.nlanes - 1;
    if (lb >= I.nlanes) lb = I.nlanes - 1;
    return mk(__ldg(&mt.lam[la]) + mt.alpha, __ldg(&mt.lam[lb]) + mt.alpha);
  };
  // forward: x_i -= l_{i-2} x_{i-2},  l_j = low_j / dia'_j
  chain_solve(b, I.r0, ch, true, [=](int t, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    Coef cf;
    cf.s = mk(1, 1);
    cf.r = mk(0, 0);
    cf.p = mk(0, 0);
    if (i >= 2) {
      const cplx mu = mu_of(t);
      const double al = __ldg(&mt.a_low[i - 2]), cl = __ldg(&mt.c_low[i - 2]);
      const cplx inv = lane_ptr(b, r1, t)[padi(slot_of(L, i - 2))];
      cf.p = mk(-fma(mu.x, cl, al) * inv.x, -fma(mu.y, cl, al) * inv.y);
    }
    return cf;
  });
  // backward: x_i = (x_i - up1'_i x_{i+2} - up2_i x_{i+4}) / dia'_i
  chain_solve(b, I.r0, ch, false, [=](int t, int c, int mm) -> Coef {
    const int i = 2 * mm + c;
    const cplx mu = mu_of(t);
    const cplx inv = lane_ptr(b, r1, t)[padi(slot_of(L, i))];
    cplx u1 = mk(0, 0), u2 = mk(0, 0);
    if (i < n - 2) {
      const double a = __ldg(&mt.a_up1[i]), cc = __ldg(&mt.c_up1[i]);
      u1 = mk(fma(mu.x, cc, a), fma(mu.y, cc, a));
      if (i >= 2) {
        const double al = __ldg(&mt.a_low[i - 2]), cl = __ldg(&mt.c_low[i - 2]);
        const double a2 = __ldg(&mt.a_up2[i - 2]), c2 = __ldg(&mt.c_up2[i - 2]);
        const cplx invm = lane_ptr(b, r1, t)[padi(slot_of(L, i - 2))];
        const cplx lw = mk(fma(mu.x, cl, al) * invm.x, fma(mu.y, cl, al) * invm.y);
        if (i - 2 < n - 4) {
          u1.x = fma(-lw.x, fma(mu.x, c2, a2), u1.x);
          u1.y = fma(-lw.y, fma(mu.y, c2, a2), u1.y);
        }
      }
    }
    if (i < n - 4) {
      const double a = __ldg(&mt.a_up2[i]), cc = __ldg(&mt.c_up2[i]);
      u2 = mk(fma(mu.x, cc, a), fma(mu.y, cc, a));
    }
    Coef cf;
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
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const DctPlan P = *(const DctPlan*)I.p0;
  const int n = P.n, N = n - 1;
  const bool backward = I.i0 != 0;
  const Lay lin = I.lay;
  const int npairs = N / 2 + 1;
  const int nwarp = (b.nthr + 31) >> 5;
  const int group = P.fft.pow2 ? b.T : b.wbT;
  for (int g = 0; g < b.T; g += group) {
    const int ns = min(group, b.T - g);
    // ---- pre-combine, in place at the source slots (the first FFT pass
    //      reads through the layout map; Bluestein writes to the work buffer)
    for (int tt = 0; tt < ns; ++tt) {
      const int t = g + tt;
      cplx* x = lane_ptr(b, I.r0, t);
      cplx f1 = mk(0, 0);
      for (int j = b.tid; j < npairs; j += b.nthr) {
        const int jm = N - j;
        const int sj = padi(slot_of(lin, j)), sm = padi(slot_of(lin, jm));
        cplx a = x[sj], c = x[sm];
        if (backward) {  // c_k * (-1)^k / 2, ends doubled
          double ga = (j & 1) ? -0.5 : 0.5, gc = (jm & 1) ? -0.5 : 0.5;
          if (j == 0) ga *= 2.0;
          if (jm == N) gc *= 2.0;
          a = cscale(a, ga);
          c = cscale(c, gc);
        }
        const cplx sc = __ldg(&P.sc[j]);
        const cplx sum = cadd(a, c), dif = csub(a, c);
        const cplx za = mk(fma(-sc.x, dif.x, 0.5 * sum.x), fma(-sc.x, dif.y, 0.5 * sum.y));
        const cplx zb = mk(fma(sc.x, dif.x, 0.5 * sum.x), fma(sc.x, dif.y, 0.5 * sum.y));
        const double w = (j == 0) ? 0.5 * sc.y : sc.y;
        f1.x = fma(w, dif.x, f1.x);
        f1.y = fma(w, dif.y, f1.y);
        if (P.fft.pow2) {
          x[sj] = za;
          if (jm != j && j != 0) x[sm] = zb;
        } else {
          cplx* wbuf = b.wb + (size_t)tt * b.wbP;
          wbuf[padi(j)] = cmul(za, __ldg(&P.fft.chirp[j]));
          if (jm != j && j != 0) wbuf[padi(jm)] = cmul(zb, __ldg(&P.fft.chirp[jm]));
        }
      }
      for (int o = 16; o > 0; o >>= 1) {
        f1.x += __shfl_down_sync(0xffffffffu, f1.x, o);
        f1.y += __shfl_down_sync(0xffffffffu, f1.y, o);
      }
      if ((b.tid & 31) == 0) b.scr[t * 32 + (b.tid >> 5)] = f1;
    }
    __syncthreads();
    // ---- complex DFT of length N ----------------------------------------
    if (P.fft.pow2) {
      fft_run(b.tid, b.nthr, (I.r0 * b.T + g) * b.capP, b.capP, ns, N, P.fft.tw, false, 1.0, lin, true);
    } else {
      const int Lb = P.fft.Lb;
      for (int idx = b.tid; idx < ns * (Lb - N); idx += b.nthr) {
        const int t = idx / (Lb - N), j = N + idx % (Lb - N);
        b.wb[(size_t)t * b.wbP + padi(j)] = mk(0, 0);
      }
      __syncthreads();
      fft_run(b, b.wb_off, b.wbP, ns, Lb, P.fft.tw, false, 1.0);
      for (int idx = b.tid; idx < ns * Lb; idx += b.nthr) {
        const int t = idx / Lb, j = idx % Lb;
        cplx* w = &b.wb[(size_t)t * b.wbP + padi(j)];
        *w = cmul(*w, __ldg(&P.fft.bhat[j]));
      }
      __syncthreads();
      fft_run(b, b.wb_off, b.wbP, ns, Lb, P.fft.tw, true, 1.0);
      for (int idx = b.tid; idx < ns * N; idx += b.nthr) {
        const int t = idx / N, j = idx % N;
        lane_ptr(b, I.r0, g + t)[padi(j)] = cmul(b.wb[(size_t)t * b.wbP + padi(j)], __ldg(&P.fft.chirp[j]));
      }
      __syncthreads();
    }
  }
  // ---- recombine: X_{2k} = Z_k + Z_{N-k};  D_k = i (Z_k - Z_{N-k}) -------
  const double he = backward ? 1.0 : 1.0 / (double)N;  // even outputs: (+1)/N
  const int Ko = (N - 1) / 2;                           // odd outputs O_0..O_Ko
  for (int idx = b.tid; idx < b.T * npairs; idx += b.nthr) {
    const int t = idx / npairs, k = idx % npairs;
    cplx* x = lane_ptr(b, I.r0, t);
    const cplx zk = x[padi(k)];
    const cplx zm = (k == 0) ? zk : x[padi(N - k)];
    cplx e = cadd(zk, zm);
    double h = he;
    if (!backward && (k == 0 || 2 * k == N)) h *= 0.5;
    x[padi(k)] = cscale(e, h);
    if (k >= 1 && k <= Ko) x[padi(N - k)] = mk(-(zk.y - zm.y), zk.x - zm.x);
    if (k == 0) {
      cplx f1 = mk(0, 0);
      for (int w = 0; w < nwarp; ++w) f1 = cadd(f1, b.scr[t * 32 + w]);
      x[padi(N)] = cscale(f1, 2.0);
    }
  }
  __syncthreads();
  // ---- odd outputs: prefix sum along slots N, N-1, ... -------------------
  Chains ch;
  ch.nch = 1;
  ch.c0[0] = N;
  ch.cs[0] = -1;
  ch.M[0] = Ko + 1;
  ch.c0[1] = 0;
  ch.cs[1] = 0;
  ch.M[1] = 0;
  chain_solve(b, I.r0, ch, true, [=](int, int, int) -> Coef {
    Coef cf;
    cf.s = mk(1, 1);
    cf.p = mk(1, 1);
    cf.r = mk(0, 0);
    return cf;
  });
  if (!backward) {  // odd outputs: (-1)/N, last one halved when N is odd
    const double ho = -1.0 / (double)N;
    for (int idx = b.tid; idx < b.T * (Ko + 1); idx += b.nthr) {
      const int t = idx / (Ko + 1), k = idx % (Ko + 1);
      cplx* x = &lane_ptr(b, I.r0, t)[padi(N - k)];
      *x = cscale(*x, (2 * k + 1 == N) ? 0.5 * ho : ho);
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
  cplx* dst = (cplx*)I.p0;
  for (int idx = b.tid; idx < b.T * m; idx += b.nthr) {
    const int k = idx / b.T, t = idx % b.T;
    const int ca = 2 * (b.unit0 + t), cb = ca + 1;
    const cplx* z = lane_ptr(b, I.r0, t);
    const cplx zk = z[padi(k)];
    const cplx zm = cconj(z[padi((n - k) % n)]);
    const cplx s = cadd(zk, zm), d = csub(zk, zm);
    double h = 0.5 * I.s0;
    if ((I.flags & LF_CUT) && k >= I.i0) h = 0.0;
    if (ca < I.nlanes) dst[(size_t)k * I.ld + ca] = mk(h * s.x, h * s.y);
    if (cb < I.nlanes) dst[(size_t)k * I.ld + cb] = mk(h * d.y, -h * d.x);
  }
  __syncthreads();
}

RP_DEVNI void op_irfft(const Program* __restrict__ pg, int pc) {
  const Blk b = make_blk(pg);
  const Instr I = pg->ins[pc];
  const FftPlan P = *(const FftPlan*)I.p1;
  const int n = P.L, m = n / 2 + 1;
  const cplx* src = (const cplx*)I.p0;
  for (int idx = b.tid; idx < b.T * m; idx += b.nthr) {
    const int k = idx / b.T, t = idx % b.T;
    const int ca = 2 * (b.unit0 + t), cb = ca + 1;
    cplx xa = (ca < I.nlanes) ? src[(size_t)k * I.ld + ca] : mk(0, 0);
    cplx xb = (cb < I.nlanes) ? src[(size_t)k * I.ld + cb] : mk(0, 0);
    if (I.flags & LF_MULIK) {
      const double kk = (double)k * I.s0;
      xa = mk(-kk * xa.y, kk * xa.x);
      xb = mk(-kk * xb.y, kk * xb.x);
    } else {
      xa = cscale(xa, I.s0);
      xb = cscale(xb, I.s0);
    }
    if (k == 0 || 2 * k == n) {
      xa.y = 0.0;
      xb.y = 0.0;
    }
    cplx* z = lane_ptr(b, I.r0, t);
    z[padi(k)] = mk(xa.x - xb.y, xa.y + xb.x);  // X_a + i X_b
    if (k > 0 && 2 * k != n) z[padi(n - k)] = mk(xa.x + xb.y, -xa.y + xb.x);  // conj(X_a) + i conj(X_b)
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
  for (int pc = 0; pc < ninstr; ++pc) {
    const int op = pg->ins[pc].op;
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
  }
}

}  // namespace rp
