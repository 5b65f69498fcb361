// progbuild.cu -- device arrays and the lane-program builder (model.h).
#include <algorithm>

#include "model.h"

namespace rp {

void Arr::alloc(int r, int c, bool cx) {
  rows = r;
  cols = c;
  cplx = cx;
  ld = cx ? ((c + 3) & ~3) : ((c + 7) & ~7);
  buf = DevBuf((size_t)rows * ld * elem_bytes());
}
void Arr::upload(const double* host, cudaStream_t s) {
  const size_t w = (size_t)cols * elem_bytes();
  rt::h2d_2d(buf.p, (size_t)ld * elem_bytes(), host, w, w, rows, s);
}
void Arr::download(double* host, cudaStream_t s) const {
  const size_t w = (size_t)cols * elem_bytes();
  rt::d2h_2d(host, w, buf.p, (size_t)ld * elem_bytes(), w, rows, s);
}
void Arr::zero(cudaStream_t s) { rt::dzero(buf.p, buf.bytes, s); }

void Built::dump_prof(const char* name) const {
  if (!dprof.p) return;
  static const char* names[] = {"end", "ld", "st", "zero", "copy", "axpy", "scale", "mulpw", "cut", "mulik", "toortho",
                                "fromortho", "diff", "dct", "rfft", "irfft", "bandmv", "fdma", "fdmamode", "setzero00"};
  long long c[64];
  rt::d2h(c, dprof.p, sizeof(c), 0);
  rt::sync(0);
  long long tot = 0;
  for (int i = 0; i < 20; ++i) tot += c[i];
  fprintf(stderr, "  [opprof] %-26s blocks %d thr %d smem %d: total %lld cyc:", name, nblocks, nthreads, smem, tot);
  for (int i = 0; i < 20; ++i)
    if (c[i]) fprintf(stderr, " %s=%.1f%%", names[i], 100.0 * c[i] / (double)tot);
  if (c[32] + c[40])
    fprintf(stderr, " | dct: pre=%lld fft=%lld recomb=%lld scan=%lld | ld: prologue=%lld loop=%lld bar=%lld", c[32], c[33], c[34],
            c[35], c[40], c[41], c[42]);
  fprintf(stderr, "\n");
}

void Built::launch(cudaStream_t s) const {
  if (!valid) throw Error(RP_ERR_INTERNAL, "launch of an unbuilt lane program");
  launch_lane_programs(dprog.as<Program>(), 1, nblocks, nthreads, smem, s);
}

// --------------------------------------------------------------------------
static int slot_h(const Lay& L, int i) {
  int h = i >> 1;
  return (i & 1) ? L.o0 + L.so * h : L.e0 + L.se * h;
}
static int padi_h(int s) { return s + (s >> 5); }

ProgBuilder::ProgBuilder(int axis, int nunits) {
  memset(&p_, 0, sizeof(p_));
  p_.axis = axis;
  p_.nunits = nunits;
}

Instr& ProgBuilder::add(int op) {
  if (p_.ninstr >= RP_MAX_INSTR) throw Error(RP_ERR_INTERNAL, "lane program too long");
  Instr& I = p_.ins[p_.ninstr++];
  memset(&I, 0, sizeof(I));
  I.op = op;
  I.i1 = -1;
  I.s0 = 1.0;
  return I;
}

void ProgBuilder::touch(int r, Lay lay, int n) {
  p_.nreg = std::max(p_.nreg, r + 1);
  int mx = 0;
  for (int i : {0, 1, n - 2, n - 1})
    if (i >= 0 && i < n) {
      int s = slot_h(lay, i);
      if (s < 0) throw Error(RP_ERR_INTERNAL, "negative lane slot");
      mx = std::max(mx, s);
    }
  p_.cap = std::max(p_.cap, mx + 1);
}

void ProgBuilder::ld(int r, ArrRef a, int n, Lay lay, double s, int flags, int shift, const double* lanecoef,
                     int zfill, int lane2) {
  Instr& I = add(OP_LD);
  I.r0 = r;
  I.n = n;
  I.n2 = zfill;
  I.flags = flags | (a.cplx ? LF_COMPLEX : 0);
  if (lanecoef) I.flags |= LF_LANECOEF;
  if (lane2 >= 0) {
    I.flags |= LF_LANE2;
    I.i2 = lane2;
  }
  I.shift = shift;
  I.lay = lay;
  I.p0 = a.p;
  I.p1 = lanecoef;
  I.ld = a.ld;
  I.nlanes = (p_.axis == AXIS_Y) ? a.rows : a.cols;
  I.s0 = s;
  bytes_unit_ += (double)n * ((flags & LF_BCAST) ? 8.0 : 16.0);
  touch(r, lay, std::max(n, zfill));
}

void ProgBuilder::st(int r, ArrRef a, int n, Lay lay, double s, int flags, int cut_i, int cut_lane, int lane2) {
  Instr& I = add(OP_ST);
  I.r0 = r;
  I.n = n;
  I.flags = flags | (a.cplx ? LF_COMPLEX : 0);
  if (cut_i >= 0 || cut_lane >= 0) {
    I.flags |= LF_CUT;
    I.i0 = cut_i >= 0 ? cut_i : (1 << 30);
    I.i1 = cut_lane;
  }
  if (lane2 >= 0) {
    I.flags |= LF_LANE2;
    I.i2 = lane2;
  }
  I.lay = lay;
  I.p0 = a.p;
  I.ld = a.ld;
  I.nlanes = (p_.axis == AXIS_Y) ? a.rows : a.cols;
  I.s0 = s;
  bytes_unit_ += (double)n * 16.0 * ((flags & LF_ACC) ? 2.0 : 1.0);
  touch(r, lay, n);
}

void ProgBuilder::toortho(int r, const Base& b, Lay lay) {
  if (!b.is_composite()) return;
  Instr& I = add(OP_TOORTHO);
  I.r0 = r;
  I.n = b.n;
  I.lay = lay;
  I.p0 = b.d_sd.p;
  I.p1 = b.d_sl.p;
  touch(r, lay, b.n);
}

void ProgBuilder::fromortho(int r, const Base& b, Lay lay) {
  if (!b.is_composite()) return;
  Instr& I = add(OP_FROMORTHO);
  I.r0 = r;
  I.n = b.n;
  I.lay = lay;
  I.p0 = b.d_sd.p;
  I.p1 = b.d_sl.p;
  I.p2 = b.d_tdma.p;
  touch(r, lay, b.n);
}

Lay ProgBuilder::diff(int r, int n, Lay lay, int times, double scale) {
  if (times <= 0) {
    if (scale != 1.0) this->scale(r, n, scale, lay);
    return lay;
  }
  Instr& I = add(OP_DIFF);
  I.r0 = r;
  I.n = n;
  I.lay = lay;
  I.i0 = times;
  I.s0 = scale;
  touch(r, lay, n);
  Lay L = lay;
  for (int k = 0; k < times; ++k) {
    L = lay_after_diff(L);
    touch(r, L, n);
  }
  return L;
}

Lay ProgBuilder::dct(int r, const Base& b, Lay lay, bool backward) {
  Instr& I = add(OP_DCT);
  I.r0 = r;
  I.n = b.n;
  I.lay = lay;
  I.i0 = backward ? 1 : 0;
  I.p0 = b.d_dct.p;
  touch(r, lay, b.n);
  Lay out = lay_split(b.n - 1);
  touch(r, out, b.n);
  touch(r, lay_natural(), b.n);
  fftlen_ = std::max(fftlen_, b.fft_len());
  if (!b.fft.plan.pow2) wbcap_ = std::max(wbcap_, b.fft.plan.Lb);
  return out;
}

void ProgBuilder::bandmv(int r, const Base& b, Lay lay) {
  Instr& I = add(OP_BANDMV);
  I.r0 = r;
  I.n = b.n;
  I.lay = lay;
  I.p0 = b.d_b2lo.p;
  I.p1 = b.d_b2di.p;
  I.p2 = b.d_b2up.p;
  touch(r, lay, b.n);
}

void ProgBuilder::fdma(int r, const FdmaDev& f, Lay lay) {
  Instr& I = add(OP_FDMA);
  I.r0 = r;
  I.n = f.n;
  I.lay = lay;
  I.p0 = f.tab.p;
  touch(r, lay, f.n);
}

void ProgBuilder::fdmamode(int r, int rinv, const FdmaModeDev& f, Lay lay, bool complex_lanes) {
  Instr& I = add(OP_FDMAMODE);
  I.r0 = r;
  I.r1 = rinv;
  I.n = f.n;
  I.lay = lay;
  I.p0 = f.tab.p;
  I.nlanes = f.nlanes;
  I.flags = complex_lanes ? LF_COMPLEX : 0;
  touch(r, lay, f.n);
  touch(rinv, lay, f.n);
}

void ProgBuilder::copy(int rd, int rs, int n, Lay ld, Lay ls) {
  Instr& I = add(OP_COPY);
  I.r0 = rd;
  I.r1 = rs;
  I.n = n;
  I.lay = ld;
  I.lay2 = ls;
  touch(rd, ld, n);
  touch(rs, ls, n);
}
void ProgBuilder::axpy(int rd, int rs, int n, double a, Lay ld, Lay ls) {
  Instr& I = add(OP_AXPY);
  I.r0 = rd;
  I.r1 = rs;
  I.n = n;
  I.s0 = a;
  I.lay = ld;
  I.lay2 = ls;
  touch(rd, ld, n);
  touch(rs, ls, n);
}
void ProgBuilder::scale(int r, int n, double a, Lay lay) {
  Instr& I = add(OP_SCALE);
  I.r0 = r;
  I.n = n;
  I.s0 = a;
  I.lay = lay;
  touch(r, lay, n);
}
void ProgBuilder::mulpw(int rd, int ra, int rb, int n, Lay ld, Lay lab, bool acc) {
  Instr& I = add(OP_MULPW);
  I.r0 = rd;
  I.r1 = ra;
  I.r2 = rb;
  I.n = n;
  I.lay = ld;
  I.lay2 = lab;
  I.flags = acc ? LF_ACC : 0;
  touch(rd, ld, n);
  touch(ra, lab, n);
  touch(rb, lab, n);
}
void ProgBuilder::cut(int r, int n, int from, Lay lay) {
  Instr& I = add(OP_CUT);
  I.r0 = r;
  I.n = n;
  I.i0 = from;
  I.lay = lay;
  touch(r, lay, n);
}
void ProgBuilder::mulik(int r, int n, double a, Lay lay, bool elem_k) {
  Instr& I = add(OP_MULIK);
  I.r0 = r;
  I.n = n;
  I.s0 = a;
  I.lay = lay;
  I.flags = elem_k ? LF_ELEMK : 0;
  touch(r, lay, n);
}
void ProgBuilder::zero(int r, int n, Lay lay) {
  Instr& I = add(OP_ZERO);
  I.r0 = r;
  I.n = n;
  I.lay = lay;
  touch(r, lay, n);
}
void ProgBuilder::setzero00(int r, Lay lay, bool complex_lanes) {
  Instr& I = add(OP_SETZERO00);
  I.r0 = r;
  I.lay = lay;
  I.flags = complex_lanes ? LF_COMPLEX : 0;
}
void ProgBuilder::rfft_st(int r0, ArrRef dst, const Base& b, double s, int cut_k) {
  Instr& I = add(OP_RFFT);
  I.r0 = r0;
  I.n = b.n;
  I.p0 = dst.p;
  I.p1 = b.d_fft.p;
  I.ld = dst.ld;
  I.nlanes = dst.cols;
  I.s0 = s;
  if (cut_k >= 0) {
    I.flags |= LF_CUT;
    I.i0 = cut_k;
  }
  bytes_unit_ += (double)b.m * 32.0;
  touch(r0, lay_natural(), b.n);
  fftlen_ = std::max(fftlen_, b.fft_len());
  if (!b.fft.plan.pow2) wbcap_ = std::max(wbcap_, b.fft.plan.Lb);
}
void ProgBuilder::irfft_ld(int r0, ArrRef src, const Base& b, double s, bool mulik) {
  Instr& I = add(OP_IRFFT);
  I.r0 = r0;
  I.n = b.n;
  I.p0 = src.p;
  I.p1 = b.d_fft.p;
  I.ld = src.ld;
  I.nlanes = src.cols;
  I.s0 = s;
  if (mulik) I.flags |= LF_MULIK;
  bytes_unit_ += (double)b.m * 32.0;
  touch(r0, lay_natural(), b.n);
  fftlen_ = std::max(fftlen_, b.fft_len());
  if (!b.fft.plan.pow2) wbcap_ = std::max(wbcap_, b.fft.plan.Lb);
}

static int pow2_ceil(int v) {
  int p = 1;
  while (p < v) p <<= 1;
  return p;
}

Built ProgBuilder::build() {
  Program& p = p_;
  const int capP = padi_h(p.cap) + 1;
  const int nck = ((p.cap + 1) / 2 + 1 + RP_CHUNK_HOST - 1) / RP_CHUNK_HOST;
  const int wbP = wbcap_ ? padi_h(wbcap_) + 1 : 0;
  const int budget = RP_MAX_SMEM - 1024;
  int bestT = 0, bestWb = 0, bestSmem = 0;
  const int candX[] = {4, 2, 1}, candY[] = {2, 1};
  const int* cand = (p.axis == AXIS_X) ? candX : candY;
  const int ncand = (p.axis == AXIS_X) ? 3 : 2;
  // tuning knobs (development): RUSTPDE_B200_SOFT = soft smem budget fraction in percent for row programs,
  // RUSTPDE_B200_TMULT = multiply the thread count by this factor
  const char* e_soft = getenv("RUSTPDE_B200_SOFT");
  const int soft_pct = e_soft ? atoi(e_soft) : 50;
  const int soft = (p.axis == AXIS_Y) ? (int)((long long)budget * soft_pct / 100) : budget;  // rows: room for 2 blocks per SM
  for (int pass = 0; pass < 2 && !bestT; ++pass)
    for (int c = 0; c < ncand && !bestT; ++c) {
      const int T = cand[c];
      if (T > 1 && (p.nunits + T - 1) / T < 1) continue;
      for (int wbT = wbcap_ ? T : 0; wbT >= (wbcap_ ? 1 : 0); wbT = (wbT > 1 ? wbT / 2 : wbT - 1)) {
        const long long scr = (long long)T * 32 + 160;  // DCT partial sums + warp totals of the scans
        const long long bytes = 16LL * ((long long)p.nreg * T * capP + (long long)wbT * wbP + scr);
        if (bytes <= (pass == 0 ? soft : budget)) {
          bestT = T;
          bestWb = wbT;
          bestSmem = (int)bytes;
          break;
        }
        if (wbT <= 1) break;
      }
    }
  if (!bestT) throw Error(RP_ERR_SHAPE, "lane too long for shared memory (single-block lane kernels)");
  p.T = bestT;
  p.wb_cap = wbcap_;
  p.wb_T = bestWb;
  p.smem_bytes = bestSmem;
  int nthr = std::min(512, std::max(64, pow2_ceil((fftlen_ + 7) / 8)));
  const int want = pow2_ceil(std::max(1, bestT * p.cap / 8));
  while (nthr < 256 && nthr < want) nthr <<= 1;
  if (const char* e_tm = getenv("RUSTPDE_B200_TMULT")) nthr = std::min(512, nthr * atoi(e_tm));
  while (nthr < 64 * bestT) nthr <<= 1;  // scans: >= one warp per parity chain
  (void)nck;
  if (nthr > 512) throw Error(RP_ERR_SHAPE, "lane too long for a 512-thread block");
  p.nthreads = nthr;
  Built b;
  p.prof = nullptr;
  if (getenv("RUSTPDE_B200_OPPROF")) {
    b.dprof = DevBuf(64 * sizeof(long long));
    p.prof = b.dprof.as<long long>();
  }
  b.dprog = upload_struct(p);
  b.nblocks = (p.nunits + p.T - 1) / p.T;
  b.nthreads = nthr;
  b.smem = bestSmem;
  b.bytes = bytes_unit_ * (double)p.nunits;
  b.valid = true;
  return b;
}

}  // namespace rp
