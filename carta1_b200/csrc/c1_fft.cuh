// c1_fft.cuh -- warp-level radix-2 DIT FFT in registers, bit-exact with the reference's
// in-place Float32Array FFT (codec/transforms/fft.js:14-68).
//
// The reference stores every butterfly output to a Float32Array; all arithmetic between
// stores is binary64.  Here every value stays in a binary64 register that holds an exactly
// f32-representable number: rnd32() performs the Float32Array store's rounding inside the
// FP64 pipe.  B200 converts f64<->f32 at 16 lanes/clk/SM (quarter of the DADD rate, measured
// in profiles/r01_ubench_fp64_pipe.txt), so a cvt round trip per store would make the
// conversion pipe the bottleneck; the add/subtract of a power of two below costs two DADD.
//
// Butterflies are lane-local.  Before stage s (pairs at distance 2^s) lane b owns the two
// array positions whose index is b with a bit inserted at position s; after the stage the
// lanes b and b ^ 2^s swap one value each to set up stage s+1.
#pragma once

#include "c1_common.cuh"

namespace c1 {

// Float32Array store of a binary64 value, result widened back to binary64 (exact).
//
// FastRound: |v| + C - C, with C = 2^(max(e, -126) + 29) and e the exponent of v, rounds |v|
// to the binary32 grid with round-to-nearest-even: C's ulp is the f32 ulp of v's binade
// (2^-149 for everything below 2^-126, so f32 subnormals and zero come out right too) and the
// sum stays inside C's binade; the sign of v is put back afterwards, which also keeps -0.
// The only inputs it cannot round are |v| >= 2^127 (f32 overflow rounding), infinities and
// NaN.  Callers bound the magnitude of a transform's inputs up front (kFastRoundInputLimit:
// every transform here has a gain below 2^16) and use ExactRound, the cvt round trip,
// otherwise.  A per-value branch would be if-converted by ptxas into "always pay both
// conversions", which is the cost this avoids.
struct FastRound {
  // Carriers for C: 64-bit registers whose low word stays zero for the whole kernel; a rounding
  // only rewrites the high word (no per-rounding move to zero a fresh low word).  Two of them so
  // that the roundings of a butterfly do not serialise on one register.
  double cz0 = 0.0, cz2 = 0.0;
  int clamp;  // exponent field of 2^(-126 + 29), kept in a register so add + max fuse (VIADDMNMX)
  __device__ __forceinline__ FastRound() { asm("mov.u32 %0, 0x39E00000;" : "=r"(clamp)); }
  __device__ __forceinline__ double round_with(double v, double &cz) const {
    const int h = __double2hiint(v);
    const int ce = max((h & 0x7FF00000) + 0x01D00000, clamp);  // exponent field of C: max(e, -126) + 29
    asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {lo, %1};\n\t}" : "+d"(cz) : "r"(ce));
    const double r = (fabs(v) + cz) - cz;  // r >= +0
    return copysign(r, v);
  }
  __device__ __forceinline__ double r0(double v) { return round_with(v, cz0); }
  // Two roundings in four go through the conversion pipe instead (cvt.rn.f32.f64 + cvt.f64.f32 is
  // the rounding by definition): two issue slots instead of eight issue clocks each, and the XU
  // pipe (one conversion per 8 clocks per sub-partition) has room for half of the roundings --
  // measured: 0 of 4 -> 1.67 ms, 1 -> 1.61, 2 -> 1.57, 3 -> 1.59 (MDCT kernels, 1 h stereo).
  __device__ __forceinline__ double r1(double v) { return (double)(float)v; }
  __device__ __forceinline__ double r2(double v) { return round_with(v, cz2); }
  __device__ __forceinline__ double r3(double v) { return (double)(float)v; }
};
struct ExactRound {
  __device__ __forceinline__ double operator()(double v) const { return (double)(float)v; }
  __device__ __forceinline__ double r0(double v) const { return (double)(float)v; }
  __device__ __forceinline__ double r1(double v) const { return (double)(float)v; }
  __device__ __forceinline__ double r2(double v) const { return (double)(float)v; }
  __device__ __forceinline__ double r3(double v) const { return (double)(float)v; }
};
// high word of 2^100: transforms whose inputs stay below it cannot reach 2^127 internally
constexpr unsigned kFastRoundInputLimit = 0x46300000u;
__device__ __forceinline__ unsigned abs_hi_word(double v) { return (unsigned)__double2hiint(v) & 0x7FFFFFFFu; }

struct Cplx {
  double re, im;
};

// fft.js:46-60 on (even, odd) = (a, b) with twiddle w
template <typename R>
__device__ __forceinline__ void butterfly(Cplx &a, Cplx &b, const double2 w, R &rnd) {
  const double tr = b.re * w.x - b.im * w.y;
  const double ti = b.re * w.y + b.im * w.x;
  const double er = a.re, ei = a.im;
  a.re = rnd.r0(er + tr);
  a.im = rnd.r1(ei + ti);
  b.re = rnd.r2(er - tr);
  b.im = rnd.r3(ei - ti);
}

// Lanes l and l ^ h re-pair their values: the lane with bit h clear keeps a and receives the
// partner's a as its new b; the lane with bit h set keeps b and receives the partner's b as
// its new a.
__device__ __forceinline__ void repair(Cplx &a, Cplx &b, int h, int lane) {
  const bool up = (lane & h) != 0;
  const double sr = up ? a.re : b.re;
  const double si = up ? a.im : b.im;
  const double rr = __shfl_xor_sync(0xffffffffu, sr, h);
  const double ri = __shfl_xor_sync(0xffffffffu, si, h);
  if (up) { a.re = rr; a.im = ri; } else { b.re = rr; b.im = ri; }
}

// kLaneBits = 3: four independent 16-point FFTs per warp (lane groups of 8)
// kLaneBits = 5: one 64-point FFT per warp
// On entry lane g (index within its group) holds, in bit-reversed array order, positions
// 2g (a) and 2g+1 (b), i.e. natural input indices brev(g) and brev(g) + N/2.  On exit a and
// b hold natural output indices g and g + N/2.
template <int kLaneBits, typename R>
__device__ __forceinline__ void warp_fft_regs(Cplx &a, Cplx &b, const double2 *__restrict__ tw, int lane,
                                              R &rnd) {
  const int g = lane & ((1 << kLaneBits) - 1);
#pragma unroll
  for (int s = 0; s <= kLaneBits; s++) {
    const int h = 1 << s;
    butterfly(a, b, __ldg(&tw[h - 1 + (g & (h - 1))]), rnd);
    if (s < kLaneBits) repair(a, b, h, lane);
  }
}

// One 128-point FFT per warp: rows 0/1 are the lower/upper half of the bit-reversed array.
// Row r holds positions 64r + 2l (a), 64r + 2l + 1 (b) = natural inputs 64j + 2*brev5(l) + r.
// On exit a0, b0, a1, b1 hold natural outputs l, l+32, l+64, l+96.
template <typename R>
__device__ __forceinline__ void warp_fft128_regs(Cplx &a0, Cplx &b0, Cplx &a1, Cplx &b1,
                                                 const double2 *__restrict__ tw, int lane, R &rnd) {
#pragma unroll
  for (int s = 0; s <= 5; s++) {
    const int h = 1 << s;
    const double2 w = __ldg(&tw[h - 1 + (lane & (h - 1))]);
    butterfly(a0, b0, w, rnd);
    butterfly(a1, b1, w, rnd);
    if (s < 5) { repair(a0, b0, h, lane); repair(a1, b1, h, lane); }
  }
  butterfly(a0, a1, __ldg(&tw[63 + lane]), rnd);
  butterfly(b0, b1, __ldg(&tw[63 + 32 + lane]), rnd);
}

__device__ __forceinline__ int brev_bits(int x, int bits) { return (int)(__brev((unsigned)x) >> (32 - bits)); }

// ------------------------------------------------------------------------------------
// In-thread formulation for the long blocks (FFT64 / FFT128): every thread owns 8 complex
// values and runs three radix-2 stages on them without talking to anybody; between such
// passes the values are transposed through shared memory.  The butterflies, their inputs
// and their order of roundings are exactly the reference's (fft.js:35-66) -- only which
// thread executes which butterfly changes.
//
//   pass A: thread t owns array positions 8t + j           -> stages 0,1,2 (twiddles are
//           compile-time entries of the recurrence table: constant-bank operands)
//   pass B: thread (b6, u) owns positions 64 b6 + 8m + u   -> stages 3,4,5
//   pass C: (FFT128 only) thread (u, h) owns positions 8m + u and 64 + 8m + u, m in 4h..4h+3
//           -> stage 6
// A warp is specialised by ROLE so that all of this geometry is compile-time and every
// address is one per-lane base plus an immediate:
//   role 0 (kFft = 64):  8 threads per transform, 4 transforms per warp (low and mid band of two
//                        consecutive sound units)
//   role 1 (kFft = 128): 16 threads per transform, 2 transforms per warp (high band of the two)
//
// Transpose buffer: complex position p lives at 16-byte slot p + (p >> 3); every quarter-warp
// access of the three layouts above then touches 8 distinct 16-byte bank groups.
// ------------------------------------------------------------------------------------
// slot of (cos, sin) pair s in the padded pre/post-twiddle tables of the MDCT / IMDCT kernels: one empty slot after
// every eight, so that the bit-reversed lanes of a quarter-warp touch 8 different 16-byte bank groups
__host__ __device__ constexpr int pad8(int s) { return s + (s >> 3); }

static __constant__ double2 c_fft_tw[255];  // DevTables::fft_tw, one copy per translation unit

template <typename R>
__device__ __forceinline__ void fft8_pass_a(Cplx (&v)[8], R &rnd) {
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int d = 1 << s;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (!(j & d)) butterfly(v[j], v[j + d], c_fft_tw[d - 1 + (j & (d - 1))], rnd);
  }
}

// v[m] = position base + 8m + u; stage 3 + ls pairs m with m + 2^ls, twiddle index
// (pos & (half - 1)) = 8 (m & (2^ls - 1)) + u with half = 8 * 2^ls.  tw_u = table + u.
template <typename R>
__device__ __forceinline__ void fft8_pass_b(Cplx (&v)[8], const double2 *__restrict__ tw_u, R &rnd) {
#pragma unroll
  for (int ls = 0; ls < 3; ls++) {
    const int d = 1 << ls, half = 8 << ls;
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (!(m & d)) butterfly(v[m], v[m + d], tw_u[half - 1 + 8 * (m & (d - 1))], rnd);
  }
}

// v[k] = position 8 (4h + k) + u, v[4 + k] = that + 64: stage 6, twiddle index 8 (4h + k) + u.
// tw_uh = table + 32 h + u.
template <typename R>
__device__ __forceinline__ void fft8_pass_c(Cplx (&v)[8], const double2 *__restrict__ tw_uh, R &rnd) {
#pragma unroll
  for (int k = 0; k < 4; k++) butterfly(v[k], v[4 + k], tw_uh[63 + 8 * k], rnd);
}

template <int kRole>
struct LongGeom {
  static constexpr int kFft = kRole == 0 ? 64 : 128;    // complex FFT size
  static constexpr int kN = 4 * kFft;                   // MDCT size (256 / 512)
  static constexpr int kSize = kN / 2;                  // band samples per frame (128 / 256)
  static constexpr int kLanes = kFft / 8;               // threads per transform
  static constexpr int kPerWarp = 32 / kLanes;          // transforms per warp (4 / 2)
  static constexpr int kSlots = kFft + kFft / 8;        // 16-byte slots of the transpose buffer
  int x;       // transform index inside the warp
  int t;       // thread index inside the transform
  int rev_t;   // bit reversal of t
  int u, b6;   // pass B: positions 64 b6 + 8m + u
  __device__ __forceinline__ explicit LongGeom(int lane) {
    x = lane / kLanes;
    t = lane % kLanes;
    rev_t = (int)(__brev((unsigned)t) >> (kRole == 0 ? 29 : 28));
    u = t & 7;
    b6 = t >> 3;
  }
  // natural FFT-input index held at array position 8t + j (bit-reversed order, fft.js:21-32):
  // q = brev3(j) * kLanes + rev_t
  static __device__ __forceinline__ constexpr int q_step(int j) {
    return (((j & 1) << 2) | (j & 2) | ((j >> 2) & 1)) * kLanes;
  }
  // natural-order output index of v[k] after fft_long_inthread, minus the per-lane part
  // out_base() = u (role 0) or 32 b6 + u (role 1)
  __device__ __forceinline__ int out_base() const { return kRole == 0 ? u : 32 * b6 + u; }
  static __device__ __forceinline__ constexpr int out_step(int k) {
    return kRole == 0 ? 8 * k : 8 * (k & 3) + 64 * (k >> 2);
  }
};

// Runs passes A..C on v (pass-A layout on entry: v[j] = position 8t + j); on exit v[k] holds the
// natural-order output of index out_base() + out_step(k).  xbuf: this transform's transpose
// buffer (LongGeom::kSlots double2).  Every lane runs every instruction.
template <int kRole, typename R>
__device__ __forceinline__ void fft_long_inthread(Cplx (&v)[8], const LongGeom<kRole> &G, double2 *xbuf,
                                                  const double2 *__restrict__ tw, R &rnd) {
  fft8_pass_a(v, rnd);
  __syncwarp();
  double2 *put_a = xbuf + 9 * G.t;  // slot(8t + j) = 9t + j
#pragma unroll
  for (int j = 0; j < 8; j++) put_a[j] = make_double2(v[j].re, v[j].im);
  __syncwarp();
  double2 *at_b = xbuf + 72 * G.b6 + G.u;  // slot(64 b6 + 8m + u) = 72 b6 + 9m + u
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const double2 z = at_b[9 * m];
    v[m].re = z.x;
    v[m].im = z.y;
  }
  fft8_pass_b(v, tw + G.u, rnd);
  if (kRole == 1) {
    __syncwarp();
#pragma unroll
    for (int m = 0; m < 8; m++) at_b[9 * m] = make_double2(v[m].re, v[m].im);
    __syncwarp();
    const double2 *at_c = xbuf + 36 * G.b6 + G.u;  // slot(8 (4h + k) + u) = 36 h + 9k + u; +64 -> +72
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 lo = at_c[9 * k], hi = at_c[9 * k + 72];
      v[k].re = lo.x;
      v[k].im = lo.y;
      v[4 + k].re = hi.x;
      v[4 + k].im = hi.y;
    }
    fft8_pass_c(v, tw + 32 * G.b6 + G.u, rnd);
  }
}

}  // namespace c1
