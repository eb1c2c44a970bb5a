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
  __device__ __forceinline__ double operator()(double v) const {
    const int h = __double2hiint(v);
    // exponent field of C: max(e, -126) + 29; the low word of C is zero
    const unsigned ce = max((unsigned)h & 0x7FF00000u, 0x38100000u) + 0x01D00000u;
    const double c = __hiloint2double((int)ce, 0);
    const double r = (fabs(v) + c) - c;  // r >= +0: its sign bit is clear
    return __hiloint2double(__double2hiint(r) | (h & (int)0x80000000u), __double2loint(r));
  }
};
struct ExactRound {
  __device__ __forceinline__ double operator()(double v) const { return (double)(float)v; }
};
// high word of 2^100: transforms whose inputs stay below it cannot reach 2^127 internally
constexpr unsigned kFastRoundInputLimit = 0x46300000u;
__device__ __forceinline__ unsigned abs_hi_word(double v) { return (unsigned)__double2hiint(v) & 0x7FFFFFFFu; }

struct Cplx {
  double re, im;
};

// fft.js:46-60 on (even, odd) = (a, b) with twiddle w
template <typename R>
__device__ __forceinline__ void butterfly(Cplx &a, Cplx &b, const double2 w, const R &rnd) {
  const double tr = b.re * w.x - b.im * w.y;
  const double ti = b.re * w.y + b.im * w.x;
  const double er = a.re, ei = a.im;
  a.re = rnd(er + tr);
  a.im = rnd(ei + ti);
  b.re = rnd(er - tr);
  b.im = rnd(ei - ti);
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
                                              const R &rnd) {
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
                                                 const double2 *__restrict__ tw, int lane, const R &rnd) {
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
// FFT64 uses 8 threads, FFT128 16: one warp transforms the low, mid and high band of a
// sound unit at once (lanes 0-7, 8-15, 16-31).
//
// Transpose buffer: complex position p lives at 16-byte slot p + (p >> 3); every quarter-warp
// access of the three layouts above then touches 8 distinct 16-byte bank groups.
// ------------------------------------------------------------------------------------
static __constant__ double2 c_fft_tw[255];  // DevTables::fft_tw, one copy per translation unit

constexpr int kXposeSlots64 = 64 + 8, kXposeSlots128 = 128 + 16;
__device__ __forceinline__ int xpose_slot(int p) { return p + (p >> 3); }

template <typename R>
__device__ __forceinline__ void fft8_pass_a(Cplx (&v)[8], const R &rnd) {
#pragma unroll
  for (int s = 0; s < 3; s++) {
    const int d = 1 << s;
#pragma unroll
    for (int j = 0; j < 8; j++)
      if (!(j & d)) butterfly(v[j], v[j + d], c_fft_tw[d - 1 + (j & (d - 1))], rnd);
  }
}

// v[m] = position base + 8m + u; stage 3 + ls pairs m with m + 2^ls, twiddle index
// (pos & (half - 1)) = 8 (m & (2^ls - 1)) + u with half = 8 * 2^ls
template <typename R>
__device__ __forceinline__ void fft8_pass_b(Cplx (&v)[8], int u, const double2 *__restrict__ tw, const R &rnd) {
#pragma unroll
  for (int ls = 0; ls < 3; ls++) {
    const int d = 1 << ls, half = 8 << ls;
#pragma unroll
    for (int m = 0; m < 8; m++)
      if (!(m & d)) butterfly(v[m], v[m + d], __ldg(&tw[half - 1 + 8 * (m & (d - 1)) + u]), rnd);
  }
}

// v[k] = position 8 (4h + k) + u, v[4 + k] = that + 64: stage 6, twiddle index 8 (4h + k) + u
template <typename R>
__device__ __forceinline__ void fft8_pass_c(Cplx (&v)[8], int u, int h, const double2 *__restrict__ tw, const R &rnd) {
#pragma unroll
  for (int k = 0; k < 4; k++) butterfly(v[k], v[4 + k], __ldg(&tw[63 + 8 * (4 * h + k) + u]), rnd);
}

__device__ __forceinline__ void xpose_put(double2 *buf, int p, const Cplx &z) { buf[xpose_slot(p)] = make_double2(z.re, z.im); }
__device__ __forceinline__ Cplx xpose_get(const double2 *buf, int p) {
  const double2 t = buf[xpose_slot(p)];
  Cplx z;
  z.re = t.x;
  z.im = t.y;
  return z;
}

// Lane geometry of the three concurrent long-block transforms of one sound unit.
struct LongLanes {
  int band;      // 0 low, 1 mid, 2 high
  int t;         // thread index inside the band's transform (0..7 or 0..15)
  int nl;        // threads of the transform = FFT size / 8
  int rev_t;     // bit reversal of t over log2(nl) bits
  int u, b6;     // pass B: positions 64 b6 + 8m + u
  __device__ __forceinline__ explicit LongLanes(int lane) {
    band = lane < 8 ? 0 : (lane < 16 ? 1 : 2);
    t = band == 2 ? lane - 16 : (lane & 7);
    nl = band == 2 ? 16 : 8;
    rev_t = band == 2 ? brev_bits(t, 4) : brev_bits(t, 3);
    u = t & 7;
    b6 = t >> 3;
  }
  // natural FFT-input index held at array position 8t + j (bit-reversed order, fft.js:21-32)
  __device__ __forceinline__ int q_of(int j) const { return (int)(__brev((unsigned)j) >> 29) * nl + rev_t; }
};

// Runs passes A..C on v (pass-A layout on entry: v[j] = position 8t + j); on exit v[k] holds the
// natural-order output whose index is out_index(k).  xbuf: this transform's transpose buffer.
template <typename R>
__device__ __forceinline__ void fft_long_inthread(Cplx (&v)[8], const LongLanes &G, double2 *xbuf,
                                                  const double2 *__restrict__ tw, bool active, const R &rnd) {
  fft8_pass_a(v, rnd);
  __syncwarp();
  if (active) {
#pragma unroll
    for (int j = 0; j < 8; j++) xpose_put(xbuf, 8 * G.t + j, v[j]);
  }
  __syncwarp();
  if (active) {
#pragma unroll
    for (int m = 0; m < 8; m++) v[m] = xpose_get(xbuf, 64 * G.b6 + 8 * m + G.u);
  }
  fft8_pass_b(v, G.u, tw, rnd);
  if (G.band == 2) {  // lanes 16-31: one more stage
    __syncwarp(0xffff0000u);
    if (active) {
#pragma unroll
      for (int m = 0; m < 8; m++) xpose_put(xbuf, 64 * G.b6 + 8 * m + G.u, v[m]);
    }
    __syncwarp(0xffff0000u);
    if (active) {
#pragma unroll
      for (int k = 0; k < 4; k++) {
        v[k] = xpose_get(xbuf, 8 * (4 * G.b6 + k) + G.u);
        v[4 + k] = xpose_get(xbuf, 64 + 8 * (4 * G.b6 + k) + G.u);
      }
    }
    fft8_pass_c(v, G.u, G.b6, tw, rnd);
  }
}
// natural-order output index of v[k] after fft_long_inthread
__device__ __forceinline__ int long_out_index(const LongLanes &G, int k) {
  if (G.band == 2) return 8 * (4 * G.b6 + (k & 3)) + G.u + 64 * (k >> 2);
  return 8 * k + G.u;
}

}  // namespace c1
