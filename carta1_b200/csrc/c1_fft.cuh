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
    const unsigned h = (unsigned)__double2hiint(v);
    const unsigned a = h & 0x7FFFFFFFu;
    const unsigned ce = (max(a, 0x38100000u) & 0x7FF00000u) + 0x01D00000u;
    const double c = __hiloint2double((int)ce, 0);
    const double r = (__hiloint2double((int)a, __double2loint(v)) + c) - c;
    return __hiloint2double((int)((unsigned)__double2hiint(r) | (h & 0x80000000u)), __double2loint(r));
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

}  // namespace c1
