// c1_decode.cu -- ATRAC1 decode kernels for sm_100a.
//
//   K5 unpack_dequant_kernel  212-byte unit -> 512 coefficients + block modes
//                             (serialization.js:111-176, decoder.js:52-98, quantization.js:65-78)
//   K6 imdct_kernel           coefficients -> per-band IMDCT middle halves
//                             (mdct.js:139-211, decoder.js:175-306 up to the overlap-add)
//   K7 synth_kernel           overlap-add (mdct.js:230-245) + two-stage QMF synthesis
//                             (qmf.js:60-105, decoder.js:360-388) -> PCM (f32 or WAV int16)
// Frame f of a row depends on units f-1 and f only: every kernel
// treats the row as starting from the decoder's zero state and K7 recomputes the halo.
#include "c1_common.cuh"
#include "c1_fft.cuh"
#include "c1_launch.h"

#include <algorithm>

namespace c1 {

// ------------------------------------------------------------------------------------
// K5+K6 fused: one warp per sound unit.  The unit is unpacked and dequantised into a
// shared-memory row of 512 coefficients, each band is inverse-transformed in registers
// (c1_fft.cuh) and the IMDCT middle halves replace the coefficients in the same row.
// Only inv[N/4 .. N/4 + N/2) is ever used by the decoder (decoder.js:186-194,268-276).
// ------------------------------------------------------------------------------------
// unpackBits (bitstream.js:48-69) on big-endian words: a read that runs past byte 212
// returns only the bits that were there, unshifted.
__device__ __forceinline__ uint32_t get_bits(const uint32_t *words, int pos, int bits) {
  const int avail = kFrameBits - pos;
  if (avail <= 0) return 0;
  const int nb = bits < avail ? bits : avail;
  const int w = pos >> 5, off = pos & 31;
  const unsigned long long v = ((unsigned long long)words[w] << 32) | words[w + 1];
  return (uint32_t)((v << off) >> (64 - nb));
}

// f32((q * SF) / R) (quantization.js:75).  The division by the small integer R uses the
// correctly rounded reciprocal and one residual correction (Markstein): q0 = x*rcp,
// rem = x - q0*R exactly (fma), result = RN(q0 + rem*rcp) == RN(x / R).  R = 2^k - 1 is never
// the all-ones significand the theorem excludes; carta1_debug_selftest checks every
// (q, R, SF) against the IEEE division.
__device__ __forceinline__ double div_by_range(double x, double range, double rcp) {
  const double q0 = x * rcp;
  const double rem = fma(-q0, range, x);
  return fma(rem, rcp, q0);
}
__device__ __forceinline__ double int_to_double(int q) {  // exact, no conversion pipe
  return __hiloint2double(0x43300000, (int)((unsigned)q ^ 0x80000000u)) - 4503601774854144.0;
}

// Frame objects the 212-byte layout cannot hold (decode() accepts any object with the frame
// keys, e.g. tests/decoder.test.js:70-84): the host replays dequantizationStage's
// `coefficients.set(dequantized, position)` sequence (decoder.js:73-94) into per-position
// arrays, [unit][512] each, and the kernel dequantises per position.  bits == 0: position
// never written (stays +0).
struct ExpandedFrames {
  const int32_t *q;
  const uint8_t *sfi, *bits;
  const uint8_t *modes;  // [unit][4]
};

// mdct.js:161-170: pre-twiddle of FFT input i (natural order)
template <typename In, typename R>
__device__ __forceinline__ Cplx imdct_pre(int i, int n, const In &in, const double *__restrict__ tab,
                                          R &rnd) {
  const int i2 = 2 * i, half = n >> 1;
  const double r = -in(i2);
  const double m = -in(half - 1 - i2);
  const double c = __ldg(&tab[i2]), s = __ldg(&tab[i2 + 1]);
  Cplx z;
  z.re = rnd.r0(m * s + r * c);
  z.im = rnd.r1(m * c - r * s);
  return z;
}

// mdct.js:177-208 restricted to output[n4 .. 3*n4): out[] is that middle half
__device__ __forceinline__ void imdct_post(const Cplx z, int i, int n, const double *__restrict__ tab, float *out) {
  const int half = n >> 1, n4 = n >> 2, fft_n = n >> 2;
  const double2 cs = __ldg(reinterpret_cast<const double2 *>(tab + 2 * i));
  const double c = cs.x, s = cs.y;
  const float r1 = (float)(z.re * c + z.im * s);
  const float i1 = (float)(z.re * s - z.im * c);
  const int idx = i < (fft_n >> 1) ? 2 * i : (i - (fft_n >> 1)) * 2 + n4;
  out[half - 1 - idx] = r1;  // output[n34 - 1 - idx]
  out[idx] = i1;             // output[n4 + idx]
}

// One band of one sound unit, by one warp: x (coefficients) -> y (IMDCT middle half).
template <typename R>
__device__ __forceinline__ void imdct_band(int band, bool is_long, const float *x, float *y,
                                           const DevTables *__restrict__ T, int lane) {
  R rnd;
  const double2 *tw = T->fft_tw;
  const int size = band == 2 ? 256 : 128;
  const bool rev = band > 0;  // utils.js:42-48: un-reverse mid / high spectra
  if (is_long) {
    auto in = [&](int k) -> double { return (double)x[rev ? size - 1 - k : k]; };
    const int r5 = brev_bits(lane, 5);
    if (band < 2) {
      const double *tab = T->mdct_inv256;
      Cplx a = imdct_pre(r5, 256, in, tab, rnd);
      Cplx b = imdct_pre(32 + r5, 256, in, tab, rnd);
      warp_fft_regs<5>(a, b, tw, lane, rnd);
      imdct_post(a, lane, 256, tab, y);
      imdct_post(b, lane + 32, 256, tab, y);
    } else {
      const double *tab = T->mdct_inv512;
      Cplx a0 = imdct_pre(2 * r5, 512, in, tab, rnd);
      Cplx b0 = imdct_pre(64 + 2 * r5, 512, in, tab, rnd);
      Cplx a1 = imdct_pre(2 * r5 + 1, 512, in, tab, rnd);
      Cplx b1 = imdct_pre(64 + 2 * r5 + 1, 512, in, tab, rnd);
      warp_fft128_regs(a0, b0, a1, b1, tw, lane, rnd);
      imdct_post(a0, lane, 512, tab, y);
      imdct_post(b0, lane + 32, 512, tab, y);
      imdct_post(a1, lane + 64, 512, tab, y);
      imdct_post(b1, lane + 96, 512, tab, y);
    }
  } else {
    const double *tab = T->mdct_inv64;
    const int g = lane & 7, r3 = brev_bits(g, 3);
    for (int b0 = 0; b0 < (size >> 5); b0 += 4) {
      const int blk = b0 + (lane >> 3);
      const float *xb = x + 32 * blk;
      auto in = [&](int k) -> double { return (double)xb[rev ? 31 - k : k]; };
      Cplx a = imdct_pre(r3, 64, in, tab, rnd);
      Cplx b = imdct_pre(8 + r3, 64, in, tab, rnd);
      warp_fft_regs<3>(a, b, tw, lane, rnd);
      __syncwarp();  // y may alias x: every coefficient of these blocks is read by now
      imdct_post(a, g, 64, tab, y + 32 * blk);
      imdct_post(b, g + 8, 64, tab, y + 32 * blk);
    }
  }
}

__device__ __noinline__ void imdct_band_exact(int band, bool is_long, const float *x, float *y,
                                              const DevTables *__restrict__ T, int lane) {
  imdct_band<ExactRound>(band, is_long, x, y, T, lane);
}

// ------------------------------------------------------------------------------------
// K5: unpack + dequantise, one warp per sound unit -> 512 coefficients + block modes.
// ------------------------------------------------------------------------------------
constexpr int kUnpackWarps = 8;
struct UpBfu {          // per BFU of the unit being unpacked: what every coefficient reads (one 8-byte load)
  uint32_t base_bits;   // bit offset (malformed units: up to 16 + 520 + 16 * 512) | width << 16
  int32_t tab;          // first entry of (wl, sfi) in DevTables::deq_tab, or -1: width above 7 bits, or a unit that claims
                        // more bits than it has (then every BFU takes the general path, see `overrun`)
};
struct UpBfuWide {      // the same BFU for codes of 8 bits and more (the division path)
  double sf;            // SCALE_FACTORS[sfi]; 0 when sfi == 0 (dequantize returns zeros)
  double rcp;           // 1 / quantRange
  double range;         // quantRange = 2^(bits-1) - 1
};
struct UnpackWarpSmem {
  UpBfu bfu[52];
  UpBfuWide wide[52];
  uint32_t words[56];
};

// modes[unit * 4 + 0..2]: 1 = short blocks; modes[unit * 4 + 3]: 1 = the unit's band record is
// already in place (stateful handles: frame 0 of every row is the record kept from the
// previous call), the IMDCT kernels skip it.
__global__ void __launch_bounds__(kUnpackWarps * 32, 6)
unpack_kernel(const uint8_t *__restrict__ su, size_t su_frame_stride, size_t su_stream_stride,
              long long n_su_valid, int frames, int n_units, const DevTables *__restrict__ T,
              float *__restrict__ coefs, uint8_t *__restrict__ modes, float *__restrict__ inv,
              const float *__restrict__ prev_rec, ExpandedFrames xf) {
  __shared__ __align__(16) UnpackWarpSmem s_all[kUnpackWarps];
  __shared__ uint16_t s_bj[2][512];  // long, short: position -> (BFU << 5) | index inside the BFU
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const FormatTables &F = T->fmt;
  for (int i = tid; i < 512; i += kUnpackWarps * 32) { s_bj[0][i] = F.bj_long[i]; s_bj[1][i] = F.bj_short[i]; }
  __syncthreads();
  UnpackWarpSmem &S = s_all[warp];
  uint32_t *words = S.words;
  const int sz0 = F.specs[lane], sz1 = lane < 20 ? F.specs[lane + 32] : 0;
  const int su_skip = prev_rec ? 1 : 0;
  for (int unit = blockIdx.x * kUnpackWarps + warp; unit < n_units; unit += gridDim.x * kUnpackWarps) {
    const int stream = unit / frames, frame = unit - stream * frames;
    float *dst = coefs + (size_t)unit * 512;
    if (prev_rec && frame == 0) {
      const float4 *s4 = reinterpret_cast<const float4 *>(prev_rec + (size_t)stream * 512);
      float4 *dst4 = reinterpret_cast<float4 *>(inv + (size_t)unit * 512);
#pragma unroll
      for (int k = 0; k < 4; k++) dst4[lane + 32 * k] = s4[lane + 32 * k];
      if (lane < 4) modes[(size_t)unit * 4 + lane] = lane == 3 ? 1 : 0;
      continue;
    }
    const long long lin = (long long)(frame - su_skip) * (long long)su_frame_stride +
                          (long long)stream * (long long)su_stream_stride;
    int short_mask = 0;  // bit b: band b uses short blocks (any non-zero mode, decoder.js:82-83)
    if (xf.q) {
      // frame objects in position-expanded form (see ExpandedFrames): the general decode() input
      const size_t xu = (size_t)stream * (frames - su_skip) + (frame - su_skip);
      const int m0 = xf.modes[xu * 4], m1 = xf.modes[xu * 4 + 1], m2 = xf.modes[xu * 4 + 2];
      short_mask = (m0 != 0) | ((m1 != 0) << 1) | ((m2 != 0) << 2);
      for (int k = 0; k < 16; k++) {
        const int c = lane + 32 * k;
        const int bits = xf.bits[xu * 512 + c], sfi = xf.sfi[xu * 512 + c];
        float val = 0.0f;
        if (bits > 0 && sfi > 0) {  // quantization.js:65-78, the IEEE division itself
          const double range = (double)((1 << (bits - 1)) - 1);
          val = (float)(((double)xf.q[xu * 512 + c] * T->sf[sfi]) / range);
        }
        dst[c] = val;
      }
    } else if (lin >= n_su_valid) {  // dummy frame {nBfu: 0, blockModes: [0,0,0]} (processor.js:299-307)
      // all-zero coefficients: every IMDCT output is +0 or -0; the transform runs for the signs
      for (int k = 0; k < 16; k++) dst[lane + 32 * k] = 0.0f;
    } else {
      const uint32_t *src = reinterpret_cast<const uint32_t *>(su + (size_t)lin * kSuBytes);
      const uint32_t w0 = __ldg(src + lane), w1 = lane < kSuWords - 32 ? __ldg(src + 32 + lane) : 0u;
      __syncwarp();  // the previous unit's reads of S are done
      words[lane] = __byte_perm(w0, 0, 0x0123);
      if (lane < 24) words[32 + lane] = __byte_perm(w1, 0, 0x0123);
      __syncwarp();
      const uint32_t header = words[0] >> 16;  // serialization.js:118-126
      const int m0 = 2 - (int)((header >> 14) & 3), m1 = 2 - (int)((header >> 12) & 3),
                m2 = 3 - (int)((header >> 10) & 3);
      const int idx = (header >> 5) & 7;
      const int n = idx == 0 ? 20 : 24 + 4 * idx;  // BFU_AMOUNTS
      int run = 16 + 10 * n;
#pragma unroll
      for (int h = 0; h < 2; h++) {  // word lengths, scale factors, bit offsets (exclusive scan)
        const int b = lane + 32 * h;
        int wl = 0, sfi = 0;
        if (b < n) {
          wl = (int)get_bits(words, 16 + 4 * b, 4);
          sfi = (int)get_bits(words, 16 + 4 * n + 6 * b, 6);
        }
        const int bits = wl_bits(wl);
        const int cost = bits * (h == 0 ? sz0 : sz1);
        int incl = cost;
#pragma unroll
        for (int d = 1; d < 32; d <<= 1) {
          const int t = __shfl_up_sync(0xffffffffu, incl, d);
          if (lane >= d) incl += t;
        }
        if (b < 52) {
          UpBfu &r = S.bfu[b];
          UpBfuWide &rw = S.wide[b];
          const int range = (1 << wl) - 1;  // 2^(bits-1) - 1
          rw.sf = sfi ? __ldg(&T->sf[sfi]) : 0.0;
          rw.rcp = __ldg(&T->rcp_range[wl]);
          rw.range = (double)range;
          r.base_bits = (uint32_t)(run + incl - cost) | ((uint32_t)bits << 16);
          r.tab = wl >= 1 && wl <= kDeqMaxWl ? deq_off(wl) + (sfi << bits) : -1;
        }
        run += __shfl_sync(0xffffffffu, incl, 31);
      }
      __syncwarp();
      short_mask = (m0 != 0) | ((m1 != 0) << 1) | ((m2 != 0) << 2);
      // `run` is the end of the coefficient bits: a unit that claims more than the 1696 it has
      // (malformed input) needs unpackBits' behaviour at the end of the buffer (bitstream.js:55-68)
      const bool overrun = run > kFrameBits;
      if (overrun) {  // rare: no table look-ups for this unit, the coefficient loop then tests one field only
        if (lane < 26) { S.bfu[lane].tab = -1; S.bfu[lane + 26].tab = -1; }
        __syncwarp();
      }
      const uint16_t *bj0 = s_bj[m0 != 0] + lane, *bj1 = s_bj[m1 != 0] + lane, *bj2 = s_bj[m2 != 0] + lane;
#pragma unroll
      for (int k = 0; k < 16; k++) {  // serialization.js:153-166 + decoder.js:65-94
        const uint32_t bj = (k < 4 ? bj0 : (k < 8 ? bj1 : bj2))[32 * k];
        const UpBfu r = S.bfu[bj >> 5];
        const int bits = (int)(r.base_bits >> 16);
        float val = 0.0f;
        if (bits > 0) {
          const int pos = (int)(r.base_bits & 0xFFFFu) + (int)(bj & 31u) * bits;
          if (r.tab >= 0) {
            // short codes (2..7 bits): f32((q * SF) / R) was tabulated on the host for every bit pattern
            const int w = pos >> 5, off = pos & 31;
            const uint32_t top = __funnelshift_l(words[w + 1], words[w], off);  // bits pos.. at the top
            val = __ldg(&T->deq_tab[r.tab + (int)(top >> (32 - bits))]);
          } else {
            int q;
            if (!overrun) {
              const int w = pos >> 5, off = pos & 31;
              const uint32_t top = __funnelshift_l(words[w + 1], words[w], off);
              q = (int)top >> (32 - bits);  // sign-extending (bitstream.js:78-82)
            } else {
              const int v = (int)get_bits(words, pos, bits);
              q = v >= (1 << (bits - 1)) ? v - (1 << bits) : v;
            }
            const UpBfuWide rw = S.wide[bj >> 5];
            if (rw.sf != 0.0) val = (float)div_by_range(int_to_double(q) * rw.sf, rw.range, rw.rcp);
          }
        }
        dst[lane + 32 * k] = val;
      }
    }
    if (lane < 4) modes[(size_t)unit * 4 + lane] = lane == 3 ? 0 : (uint8_t)((short_mask >> lane) & 1);
  }
}

// ------------------------------------------------------------------------------------
// deserializeFrame in batches (serialization.js:111-176; the frame dump of bin/cli.js:567-677 runs it over
// a whole file): one warp per sound unit -> BFU count, block modes, word-length and scale-factor indices
// and the quantised integers in bitstream order (BFU b's SPECS_PER_BFU[b] values at BFU_START_LONG[b],
// which is the running sum of the sizes).  Reads past byte 212 follow unpackBits (bitstream.js:55-68).
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(kUnpackWarps * 32)
deserialize_kernel(const uint8_t *__restrict__ su, int n_units, const DevTables *__restrict__ T,
                   uint8_t *__restrict__ n_bfu_out, int8_t *__restrict__ modes_out, uint8_t *__restrict__ wl_out,
                   uint8_t *__restrict__ sfi_out, int32_t *__restrict__ q_out) {
  __shared__ uint32_t s_words[kUnpackWarps][56];
  __shared__ uint32_t s_base[kUnpackWarps][52];  // bit offset | width << 16
  __shared__ uint16_t s_bj[512];                 // bitstream position -> (BFU << 5) | index inside the BFU
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const FormatTables &F = T->fmt;
  for (int i = tid; i < 512; i += kUnpackWarps * 32) s_bj[i] = F.bj_long[i];
  __syncthreads();
  uint32_t *words = s_words[warp], *base = s_base[warp];
  const int sz0 = F.specs[lane], sz1 = lane < 20 ? F.specs[lane + 32] : 0;
  for (int unit = blockIdx.x * kUnpackWarps + warp; unit < n_units; unit += gridDim.x * kUnpackWarps) {
    const uint32_t *src = reinterpret_cast<const uint32_t *>(su + (size_t)unit * kSuBytes);
    const uint32_t w0 = __ldg(src + lane), w1 = lane < kSuWords - 32 ? __ldg(src + 32 + lane) : 0u;
    __syncwarp();
    words[lane] = __byte_perm(w0, 0, 0x0123);
    if (lane < 24) words[32 + lane] = __byte_perm(w1, 0, 0x0123);
    __syncwarp();
    const uint32_t header = words[0] >> 16;
    const int idx = (header >> 5) & 7;
    const int n = idx == 0 ? 20 : 24 + 4 * idx;  // BFU_AMOUNTS
    if (lane == 0) {
      n_bfu_out[unit] = (uint8_t)n;
      modes_out[unit * 3 + 0] = (int8_t)(2 - (int)((header >> 14) & 3));
      modes_out[unit * 3 + 1] = (int8_t)(2 - (int)((header >> 12) & 3));
      modes_out[unit * 3 + 2] = (int8_t)(3 - (int)((header >> 10) & 3));
    }
    int run = 16 + 10 * n;
#pragma unroll
    for (int h = 0; h < 2; h++) {
      const int b = lane + 32 * h;
      int wl = 0, sfi = 0;
      if (b < n) {
        wl = (int)get_bits(words, 16 + 4 * b, 4);
        sfi = (int)get_bits(words, 16 + 4 * n + 6 * b, 6);
      }
      const int bits = wl_bits(wl);
      const int cost = bits * (h == 0 ? sz0 : sz1);
      int incl = cost;
#pragma unroll
      for (int d = 1; d < 32; d <<= 1) {
        const int t = __shfl_up_sync(0xffffffffu, incl, d);
        if (lane >= d) incl += t;
      }
      if (b < 52) {
        base[b] = (uint32_t)(run + incl - cost) | ((uint32_t)bits << 16);
        wl_out[(size_t)unit * 52 + b] = (uint8_t)wl;
        sfi_out[(size_t)unit * 52 + b] = (uint8_t)sfi;
      }
      run += __shfl_sync(0xffffffffu, incl, 31);
    }
    __syncwarp();
#pragma unroll 4
    for (int k = 0; k < 16; k++) {
      const uint32_t bj = s_bj[lane + 32 * k];
      const uint32_t r = base[bj >> 5];
      const int bits = (int)(r >> 16);
      int q = 0;
      if (bits > 0) {
        const int v = (int)get_bits(words, (int)(r & 0xFFFFu) + (int)(bj & 31u) * bits, bits);
        q = v >= (1 << (bits - 1)) ? v - (1 << bits) : v;  // bitstream.js:78-82
      }
      q_out[(size_t)unit * 512 + lane + 32 * k] = q;
    }
  }
}

cudaError_t launch_deserialize(const uint8_t *d_su, int n_units, const DevTables *tables, uint8_t *n_bfu, int8_t *modes,
                               uint8_t *wl, uint8_t *sfi, int32_t *q, cudaStream_t st, Prof *prof) {
  if (n_units <= 0) return cudaSuccess;
  deserialize_kernel<<<std::min((n_units + kUnpackWarps - 1) / kUnpackWarps, persistent_ctas(4)), kUnpackWarps * 32, 0, st>>>(
      d_su, n_units, tables, n_bfu, modes, wl, sfi, q);
  prof->launches++;
  return cudaGetLastError();
}

// ------------------------------------------------------------------------------------
// K6: IMDCT.  Long blocks use the in-thread passes of c1_fft.cuh.  A warp task is a pair of
// consecutive sound units and a ROLE: role 0 transforms the low and mid bands of both units
// (4 x IMDCT256), role 1 their high bands (2 x IMDCT512); rows holds the task's coefficients,
// [transform][kSize].
//
// The output goes straight into the band record K7 reads, per band
//   [head16 | tail16 | time-domain samples 32..size)
// Samples >= 32 depend on this unit only (decoder.js:203-226,244-300): in long mode they are the
// IMDCT middle half y shifted by 16 (record[p] = y[p - 16]), head16 = y[0..16) and
// tail16 = y[size-16..size) feed the overlap-add that K7 finishes with the previous unit's tail.
// The record overwrites the coefficients of the same band (every coefficient is read before
// the first transpose, every output written after the last).  Transforms whose bit in
// long_mask is clear run the same instructions and do not store.
// ------------------------------------------------------------------------------------
// Shared-memory rows (bank conflicts, profiles/r02_mdct_shared_conflicts.txt): the pre-twiddle reads x[2q] and
// x[size-1-2q], the post-twiddle writes record positions of one parity per store, so a row is kept de-interleaved:
// position p at plane p & 1 (even positions first, odd ones kSize / 2 further), index p >> 1, both for the
// coefficients read and for the record written over them.  The row stride spreads the transforms of a warp over
// the banks; the pre/post table holds (cos, sin) pair s at slot s + (s >> 3) (as in the MDCT kernels).
template <int kRole>
struct ImdctLayout {
  static constexpr int kSize = LongGeom<kRole>::kSize;
  static constexpr int kRow = kSize + (kRole == 0 ? 8 : 16);  // floats: 136 / 272
};

template <int kRole, typename R>
__device__ __forceinline__ void imdct_long_task(unsigned long_mask, float *rows, double2 *xbuf_all,
                                                const double2 *tab2, const double2 *tw, int lane) {
  using G = LongGeom<kRole>;
  constexpr int kSize = G::kSize;
  R rnd;
  const G g(lane);
  float *xb = rows + g.x * ImdctLayout<kRole>::kRow;
  const bool rev = kRole == 1 || (g.x & 1);  // utils.js:42-48: un-reverse mid / high spectra
  const float *pa = xb + g.rev_t, *pb = xb + (kSize - 1) - g.rev_t;  // x[2q] = E[q], x[size-1-2q] = O[size/2-1-q]
  const double2 *ptab = tab2 + pad8(g.rev_t);
  Cplx v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {  // mdct.js:161-170 for FFT input q = q_step(j) + rev_t
    const int q = G::q_step(j);
    const float fa = pa[q], fb = pb[-q];
    const double r = -(double)(rev ? fb : fa);
    const double m = -(double)(rev ? fa : fb);
    const double2 cs = ptab[pad8(q)];
    v[j].re = rnd.r0(m * cs.y + r * cs.x);
    v[j].im = rnd.r1(m * cs.x - r * cs.y);
  }
  fft_long_inthread<kRole>(v, g, xbuf_all + g.x * G::kSlots, tw, rnd);
  if ((long_mask >> g.x) & 1) {
    // mdct.js:177-208 restricted to output[n/4 .. 3n/4): FFT output i gives y[2i] and y[size-1-2i].
    // Record position of y[q]: q + 16, except q < 16 -> q (head) and q >= size - 16 -> q - size + 32
    // (tail).  2i < 16 only for k == 0 and 2i >= size - 16 only for k == 7 (role 1: on the lanes
    // with b6 == 0 / b6 == 1 respectively).  y[2i] lands on an even position P (stored at E[P / 2]),
    // y[size-1-2i] on an odd one (O[(P - 1) / 2]).
    const int ib = g.out_base();
    const double2 *pt = tab2 + pad8(ib);
    float *pe = xb + ib + 8, *po = xb + kSize + 7 - ib;
    const bool first = kRole == 0 || g.b6 == 0, last = kRole == 0 || g.b6 == 1;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int st = G::out_step(k);
      const double2 cs = pt[pad8(st)];
      const float y1 = (float)(v[k].re * cs.x + v[k].im * cs.y);  // y[size - 1 - 2i]
      const float y0 = (float)(v[k].re * cs.y - v[k].im * cs.x);  // y[2i]
      int d0 = st, d1 = -st;
      if (k == 0 && first) { d0 -= 8; d1 += 8 - kSize / 2; }  // y[2i] is head, y[size-1-2i] is tail
      if (k == 7 && last) { d0 += 8 - kSize / 2; d1 -= 8; }   // y[2i] is tail, y[size-1-2i] is head
      pe[d0] = y0;
      po[d1] = y1;
    }
  }
}

template <int kRole>
__device__ __noinline__ void imdct_long_task_exact(unsigned long_mask, float *rows, double2 *xbuf_all,
                                                   const double2 *tab2, const double2 *tw, int lane) {
  imdct_long_task<kRole, ExactRound>(long_mask, rows, xbuf_all, tab2, tw, lane);
}

// Short blocks of one band (rare: out of line): the coefficients are gathered from the de-interleaved row into
// natural order (scratch[0 .. size)), transformed in place, the record is assembled in scratch[256 ..)
// (block-to-block overlap-add, mdct.js:230-245 with prev = second half of the previous block) and scattered back
// into the row.
__device__ __noinline__ void imdct_short_band(int band, float *row, float *scratch, bool fast,
                                              const DevTables *__restrict__ T, int lane) {
  const int size = band == 2 ? 256 : 128;
  float *x = scratch, *stage = scratch + 256;
  for (int p = lane; p < size; p += 32) x[p] = row[(p & 1) * (size >> 1) + (p >> 1)];
  __syncwarp();
  if (fast) imdct_band<FastRound>(band, false, x, x, T, lane);
  else imdct_band<ExactRound>(band, false, x, x, T, lane);
  __syncwarp();
  const float *v = x;
  if (lane < 16) stage[lane] = v[lane];
  else stage[lane] = v[size - 32 + lane];
  for (int p = 32 + lane; p < size; p += 32) {
    const int q = p & 31, blk = p - q;
    const int i = q < 16 ? q : 31 - q;
    const double pv = (double)v[blk - 16 + i], cv = (double)v[blk + 15 - i];
    const double w1 = __ldg(&T->win[i]), w2 = __ldg(&T->win[31 - i]);
    stage[p] = q < 16 ? (float)(pv * w2 - cv * w1) : (float)(pv * w1 + cv * w2);
  }
  __syncwarp();
  for (int p = lane; p < size; p += 32) row[(p & 1) * (size >> 1) + (p >> 1)] = stage[p];
  __syncwarp();
}

#ifndef C1_IMDCT_WARPS
#define C1_IMDCT_WARPS 8
#define C1_IMDCT_CTAS 3
#endif
constexpr int kImdctWarps = C1_IMDCT_WARPS, kImdctCtasPerSm = C1_IMDCT_CTAS;
struct ImdctWarpSmem {
  double2 xbuf[4 * LongGeom<0>::kSlots];  // transposes of the long-block FFTs (also the short-block scratch)
  float rows[544];                         // role 0: 4 x 136, role 1: 2 x 272, de-interleaved (ImdctLayout)
};
static_assert(4 * LongGeom<0>::kSlots == 2 * LongGeom<1>::kSlots, "xbuf");
static_assert(4 * ImdctLayout<0>::kRow <= 544 && 2 * ImdctLayout<1>::kRow <= 544, "rows");
struct ImdctTables {  // staged once per CTA: pre/post table (N/4 pairs, pair s at slot pad8(s)), FFT twiddles of stages 3..
  double2 tab2[144];
  double2 tw[128];
};
constexpr size_t kImdctSmemBytes = sizeof(ImdctWarpSmem) * kImdctWarps + sizeof(ImdctTables);

// One kernel per role (see mdct_kernel): the hot loop stays inside the instruction cache.
template <int kRole>
__global__ void __launch_bounds__(kImdctWarps * 32, kImdctCtasPerSm)
imdct_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes, int n_units,
             const DevTables *__restrict__ T, float *__restrict__ inv) {
  using G = LongGeom<kRole>;
  constexpr int kSize = G::kSize, kPer = G::kPerWarp / 2;  // transforms per unit
  constexpr int kOff = kRole == 0 ? 0 : 256;              // first coefficient of the role
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  ImdctWarpSmem &S = reinterpret_cast<ImdctWarpSmem *>(smem_raw)[warp];
  ImdctTables &ST = *reinterpret_cast<ImdctTables *>(smem_raw + sizeof(ImdctWarpSmem) * kImdctWarps);
  {
    const double *tab = kRole == 0 ? T->mdct_inv256 : T->mdct_inv512;
    for (int i = threadIdx.x; i < G::kN / 4; i += blockDim.x) ST.tab2[pad8(i)] = make_double2(tab[2 * i], tab[2 * i + 1]);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ST.tw[i] = T->fft_tw[i];
  }
  __syncthreads();
  const int n_pairs = (n_units + 1) >> 1;
  for (int pair = blockIdx.x * kImdctWarps + warp; pair < n_pairs; pair += gridDim.x * kImdctWarps) {
    __syncwarp();
    {  // the next pair's 2 x 1 KB of coefficients (this role's half of each unit): 16 lines, one per lane
      const int next = pair + gridDim.x * kImdctWarps;
      const int un = 2 * next + (lane >> 3);
      if (lane < 16 && un < n_units) prefetch_l2(coefs + (size_t)un * 512 + kOff + 32 * (lane & 7));
    }
    const int u0 = 2 * pair;
    bool live[2];
    unsigned long_mask = 0, short_mask = 0;
#pragma unroll
    for (int s = 0; s < 2; s++) {
      live[s] = u0 + s < n_units;
      if (live[s]) {
        const uchar4 m = *reinterpret_cast<const uchar4 *>(modes + (size_t)(u0 + s) * 4);
        live[s] = m.w == 0;
        const unsigned sm = kRole == 0 ? (unsigned)(m.x != 0) | ((unsigned)(m.y != 0) << 1) : (unsigned)(m.z != 0);
        if (live[s]) {
          short_mask |= sm << (s * kPer);
          long_mask |= (~sm & ((1u << kPer) - 1u)) << (s * kPer);
        }
      }
    }
    if (!live[0] && !live[1]) continue;
    // 256 coefficients of the role per unit = 2 float4 per lane and unit; positions 4 lane .. 4 lane + 3 of a row go
    // to E[2 lane], O[2 lane], E[2 lane + 1], O[2 lane + 1]
    constexpr int kRow = ImdctLayout<kRole>::kRow;
    unsigned big = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float4 x = live[k >> 1]
                           ? __ldg(reinterpret_cast<const float4 *>(coefs + (size_t)(u0 + (k >> 1)) * 512 + kOff) + lane + 32 * (k & 1))
                           : make_float4(0.f, 0.f, 0.f, 0.f);
      float *row = kRole == 0 ? S.rows + k * kRow + 2 * lane : S.rows + (k >> 1) * kRow + 64 * (k & 1) + 2 * lane;
      *reinterpret_cast<float2 *>(row) = make_float2(x.x, x.z);
      *reinterpret_cast<float2 *>(row + kSize / 2) = make_float2(x.y, x.w);
      big = max(big, max(max(__float_as_uint(x.x) & 0x7FFFFFFFu, __float_as_uint(x.y) & 0x7FFFFFFFu),
                         max(__float_as_uint(x.z) & 0x7FFFFFFFu, __float_as_uint(x.w) & 0x7FFFFFFFu)));
    }
    __syncwarp();
    // 0x71800000 is 2^100 as binary32: below it FastRound is exact for every transform value
    const bool fast = __reduce_max_sync(0xffffffffu, big) < 0x71800000u;
    if (long_mask) {
      if (fast) imdct_long_task<kRole, FastRound>(long_mask, S.rows, S.xbuf, ST.tab2, ST.tw, lane);
      else imdct_long_task_exact<kRole>(long_mask, S.rows, S.xbuf, ST.tab2, ST.tw, lane);
      __syncwarp();
    }
    if (short_mask) {
      for (int x = 0; x < G::kPerWarp; x++)
        if ((short_mask >> x) & 1)
          imdct_short_band(kRole == 0 ? x % kPer : 2, S.rows + x * kRow, reinterpret_cast<float *>(S.xbuf), fast, T, lane);
    }
#pragma unroll
    for (int k = 0; k < 4; k++)
      if (live[k >> 1]) {
        const float *row = kRole == 0 ? S.rows + k * kRow + 2 * lane : S.rows + (k >> 1) * kRow + 64 * (k & 1) + 2 * lane;
        const float2 e = *reinterpret_cast<const float2 *>(row), o = *reinterpret_cast<const float2 *>(row + kSize / 2);
        reinterpret_cast<float4 *>(inv + onchip_row((size_t)(u0 + (k >> 1))) * 512 + kOff)[lane + 32 * (k & 1)] = make_float4(e.x, o.x, e.y, o.y);
      }
  }
}

// ------------------------------------------------------------------------------------
// Time-domain band sample p from the band records of this and the previous unit
// (decoder.js:196-206,255-263; mdct.js:230-245): rec = [head16 | tail16 | samples 32..size).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float band_sample(const float *__restrict__ rec, const float *__restrict__ rec_prev,
                                             int p, const double *__restrict__ win) {
  if (p >= 32) return rec[p];
  const int i = p < 16 ? p : 31 - p;
  const double pv = rec_prev ? (double)rec_prev[16 + i] : 0.0;
  const double cv = (double)rec[15 - i];
  const double w1 = win[i], w2 = win[31 - i];
  if (p < 16) return (float)(pv * w2 - cv * w1);
  return (float)(pv * w1 + cv * w2);
}

__global__ void __launch_bounds__(256)
bands_time_kernel(const float *__restrict__ inv, const uint8_t *__restrict__ modes, int frames, int n_units,
                  const DevTables *__restrict__ T, float *__restrict__ bands) {
  const int unit = blockIdx.x;
  if (unit >= n_units) return;
  const int frame = unit % frames;
  for (int c = threadIdx.x; c < 512; c += 256) {
    const int band = band_of_coef(c);
    const int off = band == 0 ? 0 : band == 1 ? 128 : 256;
    const int size = band == 2 ? 256 : 128;
    const float *f = inv + (size_t)unit * 512 + off;
    bands[(size_t)unit * 512 + c] =
        band_sample(f, frame > 0 ? f - 512 : nullptr, c - off, T->win);
  }
}

// ------------------------------------------------------------------------------------
// K7: overlap-add + two-stage QMF synthesis, streamed: a warp walks a run of consecutive frames
// of one row and carries the decoder state (IMDCT tails, QMF delay lines) in shared memory, as
// the reference carries it from frame to frame (decoder.js:360-388, qmf.js:60-105).
//   S[m], D[m]   = f32(.5(L[m] +- M[m]))                              (qmf.js:77-83)
//   out2[2i]     = f32(sum_j D[i-23+j] * ODD[j]),  out2[2i+1] = f32(sum_j S[i-23+j] * EVEN[j])
//   S1[n], D1[n] = f32(.5(out2[n] +- H[n-39]))                        (decoder.js:362-367)
//   pcm[2n]      = f32(sum_j D1[n-23+j] * ODD[j]), pcm[2n+1]  = f32(sum_j S1[n-23+j] * EVEN[j])
// A run that does not start at the row start is primed from the band record of the unit
// before it: the state after a frame depends on that unit alone.
// Rings: S/D with 24 entries of history, a lane produces 4 consecutive i; S1/D1 with 24, 8 consecutive n
// per lane; the delayed high band with 40.  A ring whose lanes produce kR outputs each keeps element e at
// index e + 2 (e / kR): groups of kR doubles, 2 doubles of padding after each.  Lane t's register window
// (elements kR t .. kR t + kR + 23) then starts at (kR + 2) t, its even/odd element pairs are 16-byte aligned
// (one LDS.128 per pair: 14 / 16 loads per window instead of 27 / 31), and a quarter-warp's loads touch eight
// different 16-byte bank groups ((kR + 2) / 2 = 3 / 5 is odd).  A lane's kR fresh elements are one whole group.
// ------------------------------------------------------------------------------------
#ifndef C1_SY_WARPS
#define C1_SY_WARPS 8
#define C1_SY_CTAS 2
#endif
constexpr int kSyWarps = C1_SY_WARPS, kSyCtasPerSm = C1_SY_CTAS;
constexpr int kSyRingA = 228;     // > ring_at<4>(24 + 128 - 1), even
constexpr int kSyRingB = 348;     // > ring_at<8>(24 + 256 - 1), even
static_assert(ring_at<4>(151) < kSyRingA && ring_at<8>(279) < kSyRingB && kSyRingA % 2 == 0 && kSyRingB % 2 == 0, "ring sizes");
struct SyWarpSmem {
  double a[2][kSyRingA];         // D, S of stage 2
  double b[2][kSyRingB];         // D1, S1 of stage 1
  float hd[40 + 256];            // high band with 40 entries of history (39 used)
  float td[512];                 // time-domain band frame: low 128 | mid 128 | high 256
  float tail[48];                // tail16 of the previous unit's three band records
  unsigned long long mbar, pad;  // the warp's mbarrier: completion of the copy into td
};
constexpr size_t kSySmemBytes = sizeof(SyWarpSmem) * kSyWarps;

__constant__ double c_syn_even[24];
__constant__ double c_syn_odd[24];

cudaError_t upload_decode_constants(const DevTables *host_tables) {
  cudaError_t e = cudaMemcpyToSymbol(c_syn_even, host_tables->qmf_even, 24 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_syn_odd, host_tables->qmf_odd, 24 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_fft_tw, host_tables->fft_tw, sizeof(host_tables->fft_tw));
  return e;
}

// acc[r] = sum_j seq[kR*t + 1 + r + j] * taps[j], j ascending, r = 0..kR-1: window element i = r + j + 1 of
// lane t sits at (kR + 2) t + ring_at<kR>(i); the window is loaded as (even, odd) pairs
template <int kR>
__device__ __forceinline__ void fir_synthesis(const double *__restrict__ seq, int t, const double *taps,
                                              double (&acc)[kR]) {
  const double *base = seq + (kR + 2) * t;
  double w[kR + 24];
#pragma unroll
  for (int i = 0; i < kR + 24; i += 2) {
    const double2 v = *reinterpret_cast<const double2 *>(base + ring_at<kR>(i));
    w[i] = v.x;
    w[i + 1] = v.y;
  }
#pragma unroll
  for (int r = 0; r < kR; r++) acc[r] = 0.0;
#pragma unroll
  for (int j = 0; j < 24; j++) {
    const double c = taps[j];
#pragma unroll
    for (int r = 0; r < kR; r++) acc[r] = fma(w[r + j + 1], c, acc[r]);
  }
}

// The band record of a unit (512 floats, 2 KB, 16-byte aligned: context scratch) is fetched straight into S.td with one
// TMA bulk copy issued by lane 0 while the frame before it is being filtered; completion is a phase of the warp's
// mbarrier.  S.td is free from the moment the previous unit's merge has read it (the caller's warp barrier after
// sy_load_unit) until sy_load_unit of this unit waits for the copy.
__device__ __forceinline__ void sy_prefetch(SyWarpSmem &S, const float *__restrict__ rec, int lane) {
  if (lane == 0) {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' accesses to td, before the TMA unit rewrites it
    bulk_g2s((uint32_t)__cvta_generic_to_shared(S.td), rec, 2048u, (uint32_t)__cvta_generic_to_shared(&S.mbar));
  }
}

// f32(0.5 * ((double)a +- (double)b)) (qmf.js:77-83, decoder.js:362-367) without leaving binary32.  The reference's value is
// RN24(RN53(0.5 a +- 0.5 b)); rounding first to 53 and then to 24 bits is innocuous for a sum of two 24-bit numbers
// (53 >= 2 * 24 + 2), so it equals RN24(0.5 a +- 0.5 b), which is what one FMA computes when 0.5 b is exact: b zero, or
// |b| >= 2^-125 (also true for infinities and NaN).  merge_exact_in_f32 says whether that holds; the caller falls back to the
// binary64 expression otherwise.  Saves four conversions per pair on the XU pipe (16 lanes per clock per SM).
__device__ __forceinline__ unsigned merge_key(float b) { return (__float_as_uint(b) & 0x7FFFFFFFu) - 1u; }  // 0 -> 0xFFFFFFFF
__device__ __forceinline__ bool merge_exact_in_f32(unsigned min_key) { return min_key >= 0x00FFFFFFu; }
__device__ __forceinline__ void merge_f32(float a, float b, float &sum, float &dif) {
  const float hb = 0.5f * b;
  sum = fmaf(0.5f, a, hb);
  dif = fmaf(0.5f, a, -hb);
}

struct SyWarpSmem;
__device__ __noinline__ void sy_merge_lm_f64(SyWarpSmem &S, int lane);
__device__ __noinline__ void sy_merge_hd_f64(SyWarpSmem &S, int lane, float x0, float x1, float x2, float x3, float x4, float x5,
                                             float x6, float x7);

// Band record of one unit (already on its way into S.td, sy_prefetch) -> time-domain frame in S.td (first 32 samples
// of a band: overlap-add with the previous unit's tail, mdct.js:230-245), merged low/mid -> S/D ring (elements
// 24..151), high -> hd ring (40..295); the unit's tails replace the previous ones.
__device__ __forceinline__ void sy_load_unit(SyWarpSmem &S, const double w1, const double w2, int lane, uint32_t &parity) {
  mbar_wait((uint32_t)__cvta_generic_to_shared(&S.mbar), parity);
  parity ^= 1u;
  // lane p finishes sample p of each band: i = p < 16 ? p : 31 - p, w1 = WIN[i], w2 = WIN[31 - i]
  const int i = lane < 16 ? lane : 31 - lane;
  __syncwarp();
  float ola[3], tl[3];
#pragma unroll
  for (int b = 0; b < 3; b++) {
    const int off = b == 0 ? 0 : b == 1 ? 128 : 256;
    const double pv = (double)S.tail[16 * b + i];
    const double cv = (double)S.td[off + 15 - i];
    ola[b] = lane < 16 ? (float)(pv * w2 - cv * w1) : (float)(pv * w1 + cv * w2);
    tl[b] = S.td[off + lane];  // record[16..32) is tail16 (lanes >= 16)
  }
  __syncwarp();  // every lane has read the old tails and the head of the record
#pragma unroll
  for (int b = 0; b < 3; b++) {
    const int off = b == 0 ? 0 : b == 1 ? 128 : 256;
    if (lane >= 16) S.tail[16 * b + lane - 16] = tl[b];
    S.td[off + lane] = ola[b];
  }
  __syncwarp();
  // merge low / mid (qmf.js:77-83): m = 4 lane + r -> element 24 + m = 4 (lane + 6) + r
  {
    const float4 l4 = reinterpret_cast<const float4 *>(S.td)[lane], m4 = reinterpret_cast<const float4 *>(S.td + 128)[lane];
    const float l[4] = {l4.x, l4.y, l4.z, l4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
    const bool f32_ok = merge_exact_in_f32(min(min(merge_key(m[0]), merge_key(m[1])), min(merge_key(m[2]), merge_key(m[3]))));
    if (__all_sync(0xffffffffu, f32_ok)) {
      float sum[4], dif[4];
#pragma unroll
      for (int r = 0; r < 4; r++) merge_f32(l[r], m[r], sum[r], dif[r]);
      // elements 24 + 4 lane + r: the whole group lane + 6
      double2 *da = reinterpret_cast<double2 *>(S.a[0] + 6 * (lane + 6)), *sa = reinterpret_cast<double2 *>(S.a[1] + 6 * (lane + 6));
      da[0] = make_double2((double)dif[0], (double)dif[1]);
      da[1] = make_double2((double)dif[2], (double)dif[3]);
      sa[0] = make_double2((double)sum[0], (double)sum[1]);
      sa[1] = make_double2((double)sum[2], (double)sum[3]);
    } else {
      sy_merge_lm_f64(S, lane);  // binary32-subnormal mid-band samples somewhere in the warp: rare, out of line
    }
  }
  // high band into its delay ring
  {
    // (a plain copy of 256 floats: lane l moves 16-byte chunks l and l + 32, unit stride across the warp)
    const float4 h0 = reinterpret_cast<const float4 *>(S.td + 256)[lane], h1 = reinterpret_cast<const float4 *>(S.td + 256)[lane + 32];
    float4 *d = reinterpret_cast<float4 *>(S.hd + 40);
    d[lane] = h0;
    d[lane + 32] = h1;
  }
}

// stage 2 + merge with the delayed high band -> D1/S1 ring (elements 24..279)
__device__ __forceinline__ void sy_stage2(SyWarpSmem &S, int lane) {
  double ev[4], od[4];
  fir_synthesis<4>(S.a[0], lane, c_syn_odd, od);   // out2[2i],   i = 4 lane + r
  fir_synthesis<4>(S.a[1], lane, c_syn_even, ev);  // out2[2i+1]
  // H[n - 39] = hd[8 lane + c + 1], c = 0..7: three aligned 16-byte reads instead of eight 8-way conflicting ones
  const float4 ha = reinterpret_cast<const float4 *>(S.hd)[2 * lane], hb = reinterpret_cast<const float4 *>(S.hd)[2 * lane + 1],
               hc = reinterpret_cast<const float4 *>(S.hd)[2 * lane + 2];
  const float hv[8] = {ha.y, ha.z, ha.w, hb.x, hb.y, hb.z, hb.w, hc.x};
  unsigned key = merge_key(hv[0]);
#pragma unroll
  for (int c = 1; c < 8; c++) key = min(key, merge_key(hv[c]));
  float x[8];
#pragma unroll
  for (int c = 0; c < 8; c++) x[c] = (float)((c & 1) ? ev[c >> 1] : od[c >> 1]);  // n = 8 lane + c
  if (__all_sync(0xffffffffu, merge_exact_in_f32(key))) {
    float s1[8], d1[8];
#pragma unroll
    for (int c = 0; c < 8; c++) merge_f32(x[c], hv[c], s1[c], d1[c]);
    // elements 24 + n = 8 (lane + 3) + c: the whole group lane + 3
    double2 *db = reinterpret_cast<double2 *>(S.b[0] + 10 * (lane + 3)), *sb = reinterpret_cast<double2 *>(S.b[1] + 10 * (lane + 3));
#pragma unroll
    for (int q = 0; q < 4; q++) {
      db[q] = make_double2((double)d1[2 * q], (double)d1[2 * q + 1]);
      sb[q] = make_double2((double)s1[2 * q], (double)s1[2 * q + 1]);
    }
  } else {
    sy_merge_hd_f64(S, lane, x[0], x[1], x[2], x[3], x[4], x[5], x[6], x[7]);  // rare, out of line
  }
}

// The two merges as the reference writes them (binary64), for warps that hold binary32-subnormal samples.
__device__ __noinline__ void sy_merge_lm_f64(SyWarpSmem &S, int lane) {
  const float4 l4 = reinterpret_cast<const float4 *>(S.td)[lane], m4 = reinterpret_cast<const float4 *>(S.td + 128)[lane];
  const float l[4] = {l4.x, l4.y, l4.z, l4.w}, m[4] = {m4.x, m4.y, m4.z, m4.w};
#pragma unroll
  for (int r = 0; r < 4; r++) {
    const float sum = (float)(0.5 * ((double)l[r] + (double)m[r]));
    const float dif = (float)(0.5 * ((double)l[r] - (double)m[r]));
    S.a[0][6 * (lane + 6) + r] = (double)dif;
    S.a[1][6 * (lane + 6) + r] = (double)sum;
  }
}
__device__ __noinline__ void sy_merge_hd_f64(SyWarpSmem &S, int lane, float x0, float x1, float x2, float x3, float x4, float x5,
                                             float x6, float x7) {
  const float x[8] = {x0, x1, x2, x3, x4, x5, x6, x7};
#pragma unroll
  for (int c = 0; c < 8; c++) {
    const float h = S.hd[8 * lane + c + 1];
    const float s1 = (float)(0.5 * ((double)x[c] + (double)h));
    const float d1 = (float)(0.5 * ((double)x[c] - (double)h));
    S.b[0][10 * (lane + 3) + c] = (double)d1;
    S.b[1][10 * (lane + 3) + c] = (double)s1;
  }
}

__device__ __forceinline__ void sy_shift(SyWarpSmem &S, int lane, int keep_a_at, int keep_b_at) {
  // element `lane` of the a / b rings sits at keep_a_at / keep_b_at; elements 128 + lane / 256 + lane 32 groups further
  double keep_a[2], keep_b[2];
  if (lane < 24) {
#pragma unroll
    for (int p = 0; p < 2; p++) {
      keep_a[p] = S.a[p][keep_a_at + 32 * 6];
      keep_b[p] = S.b[p][keep_b_at + 32 * 10];
    }
  }
  const float h0 = S.hd[256 + lane], h1 = lane < 8 ? S.hd[288 + lane] : 0.0f;
  __syncwarp();
  if (lane < 24) {
#pragma unroll
    for (int p = 0; p < 2; p++) {
      S.a[p][keep_a_at] = keep_a[p];
      S.b[p][keep_b_at] = keep_b[p];
    }
  }
  S.hd[lane] = h0;
  if (lane < 8) S.hd[32 + lane] = h1;
}

template <int kFmt>  // 0: f32 planar rows, 1: s16 interleaved (processor.js:382-389)
__global__ void __launch_bounds__(kSyWarps * 32, kSyCtasPerSm)
synth_kernel(const float *__restrict__ inv, int frames, int halo, int n_streams, int run_len,
             const DevTables *__restrict__ T, void *__restrict__ pcm_v, size_t row_stride, int n_ch,
             float *__restrict__ save_rec) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  SyWarpSmem &S = reinterpret_cast<SyWarpSmem *>(smem_raw)[warp];
  uint32_t parity = 0;
  if (lane == 0) mbar_init((uint32_t)__cvta_generic_to_shared(&S.mbar));
  __syncwarp();
  const int wi = lane < 16 ? lane : 31 - lane;
  const double w1 = T->win[wi], w2 = T->win[31 - wi];
  const int keep_a_at = ring_at<4>(lane), keep_b_at = ring_at<8>(lane);
  const int out_frames = frames - halo;
  const int runs_per_row = (out_frames + run_len - 1) / run_len;
  const int n_runs = runs_per_row * n_streams;
  for (int run = blockIdx.x * kSyWarps + warp; run < n_runs; run += gridDim.x * kSyWarps) {
    const int stream = run / runs_per_row;
    const int f0 = halo + (run - stream * runs_per_row) * run_len;
    const int f1 = min(f0 + run_len, frames);
    const size_t row0 = (size_t)stream * frames;  // unit index of the row's frame 0
    __syncwarp();
    sy_prefetch(S, inv + onchip_row(row0 + (f0 > 0 ? f0 - 1 : f0)) * 512, lane);
    // silent state (new BufferPool, buffers.js:31-35,67-72)
    if (lane < 24) {
#pragma unroll
      for (int p = 0; p < 2; p++) {
        S.a[p][keep_a_at] = 0.0;
        S.b[p][keep_b_at] = 0.0;
      }
    }
    S.hd[lane] = 0.0f;
    if (lane < 8) S.hd[32 + lane] = 0.0f;
    S.tail[lane] = 0.0f;
    if (lane < 16) S.tail[32 + lane] = 0.0f;
    __syncwarp();
    if (f0 > 0) {
      // prime from unit f0 - 1: only samples >= 32 of its bands reach the state (they do not depend
      // on the tails before it), so the zero state above is as good as the true one
      sy_load_unit(S, w1, w2, lane, parity);
      __syncwarp();
      sy_prefetch(S, inv + onchip_row(row0 + f0) * 512, lane);
      sy_stage2(S, lane);
      __syncwarp();
      sy_shift(S, lane, keep_a_at, keep_b_at);
    }
    for (int f = f0; f < f1; f++) {
      __syncwarp();
      sy_load_unit(S, w1, w2, lane, parity);
      __syncwarp();
      if (f + 1 < f1) sy_prefetch(S, inv + onchip_row(row0 + f + 1) * 512, lane);
      sy_stage2(S, lane);
      __syncwarp();
      {  // stage 1 -> PCM: lane covers samples [16 lane, 16 lane + 16) of the frame
        double ev[8], od[8];
        fir_synthesis<8>(S.b[0], lane, c_syn_odd, od);
        fir_synthesis<8>(S.b[1], lane, c_syn_even, ev);
        const size_t sample = (size_t)(f - halo) * 512 + 16 * lane;
        if (kFmt == 0) {
          float *dstf = static_cast<float *>(pcm_v) + (size_t)stream * row_stride + sample;
          if ((reinterpret_cast<uintptr_t>(dstf) & 15) == 0) {
            float4 *dst = reinterpret_cast<float4 *>(dstf);
#pragma unroll
            for (int q = 0; q < 4; q++)
              dst[q] = make_float4((float)od[2 * q], (float)ev[2 * q], (float)od[2 * q + 1], (float)ev[2 * q + 1]);
          } else {
#pragma unroll
            for (int q = 0; q < 8; q++) { dstf[2 * q] = (float)od[q]; dstf[2 * q + 1] = (float)ev[q]; }
          }
        } else {
          short *dst = static_cast<short *>(pcm_v);
#pragma unroll
          for (int r = 0; r < 16; r++) {
            double d = (double)(float)((r & 1) ? ev[r >> 1] : od[r >> 1]);  // Math.max(-1, Math.min(1, x))
            d = d > 1.0 ? 1.0 : d;
            d = d < -1.0 ? -1.0 : d;
            const double w = d < 0.0 ? d * 32768.0 : d * 32767.0;
            dst[(sample + r) * n_ch + stream] = (short)(isnan(w) ? 0 : __double2int_rz(w));
          }
        }
      }
      __syncwarp();
      sy_shift(S, lane, keep_a_at, keep_b_at);
    }
    if (save_rec && f1 == frames) {  // the row's last unit: its band record is the decoder state of the next call
      const float4 *src = reinterpret_cast<const float4 *>(inv + (row0 + frames - 1) * 512);
      float4 *dst = reinterpret_cast<float4 *>(save_rec + (size_t)stream * 512);
#pragma unroll
      for (int k = 0; k < 4; k++) dst[lane + 32 * k] = src[lane + 32 * k];
    }
  }
}

// ------------------------------------------------------------------------------------
// Self-test of the two arithmetic shortcuts against the IEEE operations they replace:
// div_by_range vs '/', for every (word length, quantised value, scale factor); FastRound vs the
// cvt round trip on values at and around f32 rounding ties in every binade.
// ------------------------------------------------------------------------------------
__global__ void selftest_kernel(const DevTables *__restrict__ T, unsigned long long *bad) {
  const long long tid = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  const long long stride = (long long)gridDim.x * blockDim.x;
  unsigned long long local = 0;
  // (a) division: bits 2..16, q in [-2^(bits-1), 2^(bits-1)), sfi 1..63
  for (int bits = 2; bits <= 16; bits++) {
    const int range_i = (1 << (bits - 1)) - 1;
    const double range = (double)range_i;
    const double rcp = 1.0 / range;
    const long long total = (long long)(1 << bits) * 63;
    for (long long k = tid; k < total; k += stride) {
      const int q = (int)(k / 63) - (1 << (bits - 1));
      const int sfi = (int)(k % 63) + 1;
      const double x = int_to_double(q) * T->sf[sfi];
      const double want = ((double)q * T->sf[sfi]) / range;
      const double got = div_by_range(x, range, rcp);
      if (__double_as_longlong(want) != __double_as_longlong(got)) local++;
    }
  }
  // (a') the host-built dequantisation table against the device's own IEEE division, every entry
  for (int wl = 1; wl <= kDeqMaxWl; wl++) {
    const int bits = wl + 1;
    const double range = (double)((1 << wl) - 1);
    for (long long k = tid; k < (64ll << bits); k += stride) {
      const int sfi = (int)(k >> bits), code = (int)(k & ((1 << bits) - 1));
      const int q = code >= (1 << (bits - 1)) ? code - (1 << bits) : code;
      const float want = sfi ? (float)(((double)q * T->sf[sfi]) / range) : 0.0f;
      if (__float_as_uint(want) != __float_as_uint(T->deq_tab[deq_off(wl) + (sfi << bits) + code])) local++;
    }
  }
  // (b) rounding: f32 patterns stepped through all exponents, offsets of k/8 f32-ulp
  for (long long k = tid; k < (1ll << 24); k += stride) {
    const unsigned fb = (unsigned)(k * 251u) ^ (unsigned)(k << 9);
    const float f = __uint_as_float(fb);
    if (!isfinite(f)) continue;
    const double d = (double)f;
    const double ulp = (double)__uint_as_float(((fb & 0x7F800000u) ? (fb & 0x7F800000u) : 0x00800000u)) * 1.1920928955078125e-07;
#pragma unroll
    for (int j = -9; j <= 9; j++) {
      const double v = d + ulp * (0.125 * j) + ((j & 1) ? ulp * 1e-9 : 0.0);
      const double want = (double)(float)v;
      FastRound fr;
      const double got = fr.r0(v);
      // exact for everything below 2^127, zeros and f32 subnormals included
      if (abs_hi_word(v) < 0x47E00000u && __double_as_longlong(want) != __double_as_longlong(got)) local++;
    }
  }
  if (local) atomicAdd(bad, local);
}

cudaError_t launch_selftest(const DevTables *tables, unsigned long long *d_bad, cudaStream_t st) {
  selftest_kernel<<<592, 256, 0, st>>>(tables, d_bad);
  return cudaGetLastError();
}

cudaError_t launch_decode(const DecodeLaunch &L, cudaStream_t st, Prof *prof) {
  const int n_units = L.n_streams * L.frames_total;
  if (n_units == 0) return cudaSuccess;
  prof->begin(K_UNPACK, st);
  unpack_kernel<<<std::min((n_units + kUnpackWarps - 1) / kUnpackWarps, resident_ctas((const void *)unpack_kernel, kUnpackWarps * 32, 0)),
                  kUnpackWarps * 32, 0, st>>>(
      L.su, L.su_frame_stride, L.su_stream_stride, L.n_su_valid, L.frames_total, n_units, L.tables, L.coefs,
      L.modes, L.inv, L.prev_rec, ExpandedFrames{L.x_q, L.x_sfi, L.x_bits, L.x_modes});
  prof->end(K_UNPACK, st);
  {
    cudaError_t e1 = cudaFuncSetAttribute(imdct_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kImdctSmemBytes);
    if (e1 == cudaSuccess)
      e1 = cudaFuncSetAttribute(imdct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kImdctSmemBytes);
    if (e1 != cudaSuccess) return e1;
    const int n_pairs = (n_units + 1) / 2;
    const int grid = std::min((n_pairs + kImdctWarps - 1) / kImdctWarps, persistent_ctas(kImdctCtasPerSm));
    prof->begin(K_IMDCT, st);
    const bool fork = fork_roles(L.fj, grid, persistent_ctas(kImdctCtasPerSm));
    if (fork && (e1 = fork_begin(L.fj, st)) != cudaSuccess) return e1;
    imdct_kernel<0><<<grid, kImdctWarps * 32, kImdctSmemBytes, st>>>(L.coefs, L.modes, n_units, L.tables, L.inv);
    imdct_kernel<1><<<grid, kImdctWarps * 32, kImdctSmemBytes, fork ? L.fj->aux : st>>>(L.coefs, L.modes, n_units, L.tables, L.inv);
    if (fork && (e1 = fork_end(L.fj, st)) != cudaSuccess) return e1;
    prof->launches++;
    prof->end(K_IMDCT, st);
  }
  if (L.bands_dbg) {
    prof->begin(K_BANDS_TIME, st);
    bands_time_kernel<<<n_units, 256, 0, st>>>(L.inv, L.modes, L.frames_total, n_units, L.tables, L.bands_dbg);
    prof->end(K_BANDS_TIME, st);
  }
  if (L.pcm && L.frames_total > L.halo_frames) {
    cudaError_t e0 = cudaFuncSetAttribute(synth_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSySmemBytes);
    if (e0 == cudaSuccess)
      e0 = cudaFuncSetAttribute(synth_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSySmemBytes);
    if (e0 != cudaSuccess) return e0;
    const int run_len = pick_run_len(L.frames_total - L.halo_frames, L.n_streams, persistent_ctas(kSyCtasPerSm) * kSyWarps);
    const int n_runs = ((L.frames_total - L.halo_frames + run_len - 1) / run_len) * L.n_streams;
    const int grid = std::min((n_runs + kSyWarps - 1) / kSyWarps, persistent_ctas(kSyCtasPerSm));
    prof->begin(K_SYNTH, st);
    if (L.pcm_fmt == 0)
      synth_kernel<0><<<grid, kSyWarps * 32, kSySmemBytes, st>>>(L.inv, L.frames_total, L.halo_frames, L.n_streams, run_len,
                                                                  L.tables, L.pcm, L.row_stride, L.n_ch_interleave, L.save_rec);
    else
      synth_kernel<1><<<grid, kSyWarps * 32, kSySmemBytes, st>>>(L.inv, L.frames_total, L.halo_frames, L.n_streams, run_len,
                                                                  L.tables, L.pcm, L.row_stride, L.n_ch_interleave, L.save_rec);
    prof->end(K_SYNTH, st);
  }
  return cudaGetLastError();
}

}  // namespace c1
