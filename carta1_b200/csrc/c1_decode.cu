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
#include "c1_launch.h"

namespace c1 {

__device__ __forceinline__ int bitrev_d(int x, int log2n) { return (int)(__brev((unsigned)x) >> (32 - log2n)); }

__device__ __forceinline__ void warp_fft_d(float *re, float *im, int n, const double2 *__restrict__ tw,
                                           int lane) {  // fft.js:35-66
  for (int half = 1; half < n; half <<= 1) {
    for (int b = lane; b < (n >> 1); b += 32) {
      const int k = b & (half - 1);
      const int e = ((b - k) << 1) + k;
      const int o = e + half;
      const double2 w = tw[half - 1 + k];
      const double er = re[e], ei = im[e], orr = re[o], oi = im[o];
      const double tr = orr * w.x - oi * w.y;
      const double ti = orr * w.y + oi * w.x;
      re[e] = (float)(er + tr);
      im[e] = (float)(ei + ti);
      re[o] = (float)(er - tr);
      im[o] = (float)(ei - ti);
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// K5: one warp per sound unit.
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

__global__ void __launch_bounds__(128)
unpack_dequant_kernel(const uint8_t *__restrict__ su, size_t su_frame_stride, size_t su_stream_stride,
                      long long n_su_valid, int frames, int n_units, const DevTables *__restrict__ T,
                      float *__restrict__ coefs, uint8_t *__restrict__ modes) {
  __shared__ uint32_t s_words[4][56];
  __shared__ uint16_t s_base[4][52];
  __shared__ uint8_t s_wl[4][52], s_sfi[4][52];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + warp;
  if (unit >= n_units) return;
  const int stream = unit / frames, frame = unit - stream * frames;
  const long long lin = (long long)frame * (long long)su_frame_stride + (long long)stream * (long long)su_stream_stride;
  float *dst = coefs + (size_t)unit * 512;
  const FormatTables &F = T->fmt;
  if (lin >= n_su_valid) {  // dummy frame {nBfu: 0, blockModes: [0,0,0]} (processor.js:299-307)
    for (int k = 0; k < 16; k++) dst[lane + 32 * k] = 0.0f;
    if (lane < 4) modes[(size_t)unit * 4 + lane] = 0;
    return;
  }
  uint32_t *words = s_words[warp];
  const uint32_t *src = reinterpret_cast<const uint32_t *>(su + (size_t)lin * kSuBytes);
  for (int i = lane; i < 56; i += 32) words[i] = i < kSuWords ? __byte_perm(src[i], 0, 0x0123) : 0u;
  __syncwarp();
  const uint32_t header = words[0] >> 16;
  const int m0 = 2 - (int)((header >> 14) & 3), m1 = 2 - (int)((header >> 12) & 3),
            m2 = 3 - (int)((header >> 10) & 3);
  const int idx = (header >> 5) & 7;
  const int n = idx == 0 ? 20 : 24 + 4 * idx;  // BFU_AMOUNTS
  for (int i = lane; i < n; i += 32) {
    s_wl[warp][i] = (uint8_t)get_bits(words, 16 + 4 * i, 4);
    s_sfi[warp][i] = (uint8_t)get_bits(words, 16 + 4 * n + 6 * i, 6);
  }
  __syncwarp();
  if (lane == 0) {
    int pos = 16 + 10 * n;
    for (int b = 0; b < n; b++) {
      s_base[warp][b] = (uint16_t)pos;
      pos += wl_bits(s_wl[warp][b]) * (int)F.specs[b];
    }
  }
  __syncwarp();
  if (lane < 4) modes[(size_t)unit * 4 + lane] = (uint8_t)(lane == 0 ? m0 != 0 : lane == 1 ? m1 != 0 : lane == 2 ? m2 != 0 : 0);
  for (int k = 0; k < 16; k++) {
    const int c = lane + 32 * k;
    const int mode = c < 128 ? m0 : (c < 256 ? m1 : m2);
    const int b = mode == 0 ? F.bfu_of_long[c] : F.bfu_of_short[c];
    float val = 0.0f;
    if (b < n) {
      const int bits = wl_bits(s_wl[warp][b]);
      if (bits > 0) {
        const int j = c - (mode == 0 ? F.start_long[b] : F.start_short[b]);
        const int v = (int)get_bits(words, (int)s_base[warp][b] + j * bits, bits);
        const int q = v >= (1 << (bits - 1)) ? v - (1 << bits) : v;  // bitstream.js:78-82
        const int sfi = s_sfi[warp][b];
        if (sfi != 0) {
          const int range = (1 << (bits - 1)) - 1;
          val = (float)(((double)q * T->sf[sfi]) / (double)range);  // quantization.js:75
        }
      }
    }
    dst[c] = val;
  }
}

// ------------------------------------------------------------------------------------
// K6: IMDCT, one warp per sound unit.  Only inv[N/4 .. N/4 + N/2) is ever used by the
// decoder (decoder.js:186-194,268-276), so only that half is produced.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void imdct_warp(const float *__restrict__ in, bool reverse, int n, int lg_fft,
                                           const double *__restrict__ tab, const double2 *__restrict__ tw,
                                           float *re, float *im, float *__restrict__ out, int lane) {
  const int n4 = n >> 2, half = n >> 1, fft_n = n >> 2;
  for (int i = lane; i < fft_n; i += 32) {  // mdct.js:161-170
    const int i2 = 2 * i;
    const int ia = reverse ? half - 1 - i2 : i2;
    const int ib = reverse ? i2 : half - 1 - i2;
    const double r = -(double)in[ia];
    const double m = -(double)in[ib];
    const double c = tab[i2], s = tab[i2 + 1];
    const int q = bitrev_d(i, lg_fft);
    re[q] = (float)(m * s + r * c);
    im[q] = (float)(m * c - r * s);
  }
  __syncwarp();
  warp_fft_d(re, im, fft_n, tw, lane);
  for (int i = lane; i < fft_n; i += 32) {  // mdct.js:177-208, restricted to [n4, 3*n4)
    const int i2 = 2 * i;
    const double c = tab[i2], s = tab[i2 + 1];
    const double r = re[i], m = im[i];
    const float r1 = (float)(r * c + m * s);
    const float i1 = (float)(r * s - m * c);
    if (i < (fft_n >> 1)) {
      out[half - 1 - i2] = r1;  // output[n34 - 1 - i2]
      out[i2] = i1;             // output[n4 + i2]
    } else {
      const int idx = (i - (fft_n >> 1)) * 2 + n4;
      out[half - 1 - idx] = r1;  // output[n34 - 1 - idx]
      out[idx] = i1;             // output[n4 + idx]
    }
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128)
imdct_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes, int n_units,
             const DevTables *__restrict__ T, float *__restrict__ inv) {
  __shared__ float s_re[4][128], s_im[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int unit = blockIdx.x * 4 + warp;
  if (unit >= n_units) return;
  float *re = s_re[warp], *im = s_im[warp];
  for (int band = 0; band < 3; band++) {
    const int size = band == 2 ? 256 : 128;
    const int off = band == 0 ? 0 : band == 1 ? 128 : 256;
    const float *src = coefs + (size_t)unit * 512 + off;
    float *dst = inv + (size_t)unit * 512 + off;
    if (modes[(size_t)unit * 4 + band] == 0) {
      imdct_warp(src, band > 0, band == 2 ? 512 : 256, band == 2 ? 7 : 6,
                 band == 2 ? T->mdct_inv512 : T->mdct_inv256, T->fft_tw, re, im, dst, lane);
    } else {
      for (int b = 0; b < (size >> 5); b++)
        imdct_warp(src + 32 * b, band > 0, 64, 4, T->mdct_inv64, T->fft_tw, re, im, dst + 32 * b, lane);
    }
  }
}

// ------------------------------------------------------------------------------------
// Overlap-add as a pure function of the IMDCT halves (decoder.js:175-306, mdct.js:230-245).
// inv_f / inv_prev: this / the previous frame's band slice of `inv`; p: position in band.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ float band_sample(const float *__restrict__ inv_f, const float *__restrict__ inv_prev,
                                             int size, bool is_long, int p, const double *__restrict__ win) {
  if (is_long && p >= 32) return inv_f[p - 16];
  const int q = is_long ? p : (p & 31);
  const int blk = p - q;
  const int i = q < 16 ? q : 31 - q;
  const float pv = blk == 0 ? (inv_prev ? inv_prev[size - 16 + i] : 0.0f) : inv_f[blk - 16 + i];
  const float cv = inv_f[blk + 15 - i];
  const double w1 = win[i], w2 = win[31 - i];
  if (q < 16) return (float)((double)pv * w2 - (double)cv * w1);
  return (float)((double)pv * w1 + (double)cv * w2);
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
        band_sample(f, frame > 0 ? f - 512 : nullptr, size, modes[(size_t)unit * 4 + band] == 0, c - off, T->win);
  }
}

// ------------------------------------------------------------------------------------
// K7: overlap-add + QMF synthesis for a tile of frames of one row.
//   S[m], D[m]   = f32(.5(L[m] +- M[m]))                              (qmf.js:77-83)
//   out2[2i]     = f32(sum_j D[i-23+j] * ODD[j]),  out2[2i+1] = f32(sum_j S[i-23+j] * EVEN[j])
//   S1[n], D1[n] = f32(.5(out2[n] +- H[n-39]))                        (decoder.js:362-367)
//   pcm[2n]      = f32(sum_j D1[n-23+j] * ODD[j]), pcm[2n+1]  = f32(sum_j S1[n-23+j] * EVEN[j])
// Negative indices are the zero-initialised delay lines.  As in K1 the sequences live in
// shared memory as binary64, row kk&7 / column kk>>3, and every thread produces 8 outputs
// of each polyphase from a 31-value register window.
// ------------------------------------------------------------------------------------
constexpr int kSynTile = 8;
constexpr int kSynThreads = 256;
constexpr int kSynS2Threads = 16 * kSynTile + 2;
constexpr int kSynStrideA = 146;  // >= (8*kSynS2Threads + 31)/8 + 1, == 2 mod 16
constexpr int kSynStrideB = 274;  // >= (8*32*kSynTile + 31)/8 + 1,   == 2 mod 16
constexpr int kSynHd = 256 * kSynTile + 24;

__constant__ double c_syn_even[24];
__constant__ double c_syn_odd[24];

cudaError_t upload_decode_constants(const double *even24, const double *odd24) {
  cudaError_t e = cudaMemcpyToSymbol(c_syn_even, even24, 24 * sizeof(double));
  if (e != cudaSuccess) return e;
  return cudaMemcpyToSymbol(c_syn_odd, odd24, 24 * sizeof(double));
}

// acc[r] = sum_j w[8t + 1 + r + j] * taps[j], j ascending, for r = 0..7 (thread t)
template <int kStride>
__device__ __forceinline__ void fir8_synthesis(const double *__restrict__ seq, int t, const double *taps,
                                               double (&acc)[8]) {
#pragma unroll
  for (int r = 0; r < 8; r++) acc[r] = 0.0;
#pragma unroll
  for (int j = 0; j < 24; j++) {
    const double c = taps[j];
#pragma unroll
    for (int r = 0; r < 8; r++) {
      const int i = r + j + 1;
      acc[r] = fma(seq[(i & 7) * kStride + t + (i >> 3)], c, acc[r]);
    }
  }
}

template <int kFmt>  // 0: f32 planar rows, 1: s16 interleaved (processor.js:382-389)
__global__ void __launch_bounds__(kSynThreads)
synth_kernel(const float *__restrict__ inv, const uint8_t *__restrict__ modes, int frames, int halo,
             const DevTables *__restrict__ T, void *__restrict__ pcm_v, size_t row_stride, int n_ch) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  double *sa = reinterpret_cast<double *>(smem_raw);  // [2][8*kSynStrideA]: D, S of stage 2
  double *sb = sa + 2 * 8 * kSynStrideA;               // [2][8*kSynStrideB]: D1, S1 of stage 1
  float *hd = reinterpret_cast<float *>(sb + 2 * 8 * kSynStrideB);  // delayed high band, n >= 256*f0 - 24
  __shared__ double win[32];
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * kSynTile;
  const int stream = blockIdx.y;
  if (tid < 32) win[tid] = T->win[tid];
  __syncthreads();
  const float *inv_row = inv + (size_t)stream * frames * 512;
  const uint8_t *mode_row = modes + (size_t)stream * frames * 4;
  const int f_end = min(f0 + kSynTile, frames);

  // merged low/mid pairs: kk -> m = 128*f0 - 40 + kk
  const int m_lo = 128 * f0 - 40;
  for (int kk = tid; kk < 8 * kSynS2Threads + 32; kk += kSynThreads) {
    const int m = m_lo + kk;
    float sum = 0.0f, dif = 0.0f;
    if (m >= 0 && m < 128 * f_end) {
      const int fr = m >> 7, p = m & 127;
      const float *fl = inv_row + (size_t)fr * 512;
      const float l = band_sample(fl, fr > 0 ? fl - 512 : nullptr, 128, mode_row[fr * 4 + 0] == 0, p, win);
      const float h = band_sample(fl + 128, fr > 0 ? fl - 384 : nullptr, 128, mode_row[fr * 4 + 1] == 0, p, win);
      sum = (float)(0.5 * ((double)l + (double)h));
      dif = (float)(0.5 * ((double)l - (double)h));
    }
    const int at = (kk & 7) * kSynStrideA + (kk >> 3);
    sa[at] = (double)dif;
    sa[8 * kSynStrideA + at] = (double)sum;
  }
  // delayed high band: hd[q] = H[n - 39], n = 256*f0 - 24 + q
  for (int q = tid; q < kSynHd; q += kSynThreads) {
    const int g = 256 * f0 - 24 + q - 39;
    float h = 0.0f;
    if (g >= 0 && g < 256 * f_end) {
      const int fr = g >> 8, p = g & 255;
      const float *fh = inv_row + (size_t)fr * 512 + 256;
      h = band_sample(fh, fr > 0 ? fh - 512 : nullptr, 256, mode_row[fr * 4 + 2] == 0, p, win);
    }
    hd[q] = h;
  }
  __syncthreads();
  // stage 2 (low + mid -> 256-rate signal) and the merge with the delayed high band
  if (tid < kSynS2Threads) {
    double ev[8], od[8];
    fir8_synthesis<kSynStrideA>(sa, tid, c_syn_odd, od);                     // out2[2i]
    fir8_synthesis<kSynStrideA>(sa + 8 * kSynStrideA, tid, c_syn_even, ev);  // out2[2i+1]
#pragma unroll
    for (int r = 0; r < 8; r++) {
#pragma unroll
      for (int par = 0; par < 2; par++) {
        const int n = 2 * (128 * f0 - 16 + 8 * tid + r) + par;
        const int kk1 = n - (256 * f0 - 24);
        if (kk1 < 0) continue;
        float s1 = 0.0f, d1 = 0.0f;
        if (n >= 0 && n < 256 * f_end) {
          const float x = (float)(par ? ev[r] : od[r]);
          const float h = hd[kk1];
          s1 = (float)(0.5 * ((double)x + (double)h));
          d1 = (float)(0.5 * ((double)x - (double)h));
        }
        const int at = (kk1 & 7) * kSynStrideB + (kk1 >> 3);
        sb[at] = (double)d1;
        sb[8 * kSynStrideB + at] = (double)s1;
      }
    }
  }
  __syncthreads();
  // stage 1 -> PCM: thread t covers samples [16t, 16t+16) of the tile
  {
    const int fr = f0 + (tid >> 5);
    if (fr < frames && fr >= halo) {
      double ev[8], od[8];
      fir8_synthesis<kSynStrideB>(sb, tid, c_syn_odd, od);
      fir8_synthesis<kSynStrideB>(sb + 8 * kSynStrideB, tid, c_syn_even, ev);
      const size_t sample = (size_t)(fr - halo) * 512 + 16 * (tid & 31);
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
  }
}
constexpr size_t kSynSmemBytes =
    (size_t)(2 * 8 * kSynStrideA + 2 * 8 * kSynStrideB) * sizeof(double) + (size_t)kSynHd * sizeof(float);

cudaError_t launch_decode(const DecodeLaunch &L, cudaStream_t st, Prof *prof) {
  const int n_units = L.n_streams * L.frames_total;
  if (n_units == 0) return cudaSuccess;
  prof->begin(K_UNPACK_DEQUANT, st);
  unpack_dequant_kernel<<<(n_units + 3) / 4, 128, 0, st>>>(L.su, L.su_frame_stride, L.su_stream_stride,
                                                         L.n_su_valid, L.frames_total, n_units, L.tables,
                                                         L.coefs, L.modes);
  prof->end(K_UNPACK_DEQUANT, st);
  prof->begin(K_IMDCT, st);
  imdct_kernel<<<(n_units + 3) / 4, 128, 0, st>>>(L.coefs, L.modes, n_units, L.tables, L.inv);
  prof->end(K_IMDCT, st);
  if (L.bands_dbg) {
    prof->begin(K_BANDS_TIME, st);
    bands_time_kernel<<<n_units, 256, 0, st>>>(L.inv, L.modes, L.frames_total, n_units, L.tables, L.bands_dbg);
    prof->end(K_BANDS_TIME, st);
  }
  if (L.pcm) {
    dim3 grid((L.frames_total + kSynTile - 1) / kSynTile, L.n_streams);
    cudaError_t e0 = cudaFuncSetAttribute(synth_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSynSmemBytes);
    if (e0 == cudaSuccess)
      e0 = cudaFuncSetAttribute(synth_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kSynSmemBytes);
    if (e0 != cudaSuccess) return e0;
    prof->begin(K_SYNTH, st);
    if (L.pcm_fmt == 0)
      synth_kernel<0><<<grid, kSynThreads, kSynSmemBytes, st>>>(L.inv, L.modes, L.frames_total, L.halo_frames,
                                                                 L.tables, L.pcm, L.row_stride, L.n_ch_interleave);
    else
      synth_kernel<1><<<grid, kSynThreads, kSynSmemBytes, st>>>(L.inv, L.modes, L.frames_total, L.halo_frames,
                                                                 L.tables, L.pcm, L.row_stride, L.n_ch_interleave);
    prof->end(K_SYNTH, st);
  }
  return cudaGetLastError();
}

}  // namespace c1
