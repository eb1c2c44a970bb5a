// c1_encode.cu -- ATRAC1 encode kernels for sm_100a.
//
// Pipeline (all stateless over a row of PCM that starts at silence):
//   K1 qmf_analysis_kernel    PCM -> low/mid/high bands            (qmf.js:19-50, encoder.js:69-95)
//   K2 band_mags_kernel       bands -> transient magnitude spectra (transient.js:17-35)
//      transient_modes_kernel magnitudes(f), magnitudes(f-1) -> block modes (transient.js:44-226)
//   K3 mdct_kernel            bands(f), tail(f-1), modes -> 512 coefficients (encoder.js:170-349)
//   K4 alloc_quant_pack_kernel coefficients -> 212-byte sound unit
//                             (quantization.js:34-149, bitallocation.js:74-341, serialization.js:41-98)
// Arithmetic: binary64, one rounding per reference operator, binary32 at every typed-array
// store.  Compiled with -fmad=false.
#include "c1_common.cuh"
#include "c1_fdlibm.cuh"
#include "c1_launch.h"

namespace c1 {

// ------------------------------------------------------------------------------------
// Warp FFT in shared memory: radix-2 DIT, input already in bit-reversed order, f32 store
// after every butterfly (fft.js:35-66).  tw[h - 1 + k] is the k-th recurrence twiddle of
// half-stride h.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void warp_fft(float *re, float *im, int n, const double2 *__restrict__ tw,
                                         int lane) {
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

__device__ __forceinline__ int bitrev(int x, int log2n) { return (int)(__brev((unsigned)x) >> (32 - log2n)); }

// ------------------------------------------------------------------------------------
// K1: two-stage QMF analysis over a tile of frames, with halos recomputed per CTA.
//   S1lo/S1hi[n] = f32(E +- O), E = sum_j x[2n+1-2j]*EVEN[j], O = sum_j x[2n-2j]*ODD[j]
//   L/M[m]       = the same filter applied to S1lo;  H[n] = S1hi[n-39].
// fma() is exact-product here (both factors are widened f32), so it equals mul-then-add.
// ------------------------------------------------------------------------------------
constexpr int kQmfTile = 4;  // frames per CTA

template <int kFmt>  // 0: f32 planar rows, 1: s16 interleaved
__global__ void __launch_bounds__(256)
qmf_analysis_kernel(const void *__restrict__ pcm_v, size_t row_stride, int n_ch, long long valid_samples,
                    int frames, const DevTables *__restrict__ T, float *__restrict__ bands) {
  __shared__ float xs[kQmfTile * 512 + 138 + 2];
  __shared__ float s1[kQmfTile * 256 + 46 + 2];
  __shared__ double ce[24], co[24];
  const int tid = threadIdx.x;
  const int f0 = blockIdx.x * kQmfTile;
  const int stream = blockIdx.y;
  if (tid < 24) { ce[tid] = T->qmf_even[tid]; co[tid] = T->qmf_odd[tid]; }
  const long long x0 = 512ll * f0 - 138;
  for (int i = tid; i < kQmfTile * 512 + 138; i += 256) {
    const long long g = x0 + i;
    float v = 0.0f;
    if (g >= 0 && g < valid_samples) {
      if (kFmt == 0) {
        v = static_cast<const float *>(pcm_v)[(size_t)stream * row_stride + (size_t)g];
      } else {  // bin/cli.js:395  readInt16LE / 32768.0 -> Float32Array
        const short s = static_cast<const short *>(pcm_v)[(size_t)g * n_ch + stream];
        v = (float)((double)s / 32768.0);
      }
    }
    xs[i] = v;
  }
  __syncthreads();
  float *out = bands + ((size_t)stream * frames) * 512;
  const int n_lo = 256 * f0 - 46;
  for (int t = tid; t < kQmfTile * 256 + 46; t += 256) {
    double e = 0.0, o = 0.0;
#pragma unroll
    for (int j = 0; j < 24; j++) {
      e = fma((double)xs[47 + 2 * t - 2 * j], ce[j], e);
      o = fma((double)xs[46 + 2 * t - 2 * j], co[j], o);
    }
    s1[t] = (float)(e + o);
    const int nh = n_lo + t + 39;  // delayed high-band index (encoder.js:84-90)
    if (nh >= 256 * f0 && nh < 256 * (f0 + kQmfTile)) {
      const int fr = nh >> 8;
      if (fr < frames) out[(size_t)fr * 512 + 256 + (nh & 255)] = (float)(e - o);
    }
  }
  __syncthreads();
  for (int u = tid; u < kQmfTile * 128; u += 256) {
    double e = 0.0, o = 0.0;
#pragma unroll
    for (int j = 0; j < 24; j++) {
      e = fma((double)s1[47 + 2 * u - 2 * j], ce[j], e);
      o = fma((double)s1[46 + 2 * u - 2 * j], co[j], o);
    }
    const int fr = f0 + (u >> 7);
    if (fr < frames) {
      out[(size_t)fr * 512 + (u & 127)] = (float)(e + o);
      out[(size_t)fr * 512 + 128 + (u & 127)] = (float)(e - o);
    }
  }
}

// ------------------------------------------------------------------------------------
// K2a: magnitude spectra for transient detection, one warp per sound unit.
// ------------------------------------------------------------------------------------
__global__ void __launch_bounds__(128)
band_mags_kernel(const float *__restrict__ bands, int n_su, const DevTables *__restrict__ T,
                 float *__restrict__ mags) {
  __shared__ float s_re[4][256], s_im[4][256];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int su = blockIdx.x * 4 + warp;
  if (su >= n_su) return;
  float *re = s_re[warp], *im = s_im[warp];
  for (int band = 0; band < 3; band++) {
    const int n = band == 2 ? 256 : 128;
    const int lg = band == 2 ? 8 : 7;
    const float *src = bands + (size_t)su * 512 + (band == 0 ? 0 : band == 1 ? 128 : 256);
    for (int i = lane; i < n; i += 32) {
      re[bitrev(i, lg)] = src[i];
      im[i] = 0.0f;
    }
    __syncwarp();
    warp_fft(re, im, n, T->fft_tw, lane);
    float *dst = mags + (size_t)su * 256 + (band == 0 ? 0 : band == 1 ? 64 : 128);
    for (int i = lane; i < (n >> 1); i += 32) {
      const double r = re[i], m = im[i];
      dst[i] = (float)sqrt(r * r + m * m);  // transient.js:31
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// K2b: transient score (transient.js:63-226), one thread per (sound unit, band).  All sums
// run serially in index order, as the reference's loops do.
// ------------------------------------------------------------------------------------
__device__ double js_max(double a, double b) {
  if (isnan(a) || isnan(b)) return nan("");
  if (a == 0.0 && b == 0.0) return signbit(a) ? b : a;
  return a > b ? a : b;
}
__device__ double js_min(double a, double b) {
  if (isnan(a) || isnan(b)) return nan("");
  if (a == 0.0 && b == 0.0) return signbit(a) ? a : b;
  return a < b ? a : b;
}

struct SpectrumFeatures {
  double flatness, hf_ratio, energy;
};

__device__ SpectrumFeatures spectrum_features(const float *__restrict__ x, int n) {
  SpectrumFeatures f;
  const double EPS = 1e-10;
  double sum_log = 0.0, sum_lin = 0.0, lo = 0.0, hi = 0.0, energy = 0.0;
  int valid = 0;
  const int mid = n >> 1;
  for (int i = 0; i < n; i++) {
    const double v = x ? (double)x[i] : 0.0;
    const double m = fabs(v);
    if (m > EPS) {
      sum_log += fd::log(m);
      sum_lin += m;
      valid++;
    }
    const double sq = v * v;
    if (i < mid) lo += sq; else hi += sq;
    energy += sq;
  }
  if (valid == 0) {
    f.flatness = 0.0;
  } else {
    const double geo = fd::exp(sum_log / valid);
    const double arith = sum_lin / valid;
    f.flatness = arith > EPS ? geo / arith : 0.0;
  }
  const double total = lo + hi;
  f.hf_ratio = total > 0.0 ? hi / total : 0.0;
  f.energy = energy;
  return f;
}

__device__ double transient_score(const float *__restrict__ cur, const float *__restrict__ prev, int n,
                                  double log1p10) {
  double flux = 0.0, cur_energy = 0.0;
  for (int i = 0; i < n; i++) {  // transient.js:92-112
    const double c = fabs((double)cur[i]);
    const double p = prev ? fabs((double)prev[i]) : 0.0;
    const double d = c - p;
    if (d > 0.0) flux += d;
    cur_energy += c * c;
  }
  double norm = sqrt(cur_energy);
  if (norm == 0.0 || isnan(norm)) norm = 1e-6;
  const double spectral_flux = flux / norm;
  const SpectrumFeatures fc = spectrum_features(cur, n);
  const SpectrumFeatures fp = spectrum_features(prev, n);
  const double flat_change = fabs(fc.flatness - fp.flatness);
  const double hf_change = fabs(fc.hf_ratio - fp.hf_ratio);
  const double ce = js_max(fc.energy, 1e-10), pe = js_max(fp.energy, 1e-10);  // :182-183
  const double db = 10.0 * fd::log10(ce / pe);
  const double e_change = js_max(0.0, db);
  const double flat_c = sqrt(flat_change);
  const double hf_c = fd::log1p(hf_change * 10.0) / log1p10;
  const double e_c = js_min(e_change / 30.0, 1.0);
  return (spectral_flux + flat_c + hf_c + e_c) / 4.0;
}

__global__ void __launch_bounds__(128)
transient_modes_kernel(const float *__restrict__ mags, int frames, int n_su,
                       const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                       uint8_t *__restrict__ modes, double *__restrict__ scores) {
  const int idx = blockIdx.x * blockDim.x + threadIdx.x;
  const int su = idx / 3, band = idx - su * 3;
  if (su >= n_su) return;
  const int frame = su % frames;
  const int n = band == 2 ? 128 : 64;
  const int off = band == 0 ? 0 : band == 1 ? 64 : 128;
  const float *cur = mags + (size_t)su * 256 + off;
  const float *prev = frame > 0 ? cur - 256 : nullptr;  // frame 0: all-zero previous spectrum
  const double score = transient_score(cur, prev, n, T->log1p10);
  // every band compares against transientThresholdLow (encoder.js:137-141); mode = t*max(b+1,2)
  const int transient = score > P->threshold;
  modes[(size_t)su * 4 + band] = (uint8_t)(transient ? (band == 2 ? 3 : 2) : 0);
  if (scores) scores[(size_t)su * 3 + band] = score;
}

// ------------------------------------------------------------------------------------
// K3: windowed MDCT, one warp per sound unit (mdct.js:54-122, encoder.js:228-316).
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void mdct_warp(const float *in, int n, int lg_fft, const double *__restrict__ tab,
                                          const double2 *__restrict__ tw, float *re, float *im,
                                          float *__restrict__ out, bool reverse, int lane) {
  const int n4 = n >> 2, n34 = 3 * n4, half = n >> 1, fft_n = n >> 2;
  for (int p = lane; p < fft_n; p += 32) {
    const int i = 2 * p;
    double r, m;
    if (i < n4) {  // mdct.js:76-89
      r = (double)in[n34 - 1 - i] + (double)in[n34 + i];
      m = (double)in[n4 + i] - (double)in[n4 - 1 - i];
    } else {       // mdct.js:91-105
      r = (double)in[n34 - 1 - i] - (double)in[i - n4];
      m = (double)in[n4 + i] + (double)in[5 * n4 - 1 - i];
    }
    const double c = tab[i], s = tab[i + 1];
    const int q = bitrev(p, lg_fft);
    re[q] = (float)(r * c + m * s);
    im[q] = (float)(m * c - r * s);
  }
  __syncwarp();
  warp_fft(re, im, fft_n, tw, lane);
  for (int i = lane; i < fft_n; i += 32) {  // mdct.js:111-119
    const double c = tab[2 * i], s = tab[2 * i + 1];
    const double r = re[i], m = im[i];
    const float o0 = (float)(-r * c - m * s);
    const float o1 = (float)(-r * s + m * c);
    int i0 = 2 * i, i1 = half - 1 - 2 * i;
    if (reverse) { i0 = half - 1 - i0; i1 = half - 1 - i1; }  // utils.js:42-48
    out[i0] = o0;
    out[i1] = o1;
  }
  __syncwarp();
}

__global__ void __launch_bounds__(128)
mdct_kernel(const float *__restrict__ bands, const uint8_t *__restrict__ modes, int frames, int n_su,
            const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
            float *__restrict__ coefs) {
  __shared__ float s_buf[4][512], s_re[4][128], s_im[4][128];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int su = blockIdx.x * 4 + warp;
  if (su >= n_su) return;
  const int frame = su % frames;
  float *buf = s_buf[warp], *re = s_re[warp], *im = s_im[warp];
  const double *win = T->win;
  for (int band = 0; band < 3; band++) {
    const int size = band == 2 ? 256 : 128;
    const int off = band == 0 ? 0 : band == 1 ? 128 : 256;
    const float *cur = bands + (size_t)su * 512 + off;
    const float *prev = frame > 0 ? cur - 512 : nullptr;
    const int mode = P->use_fixed ? P->fixed[band] : (int)modes[(size_t)su * 4 + band];
    float *dst = coefs + (size_t)su * 512 + off;
    if (mode == 0) {
      const int n = band == 2 ? 512 : 256;
      const int ws = band == 2 ? 112 : 48;  // constants.js:115-119
      for (int k = lane; k < n; k += 32) {
        float v = 0.0f;
        const int a = k - ws;
        if (a >= 0 && a < 32) {  // overlap saved by the previous frame's tail windowing
          v = prev ? (float)(win[a] * (double)prev[size - 32 + a]) : 0.0f;
        } else if (a >= 32 && a < 32 + size) {
          const int sidx = a - 32;
          const float x = cur[sidx];
          const int t = sidx - (size - 32);
          v = t >= 0 ? (float)((double)x * win[31 - t]) : x;
        }
        buf[k] = v;
      }
      __syncwarp();
      mdct_warp(buf, n, band == 2 ? 7 : 6, band == 2 ? T->mdct_fwd512 : T->mdct_fwd256, T->fft_tw, re, im,
                dst, band > 0, lane);
    } else {
      const int blocks = size >> 5;
      for (int b = 0; b < blocks; b++) {
        {
          // lanes 0..31: overlap = WIN[i] * (previous 32-sample block); block = x * WIN[31-i]
          const int i = lane;
          float src_prev;
          if (b == 0) src_prev = prev ? prev[size - 32 + i] : 0.0f;
          else src_prev = cur[32 * (b - 1) + i];
          // frame 0 / block 0 starts from the all-zero overlap buffer (buffers.js:60-65)
          buf[i] = (b == 0 && !prev) ? 0.0f : (float)(win[i] * (double)src_prev);
          buf[32 + i] = (float)((double)cur[32 * b + i] * win[31 - i]);
        }
        __syncwarp();
        mdct_warp(buf, 64, 4, T->mdct_fwd64, T->fft_tw, re, im, dst + 32 * b, band > 0, lane);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K4a: scale factors + RDO bit allocation by literal max-heap emulation
// (bitallocation.js:74-341).  One CTA = 256 threads = 32 sound units; warp w runs candidate
// BFU count BFU_AMOUNTS[w] for all 32 units (lane = unit), so the lanes of a warp execute
// loops of similar length.
//
// Heap entries are one 32-bit word: key[24:10] | wl[9:6] | bfu[5:0].  `key` is an
// order-isomorphic 15-bit image of the reference's Float32Array priority (DevEncParams::key0/
// key1): f32 exponent (8 bits) over the rank of the f32 mantissa among the 126 possible
// ones.  For wl >= 1 the priority halves exactly with every step, i.e. key -= 128.
// Heaps are stored node-major / lane-minor: a warp's 32 lanes always hit 32 distinct banks.
// ------------------------------------------------------------------------------------
constexpr int kAlSu = 32;
constexpr int kAlThreads = 256;
constexpr int kAlHeapSlots = 308;  // sum over candidates of (cand + 1)
constexpr int kAlWlSlots = 300;    // sum over candidates of cand

struct AllocRec {  // per sound unit, global scratch between K4a and K4b
  uint8_t n_bfu, pad[3];
  uint8_t wl[52];
  uint8_t sfi[52];
  uint8_t pad2[4];
};
static_assert(sizeof(AllocRec) == 112, "AllocRec layout");

struct AlSmem {
  uint32_t heap[kAlHeapSlots][32];
  uint8_t wl[kAlWlSlots][32];
  double dist[8][32];
  uint16_t key0[64], key1[64];
  uint8_t sfi[kAlSu][52];
  uint8_t mode[kAlSu][4];
  uint8_t specs[52];
  int best[kAlSu];
};

__device__ __forceinline__ void heap_sift(uint32_t (*H)[32], int lane, int start, int size, uint32_t v) {
  // bitallocation.js:314-341; v is the entry being placed, H[start] is the hole
  int i = start;
  const uint32_t vm = v | 0x3FFu;
  for (;;) {
    const int l = 2 * i + 1;
    if (l >= size) break;
    const uint32_t cl = H[l][lane];
    const uint32_t cr = H[l + 1][lane];  // slot `size` is kept zero
    const bool pl = cl > vm;
    const uint32_t m = pl ? (cl | 0x3FFu) : vm;
    const bool pr = cr > m;
    if (!(pl || pr)) break;
    H[i][lane] = pr ? cr : cl;
    i = pr ? l + 1 : l;
  }
  H[i][lane] = v;
}

__global__ void __launch_bounds__(kAlThreads)
alloc_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes, int frames, int halo,
             int n_out_frames, int n_streams, const DevTables *__restrict__ T,
             const DevEncParams *__restrict__ P, AllocRec *__restrict__ recs) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlSmem &S = *reinterpret_cast<AlSmem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_units = (long long)n_streams * n_out_frames;
  const long long unit0 = (long long)blockIdx.x * kAlSu;
  const FormatTables &F = T->fmt;

  if (tid < 64) { S.key0[tid] = P->key0[tid]; S.key1[tid] = P->key1[tid]; }
  if (tid < 52) S.specs[tid] = F.specs[tid];
  if (tid < kAlSu * 3) {
    const int u = tid / 3, b = tid - u * 3;
    const long long unit = unit0 + u;
    uint8_t m = 0;
    if (unit < n_units) {
      const size_t su = (size_t)(unit / n_out_frames) * frames + halo + (size_t)(unit % n_out_frames);
      m = P->use_fixed ? (uint8_t)(P->fixed[b] != 0) : modes[su * 4 + b];
    }
    S.mode[u][b] = m;
  }
  __syncthreads();

  // ---- phase A: scale-factor index per BFU (bitallocation.js:290-299).  The log2/ceil of
  // the reference equals 3*(E+21) + #{thresholds of the binade below max} (DevTables::sf_thr).
  for (int item = tid; item < kAlSu * 52; item += kAlThreads) {
    const int u = item / 52, b = item - u * 52;
    const long long unit = unit0 + u;
    int sfi = 0;
    if (unit < n_units) {
      const size_t su = (size_t)(unit / n_out_frames) * frames + halo + (size_t)(unit % n_out_frames);
      const int sz = F.specs[b];
      const int start = S.mode[u][band_of_bfu(b)] == 0 ? F.start_long[b] : F.start_short[b];
      const float *src = coefs + su * 512 + start;
      float mx = 0.0f;
      for (int j = 0; j < sz; j++) {
        const float a = fabsf(src[j]);
        if (a > mx) mx = a;  // NaN never wins, as in the reference
      }
      if (mx > 0.0f) {
        const uint32_t bits = __float_as_uint(mx);
        const int e3 = 3 * ((int)(bits >> 23) - 127 + 21);
        if (e3 >= 63) sfi = 63;
        else if (e3 >= 0) {
          sfi = e3 + (mx > T->sf_thr[e3]) + (mx > T->sf_thr[e3 + 1]) + (mx > T->sf_thr[e3 + 2]);
          if (sfi > 63) sfi = 63;
        }
      }
    }
    S.sfi[u][b] = (uint8_t)sfi;
  }
  __syncthreads();

  // ---- phase B: warp = candidate, lane = unit
  {
    const int u = lane;
    const int c = warp;
    const int cand = c == 0 ? 20 : 24 + 4 * c;               // BFU_AMOUNTS
    const int hbase = c == 0 ? 0 : (21 + (c - 1) * 29 + 2 * (c - 1) * (c - 2));  // prefix of (cand+1)
    const int wbase = hbase - c;                              // prefix of cand
    uint32_t (*H)[32] = &S.heap[hbase];
    uint8_t (*W)[32] = &S.wl[wbase];
    const bool live = unit0 + u < n_units;
    double total = 0.0;
    if (live) {
      int remaining = kFrameBits - 40 - 10 * cand;  // bitallocation.js:97-100
      int count = 0;
      for (int b = 0; b < cand; b++) {               // bitallocation.js:216-232
        W[b][lane] = 0;
        const uint32_t sfi = S.sfi[u][b];
        if (sfi) {
          H[count][lane] = ((uint32_t)S.key0[sfi] << 10) | (uint32_t)b;
          count++;
        }
      }
      for (int i = count; i <= cand; i++) H[i][lane] = 0;
      if (count) {
        for (int i = (count >> 1) - 1; i >= 0; i--) heap_sift(H, lane, i, count, H[i][lane]);
        int size = count;
        uint32_t e = H[0][lane];
        while (remaining > 0 && size > 0) {  // bitallocation.js:244-278
          const int b = e & 63;
          const int wl = (e >> 6) & 15;
          const int cost = (int)S.specs[b] << (wl == 0);
          bool pop = cost > remaining;
          if (!pop) {
            remaining -= cost;
            if (wl == 0) e = ((uint32_t)S.key1[S.sfi[u][b]] << 10) | (1u << 6) | (uint32_t)b;
            else e += 64u - (128u << 10);  // wl + 1, priority halves exactly
            pop = wl == 14;                // reached MAX_WORD_LENGTH_INDEX
          }
          if (pop) {
            W[b][lane] = (uint8_t)((e >> 6) & 15);
            size--;
            e = H[size][lane];
            H[size][lane] = 0;
            if (size == 0) break;
          }
          heap_sift(H, lane, 0, size, e);
          e = H[0][lane];
        }
        for (int i = 0; i < size; i++) {
          const uint32_t x = H[i][lane];
          W[x & 63][lane] = (uint8_t)((x >> 6) & 15);
        }
      }
      // total distortion of this candidate (bitallocation.js:157-190), index order
      for (int i = 0; i < 52; i++) {
        const int sfi = S.sfi[u][i];
        if (sfi == 0) continue;  // contributes +0.0 or is skipped by the reference
        const int bits = i < cand ? wl_bits(W[i][lane]) : 0;
        if (bits == 0) {
          total += (double)P->zero_bit[sfi * 8 + F.size_class[i]];
        } else {
          const double inv = __hiloint2double((1023 - bits) << 20, 0);
          total += P->bsf[sfi] * inv * (double)S.specs[i];
        }
      }
    }
    S.dist[c][u] = total;
    __syncthreads();
    if (warp == 0 && live) {  // first strict minimum over ascending candidates (:122-129)
      double min_total = __longlong_as_double(0x7ff0000000000000ll);
      int best = -1;
#pragma unroll
      for (int k = 0; k < 8; k++) {
        const double t = S.dist[k][u];
        if (t < min_total) { min_total = t; best = k; }
      }
      S.best[u] = best;
    }
    __syncthreads();
    if (live) {
      const int best = S.best[u];
      AllocRec *r = recs + (unit0 + u);
      if (best == c) {
        r->n_bfu = (uint8_t)cand;
        for (int b = 0; b < 52; b++) {
          r->wl[b] = b < cand ? W[b][lane] : 0;
          r->sfi[b] = S.sfi[u][b];
        }
      } else if (best < 0 && c == 0) {  // bitallocation.js:132-139
        r->n_bfu = 20;
        for (int b = 0; b < 52; b++) { r->wl[b] = 0; r->sfi[b] = 0; }
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K4b: quantise (quantization.js:34-56) and pack the 212-byte unit
// (serialization.js:41-98, bitstream.js:15-39).  One warp per sound unit.
// ------------------------------------------------------------------------------------
__device__ __forceinline__ void put_bits(uint32_t *words, int pos, uint32_t value, int bits) {
  const int w = pos >> 5, off = pos & 31;
  const unsigned long long v = (unsigned long long)value << (64 - off - bits);
  const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  if (hi) atomicOr(&words[w], hi);
  if (lo) atomicOr(&words[w + 1], lo);
}

constexpr int kQpWarps = 8;

__global__ void __launch_bounds__(kQpWarps * 32)
quant_pack_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes,
                  const AllocRec *__restrict__ recs, int frames, int halo, int n_out_frames, int n_streams,
                  const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                  uint8_t *__restrict__ su_out, size_t su_frame_stride, size_t su_stream_stride) {
  __shared__ uint32_t s_words[kQpWarps][56];
  __shared__ double s_nf[kQpWarps][52];
  __shared__ uint16_t s_base[kQpWarps][52];
  __shared__ uint8_t s_bits[kQpWarps][52];
  __shared__ uint8_t s_bfu_long[512], s_bfu_short[512];
  __shared__ uint16_t s_start_long[52], s_start_short[52];
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const FormatTables &F = T->fmt;
  for (int i = tid; i < 512; i += kQpWarps * 32) { s_bfu_long[i] = F.bfu_of_long[i]; s_bfu_short[i] = F.bfu_of_short[i]; }
  if (tid < 52) { s_start_long[tid] = F.start_long[tid]; s_start_short[tid] = F.start_short[tid]; }
  __syncthreads();
  const long long n_units = (long long)n_streams * n_out_frames;
  const long long unit = (long long)blockIdx.x * kQpWarps + warp;
  if (unit >= n_units) return;
  const int stream = (int)(unit / n_out_frames);
  const int frame_out = (int)(unit % n_out_frames);
  const size_t su = (size_t)stream * frames + halo + frame_out;
  const AllocRec *r = recs + unit;
  const int n = r->n_bfu;
  uint32_t *words = s_words[warp];
  for (int i = lane; i < 56; i += 32) words[i] = 0;
  int m0, m1, m2, l0, l1, l2;
  if (P->use_fixed) { m0 = P->fixed[0]; m1 = P->fixed[1]; m2 = P->fixed[2]; }
  else { m0 = modes[su * 4]; m1 = modes[su * 4 + 1]; m2 = modes[su * 4 + 2]; }
  l0 = m0 == 0; l1 = m1 == 0; l2 = m2 == 0;
  // per-BFU bit widths, offsets (exclusive scan over bits*size) and 1/step factors
  int run = 16 + 10 * n;
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int b = lane + 32 * h;
    int wl = 0, sfi = 0, sz = 0;
    if (b < 52) { wl = r->wl[b]; sfi = r->sfi[b]; sz = F.specs[b]; }
    const int bits = b < n ? wl_bits(wl) : 0;
    int incl = bits * sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (b < 52) {
      s_base[warp][b] = (uint16_t)(run + incl - bits * sz);
      s_bits[warp][b] = (uint8_t)bits;
      s_nf[warp][b] = (bits > 0 && sfi > 0) ? (double)((1 << (bits - 1)) - 1) / T->sf[sfi] : 0.0;
      if (b < n) {
        put_bits(words, 16 + 4 * b, (uint32_t)wl, 4);
        put_bits(words, 16 + 4 * n + 6 * b, (uint32_t)sfi, 6);
      }
    }
    run += __shfl_sync(0xffffffffu, incl, 31);
  }
  if (lane == 0) {
    const int idx = n == 20 ? 0 : (n - 24) / 4;
    const uint32_t header = (((uint32_t)(2 - m0) << 14) | ((uint32_t)(2 - m1) << 12) |
                             ((uint32_t)(3 - m2) << 10) | ((uint32_t)idx << 5)) & 0xFFFFu;
    atomicOr(&words[0], header << 16);
  }
  __syncwarp();
  const float *src = coefs + su * 512;
#pragma unroll 4
  for (int k = 0; k < 16; k++) {
    const int cidx = lane + 32 * k;
    const int long_mode = k < 4 ? l0 : (k < 8 ? l1 : l2);
    const int b = long_mode ? s_bfu_long[cidx] : s_bfu_short[cidx];
    const int bits = s_bits[warp][b];
    if (bits == 0) continue;
    const double nf = s_nf[warp][b];
    int q = 0;
    if (nf != 0.0) {
      const int range = (1 << (bits - 1)) - 1;
      const double x = (double)src[cidx] * nf;
      const int y = js_to_int32(x + (x >= 0.0 ? 0.5 : -0.5));
      q = y > range ? range : (y < -range ? -range : y);
    }
    const int j = cidx - (long_mode ? s_start_long[b] : s_start_short[b]);
    put_bits(words, s_base[warp][b] + j * bits, (uint32_t)q & ((1u << bits) - 1u), bits);
  }
  __syncwarp();
  uint32_t *dst = reinterpret_cast<uint32_t *>(
      su_out + ((size_t)frame_out * su_frame_stride + (size_t)stream * su_stream_stride) * kSuBytes);
  for (int i = lane; i < kSuWords; i += 32) dst[i] = __byte_perm(words[i], 0, 0x0123);
}

// ------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------
size_t alloc_rec_bytes() { return sizeof(AllocRec); }

const char *kernel_name(int id) {
  static const char *names[K_COUNT] = {"qmf_analysis", "band_mags", "transient_modes", "mdct",
                                       "alloc", "quant_pack", "unpack_dequant", "imdct", "bands_time", "synth"};
  return id >= 0 && id < K_COUNT ? names[id] : "?";
}

cudaError_t launch_encode(const EncodeLaunch &L, cudaStream_t st, Prof *prof) {
  const int frames = L.frames_total;
  const int n_su = L.n_streams * frames;
  if (n_su == 0) return cudaSuccess;
  {
    dim3 grid((frames + kQmfTile - 1) / kQmfTile, L.n_streams);
    prof->begin(K_QMF_ANALYSIS, st);
    if (L.pcm_fmt == 0)
      qmf_analysis_kernel<0><<<grid, 256, 0, st>>>(L.pcm, L.row_stride, L.n_ch_interleave, L.valid_samples,
                                                  frames, L.tables, L.bands);
    else
      qmf_analysis_kernel<1><<<grid, 256, 0, st>>>(L.pcm, L.row_stride, L.n_ch_interleave, L.valid_samples,
                                                  frames, L.tables, L.bands);
    prof->end(K_QMF_ANALYSIS, st);
  }
  if (!L.use_fixed) {
    prof->begin(K_BAND_MAGS, st);
    band_mags_kernel<<<(n_su + 3) / 4, 128, 0, st>>>(L.bands, n_su, L.tables, L.mags);
    prof->end(K_BAND_MAGS, st);
    prof->begin(K_TRANSIENT_MODES, st);
    transient_modes_kernel<<<(n_su * 3 + 127) / 128, 128, 0, st>>>(L.mags, frames, n_su, L.tables, L.params,
                                                                  L.modes, L.scores);
    prof->end(K_TRANSIENT_MODES, st);
  }
  prof->begin(K_MDCT, st);
  mdct_kernel<<<(n_su + 3) / 4, 128, 0, st>>>(L.bands, L.modes, frames, n_su, L.tables, L.params, L.coefs);
  prof->end(K_MDCT, st);
  const long long n_units = (long long)L.n_streams * L.n_out_frames;
  if (n_units > 0 && L.su_out) {
    cudaError_t e = cudaFuncSetAttribute(alloc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(AlSmem));
    if (e != cudaSuccess) return e;
    AllocRec *recs = static_cast<AllocRec *>(L.alloc_recs);
    prof->begin(K_ALLOC, st);
    alloc_kernel<<<(unsigned)((n_units + kAlSu - 1) / kAlSu), kAlThreads, sizeof(AlSmem), st>>>(
        L.coefs, L.modes, frames, L.halo_frames, L.n_out_frames, L.n_streams, L.tables, L.params, recs);
    prof->end(K_ALLOC, st);
    prof->begin(K_QUANT_PACK, st);
    quant_pack_kernel<<<(unsigned)((n_units + kQpWarps - 1) / kQpWarps), kQpWarps * 32, 0, st>>>(
        L.coefs, L.modes, recs, frames, L.halo_frames, L.n_out_frames, L.n_streams, L.tables, L.params,
        L.su_out, L.su_frame_stride, L.su_stream_stride);
    prof->end(K_QUANT_PACK, st);
  }
  return cudaGetLastError();
}

}  // namespace c1
