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
// K4: scale factors, RDO bit allocation (literal max-heap emulation), quantisation and
// 212-byte packing.  One CTA = 128 threads = 16 sound units; 8 threads per unit run the 8
// candidate BFU counts' heaps concurrently.
//
// Heap entries are one 32-bit word: rank[31:22] | sfi[21:16] | wl[11:8] | bfu[5:0].  `rank`
// is the position of the entry's f32 priority among all distinct priorities of this
// encoder (DevEncParams::rank), so comparing ranks is comparing the reference's
// Float32Array priorities, ties included.  The heap is stored node-major / thread-minor, so
// the 32 lanes of a warp always hit 32 different banks.
// ------------------------------------------------------------------------------------
constexpr int kAqSu = 16;
constexpr int kAqThreads = 128;

struct AqSmem {
  float coef[kAqSu][512];
  uint32_t heap[52][kAqThreads];
  uint8_t wl[52][kAqThreads];
  double nf[kAqSu][52];
  float zero_bit[kAqSu][52];
  uint32_t words[kAqSu][56];
  uint16_t rank[1024];
  uint16_t base[kAqSu][52];
  uint8_t sfi[kAqSu][52];
  uint8_t wlf[kAqSu][52];
  uint8_t mode[kAqSu][4];
  int nbfu[kAqSu];
};

__device__ __forceinline__ void heap_sift(uint32_t *H, int start, int size) {  // bitallocation.js:314-341
  int i = start;
  const uint32_t v = H[i * kAqThreads];
  const uint32_t vr = v >> 22;
  for (;;) {
    const int l = 2 * i + 1;
    if (l >= size) break;
    const int r = l + 1;
    const uint32_t cl = H[l * kAqThreads];
    int max_i = i;
    uint32_t max_r = vr;
    uint32_t moved = cl;
    if ((cl >> 22) > max_r) { max_i = l; max_r = cl >> 22; }
    if (r < size) {
      const uint32_t cr = H[r * kAqThreads];
      if ((cr >> 22) > max_r) { max_i = r; moved = cr; }
    }
    if (max_i == i) break;
    H[i * kAqThreads] = moved;
    i = max_i;
  }
  H[i * kAqThreads] = v;
}

__device__ __forceinline__ void put_bits(uint32_t *words, int pos, uint32_t value, int bits) {
  const int w = pos >> 5, off = pos & 31;
  const unsigned long long v = (unsigned long long)value << (64 - off - bits);
  const uint32_t hi = (uint32_t)(v >> 32), lo = (uint32_t)v;
  if (hi) atomicOr(&words[w], hi);
  if (lo) atomicOr(&words[w + 1], lo);
}

__global__ void __launch_bounds__(kAqThreads)
alloc_quant_pack_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes, int frames,
                        int halo, int n_out_frames, int n_streams, const DevTables *__restrict__ T,
                        const DevEncParams *__restrict__ P, uint8_t *__restrict__ su_out,
                        size_t su_frame_stride, size_t su_stream_stride) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AqSmem &S = *reinterpret_cast<AqSmem *>(smem_raw);
  const int tid = threadIdx.x;
  const long long n_units = (long long)n_streams * n_out_frames;
  const long long unit0 = (long long)blockIdx.x * kAqSu;
  const FormatTables &F = T->fmt;

  // ---- load: coefficients, modes, rank table
  for (int i = tid; i < 1024; i += kAqThreads) S.rank[i] = P->rank[i];
  for (int u = 0; u < kAqSu; u++) {
    const long long unit = unit0 + u;
    if (unit >= n_units) break;
    const int stream = (int)(unit / n_out_frames);
    const int frame = halo + (int)(unit % n_out_frames);
    const size_t su = (size_t)stream * frames + frame;
    const float4 v = reinterpret_cast<const float4 *>(coefs + su * 512)[tid];
    reinterpret_cast<float4 *>(S.coef[u])[tid] = v;
    if (tid < 3) S.mode[u][tid] = P->use_fixed ? (uint8_t)(P->fixed[tid] != 0) : modes[su * 4 + tid];
  }
  __syncthreads();

  // ---- phase A: scale factor per BFU (bitallocation.js:290-299 via the exact threshold
  // table) and its zero-bit distortion (bitallocation.js:83-88)
  for (int item = tid; item < kAqSu * 52; item += kAqThreads) {
    const int u = item / 52, b = item - u * 52;
    if (unit0 + u >= n_units) continue;
    const int sz = F.specs[b];
    const int start = S.mode[u][band_of_bfu(b)] == 0 ? F.start_long[b] : F.start_short[b];
    float mx = 0.0f;
    for (int j = 0; j < sz; j++) {
      const float a = fabsf(S.coef[u][start + j]);
      if (a > mx) mx = a;
    }
    int sfi = 0;
    if (mx > 0.0f) {
#pragma unroll 7
      for (int k = 0; k < 63; k++) sfi += (mx > T->sf_thr[k]);
    }
    S.sfi[u][b] = (uint8_t)sfi;
    S.zero_bit[u][b] = sfi > 0 ? (float)(P->bsf[sfi] * 2.0 * (double)sz) : 0.0f;
  }
  __syncthreads();

  // ---- phase B: one thread per (unit, candidate BFU count)
  {
    const int u = tid >> 3, c = tid & 7;
    const int cand = c == 0 ? 20 : 24 + 4 * c;  // BFU_AMOUNTS = 20,28,32,...,52
    const bool live = unit0 + u < n_units;
    double total = 0.0;
    if (live) {
      int remaining = kFrameBits - 40 - 10 * cand;  // bitallocation.js:97-100
      uint32_t *H = &S.heap[0][tid];
      int count = 0;
      for (int b = 0; b < 52; b++) S.wl[b][tid] = 0;
      for (int b = 0; b < cand; b++) {  // bitallocation.js:216-232
        const uint32_t sfi = S.sfi[u][b];
        if (sfi) {
          H[count * kAqThreads] = ((uint32_t)S.rank[sfi * 16] << 22) | (sfi << 16) | (uint32_t)b;
          count++;
        }
      }
      if (count) {
        for (int i = (count >> 1) - 1; i >= 0; i--) heap_sift(H, i, count);
        int size = count;
        while (remaining > 0 && size > 0) {  // bitallocation.js:244-278
          uint32_t e = H[0];
          const int b = e & 63;
          int wl = (e >> 8) & 15;
          const int cost = (wl == 0 ? 2 : 1) * (int)F.specs[b];
          bool pop;
          if (cost > remaining) {
            pop = true;
          } else {
            remaining -= cost;
            wl++;
            e = (e & 0xFFFFF0FFu) | ((uint32_t)wl << 8);
            if (wl < 15) {
              const uint32_t sfi = (e >> 16) & 63;
              e = (e & 0x003FFFFFu) | ((uint32_t)S.rank[sfi * 16 + wl] << 22);
              H[0] = e;
              heap_sift(H, 0, size);
              pop = false;
            } else {
              pop = true;
            }
          }
          if (pop) {  // retired entries are parked behind the live heap
            size--;
            const uint32_t last = H[size * kAqThreads];
            H[size * kAqThreads] = e;
            if (size > 0) {
              H[0] = last;
              heap_sift(H, 0, size);
            }
          }
        }
        for (int i = 0; i < count; i++) {
          const uint32_t e = H[i * kAqThreads];
          S.wl[e & 63][tid] = (uint8_t)((e >> 8) & 15);
        }
      }
      // total distortion of this candidate (bitallocation.js:157-190), index order
      for (int i = 0; i < cand; i++) {
        const int bits = wl_bits(S.wl[i][tid]);
        if (bits == 0) { total += (double)S.zero_bit[u][i]; continue; }
        const int sfi = S.sfi[u][i];
        if (sfi == 0) continue;
        const double inv = __hiloint2double((1023 - bits) << 20, 0);
        total += P->bsf[sfi] * inv * (double)F.specs[i];
      }
      for (int i = cand; i < 52; i++) total += (double)S.zero_bit[u][i];
    }
    // first strict minimum over ascending candidates (bitallocation.js:122-129)
    const double inf = __longlong_as_double(0x7ff0000000000000ll);
    double key = (live && total < inf) ? total : inf;
    int best = c;
#pragma unroll
    for (int off = 4; off >= 1; off >>= 1) {
      const double ok = __shfl_xor_sync(0xffffffffu, key, off, 8);
      const int oi = __shfl_xor_sync(0xffffffffu, best, off, 8);
      if (ok < key || (ok == key && oi < best)) { key = ok; best = oi; }
    }
    if (live) {
      if (key < inf) {
        if (best == c) {
          S.nbfu[u] = cand;
          for (int b = 0; b < 52; b++) S.wlf[u][b] = b < cand ? S.wl[b][tid] : 0;
        }
      } else if (c == 0) {  // bitallocation.js:132-139
        S.nbfu[u] = 20;
        for (int b = 0; b < 52; b++) S.wlf[u][b] = 0;
      }
    }
    __syncwarp();
    if (live && !(key < inf) && c == 0)
      for (int b = 0; b < 52; b++) S.sfi[u][b] = 0;
  }
  __syncthreads();

  // ---- phase C: quantise (quantization.js:34-56) and pack (serialization.js:41-98)
  {
    const int warp = tid >> 5, lane = tid & 31;
    for (int uu = 0; uu < 4; uu++) {
      const int u = warp * 4 + uu;
      const long long unit = unit0 + u;
      if (unit >= n_units) break;
      const int n = S.nbfu[u];
      uint32_t *words = S.words[u];
      for (int i = lane; i < 56; i += 32) words[i] = 0;
      for (int b = lane; b < 52; b += 32) {
        const int bits = wl_bits(S.wlf[u][b]);
        const int sfi = S.sfi[u][b];
        double nf = 0.0;
        if (b < n && bits > 0 && sfi > 0) nf = (double)((1 << (bits - 1)) - 1) / T->sf[sfi];
        S.nf[u][b] = nf;
      }
      if (lane == 0) {
        int pos = 16 + 10 * n;
        for (int b = 0; b < n; b++) {
          S.base[u][b] = (uint16_t)pos;
          pos += wl_bits(S.wlf[u][b]) * (int)F.specs[b];
        }
      }
      __syncwarp();
      if (lane == 0) {
        const int idx = n == 20 ? 0 : (n - 24) / 4;
        int m0, m1, m2;
        if (P->use_fixed) { m0 = P->fixed[0]; m1 = P->fixed[1]; m2 = P->fixed[2]; }
        else { m0 = S.mode[u][0]; m1 = S.mode[u][1]; m2 = S.mode[u][2]; }
        const uint32_t header = (((uint32_t)(2 - m0) << 14) | ((uint32_t)(2 - m1) << 12) |
                                 ((uint32_t)(3 - m2) << 10) | ((uint32_t)idx << 5)) & 0xFFFFu;
        atomicOr(&words[0], header << 16);
      }
      for (int i = lane; i < n; i += 32) {
        put_bits(words, 16 + 4 * i, S.wlf[u][i], 4);
        put_bits(words, 16 + 4 * n + 6 * i, S.sfi[u][i], 6);
      }
      for (int k = 0; k < 16; k++) {
        const int cidx = lane + 32 * k;
        const int long_mode = S.mode[u][band_of_coef(cidx)] == 0;
        const int b = long_mode ? F.bfu_of_long[cidx] : F.bfu_of_short[cidx];
        if (b >= n) continue;
        const int bits = wl_bits(S.wlf[u][b]);
        if (bits == 0) continue;
        int q = 0;
        if (S.sfi[u][b] != 0) {
          const int range = (1 << (bits - 1)) - 1;
          const double x = (double)S.coef[u][cidx] * S.nf[u][b];
          const int y = js_to_int32(x + (x >= 0.0 ? 0.5 : -0.5));
          q = y > range ? range : (y < -range ? -range : y);
        }
        const int j = cidx - (long_mode ? F.start_long[b] : F.start_short[b]);
        put_bits(words, S.base[u][b] + j * bits, (uint32_t)q & ((1u << bits) - 1u), bits);
      }
      __syncwarp();
      const int stream = (int)(unit / n_out_frames);
      const int frame_out = (int)(unit % n_out_frames);
      uint32_t *dst = reinterpret_cast<uint32_t *>(
          su_out + ((size_t)frame_out * su_frame_stride + (size_t)stream * su_stream_stride) * kSuBytes);
      for (int i = lane; i < kSuWords; i += 32) dst[i] = __byte_perm(words[i], 0, 0x0123);
      __syncwarp();
    }
  }
}

// ------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------
size_t alloc_quant_pack_smem_bytes() { return sizeof(AqSmem); }

const char *kernel_name(int id) {
  static const char *names[K_COUNT] = {"qmf_analysis", "band_mags", "transient_modes", "mdct",
                                       "alloc_quant_pack", "unpack_dequant", "imdct", "bands_time", "synth"};
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
    cudaError_t e = cudaFuncSetAttribute(alloc_quant_pack_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)sizeof(AqSmem));
    if (e != cudaSuccess) return e;
    prof->begin(K_ALLOC_QUANT_PACK, st);
    alloc_quant_pack_kernel<<<(unsigned)((n_units + kAqSu - 1) / kAqSu), kAqThreads, sizeof(AqSmem), st>>>(
        L.coefs, L.modes, frames, L.halo_frames, L.n_out_frames, L.n_streams, L.tables, L.params, L.su_out,
        L.su_frame_stride, L.su_stream_stride);
    prof->end(K_ALLOC_QUANT_PACK, st);
  }
  return cudaGetLastError();
}

}  // namespace c1
