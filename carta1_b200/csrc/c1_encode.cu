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
#include "c1_fft.cuh"
#include "c1_launch.h"

#include <algorithm>
#include <atomic>
#include <mutex>

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
// K1: two-stage QMF analysis, streamed: a warp walks a run of consecutive frames of one row
// and carries the filter state in shared memory, exactly as the reference carries its delay
// lines from frame to frame (qmf.js:19-50, encoder.js:69-95).
//   S1lo/S1hi[n] = f32(E +- O), E = sum_j x[2n+1-2j]*EVEN[j], O = sum_j x[2n-2j]*ODD[j]
//   L/M[m]       = the same filter applied to S1lo;  H[n] = S1hi[n-39].
// fma() is exact-product here (both factors are widened f32), so it equals mul-then-add.
//
// A run that does not start at the row start is primed from the frame before it: the state
// after a frame (46 PCM samples, 46 S1lo values, 39 S1hi values) is a function of the last
// 174 samples of that frame alone.
//
// Layout: a signal is widened to binary64 once and kept as its two polyphase sequences
// xo[k] = x[2k+1], xe[k] = x[2k], each with 24 entries of history in front.  Stage 1: a lane
// produces 8 consecutive outputs, so a 24-tap filter needs a window of 31 values for 192 DFMA;
// element e lives at row e&7, column e>>3 (row stride == 2 mod 16 doubles): the warp's window
// reads and the frame fill are both bank-conflict free.  Stage 2: 4 consecutive outputs per
// lane, window of 27, element e at row e&3, column e>>2.
// ------------------------------------------------------------------------------------
constexpr int kQaWarps = 8, kQaCtasPerSm = 2;
// frames per run: chosen per launch (pick_run_len) so that the runs fill whole waves of the persistent warps
constexpr int kQaStride1 = 50;    // >= (24 + 256) / 8, == 2 mod 16
constexpr int kQaStride2 = 40;    // >= (24 + 128) / 4
static_assert(kQaStride1 * 8 >= 280 && kQaStride1 % 16 == 2 && kQaStride2 * 4 >= 152, "ring strides");
struct QaWarpSmem {
  double x[2][8 * kQaStride1];   // odd, even polyphase of the PCM frame
  double s[2][4 * kQaStride2];   // odd, even polyphase of S1lo
  float hi[40 + 256];            // S1hi with 40 entries of history (39 used)
  float raw[512];                // the next PCM frame, in flight (TMA bulk copy) while this one is filtered
  unsigned long long mbar, pad;  // the warp's mbarrier: completion of the copy into raw
};
static_assert(sizeof(QaWarpSmem) % 16 == 0 && (sizeof(double) * (2 * 8 * kQaStride1 + 2 * 4 * kQaStride2) + 4 * 296) % 16 == 0, "raw is 16-byte aligned");
constexpr size_t kQaSmemBytes = sizeof(QaWarpSmem) * kQaWarps;

__constant__ double c_qmf_even[24];
__constant__ double c_qmf_odd[24];

cudaError_t upload_encode_constants(const DevTables *host_tables) {
  cudaError_t e = cudaMemcpyToSymbol(c_qmf_even, host_tables->qmf_even, 24 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_qmf_odd, host_tables->qmf_odd, 24 * sizeof(double));
  if (e == cudaSuccess) e = cudaMemcpyToSymbol(c_fft_tw, host_tables->fft_tw, sizeof(host_tables->fft_tw));
  return e;
}

// acc[r] = sum_j seq[kR*t + 24 + r - j] * taps[j], j ascending, r = 0..kR-1: element kR*t + i
// sits at row i & (kR-1), column t + i / kR of a kR-row ring
template <int kR, int kStride>
__device__ __forceinline__ void fir_analysis(const double *__restrict__ seq, int t, const double *taps,
                                             double (&acc)[kR]) {
#pragma unroll
  for (int r = 0; r < kR; r++) acc[r] = 0.0;
#pragma unroll
  for (int j = 0; j < 24; j++) {
    const double c = taps[j];
#pragma unroll
    for (int r = 0; r < kR; r++) {
      const int i = r - j + 24;
      acc[r] = fma(seq[(i & (kR - 1)) * kStride + t + i / kR], c, acc[r]);
    }
  }
}

// 16-byte asynchronous copy global -> shared; bytes beyond src_bytes are zero-filled.
__device__ __forceinline__ void cp_async16_zfill(void *smem_dst, const void *gsrc, int src_bytes) {
  asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
               "l"(gsrc), "r"(src_bytes) : "memory");
}

// f32 planar, 16-byte aligned rows: the frame starting at sample `first` is fetched into the warp's raw staging
// buffer while the previous frame is being filtered.  A whole frame (every frame but a row's last) is one TMA bulk
// copy of 2 KB issued by lane 0, completion on the warp's mbarrier; a partial frame takes 16-byte cp.async chunks
// with zero fill.  The caller issues it after the warp barrier that follows qa_fill (every lane has read raw).
// Returns how the fetch completes: 2 = mbarrier phase, 1 = cp.async group.
__device__ __forceinline__ int qa_prefetch(QaWarpSmem &S, const float *__restrict__ row, long long first,
                                           long long valid_samples, int lane) {
  if (first + 512 <= valid_samples) {
    if (lane == 0) {
      asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // the lanes' reads of raw, before the TMA unit rewrites it
      bulk_g2s((uint32_t)__cvta_generic_to_shared(S.raw), row + first, 2048u, (uint32_t)__cvta_generic_to_shared(&S.mbar));
    }
    return 2;
  }
#pragma unroll
  for (int k = 0; k < 4; k++) {
    const long long g = first + 4 * lane + 128 * k;
    const long long left = valid_samples - g;  // samples that exist from g on
    const int bytes = left >= 4 ? 16 : (left > 0 ? (int)left * 4 : 0);
    cp_async16_zfill(S.raw + 4 * lane + 128 * k, row + (bytes ? g : 0), bytes);
  }
  asm volatile("cp.async.commit_group;" ::: "memory");
  return 1;
}

// Per-lane ring positions, computed once per kernel (the compiler otherwise re-derives them every frame).
struct QaLane {
  int fill;   // sample 4 lane of a frame in its polyphase ring: element 24 + 2 lane -> ((2 lane) & 7) * stride + 3 + (2 lane >> 3)
  int keep_x; // element `lane` of the x rings (element 256 + lane is 32 columns further)
  int keep_s; // element `lane` of the s rings (element 128 + lane is 32 columns further)
  __device__ __forceinline__ explicit QaLane(int lane)
      : fill(((2 * lane) & 7) * kQaStride1 + 3 + ((2 * lane) >> 3)),
        keep_x((lane & 7) * kQaStride1 + (lane >> 3)),
        keep_s((lane & 3) * kQaStride2 + (lane >> 2)) {}
};

// PCM frame -> polyphase ring (elements 24..279).  Lane l owns samples 4l + 128k + c.  pending: how the frame's
// fetch into S.raw completes (qa_prefetch), 0 = not prefetched (unaligned rows, int16); parity: the mbarrier's phase.
template <int kFmt>
__device__ __forceinline__ void qa_fill(QaWarpSmem &S, const void *__restrict__ pcm_v, size_t row_off, int n_ch,
                                        int stream, long long first, long long valid_samples, int pending, uint32_t &parity,
                                        int lane, const QaLane &Q) {
  float v[4][4];
  if (kFmt == 0 && pending) {
    if (pending == 2) {
      mbar_wait((uint32_t)__cvta_generic_to_shared(&S.mbar), parity);
      parity ^= 1u;
    } else {
      asm volatile("cp.async.wait_group 0;" ::: "memory");
    }
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const float4 q = *reinterpret_cast<const float4 *>(S.raw + 4 * lane + 128 * k);
      v[k][0] = q.x; v[k][1] = q.y; v[k][2] = q.z; v[k][3] = q.w;
    }
  } else {
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const long long g = first + 4 * lane + 128 * k;
      if (kFmt == 0) {
        const float *src = static_cast<const float *>(pcm_v) + row_off + g;
#pragma unroll
        for (int c = 0; c < 4; c++) v[k][c] = g + c < valid_samples ? __ldg(src + c) : 0.0f;
      } else {  // bin/cli.js:395  readInt16LE / 32768.0 -> Float32Array
        const short *src = static_cast<const short *>(pcm_v);
#pragma unroll
        for (int c = 0; c < 4; c++)
          v[k][c] = g + c < valid_samples ? (float)((double)__ldg(src + (size_t)(g + c) * n_ch + stream) / 32768.0) : 0.0f;
      }
    }
  }
  // sample 4l + 128k + c is element e = 24 + 2l + 64k + (c >> 1) of polyphase (c & 1 ? 0 : 1); its ring position
  // (e & 7) * stride + (e >> 3) is Q.fill + stride * (c >> 1) + 8k
  double *xb = S.x[0] + Q.fill;
#pragma unroll
  for (int k = 0; k < 4; k++)
#pragma unroll
    for (int c = 0; c < 4; c++)
      xb[((c & 1) ? 0 : 8 * kQaStride1) + kQaStride1 * (c >> 1) + 8 * k] = (double)v[k][c];
}

// stage 1 of the frame in the ring: S1lo -> stage-2 ring (elements 24..151), S1hi -> hi ring (40..295)
__device__ __forceinline__ void qa_stage1(QaWarpSmem &S, int lane) {
  double e[8], o[8];
  fir_analysis<8, kQaStride1>(S.x[0], lane, c_qmf_even, e);
  fir_analysis<8, kQaStride1>(S.x[1], lane, c_qmf_odd, o);
  float hi[8];
#pragma unroll
  for (int r = 0; r < 8; r++) {
    const float lo = (float)(e[r] + o[r]);
    hi[r] = (float)(e[r] - o[r]);
    // S1lo[8 lane + r] feeds stage 2: polyphase r & 1, element 24 + 4 lane + (r >> 1) = 4 (lane + 6) + (r >> 1)
    S.s[(r & 1) ? 0 : 1][(r >> 1) * kQaStride2 + lane + 6] = (double)lo;
  }
  float4 *h4 = reinterpret_cast<float4 *>(S.hi + 40 + 8 * lane);
  h4[0] = make_float4(hi[0], hi[1], hi[2], hi[3]);
  h4[1] = make_float4(hi[4], hi[5], hi[6], hi[7]);
}

// frame end: the last 24 polyphase entries / 40 S1hi values become the history of the next frame
__device__ __forceinline__ void qa_shift(QaWarpSmem &S, int lane, const QaLane &Q) {
  double keep_x[2], keep_s[2];
  float keep_h[2];
  if (lane < 24) {
#pragma unroll
    for (int p = 0; p < 2; p++) {
      keep_x[p] = S.x[p][Q.keep_x + 32];
      keep_s[p] = S.s[p][Q.keep_s + 32];
    }
  }
  keep_h[0] = S.hi[256 + lane];
  keep_h[1] = lane < 8 ? S.hi[288 + lane] : 0.0f;
  __syncwarp();
  if (lane < 24) {
#pragma unroll
    for (int p = 0; p < 2; p++) {
      S.x[p][Q.keep_x] = keep_x[p];
      S.s[p][Q.keep_s] = keep_s[p];
    }
  }
  S.hi[lane] = keep_h[0];
  if (lane < 8) S.hi[32 + lane] = keep_h[1];
}

template <int kFmt>  // 0: f32 planar rows, 1: s16 interleaved
__global__ void __launch_bounds__(kQaWarps * 32, kQaCtasPerSm)
qmf_analysis_kernel(const void *__restrict__ pcm_v, size_t row_stride, int n_ch, long long valid_samples,
                    int frames, int n_streams, int run_len, float *__restrict__ bands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  QaWarpSmem &S = reinterpret_cast<QaWarpSmem *>(smem_raw)[warp];
  const QaLane Q(lane);
  uint32_t parity = 0;
  if (lane == 0) mbar_init((uint32_t)__cvta_generic_to_shared(&S.mbar));
  __syncwarp();
  const int runs_per_row = (frames + run_len - 1) / run_len;
  const int n_runs = runs_per_row * n_streams;
  for (int run = blockIdx.x * kQaWarps + warp; run < n_runs; run += gridDim.x * kQaWarps) {
    const int stream = run / runs_per_row;
    const int f0 = (run - stream * runs_per_row) * run_len;
    const int f1 = min(f0 + run_len, frames);
    const size_t row_off = kFmt == 0 ? (size_t)stream * row_stride : 0;
    const float *row = static_cast<const float *>(pcm_v) + row_off;  // kFmt == 0 only
    const bool async_ok = kFmt == 0 && ((reinterpret_cast<uintptr_t>(row) & 15) == 0);
    __syncwarp();
    int pending = 0;
    if (async_ok) pending = qa_prefetch(S, row, 512ll * (f0 > 0 ? f0 - 1 : f0), valid_samples, lane);
    if (f0 == 0) {  // row start: silent history (new BufferPool, buffers.js:31-42)
      if (lane < 24) {
#pragma unroll
        for (int p = 0; p < 2; p++) {
          S.x[p][Q.keep_x] = 0.0;
          S.s[p][Q.keep_s] = 0.0;
        }
      }
      S.hi[lane] = 0.0f;
      if (lane < 8) S.hi[32 + lane] = 0.0f;
    } else {        // prime the state from the frame before the run (only its last 64 S1 outputs matter)
      qa_fill<kFmt>(S, pcm_v, row_off, n_ch, stream, 512ll * (f0 - 1), valid_samples, pending, parity, lane, Q);
      __syncwarp();
      if (async_ok) pending = qa_prefetch(S, row, 512ll * f0, valid_samples, lane);
      qa_stage1(S, lane);
      __syncwarp();
      qa_shift(S, lane, Q);
    }
    for (int f = f0; f < f1; f++) {
      __syncwarp();
      qa_fill<kFmt>(S, pcm_v, row_off, n_ch, stream, 512ll * f, valid_samples, pending, parity, lane, Q);
      __syncwarp();
      if (async_ok && f + 1 < f1) pending = qa_prefetch(S, row, 512ll * (f + 1), valid_samples, lane);
      qa_stage1(S, lane);
      __syncwarp();
      float *out = bands + onchip_row((size_t)stream * frames + f) * 512;
      {  // stage 2: low / mid samples 4 lane .. 4 lane + 3
        double e[4], o[4];
        fir_analysis<4, kQaStride2>(S.s[0], lane, c_qmf_even, e);
        fir_analysis<4, kQaStride2>(S.s[1], lane, c_qmf_odd, o);
        reinterpret_cast<float4 *>(out)[lane] = make_float4((float)(e[0] + o[0]), (float)(e[1] + o[1]),
                                                            (float)(e[2] + o[2]), (float)(e[3] + o[3]));
        reinterpret_cast<float4 *>(out + 128)[lane] = make_float4((float)(e[0] - o[0]), (float)(e[1] - o[1]),
                                                                  (float)(e[2] - o[2]), (float)(e[3] - o[3]));
      }
      // high band delayed by 39 (encoder.js:84-90): H[n] = S1hi[n - 39] = ring[n + 1]; lane l emits
      // H[8l .. 8l + 7] = ring[8l + 1 .. 8l + 8] from three aligned 16-byte reads
      {
        const float4 *hp = reinterpret_cast<const float4 *>(S.hi) + 2 * lane;
        const float4 a = hp[0], b = hp[1], c = hp[2];
        float4 *o4 = reinterpret_cast<float4 *>(out + 256) + 2 * lane;
        o4[0] = make_float4(a.y, a.z, a.w, b.x);
        o4[1] = make_float4(b.y, b.z, b.w, c.x);
      }
      __syncwarp();
      qa_shift(S, lane, Q);
    }
  }
}

// ------------------------------------------------------------------------------------
// K2a: magnitude spectra and spectrum features for transient detection, one warp per sound
// unit (transient.js:17-35,116-189).  performFFT runs the complex FFT on the real band: 128
// points for low and mid (16 lanes each, concurrently), 256 for high (32 lanes), in the
// in-thread formulation of c1_fft.cuh.  Exact shortcuts that only use what the reference
// arithmetic itself produces:
//   - the imaginary parts start as +0, so the stage-0 butterflies and the k = 0 butterflies of
//     stage 1 (twiddle (1, 0)) are real: t = (b.re * 1 - (+0) * 0, b.re * 0 + (+0) * 1) = (b.re, +0);
//   - only bins below N/2 are read (transient.js:29), i.e. the "even" outputs of the last stage.
// The per-band sums of spectrum_features run serially in index order as the reference's loops
// do, one lane per accumulator; the logarithms they add are computed by all lanes first.
// ------------------------------------------------------------------------------------
template <typename R>
__device__ __forceinline__ void butterfly_real(Cplx &a, Cplx &b, R &rnd) {  // im(a) = im(b) = +0, twiddle (1, 0)
  const double er = a.re, tr = b.re;
  a.re = rnd.r0(er + tr);
  b.re = rnd.r2(er - tr);
}
template <typename R>
__device__ __forceinline__ void butterfly_even_only(Cplx &a, const Cplx &b, const double2 w, R &rnd) {
  const double tr = b.re * w.x - b.im * w.y;
  const double ti = b.re * w.y + b.im * w.x;
  a.re = rnd.r0(a.re + tr);
  a.im = rnd.r1(a.im + ti);
}

// stages 0..2 on positions 8t + j of a real signal
template <typename R>
__device__ __forceinline__ void fft8_pass_a_real(Cplx (&v)[8], R &rnd) {
#pragma unroll
  for (int j = 0; j < 8; j += 2) butterfly_real(v[j], v[j + 1], rnd);  // stage 0
  butterfly_real(v[0], v[2], rnd);                                    // stage 1, k = 0
  butterfly_real(v[4], v[6], rnd);
  butterfly(v[1], v[3], c_fft_tw[2], rnd);                            // stage 1, k = 1
  butterfly(v[5], v[7], c_fft_tw[2], rnd);
#pragma unroll
  for (int j = 0; j < 4; j++) butterfly(v[j], v[j + 4], c_fft_tw[3 + j], rnd);  // stage 2
}

#ifndef C1_TS_WARPS
#define C1_TS_WARPS 8
#define C1_TS_CTAS 3
#endif
constexpr int kTsWarps = C1_TS_WARPS, kTsCtasPerSm = C1_TS_CTAS;
struct TsWarpSmem {
  double2 xbuf[256 + 32];  // transpose buffer (slot p + p/8)
  double logm[256 + 12];   // log(magnitude) where magnitude > 1e-10 (term arrays are skewed, see ts_skew)
  float mag[256];          // low 64 | mid 64 | high 128
};
constexpr size_t kTsSmemBytes = sizeof(TsWarpSmem) * kTsWarps;

// one band's FFT on kLanes = N/8 lanes of the warp; bins [0, N/2) -> mag[]
template <int kLog2N, typename R>
__device__ __forceinline__ void transient_fft(const float *__restrict__ x, double2 *xbuf, float *mag, int t,
                                              const double2 *__restrict__ tw, R &rnd) {
  constexpr int kN = 1 << kLog2N, kLanes = kN / 8;
  const int rev_t = (int)(__brev((unsigned)t) >> (32 - (kLog2N - 3)));
  const int u = t & 7, b = t >> 3;
  Cplx v[8];
#pragma unroll
  for (int j = 0; j < 8; j++) {  // position 8t + j holds x[brev(8t + j)] = x[brev3(j) * kLanes + rev_t]
    v[j].re = (double)__ldg(x + ((((j & 1) << 2) | (j & 2) | ((j >> 2) & 1)) * kLanes) + rev_t);
    v[j].im = 0.0;
  }
  fft8_pass_a_real(v, rnd);
  __syncwarp();
#pragma unroll
  for (int j = 0; j < 8; j++) xbuf[9 * t + j] = make_double2(v[j].re, v[j].im);
  __syncwarp();
  double2 *at_b = xbuf + 72 * b + u;  // positions 64 b + 8m + u
#pragma unroll
  for (int m = 0; m < 8; m++) {
    const double2 z = at_b[9 * m];
    v[m].re = z.x;
    v[m].im = z.y;
  }
  fft8_pass_b(v, tw + u, rnd);
  __syncwarp();
#pragma unroll
  for (int m = 0; m < 8; m++) at_b[9 * m] = make_double2(v[m].re, v[m].im);
  __syncwarp();
  if (kLog2N == 7) {
    // stage 6: thread (u, h = b) pairs positions 8 (4h + k) + u and + 64; only the former is a bin < 64
    const double2 *at_c = xbuf + 36 * b + u;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const double2 lo = at_c[9 * k], hi = at_c[9 * k + 72];
      Cplx a, o;
      a.re = lo.x; a.im = lo.y; o.re = hi.x; o.im = hi.y;
      butterfly_even_only(a, o, __ldg(tw + 63 + 32 * b + 8 * k + u), rnd);
      mag[32 * b + 8 * k + u] = (float)sqrt(a.re * a.re + a.im * a.im);  // transient.js:31
    }
  } else {
    // stages 6, 7: thread (u, h = b in 0..3) owns p0 = 8 (2h + e) + u, e = 0..1, and p0 + 64 c, c = 0..3
    const double2 *at_c = xbuf + 18 * b + u;
#pragma unroll
    for (int e = 0; e < 2; e++) {
      Cplx c0, c1, c2, c3;
      { const double2 z = at_c[9 * e]; c0.re = z.x; c0.im = z.y; }
      { const double2 z = at_c[9 * e + 72]; c1.re = z.x; c1.im = z.y; }
      { const double2 z = at_c[9 * e + 144]; c2.re = z.x; c2.im = z.y; }
      { const double2 z = at_c[9 * e + 216]; c3.re = z.x; c3.im = z.y; }
      const int k6 = 16 * b + 8 * e + u;  // p & 63
      const double2 w6 = __ldg(tw + 63 + k6);
      butterfly(c0, c1, w6, rnd);  // stage 6: (p0, p0 + 64), (p0 + 128, p0 + 192)
      butterfly(c2, c3, w6, rnd);
      butterfly_even_only(c0, c2, __ldg(tw + 127 + k6), rnd);       // stage 7: (p0, p0 + 128)
      butterfly_even_only(c1, c3, __ldg(tw + 127 + 64 + k6), rnd);  //          (p0 + 64, p0 + 192)
      mag[k6] = (float)sqrt(c0.re * c0.re + c0.im * c0.im);
      mag[64 + k6] = (float)sqrt(c1.re * c1.re + c1.im * c1.im);
    }
  }
}

// Warps are specialised by role, each role its own kernel (as in K3: the unrolled transforms of both
// sizes in one kernel overflow the instruction cache, `no_instruction` stalls 0.65 per issue): role 0
// takes the low and mid bands of two sound units (four FFT128, two at a time on 16 lanes each), role 1
// their high bands (two FFT256).  S.mag holds [unit 0: 128 magnitudes | unit 1: 128 magnitudes].
template <int kRole, typename R>
__device__ __forceinline__ void transient_ffts(const float *__restrict__ bands, int su0, bool have1, TsWarpSmem &S,
                                               int lane, const DevTables *__restrict__ T) {
  R rnd;
#pragma unroll 1  // one copy of the unrolled transform: the kernel has to stay inside the instruction cache
  for (int u = 0; u < 2; u++) {
    if (u == 1 && !have1) break;
    const float *band = bands + (size_t)(su0 + u) * 512;
    if (kRole == 0)
      transient_fft<7>(band + 128 * (lane >> 4), S.xbuf + 144 * (lane >> 4), S.mag + 128 * u + 64 * (lane >> 4), lane & 15,
                       T->fft_tw, rnd);
    else
      transient_fft<8>(band + 256, S.xbuf, S.mag + 128 * u, lane, T->fft_tw, rnd);
    __syncwarp();
  }
}
template <int kRole>
__device__ __noinline__ void transient_ffts_exact(const float *__restrict__ bands, int su0, bool have1, TsWarpSmem &S,
                                                  int lane, const DevTables *__restrict__ T) {
  transient_ffts<kRole, ExactRound>(bands, su0, have1, S, lane, T);
}

struct SpectrumFeatures {  // transient.js:116-189 of one band's magnitude spectrum
  double flatness, hf_ratio, energy;
};

template <int kRole>
__global__ void __launch_bounds__(kTsWarps * 32, kTsCtasPerSm)
transient_spectrum_kernel(const float *__restrict__ bands, int n_su, const DevTables *__restrict__ T,
                          float *__restrict__ mags, SpectrumFeatures *__restrict__ feats) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  TsWarpSmem &S = reinterpret_cast<TsWarpSmem *>(smem_raw)[warp];
  constexpr int kOff = kRole == 0 ? 0 : 256;  // first band sample / 2x first magnitude of the role
  const int n_pairs = (n_su + 1) >> 1;
  for (int pair = blockIdx.x * kTsWarps + warp; pair < n_pairs; pair += gridDim.x * kTsWarps) {
    const int su0 = 2 * pair;
    const bool have1 = su0 + 1 < n_su;
    unsigned big = 0;
#pragma unroll
    for (int k = 0; k < 4; k++) {  // the role's 256 samples of each unit: 2 float4 per lane and unit
      if (k >= 2 && !have1) break;
      const float4 q = __ldg(reinterpret_cast<const float4 *>(bands + (size_t)(su0 + (k >> 1)) * 512 + kOff) + lane + 32 * (k & 1));
      big = max(big, max(max(__float_as_uint(q.x) & 0x7FFFFFFFu, __float_as_uint(q.y) & 0x7FFFFFFFu),
                         max(__float_as_uint(q.z) & 0x7FFFFFFFu, __float_as_uint(q.w) & 0x7FFFFFFFu)));
    }
    __syncwarp();
    if (__reduce_max_sync(0xffffffffu, big) < 0x71800000u) transient_ffts<kRole, FastRound>(bands, su0, have1, S, lane, T);
    else transient_ffts_exact<kRole>(bands, su0, have1, S, lane, T);
    __syncwarp();
    // magnitudes out; per-bin terms of the sums (logarithm, magnitude, square), computed by all lanes
    // (entry lane + 32k of S.mag); the term arrays reuse the transpose buffer.  The reference skips the
    // bins at or below EPS in the first two sums (transient.js:126-131); here those bins contribute +0.0,
    // which leaves a sum that started at +0.0 bit-identical (such a sum is never -0.0: an exact
    // cancellation rounds to +0.0 and fdlibm's log(1) is +0.0), so the serial loops need no masks.
    const double EPS = 1e-10;
    // Term array a (0 log, 1 magnitude, 2 square) of group g is shifted by 3 g + a doubles, so that the lanes
    // of the serial loops (one per group and accumulator, all at the same i) read 12 different bank pairs;
    // unshifted, the three arrays and the groups all start on bank 0.
    double *t_mag = reinterpret_cast<double *>(S.xbuf), *t_sq = t_mag + 272;
    constexpr int kN = kRole == 0 ? 64 : 128, kGroups = 256 / kN, kMid = kN / 2;
    const int g = (lane >> 2) < kGroups ? (lane >> 2) : 0, acc = lane & 3;
    int valid = 0;  // bins above EPS in this lane's group (entries kN g .. kN g + kN - 1)
#pragma unroll 2  // (fd::log is inlined: eight copies would be a sixth of the kernel)
    for (int k = 0; k < 8; k++) {
      const int i = lane + 32 * k;
      const int u = i >> 7;
      const float m = (u == 0 || have1) ? S.mag[i] : 0.0f;
      if (u == 0 || have1) mags[(size_t)(su0 + u) * 256 + (kOff >> 1) + (i & 127)] = m;
      const double v = (double)m, md = fabs(v);
      const bool ok = md > EPS;
      const int sk = i + 3 * (i / kN);
      S.logm[sk] = ok ? fd::log(md) : 0.0;
      t_mag[sk + 1] = ok ? md : 0.0;
      t_sq[sk + 2] = v * v;
      const unsigned okm = __ballot_sync(0xffffffffu, ok);
      if (k / (kN / 32) == g) valid += __popc(okm);
    }
    __syncwarp();
    // serial sums in index order, one lane per (unit, band, accumulator): 0 sum_log (+ valid count),
    // 1 sum_lin, 2 lo then hi, 3 energy.  Every lane runs the same loop over its own term array.
    double r0 = 0.0, r1 = 0.0;
    {
      const double *term = (acc == 0 ? S.logm : acc == 1 ? t_mag + 1 : t_sq + 2) + (kN + 3) * g;
#pragma unroll 8
      for (int i = 0; i < kMid; i++) r0 += term[i];
      double rr = acc == 2 ? 0.0 : r0;
#pragma unroll 8
      for (int i = kMid; i < kN; i++) rr += term[i];
      if (acc == 2) r1 = rr; else r0 = rr;
    }
    // gather the four lanes of a group on its first lane
    const int base = lane & ~3;
    const double sum_log = __shfl_sync(0xffffffffu, r0, base), sum_lin = __shfl_sync(0xffffffffu, r0, base + 1);
    const double lo = __shfl_sync(0xffffffffu, r0, base + 2), hi = __shfl_sync(0xffffffffu, r1, base + 2);
    const double energy = __shfl_sync(0xffffffffu, r0, base + 3);
    const int nvalid = __shfl_sync(0xffffffffu, valid, base);
    const int unit = kRole == 0 ? g >> 1 : g, band_i = kRole == 0 ? g & 1 : 2;
    if ((lane >> 2) < kGroups && acc == 0 && (unit == 0 || have1)) {
      SpectrumFeatures f;
      if (nvalid == 0) {
        f.flatness = 0.0;
      } else {
        const double geo = fd::exp(sum_log / nvalid);
        const double arith = sum_lin / nvalid;
        f.flatness = arith > EPS ? geo / arith : 0.0;
      }
      const double total = lo + hi;
      f.hf_ratio = total > 0.0 ? hi / total : 0.0;
      f.energy = energy;
      feats[(size_t)(su0 + unit) * 3 + band_i] = f;
    }
    __syncwarp();
  }
}

// ------------------------------------------------------------------------------------
// K2b: transient score (transient.js:44-112,197-226), one thread per (sound unit, band): the
// spectral-flux sum (serial, index order) and the four-feature score.
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

// A warp takes 8 consecutive sound units: their magnitude rows (and the row before the first) are
// loaded coalesced into shared memory, then lane l < 24 runs unit l / 3, band l % 3.  Row stride
// 273 floats and 8 floats of padding in front of each band keep the 24 lanes' serial walks on 24
// different banks.
constexpr int kTmWarps = 4, kTmUnits = 8, kTmRow = 273;
__device__ __forceinline__ int tm_off(int band) { return band == 0 ? 0 : band == 1 ? 64 + 8 : 128 + 16; }

__global__ void __launch_bounds__(kTmWarps * 32)
transient_modes_kernel(const float *__restrict__ mags, const SpectrumFeatures *__restrict__ feats, int frames, int halo,
                       int n_su, const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                       uint8_t *__restrict__ modes, double *__restrict__ scores, unsigned long long *__restrict__ near) {
  __shared__ float s_rows[kTmWarps][(kTmUnits + 1) * kTmRow];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  float *rows = s_rows[warp];
  const int n_groups = (n_su + kTmUnits - 1) / kTmUnits;
  for (int group = blockIdx.x * kTmWarps + warp; group < n_groups; group += gridDim.x * kTmWarps) {
    const int su0 = group * kTmUnits;
    __syncwarp();
    // rows[r] = magnitudes of unit su0 - 1 + r, r = 0..8 (64 float4 per row: lane covers two of them)
    for (int r = 0; r <= kTmUnits; r++) {
      const int su = su0 - 1 + r;
      if (su < 0 || su >= n_su) continue;
      const float4 *src = reinterpret_cast<const float4 *>(mags + (size_t)su * 256);
#pragma unroll
      for (int k = 0; k < 2; k++) {
        const int q = lane + 32 * k;  // float4 index: bins 4q .. 4q + 3
        const float4 v = __ldg(src + q);
        const int band = q < 16 ? 0 : q < 32 ? 1 : 2;
        float *dst = rows + r * kTmRow + 4 * q + 8 * band;
        dst[0] = v.x; dst[1] = v.y; dst[2] = v.z; dst[3] = v.w;
      }
    }
    __syncwarp();
    const int u = lane / 3, band = lane - 3 * u;
    const int su = su0 + u;
    if (lane < 3 * kTmUnits && su < n_su) {
      const int frame = su % frames;
      const int n = band == 2 ? 128 : 64;
      const float *cur = rows + (u + 1) * kTmRow + tm_off(band);
      const float *prev = cur - kTmRow;
      const bool has_prev = frame > 0;  // frame 0: all-zero previous spectrum
      double flux = 0.0;
      for (int i = 0; i < n; i++) {  // transient.js:92-112
        const double c = fabs((double)cur[i]);
        const double p = has_prev ? fabs((double)prev[i]) : 0.0;
        const double d = c - p;
        if (d > 0.0) flux += d;
      }
      const SpectrumFeatures fc = feats[(size_t)su * 3 + band];
      SpectrumFeatures fp;  // features of the all-zero spectrum: no valid bin, no energy
      fp.flatness = 0.0; fp.hf_ratio = 0.0; fp.energy = 0.0;
      if (has_prev) fp = feats[(size_t)(su - 1) * 3 + band];
      // the flux loop's own energy sum adds the same squares in the same order as the features' one
      double norm = sqrt(fc.energy);
      if (norm == 0.0 || isnan(norm)) norm = 1e-6;
      const double spectral_flux = flux / norm;
      const double flat_change = fabs(fc.flatness - fp.flatness);
      const double hf_change = fabs(fc.hf_ratio - fp.hf_ratio);
      const double ce = js_max(fc.energy, 1e-10), pe = js_max(fp.energy, 1e-10);  // :182-183
      const double db = 10.0 * fd::log10(ce / pe);
      const double e_change = js_max(0.0, db);
      const double flat_c = sqrt(flat_change);
      const double hf_c = fd::log1p(hf_change * 10.0) / T->log1p10;
      const double e_c = js_min(e_change / 30.0, 1.0);
      const double score = (spectral_flux + flat_c + hf_c + e_c) / 4.0;
      // every band compares against transientThresholdLow (encoder.js:137-141); mode = t*max(b+1,2)
      const int transient = score > P->threshold;
      modes[(size_t)su * 4 + band] = (uint8_t)(transient ? (band == 2 ? 3 : 2) : 0);
      if (scores) scores[(size_t)su * 3 + band] = score;
      // Close calls of `score > threshold` (transient.js:54): the score goes through log / exp / log10 / log1p, the one
      // place where a libm that is not V8's could flip a decision; counted for emitted frames so
      // the risk is observable (carta1_ctx_near_threshold).
      const double gap = fabs(score - P->threshold);
      if (near && frame >= halo && gap < 1e-9) {
        atomicAdd(near, 1ull);
        if (gap < 1e-12) atomicAdd(near + 1, 1ull);
      }
    }
  }
}

// ------------------------------------------------------------------------------------
// K3: windowed MDCT, one warp per sound unit (mdct.js:54-122, encoder.js:228-316).
// Long blocks: FFT64 (low, mid) / FFT128 (high) in registers; short blocks: four FFT16 per
// warp pass.  All values are binary64 registers holding f32-representable numbers
// (c1_fft.cuh); the only f64->f32 conversions are the final coefficient stores.
// ------------------------------------------------------------------------------------
// mdct.js:76-105: pre-twiddle of FFT input q (natural order) of an N-point MDCT
template <typename In, typename R>
__device__ __forceinline__ Cplx mdct_pre(int q, int n, const In &in, const double *__restrict__ tab,
                                         R &rnd) {
  const int i = 2 * q, n4 = n >> 2, n34 = 3 * n4;
  double r, m;
  if (i < n4) {
    r = in(n34 - 1 - i) + in(n34 + i);
    m = in(n4 + i) - in(n4 - 1 - i);
  } else {
    r = in(n34 - 1 - i) - in(i - n4);
    m = in(n4 + i) + in(5 * n4 - 1 - i);
  }
  const double c = __ldg(&tab[i]), s = __ldg(&tab[i + 1]);
  Cplx z;
  z.re = rnd.r0(r * c + m * s);
  z.im = rnd.r1(m * c - r * s);
  return z;
}

// mdct.js:111-119: FFT output i -> coefficients 2i and N/2-1-2i (spectrum reversed for
// bands 1 and 2, utils.js:42-48); `out` is the f32 staging row of this block
__device__ __forceinline__ void mdct_post(const Cplx z, int i, int n, const double *__restrict__ tab,
                                          float *out, bool reverse) {
  const int half = n >> 1;
  const double c = __ldg(&tab[2 * i]), s = __ldg(&tab[2 * i + 1]);
  const float o0 = (float)(-z.re * c - z.im * s);
  const float o1 = (float)(-z.re * s + z.im * c);
  int i0 = 2 * i, i1 = half - 1 - 2 * i;
  if (reverse) { i0 = half - 1 - i0; i1 = half - 1 - i1; }
  out[i0] = o0;
  out[i1] = o1;
}

// One band of one sound unit, by one warp.  arr holds the windowed transform input (see
// mdct_kernel).
template <typename R>
__device__ __forceinline__ void mdct_band(int band, bool is_long, const double *arr, float *out,
                                          const DevTables *__restrict__ T, int lane) {
  R rnd;
  const double2 *tw = T->fft_tw;
  const int size = band == 2 ? 256 : 128;
  const bool rev = band > 0;
  if (is_long) {
    const int ws = band == 2 ? 112 : 48;  // constants.js:115-119
    const int span = size + 32;
    auto in = [&](int k) -> double {
      const unsigned a = (unsigned)(k - ws);
      return a < (unsigned)span ? arr[a] : 0.0;
    };
    const int r5 = brev_bits(lane, 5);
    if (band < 2) {
      const double *tab = T->mdct_fwd256;
      Cplx a = mdct_pre(r5, 256, in, tab, rnd);
      Cplx b = mdct_pre(32 + r5, 256, in, tab, rnd);
      warp_fft_regs<5>(a, b, tw, lane, rnd);
      mdct_post(a, lane, 256, tab, out, rev);
      mdct_post(b, lane + 32, 256, tab, out, rev);
    } else {
      const double *tab = T->mdct_fwd512;
      Cplx a0 = mdct_pre(2 * r5, 512, in, tab, rnd);
      Cplx b0 = mdct_pre(64 + 2 * r5, 512, in, tab, rnd);
      Cplx a1 = mdct_pre(2 * r5 + 1, 512, in, tab, rnd);
      Cplx b1 = mdct_pre(64 + 2 * r5 + 1, 512, in, tab, rnd);
      warp_fft128_regs(a0, b0, a1, b1, tw, lane, rnd);
      mdct_post(a0, lane, 512, tab, out, rev);
      mdct_post(b0, lane + 32, 512, tab, out, rev);
      mdct_post(a1, lane + 64, 512, tab, out, rev);
      mdct_post(b1, lane + 96, 512, tab, out, rev);
    }
  } else {
    const double *tab = T->mdct_fwd64;
    const int g = lane & 7, r3 = brev_bits(g, 3);
    for (int b0 = 0; b0 < (size >> 5); b0 += 4) {
      const int blk = b0 + (lane >> 3);
      const double *ab = arr + 64 * blk;
      auto in = [&](int k) -> double { return ab[k]; };
      Cplx a = mdct_pre(r3, 64, in, tab, rnd);
      Cplx b = mdct_pre(8 + r3, 64, in, tab, rnd);
      warp_fft_regs<3>(a, b, tw, lane, rnd);
      mdct_post(a, g, 64, tab, out + 32 * blk, rev);
      mdct_post(b, g + 8, 64, tab, out + 32 * blk, rev);
    }
  }
}

// Conversion-based rounding for inputs beyond kFastRoundInputLimit; out of line to keep the
// hot path small.
__device__ __noinline__ void mdct_band_exact(int band, bool is_long, const double *arr, float *out,
                                             const DevTables *__restrict__ T, int lane) {
  mdct_band<ExactRound>(band, is_long, arr, out, T, lane);
}

// ---- long blocks (c1_fft.cuh, in-thread passes) --------------------------------------------
// A warp task is a pair of consecutive sound units and a ROLE: role 0 transforms the low and
// mid bands of both units (4 x MDCT256), role 1 their high bands (2 x MDCT512).
//
// mdct.js:76-105 for FFT input q (natural order).  The transform buffer is zero outside
// [windowStart, windowStart + size + 32) (encoder.js:240-247), so of the four taps of a
// pre-twiddle two are always inside; the other two are inside only for the first 8 values of
// the first loop (q < 8) and the last 8 of the second (q >= N/4 - 8).  Position 8t + j of the
// bit-reversed array holds q = brev3(j) * kLanes + brev(t): first loop for even j, second for
// odd j, edge values only for j == 0 / j == 7 (every lane of FFT64; even t / odd t of FFT128).
// Adding the +0 of an outside tap is kept where it can turn a -0 into +0 (x + 0), dropped
// where it cannot (x - 0).
//
// Shared-memory layout (bank conflicts, profiles/r02_mdct_shared_conflicts.txt): the pre-twiddle reads every
// other element of the buffer (the first tap only odd elements, the second only even ones), so the buffer is kept
// de-interleaved: element k of [overlap 32 | samples kSize] (index 0 == windowStart) lives in plane k & 1 at index
// k >> 1, and a warp's loads are unit-stride.  The pre/post table is an array of (cos, sin) pairs whose pair s sits
// at slot s + (s >> 3): the bit-reversed lanes of a quarter-warp then touch 8 different 16-byte bank groups.
//   pr = &O[(n34 - 2 - ws) / 2 - rev_t], pm = &E[(n4 - ws) / 2 + rev_t], ptab = &tab2[pad(rev_t)]

template <int kRole, int kJ, typename R>
__device__ __forceinline__ Cplx mdct_pre_long(const double *pr, const double *pm, const double2 *ptab, bool edge,
                                              R &rnd) {
  using G = LongGeom<kRole>;
  constexpr int q = G::q_step(kJ);  // FFT input q + rev_t, i.e. i = 2 (q + rev_t)
  constexpr int n8 = G::kN / 8;     // n4 elements = n8 plane entries
  double r = pr[-q], m = pm[q];
  if (kJ == 0) {         // r = in(n34-1-i) + in(n34+i), m = in(n4+i) - in(n4-1-i): taps n34+i, n4-1-i
    const double r1 = edge ? pm[q + 2 * n8] : 0.0, m1 = edge ? pr[-q - 2 * n8] : 0.0;
    r = r + r1;
    m = m - m1;
  } else if (kJ == 7) {  // r = in(n34-1-i) - in(i-n4), m = in(n4+i) + in(5*n4-1-i): taps i-n4, 5*n4-1-i
    const double r1 = edge ? pm[q - 2 * n8] : 0.0, m1 = edge ? pr[-q + 2 * n8] : 0.0;
    r = r - r1;
    m = m + m1;
  } else if ((kJ & 1) == 0) {
    r = r + 0.0;
  } else {
    m = m + 0.0;
  }
  const double2 cs = ptab[pad8(q)];
  Cplx z;
  z.re = rnd.r0(r * cs.x + m * cs.y);
  z.im = rnd.r1(m * cs.x - r * cs.y);
  return z;
}

template <int kRole, int kJ, typename R>
__device__ __forceinline__ void mdct_pre_all(Cplx (&v)[8], const double *pr, const double *pm, const double2 *ptab,
                                             int t, R &rnd) {
  if constexpr (kJ < 8) {
    // FFT128: q < 8 needs brev4(t) < 8 (t even), q >= 120 needs t odd
    const bool edge = kRole == 0 || ((t & 1) == (kJ == 7 ? 1 : 0));
    v[kJ] = mdct_pre_long<kRole, kJ>(pr, pm, ptab, edge, rnd);
    mdct_pre_all<kRole, kJ + 1>(v, pr, pm, ptab, t, rnd);
  }
}

// Shared-memory geometry of a role's warp task.
template <int kRole>
struct MdctLayout {
  static constexpr int kSize = LongGeom<kRole>::kSize;
  // doubles per plane: (kSize + 32) / 2 used, + 4 so that the buffers of the two transforms a half-warp of role 0
  // reads together start 64 bytes apart modulo 128
  static constexpr int kPlane = (kSize + 32) / 2 + 4;      // 84 / 148
  static constexpr int kBuf = 2 * kPlane;                  // doubles per transform: 168 / 296
  // coefficient rows, de-interleaved the same way (even positions | odd positions); the row stride spreads the
  // transforms of a warp over the banks (role 0: 8 floats apart modulo 32, role 1: 16)
  static constexpr int kRow = kSize + (kRole == 0 ? 8 : 16);  // floats: 136 / 272
};

// The transforms of one warp task.  arr: kPerWarp de-interleaved buffers (MdctLayout; each doubles as its
// transform's transpose buffer once the pre-twiddle has read it); out: kPerWarp de-interleaved coefficient rows.
// Transforms whose bit in long_mask is clear (short mode) run the same instructions on whatever their buffer
// holds and do not store.
template <int kRole, typename R>
__device__ __forceinline__ void mdct_long_task(unsigned long_mask, double *arr, float *out, const double2 *tab2,
                                               const double2 *tw, int lane) {
  using G = LongGeom<kRole>;
  using L = MdctLayout<kRole>;
  constexpr int kWs = kRole == 0 ? 48 : 112;  // constants.js:115-119
  constexpr int kN = G::kN;
  R rnd;
  const G g(lane);
  double *a = arr + g.x * L::kBuf;
  Cplx v[8];
  mdct_pre_all<kRole, 0>(v, a + L::kPlane + (3 * kN / 4 - 2 - kWs) / 2 - g.rev_t, a + (kN / 4 - kWs) / 2 + g.rev_t,
                         tab2 + pad8(g.rev_t), g.t, rnd);
  fft_long_inthread<kRole>(v, g, reinterpret_cast<double2 *>(a), tw, rnd);
  if ((long_mask >> g.x) & 1) {
    // mdct.js:111-119: FFT output i gives coefficients 2i and size-1-2i; spectrum reversed for the mid and high
    // bands (utils.js:42-48), which swaps the two.  Even positions live in the first half of the row, odd ones in
    // the second: position 2i at E[i], position size-1-2i at O[size/2-1-i].
    const bool reverse = kRole == 1 || (g.x & 1);
    const int ib = g.out_base();
    float *row = out + g.x * L::kRow;
    const double2 *pt = tab2 + pad8(ib);
    float *pe = row + ib, *po = row + (G::kSize - 1) - ib;
#pragma unroll
    for (int k = 0; k < 8; k++) {
      const int st = G::out_step(k);
      const double2 cs = pt[pad8(st)];
      const float c0 = (float)(-v[k].re * cs.x - v[k].im * cs.y);
      const float c1 = (float)(-v[k].re * cs.y + v[k].im * cs.x);
      pe[st] = reverse ? c1 : c0;
      po[-st] = reverse ? c0 : c1;
    }
  }
}
static_assert(LongGeom<0>::kSlots * 2 <= MdctLayout<0>::kBuf && LongGeom<1>::kSlots * 2 <= MdctLayout<1>::kBuf,
              "transpose buffers alias the input buffers");

template <int kRole>
__device__ __noinline__ void mdct_long_task_exact(unsigned long_mask, double *arr, float *out, const double2 *tab2,
                                                  const double2 *tw, int lane) {
  mdct_long_task<kRole, ExactRound>(long_mask, arr, out, tab2, tw, lane);
}

#ifndef C1_MDCT_WARPS
#define C1_MDCT_WARPS 8
#define C1_MDCT_CTAS 3
#endif
constexpr int kMdctWarps = C1_MDCT_WARPS, kMdctCtasPerSm = C1_MDCT_CTAS;
struct MdctWarpSmem {
  double guard[16]; // the two edge taps of mdct_pre_long are loaded (and discarded) up to 8 doubles in front of a
                    // buffer on lanes where they fall into the zero padding
  double arr[672];  // role 0: 4 x 168, role 1: 2 x 296 (MdctLayout); short blocks: 512 + a 256-float staging row
  float out[544];   // role 0: 4 x 136, role 1: 2 x 272, de-interleaved rows
};
static_assert(4 * MdctLayout<0>::kBuf <= 672 && 2 * MdctLayout<1>::kBuf <= 672 && 4 * MdctLayout<0>::kRow <= 544 &&
              2 * MdctLayout<1>::kRow <= 544, "MdctWarpSmem");
// twiddle tables of the role, staged once per CTA: pre/post table (N/4 (cos, sin) pairs, pair s at slot pad8(s)) and
// the FFT recurrence twiddles of stages 3.. (indices below 128)
struct MdctTables {
  double2 tab2[144];
  double2 tw[128];
};
constexpr size_t kMdctSmemBytes = sizeof(MdctWarpSmem) * kMdctWarps + sizeof(MdctTables);

// Short blocks of one band (encoder.js:279-304): block b transforms
// [WIN * previous block (32) | block * reversed WIN (32)].  c: the band's samples of this frame
// (c - 512: the previous frame's).  The coefficients are produced in natural order in a staging row behind the
// transform inputs and then scattered into the de-interleaved row `row` (even positions | odd positions).
__device__ __noinline__ void mdct_short_band(int band, const float *__restrict__ c, bool has_prev, bool fast,
                                             double *arr, float *row, const DevTables *__restrict__ T, int lane,
                                             double w_fwd, double w_rev) {
  ExactRound xr;
  const int size = band == 2 ? 256 : 128;
  float *stage = reinterpret_cast<float *>(arr + 512);
  for (int b = 0; b < (size >> 5); b++) {
    float src_prev = 0.0f;
    bool have_prev = true;
    if (b == 0) { have_prev = has_prev; if (has_prev) src_prev = c[-512 + size - 32 + lane]; }
    else src_prev = c[32 * (b - 1) + lane];
    arr[64 * b + lane] = have_prev ? xr(w_fwd * (double)src_prev) : 0.0;
    arr[64 * b + 32 + lane] = xr((double)c[32 * b + lane] * w_rev);
  }
  __syncwarp();
  if (fast) mdct_band<FastRound>(band, false, arr, stage, T, lane);
  else mdct_band<ExactRound>(band, false, arr, stage, T, lane);
  __syncwarp();
  for (int p = lane; p < size; p += 32) row[(p & 1) * (size >> 1) + (p >> 1)] = stage[p];
  __syncwarp();
}

// findScaleFactor (bitallocation.js:290-299) from the largest |coefficient| of a BFU.  The
// log2 / ceil of the reference equals 3*(E+21) + #{thresholds of the binade below max}
// (DevTables::sf_thr, exact: SURVEY.md section 0.3).  NaN never wins the max, as in the reference.
__device__ __forceinline__ int scale_factor_index(float mx, const DevTables *__restrict__ T) {
  int sfi = 0;
  if (mx > 0.0f) {
    const uint32_t bits = __float_as_uint(mx);
    const int e3 = 3 * ((int)(bits >> 23) - 127 + 21);
    if (e3 >= 63) sfi = 63;
    else if (e3 >= 0) {
      sfi = e3 + (mx > T->sf_thr[e3]) + (mx > T->sf_thr[e3 + 1]) + (mx > T->sf_thr[e3 + 2]);
      if (sfi > 63) sfi = 63;
    }
  }
  return sfi;
}

// One warp task: sound units su0 = 2 * pair and su0 + 1 (consecutive frames of a row, unless
// the row ends in between).
template <int kRole>
__device__ __forceinline__ void mdct_warp_task(int pair, const float *__restrict__ bands,
                                               const uint8_t *__restrict__ modes, int frames, int n_su,
                                               const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                                               float *__restrict__ coefs, uint8_t *__restrict__ sfi_out,
                                               MdctWarpSmem &S, const double2 *s_tab2, const double2 *s_tw, int lane,
                                               double w_fwd, double w_rev) {
  using G = LongGeom<kRole>;
  using L = MdctLayout<kRole>;
  constexpr int kSize = G::kSize, kPer = G::kPerWarp / 2;  // transforms per unit
  constexpr int kOff = kRole == 0 ? 0 : 256;              // first band sample of the role
  double *arr = S.arr;
  float *out = S.out;
  const int su0 = 2 * pair;
  const bool have1 = su0 + 1 < n_su;
  const int frame0 = su0 % frames;
  const bool cont1 = have1 && frame0 + 1 < frames;  // unit 1 continues unit 0's row
  const float *cur0 = bands + onchip_row((size_t)su0) * 512 + kOff;
  ExactRound xr;
  // block modes of this role's bands: transform x = unit * kPer + band
  int mode[G::kPerWarp];
  unsigned long_mask = 0;
#pragma unroll
  for (int x = 0; x < G::kPerWarp; x++) {
    const int unit = x / kPer, band = kRole == 0 ? x % kPer : 2;
    mode[x] = 1;
    if (unit == 0 || have1) mode[x] = P->use_fixed ? P->fixed[band] : (int)modes[(size_t)(su0 + unit) * 4 + band];
    long_mask |= (mode[x] == 0 ? 1u : 0u) << x;
  }
  unsigned big = 0;  // largest |input| (binary32 bits) of this task's transforms
  {
    // 256 samples of the role per unit = 64 float4 = 2 per lane and unit
    float4 x[4];
#pragma unroll
    for (int k = 0; k < 4; k++) {
      const bool ok = k < 2 || have1;
      x[k] = ok ? __ldg(reinterpret_cast<const float4 *>(cur0 + 512 * (k >> 1)) + lane + 32 * (k & 1))
                : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float pv[kPer];  // tails of the frame before unit 0
#pragma unroll
    for (int b = 0; b < kPer; b++) pv[b] = frame0 > 0 ? __ldg(cur0 - 512 + b * kSize + kSize - 32 + lane) : 0.0f;
#pragma unroll
    for (int k = 0; k < 4; k++) {
      // role 0: k = transform (unit k>>1, band k&1), samples 4*lane ..; role 1: transform k>>1, samples 128*(k&1) + 4*lane ..
      // sample s is buffer element 32 + s: even samples at E[16 + s/2], odd ones at O[16 + s/2]
      const int at = kRole == 0 ? k * L::kBuf + 16 + 2 * lane : (k >> 1) * L::kBuf + 16 + 64 * (k & 1) + 2 * lane;
      *reinterpret_cast<double2 *>(arr + at) = make_double2((double)x[k].x, (double)x[k].z);
      *reinterpret_cast<double2 *>(arr + at + L::kPlane) = make_double2((double)x[k].y, (double)x[k].w);
      big = max(big, max(max(__float_as_uint(x[k].x) & 0x7FFFFFFFu, __float_as_uint(x[k].y) & 0x7FFFFFFFu),
                         max(__float_as_uint(x[k].z) & 0x7FFFFFFFu, __float_as_uint(x[k].w) & 0x7FFFFFFFu)));
    }
    __syncwarp();
    // tail windowing (encoder.js:309-316): the last 32 samples v of a frame become v * WIN[31-i] in
    // this frame's buffer and WIN[i] * v in the overlap slot of the next frame's.  Lane i owns buffer elements
    // i (overlap slot) and kSize + i (tail): plane i & 1, indices i >> 1 and kSize / 2 + (i >> 1).
    const int pl = (lane & 1) * L::kPlane, lo = lane >> 1, hi = kSize / 2 + (lane >> 1);
#pragma unroll
    for (int b = 0; b < kPer; b++) {
      double *a0 = arr + b * L::kBuf + pl, *a1 = arr + (kPer + b) * L::kBuf + pl;
      const double v0 = a0[hi], v1 = a1[hi];
      a0[lo] = frame0 > 0 ? xr(w_fwd * (double)pv[b]) : 0.0;
      a0[hi] = xr(v0 * w_rev);
      a1[lo] = cont1 ? xr(w_fwd * v0) : 0.0;
      a1[hi] = xr(v1 * w_rev);
      big = max(big, __float_as_uint(pv[b]) & 0x7FFFFFFFu);
    }
    __syncwarp();
  }
  // 0x71800000 is 2^100 as binary32: below it no transform value can reach the binary32 overflow
  // threshold, the one case FastRound cannot round
  const bool fast = __reduce_max_sync(0xffffffffu, big) < 0x71800000u;
  if (long_mask) {
    if (fast) mdct_long_task<kRole, FastRound>(long_mask, arr, out, s_tab2, s_tw, lane);
    else mdct_long_task_exact<kRole>(long_mask, arr, out, s_tab2, s_tw, lane);
    __syncwarp();
  }
  // ---- short blocks, one band at a time (rare: out of line)
  if (long_mask != (have1 ? (1u << G::kPerWarp) - 1u : (1u << kPer) - 1u)) {
    for (int x = 0; x < G::kPerWarp; x++) {
      const int unit = x / kPer, bsel = x % kPer;
      if (((long_mask >> x) & 1) || (unit == 1 && !have1)) continue;
      mdct_short_band(kRole == 0 ? bsel : 2, cur0 + 512 * unit + bsel * kSize, unit == 0 ? frame0 > 0 : cont1, fast,
                      arr, out + x * L::kRow, T, lane, w_fwd, w_rev);
    }
  }
  // scale-factor index of every BFU of the role (groupIntoBFUs, quantization.js:106-149 +
  // findScaleFactor): role 0 owns BFUs 0..35 of both units, role 1 BFUs 36..51
  {
    constexpr int kBfus = kRole == 0 ? 36 : 16, kFirst = kRole == 0 ? 0 : 36;
    const FormatTables &F = T->fmt;
    for (int item = lane; item < 2 * kBfus; item += 32) {
      const int unit = item / kBfus, b = kFirst + item % kBfus;
      if (unit == 1 && !have1) break;
      const int x = kRole == 0 ? unit * 2 + (b >= 20) : unit;
      const int start = (mode[x] == 0 ? F.start_long[b] : F.start_short[b]) - (kRole == 0 ? (b >= 20 ? 128 : 0) : 256);
      const int sz = F.specs[b];
      // positions start .. start + sz - 1 of the row: the even ones are E[(start + 1) >> 1 ..], the odd ones
      // O[start >> 1 ..]
      const float *ce = out + x * L::kRow + ((start + 1) >> 1);
      const float *co = out + x * L::kRow + kSize / 2 + (start >> 1);
      const int ne = ((start + sz + 1) >> 1) - ((start + 1) >> 1), no = ((start + sz) >> 1) - (start >> 1);
      // max |c| (NaN never wins, as with the reference's `if (a > max)`): one FMNMX per coefficient
      constexpr int kMaxHalf = kRole == 0 ? 5 : 10;
      float mx = 0.0f;
#pragma unroll
      for (int j = 0; j < kMaxHalf; j++) {
        if (j < ne) mx = fmaxf(mx, fabsf(ce[j]));
        if (j < no) mx = fmaxf(mx, fabsf(co[j]));
      }
      sfi_out[(size_t)(su0 + unit) * 64 + b] = (uint8_t)scale_factor_index(mx, T);
    }
  }
  // 256 coefficients of the role per unit, back to natural order: positions 4 lane .. 4 lane + 3 of a row are
  // E[2 lane], O[2 lane], E[2 lane + 1], O[2 lane + 1]
#pragma unroll
  for (int k = 0; k < 4; k++) {
    if (k < 2 || have1) {
      const float *row = kRole == 0 ? out + k * L::kRow + 2 * lane : out + (k >> 1) * L::kRow + 64 * (k & 1) + 2 * lane;
      const float2 e = *reinterpret_cast<const float2 *>(row), o = *reinterpret_cast<const float2 *>(row + kSize / 2);
      reinterpret_cast<float4 *>(coefs + (size_t)(su0 + (k >> 1)) * 512 + kOff)[lane + 32 * (k & 1)] = make_float4(e.x, o.x, e.y, o.y);
    }
  }
  __syncwarp();
}

// One kernel per role: the fully unrolled transforms of one role are ~30 KB of code; keeping a
// single role per launch keeps the hot loop inside the instruction cache.
template <int kRole>
__global__ void __launch_bounds__(kMdctWarps * 32, kMdctCtasPerSm)
mdct_kernel(const float *__restrict__ bands, const uint8_t *__restrict__ modes, int frames, int n_su,
            const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
            float *__restrict__ coefs, uint8_t *__restrict__ sfi_out) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  MdctWarpSmem &S = reinterpret_cast<MdctWarpSmem *>(smem_raw)[warp];
  MdctTables &ST = *reinterpret_cast<MdctTables *>(smem_raw + sizeof(MdctWarpSmem) * kMdctWarps);
  {
    const double *tab = kRole == 0 ? T->mdct_fwd256 : T->mdct_fwd512;
    for (int i = threadIdx.x; i < LongGeom<kRole>::kN / 4; i += blockDim.x) ST.tab2[pad8(i)] = make_double2(tab[2 * i], tab[2 * i + 1]);
    for (int i = threadIdx.x; i < 128; i += blockDim.x) ST.tw[i] = T->fft_tw[i];
  }
  __syncthreads();
  const double w_fwd = T->win[lane], w_rev = T->win[31 - lane];  // WINDOW_SHORT[i], [31 - i]
  const int n_pairs = (n_su + 1) >> 1;
  // persistent warps: the grid is sized to the machine and every warp walks the unit pairs
  for (int pair = blockIdx.x * kMdctWarps + warp; pair < n_pairs; pair += gridDim.x * kMdctWarps) {
    {  // the next pair's 2 x 1 KB of band samples (this role's half of each unit): 16 lines, one per lane
      const int next = pair + gridDim.x * kMdctWarps;
      const int su = 2 * next + (lane >> 3);
      if (lane < 16 && su < n_su) prefetch_l2(bands + onchip_row((size_t)su) * 512 + (kRole == 0 ? 0 : 256) + 32 * (lane & 7));
    }
    mdct_warp_task<kRole>(pair, bands, modes, frames, n_su, T, P, coefs, sfi_out, S, ST.tab2, ST.tw, lane, w_fwd, w_rev);
  }
}

// ------------------------------------------------------------------------------------
// K4a: scale factors + RDO bit allocation (bitallocation.js:74-341).
//
// The reference runs a greedy max-heap bit spend for each of the 8 candidate BFU counts and
// keeps the first candidate with the strictly smallest total distortion.  Ties between equal
// Float32Array priorities are resolved by the heap's structure, so the heap is emulated
// literally (same array layout, same sift-down, same strict comparisons).
//
// Candidate pruning (exact): candidate n leaves BFUs >= n uncoded, which alone costs
// tail(n) = sum_{i>=n} zeroBit[i].  The 52-BFU candidate has no tail.  If tail(n), deflated
// by the worst-case rounding of the reference's own summation, already exceeds the 52-BFU
// total, candidate n can neither win nor tie and is not run.  On broadband material fewer
// than 1 % of the smaller candidates survive; on silence all do (and cost nothing).
//
// One warp = 32 sound units, one per lane; warps are independent and persistent.  Pass 1: lane u
// runs the 52-BFU candidate of unit u.  Survivors of the warp's units are compacted into a list
// and run 32 at a time.
//
// Heap entries are one 32-bit word: key[29:15] | size[14:10] | wl[9:6] | bfu[5:0].  `key` is an
// order-isomorphic 15-bit image of the reference's f32 priority (DevEncParams::key0/key1):
// f32 exponent (8 bits) over the rank of the f32 mantissa among the 126 possible ones.  For
// wl >= 1 the priority halves exactly with every step, i.e. key -= 128.  Heaps are stored
// node-major / thread-minor, so a warp's 32 lanes always hit 32 distinct banks.
// ------------------------------------------------------------------------------------
constexpr int kAlWarps = 4;

struct AllocRec {  // per sound unit, global scratch between K4a and K4b
  uint8_t n_bfu, pad[3];
  uint8_t wl[52];
  uint8_t sfi[52];
  uint8_t pad2[4];
};
static_assert(sizeof(AllocRec) == 112 && offsetof(AllocRec, wl) == 4 && offsetof(AllocRec, sfi) == 56, "AllocRec layout");

struct AllocCand {  // per sound unit: results of surviving smaller candidates
  double total[7];
  uint8_t wl[7][52];
  uint8_t pad[4];
};
static_assert(sizeof(AllocCand) == 424, "AllocCand layout");

struct AlWarpSmem {  // one warp = 32 sound units, one per lane
  uint32_t heap[53][32];
  uint8_t wl[52][32];
  uint8_t sfi[32][52];
  uint8_t list[32 * 7];  // surviving (lane << 3 | candidate) pairs
};
struct AlSmem {
  AlWarpSmem w[kAlWarps];
  double bsf[64];          // DevEncParams::bsf
  float zero_bit[64 * 8];  // DevEncParams::zero_bit
  uint16_t key0[64], key1[64];
  uint8_t specs[52], size_class[52];
};

// Shared-memory byte addresses (32-bit) keep the sift loop free of 64-bit pointer math.
__device__ __forceinline__ uint32_t lds32(uint32_t addr) {
  uint32_t v;
  asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(addr));
  return v;
}
__device__ __forceinline__ void sts32(uint32_t addr, uint32_t v) {
  asm volatile("st.shared.u32 [%0], %1;" ::"r"(addr), "r"(v) : "memory");
}
constexpr uint32_t kNode = 32 * 4;  // byte stride between heap nodes

constexpr uint32_t kLow = 0x7FFFu;  // everything below the key: size[14:10] | wl[9:6] | bfu[5:0]
constexpr int kKeyShift = 15;

// bitallocation.js:314-341.  `hole` is the address of the node being filled, `end` the
// address of node `size`, `h0` the address of node 0; v is the entry being placed.
__device__ __forceinline__ void heap_sift(uint32_t h0, uint32_t hole, uint32_t end, uint32_t v) {
  const uint32_t vm = v | kLow;
  for (;;) {
    const uint32_t l = 2u * hole - h0 + kNode;  // node 2i+1
    if (l >= end) break;
    const uint32_t cl = lds32(l);
    const uint32_t cr = lds32(l + kNode);  // node `size` is kept zero
    const bool pl = cl > vm;
    const uint32_t m = pl ? (cl | kLow) : vm;
    const bool pr = cr > m;
    if (!(pl || pr)) break;
    sts32(hole, pr ? cr : cl);
    hole = pr ? l + kNode : l;
  }
  sts32(hole, v);
}

// The same from the root, for the greedy loop, where this walk is the critical path of every step:
// the four grandchildren are fetched while the two children are compared (a level costs the
// compare-select chain instead of a shared-memory round trip on top of it), and the entry that ends
// up at the root is returned instead of being read back.  Nodes past `end` are not heap entries
// (stale), so addresses are clamped to node `end`, which is kept zero.
__device__ __forceinline__ uint32_t heap_sift_root(uint32_t h0, uint32_t end, uint32_t v) {
  const uint32_t vm = v | kLow;
  uint32_t hole = h0, l = h0 + kNode, root = v;
  if (l < end) {
    uint32_t cl = lds32(l), cr = lds32(l + kNode);
    for (;;) {
      const uint32_t ll = 2u * l - h0 + kNode;  // children of l: ll, ll + 1; of l + 1: ll + 2, ll + 3 (in nodes)
      const uint32_t g0 = lds32(min(ll, end)), g1 = lds32(min(ll + kNode, end));
      const uint32_t g2 = lds32(min(ll + 2 * kNode, end)), g3 = lds32(min(ll + 3 * kNode, end));
      const bool pl = cl > vm;
      const uint32_t m = pl ? (cl | kLow) : vm;
      const bool pr = cr > m;
      if (!(pl || pr)) break;
      const uint32_t up = pr ? cr : cl;
      sts32(hole, up);
      if (hole == h0) root = up;
      hole = pr ? l + kNode : l;
      l = pr ? ll + 2 * kNode : ll;
      if (l >= end) break;
      cl = pr ? g2 : g0;
      cr = pr ? g3 : g1;
    }
  }
  sts32(hole, v);
  return root;
}

// Greedy spend for one candidate (bitallocation.js:203-281) followed by its total
// distortion (:157-190).  col = this lane's column in the heap / wl arrays.
__device__ __forceinline__ double run_candidate(AlWarpSmem &S, const AlSmem &C, const DevEncParams *__restrict__ P,
                                                const FormatTables &F, const uint8_t *sfi_row, int cand,
                                                int col) {
  const uint32_t h0 = (uint32_t)__cvta_generic_to_shared(&S.heap[0][col]);
  uint8_t *W = &S.wl[0][col];
  int remaining = kFrameBits - 40 - 10 * cand;  // bitallocation.js:97-100
  int count = 0;
  int min_sz = 32;  // smallest BFU in the heap: nothing can be bought for less
  for (int b = 0; b < cand; b++) {  // bitallocation.js:216-232
    W[b * 32] = 0;
    const uint32_t sfi = sfi_row[b];
    if (sfi) {
      const uint32_t sz = C.specs[b];
      sts32(h0 + count * kNode, ((uint32_t)C.key0[sfi] << kKeyShift) | (sz << 10) | (uint32_t)b);
      count++;
      min_sz = min(min_sz, (int)sz);
    }
  }
  sts32(h0 + count * kNode, 0);
  if (count) {
    uint32_t end = h0 + count * kNode;
    for (int i = (count >> 1) - 1; i >= 0; i--) heap_sift(h0, h0 + i * kNode, end, lds32(h0 + i * kNode));
    uint32_t e = lds32(h0);
    // bitallocation.js:244-278 (size > 0 is the loop's other exit).  Once fewer bits remain than
    // the smallest BFU costs, every further iteration of the reference's loop is a pop that
    // changes nothing: stop there.
    while (remaining >= min_sz) {
      const int b = e & 63;
      const int wl = (e >> 6) & 15;
      const int cost = (int)((e >> 10) & 31u) << (wl == 0);
      bool pop = cost > remaining;
      if (!pop) {
        remaining -= cost;
        if (wl == 0) e = ((uint32_t)C.key1[sfi_row[b]] << kKeyShift) | (e & (31u << 10)) | (1u << 6) | (uint32_t)b;
        else e += 64u - (128u << kKeyShift);  // wl + 1, priority halves exactly
        pop = wl == 14;                // reached MAX_WORD_LENGTH_INDEX
      }
      if (pop) {
        W[b * 32] = (uint8_t)((e >> 6) & 15);
        end -= kNode;
        if (end == h0) break;
        e = lds32(end);
        sts32(end, 0);
      }
      e = heap_sift_root(h0, end, e);
    }
    for (uint32_t a = h0; a < end; a += kNode) {  // entries still in the heap keep their wl
      const uint32_t x = lds32(a);
      W[(x & 63) * 32] = (uint8_t)((x >> 6) & 15);
    }
  }
  // bitallocation.js:157-190.  A BFU with scale-factor index 0 adds +0.0 or is skipped by the reference;
  // `total` starts at +0.0 and only ever grows, so adding +0.0 is the same: the loop is branch-free and
  // only the additions are serial.
  double total = 0.0;
#pragma unroll 4
  for (int i = 0; i < 52; i++) {
    const int sfi = sfi_row[i];
    const int bits = i < cand ? wl_bits(W[i * 32]) : 0;
    const double zb = (double)C.zero_bit[sfi * 8 + C.size_class[i]];
    const double inv = __hiloint2double((1023 - bits) << 20, 0);
    const double coded = C.bsf[sfi] * inv * (double)C.specs[i];
    total += sfi == 0 ? 0.0 : (bits == 0 ? zb : coded);
  }
  return total;
}

// One group of 32 consecutive output units, one unit per lane (the body of K4a).
__device__ __forceinline__ void alloc_group(AlWarpSmem &S, const AlSmem &C, const uint8_t *__restrict__ sfi_all, int frames,
                                            int halo, int n_out_frames, long long n_units, const FormatTables &F,
                                            const DevEncParams *__restrict__ P, AllocRec *recs, AllocCand *cands,
                                            long long unit0, int group_lanes, int lane) {
  __syncwarp();
  // ---- phase A: the scale-factor indices the MDCT kernels left per unit (64-byte records)
  for (int item = lane; item < group_lanes * 13; item += 32) {
    const int u = item / 13, w = item - u * 13;
    const long long unit = unit0 + u;
    uint32_t v = 0;
    if (unit < n_units) {
      // (unit counts stay far below 2^32: a unit owns 2 KB of coefficients; 32-bit division is a tenth of the 64-bit one)
      const unsigned st = (unsigned)unit / (unsigned)n_out_frames, fo = (unsigned)unit - st * (unsigned)n_out_frames;
      const size_t su = (size_t)st * frames + halo + fo;
      v = __ldg(reinterpret_cast<const uint32_t *>(sfi_all + su * 64) + w);
    }
    reinterpret_cast<uint32_t *>(&S.sfi[u][0])[w] = v;
  }
  __syncwarp();
  // ---- pass 1: the 52-BFU candidate of the lane's unit; its result is the provisional record
  const bool live = lane < group_lanes && unit0 + lane < n_units;
  double total52 = 0.0;
  uint32_t survive = 0;
  if (live) {
    total52 = run_candidate(S, C, P, F, S.sfi[lane], 52, lane);
    AllocRec *r = recs + (unit0 + lane);
    {  // the record, four BFUs per 32-bit store: n_bfu | wl[52] | sfi[52]
      uint32_t *rw = reinterpret_cast<uint32_t *>(r);
      rw[0] = 52u;
      const uint32_t *srow = reinterpret_cast<const uint32_t *>(&S.sfi[lane][0]);
#pragma unroll
      for (int q = 0; q < 13; q++) {
        rw[1 + q] = (uint32_t)S.wl[4 * q][lane] | ((uint32_t)S.wl[4 * q + 1][lane] << 8) |
                    ((uint32_t)S.wl[4 * q + 2][lane] << 16) | ((uint32_t)S.wl[4 * q + 3][lane] << 24);
        rw[14 + q] = srow[q];
      }
    }
    // Candidate pruning (exact): candidate n leaves BFUs >= n uncoded, which alone costs
    // tail(n) = sum_{i>=n} zeroBit[i]; if that, deflated by the worst-case rounding of the
    // reference's own 52-term summation (1 - 2^-40), already exceeds the 52-BFU total, the
    // candidate can neither win nor tie and is not run.  (zeroBit is +0.0 where the index is 0.)
    double tail = 0.0;
#pragma unroll
    for (int c = 6; c >= 0; c--) {
      const int bound = c == 0 ? 20 : 24 + 4 * c;  // BFU_AMOUNTS[c]
      const int upper = c == 6 ? 52 : 28 + 4 * c;  // BFU_AMOUNTS[c + 1]
      for (int i = upper - 1; i >= bound; i--) {
        const int sfi = S.sfi[lane][i];
        tail += sfi ? (double)C.zero_bit[sfi * 8 + C.size_class[i]] : 0.0;
      }
      if (!(tail * (1.0 - 9.094947017729282e-13) > total52)) survive |= 1u << c;
    }
  }
  // ---- compact the surviving (unit, candidate) pairs of the warp and run them 32 at a time
  const int n_mine = __popc(survive);
  int incl = n_mine;
#pragma unroll
  for (int d = 1; d < 32; d <<= 1) {
    const int t = __shfl_up_sync(0xffffffffu, incl, d);
    if (lane >= d) incl += t;
  }
  const int n_list = __shfl_sync(0xffffffffu, incl, 31);
  {
    int at = incl - n_mine;
    for (uint32_t m = survive; m; m &= m - 1) S.list[at++] = (uint8_t)((lane << 3) | (__ffs(m) - 1));
  }
  __syncwarp();
  for (int base = 0; base < n_list; base += 32) {  // uniform trip count
    const int k = base + lane;
    if (k < n_list) {
      const int u = S.list[k] >> 3, c = S.list[k] & 7;
      const int cand = c == 0 ? 20 : 24 + 4 * c;
      const double total = run_candidate(S, C, P, F, S.sfi[u], cand, lane);
      AllocCand *ac = cands + (unit0 + u);
      ac->total[c] = total;
      for (int b = 0; b < cand; b++) ac->wl[c][b] = S.wl[b][lane];
    }
    __syncwarp();
  }
  __threadfence_block();
  __syncwarp();
  // ---- first strict minimum over ascending candidates (bitallocation.js:91-130)
  if (live && survive) {
    const AllocCand *ac = cands + (unit0 + lane);
    double min_total = __longlong_as_double(0x7ff0000000000000ll);
    int best = -1;
    for (int c = 0; c < 7; c++) {
      if (!(survive >> c & 1)) continue;
      const double t = ac->total[c];
      if (t < min_total) { min_total = t; best = c; }
    }
    if (total52 < min_total) best = 7;
    AllocRec *r = recs + (unit0 + lane);
    if (best < 0) {  // bitallocation.js:132-139
      r->n_bfu = 20;
      for (int b = 0; b < 52; b++) { r->wl[b] = 0; r->sfi[b] = 0; }
    } else if (best < 7) {
      const int cand = best == 0 ? 20 : 24 + 4 * best;
      r->n_bfu = (uint8_t)cand;
      for (int b = 0; b < 52; b++) r->wl[b] = b < cand ? ac->wl[best][b] : 0;
    }
  } else if (live && !(total52 < __longlong_as_double(0x7ff0000000000000ll))) {
    AllocRec *r = recs + (unit0 + lane);  // no finite candidate at all: the reference's fallback
    r->n_bfu = 20;
    for (int b = 0; b < 52; b++) { r->wl[b] = 0; r->sfi[b] = 0; }
  }
}

__device__ __forceinline__ void alloc_stage_tables(AlSmem &C, const DevEncParams *__restrict__ P, const FormatTables &F, int tid) {
  if (tid < 64) { C.key0[tid] = P->key0[tid]; C.key1[tid] = P->key1[tid]; C.bsf[tid] = P->bsf[tid]; }
  if (tid < 52) { C.specs[tid] = F.specs[tid]; C.size_class[tid] = F.size_class[tid]; }
  for (int i = tid; i < 64 * 8; i += kAlWarps * 32) C.zero_bit[i] = P->zero_bit[i];
}

// Warps are independent (no CTA barrier after the tables are staged) and persistent: a warp
// walks groups of 32 consecutive output units, one unit per lane.
__global__ void __launch_bounds__(kAlWarps * 32)
alloc_kernel(const uint8_t *__restrict__ sfi_all, int frames, int halo, int n_out_frames, int n_streams,
             int group_lanes, const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
             AllocRec *__restrict__ recs, AllocCand *__restrict__ cands) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  AlSmem &C = *reinterpret_cast<AlSmem *>(smem_raw);
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const long long n_units = (long long)n_streams * n_out_frames;
  alloc_stage_tables(C, P, T->fmt, tid);
  __syncthreads();
  const long long n_groups = (n_units + group_lanes - 1) / group_lanes;
  for (long long group = (long long)blockIdx.x * kAlWarps + warp; group < n_groups; group += (long long)gridDim.x * kAlWarps)
    alloc_group(C.w[warp], C, sfi_all, frames, halo, n_out_frames, n_units, T->fmt, P, recs, cands, group * group_lanes,
                group_lanes, lane);
}

// ------------------------------------------------------------------------------------
// K4b: quantise (quantization.js:34-56) and pack the 212-byte unit
// (serialization.js:41-98, bitstream.js:15-39).  One warp per sound unit, persistent.
// Lane l owns coefficients l + 32k: all 16 are loaded before anything else happens; a
// position table gives (BFU, index inside the BFU) in one load, a 16-byte record per BFU the
// norm factor, bit offset, width and range in another.
// ------------------------------------------------------------------------------------
// OR the low `bits` bits of value into the big-endian bit image at bit position pos (MSB first,
// bitstream.js:15-39).  The code is aligned to the top of a word, split at the word boundary with
// two funnel shifts and merged with predicated shared-memory reductions (no branches).
__device__ __forceinline__ void red_or_shared(uint32_t addr, uint32_t v) {
  asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %1, 0;\n\t@p red.shared.or.b32 [%0], %1;\n\t}" ::"r"(addr), "r"(v) : "memory");
}
__device__ __forceinline__ void put_bits(uint32_t *words, int pos, uint32_t value, int bits) {
  const uint32_t t = value << (32 - bits);  // also drops the sign-extension bits of a negative code
  const int off = pos & 31;
  const uint32_t addr = (uint32_t)__cvta_generic_to_shared(words) + ((pos >> 5) << 2);
  red_or_shared(addr, t >> off);
  red_or_shared(addr + 4, __funnelshift_r(0u, t, off));  // off == 0: nothing spills over
}

constexpr int kQpWarps = 8;
struct QpBfu {          // per BFU of the unit being packed: one 16-byte load per coefficient
  double norm;          // quantRange / scaleFactor, 0 when the BFU carries no bits or sfi == 0
  uint32_t base_range;  // bit offset | quantRange << 16
  uint32_t bits;        // width of a code, 0 when nothing of the BFU goes into the image
};
static_assert(sizeof(QpBfu) == 16, "QpBfu is one 16-byte record");
struct QpWarpSmem {
  QpBfu bfu[52];
  uint32_t words[56];
};

// The 16 coefficients of a lane (coefficient lane + 32 k): quantise (quantization.js:34-56) and merge the
// codes into the unit's bit image.  kWrap: some BFU sits at the top scale factor, |coefficient| may exceed
// it and the reference's `| 0` (ToInt32) may wrap: that variant follows ToInt32 literally.
template <bool kWrap>
__device__ __forceinline__ void qp_coefs(QpWarpSmem &S, const float (&c)[16], const uint16_t *bj0, const uint16_t *bj1,
                                         const uint16_t *bj2) {
  uint32_t *words = S.words;
  double half = 0.0;  // +-0.5 with the sign of x: only the high word is rewritten per coefficient (the low word stays 0)
#pragma unroll
  for (int k = 0; k < 16; k++) {
    const uint32_t bj = (k < 4 ? bj0 : (k < 8 ? bj1 : bj2))[32 * k];
    const uint4 rw = *reinterpret_cast<const uint4 *>(&S.bfu[bj >> 5]);  // norm (lo, hi), base_range, bits
    const int bits = (int)rw.w;
    if (bits) {
      const int range = (int)(rw.z >> 16);
      // x = c * normFactor; y = (x + (x >= 0 ? 0.5 : -0.5)) | 0; clamp to +-range (norm > 0: uncoded BFUs were skipped)
      const double x = (double)c[k] * __hiloint2double((int)rw.y, (int)rw.x);
      int y;
      if (kWrap) {
        y = js_to_int32(x + (x >= 0.0 ? 0.5 : -0.5));
      } else {  // x + copysign(0.5, x): x >= 0 is true for -0 and so is a clear sign bit; x is never NaN here (finite c, norm)
        const int hs = (__double2hiint(x) & (int)0x80000000) | 0x3FE00000;
        asm("{\n\t.reg .b32 lo, hi;\n\tmov.b64 {lo, hi}, %0;\n\tmov.b64 %0, {lo, %1};\n\t}" : "+d"(half) : "r"(hs));
        y = __double2int_rz(x + half);
      }
      const int q = min(max(y, -range), range);
      put_bits(words, (int)(rw.z & 0xFFFFu) + (int)(bj & 31u) * bits, (uint32_t)q, bits);
    }
  }
}
__device__ __noinline__ void qp_coefs_wrap(QpWarpSmem &S, const float (&c)[16], const uint16_t *bj0, const uint16_t *bj1,
                                           const uint16_t *bj2) {
  qp_coefs<true>(S, c, bj0, bj1, bj2);
}

// One sound unit by one warp (the body of K4b).
__device__ __forceinline__ void qp_unit(QpWarpSmem &S, const uint16_t (*s_bj)[512], int sz0, int sz1,
                                        const float *__restrict__ coefs, const uint8_t *__restrict__ modes,
                                        const AllocRec *recs, int frames, int halo, int n_out_frames,
                                        const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                                        uint8_t *__restrict__ su_out, size_t su_frame_stride, size_t su_stream_stride,
                                        long long unit, int lane) {
  uint32_t *words = S.words;
  const int stream = (int)((unsigned)unit / (unsigned)n_out_frames);  // unit counts stay far below 2^32
  const int frame_out = (int)((unsigned)unit - (unsigned)stream * (unsigned)n_out_frames);
  const size_t su = (size_t)stream * frames + halo + frame_out;
  const float *src = coefs + su * 512;
  float c[16];
#pragma unroll
  for (int k = 0; k < 16; k++) c[k] = __ldg(src + lane + 32 * k);
  const AllocRec *r = recs + unit;
  const int n = r->n_bfu;
  int m0, m1, m2;
  if (P->use_fixed) { m0 = P->fixed[0]; m1 = P->fixed[1]; m2 = P->fixed[2]; }
  else { m0 = modes[su * 4]; m1 = modes[su * 4 + 1]; m2 = modes[su * 4 + 2]; }
  __syncwarp();  // the previous unit's words have been stored
  for (int i = lane; i < 56; i += 32) words[i] = 0;
  __syncwarp();
  // per-BFU records: widths, bit offsets (exclusive scan over bits * size), norm factors; the
  // word-length and scale-factor fields of the unit
  int run = 16 + 10 * n;
  bool wrap = false;  // a BFU at the top scale factor: |coefficient| may exceed it, ToInt32 may wrap
#pragma unroll
  for (int h = 0; h < 2; h++) {
    const int b = lane + 32 * h;
    const int sz = h == 0 ? sz0 : sz1;
    int wl = 0, sfi = 0;
    if (b < 52) { wl = r->wl[b]; sfi = r->sfi[b]; }
    const int bits = b < n ? wl_bits(wl) : 0;
    int incl = bits * sz;
#pragma unroll
    for (int d = 1; d < 32; d <<= 1) {
      const int t = __shfl_up_sync(0xffffffffu, incl, d);
      if (lane >= d) incl += t;
    }
    if (b < 52) {
      const int base = run + incl - bits * sz;
      const bool coded = bits > 0 && sfi > 0;
      S.bfu[b].norm = coded ? __ldg(&T->norm[wl][sfi]) : 0.0;
      // a BFU that carries bits but has scale-factor index 0 quantises to zeros (quantization.js:37-40):
      // nothing to merge into the image, so its width field is 0 and the coefficient loop skips it
      S.bfu[b].base_range = (uint32_t)base | ((uint32_t)((1 << wl) - 1) << 16);
      S.bfu[b].bits = (uint32_t)(coded ? bits : 0);
      wrap |= coded && sfi == 63;
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
  wrap = __any_sync(0xffffffffu, wrap);
  __syncwarp();
  const uint16_t *bj0 = s_bj[m0 != 0] + lane, *bj1 = s_bj[m1 != 0] + lane, *bj2 = s_bj[m2 != 0] + lane;
  if (wrap) qp_coefs_wrap(S, c, bj0, bj1, bj2);  // rare (a BFU at the top scale factor), out of line
  else qp_coefs<false>(S, c, bj0, bj1, bj2);
  __syncwarp();
  uint32_t *dst = reinterpret_cast<uint32_t *>(
      su_out + ((size_t)frame_out * su_frame_stride + (size_t)stream * su_stream_stride) * kSuBytes);
  for (int i = lane; i < kSuWords; i += 32) dst[i] = __byte_perm(words[i], 0, 0x0123);
}

__global__ void __launch_bounds__(kQpWarps * 32)
quant_pack_kernel(const float *__restrict__ coefs, const uint8_t *__restrict__ modes,
                  const AllocRec *__restrict__ recs, int frames, int halo, int n_out_frames, int n_streams,
                  const DevTables *__restrict__ T, const DevEncParams *__restrict__ P,
                  uint8_t *__restrict__ su_out, size_t su_frame_stride, size_t su_stream_stride) {
  __shared__ __align__(16) QpWarpSmem s_warp[kQpWarps];
  __shared__ uint16_t s_bj[2][512];  // long, short
  const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
  const FormatTables &F = T->fmt;
  for (int i = tid; i < 512; i += kQpWarps * 32) { s_bj[0][i] = F.bj_long[i]; s_bj[1][i] = F.bj_short[i]; }
  __syncthreads();
  const int sz0 = F.specs[lane], sz1 = lane < 20 ? F.specs[lane + 32] : 0;
  const long long n_units = (long long)n_streams * n_out_frames;
  for (long long unit = (long long)blockIdx.x * kQpWarps + warp; unit < n_units; unit += (long long)gridDim.x * kQpWarps) {
    qp_unit(s_warp[warp], s_bj, sz0, sz1, coefs, modes, recs, frames, halo, n_out_frames, T, P, su_out, su_frame_stride,
            su_stream_stride, unit, lane);
  }
}

// ------------------------------------------------------------------------------------
// Host-side launchers
// ------------------------------------------------------------------------------------
size_t alloc_rec_bytes() { return sizeof(AllocRec) + sizeof(AllocCand); }

const char *kernel_name(int id) {
  static const char *names[K_COUNT] = {"qmf_analysis", "band_mags", "transient_modes", "mdct",
                                       "alloc", "quant_pack", "unpack_dequant", "imdct", "bands_time", "synth"};
  return id >= 0 && id < K_COUNT ? names[id] : "?";
}

// Frames per run of the streaming QMF kernels.  A warp carries the filter state across a run, and a
// run that does not start its row first re-derives that state from the frame before it (about 0.6 of
// a frame's work).  With `warps` persistent warps the kernel lasts waves x (run + 0.6) frame times,
// waves = ceil(runs / warps): pick the run length that minimises it (e.g. 33 instead of 32 for the
// 1 h stereo workload: 8 full waves instead of 8.2, i.e. 9).  A launch that fits one wave whatever the run length
// (the stateful handles' few frames per call) gets runs of one frame: the call waits for the longest run.
int pick_run_len(int frames, int n_streams, int warps) {
  int best = 32;
  double best_cost = 1e300;
  for (int len = 1; len <= 96; len++) {
    const long long runs = (long long)((frames + len - 1) / len) * n_streams;
    const long long waves = (runs + warps - 1) / warps;
    const double cost = (double)waves * (std::min(len, frames) + 0.6);
    if (cost < best_cost - 1e-9) { best_cost = cost; best = len; }
  }
  return best;
}

// CTAs of a persistent kernel: as many as are resident at once (a larger grid would run a second,
// partly empty wave; the kernels stride over their work lists).  Cached per kernel.
int resident_ctas(const void *kernel, int threads, size_t dyn_smem) {
  struct Entry { const void *k; int dev; int ctas; };
  static std::mutex mu;  // contexts launch from different threads at the same time
  static Entry cache[64];
  static int n_cache = 0;
  int dev = 0;
  cudaGetDevice(&dev);
  std::lock_guard<std::mutex> lock(mu);
  for (int i = 0; i < n_cache; i++)
    if (cache[i].k == kernel && cache[i].dev == dev) return cache[i].ctas;
  int per_sm = 0;
  if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, kernel, threads, dyn_smem) != cudaSuccess || per_sm <= 0) {
    cudaGetLastError();
    per_sm = 1;
  }
  const int ctas = persistent_ctas(per_sm);
  if (n_cache < 64) cache[n_cache++] = {kernel, dev, ctas};
  return ctas;
}

// CTAs of a persistent kernel: every SM of the current device filled to `per_sm` resident CTAs.
// Role kernels side by side when both grids together leave the machine room (see ForkJoin); CARTA1_NO_FORK=1 keeps
// everything on one stream (an A/B switch for measurements).
bool fork_roles(const ForkJoin *fj, int grid, int resident) {
  static const bool off = getenv("CARTA1_NO_FORK") != nullptr;
  return fj && fj->aux && !off && 2 * grid <= resident;
}

int persistent_ctas(int per_sm) {
  static std::atomic<int> sms_of[64];  // per device; 0 = not queried yet (a racing double query stores the same value)
  int dev = 0;
  cudaGetDevice(&dev);
  int sms = sms_of[dev & 63].load(std::memory_order_relaxed);
  if (!sms) {
    if (cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sms <= 0) sms = 148;
    sms_of[dev & 63].store(sms, std::memory_order_relaxed);
  }
  return sms * per_sm;
}

cudaError_t launch_encode(const EncodeLaunch &L, cudaStream_t st, Prof *prof) {
  const int frames = L.frames_total;
  const int n_su = L.n_streams * frames;
  if (n_su == 0) return cudaSuccess;
  {
    cudaError_t e0 = cudaFuncSetAttribute(qmf_analysis_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQaSmemBytes);
    if (e0 == cudaSuccess)
      e0 = cudaFuncSetAttribute(qmf_analysis_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kQaSmemBytes);
    if (e0 != cudaSuccess) return e0;
    const int run_len = pick_run_len(frames, L.n_streams, persistent_ctas(kQaCtasPerSm) * kQaWarps);
    const int n_runs = ((frames + run_len - 1) / run_len) * L.n_streams;
    const int grid = std::min((n_runs + kQaWarps - 1) / kQaWarps, persistent_ctas(kQaCtasPerSm));
    prof->begin(K_QMF_ANALYSIS, st);
    if (L.pcm_fmt == 0)
      qmf_analysis_kernel<0><<<grid, kQaWarps * 32, kQaSmemBytes, st>>>(L.pcm, L.row_stride, L.n_ch_interleave,
                                                                         L.valid_samples, frames, L.n_streams, run_len, L.bands);
    else
      qmf_analysis_kernel<1><<<grid, kQaWarps * 32, kQaSmemBytes, st>>>(L.pcm, L.row_stride, L.n_ch_interleave,
                                                                         L.valid_samples, frames, L.n_streams, run_len, L.bands);
    prof->end(K_QMF_ANALYSIS, st);
  }
  if (!L.use_fixed) {
    cudaError_t e2 = cudaFuncSetAttribute(transient_spectrum_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmemBytes);
    if (e2 == cudaSuccess)
      e2 = cudaFuncSetAttribute(transient_spectrum_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kTsSmemBytes);
    if (e2 != cudaSuccess) return e2;
    prof->begin(K_BAND_MAGS, st);
    {
      const int n_ts_pairs = (n_su + 1) / 2;
      const int grid = std::min((n_ts_pairs + kTsWarps - 1) / kTsWarps, persistent_ctas(kTsCtasPerSm));
      const bool fork = fork_roles(L.fj, grid, persistent_ctas(kTsCtasPerSm));
      if (fork && (e2 = fork_begin(L.fj, st)) != cudaSuccess) return e2;
      transient_spectrum_kernel<0><<<grid, kTsWarps * 32, kTsSmemBytes, st>>>(L.bands, n_su, L.tables, L.mags,
                                                                              static_cast<SpectrumFeatures *>(L.feats));
      transient_spectrum_kernel<1><<<grid, kTsWarps * 32, kTsSmemBytes, fork ? L.fj->aux : st>>>(
          L.bands, n_su, L.tables, L.mags, static_cast<SpectrumFeatures *>(L.feats));
      if (fork && (e2 = fork_end(L.fj, st)) != cudaSuccess) return e2;
      prof->launches++;
    }
    prof->end(K_BAND_MAGS, st);
    prof->begin(K_TRANSIENT_MODES, st);
    const int n_tm_groups = (n_su + kTmUnits - 1) / kTmUnits;
    transient_modes_kernel<<<std::min((n_tm_groups + kTmWarps - 1) / kTmWarps,
                                      resident_ctas((const void *)transient_modes_kernel, kTmWarps * 32, 0)),
                             kTmWarps * 32, 0, st>>>(L.mags, static_cast<const SpectrumFeatures *>(L.feats),
                                                                  frames, L.halo_frames, n_su, L.tables, L.params, L.modes, L.scores,
                                                                  L.near_counts);
    prof->end(K_TRANSIENT_MODES, st);
  }
  prof->begin(K_MDCT, st);
  {
    cudaError_t e1 = cudaFuncSetAttribute(mdct_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMdctSmemBytes);
    if (e1 == cudaSuccess)
      e1 = cudaFuncSetAttribute(mdct_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)kMdctSmemBytes);
    if (e1 != cudaSuccess) return e1;
    const int n_pairs = (n_su + 1) / 2;
    const int grid = std::min((n_pairs + kMdctWarps - 1) / kMdctWarps, persistent_ctas(kMdctCtasPerSm));
    const bool fork = fork_roles(L.fj, grid, persistent_ctas(kMdctCtasPerSm));
    if (fork && (e1 = fork_begin(L.fj, st)) != cudaSuccess) return e1;
    mdct_kernel<0><<<grid, kMdctWarps * 32, kMdctSmemBytes, st>>>(L.bands, L.modes, frames, n_su, L.tables, L.params, L.coefs, L.sfi);
    mdct_kernel<1><<<grid, kMdctWarps * 32, kMdctSmemBytes, fork ? L.fj->aux : st>>>(L.bands, L.modes, frames, n_su, L.tables,
                                                                                   L.params, L.coefs, L.sfi);
    if (fork && (e1 = fork_end(L.fj, st)) != cudaSuccess) return e1;
    prof->launches++;
  }
  prof->end(K_MDCT, st);
  const long long n_units = (long long)L.n_streams * L.n_out_frames;
  if (n_units > 0 && L.su_out) {
    AllocRec *recs = static_cast<AllocRec *>(L.alloc_recs);
    AllocCand *cands = reinterpret_cast<AllocCand *>(recs + n_units);
    {
      cudaError_t e = cudaFuncSetAttribute(alloc_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sizeof(AlSmem));
      if (e != cudaSuccess) return e;
      // A lane walks its unit's heap serially and the lanes of a warp diverge, so a warp takes as long as
      // its units together need; small batches therefore use fewer lanes per warp, as few as still leave
      // no more groups than warps are resident (cfg1's 1,724 units: one lane each instead of 54 full warps).
      const int ctas = resident_ctas((const void *)alloc_kernel, kAlWarps * 32, sizeof(AlSmem));
      int group_lanes = 1;
      while (group_lanes < 32 && (n_units + group_lanes - 1) / group_lanes > (long long)ctas * kAlWarps) group_lanes *= 2;
      const long long n_groups = (n_units + group_lanes - 1) / group_lanes;
      prof->begin(K_ALLOC, st);
      alloc_kernel<<<(unsigned)std::min<long long>((n_groups + kAlWarps - 1) / kAlWarps, ctas), kAlWarps * 32, sizeof(AlSmem), st>>>(
          L.sfi, frames, L.halo_frames, L.n_out_frames, L.n_streams, group_lanes, L.tables, L.params, recs, cands);
      prof->end(K_ALLOC, st);
      prof->begin(K_QUANT_PACK, st);
      quant_pack_kernel<<<(unsigned)std::min<long long>((n_units + kQpWarps - 1) / kQpWarps,
                                                        resident_ctas((const void *)quant_pack_kernel, kQpWarps * 32, 0)),
                          kQpWarps * 32, 0, st>>>(
          L.coefs, L.modes, recs, frames, L.halo_frames, L.n_out_frames, L.n_streams, L.tables, L.params,
          L.su_out, L.su_frame_stride, L.su_stream_stride);
      prof->end(K_QUANT_PACK, st);
    }
  }
  return cudaGetLastError();
}

}  // namespace c1
