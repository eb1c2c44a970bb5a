// c1_abi.cu -- the extern "C" boundary declared in include/carta1_b200.h: context, tables,
// workspaces, chunked host<->device staging and the stateful stream handles.
#include <math.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <algorithm>
#include <atomic>
#include <chrono>
#include <mutex>
#include <string>
#include <thread>
#include <vector>

#include "../../include/carta1_b200.h"
#include "c1_common.cuh"
#include "c1_launch.h"

using namespace c1;

namespace {

// last error of the calls that have no context to keep it in (carta1_ctx_create, carta1_aea_parse_header):
// per thread, so that two threads creating contexts do not write one string
thread_local std::string g_create_error;

// The kernels read the QMF taps and the FFT twiddles of stages 0..2 from __constant__ memory, which is one
// copy per DEVICE, not per context.  The taps are literals of the format (constants.js:74-107) and never
// differ; the twiddles derive from carta1_tables::fft_w, which a host may inject.  Contexts on one device
// therefore have to agree on fft_w: the first context uploads, later ones are refused if theirs differ
// (CARTA1_ERR_ARG), and the upload only ever happens while no context of the device is alive, i.e. while
// none of its kernels can be running.
struct DeviceConstants {
  int live = 0;                 // contexts alive on the device
  bool loaded = false;
  double fft_w[8][2];
};
std::mutex g_const_mutex;
DeviceConstants g_const[64];

// ---------------------------------------------------------------- host tables
const int kSpecs[52] = {8, 8, 8, 8, 4, 4, 4, 4, 8, 8, 8, 8, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 6, 7, 7,
                        7, 7, 9, 9, 9, 9, 10, 10, 10, 10, 12, 12, 12, 12, 12, 12, 12, 12, 20, 20, 20,
                        20, 20, 20, 20, 20};
const int kStartLong[52] = {0, 8, 16, 24, 32, 36, 40, 44, 48, 56, 64, 72, 80, 86, 92, 98, 104, 110, 116,
                            122, 128, 134, 140, 146, 152, 159, 166, 173, 180, 189, 198, 207, 216, 226,
                            236, 246, 256, 268, 280, 292, 304, 316, 328, 340, 352, 372, 392, 412, 432,
                            452, 472, 492};
const int kStartShort[52] = {0, 32, 64, 96, 8, 40, 72, 104, 12, 44, 76, 108, 20, 52, 84, 116, 26, 58, 90,
                             122, 128, 160, 192, 224, 134, 166, 198, 230, 141, 173, 205, 237, 150, 182,
                             214, 246, 256, 288, 320, 352, 384, 416, 448, 480, 268, 300, 332, 364, 396,
                             428, 460, 492};
// QMF prototype, decimal literals of codec/core/constants.js:74-80 (stored as Float32Array)
const double kQmfLiterals[24] = {
    -0.00001461907, -0.00009205479, -0.000056157569, 0.00030117269, 0.0002422519,
    -0.00085293897, -0.0005205574,  0.0020340169,    0.00078333891, -0.0042153862,
    -0.00075614988, 0.0078402944,   -0.000061169922, -0.01344162,   0.0024626821,
    0.021736089,    -0.007801671,   -0.034090221,    0.01880949,    0.054326009,
    -0.043596379,   -0.099384367,   0.13207909,      0.46424159};

void fill_mdct_table(double *tab, int size, double scale) {  // mdct.js:21-37
  const double alpha = (2.0 * M_PI) / (8.0 * size);
  const double omega = (2.0 * M_PI) / size;
  const double root = sqrt(scale / size);
  for (int i = 0; i < size / 4; i++) {
    const double angle = omega * i + alpha;
    tab[2 * i] = root * cos(angle);
    tab[2 * i + 1] = root * sin(angle);
  }
}

// fdlibm log1p(10): the k != 0 branch of s_log1p.c evaluated for x = 10 (host, no FMA).
double fdlibm_log1p_10() {
  const double ln2_hi = 6.93147180369123816490e-01, ln2_lo = 1.90821492927058770002e-10,
               Lp1 = 6.666666666666735130e-01, Lp2 = 3.999999999940941908e-01,
               Lp3 = 2.857142874366239149e-01, Lp4 = 2.222219843214978396e-01,
               Lp5 = 1.818357216161805012e-01, Lp6 = 1.531383769920937332e-01,
               Lp7 = 1.479819860511658591e-01;
  // u = 11 = 1.375 * 2^3, hu = 0x60000 < 0x6a09e -> k = 3, f = 0.375, c = 0
  volatile double f = 0.375;
  const int k = 3;
  const double c = 0.0;
  volatile double hfsq = 0.5 * f * f;
  volatile double s = f / (2.0 + f);
  volatile double z = s * s;
  volatile double R = z * (Lp1 + z * (Lp2 + z * (Lp3 + z * (Lp4 + z * (Lp5 + z * (Lp6 + z * Lp7))))));
  return k * ln2_hi - ((hfsq - (s * (hfsq + R) + (k * ln2_lo + c))) - f);
}

void build_dev_tables(const carta1_tables &t, DevTables *d) {
  memset(d, 0, sizeof(*d));
  float window[48];
  for (int i = 0; i < 24; i++) {  // constants.js:83-90
    const float c = (float)kQmfLiterals[i];
    window[i] = (float)((double)c * 2.0);
    window[47 - i] = window[i];
  }
  for (int i = 0; i < 24; i++) {  // constants.js:93-107
    d->qmf_even[i] = (double)window[2 * i];
    d->qmf_odd[i] = (double)window[2 * i + 1];
  }
  memcpy(d->win, t.window_short, sizeof d->win);
  memcpy(d->sf, t.scale_factors, sizeof d->sf);
  memcpy(d->mdct_fwd64, t.mdct_fwd64, sizeof d->mdct_fwd64);
  memcpy(d->mdct_fwd256, t.mdct_fwd256, sizeof d->mdct_fwd256);
  memcpy(d->mdct_fwd512, t.mdct_fwd512, sizeof d->mdct_fwd512);
  memcpy(d->mdct_inv64, t.mdct_inv64, sizeof d->mdct_inv64);
  memcpy(d->mdct_inv256, t.mdct_inv256, sizeof d->mdct_inv256);
  memcpy(d->mdct_inv512, t.mdct_inv512, sizeof d->mdct_inv512);
  // FFT twiddles by the recurrence of fft.js:42-65 (restarts for every group, so it is a
  // pure per-stride table).  volatile keeps the host compiler from contracting.
  for (int lv = 0; lv < 8; lv++) {
    const int half = 1 << lv;
    const double wr = t.fft_w[lv][0], wi = t.fft_w[lv][1];
    volatile double tr = 1.0, ti = 0.0;
    for (int k = 0; k < half; k++) {
      d->fft_tw[half - 1 + k] = make_double2(tr, ti);
      volatile double a = tr * wr, b = ti * wi, c = tr * wi, e = ti * wr;
      const double nr = a - b;
      ti = c + e;
      tr = nr;
    }
  }
  // Exact findScaleFactor thresholds: thr[k] = largest f32 <= 2^(k/3 - 21), derived with
  // integer arithmetic (cube roots of 2 and 4 to 24 bits).
  uint32_t root[3] = {1u << 23, 0, 0};
  for (int r = 1; r <= 2; r++) {
    uint32_t lo = 1u << 23, hi = (1u << 24) - 1;
    const unsigned __int128 target = (unsigned __int128)(r == 1 ? 2 : 4) << 69;
    while (lo < hi) {
      const uint32_t mid = lo + (hi - lo + 1) / 2;
      const unsigned __int128 cube = (unsigned __int128)mid * mid * mid;
      if (cube <= target) lo = mid; else hi = mid - 1;
    }
    root[r] = lo;
  }
  for (int k = 0; k < 63; k++) d->sf_thr[k] = (float)ldexp((double)root[k % 3], k / 3 - 21 - 23);
  d->sf_thr[63] = INFINITY;
  d->log1p10 = fdlibm_log1p_10();
  for (int wl = 1; wl < 16; wl++) {
    volatile double range = (double)((1 << wl) - 1);
    d->rcp_range[wl] = 1.0 / range;
    if (wl >= 1 && wl <= kDeqMaxWl) {
      const int bits = wl + 1;
      for (int sfi = 0; sfi < 64; sfi++)
        for (int code = 0; code < (1 << bits); code++) {
          const int q = code >= (1 << (bits - 1)) ? code - (1 << bits) : code;  // bitstream.js:78-82
          const volatile double prod = (double)q * t.scale_factors[sfi];      // quantization.js:75: two roundings
          d->deq_tab[deq_off(wl) + (sfi << bits) + code] = sfi ? (float)(prod / range) : 0.0f;
        }
    }
  }
  for (int wl = 1; wl < 16; wl++)
    for (int i = 0; i < 64; i++) {
      volatile double range = (double)((1 << wl) - 1);
      d->norm[wl][i] = range / t.scale_factors[i];
    }
  for (int b = 0; b < 52; b++) {
    d->fmt.specs[b] = (uint8_t)kSpecs[b];
    d->fmt.start_long[b] = (uint16_t)kStartLong[b];
    d->fmt.start_short[b] = (uint16_t)kStartShort[b];
    for (int j = 0; j < kSpecs[b]; j++) {
      d->fmt.bfu_of_long[kStartLong[b] + j] = (uint8_t)b;
      d->fmt.bfu_of_short[kStartShort[b] + j] = (uint8_t)b;
      d->fmt.bj_long[kStartLong[b] + j] = (uint16_t)((b << 5) | j);
      d->fmt.bj_short[kStartShort[b] + j] = (uint16_t)((b << 5) | j);
    }
    static const int kSizes[8] = {4, 6, 7, 8, 9, 10, 12, 20};
    for (int c = 0; c < 8; c++)
      if (kSizes[c] == kSpecs[b]) d->fmt.size_class[b] = (uint8_t)c;
  }
}

// DISTORTION_DELTA_FACTORS / WORD_LENGTH_DELTA_BITS (constants.js:162-179)
double ddf(int i) { return i == 0 ? 2.0 - 0.25 : ldexp(1.0, -(i + 1)) - ldexp(1.0, -(i + 2)); }
int delta_bits(int i) { return i == 0 ? 2 : 1; }

// Returns false if the priorities cannot be represented by the 15-bit keys (only possible
// with pathological injected scale factors: zero, negative, non-finite or f32-subnormal).
bool build_enc_params(const carta1_tables &t, const carta1_enc_opts &o, DevEncParams *p) {
  static const int kSizes[8] = {4, 6, 7, 8, 9, 10, 12, 20};
  memset(p, 0, sizeof(*p));
  p->threshold = o.transient_threshold_low;
  for (int i = 0; i < 64; i++) {  // bitallocation.js:46-61
    if (o.biased_scale_factors) p->bsf[i] = o.biased_scale_factors[i];
    else p->bsf[i] = (o.allocation_bias == 1.0) ? t.scale_factors[i] : pow(t.scale_factors[i], o.allocation_bias);
  }
  p->use_fixed = o.use_fixed_block_modes ? 1 : 0;
  for (int i = 0; i < 3; i++) p->fixed[i] = o.fixed_block_modes[i];
  for (int s = 0; s < 64; s++)
    for (int c = 0; c < 8; c++) {  // bitallocation.js:86-87
      volatile double twice = p->bsf[s] * 2.0;
      p->zero_bit[s * 8 + c] = (float)(twice * (double)kSizes[c]);
    }
  // Heap priorities as the reference stores them: f32((bsf[sfi] * DDF[wl]) / deltaBits[wl])
  // (bitallocation.js:226-230,266-269).
  uint32_t pr[64][15];
  std::vector<uint32_t> mant;
  for (int s = 1; s < 64; s++) {
    for (int w = 0; w < 15; w++) {
      volatile double dd = p->bsf[s] * ddf(w);
      const float f = (float)(dd / (double)delta_bits(w));
      uint32_t bits;
      memcpy(&bits, &f, 4);
      const uint32_t e = (bits >> 23) & 0xFF;
      if ((bits >> 31) || e == 0 || e == 0xFF) return false;
      pr[s][w] = bits;
      if (w >= 2 && bits != pr[s][w - 1] - (1u << 23)) return false;  // halves exactly
    }
    mant.push_back(pr[s][0] & 0x7FFFFF);
    mant.push_back(pr[s][1] & 0x7FFFFF);
  }
  std::sort(mant.begin(), mant.end());
  mant.erase(std::unique(mant.begin(), mant.end()), mant.end());
  for (int s = 1; s < 64; s++)
    for (int w = 0; w < 2; w++) {
      const uint32_t m = (uint32_t)(std::lower_bound(mant.begin(), mant.end(), pr[s][w] & 0x7FFFFF) - mant.begin());
      const uint16_t key = (uint16_t)((((pr[s][w] >> 23) & 0xFF) << 7) | m);
      (w == 0 ? p->key0 : p->key1)[s] = key;
    }
  return true;
}

// Pageable transfers below this go straight to cudaMemcpyAsync (CARTA1_BOUNCE_MIN_BYTES overrides it: tests).
size_t bounce_min_bytes() {
  const char *v = getenv("CARTA1_BOUNCE_MIN_BYTES");
  return v && *v ? (size_t)strtoull(v, nullptr, 10) : (size_t)8 << 20;
}
constexpr int kSlots = 4;       // PCM staging slots of the pipelined host entry points
constexpr int kUnitSlots = 32;  // sound-unit input slots of the decoder (a tenth of the PCM bytes): deep enough that
                                // the uploads of a whole hour are queued before anything else competes for the copy engine

// Development aid (CARTA1_TRACE_PASSES=1): device timestamps of every pass's H2D, compute and D2H
// phases, printed to stderr when the call returns.
struct PassTrace {
  bool on = false;
  cudaEvent_t t0 = nullptr;
  std::vector<cudaEvent_t> ev;  // 6 per pass: h2d begin/end, compute begin/end, d2h begin/end
  const char *what = "";
  void start(const char *w, cudaStream_t st) {
    const char *v = getenv("CARTA1_TRACE_PASSES");
    on = v && *v && *v != '0';
    what = w;
    if (!on) return;
    cudaEventCreate(&t0);
    cudaEventRecord(t0, st);
  }
  void mark(cudaStream_t st) {
    if (!on) return;
    cudaEvent_t e;
    cudaEventCreate(&e);
    cudaEventRecord(e, st);
    ev.push_back(e);
  }
  void report() {
    if (!on) return;
    for (size_t i = 0; i + 5 < ev.size(); i += 6) {
      float t[6];
      for (int k = 0; k < 6; k++) cudaEventElapsedTime(&t[k], t0, ev[i + k]);
      fprintf(stderr, "[carta1 %s] pass %2zu  h2d %7.2f-%7.2f  compute %7.2f-%7.2f  d2h %7.2f-%7.2f ms\n", what, i / 6,
              t[0], t[1], t[2], t[3], t[4], t[5]);
    }
    for (cudaEvent_t e : ev) cudaEventDestroy(e);
    cudaEventDestroy(t0);
    ev.clear();
    on = false;
  }
};

// Bumped whenever a device buffer moves: captured launch sequences (GraphCache) hold raw pointers.
// Process-wide and atomic: contexts are used from different threads at the same time.
static std::atomic<unsigned long long> g_alloc_generation{0};

struct DevBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    g_alloc_generation++;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaMalloc(&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFree(p); p = nullptr; cap = 0; }
};

// Pinned host memory owned by the context: bounce buffers for callers whose arrays are pageable.
struct HostBuf {
  void *p = nullptr;
  size_t cap = 0;
  cudaError_t ensure(size_t bytes) {
    if (bytes <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    const size_t want = bytes + (bytes >> 3) + 256;
    cudaError_t e = cudaHostAlloc(&p, want, cudaHostAllocDefault);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() { if (p) cudaFreeHost(p); p = nullptr; cap = 0; }
};

// memcpy split over a few host threads (one thread moves ~10 GB/s; the PCIe link takes 55).
void parallel_memcpy(void *dst, const void *src, size_t bytes) {
  const unsigned hw = std::thread::hardware_concurrency();
  int nt = (int)std::min<size_t>(std::max(1u, std::min(8u, hw / 2)), bytes >> 21);  // >= 2 MiB per thread
  if (nt <= 1) { memcpy(dst, src, bytes); return; }
  const size_t piece = ((bytes / (size_t)nt) + 4095) & ~(size_t)4095;
  std::vector<std::thread> th;
  for (int i = 1; i < nt; i++) {
    const size_t a = std::min(bytes, piece * (size_t)i), b = std::min(bytes, a + piece);
    if (b > a) th.emplace_back([=] { memcpy((char *)dst + a, (const char *)src + a, b - a); });
  }
  memcpy(dst, src, std::min(bytes, piece));
  for (auto &t : th) t.join();
}

// The stateful handles are called with a few frames at a time: eight to eleven small kernels whose launch
// gaps are most of the call (131 us for one frame).  The kernel sequence of a call shape (frames per call) is
// captured into a CUDA graph the second time the shape is seen and replayed afterwards; the host<->device
// copies of the caller's buffers stay outside it.  A moved scratch buffer (g_alloc_generation) or a profiling
// session invalidates / bypasses the graph.  Any failure while capturing disables graphs for the handle.
struct GraphCache {
  int shape = -1;            // frames per call (decoder: with a previous unit) the graph was captured for
  int seen = 0;              // consecutive calls with this shape
  unsigned long long generation = 0;
  cudaGraphExec_t exec = nullptr;
  unsigned long long kernels = 0;  // launches one replay stands for
  bool disabled = false;
  void drop() {
    if (exec) cudaGraphExecDestroy(exec);
    exec = nullptr;
  }
  // 0: run eagerly, 1: replay, 2: capture this call
  int plan(int new_shape, bool profiling) {
    if (disabled || profiling || getenv("CARTA1_NO_GRAPHS")) return 0;
    if (new_shape != shape) { drop(); shape = new_shape; seen = 0; }
    seen++;
    if (exec && generation == g_alloc_generation) return 1;
    drop();
    return seen >= 2 ? 2 : 0;
  }
};

}  // namespace

struct carta1_ctx {
  // Every compute entry point holds this for the duration of the call: the context's stream, scratch buffers, graph
  // caches and error text are shared by the handles created from it, so calls on one context (from any number of
  // threads, on any of its handles) run one after the other; concurrency is one context per thread.
  std::recursive_mutex call_mu;
  int device = 0;
  cudaStream_t stream = nullptr;
  carta1_tables tables;
  DevTables *d_tables = nullptr;
  DevEncParams *d_params = nullptr;  // params of the most recent whole-buffer/device call
  std::string err;
  Prof prof;
  // cache of the parameters resident in d_params
  bool params_valid = false;
  carta1_enc_opts params_opts;
  double params_bsf[64];
  DevBuf bands, mags, feats, modes, coefs, sfi, inv, scores, dbg, recs;
  unsigned long long *d_near = nullptr;  // transient close calls (< 1e-9, < 1e-12), see carta1_ctx_near_threshold
  uint64_t near_decisions = 0;           // block-mode decisions taken (3 per emitted unit of an auto-mode call)
  // Host entry points: passes rotate through kSlots staging slots so that the H2D copy of pass i+1 and the D2H
  // copy of pass i-1 run while pass i computes (three streams, events between them).
  DevBuf stage_pcm[kSlots], stage_su[kUnitSlots];
  cudaStream_t h2d = nullptr, d2h = nullptr;
  cudaStream_t small = nullptr;  // highest priority: the sound-unit side of a host call (see small_copy)
  ForkJoin fj;                   // small launches run the two role kernels of a transform side by side (c1_launch.h)
  cudaEvent_t ev_hist = nullptr; // carta1_enc_frames: the history rows of the next call have been saved (on fj.aux)
  cudaEvent_t ev_in[kSlots] = {}, ev_comp[kSlots] = {}, ev_out[kSlots] = {};
  cudaEvent_t ev_uin[kUnitSlots] = {}, ev_ucomp[kUnitSlots] = {};
  // pageable caller buffers are staged through these (filled / drained by parallel_memcpy, one pass behind)
  HostBuf bounce_pcm[kSlots], bounce_su[kSlots];
  cudaEvent_t ev_bin[kSlots] = {};  // the H2D that read bounce slot i has finished
  size_t max_units_per_pass = 1u << 16;  // frames*channels per pass of the chunked host entry points
  bool holds_constants = false;          // counted in g_const[device].live
};

struct carta1_encoder {
  carta1_ctx *ctx;
  int n_streams;
  carta1_enc_opts opts;
  double bsf_copy[64];
  DevEncParams *d_params = nullptr;
  float *d_hist = nullptr;  // [n_streams][1024]: the last two frames of PCM per stream
  DevBuf work, su;
  GraphCache graph;
};

struct carta1_decoder {
  carta1_ctx *ctx;
  int n_streams;
  bool has_prev = false;
  float *d_rec = nullptr;  // [n_streams][512]: band record (IMDCT output) of the previous unit
  DevBuf work, pcm;
  GraphCache graph;
};

namespace {

struct CtxLock {
  std::recursive_mutex *m;
  explicit CtxLock(carta1_ctx *ctx) : m(ctx ? &ctx->call_mu : nullptr) { if (m) m->lock(); }
  ~CtxLock() { if (m) m->unlock(); }
  CtxLock(const CtxLock &) = delete;
  CtxLock &operator=(const CtxLock &) = delete;
};

int fail(carta1_ctx *ctx, int code, const std::string &msg) {
  if (ctx) ctx->err = msg; else g_create_error = msg;
  return code;
}
int cuda_fail(carta1_ctx *ctx, cudaError_t e, const char *what) {
  return fail(ctx, CARTA1_ERR_CUDA, std::string(what) + ": " + cudaGetErrorString(e));
}
#define CU(ctx, call)                                              \
  do {                                                             \
    cudaError_t e_ = (call);                                       \
    if (e_ != cudaSuccess) return cuda_fail((ctx), e_, #call);     \
  } while (0)

int upload_params(carta1_ctx *ctx, const carta1_enc_opts *opts, DevEncParams *d_params) {
  carta1_enc_opts o;
  if (opts) o = *opts; else carta1_default_enc_opts(&o);
  if (d_params == ctx->d_params) {  // skip the rebuild + upload when nothing changed
    const carta1_enc_opts &c = ctx->params_opts;
    bool same = ctx->params_valid && c.transient_threshold_low == o.transient_threshold_low &&
                c.allocation_bias == o.allocation_bias && c.use_fixed_block_modes == o.use_fixed_block_modes &&
                memcmp(c.fixed_block_modes, o.fixed_block_modes, sizeof c.fixed_block_modes) == 0 &&
                (c.biased_scale_factors != nullptr) == (o.biased_scale_factors != nullptr);
    if (same && o.biased_scale_factors)
      same = memcmp(ctx->params_bsf, o.biased_scale_factors, sizeof ctx->params_bsf) == 0;
    if (same) return CARTA1_OK;
    ctx->params_valid = false;  // until the new parameters are on the device
  }
  DevEncParams hp;
  if (!build_enc_params(ctx->tables, o, &hp))
    return fail(ctx, CARTA1_ERR_ARG, "carta1: biased scale factors must be positive, finite and normal in binary32");
  CU(ctx, cudaMemcpyAsync(d_params, &hp, sizeof(hp), cudaMemcpyHostToDevice, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));  // hp lives on this stack frame
  if (d_params == ctx->d_params) {
    ctx->params_opts = o;
    if (o.biased_scale_factors) memcpy(ctx->params_bsf, o.biased_scale_factors, sizeof ctx->params_bsf);
    ctx->params_valid = true;
  }
  return CARTA1_OK;
}

int ensure_encode_scratch(carta1_ctx *ctx, size_t units, bool auto_modes) {
  CU(ctx, ctx->bands.ensure(units * 512 * sizeof(float)));
  CU(ctx, ctx->coefs.ensure(units * 512 * sizeof(float)));
  CU(ctx, ctx->sfi.ensure(units * 64));
  CU(ctx, ctx->modes.ensure(units * 4));
  if (auto_modes) CU(ctx, ctx->mags.ensure(units * 256 * sizeof(float)));
  if (auto_modes) CU(ctx, ctx->feats.ensure(units * 9 * sizeof(double)));
  return CARTA1_OK;
}
int ensure_decode_scratch(carta1_ctx *ctx, size_t units) {
  CU(ctx, ctx->inv.ensure(units * 512 * sizeof(float)));
  CU(ctx, ctx->coefs.ensure(units * 512 * sizeof(float)));
  CU(ctx, ctx->modes.ensure(units * 4));
  return CARTA1_OK;
}

}  // namespace

// Runs `body` (kernel launches on ctx->stream only, returning a CARTA1 code) as planned by the cache.
template <typename Body>
static int run_graphed(carta1_ctx *ctx, GraphCache &g, int shape, Body body) {
  const int plan = g.plan(shape, ctx->prof.on);
  if (plan == 1) {
    CU(ctx, cudaGraphLaunch(g.exec, ctx->stream));
    ctx->prof.launches += g.kernels;
    return CARTA1_OK;
  }
  if (plan == 2) {
    const unsigned long long before = ctx->prof.launches;
    static const bool trace = getenv("CARTA1_TRACE_GRAPH") != nullptr;  // development aid: where the capturing call spends its time
    const auto t0 = std::chrono::steady_clock::now();
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      const int rc = body();
      cudaGraph_t graph = nullptr;
      const cudaError_t ee = cudaStreamEndCapture(ctx->stream, &graph);
      const auto t1 = std::chrono::steady_clock::now();
      cudaGraphExec_t exec = nullptr;
      if (rc == CARTA1_OK && ee == cudaSuccess && graph && cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        const auto t2 = std::chrono::steady_clock::now();
        cudaGraphDestroy(graph);
        g.exec = exec;
        g.generation = g_alloc_generation;
        g.kernels = ctx->prof.launches - before;
        CU(ctx, cudaGraphLaunch(g.exec, ctx->stream));
        if (trace) {
          cudaStreamSynchronize(ctx->stream);
          const auto t3 = std::chrono::steady_clock::now();
          auto ms = [](auto a, auto b) { return std::chrono::duration<double, std::milli>(b - a).count(); };
          fprintf(stderr, "[carta1] graph of shape %d: capture %.3f ms, instantiate %.3f ms, first launch + sync %.3f ms\n", shape,
                  ms(t0, t1), ms(t1, t2), ms(t2, t3));
        }
        return CARTA1_OK;
      }
      if (graph) cudaGraphDestroy(graph);
      cudaGetLastError();
      ctx->prof.launches = before;
    }
    g.disabled = true;  // nothing ran while capturing: fall through to the eager path
  }
  return body();
}

extern "C" {

int carta1_abi_version(void) { return 1; }

void carta1_default_tables(carta1_tables *t) {
  for (int i = 0; i < 32; i++) t->window_short[i] = sin(((i + 0.5) * M_PI) / 64);  // constants.js:60-66
  for (int i = 0; i < 64; i++) t->scale_factors[i] = pow(2.0, i / 3.0 - 21);       // constants.js:144-150
  fill_mdct_table(t->mdct_fwd64, 64, 0.5);                                         // mdct.js:215-221
  fill_mdct_table(t->mdct_fwd256, 256, 0.5);
  fill_mdct_table(t->mdct_fwd512, 512, 1.0);
  fill_mdct_table(t->mdct_inv64, 64, 64 * 8);
  fill_mdct_table(t->mdct_inv256, 256, 256 * 8);
  fill_mdct_table(t->mdct_inv512, 512, 512 * 4);
  for (int k = 0; k < 8; k++) {  // fft.js:36-39
    const int stride = 2 << k;
    const double angle = (-2 * M_PI) / stride;
    t->fft_w[k][0] = cos(angle);
    t->fft_w[k][1] = sin(angle);
  }
}

void carta1_default_enc_opts(carta1_enc_opts *o) {  // options.js:17-23
  memset(o, 0, sizeof(*o));
  o->transient_threshold_low = 1.0;
  o->allocation_bias = 1.0;
  o->use_fixed_block_modes = 0;
  o->biased_scale_factors = nullptr;
}

int carta1_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) return 0;
  return n;
}

int carta1_ctx_create(int device, const carta1_tables *tables, carta1_ctx **out) {
  if (!out) return fail(nullptr, CARTA1_ERR_ARG, "carta1_ctx_create: out is NULL");
  *out = nullptr;
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess || n == 0)
    return fail(nullptr, CARTA1_ERR_CUDA,
                std::string("carta1_b200 needs a CUDA device (sm_100a); none usable: ") +
                    (e != cudaSuccess ? cudaGetErrorString(e) : "device count is 0"));
  if (device < 0 || device >= n) return fail(nullptr, CARTA1_ERR_ARG, "carta1_ctx_create: bad device index");
  e = cudaSetDevice(device);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaSetDevice");
  carta1_ctx *ctx = new carta1_ctx();
  ctx->device = device;
  if (tables) ctx->tables = *tables; else carta1_default_tables(&ctx->tables);
  DevTables *ht = new DevTables();
  build_dev_tables(ctx->tables, ht);
  e = cudaStreamCreateWithFlags(&ctx->stream, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->h2d, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->d2h, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaStreamCreateWithFlags(&ctx->fj.aux, cudaStreamNonBlocking);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fj.fork, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->fj.join, cudaEventDisableTiming);
  if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_hist, cudaEventDisableTiming);
  if (e == cudaSuccess) {
    int lo = 0, hi = 0;
    cudaDeviceGetStreamPriorityRange(&lo, &hi);
    e = cudaStreamCreateWithPriority(&ctx->small, cudaStreamNonBlocking, hi);
  }
  for (int i = 0; i < kSlots && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&ctx->ev_in[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_comp[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_out[i], cudaEventDisableTiming);
  }
  for (int i = 0; i < kSlots && e == cudaSuccess; i++) e = cudaEventCreateWithFlags(&ctx->ev_bin[i], cudaEventDisableTiming);
  for (int i = 0; i < kUnitSlots && e == cudaSuccess; i++) {
    e = cudaEventCreateWithFlags(&ctx->ev_uin[i], cudaEventDisableTiming);
    if (e == cudaSuccess) e = cudaEventCreateWithFlags(&ctx->ev_ucomp[i], cudaEventDisableTiming);
  }
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_tables, sizeof(DevTables));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_params, sizeof(DevEncParams));
  if (e == cudaSuccess) e = cudaMalloc(&ctx->d_near, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemset(ctx->d_near, 0, 2 * sizeof(unsigned long long));
  if (e == cudaSuccess) e = cudaMemcpy(ctx->d_tables, ht, sizeof(DevTables), cudaMemcpyHostToDevice);
  bool tables_clash = false;
  if (e == cudaSuccess) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    DeviceConstants &dc = g_const[device & 63];
    const bool same = dc.loaded && memcmp(dc.fft_w, ctx->tables.fft_w, sizeof dc.fft_w) == 0;
    if (dc.live > 0 && !same) {
      tables_clash = true;
    } else {
      if (!same) {  // no context of this device is alive: nothing of ours is running on it
        dc.loaded = false;
        e = cudaDeviceSynchronize();
        if (e == cudaSuccess) e = upload_encode_constants(ht);
        if (e == cudaSuccess) e = upload_decode_constants(ht);
        if (e == cudaSuccess) {
          memcpy(dc.fft_w, ctx->tables.fft_w, sizeof dc.fft_w);
          dc.loaded = true;
        }
      }
      if (e == cudaSuccess) { dc.live++; ctx->holds_constants = true; }
    }
  }
  delete ht;
  if (tables_clash) {
    carta1_ctx_destroy(ctx);
    return fail(nullptr, CARTA1_ERR_ARG,
                "carta1_ctx_create: a live context on this device was created with different fft_w tables "
                "(the FFT twiddles live in per-device constant memory)");
  }
  if (e != cudaSuccess) {
    cuda_fail(nullptr, e, "carta1_ctx_create");
    carta1_ctx_destroy(ctx);
    return CARTA1_ERR_CUDA;
  }
  {
    // The first cudaGraphInstantiate of a process initialises the driver's graph machinery and takes 0.4 - 20 ms
    // (measured, CARTA1_TRACE_GRAPH); later ones take 0.08 ms.  Pay it here, not in the second frame call of the first
    // stateful handle.  Failures are ignored: graphs are an optimisation (GraphCache falls back to plain launches).
    cudaGraph_t graph = nullptr;
    cudaGraphExec_t exec = nullptr;
    if (cudaStreamBeginCapture(ctx->stream, cudaStreamCaptureModeThreadLocal) == cudaSuccess) {
      cudaMemsetAsync(ctx->d_near, 0, 2 * sizeof(unsigned long long), ctx->stream);
      if (cudaStreamEndCapture(ctx->stream, &graph) == cudaSuccess && graph &&
          cudaGraphInstantiate(&exec, graph, 0) == cudaSuccess) {
        cudaGraphLaunch(exec, ctx->stream);
        cudaStreamSynchronize(ctx->stream);
      }
    }
    if (exec) cudaGraphExecDestroy(exec);
    if (graph) cudaGraphDestroy(graph);
    cudaGetLastError();
  }
  *out = ctx;
  return CARTA1_OK;
}

void carta1_ctx_destroy(carta1_ctx *ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  ctx->bands.release(); ctx->mags.release(); ctx->feats.release(); ctx->modes.release(); ctx->coefs.release(); ctx->sfi.release();
  if (ctx->h2d) cudaStreamSynchronize(ctx->h2d);
  if (ctx->d2h) cudaStreamSynchronize(ctx->d2h);
  if (ctx->small) cudaStreamSynchronize(ctx->small);
  ctx->inv.release(); ctx->scores.release(); ctx->dbg.release(); ctx->recs.release();
  for (int i = 0; i < kUnitSlots; i++) {
    ctx->stage_su[i].release();
    if (ctx->ev_uin[i]) cudaEventDestroy(ctx->ev_uin[i]);
    if (ctx->ev_ucomp[i]) cudaEventDestroy(ctx->ev_ucomp[i]);
  }
  for (int i = 0; i < kSlots; i++) {
    ctx->bounce_pcm[i].release(); ctx->bounce_su[i].release();
    if (ctx->ev_bin[i]) cudaEventDestroy(ctx->ev_bin[i]);
    ctx->stage_pcm[i].release();
    if (ctx->ev_in[i]) cudaEventDestroy(ctx->ev_in[i]);
    if (ctx->ev_comp[i]) cudaEventDestroy(ctx->ev_comp[i]);
    if (ctx->ev_out[i]) cudaEventDestroy(ctx->ev_out[i]);
  }
  if (ctx->ev_hist) cudaEventDestroy(ctx->ev_hist);
  if (ctx->fj.fork) cudaEventDestroy(ctx->fj.fork);
  if (ctx->fj.join) cudaEventDestroy(ctx->fj.join);
  if (ctx->fj.aux) cudaStreamDestroy(ctx->fj.aux);
  if (ctx->h2d) cudaStreamDestroy(ctx->h2d);
  if (ctx->d2h) cudaStreamDestroy(ctx->d2h);
  if (ctx->small) cudaStreamDestroy(ctx->small);
  if (ctx->d_tables) cudaFree(ctx->d_tables);
  if (ctx->d_params) cudaFree(ctx->d_params);
  if (ctx->d_near) cudaFree(ctx->d_near);
  if (ctx->stream) cudaStreamDestroy(ctx->stream);
  if (ctx->holds_constants) {
    std::lock_guard<std::mutex> lock(g_const_mutex);
    g_const[ctx->device & 63].live--;
  }
  delete ctx;
}

const char *carta1_last_error(const carta1_ctx *ctx) { return ctx ? ctx->err.c_str() : g_create_error.c_str(); }
int carta1_ctx_sync(carta1_ctx *ctx) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}
void *carta1_ctx_stream(carta1_ctx *ctx) { return ctx ? (void *)ctx->stream : nullptr; }
uint64_t carta1_ctx_launch_count(const carta1_ctx *ctx) { return ctx ? ctx->prof.launches : 0; }

int carta1_kernel_count(void) { return K_COUNT; }
const char *carta1_kernel_name(int id) { return kernel_name(id); }

int carta1_ctx_profile(carta1_ctx *ctx, int enable) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  ctx->prof.on = enable != 0;
  return CARTA1_OK;
}

int carta1_ctx_profile_read(carta1_ctx *ctx, double *ms_out, uint64_t *count_out, int n) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  for (int id = 0; id < K_COUNT; id++) {
    double total = 0.0;
    for (auto &pr : ctx->prof.ev[id]) {
      float ms = 0.0f;
      cudaEventElapsedTime(&ms, pr.first, pr.second);
      total += ms;
      cudaEventDestroy(pr.first);
      cudaEventDestroy(pr.second);
    }
    if (id < n) {
      if (ms_out) ms_out[id] = total;
      if (count_out) count_out[id] = ctx->prof.ev[id].size();
    }
    ctx->prof.ev[id].clear();
  }
  return CARTA1_OK;
}
size_t carta1_frame_count(size_t n_samples) { return (n_samples + 511) / 512; }

int carta1_host_alloc(size_t bytes, void **out) {
  if (!out) return CARTA1_ERR_ARG;
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, bytes ? bytes : 1, cudaHostAllocPortable);
  if (e != cudaSuccess) return cuda_fail(nullptr, e, "cudaHostAlloc");
  return CARTA1_OK;
}
void carta1_host_free(void *p) { if (p) cudaFreeHost(p); }

int carta1_ctx_set_max_units_per_pass(carta1_ctx *ctx, size_t units) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  ctx->max_units_per_pass = units ? units : (size_t)1 << 16;
  return CARTA1_OK;
}

// ------------------------------------------------------------------ device-resident
static int encode_device_impl(carta1_ctx *ctx, const void *d_pcm, int pcm_fmt, size_t row_stride,
                              int n_ch_interleave, int n_streams, size_t valid_samples, size_t halo_frames,
                              size_t n_frames, const DevEncParams *d_params, bool use_fixed, uint8_t *d_su,
                              size_t su_frame_stride, size_t su_stream_stride, float *dbg_bands,
                              float *dbg_mags, uint8_t *dbg_modes, float *dbg_coefs, double *dbg_scores = nullptr) {
  const size_t frames_total = halo_frames + n_frames;
  const size_t units = frames_total * (size_t)n_streams;
  if (units == 0) return CARTA1_OK;
  if (units > 0x7fffffffull / 4) return fail(ctx, CARTA1_ERR_ARG, "carta1: too many sound units in one launch");
  int rc = ensure_encode_scratch(ctx, units, !use_fixed);
  if (rc) return rc;
  CU(ctx, ctx->recs.ensure((size_t)n_streams * n_frames * alloc_rec_bytes()));
  EncodeLaunch L;
  memset(&L, 0, sizeof L);
  L.pcm = d_pcm; L.pcm_fmt = pcm_fmt; L.row_stride = row_stride; L.n_ch_interleave = n_ch_interleave;
  L.valid_samples = (long long)valid_samples;
  L.n_streams = n_streams; L.frames_total = (int)frames_total; L.halo_frames = (int)halo_frames;
  L.n_out_frames = (int)n_frames; L.use_fixed = use_fixed ? 1 : 0;
  L.tables = ctx->d_tables; L.params = d_params;
  L.bands = dbg_bands ? dbg_bands : (float *)ctx->bands.p;
  L.mags = dbg_mags ? dbg_mags : (float *)ctx->mags.p;
  L.modes = dbg_modes ? dbg_modes : (uint8_t *)ctx->modes.p;
  L.coefs = dbg_coefs ? dbg_coefs : (float *)ctx->coefs.p;
  L.scores = dbg_scores;
  L.near_counts = ctx->d_near;
  if (!use_fixed) ctx->near_decisions += 3ull * (uint64_t)n_streams * n_frames;
  L.feats = ctx->feats.p;
  L.sfi = (uint8_t *)ctx->sfi.p;
  L.alloc_recs = ctx->recs.p;
  L.su_out = d_su; L.su_frame_stride = su_frame_stride; L.su_stream_stride = su_stream_stride;
  L.fj = &ctx->fj;
  CU(ctx, launch_encode(L, ctx->stream, &ctx->prof));
  return CARTA1_OK;
}

static int decode_device_impl(carta1_ctx *ctx, const uint8_t *d_su, size_t su_frame_stride,
                              size_t su_stream_stride, size_t n_su_valid, int n_streams, size_t halo_frames,
                              size_t n_frames, void *d_pcm, int pcm_fmt, size_t row_stride, int n_ch_interleave,
                              float *dbg_coefs, float *dbg_bands, const float *prev_rec = nullptr,
                              const int32_t *x_q = nullptr, const uint8_t *x_sfi = nullptr,
                              const uint8_t *x_bits = nullptr, const uint8_t *x_modes = nullptr, float *save_rec = nullptr) {
  const size_t frames_total = halo_frames + n_frames;
  const size_t units = frames_total * (size_t)n_streams;
  if (units == 0) return CARTA1_OK;
  if (units > 0x7fffffffull / 4) return fail(ctx, CARTA1_ERR_ARG, "carta1: too many sound units in one launch");
  int rc = ensure_decode_scratch(ctx, units);
  if (rc) return rc;
  DecodeLaunch L;
  memset(&L, 0, sizeof L);
  L.su = d_su; L.su_frame_stride = su_frame_stride; L.su_stream_stride = su_stream_stride;
  L.n_su_valid = (long long)n_su_valid; L.n_streams = n_streams; L.frames_total = (int)frames_total;
  L.halo_frames = (int)halo_frames; L.n_out_frames = (int)n_frames; L.tables = ctx->d_tables;
  L.coefs = dbg_coefs ? dbg_coefs : (float *)ctx->coefs.p;
  L.modes = (uint8_t *)ctx->modes.p;
  L.inv = (float *)ctx->inv.p;
  L.bands_dbg = dbg_bands;
  L.prev_rec = prev_rec;
  L.save_rec = save_rec;
  L.x_q = x_q; L.x_sfi = x_sfi; L.x_bits = x_bits; L.x_modes = x_modes;
  L.pcm = d_pcm; L.pcm_fmt = pcm_fmt; L.row_stride = row_stride; L.n_ch_interleave = n_ch_interleave;
  L.fj = &ctx->fj;
  CU(ctx, launch_decode(L, ctx->stream, &ctx->prof));
  return CARTA1_OK;
}

int carta1_encode_device(carta1_ctx *ctx, const float *d_pcm, size_t row_stride, int n_streams,
                         size_t valid_samples, size_t halo_frames, size_t n_frames,
                         const carta1_enc_opts *opts, uint8_t *d_su, size_t su_frame_stride,
                         size_t su_stream_stride, int sync) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (!d_pcm || !d_su || n_streams <= 0) return fail(ctx, CARTA1_ERR_ARG, "carta1_encode_device: bad argument");
  if (halo_frames == 1) return fail(ctx, CARTA1_ERR_ARG, "carta1_encode_device: halo_frames must be 0 or >= 2");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = upload_params(ctx, opts, ctx->d_params);
  if (rc) return rc;
  const bool fixed = opts && opts->use_fixed_block_modes;
  rc = encode_device_impl(ctx, d_pcm, 0, row_stride, 1, n_streams, valid_samples, halo_frames, n_frames,
                          ctx->d_params, fixed, d_su, su_frame_stride, su_stream_stride, nullptr, nullptr,
                          nullptr, nullptr);
  if (rc) return rc;
  if (sync) CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}

int carta1_decode_device(carta1_ctx *ctx, const uint8_t *d_su, size_t su_frame_stride,
                         size_t su_stream_stride, size_t n_su_valid, int n_streams, size_t halo_frames,
                         size_t n_frames, float *d_pcm, size_t row_stride, int sync) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (!d_su || !d_pcm || n_streams <= 0) return fail(ctx, CARTA1_ERR_ARG, "carta1_decode_device: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = decode_device_impl(ctx, d_su, su_frame_stride, su_stream_stride, n_su_valid, n_streams, halo_frames,
                              n_frames, d_pcm, 0, row_stride, 1, nullptr, nullptr);
  if (rc) return rc;
  if (sync) CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}

// ------------------------------------------------------------------ whole buffers (host)
// The sound-unit side of a host call is a tenth of the PCM side and travels in the opposite
// direction.  When an encode and a decode call are in flight together (two contexts), a copy-engine
// D2H of the units queues behind the other call's whole backlog of PCM D2H copies (measured with
// CARTA1_TRACE_PASSES, tools/e2e_probe.py: every encode D2H waited ~17 ms), so the encoder's
// pipeline drains.  The encoder therefore writes its units with a copy *kernel* straight into the
// caller's buffer when that buffer is pinned (mapped into the device address space by UVA); posted
// PCIe writes from the SMs are not held up by the copy engines.  The reverse does not hold: SM
// *reads* of pinned memory starve behind a copy engine's H2D stream (measured: 24 ms for 14 MB),
// so the decoder's unit upload stays on a copy engine, on a highest-priority stream.  Pageable
// buffers always take cudaMemcpyAsync.
__global__ void __launch_bounds__(256) small_copy_kernel(uint4 *__restrict__ dst, const uint4 *__restrict__ src, size_t n16,
                                                         uint8_t *__restrict__ dst_tail, const uint8_t *__restrict__ src_tail, int n_tail) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < n16; i += stride) dst[i] = src[i];
  if (blockIdx.x == 0 && (int)threadIdx.x < n_tail) dst_tail[threadIdx.x] = src_tail[threadIdx.x];
}

// Strided row copy on the device (float4 granules): dst[r][0..width) = src[r][0..width).  The stateful handles
// move one short row per stream and call; a copy-engine 2D copy pays per row, this is one small launch.
__global__ void __launch_bounds__(256) copy_rows_kernel(float4 *__restrict__ dst, size_t dst_pitch4, const float4 *__restrict__ src,
                                                        size_t src_pitch4, int width4, size_t n_rows) {
  const size_t total = n_rows * (size_t)width4;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const size_t r = i / (size_t)width4;
    const int c = (int)(i - r * (size_t)width4);
    dst[r * dst_pitch4 + c] = src[r * src_pitch4 + c];
  }
}
// widths and pitches in floats, all multiples of 4, both pointers 16-byte aligned
static cudaError_t copy_rows(float *dst, size_t dst_pitch, const float *src, size_t src_pitch, int width, size_t n_rows,
                             cudaStream_t st, Prof *prof) {
  if (!n_rows) return cudaSuccess;
  const size_t total4 = n_rows * (size_t)(width / 4);
  const int grid = (int)std::min<size_t>(1184, (total4 + 255) / 256);
  copy_rows_kernel<<<grid, 256, 0, st>>>((float4 *)dst, dst_pitch / 4, (const float4 *)src, src_pitch / 4, width / 4, n_rows);
  prof->launches++;
  return cudaGetLastError();
}

// Device-visible alias of a pinned host pointer, or nullptr.
static void *mapped_alias(const void *host) {
  cudaPointerAttributes a;
  if (cudaPointerGetAttributes(&a, host) != cudaSuccess) { cudaGetLastError(); return nullptr; }
  if (a.type != cudaMemoryTypeHost || !a.devicePointer) return nullptr;
  return a.devicePointer;
}

// Development switch CARTA1_SMALL_COPY: 0 copy engines on the h2d/d2h streams, 1 copy engines on the
// priority stream, 2 copy kernel in both directions, 3 (default) copy kernel for writes to the host only.
static int small_copy_mode() {
  static const int mode = [] {  // initialised once, thread-safe
    const char *v = getenv("CARTA1_SMALL_COPY");
    return v && *v ? atoi(v) : 3;
  }();
  return mode;
}

// Copies `bytes` between the caller's unit buffer and a staging slot on stream st.
static cudaError_t small_copy(void *dst, const void *src, size_t bytes, cudaMemcpyKind kind, void *dst_alias,
                              const void *src_alias, cudaStream_t st, Prof *prof) {
  const bool to_host = kind == cudaMemcpyDeviceToHost;
  void *d = to_host ? dst_alias : dst;
  const void *s = to_host ? src : src_alias;
  const int mode = small_copy_mode();
  if (mode < 2 || (mode == 3 && !to_host) || !d || !s || (((uintptr_t)d | (uintptr_t)s) & 15))
    return cudaMemcpyAsync(dst, src, bytes, kind, st);
  const size_t n16 = bytes / 16;
  const int grid = (int)std::min<size_t>(296, (n16 + 255) / 256 + 1);
  small_copy_kernel<<<grid, 256, 0, st>>>((uint4 *)d, (const uint4 *)s, n16, (uint8_t *)d + n16 * 16,
                                          (const uint8_t *)s + n16 * 16, (int)(bytes - n16 * 16));
  prof->launches++;
  return cudaGetLastError();
}

// A pipelined host call that fails half-way returns while earlier passes may still be copying into or out of
// the caller's buffers and the context's slots: wait for all four streams before the caller gets the error.
static void quiesce(carta1_ctx *ctx) {
  if (!ctx || cudaSetDevice(ctx->device) != cudaSuccess) return;
  if (ctx->h2d) cudaStreamSynchronize(ctx->h2d);
  if (ctx->stream) cudaStreamSynchronize(ctx->stream);
  if (ctx->small) cudaStreamSynchronize(ctx->small);
  if (ctx->d2h) cudaStreamSynchronize(ctx->d2h);
  cudaGetLastError();
}

// halo0: frames of real history at the head of the caller's PCM that are not emitted (0, or >= 2: a shard of a
// longer stream); the whole-buffer entry points pass 0.
static int encode_host_body(carta1_ctx *ctx, const float *const *channels, const int16_t *interleaved,
                            int n_ch, size_t n_samples, size_t halo0, const carta1_enc_opts *opts, uint8_t *su_out,
                            size_t su_capacity_bytes, size_t *n_su_out) {
  if (!ctx) return CARTA1_ERR_ARG;
  // processor.js:598-604
  if ((n_ch != 1 && n_ch != 2) || (!channels && !interleaved))
    return fail(ctx, CARTA1_ERR_ARG, "ATRAC1 encoding requires one or two Float32 channels");
  if (channels)
    for (int c = 0; c < n_ch; c++)
      if (!channels[c] && n_samples)
        return fail(ctx, CARTA1_ERR_ARG, "ATRAC1 encoding requires one or two Float32 channels");
  const size_t frames = carta1_frame_count(n_samples);  // halo included
  if (halo0 == 1) return fail(ctx, CARTA1_ERR_ARG, "carta1_encode_pcm_shard: halo_frames must be 0 or >= 2");
  const size_t n_su = (frames > halo0 ? frames - halo0 : 0) * (size_t)n_ch;
  if (n_su_out) *n_su_out = n_su;
  if (n_su == 0) return CARTA1_OK;
  if (!su_out || su_capacity_bytes < n_su * CARTA1_SU_BYTES)
    return fail(ctx, CARTA1_ERR_ARG, "carta1_encode_pcm: output buffer too small");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = upload_params(ctx, opts, ctx->d_params);
  if (rc) return rc;
  const bool fixed = opts && opts->use_fixed_block_modes;
  const size_t chunk = std::max<size_t>(2, ctx->max_units_per_pass / (size_t)n_ch);
  // size every buffer for the largest pass up front: nothing may be reallocated while passes are in flight
  const size_t max_span = (std::min(frames, chunk) + 2) * 512;
  const size_t in_elem = channels ? sizeof(float) : sizeof(int16_t);
  for (int sl = 0; sl < kSlots; sl++) {
    CU(ctx, ctx->stage_pcm[sl].ensure((size_t)n_ch * max_span * in_elem));
    CU(ctx, ctx->stage_su[sl].ensure(std::min(frames, chunk) * (size_t)n_ch * CARTA1_SU_BYTES));
  }
  rc = ensure_encode_scratch(ctx, (std::min(frames, chunk) + 2) * (size_t)n_ch, !fixed);
  if (rc) return rc;
  CU(ctx, ctx->recs.ensure(std::min(frames, chunk) * (size_t)n_ch * alloc_rec_bytes()));
  size_t pass = 0;
  uint8_t *su_alias = (uint8_t *)mapped_alias(su_out);
  cudaStream_t s_out = small_copy_mode() == 0 ? ctx->d2h : ctx->small;
  // Pageable caller arrays (anything a host runtime did not pin): cudaMemcpyAsync would stage them through
  // the driver on one thread and block this one (measured 125 ms instead of 24 for the 1 h stereo encode),
  // so they go through the context's pinned bounce slots, filled by parallel_memcpy while earlier passes run.
  const size_t kBounceMinBytes = bounce_min_bytes();
  const size_t pcm_bytes = n_samples * (size_t)n_ch * in_elem;
  bool bounce_in = pcm_bytes >= kBounceMinBytes;
  if (bounce_in) {
    bool pinned = true;
    if (channels) { for (int c = 0; c < n_ch; c++) pinned = pinned && mapped_alias(channels[c]); }
    else pinned = mapped_alias(interleaved) != nullptr;
    bounce_in = !pinned;
  }
  const bool bounce_out = !su_alias && n_su * CARTA1_SU_BYTES >= kBounceMinBytes / 8;
  for (int sl = 0; sl < kSlots; sl++) {
    if (bounce_in) CU(ctx, ctx->bounce_pcm[sl].ensure((size_t)n_ch * max_span * in_elem));
    if (bounce_out) CU(ctx, ctx->bounce_su[sl].ensure(std::min(frames, chunk) * (size_t)n_ch * CARTA1_SU_BYTES));
  }
  struct Pending { bool on = false; int sl = 0; size_t off = 0, bytes = 0; } pend;  // unit bounce slot to drain
  auto drain = [&]() -> cudaError_t {
    if (!pend.on) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(ctx->ev_out[pend.sl]);
    if (e == cudaSuccess) parallel_memcpy(su_out + pend.off, ctx->bounce_su[pend.sl].p, pend.bytes);
    pend.on = false;
    return e;
  };
  PassTrace tr;
  tr.start("encode", ctx->h2d);
  for (size_t a = halo0; a < frames; a += chunk, pass++) {
    const int sl = (int)(pass % kSlots);
    const size_t b = std::min(frames, a + chunk);
    const size_t halo = a >= 2 ? 2 : 0;  // a is 0, halo0 (>= 2) or >= chunk (>= 2)
    const size_t first = a - halo;
    const size_t span = (b - first) * 512;                       // samples staged per row
    const size_t have = std::min(n_samples - first * 512, span); // samples that exist
    const void *src_rows[2] = {nullptr, nullptr};                // what the H2D copies read
    if (bounce_in) {
      if (pass >= (size_t)kSlots) CU(ctx, cudaEventSynchronize(ctx->ev_bin[sl]));  // the H2D of pass - kSlots has read the slot
      char *bb = (char *)ctx->bounce_pcm[sl].p;
      if (channels) {
        for (int c = 0; c < n_ch; c++) {
          parallel_memcpy(bb + (size_t)c * span * sizeof(float), channels[c] + first * 512, have * sizeof(float));
          src_rows[c] = bb + (size_t)c * span * sizeof(float);
        }
      } else {
        parallel_memcpy(bb, interleaved + first * 512 * (size_t)n_ch, have * (size_t)n_ch * sizeof(int16_t));
        src_rows[0] = bb;
      }
    } else if (channels) {
      for (int c = 0; c < n_ch; c++) src_rows[c] = channels[c] + first * 512;
    } else {
      src_rows[0] = interleaved + first * 512 * (size_t)n_ch;
    }
    // H2D of this pass: its device slot was last read by the compute of pass - kSlots
    if (pass >= (size_t)kSlots) CU(ctx, cudaStreamWaitEvent(ctx->h2d, ctx->ev_comp[sl], 0));
    tr.mark(ctx->h2d);
    if (channels) {
      for (int c = 0; c < n_ch; c++)
        CU(ctx, cudaMemcpyAsync((float *)ctx->stage_pcm[sl].p + (size_t)c * span, src_rows[c], have * sizeof(float),
                                cudaMemcpyHostToDevice, ctx->h2d));
    } else {
      CU(ctx, cudaMemcpyAsync(ctx->stage_pcm[sl].p, src_rows[0], have * (size_t)n_ch * sizeof(int16_t),
                              cudaMemcpyHostToDevice, ctx->h2d));
    }
    CU(ctx, cudaEventRecord(ctx->ev_in[sl], ctx->h2d));
    if (bounce_in) CU(ctx, cudaEventRecord(ctx->ev_bin[sl], ctx->h2d));
    tr.mark(ctx->h2d);
    // compute: needs the input, and its output slot drained by the D2H of pass - kSlots
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_in[sl], 0));
    if (pass >= (size_t)kSlots) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[sl], 0));
    tr.mark(ctx->stream);
    const size_t out_units = (b - a) * (size_t)n_ch;
    rc = encode_device_impl(ctx, ctx->stage_pcm[sl].p, channels ? 0 : 1, span, n_ch, n_ch, have, halo, b - a,
                            ctx->d_params, fixed, (uint8_t *)ctx->stage_su[sl].p, (size_t)n_ch, 1, nullptr,
                            nullptr, nullptr, nullptr);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->ev_comp[sl], ctx->stream));
    tr.mark(ctx->stream);
    CU(ctx, cudaStreamWaitEvent(s_out, ctx->ev_comp[sl], 0));
    tr.mark(s_out);
    const size_t off = (a - halo0) * (size_t)n_ch * CARTA1_SU_BYTES;
    if (bounce_out) {
      uint8_t *bs = (uint8_t *)ctx->bounce_su[sl].p;  // drained (pass - 1 at the latest) before it is reused
      CU(ctx, small_copy(bs, ctx->stage_su[sl].p, out_units * CARTA1_SU_BYTES, cudaMemcpyDeviceToHost,
                         (uint8_t *)mapped_alias(bs), nullptr, s_out, &ctx->prof));
    } else {
      CU(ctx, small_copy(su_out + off, ctx->stage_su[sl].p, out_units * CARTA1_SU_BYTES, cudaMemcpyDeviceToHost,
                         su_alias ? su_alias + off : nullptr, nullptr, s_out, &ctx->prof));
    }
    CU(ctx, cudaEventRecord(ctx->ev_out[sl], s_out));
    tr.mark(s_out);
    if (bounce_out) {
      CU(ctx, drain());  // the previous pass, while this one runs
      pend.on = true; pend.sl = sl; pend.off = off; pend.bytes = out_units * CARTA1_SU_BYTES;
    }
  }
  CU(ctx, drain());
  CU(ctx, cudaStreamSynchronize(s_out));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  tr.report();
  return CARTA1_OK;
}

static int encode_host_impl(carta1_ctx *ctx, const float *const *channels, const int16_t *interleaved,
                            int n_ch, size_t n_samples, const carta1_enc_opts *opts, uint8_t *su_out,
                            size_t su_capacity_bytes, size_t *n_su_out, size_t halo0 = 0) {
  const int rc = encode_host_body(ctx, channels, interleaved, n_ch, n_samples, halo0, opts, su_out, su_capacity_bytes, n_su_out);
  if (rc != CARTA1_OK) quiesce(ctx);
  return rc;
}

int carta1_encode_pcm(carta1_ctx *ctx, const float *const *channels, int n_ch, size_t n_samples,
                      const carta1_enc_opts *opts, uint8_t *su_out, size_t su_capacity_bytes,
                      size_t *n_su_out) {
  CtxLock ctx_lock(ctx);
  return encode_host_impl(ctx, channels, nullptr, n_ch, n_samples, opts, su_out, su_capacity_bytes, n_su_out);
}

int carta1_encode_pcm_s16(carta1_ctx *ctx, const int16_t *interleaved, int n_ch, size_t n_samples,
                          const carta1_enc_opts *opts, uint8_t *su_out, size_t su_capacity_bytes,
                          size_t *n_su_out) {
  CtxLock ctx_lock(ctx);
  return encode_host_impl(ctx, nullptr, interleaved, n_ch, n_samples, opts, su_out, su_capacity_bytes, n_su_out);
}

int carta1_encode_pcm_shard(carta1_ctx *ctx, const float *const *channels, int n_ch, size_t n_samples,
                            size_t halo_frames, const carta1_enc_opts *opts, uint8_t *su_out,
                            size_t su_capacity_bytes, size_t *n_su_out) {
  CtxLock ctx_lock(ctx);
  return encode_host_impl(ctx, channels, nullptr, n_ch, n_samples, opts, su_out, su_capacity_bytes, n_su_out, halo_frames);
}

// halo0: frames of sound units at the head of `su` that are history only (a shard of a longer file); their
// PCM is not written, channels_out[c] starts at the first emitted frame.
static int decode_host_body(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch, size_t halo0,
                            float *const *channels_out, int16_t *interleaved_out) {
  if (!ctx) return CARTA1_ERR_ARG;
  if (n_ch != 1 && n_ch != 2) {  // processor.js:147-157
    char msg[64];
    snprintf(msg, sizeof msg, "Unsupported channel count: %d", n_ch);
    return fail(ctx, CARTA1_ERR_ARG, msg);
  }
  if (n_su && !su) return fail(ctx, CARTA1_ERR_ARG, "ATRAC1 decoding requires AEA bytes or a Blob");
  const size_t frames = (n_su + (size_t)n_ch - 1) / (size_t)n_ch;  // halo included
  if (frames <= halo0) return CARTA1_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t chunk = std::max<size_t>(2, ctx->max_units_per_pass / (size_t)n_ch);
  const size_t max_frames = std::min(frames - halo0, chunk);
  const size_t out_elem = channels_out ? sizeof(float) : sizeof(int16_t);
  const size_t n_passes = (frames - halo0 + chunk - 1) / chunk;
  for (int sl = 0; sl < kSlots; sl++) CU(ctx, ctx->stage_pcm[sl].ensure((size_t)n_ch * max_frames * 512 * out_elem));
  for (size_t us = 0; us < std::min<size_t>(n_passes, kUnitSlots); us++)
    CU(ctx, ctx->stage_su[us].ensure((max_frames + 1) * (size_t)n_ch * CARTA1_SU_BYTES));
  int rc = ensure_decode_scratch(ctx, (max_frames + 1) * (size_t)n_ch);
  if (rc) return rc;
  size_t pass = 0;
  const uint8_t *su_alias = (const uint8_t *)mapped_alias(su);
  cudaStream_t s_in = small_copy_mode() == 0 ? ctx->h2d : ctx->small;
  // pageable caller arrays go through the pinned bounce slots (see encode_host_impl)
  const size_t kBounceMinBytes = bounce_min_bytes();
  const bool bounce_in = !su_alias && n_su * CARTA1_SU_BYTES >= kBounceMinBytes / 8;
  const size_t pcm_bytes = (frames - halo0) * 512 * (size_t)n_ch * out_elem;
  bool bounce_out = pcm_bytes >= kBounceMinBytes;
  if (bounce_out) {
    bool pinned = true;
    if (channels_out) { for (int c = 0; c < n_ch; c++) pinned = pinned && channels_out[c] && mapped_alias(channels_out[c]); }
    else pinned = mapped_alias(interleaved_out) != nullptr;
    bounce_out = !pinned;
  }
  for (int sl = 0; sl < kSlots; sl++) {
    if (bounce_in) CU(ctx, ctx->bounce_su[sl].ensure((max_frames + 1) * (size_t)n_ch * CARTA1_SU_BYTES));
    if (bounce_out) CU(ctx, ctx->bounce_pcm[sl].ensure((size_t)n_ch * max_frames * 512 * out_elem));
  }
  struct Pending { bool on = false; int sl = 0; size_t a = 0, span = 0; } pend;  // PCM bounce slot to drain
  auto drain = [&]() -> cudaError_t {
    if (!pend.on) return cudaSuccess;
    cudaError_t e = cudaEventSynchronize(ctx->ev_out[pend.sl]);
    if (e == cudaSuccess) {
      const char *bb = (const char *)ctx->bounce_pcm[pend.sl].p;
      if (channels_out) {
        for (int c = 0; c < n_ch; c++)
          parallel_memcpy(channels_out[c] + (pend.a - halo0) * 512, bb + (size_t)c * pend.span * sizeof(float), pend.span * sizeof(float));
      } else {
        parallel_memcpy(interleaved_out + (pend.a - halo0) * 512 * (size_t)n_ch, bb, (size_t)n_ch * pend.span * sizeof(int16_t));
      }
    }
    pend.on = false;
    return e;
  };
  PassTrace tr;
  tr.start("decode", s_in);
  for (size_t a = halo0; a < frames; a += chunk, pass++) {
    const int sl = (int)(pass % kSlots), us = (int)(pass % kUnitSlots);
    const size_t b = std::min(frames, a + chunk);
    const size_t halo = a >= 1 ? 1 : 0;
    const size_t first = a - halo;
    const size_t want_units = (b - first) * (size_t)n_ch;
    const size_t have_units = std::min(n_su - first * (size_t)n_ch, want_units);
    const size_t off = first * (size_t)n_ch * CARTA1_SU_BYTES;
    const uint8_t *src = su + off, *src_alias = su_alias ? su_alias + off : nullptr;
    if (bounce_in) {
      if (pass >= (size_t)kSlots) CU(ctx, cudaEventSynchronize(ctx->ev_bin[sl]));
      parallel_memcpy(ctx->bounce_su[sl].p, su + off, have_units * CARTA1_SU_BYTES);
      src = (const uint8_t *)ctx->bounce_su[sl].p;
      src_alias = (const uint8_t *)mapped_alias(src);
    }
    if (pass >= (size_t)kUnitSlots) CU(ctx, cudaStreamWaitEvent(s_in, ctx->ev_ucomp[us], 0));
    tr.mark(s_in);
    CU(ctx, small_copy(ctx->stage_su[us].p, src, have_units * CARTA1_SU_BYTES, cudaMemcpyHostToDevice, nullptr, src_alias,
                       s_in, &ctx->prof));
    CU(ctx, cudaEventRecord(ctx->ev_uin[us], s_in));
    if (bounce_in) CU(ctx, cudaEventRecord(ctx->ev_bin[sl], s_in));
    tr.mark(s_in);
    CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_uin[us], 0));
    if (pass >= (size_t)kSlots) CU(ctx, cudaStreamWaitEvent(ctx->stream, ctx->ev_out[sl], 0));
    tr.mark(ctx->stream);
    const size_t span = (b - a) * 512;
    rc = decode_device_impl(ctx, (const uint8_t *)ctx->stage_su[us].p, (size_t)n_ch, 1, have_units, n_ch, halo,
                            b - a, ctx->stage_pcm[sl].p, channels_out ? 0 : 1, span, n_ch, nullptr, nullptr);
    if (rc) return rc;
    CU(ctx, cudaEventRecord(ctx->ev_comp[sl], ctx->stream));
    CU(ctx, cudaEventRecord(ctx->ev_ucomp[us], ctx->stream));
    tr.mark(ctx->stream);
    CU(ctx, cudaStreamWaitEvent(ctx->d2h, ctx->ev_comp[sl], 0));
    tr.mark(ctx->d2h);
    if (bounce_out) {  // the slot was drained at pass - 3 at the latest
      CU(ctx, cudaMemcpyAsync(ctx->bounce_pcm[sl].p, ctx->stage_pcm[sl].p, (size_t)n_ch * span * out_elem,
                              cudaMemcpyDeviceToHost, ctx->d2h));
    } else if (channels_out) {
      for (int c = 0; c < n_ch; c++)
        CU(ctx, cudaMemcpyAsync(channels_out[c] + (a - halo0) * 512, (float *)ctx->stage_pcm[sl].p + (size_t)c * span,
                                span * sizeof(float), cudaMemcpyDeviceToHost, ctx->d2h));
    } else {
      CU(ctx, cudaMemcpyAsync(interleaved_out + (a - halo0) * 512 * (size_t)n_ch, ctx->stage_pcm[sl].p,
                              (size_t)n_ch * span * sizeof(int16_t), cudaMemcpyDeviceToHost, ctx->d2h));
    }
    CU(ctx, cudaEventRecord(ctx->ev_out[sl], ctx->d2h));
    tr.mark(ctx->d2h);
    if (bounce_out) {
      CU(ctx, drain());  // the previous pass, while this one runs
      pend.on = true; pend.sl = sl; pend.a = a; pend.span = span;
    }
  }
  CU(ctx, drain());
  CU(ctx, cudaStreamSynchronize(ctx->d2h));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  tr.report();
  return CARTA1_OK;
}

static int decode_host_impl(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch,
                            float *const *channels_out, int16_t *interleaved_out, size_t halo0 = 0) {
  const int rc = decode_host_body(ctx, su, n_su, n_ch, halo0, channels_out, interleaved_out);
  if (rc != CARTA1_OK) quiesce(ctx);
  return rc;
}

int carta1_decode_su(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch, float *const *channels_out) {
  CtxLock ctx_lock(ctx);
  if (ctx && !channels_out) return fail(ctx, CARTA1_ERR_ARG, "carta1_decode_su: channels_out is NULL");
  return decode_host_impl(ctx, su, n_su, n_ch, channels_out, nullptr);
}
int carta1_decode_su_shard(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch, size_t halo_frames,
                           float *const *channels_out) {
  CtxLock ctx_lock(ctx);
  if (ctx && !channels_out) return fail(ctx, CARTA1_ERR_ARG, "carta1_decode_su_shard: channels_out is NULL");
  return decode_host_impl(ctx, su, n_su, n_ch, channels_out, nullptr, halo_frames);
}
int carta1_decode_su_s16(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch, int16_t *interleaved_out) {
  CtxLock ctx_lock(ctx);
  if (ctx && !interleaved_out) return fail(ctx, CARTA1_ERR_ARG, "carta1_decode_su_s16: output is NULL");
  return decode_host_impl(ctx, su, n_su, n_ch, nullptr, interleaved_out);
}

// ------------------------------------------------------------------ stateful closures
int carta1_enc_create(carta1_ctx *ctx, const carta1_enc_opts *opts, int n_streams, carta1_encoder **out) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (!out || n_streams <= 0) return fail(ctx, CARTA1_ERR_ARG, "carta1_enc_create: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  carta1_encoder *e = new carta1_encoder();
  e->ctx = ctx;
  e->n_streams = n_streams;
  if (opts) e->opts = *opts; else carta1_default_enc_opts(&e->opts);
  if (e->opts.biased_scale_factors) {
    memcpy(e->bsf_copy, e->opts.biased_scale_factors, sizeof e->bsf_copy);
    e->opts.biased_scale_factors = e->bsf_copy;
  }
  cudaError_t ce = cudaMalloc(&e->d_params, sizeof(DevEncParams));
  if (ce == cudaSuccess) ce = cudaMalloc(&e->d_hist, (size_t)n_streams * 1024 * sizeof(float));
  if (ce == cudaSuccess) ce = cudaMemsetAsync(e->d_hist, 0, (size_t)n_streams * 1024 * sizeof(float), ctx->stream);
  if (ce != cudaSuccess) { carta1_enc_destroy(e); return cuda_fail(ctx, ce, "carta1_enc_create"); }
  int rc = upload_params(ctx, &e->opts, e->d_params);
  if (rc) { carta1_enc_destroy(e); return rc; }
  *out = e;
  return CARTA1_OK;
}

void carta1_enc_destroy(carta1_encoder *e) {
  CtxLock ctx_lock(e ? e->ctx : nullptr);
  if (!e) return;
  cudaSetDevice(e->ctx->device);
  cudaStreamSynchronize(e->ctx->stream);
  if (e->d_params) cudaFree(e->d_params);
  if (e->d_hist) cudaFree(e->d_hist);
  e->work.release(); e->su.release();
  e->graph.drop();
  delete e;
}

int carta1_enc_reset(carta1_encoder *e) {
  CtxLock ctx_lock(e ? e->ctx : nullptr);
  if (!e) return CARTA1_ERR_ARG;
  carta1_ctx *ctx = e->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, cudaMemsetAsync(e->d_hist, 0, (size_t)e->n_streams * 1024 * sizeof(float), ctx->stream));
  return CARTA1_OK;
}

// The encoder's carried state (BufferPool: QMF delays, MDCT overlap, previous magnitude
// spectra) is a function of the last 650 PCM samples, so the handle
// keeps the last two frames of PCM per stream and re-derives it.
int carta1_enc_frames(carta1_encoder *e, const float *pcm, int n_frames, uint8_t *su_out) {
  CtxLock ctx_lock(e ? e->ctx : nullptr);
  if (!e) return CARTA1_ERR_ARG;
  carta1_ctx *ctx = e->ctx;
  if (n_frames < 0 || (n_frames && (!pcm || !su_out))) return fail(ctx, CARTA1_ERR_ARG, "carta1_enc_frames: bad argument");
  if (n_frames == 0) return CARTA1_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t ns = (size_t)e->n_streams;
  const size_t row = (size_t)(n_frames + 2) * 512;
  CU(ctx, e->work.ensure(ns * row * sizeof(float)));
  CU(ctx, e->su.ensure(ns * (size_t)n_frames * CARTA1_SU_BYTES));
  float *w = (float *)e->work.p;
  CU(ctx, cudaMemcpy2DAsync(w + 1024, row * sizeof(float), pcm, (size_t)n_frames * 512 * sizeof(float),
                            (size_t)n_frames * 512 * sizeof(float), ns, cudaMemcpyHostToDevice, ctx->stream));
  int rc = run_graphed(ctx, e->graph, n_frames, [&]() -> int {
    CU(ctx, copy_rows(w, row, e->d_hist, 1024, 1024, ns, ctx->stream, &ctx->prof));
    // the history of the next call (the last two frames of w) is saved beside the kernels, not behind them: it only
    // reads w, as they do; the handle's second stream carries it and joins before the call returns
    CU(ctx, fork_begin(&ctx->fj, ctx->stream));
    CU(ctx, copy_rows(e->d_hist, 1024, w + (size_t)n_frames * 512, row, 1024, ns, ctx->fj.aux, &ctx->prof));
    cudaEvent_t hist_saved = ctx->ev_hist;
    CU(ctx, cudaEventRecord(hist_saved, ctx->fj.aux));
    const int r = encode_device_impl(ctx, w, 0, row, 1, e->n_streams, row, 2, (size_t)n_frames, e->d_params,
                                     e->opts.use_fixed_block_modes != 0, (uint8_t *)e->su.p, 1, (size_t)n_frames,
                                     nullptr, nullptr, nullptr, nullptr);
    CU(ctx, cudaStreamWaitEvent(ctx->stream, hist_saved, 0));  // joins the second stream (also after a failed launch)
    if (r) return r;
    return CARTA1_OK;
  });
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(su_out, e->su.p, ns * (size_t)n_frames * CARTA1_SU_BYTES, cudaMemcpyDeviceToHost,
                          ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}

int carta1_dec_create(carta1_ctx *ctx, int n_streams, carta1_decoder **out) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (!out || n_streams <= 0) return fail(ctx, CARTA1_ERR_ARG, "carta1_dec_create: bad argument");
  CU(ctx, cudaSetDevice(ctx->device));
  carta1_decoder *d = new carta1_decoder();
  d->ctx = ctx;
  d->n_streams = n_streams;
  cudaError_t ce = cudaMalloc(&d->d_rec, (size_t)n_streams * 512 * sizeof(float));
  if (ce != cudaSuccess) { delete d; return cuda_fail(ctx, ce, "carta1_dec_create"); }
  *out = d;
  return CARTA1_OK;
}

void carta1_dec_destroy(carta1_decoder *d) {
  CtxLock ctx_lock(d ? d->ctx : nullptr);
  if (!d) return;
  cudaSetDevice(d->ctx->device);
  cudaStreamSynchronize(d->ctx->stream);
  if (d->d_rec) cudaFree(d->d_rec);
  d->work.release(); d->pcm.release();
  d->graph.drop();
  delete d;
}

int carta1_dec_reset(carta1_decoder *d) {
  CtxLock ctx_lock(d ? d->ctx : nullptr);
  if (!d) return CARTA1_ERR_ARG;
  d->has_prev = false;
  return CARTA1_OK;
}

// The decoder's carried state (QMF delays, IMDCT tails) is a function of the previous sound
// unit alone; the handle keeps that unit's band record (its IMDCT
// output, 512 floats per stream), which is all the synthesis kernel reads of it.
static int dec_frames_impl(carta1_decoder *d, const uint8_t *su, const int32_t *x_q, const uint8_t *x_sfi,
                           const uint8_t *x_bits, const int32_t *x_modes, int n_frames, float *pcm_out) {
  carta1_ctx *ctx = d->ctx;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t ns = (size_t)d->n_streams;
  const size_t nf = (size_t)n_frames;
  const size_t halo = d->has_prev ? 1 : 0;
  CU(ctx, d->pcm.ensure(ns * nf * 512 * sizeof(float)));
  const uint8_t *d_su = nullptr;
  const int32_t *dq = nullptr;
  const uint8_t *dsfi = nullptr, *dbits = nullptr, *dmodes = nullptr;
  if (su) {
    CU(ctx, d->work.ensure(ns * nf * CARTA1_SU_BYTES));
    CU(ctx, cudaMemcpyAsync(d->work.p, su, ns * nf * CARTA1_SU_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    d_su = (const uint8_t *)d->work.p;  // [stream][frame][212]
  } else {
    const size_t n = ns * nf;
    CU(ctx, d->work.ensure(n * 512 * 6 + n * 4 + 64));
    uint8_t *w = (uint8_t *)d->work.p;
    std::vector<uint8_t> m(n * 4, 0);
    for (size_t i = 0; i < n; i++)
      for (int b = 0; b < 3; b++) m[i * 4 + b] = x_modes[i * 3 + b] != 0;
    CU(ctx, cudaMemcpyAsync(w, x_q, n * 512 * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(w + n * 2048, x_sfi, n * 512, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(w + n * 2560, x_bits, n * 512, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaMemcpyAsync(w + n * 3072, m.data(), n * 4, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));  // m lives on this stack frame
    dq = (const int32_t *)w; dsfi = w + n * 2048; dbits = w + n * 2560; dmodes = w + n * 3072;
  }
  auto body = [&]() -> int {
    // K7 leaves the band record of every stream's last unit in d_rec (DecodeLaunch::save_rec): no row copy behind it
    return decode_device_impl(ctx, d_su, 1, nf, ns * nf, d->n_streams, halo, nf, d->pcm.p, 0, nf * 512, 1,
                              nullptr, nullptr, halo ? d->d_rec : nullptr, dq, dsfi, dbits, dmodes, d->d_rec);
  };
  // sound units with a previous unit in the handle: the steady state of a stream, worth a graph
  const int rc = su && halo ? run_graphed(ctx, d->graph, n_frames, body) : body();
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(pcm_out, d->pcm.p, ns * nf * 512 * sizeof(float), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  d->has_prev = true;
  return CARTA1_OK;
}

int carta1_dec_frames(carta1_decoder *d, const uint8_t *su, int n_frames, float *pcm_out) {
  CtxLock ctx_lock(d ? d->ctx : nullptr);
  if (!d) return CARTA1_ERR_ARG;
  if (n_frames < 0 || (n_frames && (!su || !pcm_out)))
    return fail(d->ctx, CARTA1_ERR_ARG, "carta1_dec_frames: bad argument");
  if (n_frames == 0) return CARTA1_OK;
  return dec_frames_impl(d, su, nullptr, nullptr, nullptr, nullptr, n_frames, pcm_out);
}

int carta1_dec_frames_expanded(carta1_decoder *d, const int32_t *q, const uint8_t *sfi, const uint8_t *bits,
                               const int32_t *modes, int n_frames, float *pcm_out) {
  CtxLock ctx_lock(d ? d->ctx : nullptr);
  if (!d) return CARTA1_ERR_ARG;
  if (n_frames < 0 || (n_frames && (!q || !sfi || !bits || !modes || !pcm_out)))
    return fail(d->ctx, CARTA1_ERR_ARG, "carta1_dec_frames_expanded: bad argument");
  if (n_frames == 0) return CARTA1_OK;
  const size_t n = (size_t)d->n_streams * (size_t)n_frames * 512;
  for (size_t i = 0; i < n; i++)
    if (sfi[i] > 63 || bits[i] == 1 || bits[i] > 16)
      return fail(d->ctx, CARTA1_ERR_ARG, "carta1_dec_frames_expanded: scale-factor index or bit width out of range");
  return dec_frames_impl(d, nullptr, q, sfi, bits, modes, n_frames, pcm_out);
}

// ------------------------------------------------------------------ frame dump
int carta1_deserialize_units(carta1_ctx *ctx, const uint8_t *su, size_t n_su, uint8_t *n_bfu, int8_t *block_modes,
                             uint8_t *wl, uint8_t *sfi, int32_t *q) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (n_su == 0) return CARTA1_OK;
  if (!su || !n_bfu || !block_modes || !wl || !sfi || !q) return fail(ctx, CARTA1_ERR_ARG, "carta1_deserialize_units: NULL argument");
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t chunk = std::max<size_t>(1, ctx->max_units_per_pass);
  const size_t m = std::min(n_su, chunk);
  // one buffer: units | q | wl | sfi | modes | n_bfu
  const size_t o_q = (m * CARTA1_SU_BYTES + 255) & ~(size_t)255, o_wl = o_q + m * 2048, o_sfi = o_wl + m * 52,
               o_modes = o_sfi + m * 52, o_n = o_modes + m * 3;
  CU(ctx, ctx->dbg.ensure(o_n + m));
  uint8_t *d = (uint8_t *)ctx->dbg.p;
  for (size_t a = 0; a < n_su; a += chunk) {
    const size_t k = std::min(chunk, n_su - a);
    CU(ctx, cudaMemcpyAsync(d, su + a * CARTA1_SU_BYTES, k * CARTA1_SU_BYTES, cudaMemcpyHostToDevice, ctx->stream));
    CU(ctx, launch_deserialize(d, (int)k, ctx->d_tables, d + o_n, (int8_t *)(d + o_modes), d + o_wl, d + o_sfi,
                               (int32_t *)(d + o_q), ctx->stream, &ctx->prof));
    CU(ctx, cudaMemcpyAsync(q + a * 512, d + o_q, k * 2048, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(wl + a * 52, d + o_wl, k * 52, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(sfi + a * 52, d + o_sfi, k * 52, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(block_modes + a * 3, d + o_modes, k * 3, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaMemcpyAsync(n_bfu + a, d + o_n, k, cudaMemcpyDeviceToHost, ctx->stream));
    CU(ctx, cudaStreamSynchronize(ctx->stream));
  }
  return CARTA1_OK;
}

// ------------------------------------------------------------------ stage taps
int carta1_debug_encode_stages(carta1_ctx *ctx, const float *pcm, size_t n_samples, const carta1_enc_opts *opts,
                               float *bands, float *mags, int32_t *modes, float *coefs, uint8_t *su) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  const size_t frames = carta1_frame_count(n_samples);
  if (frames == 0) return CARTA1_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = upload_params(ctx, opts, ctx->d_params);
  if (rc) return rc;
  const bool fixed = opts && opts->use_fixed_block_modes;
  // layout: pcm | bands | mags | coefs | su | modes
  const size_t o_bands = frames * 512, o_mags = o_bands + frames * 512, o_coefs = o_mags + frames * 256,
               o_su = o_coefs + frames * 512, o_modes = (o_su + (frames * 212 + 3) / 4 + 4) & ~(size_t)3,
               total = o_modes + frames + 1;
  CU(ctx, ctx->dbg.ensure(total * sizeof(float)));
  float *base = (float *)ctx->dbg.p;
  CU(ctx, cudaMemsetAsync(base, 0, total * sizeof(float), ctx->stream));
  CU(ctx, cudaMemcpyAsync(base, pcm, n_samples * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  rc = encode_device_impl(ctx, base, 0, frames * 512, 1, 1, n_samples, 0, frames, ctx->d_params, fixed,
                          (uint8_t *)(base + o_su), 1, frames, base + o_bands, base + o_mags,
                          (uint8_t *)(base + o_modes), base + o_coefs);
  if (rc) return rc;
  if (bands) CU(ctx, cudaMemcpyAsync(bands, base + o_bands, frames * 512 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (mags) CU(ctx, cudaMemcpyAsync(mags, base + o_mags, frames * 256 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (coefs) CU(ctx, cudaMemcpyAsync(coefs, base + o_coefs, frames * 512 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (su) CU(ctx, cudaMemcpyAsync(su, base + o_su, frames * 212, cudaMemcpyDeviceToHost, ctx->stream));
  std::vector<uint8_t> m(frames * 4);
  CU(ctx, cudaMemcpyAsync(m.data(), base + o_modes, frames * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  if (modes)
    for (size_t f = 0; f < frames; f++)
      for (int b = 0; b < 3; b++)
        modes[f * 3 + b] = fixed ? opts->fixed_block_modes[b] : (int32_t)m[f * 4 + b];
  return CARTA1_OK;
}

int carta1_debug_decode_stages(carta1_ctx *ctx, const uint8_t *su, size_t n_su, float *coefs, float *bands,
                               float *pcm) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  if (n_su == 0) return CARTA1_OK;
  CU(ctx, cudaSetDevice(ctx->device));
  const size_t o_coefs = ((n_su * 212 + 3) / 4 + 4) & ~(size_t)3, o_bands = o_coefs + n_su * 512, o_pcm = o_bands + n_su * 512,
               total = o_pcm + n_su * 512;
  CU(ctx, ctx->dbg.ensure(total * sizeof(float)));
  float *base = (float *)ctx->dbg.p;
  CU(ctx, cudaMemcpyAsync(base, su, n_su * 212, cudaMemcpyHostToDevice, ctx->stream));
  int rc = decode_device_impl(ctx, (const uint8_t *)base, 1, n_su, n_su, 1, 0, n_su, base + o_pcm, 0, n_su * 512, 1,
                              base + o_coefs, base + o_bands);
  if (rc) return rc;
  if (coefs) CU(ctx, cudaMemcpyAsync(coefs, base + o_coefs, n_su * 512 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (bands) CU(ctx, cudaMemcpyAsync(bands, base + o_bands, n_su * 512 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  if (pcm) CU(ctx, cudaMemcpyAsync(pcm, base + o_pcm, n_su * 512 * 4, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}

int carta1_debug_transient_scores(carta1_ctx *ctx, const float *pcm, size_t n_samples, const carta1_enc_opts *opts,
                                  double *scores) {
  CtxLock ctx_lock(ctx);
  if (!ctx) return CARTA1_ERR_ARG;
  const size_t frames = carta1_frame_count(n_samples);
  if (frames == 0) return CARTA1_OK;
  if (!pcm || !scores) return fail(ctx, CARTA1_ERR_ARG, "carta1_debug_transient_scores: NULL argument");
  if (opts && opts->use_fixed_block_modes)
    return fail(ctx, CARTA1_ERR_ARG, "carta1_debug_transient_scores: fixed block modes compute no score");
  CU(ctx, cudaSetDevice(ctx->device));
  int rc = upload_params(ctx, opts, ctx->d_params);
  if (rc) return rc;
  // layout: pcm | scores (doubles, 8-byte aligned)
  const size_t o_scores = (frames * 512 + 1) & ~(size_t)1;
  CU(ctx, ctx->dbg.ensure(o_scores * sizeof(float) + frames * 3 * sizeof(double)));
  float *base = (float *)ctx->dbg.p;
  CU(ctx, cudaMemsetAsync(base, 0, o_scores * sizeof(float), ctx->stream));
  CU(ctx, cudaMemcpyAsync(base, pcm, n_samples * sizeof(float), cudaMemcpyHostToDevice, ctx->stream));
  rc = encode_device_impl(ctx, base, 0, frames * 512, 1, 1, n_samples, 0, frames, ctx->d_params, false, nullptr, 1, frames,
                          nullptr, nullptr, nullptr, nullptr, (double *)(base + o_scores));
  if (rc) return rc;
  CU(ctx, cudaMemcpyAsync(scores, base + o_scores, frames * 3 * sizeof(double), cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  return CARTA1_OK;
}

int carta1_ctx_near_threshold(carta1_ctx *ctx, uint64_t counts[3], int reset) {
  CtxLock ctx_lock(ctx);
  if (!ctx || !counts) return CARTA1_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  unsigned long long near[2] = {0, 0};
  CU(ctx, cudaMemcpyAsync(near, ctx->d_near, sizeof near, cudaMemcpyDeviceToHost, ctx->stream));
  if (reset) CU(ctx, cudaMemsetAsync(ctx->d_near, 0, sizeof near, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  counts[0] = ctx->near_decisions;
  counts[1] = near[0];
  counts[2] = near[1];
  if (reset) ctx->near_decisions = 0;
  return CARTA1_OK;
}

int carta1_debug_selftest(carta1_ctx *ctx, uint64_t *mismatches) {
  CtxLock ctx_lock(ctx);
  if (!ctx || !mismatches) return CARTA1_ERR_ARG;
  CU(ctx, cudaSetDevice(ctx->device));
  CU(ctx, ctx->dbg.ensure(64));
  CU(ctx, cudaMemsetAsync(ctx->dbg.p, 0, 8, ctx->stream));
  CU(ctx, launch_selftest(ctx->d_tables, (unsigned long long *)ctx->dbg.p, ctx->stream));
  unsigned long long bad = 0;
  CU(ctx, cudaMemcpyAsync(&bad, ctx->dbg.p, 8, cudaMemcpyDeviceToHost, ctx->stream));
  CU(ctx, cudaStreamSynchronize(ctx->stream));
  *mismatches = bad;
  return CARTA1_OK;
}

// ------------------------------------------------------------------ AEA container
int carta1_aea_write_header(const char *title, uint32_t su_count, int n_ch, uint8_t out[CARTA1_AEA_HEADER_BYTES]) {
  // serialization.js:190-211
  if (!out) return CARTA1_ERR_ARG;
  memset(out, 0, CARTA1_AEA_HEADER_BYTES);
  out[1] = 0x08;
  if (title) {
    size_t n = strlen(title);
    if (n > 255) n = 255;
    memcpy(out + 4, title, n);
  }
  out[260] = (uint8_t)su_count;
  out[261] = (uint8_t)(su_count >> 8);
  out[262] = (uint8_t)(su_count >> 16);
  out[263] = (uint8_t)(su_count >> 24);
  out[264] = (uint8_t)n_ch;
  return CARTA1_OK;
}

int carta1_aea_parse_header(const uint8_t *hdr, size_t len, char title_out[257], uint32_t *su_count, int *n_ch) {
  // serialization.js:222-253
  if (!hdr || len != CARTA1_AEA_HEADER_BYTES) return fail(nullptr, CARTA1_ERR_ARG, "Header must be 2048 bytes");
  if (hdr[0] != 0 || hdr[1] != 8 || hdr[2] != 0 || hdr[3] != 0) return fail(nullptr, CARTA1_ERR_ARG, "Invalid AEA file");
  size_t end = 4;
  while (end < CARTA1_AEA_HEADER_BYTES && hdr[end] != 0) end++;
  size_t tl = end == CARTA1_AEA_HEADER_BYTES ? 256 : end - 4;
  if (tl > 256) tl = 256;
  if (title_out) { memcpy(title_out, hdr + 4, tl); title_out[tl] = 0; }
  if (su_count) *su_count = (uint32_t)hdr[260] | ((uint32_t)hdr[261] << 8) | ((uint32_t)hdr[262] << 16) | ((uint32_t)hdr[263] << 24);
  if (n_ch) *n_ch = hdr[264];
  return CARTA1_OK;
}

}  // extern "C"
