// c1_launch.h -- host-side launch descriptors shared by the ABI layer and the kernel files.
#pragma once

#include <cuda_runtime.h>
#include <stddef.h>
#include <stdint.h>

#include <utility>
#include <vector>

namespace c1 {

struct DevTables;
struct DevEncParams;

// A second stream and two events: the two role kernels of a transform (low + mid bands / high band) are independent
// of each other, so a launch that leaves most of the machine idle (the stateful handles' few frames per call) runs
// them side by side: fork from the launching stream, join back (also inside a stream capture, where this becomes a
// fork in the graph).  Large launches fill the machine with one role and stay on one stream.
struct ForkJoin {
  cudaStream_t aux = nullptr;
  cudaEvent_t fork = nullptr, join = nullptr;
};
bool fork_roles(const ForkJoin *fj, int grid, int resident);  // true: role 1 goes to fj->aux
inline cudaError_t fork_begin(const ForkJoin *fj, cudaStream_t st) {
  cudaError_t e = cudaEventRecord(fj->fork, st);
  return e != cudaSuccess ? e : cudaStreamWaitEvent(fj->aux, fj->fork, 0);
}
inline cudaError_t fork_end(const ForkJoin *fj, cudaStream_t st) {
  cudaError_t e = cudaEventRecord(fj->join, fj->aux);
  return e != cudaSuccess ? e : cudaStreamWaitEvent(st, fj->join, 0);
}

// One encode pass over `n_streams` rows of (halo_frames + n_out_frames) frames each.
struct EncodeLaunch {
  const void *pcm;          // f32 planar rows (pcm_fmt 0) or s16 interleaved (pcm_fmt 1)
  int pcm_fmt;
  size_t row_stride;        // floats between rows (pcm_fmt 0)
  int n_ch_interleave;      // channel count of the interleaved s16 source (pcm_fmt 1)
  long long valid_samples;  // samples present per row; the rest reads as zero
  int n_streams;
  int frames_total;         // halo_frames + n_out_frames
  int halo_frames;
  int n_out_frames;
  int use_fixed;
  const DevTables *tables;
  const DevEncParams *params;
  // scratch, [n_streams][frames_total][...]
  float *bands;             // 512 per unit
  float *mags;              // 256 per unit (auto modes only)
  void *feats;              // 3 x {flatness, hf ratio, energy} doubles per unit (auto modes only)
  uint8_t *modes;           // 4 per unit   (auto modes only)
  double *scores;           // 3 per unit, optional
  unsigned long long *near_counts;  // optional: [0] decisions with |score - threshold| < 1e-9, [1] < 1e-12 (emitted frames)
  float *coefs;             // 512 per unit
  uint8_t *sfi;             // 64 per unit: scale-factor index per BFU (52 used)
  void *alloc_recs;         // alloc_rec_bytes() per emitted unit
  // output
  uint8_t *su_out;          // may be NULL (stage taps only)
  size_t su_frame_stride, su_stream_stride;
  const ForkJoin *fj;       // optional
};

struct DecodeLaunch {
  const uint8_t *su;        // unit (f, s) at su + (f*su_frame_stride + s*su_stream_stride)*212
  size_t su_frame_stride, su_stream_stride;
  long long n_su_valid;     // linear unit indices >= this decode as the dummy frame
  // alternative input: position-expanded frame objects, [n_streams][n_out_frames][512] (su unused)
  const int32_t *x_q;
  const uint8_t *x_sfi, *x_bits, *x_modes;
  // stateful handles: [n_streams][512] band records of the unit before the call; when set,
  // frame 0 of every row is that record and the su / expanded arrays start at frame 1
  const float *prev_rec;
  // stateful handles: when set, K7 also stores the band record of every row's LAST unit here ([n_streams][512]), the
  // prev_rec of the next call (K5 has read this call's prev_rec long before)
  float *save_rec;
  int n_streams;
  int frames_total;         // halo_frames + n_out_frames
  int halo_frames;
  int n_out_frames;
  const DevTables *tables;
  // scratch, [n_streams][frames_total][...]
  float *coefs;             // 512 per unit: dequantised coefficients
  uint8_t *modes;           // 4 per unit
  float *inv;               // 512 per unit: IMDCT output before overlap-add
  float *bands_dbg;         // optional 512 per unit: time-domain bands (stage tap)
  // output
  void *pcm;                // f32 planar rows (pcm_fmt 0) or s16 interleaved (pcm_fmt 1); may be NULL
  int pcm_fmt;
  size_t row_stride;
  int n_ch_interleave;
  const ForkJoin *fj;       // optional
};

// Launch accounting and optional per-kernel CUDA-event timing (bench.py's roofline leg).
enum KernelId {
  K_QMF_ANALYSIS = 0, K_BAND_MAGS, K_TRANSIENT_MODES, K_MDCT, K_ALLOC, K_QUANT_PACK,
  K_UNPACK, K_IMDCT, K_BANDS_TIME, K_SYNTH, K_COUNT
};
const char *kernel_name(int id);

struct Prof {
  bool on = false;
  uint64_t launches = 0;
  std::vector<std::pair<cudaEvent_t, cudaEvent_t>> ev[K_COUNT];
  void begin(int id, cudaStream_t st) {
    launches++;
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, st);
    ev[id].push_back({a, b});
  }
  void end(int id, cudaStream_t st) {
    if (on) cudaEventRecord(ev[id].back().second, st);
  }
};

size_t alloc_rec_bytes();
int pick_run_len(int frames, int n_streams, int warps);  // frames per run of the streaming QMF kernels
int persistent_ctas(int per_sm);
int resident_ctas(const void *kernel, int threads, size_t dyn_smem);  // SM count x CTAs of `kernel` resident per SM  // SM count of the current device x per_sm
// QMF taps and FFT twiddles go to __constant__ memory of the current device (once per context).
cudaError_t upload_encode_constants(const DevTables *host_tables);
cudaError_t upload_decode_constants(const DevTables *host_tables);
cudaError_t launch_encode(const EncodeLaunch &L, cudaStream_t st, Prof *prof);
cudaError_t launch_decode(const DecodeLaunch &L, cudaStream_t st, Prof *prof);
// batched deserializeFrame: [n] BFU counts, [n][3] block modes, [n][52] indices, [n][512] integers in bitstream order
cudaError_t launch_deserialize(const uint8_t *d_su, int n_units, const DevTables *tables, uint8_t *n_bfu, int8_t *modes,
                               uint8_t *wl, uint8_t *sfi, int32_t *q, cudaStream_t st, Prof *prof);
cudaError_t launch_selftest(const DevTables *tables, unsigned long long *d_bad, cudaStream_t st);

}  // namespace c1
