/*
 * carta1_b200.h -- C ABI of the B200-native ATRAC1 encode/decode hot path.
 *
 * This is the drop-in boundary for aynik/carta1's hot path.  The
 * reference has no FFI: its boundary is the ES-module surface of codec/index.js:26-47.
 * Every entry point below names the reference interface it replaces (paths relative to
 * the reference checkout).  The N-API shim in carta1_b200/napi/ binds exactly these
 * symbols; tests drive them through ctypes.
 *
 * Conventions
 *   - plain pointers and sizes only; all functions return 0 on success, non-zero on error
 *     (carta1_last_error() gives the message; messages mirror the reference's throws).
 *   - "sound unit" (SU) = 212 bytes = one mono frame of 512 samples (codec/core/constants.js:7,19).
 *   - calls are blocking and thread-safe.  A context owns one stream, its scratch buffers and the
 *     state of the handles created from it: calls on one context, or on encoders / decoders created
 *     from it, are serialised by the library (any thread may make them).  Concurrency is one context
 *     per thread: an encode call and a decode call on two contexts overlap on the PCIe link.
 *   - there is NO CPU fallback: every compute entry point fails if no sm_100 device /
 *     kernel image is available.
 */
#ifndef CARTA1_B200_H
#define CARTA1_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define CARTA1_FRAME_SAMPLES 512
#define CARTA1_SU_BYTES 212
#define CARTA1_AEA_HEADER_BYTES 2048

#define CARTA1_OK 0
#define CARTA1_ERR_ARG 1      /* bad argument (message mirrors the reference's throw) */
#define CARTA1_ERR_CUDA 2     /* CUDA runtime failure or no usable device */
#define CARTA1_ERR_ALLOC 3

typedef struct carta1_ctx carta1_ctx;
typedef struct carta1_encoder carta1_encoder;
typedef struct carta1_decoder carta1_decoder;

/* Every libm-derived table of the reference, so that a JS host can upload the values its
 * own Math.* produces.  NULL anywhere a table pointer is accepted
 * means "host libm defaults" (carta1_default_tables). */
typedef struct carta1_tables {
  double window_short[32];   /* WINDOW_SHORT, codec/core/constants.js:60-66 */
  double scale_factors[64];  /* SCALE_FACTORS, codec/core/constants.js:144-150 */
  double mdct_fwd64[32];     /* MDCTBase.sinCosTable of mdct64/256/512, codec/transforms/mdct.js:27-36,215-217 */
  double mdct_fwd256[128];
  double mdct_fwd512[256];
  double mdct_inv64[32];     /* ... of imdct64/256/512, codec/transforms/mdct.js:219-221 */
  double mdct_inv256[128];
  double mdct_inv512[256];
  double fft_w[8][2];        /* (cos, sin)(-2*pi/stride), stride = 2<<k, codec/transforms/fft.js:37-39 */
} carta1_tables;

/* EncoderOptions as read by the hot path (codec/core/options.js:16-23;
 * codec/pipeline/encoder.js:131-141 reads transientThresholdLow for all three bands). */
typedef struct carta1_enc_opts {
  double transient_threshold_low; /* EncoderOptions.transientThresholdLow */
  double allocation_bias;         /* EncoderOptions.allocationBias */
  int32_t use_fixed_block_modes;  /* EncoderOptions.fixedBlockModes != null */
  int32_t fixed_block_modes[3];
  /* pow(SCALE_FACTORS[i], allocationBias) as the host computes it
   * (codec/coding/bitallocation.js:46-61); NULL = computed natively with libm pow. */
  const double *biased_scale_factors;
} carta1_enc_opts;

/* ---- library / context ---------------------------------------------------------- */
int carta1_abi_version(void);
void carta1_default_tables(carta1_tables *out);
void carta1_default_enc_opts(carta1_enc_opts *out); /* defaults of codec/core/options.js:17-23 */
/* One context per GPU.  Uploads the tables, creates the stream. */
int carta1_ctx_create(int device, const carta1_tables *tables, carta1_ctx **out);
void carta1_ctx_destroy(carta1_ctx *ctx);
const char *carta1_last_error(const carta1_ctx *ctx); /* ctx may be NULL: last create error */
int carta1_device_count(void);

/* ---- whole-buffer entry points (host memory) -------------------------------------
 * Replace encodeAeaPcm / decodeAeaPcm bodies (codec/io/processor.js:597-654) and the
 * AudioProcessor.encodeStream / decodeStream loops (:69-237) when the caller has the
 * whole signal.  */
size_t carta1_frame_count(size_t n_samples); /* frameBufferToFrames, processor.js:246-279 */
/* The host entry points below stage at most `units` sound units (frames x channels) per pass
 * through device memory, re-reading a 2-frame PCM halo (decode: 1 unit) at every pass boundary
 * (SURVEY.md Appendix B); passes rotate through four staging slots (copies overlap compute).
 * 0 restores the default of 2^16.  Results never depend on it. */
int carta1_ctx_set_max_units_per_pass(carta1_ctx *ctx, size_t units);

/* Caller buffers may be pinned or pageable.  Pinned buffers (carta1_host_alloc, cudaHostAlloc,
 * cudaHostRegister) are read and written in place by the copy engines and kernels: 1 h of stereo
 * PCM encodes in 24 ms.  Pageable buffers of 8 MiB and more are staged through pinned bounce slots
 * owned by the context, filled and drained by a multi-threaded memcpy one pass behind: 44 ms.  An
 * encode call and a decode call on two contexts (two host threads) overlap on the PCIe link.
 * carta1_host_alloc returns page-locked host memory for a host runtime to build its typed arrays
 * on (Node: napi_create_external_arraybuffer); it has no counterpart in the reference, whose
 * Float32Array / Uint8Array results are plain allocations (processor.js:641-654). */
int carta1_host_alloc(size_t bytes, void **out);
void carta1_host_free(void *p);

/* channels[c] points to n_samples f32 samples (planar; caller zero-pads the shorter stereo
 * channel as frameBufferToFrames does).  Writes frame_count*n_ch sound units interleaved
 * L,R,L,R (createAeaBlob order, processor.js:317-339).  */
int carta1_encode_pcm(carta1_ctx *ctx, const float *const *channels, int n_ch, size_t n_samples,
                      const carta1_enc_opts *opts, uint8_t *su_out, size_t su_capacity_bytes,
                      size_t *n_su_out);
/* su: n_su interleaved units.  channels_out[c] receives ceil(n_su/n_ch)*512 samples.  A
 * missing last right unit decodes as the dummy frame of processor.js:299-307. */
int carta1_decode_su(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch,
                     float *const *channels_out);
/* Same, fused with the WAV int16 conversions:
 * int16 interleaved in -> /32768.0 (bin/cli.js:395); out -> clamp, x32767/x32768, truncate
 * (codec/io/processor.js:382-389), interleaved. */
int carta1_encode_pcm_s16(carta1_ctx *ctx, const int16_t *interleaved, int n_ch, size_t n_samples,
                          const carta1_enc_opts *opts, uint8_t *su_out, size_t su_capacity_bytes,
                          size_t *n_su_out);
int carta1_decode_su_s16(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch,
                         int16_t *interleaved_out);

/* One shard of a longer stream through host memory (BASELINE configs[4]: a rank's contiguous frame range of
 * a stream; the loop being cut is codec/io/processor.js:97-136).  channels[c] holds
 * n_samples samples of which the first halo_frames * 512 are history that is read but not emitted:
 * halo_frames is 0 (the shard starts the stream) or >= 2.  Writes (frame_count(n_samples) - halo_frames)
 * * n_ch units; they equal the corresponding units of the whole-stream call bit for bit. */
int carta1_encode_pcm_shard(carta1_ctx *ctx, const float *const *channels, int n_ch, size_t n_samples,
                            size_t halo_frames, const carta1_enc_opts *opts, uint8_t *su_out,
                            size_t su_capacity_bytes, size_t *n_su_out);
/* su holds n_su interleaved units of which the first halo_frames * n_ch are history (0: the shard starts
 * the file, else >= 1; the loop being cut is processor.js:159-237).  channels_out[c] receives
 * (ceil(n_su / n_ch) - halo_frames) * 512 samples, equal to that span of the whole-file call. */
int carta1_decode_su_shard(carta1_ctx *ctx, const uint8_t *su, size_t n_su, int n_ch, size_t halo_frames,
                           float *const *channels_out);

/* ---- stateful, batched frame closures --------------------------------------------
 * n_streams independent encode() closures (codec/pipeline/encoder.js:438-450) advanced
 * together: pcm is [n_streams][n_frames][512] f32, su_out is [n_streams][n_frames][212].
 * n_streams == 1, n_frames == 1 is the reference closure call. */
int carta1_enc_create(carta1_ctx *ctx, const carta1_enc_opts *opts, int n_streams,
                      carta1_encoder **out);
void carta1_enc_destroy(carta1_encoder *enc);
int carta1_enc_reset(carta1_encoder *enc); /* new BufferPool() */
int carta1_enc_frames(carta1_encoder *enc, const float *pcm, int n_frames, uint8_t *su_out);
/* n_streams independent decode() closures (codec/pipeline/decoder.js:408-411). */
int carta1_dec_create(carta1_ctx *ctx, int n_streams, carta1_decoder **out);
void carta1_dec_destroy(carta1_decoder *dec);
int carta1_dec_reset(carta1_decoder *dec);
int carta1_dec_frames(carta1_decoder *dec, const uint8_t *su, int n_frames, float *pcm_out);
/* The same closures fed with frame OBJECTS that the 212-byte layout cannot hold: decode()
 * accepts any object with the frame keys (tests/decoder.test.js:70-98: nBfu 0, arbitrary
 * quantizedCoefficients lengths).  The host replays dequantizationStage's
 * `coefficients.set(dequantized, position)` sequence (codec/pipeline/decoder.js:73-94) into
 * per-position arrays [n_streams][n_frames][512]: q (quantised value), sfi (scale-factor index,
 * 0..63), bits (WORD_LENGTH_BITS value 0 or 2..16; 0 = position never written); modes is
 * [n_streams][n_frames][3] (any non-zero = short blocks, decoder.js:82-83). */
int carta1_dec_frames_expanded(carta1_decoder *dec, const int32_t *q, const uint8_t *sfi,
                               const uint8_t *bits, const int32_t *modes, int n_frames,
                               float *pcm_out);

/* ---- device-resident entry points (bench `value`, multi-GPU shards) ---------------
 * All pointers are device pointers on the context's GPU; work is enqueued on the
 * context's stream and the call returns after enqueueing unless sync != 0.
 *
 * d_pcm: [n_streams] rows of `row_stride` floats; each row holds (halo_frames + n_frames)
 * frames.  halo_frames is 0 (row starts at the stream start: history is silence) or >= 2
 * (the first halo_frames frames are real history and are not emitted) -- SURVEY.md
 * Appendix B.  valid_samples = samples present per row (the rest of the last frame is
 * treated as zero).  Unit s of frame f is written at d_su + (f*su_frame_stride +
 * s*su_stream_stride)*212.  */
int carta1_encode_device(carta1_ctx *ctx, const float *d_pcm, size_t row_stride, int n_streams,
                         size_t valid_samples, size_t halo_frames, size_t n_frames,
                         const carta1_enc_opts *opts, uint8_t *d_su, size_t su_frame_stride,
                         size_t su_stream_stride, int sync);
/* d_su holds (halo_frames + n_frames) frames of units addressed as above; halo_frames is 0
 * (stream start) or >= 1.  n_su_valid: units with linear index >= n_su_valid decode as the
 * dummy frame.  PCM of frame f, stream s goes to d_pcm + s*row_stride + f*512. */
int carta1_decode_device(carta1_ctx *ctx, const uint8_t *d_su, size_t su_frame_stride,
                         size_t su_stream_stride, size_t n_su_valid, int n_streams,
                         size_t halo_frames, size_t n_frames, float *d_pcm, size_t row_stride,
                         int sync);
int carta1_ctx_sync(carta1_ctx *ctx);
void *carta1_ctx_stream(carta1_ctx *ctx); /* cudaStream_t the work is enqueued on */
/* Kernels launched by this context since creation (bench.py's gpu_launches). */
uint64_t carta1_ctx_launch_count(const carta1_ctx *ctx);

/* Per-kernel CUDA-event timing (bench.py's roofline leg).  While enabled every kernel launch
 * is bracketed by events on the context's stream; profile_read synchronises, returns the
 * summed milliseconds and launch counts per kernel id (0 .. carta1_kernel_count()-1) and
 * clears the records. */
int carta1_kernel_count(void);
const char *carta1_kernel_name(int id);
int carta1_ctx_profile(carta1_ctx *ctx, int enable);
int carta1_ctx_profile_read(carta1_ctx *ctx, double *ms_out, uint64_t *count_out, int n);

/* ---- stage-level taps for parity tests (host memory, small inputs) ---------------
 * Run the encode path on one row of PCM and return the intermediates the oracle also
 * exposes: bands [n_frames][512], transient magnitudes [n_frames][256], block modes
 * [n_frames][3], MDCT coefficients [n_frames][512].  Any output pointer may be NULL. */
int carta1_debug_encode_stages(carta1_ctx *ctx, const float *pcm, size_t n_samples,
                               const carta1_enc_opts *opts, float *bands, float *mags,
                               int32_t *modes, float *coefs, uint8_t *su);
/* Decode taps: dequantised coefficients [n][512], time-domain bands after IMDCT+overlap
 * [n][512], PCM [n][512]. */
int carta1_debug_decode_stages(carta1_ctx *ctx, const uint8_t *su, size_t n_su, float *coefs,
                               float *bands, float *pcm);

/* Transient scores [n_frames][3] (low, mid, high) of one row of PCM with auto block modes: the value
 * detectTransient compares with the threshold (codec/analysis/transient.js:44-55,197-226). */
int carta1_debug_transient_scores(carta1_ctx *ctx, const float *pcm, size_t n_samples,
                                  const carta1_enc_opts *opts, double *scores);

/* Close calls of the block-mode decision `score > threshold` (codec/analysis/transient.js:54), the one
 * place where a libm that differs from V8's by an ulp could change emitted bytes:
 * counts[0] = decisions taken (3 per emitted sound unit of every auto-block-mode call), counts[1] = those
 * with |score - threshold| < 1e-9, counts[2] = those with |score - threshold| < 1e-12, since the context
 * was created or last reset.  Synchronises the context's stream.  No counterpart in the reference. */
int carta1_ctx_near_threshold(carta1_ctx *ctx, uint64_t counts[3], int reset);

/* Device self-test of the kernels' exact arithmetic shortcuts (reciprocal-based division,
 * in-FP64 rounding to binary32) against the IEEE operations they replace; *mismatches must
 * come back 0. */
int carta1_debug_selftest(carta1_ctx *ctx, uint64_t *mismatches);

/* ---- AEA container (codec/io/serialization.js:190-253) --------------------------- */
int carta1_aea_write_header(const char *title_utf8, uint32_t su_count, int n_ch,
                            uint8_t out[CARTA1_AEA_HEADER_BYTES]);
/* returns CARTA1_ERR_ARG with "Header must be 2048 bytes" / "Invalid AEA file". */
int carta1_aea_parse_header(const uint8_t *hdr, size_t len, char title_out[257],
                            uint32_t *su_count, int *n_ch);

/* ---- frame dump -------------------------------------------------------------------
 * deserializeFrame (codec/io/serialization.js:111-176) over n_su sound units at once: what the
 * `--json` dump of bin/cli.js:567-677 runs over a whole file.  Per unit i:
 *   n_bfu[i]            nBfu (BFU_AMOUNTS[(header >> 5) & 7])
 *   block_modes[3 i..]  blockModes as stored: 2 - field, 2 - field, 3 - field (may be negative)
 *   wl[52 i + b], sfi[52 i + b]   wordLengthIndices / scaleFactorIndices of BFU b (0 for b >= nBfu)
 *   q[512 i + p]        quantizedCoefficients in bitstream order: BFU b's SPECS_PER_BFU[b] integers
 *                       at p = BFU_START_LONG[b] + j (the running sum of the sizes); 0 where the
 *                       BFU carries no bits or b >= nBfu
 * Units that claim more bits than they have follow unpackBits past the end (bitstream.js:55-68). */
int carta1_deserialize_units(carta1_ctx *ctx, const uint8_t *su, size_t n_su, uint8_t *n_bfu,
                             int8_t *block_modes, uint8_t *wl, uint8_t *sfi, int32_t *q);

#ifdef __cplusplus
}
#endif
#endif
