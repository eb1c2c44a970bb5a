/*
 * carta1_oracle.h -- CPU ORACLE for the ATRAC1 encode/decode hot path of aynik/carta1.
 *
 * THIS IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.  It is a plain-C restatement of the
 * reference's JavaScript algorithm (citations are relative to /root/reference/).
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
 * legs may build, load or call it.  The product path (carta1_b200/) never does.
 *
 * PARITY STATUS: PINNED against outputs of the reference itself, run in the build image.
 * The reference ships no golden vectors and the image has no Node, but
 * Nsight Compute ships Qt 6.6.3, whose QJSEngine is a complete ECMAScript engine:
 * tools/ref_run_qjs.py lets it import the reference's own modules from /root/reference
 * (every file that holds codec arithmetic byte for byte as shipped) and records, under
 * tests/golden/ref/, the AEA bytes and decoded PCM of encodeAeaPcm / decodeAeaPcm for the
 * golden inputs, what the reference's stage closures hand to each other, known answers of
 * each function alone, the parity suite's 122 inputs and seconds-long runs.  This oracle
 * equals every one of them bit for bit (tests/test_reference_pin.py), and so does the CUDA
 * path.  What the engine does not pin is V8's libm: Qt's engine calls the host's glibc, V8
 * carries fdlibm ports.  libm enters in two places only: the sin/cos/pow tables, which are
 * inputs (c1o_tables; the dump's are used for the comparison), and the transient score's
 * log/exp/log10/log1p, where this file carries the fdlibm port by default and the host's
 * libm behind c1o_set_host_libm(1) (bit-identical scores with the dump's engine).
 * It is also pinned against every exact known-answer value the reference's tests hold
 * (tests/bitstream.test.js:6-71, tests/mdct.test.js:22-33) and against all structural /
 * tolerance assertions of the reference test-suite (tests/test_oracle_reference_suite.py).
 *
 * Numerical contract: every typed-array store is a round to
 * binary32, every expression between stores is IEEE binary64 evaluated one operator at a
 * time in JS source order.  Build with -ffp-contract=off; no FMA is ever formed.
 */
#ifndef CARTA1_ORACLE_H
#define CARTA1_ORACLE_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define C1O_FRAME 512
#define C1O_SU_BYTES 212
#define C1O_NUM_BFU 52
#define C1O_AEA_HEADER 2048

/* Every libm-derived table of the reference.  Injectable so that a JS host can supply
 * the values its own Math.* produces. */
typedef struct c1o_tables {
  double window_short[32];   /* codec/core/constants.js:60-66 */
  double scale_factors[64];  /* codec/core/constants.js:144-150 */
  double mdct_fwd64[32];     /* codec/transforms/mdct.js:27-36, instances :215-217 */
  double mdct_fwd256[128];
  double mdct_fwd512[256];
  double mdct_inv64[32];     /* codec/transforms/mdct.js:219-221 */
  double mdct_inv256[128];
  double mdct_inv512[256];
  double fft_w[8][2];        /* codec/transforms/fft.js:37-39; stride 2<<k: cos, sin of -2pi/stride */
} c1o_tables;

/* codec/core/options.js:16-23 restricted to what the hot path reads
 * (codec/pipeline/encoder.js:131-141, codec/coding/bitallocation.js:46-61). */
typedef struct c1o_options {
  double transient_threshold; /* transientThresholdLow: used for ALL bands */
  double allocation_bias;
  int use_fixed_modes;        /* fixedBlockModes != null */
  int fixed_modes[3];
  double biased_sf[64];       /* pow(SCALE_FACTORS[i], bias), identity when bias == 1 */
} c1o_options;

/* The frame object exchanged by encode()/decode() closures
 * (codec/pipeline/encoder.js:410-416). q[b][j] = quantizedCoefficients[b][j]. */
typedef struct c1o_frame {
  int n_bfu;
  int modes[3];
  int sfi[C1O_NUM_BFU];
  int wl[C1O_NUM_BFU];
  int q[C1O_NUM_BFU][20];
} c1o_frame;

/* Persistent encoder state == the parts of BufferPool that survive a frame
 * (codec/core/buffers.js:31-42,60-65). */
typedef struct c1o_encoder {
  const c1o_tables *T;
  c1o_options opt;
  float delay_low[46], delay_mid[46], delay_high[39];
  float overlap[3][32];
  float prev_mag[3][128];
} c1o_encoder;

/* Persistent decoder state (codec/core/buffers.js:31-35,67-72): QMF delays and the
 * last 16 entries of each imdctOverlap buffer (the only part read back,
 * codec/pipeline/decoder.js:203-206,255-258). */
typedef struct c1o_decoder {
  const c1o_tables *T;
  float delay_low[46], delay_mid[46], delay_high[39];
  float tail[3][16];
} c1o_decoder;

/* Optional capture of intermediates for stage-level parity tests. */
typedef struct c1o_enc_debug {
  float bands[512];   /* low128 | mid128 | high256, before tail windowing */
  float mags[256];    /* transient magnitudes low64 | mid64 | high128 (auto mode only) */
  double score[3];    /* transient score per band (auto mode only) */
  float coefs[512];   /* MDCT coefficients */
} c1o_enc_debug;

typedef struct c1o_dec_debug {
  float coefs[512];
  float bands[512];   /* after IMDCT + overlap-add */
} c1o_dec_debug;

/* ---- tables / options ---- */
void c1o_default_tables(c1o_tables *t);          /* glibc libm */
void c1o_fdlibm_tables(c1o_tables *t);           /* the same with fdlibm's sin / cos / pow (V8 <= 11.3): fdlibm_trig_pow.c */
double c1o_fd_sin(double x);
double c1o_fd_cos(double x);
double c1o_fd_pow(double x, double y);
int c1o_fd_selfcheck(void);                      /* 0 = every fdlibm constant's decimal and bit pattern agree */
const float *c1o_qmf_even(void);                 /* 24 taps */
const float *c1o_qmf_odd(void);
const int *c1o_specs_per_bfu(void);
const int *c1o_bfu_start_long(void);
const int *c1o_bfu_start_short(void);
void c1o_options_init(c1o_options *o, const c1o_tables *t, double threshold,
                      double bias, const int *fixed_modes /* NULL = auto */);

/* ---- unit functions (each mirrors one reference function) ---- */
void c1o_qmf_analysis(const float *in, int n, float *delay46, float *lo, float *hi);
void c1o_qmf_synthesis(const float *lo, const float *hi, int n_sub, float *delay46, float *out);
void c1o_fft(float *re, float *im, int n, const c1o_tables *t);
void c1o_mdct(const c1o_tables *t, int size, const float *in, float *out);   /* size 64|256|512 */
void c1o_imdct(const c1o_tables *t, int size, const float *in, float *out);
void c1o_overlap_add(const float *prev, const float *curr, int size, const double *window, float *out);
void c1o_perform_fft(const float *samples, int n_samples, int fft_size, const c1o_tables *t, float *mag);
double c1o_transient_score(const float *cur, const float *prev, int n);
int c1o_find_scale_factor(const float *coefs, int n);
int c1o_find_scale_factor_table(float max_abs);  /* exact threshold-table variant */
const float *c1o_sf_thresholds(void);            /* 63 f32 thresholds */
void c1o_allocate_bits(const float *coefs, const int *modes, const c1o_options *o,
                       int *n_bfu, int *sfi52, int *wl52);
const double *c1o_debug_last_totals(void); /* per-candidate totals of the last c1o_allocate_bits */
void c1o_quantize(const float *c, int n, int sfi, int bits, const c1o_tables *t, int *out);
void c1o_dequantize(const int *q, int n, int sfi, int bits, const c1o_tables *t, float *out);
void c1o_pack_bits(uint8_t *buf, size_t buf_len, int bit_pos, int value, int bit_count);
int c1o_unpack_bits(const uint8_t *buf, size_t buf_len, int bit_pos, int bit_count);
int c1o_unpack_signed_bits(const uint8_t *buf, size_t buf_len, int bit_pos, int bit_count);
void c1o_serialize_frame(const c1o_frame *f, uint8_t out[C1O_SU_BYTES]);
void c1o_deserialize_frame(const uint8_t in[C1O_SU_BYTES], c1o_frame *f);
void c1o_set_host_libm(int on); /* transient detector: 1 = host libm (the dump's engine), 0 = fdlibm port (V8), default */
double c1o_log(double x);
double c1o_exp(double x);
double c1o_log10(double x);
double c1o_log1p(double x);
int32_t c1o_to_int32(double x);

/* ---- frame closures ---- */
void c1o_encoder_init(c1o_encoder *e, const c1o_tables *t, const c1o_options *o);
void c1o_decoder_init(c1o_decoder *d, const c1o_tables *t);
void c1o_encode_frame(c1o_encoder *e, const float pcm[C1O_FRAME], c1o_frame *out, c1o_enc_debug *dbg);
void c1o_decode_frame(c1o_decoder *d, const c1o_frame *in, float pcm[C1O_FRAME], c1o_dec_debug *dbg);

/* ---- whole-buffer helpers (codec/io/processor.js:246-279,317-339,597-654) ---- */
/* Frames needed for n_samples (zero-padded last frame). */
size_t c1o_frame_count(size_t n_samples);
/* Encode planar channels -> interleaved sound units (L,R,L,R...). su_out must hold
 * frame_count*n_ch*212 bytes.  Frames [frame_begin, frame_end) only; the encoder is
 * cold-started `warmup` frames earlier (2 is exact). */
void c1o_encode_pcm_range(const c1o_tables *t, const c1o_options *o, const float *const *ch,
                          int n_ch, size_t n_samples, size_t frame_begin, size_t frame_end,
                          uint8_t *su_out);
/* Decode interleaved sound units -> planar PCM; frames [frame_begin, frame_end). */
void c1o_decode_su_range(const c1o_tables *t, const uint8_t *su, size_t n_su, int n_ch,
                         size_t frame_begin, size_t frame_end, float *const *ch_out);
void c1o_aea_header(const char *title, uint32_t su_count, int n_ch, uint8_t out[C1O_AEA_HEADER]);
int c1o_aea_parse(const uint8_t *hdr, size_t len, char title[257], uint32_t *su_count, int *n_ch);
/* codec/io/processor.js:382-389 (float -> int16) and bin/cli.js:394-404 (int16 -> float). */
void c1o_pcm_to_int16(const float *in, size_t n, int16_t *out);
void c1o_int16_to_pcm(const int16_t *in, size_t n, float *out);

#ifdef __cplusplus
}
#endif
#endif
