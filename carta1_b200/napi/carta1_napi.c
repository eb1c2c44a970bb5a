/*
 * carta1_napi.c -- thin Node-API addon over the C ABI of include/carta1_b200.h.
 *
 * This is the binding a carta1 maintainer adds so that the JavaScript surface
 * (codec/index.js:26-47) runs its hot path on the GPU: index.mjs (next to this file) re-exports
 * the reference's names and routes encode()/decode(), encodeAeaPcm/decodeAeaPcm and the
 * AudioProcessor streams through the functions below.  No arithmetic happens here or in the
 * JS wrapper: typed arrays in, typed arrays out.
 *
 * Node is not present in the build image.  __graft_entry__.build() compiles this file, unchanged, against
 * node_api_min.h (hand-declared Node-API subset) and links it with tests/napi_host/napi_host.c, a minimal
 * Node-API host; tests/test_napi_host.py then calls every export the way index.mjs does (on a B200 against
 * the reference's own bytes).  It has never been loaded into Node itself; binding.gyp builds it against
 * Node's own headers.
 *
 * Exports (all synchronous unless noted):
 *   createContext(device, tables|null)                        -> ctx
 *   createEncoder(ctx, opts, nStreams) / encodeFrames(enc, Float32Array, nFrames) -> Uint8Array
 *   createDecoder(ctx, nStreams) / decodeFrames(dec, Uint8Array, nFrames)         -> Float32Array
 *   decodeFramesExpanded(dec, Int32Array q, Uint8Array sfi, Uint8Array bits, Int32Array modes, nFrames)
 *   encodePcm(ctx, Float32Array[], opts)   -> Promise<Uint8Array>     (napi_async_work)
 *   decodeSu(ctx, Uint8Array, nChannels)   -> Promise<Float32Array[]> (napi_async_work)
 *   destroy(handle)
 */
#ifdef CARTA1_NAPI_USE_NODE_HEADERS
#include <node_api.h>
#else
#include "node_api_min.h"
#endif
#include <stdlib.h>
#include <string.h>

#include "../../include/carta1_b200.h"

#define NAPI_OK(call) do { if ((call) != napi_ok) { napi_throw_error(env, NULL, "carta1_b200: N-API call failed: " #call); return NULL; } } while (0)

typedef enum { H_CTX, H_ENC, H_DEC } handle_kind;
typedef struct {
  handle_kind kind;
  void *ptr;
  carta1_ctx *ctx;  /* owning context, for error text */
} handle;

static void handle_release(handle *h) {
  if (!h || !h->ptr) return;
  if (h->kind == H_CTX) carta1_ctx_destroy((carta1_ctx *)h->ptr);
  if (h->kind == H_ENC) carta1_enc_destroy((carta1_encoder *)h->ptr);
  if (h->kind == H_DEC) carta1_dec_destroy((carta1_decoder *)h->ptr);
  h->ptr = NULL;
}
static void handle_finalize(napi_env env, void *data, void *hint) {
  (void)env; (void)hint;
  handle_release((handle *)data);
  free(data);
}
static napi_value wrap_handle(napi_env env, handle_kind kind, void *ptr, carta1_ctx *ctx) {
  handle *h = (handle *)calloc(1, sizeof *h);
  napi_value out;
  h->kind = kind; h->ptr = ptr; h->ctx = ctx;
  NAPI_OK(napi_create_external(env, h, handle_finalize, NULL, &out));
  return out;
}
static handle *get_handle(napi_env env, napi_value v, handle_kind kind) {
  handle *h = NULL;
  if (napi_get_value_external(env, v, (void **)&h) != napi_ok || !h || h->kind != kind || !h->ptr) {
    napi_throw_type_error(env, NULL, "carta1_b200: bad or destroyed handle");
    return NULL;
  }
  return h;
}
/* The reference throws Error / TypeError with fixed texts; the C ABI returns the same texts. */
static napi_value throw_abi(napi_env env, int rc, const carta1_ctx *ctx) {
  const char *msg = carta1_last_error(ctx);
  if (rc == CARTA1_ERR_ARG && strstr(msg, "requires")) napi_throw_type_error(env, NULL, msg);
  else napi_throw_error(env, NULL, msg);
  return NULL;
}

static int get_f64_array(napi_env env, napi_value obj, const char *name, double *dst, size_t n) {
  napi_value v;
  napi_typedarray_type t;
  size_t len;
  void *data;
  if (napi_get_named_property(env, obj, name, &v) != napi_ok) return 0;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, NULL, NULL) != napi_ok || t != napi_float64_array || len != n) return 0;
  memcpy(dst, data, n * sizeof(double));
  return 1;
}

/* tables: { windowShort, scaleFactors, mdctFwd64, mdctFwd256, mdctFwd512, mdctInv64, mdctInv256,
 * mdctInv512, fftW } as Float64Arrays computed by the host's own Math.* (SURVEY.md 0.3). */
static napi_value CreateContext(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  int32_t device = 0;
  napi_valuetype vt = napi_undefined;
  carta1_tables tables, *tp = NULL;
  carta1_ctx *ctx = NULL;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc > 0) NAPI_OK(napi_get_value_int32(env, argv[0], &device));
  if (argc > 1) NAPI_OK(napi_typeof(env, argv[1], &vt));
  if (vt == napi_object) {
    double fft[16];
    if (!get_f64_array(env, argv[1], "windowShort", tables.window_short, 32) ||
        !get_f64_array(env, argv[1], "scaleFactors", tables.scale_factors, 64) ||
        !get_f64_array(env, argv[1], "mdctFwd64", tables.mdct_fwd64, 32) ||
        !get_f64_array(env, argv[1], "mdctFwd256", tables.mdct_fwd256, 128) ||
        !get_f64_array(env, argv[1], "mdctFwd512", tables.mdct_fwd512, 256) ||
        !get_f64_array(env, argv[1], "mdctInv64", tables.mdct_inv64, 32) ||
        !get_f64_array(env, argv[1], "mdctInv256", tables.mdct_inv256, 128) ||
        !get_f64_array(env, argv[1], "mdctInv512", tables.mdct_inv512, 256) ||
        !get_f64_array(env, argv[1], "fftW", fft, 16)) {
      napi_throw_type_error(env, NULL, "carta1_b200: tables must hold the nine Float64Arrays");
      return NULL;
    }
    memcpy(tables.fft_w, fft, sizeof fft);
    tp = &tables;
  }
  if (carta1_ctx_create(device, tp, &ctx) != CARTA1_OK) return throw_abi(env, CARTA1_ERR_CUDA, NULL);
  return wrap_handle(env, H_CTX, ctx, ctx);
}

/* opts: { transientThresholdLow, allocationBias, fixedBlockModes|null, biasedScaleFactors?: Float64Array(64) } */
static int read_opts(napi_env env, napi_value v, carta1_enc_opts *o, double *bsf) {
  napi_valuetype vt;
  napi_value p;
  bool has = false, is_arr = false;
  carta1_default_enc_opts(o);
  if (napi_typeof(env, v, &vt) != napi_ok || vt != napi_object) return 1;
  if (napi_get_named_property(env, v, "transientThresholdLow", &p) == napi_ok) napi_get_value_double(env, p, &o->transient_threshold_low);
  if (napi_get_named_property(env, v, "allocationBias", &p) == napi_ok) napi_get_value_double(env, p, &o->allocation_bias);
  if (napi_get_named_property(env, v, "fixedBlockModes", &p) == napi_ok && napi_is_array(env, p, &is_arr) == napi_ok && is_arr) {
    uint32_t i;
    o->use_fixed_block_modes = 1;
    for (i = 0; i < 3; i++) {
      napi_value e;
      if (napi_get_element(env, p, i, &e) != napi_ok || napi_get_value_int32(env, e, &o->fixed_block_modes[i]) != napi_ok) return 0;
    }
  }
  if (napi_has_named_property(env, v, "biasedScaleFactors", &has) == napi_ok && has) {
    if (!get_f64_array(env, v, "biasedScaleFactors", bsf, 64)) return 0;
    o->biased_scale_factors = bsf;
  }
  return 1;
}

static napi_value CreateEncoder(napi_env env, napi_callback_info info) {
  size_t argc = 3;
  napi_value argv[3];
  handle *hc;
  carta1_enc_opts o;
  double bsf[64];
  int32_t n_streams = 1;
  carta1_encoder *enc = NULL;
  int rc;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2 || !(hc = get_handle(env, argv[0], H_CTX))) return NULL;
  if (!read_opts(env, argv[1], &o, bsf)) { napi_throw_type_error(env, NULL, "carta1_b200: bad encoder options"); return NULL; }
  if (argc > 2) NAPI_OK(napi_get_value_int32(env, argv[2], &n_streams));
  rc = carta1_enc_create((carta1_ctx *)hc->ptr, &o, n_streams, &enc);
  if (rc) return throw_abi(env, rc, (carta1_ctx *)hc->ptr);
  return wrap_handle(env, H_ENC, enc, (carta1_ctx *)hc->ptr);
}

static napi_value CreateDecoder(napi_env env, napi_callback_info info) {
  size_t argc = 2;
  napi_value argv[2];
  handle *hc;
  int32_t n_streams = 1;
  carta1_decoder *dec = NULL;
  int rc;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 1 || !(hc = get_handle(env, argv[0], H_CTX))) return NULL;
  if (argc > 1) NAPI_OK(napi_get_value_int32(env, argv[1], &n_streams));
  rc = carta1_dec_create((carta1_ctx *)hc->ptr, n_streams, &dec);
  if (rc) return throw_abi(env, rc, (carta1_ctx *)hc->ptr);
  return wrap_handle(env, H_DEC, dec, (carta1_ctx *)hc->ptr);
}

static int typed(napi_env env, napi_value v, napi_typedarray_type want, void **data, size_t *len) {
  napi_typedarray_type t;
  bool is = false;
  if (napi_is_typedarray(env, v, &is) != napi_ok || !is) return 0;
  if (napi_get_typedarray_info(env, v, &t, len, data, NULL, NULL) != napi_ok) return 0;
  return t == want;
}
static napi_value new_typed(napi_env env, napi_typedarray_type t, size_t elems, size_t elem_size, void **data) {
  napi_value ab, out;
  NAPI_OK(napi_create_arraybuffer(env, elems * elem_size, data, &ab));
  NAPI_OK(napi_create_typedarray(env, t, elems, ab, 0, &out));
  return out;
}

/* encodeFrames(enc, pcm: Float32Array[nStreams*nFrames*512], nFrames) -> Uint8Array[nStreams*nFrames*212]
 * (carta1_enc_frames; nStreams = nFrames = 1 is the reference's encoder(pcm) closure call). */
static napi_value EncodeFrames(napi_env env, napi_callback_info info) {
  size_t argc = 3, len = 0;
  napi_value argv[3], out;
  handle *he;
  void *pcm, *su;
  int32_t n_frames = 1;
  int rc;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2 || !(he = get_handle(env, argv[0], H_ENC))) return NULL;
  if (!typed(env, argv[1], napi_float32_array, &pcm, &len)) { napi_throw_type_error(env, NULL, "carta1_b200: pcm must be a Float32Array"); return NULL; }
  if (argc > 2) NAPI_OK(napi_get_value_int32(env, argv[2], &n_frames));
  if (n_frames <= 0 || len % ((size_t)n_frames * 512) != 0) { napi_throw_error(env, NULL, "carta1_b200: pcm length must be nStreams*nFrames*512"); return NULL; }
  out = new_typed(env, napi_uint8_array, len / 512 * CARTA1_SU_BYTES, 1, &su);
  if (!out) return NULL;
  rc = carta1_enc_frames((carta1_encoder *)he->ptr, (const float *)pcm, n_frames, (uint8_t *)su);
  if (rc) return throw_abi(env, rc, he->ctx);
  return out;
}

static napi_value DecodeFrames(napi_env env, napi_callback_info info) {
  size_t argc = 3, len = 0;
  napi_value argv[3], out;
  handle *hd;
  void *su, *pcm;
  int32_t n_frames = 1;
  int rc;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2 || !(hd = get_handle(env, argv[0], H_DEC))) return NULL;
  if (!typed(env, argv[1], napi_uint8_array, &su, &len)) { napi_throw_type_error(env, NULL, "carta1_b200: sound units must be a Uint8Array"); return NULL; }
  if (argc > 2) NAPI_OK(napi_get_value_int32(env, argv[2], &n_frames));
  if (n_frames <= 0 || len % ((size_t)n_frames * CARTA1_SU_BYTES) != 0) { napi_throw_error(env, NULL, "Frame must be 212 bytes"); return NULL; }
  out = new_typed(env, napi_float32_array, len / CARTA1_SU_BYTES * 512, 4, &pcm);
  if (!out) return NULL;
  rc = carta1_dec_frames((carta1_decoder *)hd->ptr, (const uint8_t *)su, n_frames, (float *)pcm);
  if (rc) return throw_abi(env, rc, hd->ctx);
  return out;
}

static napi_value DecodeFramesExpanded(napi_env env, napi_callback_info info) {
  size_t argc = 6, nq = 0, ns = 0, nb = 0, nm = 0;
  napi_value argv[6], out;
  handle *hd;
  void *q, *sfi, *bits, *modes, *pcm;
  int32_t n_frames = 1;
  int rc;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 6 || !(hd = get_handle(env, argv[0], H_DEC))) return NULL;
  if (!typed(env, argv[1], napi_int32_array, &q, &nq) || !typed(env, argv[2], napi_uint8_array, &sfi, &ns) ||
      !typed(env, argv[3], napi_uint8_array, &bits, &nb) || !typed(env, argv[4], napi_int32_array, &modes, &nm) ||
      nq != ns || nq != nb || nq % 512 != 0 || nm != nq / 512 * 3) {
    napi_throw_type_error(env, NULL, "carta1_b200: expanded frames are Int32Array q, Uint8Array sfi, Uint8Array bits (512 per frame) and Int32Array modes (3 per frame)");
    return NULL;
  }
  NAPI_OK(napi_get_value_int32(env, argv[5], &n_frames));
  out = new_typed(env, napi_float32_array, nq, 4, &pcm);
  if (!out) return NULL;
  rc = carta1_dec_frames_expanded((carta1_decoder *)hd->ptr, (const int32_t *)q, (const uint8_t *)sfi,
                                  (const uint8_t *)bits, (const int32_t *)modes, n_frames, (float *)pcm);
  if (rc) return throw_abi(env, rc, hd->ctx);
  return out;
}

static napi_value Destroy(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1], undef;
  handle *h = NULL;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc == 1 && napi_get_value_external(env, argv[0], (void **)&h) == napi_ok) handle_release(h);
  NAPI_OK(napi_get_undefined(env, &undef));
  return undef;
}

/* ---- whole-buffer helpers as Promises: the blocking C ABI call runs on the libuv pool ---- */
typedef struct {
  napi_async_work work;
  napi_deferred deferred;
  napi_ref keep[3];       /* input typed arrays stay alive while the worker reads them */
  int n_keep;
  carta1_ctx *ctx;
  int encode;
  /* encode */
  const float *chan[2];
  int n_ch;
  size_t n_samples;
  carta1_enc_opts opts;
  double bsf[64];
  /* decode */
  const uint8_t *su;
  size_t n_su;
  /* results: typed arrays created on the main thread before the work is queued; the library writes into
   * them (pageable memory: staged through the context's pinned bounce slots), so completion copies nothing */
  napi_ref out_ref[2];
  uint8_t *su_out;
  size_t su_cap, n_su_out;
  float *pcm_out[2];
  size_t frames;
  size_t halo_frames;     /* > 0: a shard of a longer stream (carta1_encode_pcm_shard / carta1_decode_su_shard) */
  int rc;
  char err[256];
} job;

static void job_execute(napi_env env, void *data) {
  job *j = (job *)data;
  (void)env;
  if (j->encode && j->halo_frames)
    j->rc = carta1_encode_pcm_shard(j->ctx, j->chan, j->n_ch, j->n_samples, j->halo_frames, &j->opts, j->su_out, j->su_cap, &j->n_su_out);
  else if (j->encode) j->rc = carta1_encode_pcm(j->ctx, j->chan, j->n_ch, j->n_samples, &j->opts, j->su_out, j->su_cap, &j->n_su_out);
  else if (j->halo_frames) j->rc = carta1_decode_su_shard(j->ctx, j->su, j->n_su, j->n_ch, j->halo_frames, j->pcm_out);
  else j->rc = carta1_decode_su(j->ctx, j->su, j->n_su, j->n_ch, j->pcm_out);
  if (j->rc) { strncpy(j->err, carta1_last_error(j->ctx), sizeof j->err - 1); j->err[sizeof j->err - 1] = 0; }
}

static void job_complete(napi_env env, napi_status status, void *data) {
  job *j = (job *)data;
  napi_value result = NULL, msg, err;
  int i;
  if (status == napi_ok && j->rc == 0) {
    if (j->encode) {
      if (napi_get_reference_value(env, j->out_ref[0], &result) != napi_ok) result = NULL;
    } else if (napi_create_array_with_length(env, (size_t)j->n_ch, &result) == napi_ok) {
      for (i = 0; i < j->n_ch; i++) {
        napi_value ch;
        if (napi_get_reference_value(env, j->out_ref[i], &ch) == napi_ok) napi_set_element(env, result, (uint32_t)i, ch);
      }
    }
  }
  if (result) {
    napi_resolve_deferred(env, j->deferred, result);
  } else {
    napi_create_string_utf8(env, j->rc ? j->err : "carta1_b200: async work failed", NAPI_AUTO_LENGTH, &msg);
    if (j->rc == CARTA1_ERR_ARG && strstr(j->err, "requires")) napi_create_type_error(env, NULL, msg, &err);
    else napi_create_error(env, NULL, msg, &err);
    napi_reject_deferred(env, j->deferred, err);
  }
  for (i = 0; i < j->n_keep; i++) napi_delete_reference(env, j->keep[i]);
  for (i = 0; i < 2; i++) if (j->out_ref[i]) napi_delete_reference(env, j->out_ref[i]);
  napi_delete_async_work(env, j->work);
  free(j);
}

static napi_value queue_job(napi_env env, job *j, const char *name) {
  napi_value promise, res_name;
  NAPI_OK(napi_create_promise(env, &j->deferred, &promise));
  NAPI_OK(napi_create_string_utf8(env, name, NAPI_AUTO_LENGTH, &res_name));
  NAPI_OK(napi_create_async_work(env, NULL, res_name, job_execute, job_complete, j, &j->work));
  NAPI_OK(napi_queue_async_work(env, j->work));
  return promise;
}

/* encodePcm(ctx, channels: Float32Array[1|2] (equal lengths; index.mjs zero-pads), opts[, haloFrames])
 * haloFrames (0 or >= 2): the channels start that many frames before the first frame to emit -- one shard of a
 * longer stream (carta1_encode_pcm_shard). */
static napi_value EncodePcm(napi_env env, napi_callback_info info) {
  size_t argc = 4;
  napi_value argv[4];
  handle *hc;
  uint32_t n_ch = 0, c;
  bool is_arr = false;
  job *j;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2 || !(hc = get_handle(env, argv[0], H_CTX))) return NULL;
  if (napi_is_array(env, argv[1], &is_arr) != napi_ok || !is_arr || napi_get_array_length(env, argv[1], &n_ch) != napi_ok ||
      (n_ch != 1 && n_ch != 2)) {
    napi_throw_type_error(env, NULL, "ATRAC1 encoding requires one or two Float32 channels");  /* processor.js:598-604 */
    return NULL;
  }
  j = (job *)calloc(1, sizeof *j);
  j->ctx = (carta1_ctx *)hc->ptr; j->encode = 1; j->n_ch = (int)n_ch;
  for (c = 0; c < n_ch; c++) {
    napi_value e;
    void *data;
    size_t len = 0;
    if (napi_get_element(env, argv[1], c, &e) != napi_ok || !typed(env, e, napi_float32_array, &data, &len) ||
        (c && len != j->n_samples)) {
      free(j);
      napi_throw_type_error(env, NULL, "ATRAC1 encoding requires one or two Float32 channels");
      return NULL;
    }
    j->chan[c] = (const float *)data; j->n_samples = len;
    napi_create_reference(env, e, 1, &j->keep[j->n_keep++]);
  }
  if (argc < 3 || !read_opts(env, argv[2], &j->opts, j->bsf)) carta1_default_enc_opts(&j->opts);
  if (argc >= 4) {
    uint32_t halo = 0;
    if (napi_get_value_uint32(env, argv[3], &halo) == napi_ok) j->halo_frames = halo;
  }
  {
    napi_value out;
    void *dst;
    const size_t all = carta1_frame_count(j->n_samples);
    j->su_cap = (all > j->halo_frames ? all - j->halo_frames : 0) * (size_t)n_ch * CARTA1_SU_BYTES;
    out = new_typed(env, napi_uint8_array, j->su_cap, 1, &dst);
    if (!out || napi_create_reference(env, out, 1, &j->out_ref[0]) != napi_ok) { free(j); return NULL; }
    j->su_out = (uint8_t *)dst;
  }
  return queue_job(env, j, "carta1_b200.encodePcm");
}

/* decodeSu(ctx, su: Uint8Array, channelCount[, haloFrames]); haloFrames >= 1: the units start that many frames before
 * the first frame to emit (carta1_decode_su_shard). */
static napi_value DecodeSu(napi_env env, napi_callback_info info) {
  size_t argc = 4, len = 0;
  napi_value argv[4];
  handle *hc;
  void *data;
  int32_t n_ch = 1;
  job *j;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 3 || !(hc = get_handle(env, argv[0], H_CTX))) return NULL;
  if (!typed(env, argv[1], napi_uint8_array, &data, &len)) {
    napi_throw_type_error(env, NULL, "ATRAC1 decoding requires AEA bytes or a Blob");  /* processor.js:628-637 */
    return NULL;
  }
  NAPI_OK(napi_get_value_int32(env, argv[2], &n_ch));
  if (n_ch != 1 && n_ch != 2) { napi_throw_error(env, NULL, "Unsupported channel count"); return NULL; }
  j = (job *)calloc(1, sizeof *j);
  j->ctx = (carta1_ctx *)hc->ptr; j->n_ch = n_ch; j->su = (const uint8_t *)data; j->n_su = len / CARTA1_SU_BYTES;
  napi_create_reference(env, argv[1], 1, &j->keep[j->n_keep++]);
  j->frames = (j->n_su + (size_t)n_ch - 1) / (size_t)n_ch;
  if (argc >= 4) {
    uint32_t halo = 0;
    if (napi_get_value_uint32(env, argv[3], &halo) == napi_ok) j->halo_frames = halo;
    j->frames = j->frames > j->halo_frames ? j->frames - j->halo_frames : 0;
  }
  {
    int c;
    for (c = 0; c < n_ch; c++) {
      void *dst;
      napi_value out = new_typed(env, napi_float32_array, j->frames * 512, 4, &dst);
      if (!out || napi_create_reference(env, out, 1, &j->out_ref[c]) != napi_ok) { free(j); return NULL; }
      j->pcm_out[c] = (float *)dst;
    }
  }
  return queue_job(env, j, "carta1_b200.decodeSu");
}

/* deserializeUnits(ctx, su: Uint8Array) -> [nBfu: Uint8Array(n), blockModes: Int8Array(3n), wl: Uint8Array(52n),
 * sfi: Uint8Array(52n), q: Int32Array(512n)]: deserializeFrame over a whole file (bin/cli.js:567-677), see
 * carta1_deserialize_units for the layouts. */
static napi_value DeserializeUnits(napi_env env, napi_callback_info info) {
  size_t argc = 2, len = 0, n;
  napi_value argv[2], result, v[5];
  handle *hc;
  void *su, *p[5];
  int rc, i;
  NAPI_OK(napi_get_cb_info(env, info, &argc, argv, NULL, NULL));
  if (argc < 2 || !(hc = get_handle(env, argv[0], H_CTX))) return NULL;
  if (!typed(env, argv[1], napi_uint8_array, &su, &len) || len % CARTA1_SU_BYTES) {
    napi_throw_error(env, NULL, "Frame must be 212 bytes");  /* serialization.js:112-114 */
    return NULL;
  }
  n = len / CARTA1_SU_BYTES;
  v[0] = new_typed(env, napi_uint8_array, n, 1, &p[0]);
  v[1] = new_typed(env, napi_int8_array, 3 * n, 1, &p[1]);
  v[2] = new_typed(env, napi_uint8_array, 52 * n, 1, &p[2]);
  v[3] = new_typed(env, napi_uint8_array, 52 * n, 1, &p[3]);
  v[4] = new_typed(env, napi_int32_array, 512 * n, 4, &p[4]);
  for (i = 0; i < 5; i++) if (!v[i]) return NULL;
  rc = carta1_deserialize_units((carta1_ctx *)hc->ptr, (const uint8_t *)su, n, (uint8_t *)p[0], (int8_t *)p[1],
                                (uint8_t *)p[2], (uint8_t *)p[3], (int32_t *)p[4]);
  if (rc) return throw_abi(env, rc, (carta1_ctx *)hc->ptr);
  NAPI_OK(napi_create_array_with_length(env, 5, &result));
  for (i = 0; i < 5; i++) NAPI_OK(napi_set_element(env, result, (uint32_t)i, v[i]));
  return result;
}

static napi_value Init(napi_env env, napi_value exports) {
  static const napi_property_descriptor props[] = {
      {"createContext", NULL, CreateContext, NULL, NULL, NULL, napi_default, NULL},
      {"createEncoder", NULL, CreateEncoder, NULL, NULL, NULL, napi_default, NULL},
      {"createDecoder", NULL, CreateDecoder, NULL, NULL, NULL, napi_default, NULL},
      {"encodeFrames", NULL, EncodeFrames, NULL, NULL, NULL, napi_default, NULL},
      {"decodeFrames", NULL, DecodeFrames, NULL, NULL, NULL, napi_default, NULL},
      {"decodeFramesExpanded", NULL, DecodeFramesExpanded, NULL, NULL, NULL, napi_default, NULL},
      {"encodePcm", NULL, EncodePcm, NULL, NULL, NULL, napi_default, NULL},
      {"decodeSu", NULL, DecodeSu, NULL, NULL, NULL, napi_default, NULL},
      {"deserializeUnits", NULL, DeserializeUnits, NULL, NULL, NULL, napi_default, NULL},
      {"destroy", NULL, Destroy, NULL, NULL, NULL, napi_default, NULL},
  };
  NAPI_OK(napi_define_properties(env, exports, sizeof props / sizeof props[0], props));
  return exports;
}

NAPI_MODULE(carta1_b200, Init)
