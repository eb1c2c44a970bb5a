/*
 * napi_host.c -- a minimal Node-API host, so that carta1_b200/napi/carta1_napi.c can be LINKED AND EXECUTED
 * without Node.  TEST INFRASTRUCTURE ONLY.
 *
 * Node-API is a C ABI: an addon sees its JavaScript host only through the napi_* functions.  This file implements
 * the subset the shim uses (the declarations of carta1_b200/napi/node_api_min.h, which follow node_api.h) over a
 * small tagged-value heap: undefined, numbers, strings, objects with named properties, arrays, ArrayBuffers, typed
 * arrays, externals with finalizers, errors, promises, references, and async work whose execute callback runs on
 * another thread (as libuv's pool would run it) and whose complete callback runs back on the caller's.
 * The shim is compiled unchanged together with this file into tests/napi_host/libcarta1_napi_host.so; the test
 * (tests/test_napi_host.py) plays the JavaScript side through the host_* functions below: it builds argument
 * values, calls the addon's exported functions, and inspects what comes back (values, thrown exceptions, settled
 * promises).  Values are never freed except through host_release_external (finalizer semantics).
 */
#define _GNU_SOURCE
#include <pthread.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include "../../carta1_b200/napi/node_api_min.h"

enum { V_UNDEFINED, V_NUMBER, V_STRING, V_OBJECT, V_ARRAY, V_ARRAYBUFFER, V_TYPEDARRAY, V_EXTERNAL, V_FUNCTION, V_ERROR, V_PROMISE };

typedef struct prop { char *name; napi_value value; struct prop *next; } prop;
struct napi_value__ {
  int kind;
  double num;
  char *str;
  prop *props;                          /* V_OBJECT */
  napi_value *items; size_t n_items;    /* V_ARRAY */
  void *data; size_t byte_length;       /* V_ARRAYBUFFER; V_EXTERNAL uses data */
  napi_typedarray_type ta_type; size_t ta_length; napi_value ta_buffer; size_t ta_offset; /* V_TYPEDARRAY */
  napi_finalize finalize; void *finalize_hint; int finalized; /* V_EXTERNAL */
  napi_callback cb; void *cb_data;      /* V_FUNCTION */
  int is_type_error;                    /* V_ERROR (message in str) */
  int state; napi_value settled;        /* V_PROMISE: 0 pending, 1 resolved, 2 rejected */
};
struct napi_env__ { napi_value pending; napi_value exports; };
struct napi_ref__ { napi_value value; };
struct napi_deferred__ { napi_value promise; };
struct napi_callback_info__ { size_t argc; napi_value *argv; void *data; };
struct napi_async_work__ { napi_async_execute_callback execute; napi_async_complete_callback complete; void *data; napi_env env; };

static struct napi_env__ g_env;
static napi_module *g_module;
static struct napi_value__ g_undefined = {V_UNDEFINED};

static napi_value new_value(int kind) {
  napi_value v = (napi_value)calloc(1, sizeof *v);
  v->kind = kind;
  return v;
}
static size_t elem_size(napi_typedarray_type t) {
  switch (t) {
    case napi_int8_array: case napi_uint8_array: case napi_uint8_clamped_array: return 1;
    case napi_int16_array: case napi_uint16_array: return 2;
    case napi_int32_array: case napi_uint32_array: case napi_float32_array: return 4;
    default: return 8;
  }
}
static napi_value make_error(int is_type, const char *msg) {
  napi_value e = new_value(V_ERROR);
  e->is_type_error = is_type;
  e->str = strdup(msg ? msg : "");
  return e;
}

/* ---- the napi_* functions the shim calls --------------------------------------------------------------- */
void napi_module_register(napi_module *mod) { g_module = mod; }

napi_status napi_define_properties(napi_env env, napi_value object, size_t count, const napi_property_descriptor *p) {
  (void)env;
  if (!object || object->kind != V_OBJECT) return napi_object_expected;
  for (size_t i = 0; i < count; i++) {
    napi_value v = p[i].value;
    if (p[i].method) { v = new_value(V_FUNCTION); v->cb = p[i].method; v->cb_data = p[i].data; }
    prop *q = (prop *)calloc(1, sizeof *q);
    q->name = strdup(p[i].utf8name); q->value = v; q->next = object->props; object->props = q;
  }
  return napi_ok;
}
napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data) {
  (void)env;
  if (argc) {
    const size_t cap = *argc;
    for (size_t i = 0; i < cap && argv; i++) argv[i] = i < info->argc ? info->argv[i] : &g_undefined;
    *argc = info->argc;  /* Node reports the actual count, which may exceed the capacity */
  }
  if (this_arg) *this_arg = &g_undefined;
  if (data) *data = info->data;
  return napi_ok;
}
napi_status napi_typeof(napi_env env, napi_value v, napi_valuetype *result) {
  (void)env;
  switch (v->kind) {
    case V_UNDEFINED: *result = napi_undefined; break;
    case V_NUMBER: *result = napi_number; break;
    case V_STRING: *result = napi_string; break;
    case V_EXTERNAL: *result = napi_external; break;
    case V_FUNCTION: *result = napi_function; break;
    default: *result = napi_object; break;
  }
  return napi_ok;
}
napi_status napi_get_undefined(napi_env env, napi_value *result) { (void)env; *result = &g_undefined; return napi_ok; }
napi_status napi_get_value_double(napi_env env, napi_value v, double *result) {
  (void)env; if (v->kind != V_NUMBER) return napi_number_expected; *result = v->num; return napi_ok;
}
static int32_t to_int32(double d) {  /* ECMAScript ToInt32, as napi_get_value_int32 does for finite numbers */
  if (d != d || d - d != 0) return 0;
  double t = d < 0 ? -__builtin_floor(-d) : __builtin_floor(d);
  double m = __builtin_fmod(t, 4294967296.0);
  if (m < 0) m += 4294967296.0;
  return (int32_t)(uint32_t)m;
}
napi_status napi_get_value_int32(napi_env env, napi_value v, int32_t *result) {
  (void)env; if (v->kind != V_NUMBER) return napi_number_expected; *result = to_int32(v->num); return napi_ok;
}
napi_status napi_get_value_uint32(napi_env env, napi_value v, uint32_t *result) {
  (void)env; if (v->kind != V_NUMBER) return napi_number_expected; *result = (uint32_t)to_int32(v->num); return napi_ok;
}
napi_status napi_get_value_bool(napi_env env, napi_value v, bool *result) { (void)env; (void)v; (void)result; return napi_boolean_expected; }
napi_status napi_create_double(napi_env env, double value, napi_value *result) {
  (void)env; *result = new_value(V_NUMBER); (*result)->num = value; return napi_ok;
}
napi_status napi_create_string_utf8(napi_env env, const char *str, size_t length, napi_value *result) {
  (void)env;
  *result = new_value(V_STRING);
  (*result)->str = length == NAPI_AUTO_LENGTH ? strdup(str) : strndup(str, length);
  return napi_ok;
}
static prop *find_prop(napi_value object, const char *name) {
  for (prop *q = object->props; q; q = q->next) if (!strcmp(q->name, name)) return q;
  return NULL;
}
napi_status napi_get_named_property(napi_env env, napi_value object, const char *utf8name, napi_value *result) {
  (void)env;
  if (!object || (object->kind != V_OBJECT && object->kind != V_ARRAY && object->kind != V_TYPEDARRAY)) return napi_object_expected;
  prop *q = find_prop(object, utf8name);
  *result = q ? q->value : &g_undefined;   /* a missing property reads as undefined, as in JavaScript */
  return napi_ok;
}
napi_status napi_has_named_property(napi_env env, napi_value object, const char *utf8name, bool *result) {
  (void)env;
  if (!object || object->kind != V_OBJECT) return napi_object_expected;
  *result = find_prop(object, utf8name) != NULL;
  return napi_ok;
}
napi_status napi_is_array(napi_env env, napi_value v, bool *result) { (void)env; *result = v->kind == V_ARRAY; return napi_ok; }
napi_status napi_get_array_length(napi_env env, napi_value v, uint32_t *result) {
  (void)env; if (v->kind != V_ARRAY) return napi_array_expected; *result = (uint32_t)v->n_items; return napi_ok;
}
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result) {
  (void)env;
  if (object->kind != V_ARRAY) return napi_object_expected;
  *result = index < object->n_items && object->items[index] ? object->items[index] : &g_undefined;
  return napi_ok;
}
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value) {
  (void)env;
  if (object->kind != V_ARRAY) return napi_object_expected;
  if (index >= object->n_items) {
    object->items = (napi_value *)realloc(object->items, (index + 1) * sizeof(napi_value));
    for (size_t i = object->n_items; i <= index; i++) object->items[i] = NULL;
    object->n_items = index + 1;
  }
  object->items[index] = value;
  return napi_ok;
}
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result) {
  (void)env;
  *result = new_value(V_ARRAY);
  (*result)->items = (napi_value *)calloc(length ? length : 1, sizeof(napi_value));
  (*result)->n_items = length;
  return napi_ok;
}
napi_status napi_is_typedarray(napi_env env, napi_value v, bool *result) { (void)env; *result = v->kind == V_TYPEDARRAY; return napi_ok; }
napi_status napi_get_typedarray_info(napi_env env, napi_value v, napi_typedarray_type *type, size_t *length, void **data,
                                     napi_value *arraybuffer, size_t *byte_offset) {
  (void)env;
  if (v->kind != V_TYPEDARRAY) return napi_invalid_arg;
  if (type) *type = v->ta_type;
  if (length) *length = v->ta_length;
  if (data) *data = (char *)v->ta_buffer->data + v->ta_offset;
  if (arraybuffer) *arraybuffer = v->ta_buffer;
  if (byte_offset) *byte_offset = v->ta_offset;
  return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result) {
  (void)env;
  *result = new_value(V_ARRAYBUFFER);
  (*result)->data = calloc(byte_length ? byte_length : 1, 1);   /* pageable, zero-filled, like V8's */
  (*result)->byte_length = byte_length;
  if (data) *data = (*result)->data;
  return napi_ok;
}
napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                             void *finalize_hint, napi_value *result) {
  (void)env; (void)finalize_cb; (void)finalize_hint;
  *result = new_value(V_ARRAYBUFFER);
  (*result)->data = external_data;
  (*result)->byte_length = byte_length;
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value *result) {
  (void)env;
  if (arraybuffer->kind != V_ARRAYBUFFER || byte_offset + length * elem_size(type) > arraybuffer->byte_length) return napi_invalid_arg;
  *result = new_value(V_TYPEDARRAY);
  (*result)->ta_type = type; (*result)->ta_length = length; (*result)->ta_buffer = arraybuffer; (*result)->ta_offset = byte_offset;
  return napi_ok;
}
napi_status napi_create_external(napi_env env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result) {
  (void)env;
  *result = new_value(V_EXTERNAL);
  (*result)->data = data; (*result)->finalize = finalize_cb; (*result)->finalize_hint = finalize_hint;
  return napi_ok;
}
napi_status napi_get_value_external(napi_env env, napi_value v, void **result) {
  (void)env; if (v->kind != V_EXTERNAL || v->finalized) return napi_invalid_arg; *result = v->data; return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char *code, const char *msg) { (void)code; if (!env->pending) env->pending = make_error(0, msg); return napi_ok; }
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg) { (void)code; if (!env->pending) env->pending = make_error(1, msg); return napi_ok; }
napi_status napi_create_error(napi_env env, napi_value code, napi_value msg, napi_value *result) {
  (void)env; (void)code; *result = make_error(0, msg && msg->kind == V_STRING ? msg->str : ""); return napi_ok;
}
napi_status napi_create_type_error(napi_env env, napi_value code, napi_value msg, napi_value *result) {
  (void)env; (void)code; *result = make_error(1, msg && msg->kind == V_STRING ? msg->str : ""); return napi_ok;
}
napi_status napi_create_reference(napi_env env, napi_value value, uint32_t initial_refcount, napi_ref *result) {
  (void)env; (void)initial_refcount; *result = (napi_ref)calloc(1, sizeof **result); (*result)->value = value; return napi_ok;
}
napi_status napi_get_reference_value(napi_env env, napi_ref ref, napi_value *result) { (void)env; *result = ref->value; return napi_ok; }
napi_status napi_delete_reference(napi_env env, napi_ref ref) { (void)env; free(ref); return napi_ok; }
napi_status napi_create_promise(napi_env env, napi_deferred *deferred, napi_value *promise) {
  (void)env;
  *promise = new_value(V_PROMISE);
  *deferred = (napi_deferred)calloc(1, sizeof **deferred);
  (*deferred)->promise = *promise;
  return napi_ok;
}
napi_status napi_resolve_deferred(napi_env env, napi_deferred d, napi_value v) { (void)env; d->promise->state = 1; d->promise->settled = v; free(d); return napi_ok; }
napi_status napi_reject_deferred(napi_env env, napi_deferred d, napi_value v) { (void)env; d->promise->state = 2; d->promise->settled = v; free(d); return napi_ok; }
napi_status napi_create_async_work(napi_env env, napi_value res, napi_value name, napi_async_execute_callback execute,
                                   napi_async_complete_callback complete, void *data, napi_async_work *result) {
  (void)res; (void)name;
  *result = (napi_async_work)calloc(1, sizeof **result);
  (*result)->execute = execute; (*result)->complete = complete; (*result)->data = data; (*result)->env = env;
  return napi_ok;
}
static void *run_execute(void *p) {
  napi_async_work w = (napi_async_work)p;
  w->execute(w->env, w->data);  /* on a pool thread: no napi_value may be touched here, and the shim does not */
  return NULL;
}
napi_status napi_queue_async_work(napi_env env, napi_async_work w) {
  /* libuv runs execute on a pool thread and complete on the loop thread afterwards; this host does the same,
   * only it waits for the pool thread at once instead of returning to an event loop first */
  pthread_t th;
  if (pthread_create(&th, NULL, run_execute, w) != 0) return napi_generic_failure;
  pthread_join(th, NULL);
  w->complete(env, napi_ok, w->data);
  return napi_ok;
}
napi_status napi_delete_async_work(napi_env env, napi_async_work w) { (void)env; free(w); return napi_ok; }

/* ---- the JavaScript side, for the test ------------------------------------------------------------------- */
napi_value host_load(void) {   /* what `require('carta1_b200.node')` does: run the registered init function */
  if (!g_module) return NULL;
  if (!g_env.exports) {
    g_env.exports = new_value(V_OBJECT);
    napi_value r = g_module->nm_register_func(&g_env, g_env.exports);
    if (r && r != g_env.exports) g_env.exports = r;
  }
  return g_env.exports;
}
const char *host_module_name(void) { return g_module ? g_module->nm_modname : ""; }
int host_export_names(char *buf, size_t cap) {   /* comma-separated names of exported functions */
  size_t at = 0; int n = 0;
  buf[0] = 0;
  for (prop *q = host_load()->props; q; q = q->next, n++) at += (size_t)snprintf(buf + at, at < cap ? cap - at : 0, "%s%s", n ? "," : "", q->name);
  return n;
}
napi_value host_undefined(void) { return &g_undefined; }
napi_value host_number(double d) { napi_value v; napi_create_double(&g_env, d, &v); return v; }
napi_value host_string(const char *s) { napi_value v; napi_create_string_utf8(&g_env, s, NAPI_AUTO_LENGTH, &v); return v; }
napi_value host_object(void) { return new_value(V_OBJECT); }
void host_set(napi_value obj, const char *name, napi_value v) {
  prop *q = find_prop(obj, name);
  if (q) { q->value = v; return; }
  q = (prop *)calloc(1, sizeof *q);
  q->name = strdup(name); q->value = v; q->next = obj->props; obj->props = q;
}
napi_value host_array(size_t n) { napi_value v; napi_create_array_with_length(&g_env, n, &v); return v; }
void host_array_set(napi_value arr, uint32_t i, napi_value v) { napi_set_element(&g_env, arr, i, v); }
/* a typed array over caller memory (the test keeps it alive), e.g. a numpy buffer: what a JS typed array is to an addon */
napi_value host_typedarray(int type, void *data, size_t length) {
  napi_value ab, v;
  napi_create_external_arraybuffer(&g_env, data, length * elem_size((napi_typedarray_type)type), NULL, NULL, &ab);
  napi_create_typedarray(&g_env, (napi_typedarray_type)type, length, ab, 0, &v);
  return v;
}
/* addon.<name>(...argv); NULL when the call threw (see host_take_exception) */
napi_value host_call(const char *name, size_t argc, napi_value *argv) {
  prop *q = find_prop(host_load(), name);
  if (!q || q->value->kind != V_FUNCTION) { napi_throw_type_error(&g_env, NULL, "not a function"); return NULL; }
  struct napi_callback_info__ info = {argc, argv, q->value->cb_data};
  napi_value r = q->value->cb(&g_env, &info);
  if (g_env.pending) return NULL;
  return r ? r : &g_undefined;
}
/* 0 = nothing pending, 1 = Error, 2 = TypeError; the message is copied out and the exception cleared */
int host_take_exception(char *buf, size_t cap) {
  if (!g_env.pending) return 0;
  const int kind = g_env.pending->is_type_error ? 2 : 1;
  snprintf(buf, cap, "%s", g_env.pending->str);
  g_env.pending = NULL;
  return kind;
}
int host_kind(napi_value v) { return v->kind; }
double host_number_value(napi_value v) { return v->num; }
int host_typedarray_info(napi_value v, int *type, size_t *length, void **data) {
  napi_typedarray_type t;
  if (napi_get_typedarray_info(&g_env, v, &t, length, data, NULL, NULL) != napi_ok) return 0;
  *type = (int)t;
  return 1;
}
size_t host_array_length(napi_value v) { return v->kind == V_ARRAY ? v->n_items : 0; }
napi_value host_array_get(napi_value v, uint32_t i) { napi_value r; napi_get_element(&g_env, v, i, &r); return r; }
int host_promise_state(napi_value v) { return v->kind == V_PROMISE ? v->state : -1; }
napi_value host_promise_value(napi_value v) { return v->settled; }
int host_error_info(napi_value v, char *buf, size_t cap) {   /* 1 = Error, 2 = TypeError, 0 = not an error value */
  if (v->kind != V_ERROR) return 0;
  snprintf(buf, cap, "%s", v->str);
  return v->is_type_error ? 2 : 1;
}
/* garbage collection of an external: its finalizer runs once */
void host_release_external(napi_value v) {
  if (v->kind == V_EXTERNAL && !v->finalized) {
    v->finalized = 1;
    if (v->finalize) v->finalize(&g_env, v->data, v->finalize_hint);
  }
}
