/*
 * node_api_min.h -- the subset of Node-API (node_api.h / js_native_api.h, NAPI_VERSION 8) that
 * carta1_napi.c uses, declared by hand.
 *
 * Node is not installed in the build image, so there is no node_api.h to
 * compile against.  build() compiles the shim against these declarations (and links it with tests/napi_host); a real build
 * (binding.gyp) uses Node's own header instead (CARTA1_NAPI_USE_NODE_HEADERS).
 * Names, enum values and signatures follow the stable Node-API ABI.
 */
#ifndef CARTA1_NODE_API_MIN_H
#define CARTA1_NODE_API_MIN_H
#include <stddef.h>
#include <stdint.h>
#include <stdbool.h>

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_ref__ *napi_ref;
typedef struct napi_deferred__ *napi_deferred;
typedef struct napi_callback_info__ *napi_callback_info;
typedef struct napi_async_work__ *napi_async_work;

typedef enum { napi_ok = 0, napi_invalid_arg, napi_object_expected, napi_string_expected, napi_name_expected,
               napi_function_expected, napi_number_expected, napi_boolean_expected, napi_array_expected,
               napi_generic_failure, napi_pending_exception, napi_cancelled } napi_status;
typedef enum { napi_undefined, napi_null, napi_boolean, napi_number, napi_string, napi_symbol, napi_object,
               napi_function, napi_external, napi_bigint } napi_valuetype;
typedef enum { napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
               napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array, napi_bigint64_array,
               napi_biguint64_array } napi_typedarray_type;
typedef enum { napi_default = 0 } napi_property_attributes;

typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);
typedef void (*napi_async_execute_callback)(napi_env env, void *data);
typedef void (*napi_async_complete_callback)(napi_env env, napi_status status, void *data);

typedef struct {
  const char *utf8name;
  napi_value name;
  napi_callback method, getter, setter;
  napi_value value;
  napi_property_attributes attributes;
  void *data;
} napi_property_descriptor;

typedef napi_value (*napi_addon_register_func)(napi_env env, napi_value exports);
typedef struct napi_module {
  int nm_version;
  unsigned int nm_flags;
  const char *nm_filename;
  napi_addon_register_func nm_register_func;
  const char *nm_modname;
  void *nm_priv;
  void *reserved[4];
} napi_module;

#define NAPI_AUTO_LENGTH SIZE_MAX
#ifdef __cplusplus
extern "C" {
#endif
void napi_module_register(napi_module *mod);
napi_status napi_define_properties(napi_env, napi_value object, size_t count, const napi_property_descriptor *);
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
napi_status napi_typeof(napi_env, napi_value, napi_valuetype *result);
napi_status napi_get_undefined(napi_env, napi_value *result);
napi_status napi_get_value_double(napi_env, napi_value, double *result);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t *result);
napi_status napi_get_value_uint32(napi_env, napi_value, uint32_t *result);
napi_status napi_get_value_bool(napi_env, napi_value, bool *result);
napi_status napi_create_double(napi_env, double value, napi_value *result);
napi_status napi_create_string_utf8(napi_env, const char *str, size_t length, napi_value *result);
napi_status napi_get_named_property(napi_env, napi_value object, const char *utf8name, napi_value *result);
napi_status napi_has_named_property(napi_env, napi_value object, const char *utf8name, bool *result);
napi_status napi_is_array(napi_env, napi_value, bool *result);
napi_status napi_get_array_length(napi_env, napi_value, uint32_t *result);
napi_status napi_get_element(napi_env, napi_value object, uint32_t index, napi_value *result);
napi_status napi_set_element(napi_env, napi_value object, uint32_t index, napi_value value);
napi_status napi_create_array_with_length(napi_env, size_t length, napi_value *result);
napi_status napi_is_typedarray(napi_env, napi_value, bool *result);
napi_status napi_get_typedarray_info(napi_env, napi_value typedarray, napi_typedarray_type *type, size_t *length,
                                     void **data, napi_value *arraybuffer, size_t *byte_offset);
napi_status napi_create_arraybuffer(napi_env, size_t byte_length, void **data, napi_value *result);
napi_status napi_create_external_arraybuffer(napi_env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                             void *finalize_hint, napi_value *result);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type type, size_t length, napi_value arraybuffer,
                                   size_t byte_offset, napi_value *result);
napi_status napi_create_external(napi_env, void *data, napi_finalize finalize_cb, void *finalize_hint, napi_value *result);
napi_status napi_get_value_external(napi_env, napi_value, void **result);
napi_status napi_throw_error(napi_env, const char *code, const char *msg);
napi_status napi_throw_type_error(napi_env, const char *code, const char *msg);
napi_status napi_create_error(napi_env, napi_value code, napi_value msg, napi_value *result);
napi_status napi_create_type_error(napi_env, napi_value code, napi_value msg, napi_value *result);
napi_status napi_create_reference(napi_env, napi_value value, uint32_t initial_refcount, napi_ref *result);
napi_status napi_get_reference_value(napi_env, napi_ref ref, napi_value *result);
napi_status napi_delete_reference(napi_env, napi_ref ref);
napi_status napi_create_promise(napi_env, napi_deferred *deferred, napi_value *promise);
napi_status napi_resolve_deferred(napi_env, napi_deferred deferred, napi_value resolution);
napi_status napi_reject_deferred(napi_env, napi_deferred deferred, napi_value rejection);
napi_status napi_create_async_work(napi_env, napi_value async_resource, napi_value async_resource_name,
                                   napi_async_execute_callback execute, napi_async_complete_callback complete,
                                   void *data, napi_async_work *result);
napi_status napi_queue_async_work(napi_env, napi_async_work work);
napi_status napi_delete_async_work(napi_env, napi_async_work work);
#ifdef __cplusplus
}
#endif

#define NAPI_MODULE_X(modname, regfunc)                                                       \
  static napi_module carta1_napi_module_ = {1, 0, __FILE__, regfunc, #modname, NULL, {0}};    \
  static void carta1_napi_register_(void) __attribute__((constructor));                       \
  static void carta1_napi_register_(void) { napi_module_register(&carta1_napi_module_); }
#define NAPI_MODULE(modname, regfunc) NAPI_MODULE_X(modname, regfunc)
#endif
