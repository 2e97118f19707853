/* Minimal declarations of the Node-API surface used by fmcw_napi.cc, for SYNTAX CHECKING ONLY in
 * environments without Node.js headers (this container).  Not a substitute for node_api.h. */
#ifndef FMCW_STUB_NODE_API_H
#define FMCW_STUB_NODE_API_H
#include <stddef.h>
#include <stdint.h>
typedef struct napi_env__* napi_env; typedef struct napi_value__* napi_value; typedef struct napi_ref__* napi_ref;
typedef struct napi_deferred__* napi_deferred; typedef struct napi_async_work__* napi_async_work;
typedef struct napi_callback_info__* napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef enum { napi_int16_array = 2, napi_int32_array = 5, napi_float32_array = 7, napi_float64_array = 8 } napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env, napi_callback_info);
typedef void (*napi_finalize)(napi_env, void*, void*);
typedef void (*napi_async_execute_callback)(napi_env, void*);
typedef void (*napi_async_complete_callback)(napi_env, napi_status, void*);
#define NAPI_AUTO_LENGTH ((size_t)-1)
#ifdef __cplusplus
extern "C" {
#endif
napi_status napi_get_cb_info(napi_env, napi_callback_info, size_t*, napi_value*, napi_value*, void**);
napi_status napi_has_named_property(napi_env, napi_value, const char*, bool*);
napi_status napi_get_named_property(napi_env, napi_value, const char*, napi_value*);
napi_status napi_set_named_property(napi_env, napi_value, const char*, napi_value);
napi_status napi_get_value_double(napi_env, napi_value, double*);
napi_status napi_get_value_int32(napi_env, napi_value, int32_t*);
napi_status napi_get_value_int64(napi_env, napi_value, int64_t*);
napi_status napi_get_typedarray_info(napi_env, napi_value, napi_typedarray_type*, size_t*, void**, napi_value*, size_t*);
napi_status napi_create_external(napi_env, void*, napi_finalize, void*, napi_value*);
napi_status napi_get_value_external(napi_env, napi_value, void**);
napi_status napi_throw_error(napi_env, const char*, const char*);
napi_status napi_create_arraybuffer(napi_env, size_t, void**, napi_value*);
napi_status napi_create_typedarray(napi_env, napi_typedarray_type, size_t, napi_value, size_t, napi_value*);
napi_status napi_create_string_utf8(napi_env, const char*, size_t, napi_value*);
napi_status napi_create_error(napi_env, napi_value, napi_value, napi_value*);
napi_status napi_create_object(napi_env, napi_value*);
napi_status napi_create_double(napi_env, double, napi_value*);
napi_status napi_create_promise(napi_env, napi_deferred*, napi_value*);
napi_status napi_resolve_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_reject_deferred(napi_env, napi_deferred, napi_value);
napi_status napi_create_reference(napi_env, napi_value, uint32_t, napi_ref*);
napi_status napi_delete_reference(napi_env, napi_ref);
napi_status napi_create_async_work(napi_env, napi_value, napi_value, napi_async_execute_callback, napi_async_complete_callback, void*, napi_async_work*);
napi_status napi_queue_async_work(napi_env, napi_async_work);
napi_status napi_delete_async_work(napi_env, napi_async_work);
napi_status napi_create_function(napi_env, const char*, size_t, napi_callback, void*, napi_value*);
#ifdef __cplusplus
}
#endif
#define NAPI_MODULE(name, init) extern "C" napi_value napi_register_module_v1(napi_env env, napi_value exports) { return init(env, exports); }
#endif
