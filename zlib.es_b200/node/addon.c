/*
 * addon.c — thin N-API addon over the C ABI of libzles.so (include/zles.h).
 *
 * Exposes the reference's two functions with the reference's signatures
 * (/root/reference/dist/tsc/zlib.d.ts:4-5):
 *     deflate(input: Uint8Array): Uint8Array      ->  zles_deflate
 *     inflate(input: Uint8Array): Uint8Array      ->  zles_inflate_alloc
 * plus deflateBatch / inflateBatch (arrays of Uint8Array) for the small-message shape.
 * Non-zero status codes become `throw new Error(msg)` with the reference's exact
 * message strings (zles_strerror).
 *
 * The build image has no Node.js and no node_api.h, so the handful of N-API
 * declarations used here are restated below (they are ABI-stable by design:
 * NAPI_VERSION 3).  This file is compile-checked by __graft_entry__.build();
 * it links when Node dlopen()s the resulting zles.node (napi_* symbols are
 * provided by the node binary), see INTEGRATION.md.
 */
#include <stddef.h>
#include <stdint.h>
#include <stdlib.h>

#include "zles.h"

/* ---- minimal N-API surface (node_api.h / js_native_api.h, NAPI_VERSION 3) ---- */
typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok = 0 } napi_status;
typedef enum {
  napi_int8_array, napi_uint8_array, napi_uint8_clamped_array, napi_int16_array, napi_uint16_array,
  napi_int32_array, napi_uint32_array, napi_float32_array, napi_float64_array
} napi_typedarray_type;
typedef napi_value (*napi_callback)(napi_env env, napi_callback_info info);
typedef enum { napi_default = 0 } napi_property_attributes;
typedef struct {
  const char *utf8name;
  napi_value name;
  napi_callback method;
  napi_callback getter;
  napi_callback setter;
  napi_value value;
  napi_property_attributes attributes;
  void *data;
} napi_property_descriptor;
typedef void (*napi_finalize)(napi_env env, void *finalize_data, void *finalize_hint);

extern napi_status napi_get_cb_info(napi_env env, napi_callback_info cbinfo, size_t *argc, napi_value *argv, napi_value *this_arg, void **data);
extern napi_status napi_is_typedarray(napi_env env, napi_value value, _Bool *result);
extern napi_status napi_get_typedarray_info(napi_env env, napi_value typedarray, napi_typedarray_type *type, size_t *length, void **data,
                                            napi_value *arraybuffer, size_t *byte_offset);
extern napi_status napi_create_arraybuffer(napi_env env, size_t byte_length, void **data, napi_value *result);
extern napi_status napi_create_external_arraybuffer(napi_env env, void *external_data, size_t byte_length, napi_finalize finalize_cb,
                                                    void *finalize_hint, napi_value *result);
extern napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value arraybuffer, size_t byte_offset,
                                          napi_value *result);
extern napi_status napi_throw_error(napi_env env, const char *code, const char *msg);
extern napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg);
extern napi_status napi_define_properties(napi_env env, napi_value object, size_t property_count, const napi_property_descriptor *properties);
extern napi_status napi_is_array(napi_env env, napi_value value, _Bool *result);
extern napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result);
extern napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result);
extern napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value);
extern napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result);

/* ---- helpers ---- */
static int get_bytes(napi_env env, napi_value v, const uint8_t **p, size_t *n) {
  _Bool is = 0;
  if (napi_is_typedarray(env, v, &is) != napi_ok || !is) return 0;
  napi_typedarray_type t;
  void *data = NULL;
  size_t len = 0, off = 0;
  napi_value ab;
  if (napi_get_typedarray_info(env, v, &t, &len, &data, &ab, &off) != napi_ok) return 0;
  if (t != napi_uint8_array && t != napi_uint8_clamped_array && t != napi_int8_array) return 0; /* Buffer is a Uint8Array */
  *p = (const uint8_t *)data;
  *n = len;
  return 1;
}

static napi_value throw_status(napi_env env, int rc) {
  /* the reference throws plain Error objects with these texts (src/zlib.ts:15, src/inflate.ts:32,35,50) */
  napi_throw_error(env, NULL, rc == ZLES_E_CUDA ? zles_last_cuda_error() : zles_strerror(rc));
  return NULL;
}

static napi_value make_u8(napi_env env, const uint8_t *src, size_t n) {
  void *dst = NULL;
  napi_value ab, out;
  if (napi_create_arraybuffer(env, n, &dst, &ab) != napi_ok) return NULL;
  for (size_t i = 0; i < n; i++) ((uint8_t *)dst)[i] = src[i];
  if (napi_create_typedarray(env, napi_uint8_array, n, ab, 0, &out) != napi_ok) return NULL;
  return out;
}

/* ---- deflate(input: Uint8Array): Uint8Array — replaces zlib.deflate, src/zlib.ts:25-49 ---- */
typedef int (*deflate_fn)(zles_ctx *, const uint8_t *, size_t, uint8_t *, size_t, size_t *);
static napi_value deflate_with(napi_env env, napi_callback_info info, deflate_fn fn) {
  size_t argc = 1;
  napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  const uint8_t *in;
  size_t n;
  if (argc < 1 || !get_bytes(env, argv[0], &in, &n)) { napi_throw_type_error(env, NULL, "deflate: expected a Uint8Array"); return NULL; }
  size_t cap = zles_deflate_bound(n) + 16, out_len = 0;
  void *dst = NULL;
  napi_value ab, out;
  if (napi_create_arraybuffer(env, cap, &dst, &ab) != napi_ok) return NULL;
  int rc = fn(NULL, in, n, (uint8_t *)dst, cap, &out_len);
  if (rc) return throw_status(env, rc);
  /* a view of exact length over the (slightly larger) buffer; the reference returns a fresh array of exact length */
  if (napi_create_typedarray(env, napi_uint8_array, out_len, ab, 0, &out) != napi_ok) return NULL;
  return out;
}
static napi_value js_deflate(napi_env env, napi_callback_info info) { return deflate_with(env, info, zles_deflate); }
/* the reference's deflate core (src/deflate.ts:14) and the gzip container (RFC 1952) */
static napi_value js_deflate_raw(napi_env env, napi_callback_info info) { return deflate_with(env, info, zles_deflate_raw); }
static napi_value js_gzip(napi_env env, napi_callback_info info) { return deflate_with(env, info, zles_gzip_deflate); }

/* inflateRaw / gunzip: the output size is not known in advance — the reference's first guess (10 x input, src/inflate.ts:17),
 * then once more with the size the library reports */
typedef int (*inflate_fn)(zles_ctx *, const uint8_t *, size_t, uint8_t *, size_t, size_t *);
static int inflate_raw0(zles_ctx *c, const uint8_t *in, size_t n, uint8_t *out, size_t cap, size_t *len) {
  return zles_inflate_raw(c, in, n, 0, out, cap, len);
}
static napi_value inflate_with(napi_env env, napi_callback_info info, inflate_fn fn) {
  size_t argc = 1;
  napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  const uint8_t *in;
  size_t n;
  if (argc < 1 || !get_bytes(env, argv[0], &in, &n)) { napi_throw_type_error(env, NULL, "inflate: expected a Uint8Array"); return NULL; }
  size_t cap = 10 * n + 131072, out_len = 0;
  for (int attempt = 0; attempt < 2; attempt++) {
    uint8_t *buf = (uint8_t *)malloc(cap ? cap : 1);
    if (!buf) return throw_status(env, ZLES_E_NOMEM);
    int rc = fn(NULL, in, n, buf, cap, &out_len);
    if (rc == 0) {
      napi_value out = make_u8(env, buf, out_len);
      free(buf);
      return out;
    }
    free(buf);
    if (rc != ZLES_E_OUTPUT_FULL || attempt) return throw_status(env, rc);
    cap = out_len;
  }
  return NULL;
}
static napi_value js_inflate_raw(napi_env env, napi_callback_info info) { return inflate_with(env, info, inflate_raw0); }
static napi_value js_gunzip(napi_env env, napi_callback_info info) { return inflate_with(env, info, zles_gzip_inflate); }

/* ---- inflate(input: Uint8Array): Uint8Array — replaces zlib.inflate, src/zlib.ts:11-23 ---- */
static void free_cb(napi_env env, void *data, void *hint) { (void)env; (void)hint; zles_free(data); }

static napi_value js_inflate(napi_env env, napi_callback_info info) {
  size_t argc = 1;
  napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  const uint8_t *in;
  size_t n;
  if (argc < 1 || !get_bytes(env, argv[0], &in, &n)) { napi_throw_type_error(env, NULL, "inflate: expected a Uint8Array"); return NULL; }
  uint8_t *buf = NULL;
  size_t out_len = 0;
  int rc = zles_inflate_alloc(NULL, in, n, &buf, &out_len);
  if (rc) return throw_status(env, rc);
  napi_value ab, out;
  /* like the reference, the result is a view onto a library-owned buffer (src/inflate.ts:39) */
  if (napi_create_external_arraybuffer(env, buf, out_len, free_cb, NULL, &ab) != napi_ok) {
    out = make_u8(env, buf, out_len); /* engines that forbid external buffers: copy */
    zles_free(buf);
    return out;
  }
  if (napi_create_typedarray(env, napi_uint8_array, out_len, ab, 0, &out) != napi_ok) return NULL;
  return out;
}

/* ---- batches: deflateBatch(bufs: Uint8Array[]): Uint8Array[] and inflateBatch(...) ---- */
static napi_value batch(napi_env env, napi_callback_info info, int inflate) {
  size_t argc = 1;
  napi_value argv[1];
  napi_get_cb_info(env, info, &argc, argv, NULL, NULL);
  _Bool is_arr = 0;
  if (argc < 1 || napi_is_array(env, argv[0], &is_arr) != napi_ok || !is_arr) { napi_throw_type_error(env, NULL, "expected an array of Uint8Array"); return NULL; }
  uint32_t count = 0;
  napi_get_array_length(env, argv[0], &count);
  uint64_t *in_off = (uint64_t *)calloc((size_t)count + 1, 8), *out_off = (uint64_t *)calloc((size_t)count + 1, 8);
  uint64_t *out_len = (uint64_t *)calloc((size_t)count + 1, 8);
  int32_t *status = (int32_t *)calloc((size_t)count + 1, 4);
  const uint8_t **ptr = (const uint8_t **)calloc((size_t)count + 1, sizeof(*ptr));
  napi_value result = NULL;
  uint8_t *in = NULL, *out = NULL;
  if (!in_off || !out_off || !out_len || !status || !ptr) goto done;
  for (uint32_t i = 0; i < count; i++) {
    napi_value e;
    size_t n;
    napi_get_element(env, argv[0], i, &e);
    if (!get_bytes(env, e, &ptr[i], &n)) { napi_throw_type_error(env, NULL, "expected an array of Uint8Array"); goto done; }
    in_off[i + 1] = in_off[i] + n;
    /* inflate: the reference's initial capacity is 10 x the input (src/inflate.ts:17); a stream that needs more is retried
     * on its own below, so no large constant per message is added (ten thousand 60-byte messages stay at ~6 MB) */
    out_off[i + 1] = out_off[i] + (inflate ? 10 * n + 64 : zles_deflate_bound(n));
  }
  in = (uint8_t *)malloc(in_off[count] ? in_off[count] : 1);
  out = (uint8_t *)malloc(out_off[count] ? out_off[count] : 1);
  if (!in || !out) goto done;
  for (uint32_t i = 0; i < count; i++)
    for (uint64_t k = 0; k < in_off[i + 1] - in_off[i]; k++) in[in_off[i] + k] = ptr[i][k];
  {
    int rc = inflate ? zles_inflate_batch(NULL, in, in_off, count, out, out_off, out_len, status)
                     : zles_deflate_batch(NULL, in, in_off, count, out, out_off, out_len, status);
    /* a failure of the call as a whole (no device, out of memory, a kernel fault) leaves status[] untouched: throw now */
    if (rc == ZLES_E_CUDA || rc == ZLES_E_ARG || rc == ZLES_E_NOMEM) { throw_status(env, rc); goto done; }
    if (rc) {
      /* per-buffer errors: the first one is thrown, like `inputs.map(inflate)` would; streams that only need more room
       * than the first guess fall back to the single-buffer call below */
      for (uint32_t i = 0; i < count; i++)
        if (status[i] && !(inflate && status[i] == ZLES_E_OUTPUT_FULL)) { throw_status(env, status[i]); goto done; }
    }
  }
  if (napi_create_array_with_length(env, count, &result) != napi_ok) { result = NULL; goto done; }
  for (uint32_t i = 0; i < count; i++) {
    napi_value v;
    if (inflate && status[i] == ZLES_E_OUTPUT_FULL) {
      uint8_t *buf = NULL;
      size_t len = 0;
      int rc = zles_inflate_alloc(NULL, ptr[i], (size_t)(in_off[i + 1] - in_off[i]), &buf, &len);
      if (rc) { throw_status(env, rc); result = NULL; goto done; }
      v = make_u8(env, buf, len);
      zles_free(buf);
    } else {
      v = make_u8(env, out + out_off[i], (size_t)out_len[i]);
    }
    if (!v) { result = NULL; goto done; }
    napi_set_element(env, result, i, v);
  }
done:
  free(in_off); free(out_off); free(out_len); free(status); free((void *)ptr); free(in); free(out);
  return result;
}
static napi_value js_deflate_batch(napi_env env, napi_callback_info info) { return batch(env, info, 0); }
static napi_value js_inflate_batch(napi_env env, napi_callback_info info) { return batch(env, info, 1); }

/* ZLES_DEVICES in the environment picks the GPUs the drop-in calls shard over: "all", or a comma-separated list of CUDA
 * device indices ("0,1,2,3").  Unset: device 0.  Large inputs are then sharded inside the library (zles_init, zles.h). */
static void init_devices(void) {
  const char *e = getenv("ZLES_DEVICES");
  uint32_t mask = 0;
  if (!e || !*e) return;
  if (e[0] == 'a') mask = 0xffffffffu;
  else
    for (const char *p = e; *p;) {
      char *end;
      long d = strtol(p, &end, 10);
      if (end == p) break;
      if (d >= 0 && d < 32) mask |= 1u << d;
      p = *end ? end + 1 : end;
    }
  if (mask) zles_init(mask);
}

/* module entry point looked up by Node (NAPI_MODULE_INIT expands to this symbol) */
napi_value napi_register_module_v1(napi_env env, napi_value exports) {
  init_devices();
  napi_property_descriptor props[] = {
      {"deflate", NULL, js_deflate, NULL, NULL, NULL, napi_default, NULL},
      {"inflate", NULL, js_inflate, NULL, NULL, NULL, napi_default, NULL},
      {"deflateBatch", NULL, js_deflate_batch, NULL, NULL, NULL, napi_default, NULL},
      {"inflateBatch", NULL, js_inflate_batch, NULL, NULL, NULL, napi_default, NULL},
      {"deflateRaw", NULL, js_deflate_raw, NULL, NULL, NULL, napi_default, NULL},
      {"inflateRaw", NULL, js_inflate_raw, NULL, NULL, NULL, napi_default, NULL},
      {"gzip", NULL, js_gzip, NULL, NULL, NULL, napi_default, NULL},
      {"gunzip", NULL, js_gunzip, NULL, NULL, NULL, napi_default, NULL},
  };
  napi_define_properties(env, exports, sizeof(props) / sizeof(props[0]), props);
  return exports;
}
