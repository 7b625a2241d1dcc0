/*
 * napi_mock.c — a small stand-in for the part of Node's N-API that zlib.es_b200/node/addon.c uses.
 *
 * TEST INFRASTRUCTURE ONLY.  The image has no Node.js, so the addon could only be compile-checked; linked against this
 * mock (instead of the node binary) its code actually RUNS: argument checks, the calls into libzles, the mapping of
 * status codes to thrown Errors, the batch paths.  tests/test_addon_mock.py drives it through ctypes — with the CPU
 * emulator build of the library in this container, with libzles.so on a GPU box.
 *
 * Values are plain C structs; there is no garbage collector (the tests are short-lived), finalizers of external
 * array buffers run when a result is copied out.
 */
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

typedef struct napi_env__ *napi_env;
typedef struct napi_value__ *napi_value;
typedef struct napi_callback_info__ *napi_callback_info;
typedef enum { napi_ok = 0, napi_generic_failure = 9 } napi_status;
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

enum { V_ARRAYBUFFER = 1, V_TYPEDARRAY, V_ARRAY, V_OBJECT };
typedef struct Val {
  int kind;
  uint8_t *data;        /* arraybuffer / typedarray bytes */
  size_t len;
  napi_typedarray_type ttype;
  struct Val *ab;       /* typedarray -> its buffer */
  napi_finalize fin;    /* external arraybuffer */
  void *fin_hint;
  int owned;            /* arraybuffer allocated here */
  struct Val **items;   /* array */
  size_t nitems;
} Val;

struct napi_env__ {
  int pending;          /* an exception is pending */
  int type_error;
  char msg[512];
};
struct napi_callback_info__ {
  size_t argc;
  Val **argv;
};

static struct napi_env__ g_env;
static napi_property_descriptor g_props[32];
static size_t g_nprops;

static Val *new_val(int kind) {
  Val *v = (Val *)calloc(1, sizeof(Val));
  v->kind = kind;
  return v;
}

napi_status napi_get_cb_info(napi_env env, napi_callback_info info, size_t *argc, napi_value *argv, napi_value *this_arg, void **data) {
  (void)env; (void)this_arg; (void)data;
  size_t want = *argc;
  for (size_t i = 0; i < want; i++) argv[i] = i < info->argc ? (napi_value)info->argv[i] : NULL;
  *argc = info->argc;
  return napi_ok;
}
napi_status napi_is_typedarray(napi_env env, napi_value value, _Bool *result) {
  (void)env;
  *result = value && ((Val *)value)->kind == V_TYPEDARRAY;
  return napi_ok;
}
napi_status napi_get_typedarray_info(napi_env env, napi_value ta, napi_typedarray_type *type, size_t *length, void **data, napi_value *ab, size_t *off) {
  (void)env;
  Val *v = (Val *)ta;
  if (!v || v->kind != V_TYPEDARRAY) return napi_generic_failure;
  if (type) *type = v->ttype;
  if (length) *length = v->len;
  if (data) *data = v->data;
  if (ab) *ab = (napi_value)v->ab;
  if (off) *off = 0;
  return napi_ok;
}
napi_status napi_create_arraybuffer(napi_env env, size_t n, void **data, napi_value *result) {
  (void)env;
  Val *v = new_val(V_ARRAYBUFFER);
  v->data = (uint8_t *)calloc(n ? n : 1, 1);  /* a fresh ArrayBuffer is zero-filled */
  v->len = n;
  v->owned = 1;
  if (data) *data = v->data;
  *result = (napi_value)v;
  return napi_ok;
}
napi_status napi_create_external_arraybuffer(napi_env env, void *ext, size_t n, napi_finalize fin, void *hint, napi_value *result) {
  (void)env;
  if (getenv("NAPI_MOCK_NO_EXTERNAL")) return napi_generic_failure;  /* engines that forbid external buffers (the addon then copies) */
  Val *v = new_val(V_ARRAYBUFFER);
  v->data = (uint8_t *)ext;
  v->len = n;
  v->fin = fin;
  v->fin_hint = hint;
  *result = (napi_value)v;
  return napi_ok;
}
napi_status napi_create_typedarray(napi_env env, napi_typedarray_type type, size_t length, napi_value ab, size_t off, napi_value *result) {
  (void)env;
  Val *b = (Val *)ab;
  if (!b || b->kind != V_ARRAYBUFFER || off + length > b->len) return napi_generic_failure;
  Val *v = new_val(V_TYPEDARRAY);
  v->ttype = type;
  v->data = b->data + off;
  v->len = length;
  v->ab = b;
  *result = (napi_value)v;
  return napi_ok;
}
static napi_status throw_(napi_env env, const char *msg, int type_error) {
  env->pending = 1;
  env->type_error = type_error;
  snprintf(env->msg, sizeof(env->msg), "%s", msg ? msg : "");
  return napi_ok;
}
napi_status napi_throw_error(napi_env env, const char *code, const char *msg) { (void)code; return throw_(env, msg, 0); }
napi_status napi_throw_type_error(napi_env env, const char *code, const char *msg) { (void)code; return throw_(env, msg, 1); }
napi_status napi_define_properties(napi_env env, napi_value object, size_t n, const napi_property_descriptor *props) {
  (void)env; (void)object;
  for (size_t i = 0; i < n && g_nprops < 32; i++) g_props[g_nprops++] = props[i];
  return napi_ok;
}
napi_status napi_is_array(napi_env env, napi_value value, _Bool *result) {
  (void)env;
  *result = value && ((Val *)value)->kind == V_ARRAY;
  return napi_ok;
}
napi_status napi_get_array_length(napi_env env, napi_value value, uint32_t *result) {
  (void)env;
  *result = (uint32_t)((Val *)value)->nitems;
  return napi_ok;
}
napi_status napi_get_element(napi_env env, napi_value object, uint32_t index, napi_value *result) {
  (void)env;
  Val *a = (Val *)object;
  *result = index < a->nitems ? (napi_value)a->items[index] : NULL;
  return napi_ok;
}
napi_status napi_set_element(napi_env env, napi_value object, uint32_t index, napi_value value) {
  (void)env;
  Val *a = (Val *)object;
  if (index >= a->nitems) return napi_generic_failure;
  a->items[index] = (Val *)value;
  return napi_ok;
}
napi_status napi_create_array_with_length(napi_env env, size_t length, napi_value *result) {
  (void)env;
  Val *a = new_val(V_ARRAY);
  a->items = (Val **)calloc(length ? length : 1, sizeof(Val *));
  a->nitems = length;
  *result = (napi_value)a;
  return napi_ok;
}

/* ---- the driver the Python test calls ---------------------------------------------------------------------- */
extern napi_value napi_register_module_v1(napi_env env, napi_value exports);

static napi_callback find(const char *name) {
  for (size_t i = 0; i < g_nprops; i++)
    if (strcmp(g_props[i].utf8name, name) == 0) return g_props[i].method;
  return NULL;
}
static Val *make_u8(const uint8_t *p, size_t n) {
  Val *ab = new_val(V_ARRAYBUFFER);
  ab->data = (uint8_t *)malloc(n ? n : 1);
  if (n) memcpy(ab->data, p, n);
  ab->len = n;
  ab->owned = 1;
  Val *ta = new_val(V_TYPEDARRAY);
  ta->ttype = napi_uint8_array;
  ta->data = ab->data;
  ta->len = n;
  ta->ab = ab;
  return ta;
}
static uint8_t *take_bytes(Val *ta, size_t *n) {  /* copies the result out and releases an external buffer */
  uint8_t *out = (uint8_t *)malloc(ta->len ? ta->len : 1);
  if (ta->len) memcpy(out, ta->data, ta->len);
  *n = ta->len;
  if (ta->ab && ta->ab->fin) { ta->ab->fin(&g_env, ta->ab->data, ta->ab->fin_hint); ta->ab->fin = NULL; }
  return out;
}

int mock_load(void) {
  g_nprops = 0;
  memset(&g_env, 0, sizeof(g_env));
  Val *exports = new_val(V_OBJECT);
  napi_register_module_v1(&g_env, (napi_value)exports);
  return (int)g_nprops;
}
const char *mock_export_name(int i) { return i >= 0 && (size_t)i < g_nprops ? g_props[i].utf8name : NULL; }

/* fn(Uint8Array | junk) -> Uint8Array.  Returns 0 = ok, 1 = Error thrown, 2 = TypeError thrown, -1 = no such export.
 * kind_of_arg: 0 = a Uint8Array of the given bytes, 1 = a non-array value, 2 = no argument at all */
int mock_call1(const char *name, int kind_of_arg, const uint8_t *in, size_t n, uint8_t **out, size_t *out_len, char *err, size_t errcap) {
  napi_callback fn = find(name);
  if (!fn) return -1;
  Val *arg = kind_of_arg == 0 ? make_u8(in, n) : new_val(V_OBJECT);
  Val *argv[1] = {arg};
  struct napi_callback_info__ info = {kind_of_arg == 2 ? 0u : 1u, argv};
  g_env.pending = 0;
  napi_value r = fn(&g_env, &info);
  if (g_env.pending) {
    snprintf(err, errcap, "%s", g_env.msg);
    return g_env.type_error ? 2 : 1;
  }
  if (!r || ((Val *)r)->kind != V_TYPEDARRAY) { snprintf(err, errcap, "no result"); return 3; }
  *out = take_bytes((Val *)r, out_len);
  return 0;
}

/* fn(Uint8Array[]) -> Uint8Array[]; the results come back concatenated (lens[i] each) */
int mock_call_batch(const char *name, const uint8_t *blob, const size_t *lens, uint32_t count, uint8_t **out, size_t *out_lens, char *err,
                    size_t errcap) {
  napi_callback fn = find(name);
  if (!fn) return -1;
  Val *arr = new_val(V_ARRAY);
  arr->items = (Val **)calloc(count ? count : 1, sizeof(Val *));
  arr->nitems = count;
  size_t o = 0;
  for (uint32_t i = 0; i < count; i++) { arr->items[i] = make_u8(blob + o, lens[i]); o += lens[i]; }
  Val *argv[1] = {arr};
  struct napi_callback_info__ info = {1, argv};
  g_env.pending = 0;
  napi_value r = fn(&g_env, &info);
  if (g_env.pending) {
    snprintf(err, errcap, "%s", g_env.msg);
    return g_env.type_error ? 2 : 1;
  }
  Val *res = (Val *)r;
  if (!res || res->kind != V_ARRAY || res->nitems != count) { snprintf(err, errcap, "no result"); return 3; }
  size_t total = 0;
  for (uint32_t i = 0; i < count; i++) total += res->items[i] ? res->items[i]->len : 0;
  uint8_t *buf = (uint8_t *)malloc(total ? total : 1);
  o = 0;
  for (uint32_t i = 0; i < count; i++) {
    Val *e = res->items[i];
    out_lens[i] = e ? e->len : 0;
    if (e && e->len) memcpy(buf + o, e->data, e->len);
    o += out_lens[i];
  }
  *out = buf;
  return 0;
}
void mock_free(void *p) { free(p); }
