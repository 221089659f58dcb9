/*
 * tmjx_xla_ffi.cc — XLA-FFI custom-call adapter over the C ABI (include/tmjx.h), written against XLA's FFI *C* API.
 *
 * The reference's env lives inside `jax.jit` / `lax.scan` / `pmap` (reference track_mjx/agent/mlp_ppo/ppo.py:333-340, 409);
 * a JAX host therefore calls the step as an XLA custom call (`jax.ffi.register_ffi_target` + `jax.ffi.ffi_call`).
 * The handlers add no arithmetic: they decode the call frame XLA hands them (operands, results, attributes, the stream) into
 * TmjxState / TmjxOut in the field order of include/tmjx.h and forward to tmjx_step / tmjx_forward on XLA's stream.  XLA owns every
 * buffer; state leaves are donated and aliased to the results by the caller (`input_output_aliases`), so the state is read from and
 * written to the RESULT buffers.
 *
 *   handler symbols : tmjx_step_ffi, tmjx_forward_ffi        (XLA_FFI_Handler: XLA_FFI_Error* (XLA_FFI_CallFrame*))
 *   operands        : tmjx_step_ffi: action [n_env, nu] f32, then the 25 TmjxState leaves in declaration order (donated)
 *                     tmjx_forward_ffi: the 25 TmjxState leaves
 *   results         : the 25 state leaves (aliased to the operands), then obs, reward, done, metrics (f32), cur_frame (s32)
 *   attributes      : model, clips (int64 handles from tmjx_model_create / tmjx_clips_create), flags (int64, TMJX_F_*)
 *
 * Header: jaxlib's `xla/ffi/api/c_api.h` when it is on the include path, otherwise the recalled subset in
 * tmjx_xla_ffi_c_api_min.h (this image has no jaxlib; see that file's caveat).  `tmjx_ffi_selftest` builds a call frame by hand and
 * drives a handler through it -- the "XLA side" of the contract for tests/test_gpu_ffi.py.
 */
#include <cuda_runtime.h>

#include <cstdint>
#include <cstring>
#include <string>

#include "../../include/tmjx.h"
#if __has_include("xla/ffi/api/c_api.h")
#include "xla/ffi/api/c_api.h"
#define TMJX_FFI_REAL_HEADER 1
#else
#include "tmjx_xla_ffi_c_api_min.h"
#define TMJX_FFI_REAL_HEADER 0
#endif

namespace {

constexpr int kStateLeaves = sizeof(TmjxState) / sizeof(void*);
constexpr int kOutLeaves = 5;  // obs, reward, done, metrics, cur_frame (debug taps are not exposed through XLA)

XLA_FFI_Error* make_error(const XLA_FFI_Api* api, XLA_FFI_Error_Code code, const char* msg) {
  XLA_FFI_Error_Create_Args a;
  std::memset(&a, 0, sizeof(a));
  a.struct_size = sizeof(a);
  a.message = msg;
  a.errc = code;
  return api->XLA_FFI_Error_Create(&a);
}

bool attr_i64(const XLA_FFI_Attrs& attrs, const char* name, int64_t* out) {
  const size_t len = std::strlen(name);
  for (int64_t i = 0; i < attrs.size; ++i) {
    const XLA_FFI_ByteSpan* nm = attrs.names[i];
    if (nm->len != len || std::memcmp(nm->ptr, name, len) != 0) continue;
    if (attrs.types[i] != XLA_FFI_AttrType_SCALAR) return false;
    const XLA_FFI_Scalar* sc = static_cast<const XLA_FFI_Scalar*>(attrs.attr[i]);
    if (sc->dtype == XLA_FFI_DataType_S64) { *out = *static_cast<const int64_t*>(sc->value); return true; }
    if (sc->dtype == XLA_FFI_DataType_S32) { *out = *static_cast<const int32_t*>(sc->value); return true; }
    if (sc->dtype == XLA_FFI_DataType_U32) { *out = *static_cast<const uint32_t*>(sc->value); return true; }
    return false;
  }
  return false;
}

// XLA asks a handler for its metadata by calling it with a metadata extension attached to the frame
bool answer_metadata(XLA_FFI_CallFrame* cf) {
  for (XLA_FFI_Extension_Base* e = cf->extension_start; e; e = e->next) {
    if (e->type != XLA_FFI_Extension_Metadata) continue;
    XLA_FFI_Metadata* md = reinterpret_cast<XLA_FFI_Metadata_Extension*>(e)->metadata;
    md->api_version.major_version = XLA_FFI_API_MAJOR;
    md->api_version.minor_version = XLA_FFI_API_MINOR;
    md->traits = 0;
    return true;
  }
  return false;
}

XLA_FFI_Error* run(XLA_FFI_CallFrame* cf, bool is_step) {
  if (answer_metadata(cf)) return nullptr;
  if (cf->stage != XLA_FFI_ExecutionStage_EXECUTE) return nullptr;   // nothing to instantiate / prepare / initialise
  const XLA_FFI_Api* api = cf->api;
  const int n_args = kStateLeaves + (is_step ? 1 : 0);
  if (cf->args.size != n_args) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: wrong number of operands (action + 25 state leaves)");
  if (cf->rets.size != kStateLeaves + kOutLeaves)
    return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: expected 25 state leaves + 5 outputs as results");
  for (int64_t i = 0; i < cf->args.size; ++i)
    if (cf->args.types[i] != XLA_FFI_ArgType_BUFFER) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: operands must be buffers");
  for (int64_t i = 0; i < cf->rets.size; ++i)
    if (cf->rets.types[i] != XLA_FFI_RetType_BUFFER) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: results must be buffers");
  int64_t model = 0, clips = 0, flags = 0;
  if (!attr_i64(cf->attrs, "model", &model) || !attr_i64(cf->attrs, "clips", &clips) || !attr_i64(cf->attrs, "flags", &flags))
    return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: attributes model, clips, flags (int64 scalars) are required");

  TmjxState s;
  TmjxOut o;
  std::memset(&o, 0, sizeof(o));
  void** sp = reinterpret_cast<void**>(&s);
  int n_env = 0;
  for (int i = 0; i < kStateLeaves; ++i) {
    const XLA_FFI_Buffer* b = static_cast<const XLA_FFI_Buffer*>(cf->rets.rets[i]);
    const XLA_FFI_Buffer* in = static_cast<const XLA_FFI_Buffer*>(cf->args.args[i + (is_step ? 1 : 0)]);
    if (b->data != in->data)
      return make_error(api, XLA_FFI_Error_Code_FAILED_PRECONDITION, "tmjx: state leaves must be donated (input_output_aliases={i + 1: i})");
    sp[i] = b->data;
    if (i == 0) {
      if (b->rank < 1) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: state leaves are [n_env, dim]");
      n_env = int(b->dims[0]);
    }
  }
  const XLA_FFI_Buffer* ob[kOutLeaves];
  for (int i = 0; i < kOutLeaves; ++i) ob[i] = static_cast<const XLA_FFI_Buffer*>(cf->rets.rets[kStateLeaves + i]);
  for (int i = 0; i < 4; ++i)
    if (ob[i]->dtype != XLA_FFI_DataType_F32) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: obs / reward / done / metrics are f32");
  if (ob[4]->dtype != XLA_FFI_DataType_S32) return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: cur_frame is s32");
  o.obs = static_cast<float*>(ob[0]->data);
  o.reward = static_cast<float*>(ob[1]->data);
  o.done = static_cast<float*>(ob[2]->data);
  o.metrics = static_cast<float*>(ob[3]->data);
  o.cur_frame = static_cast<int32_t*>(ob[4]->data);

  XLA_FFI_Stream_Get_Args sg;
  std::memset(&sg, 0, sizeof(sg));
  sg.struct_size = sizeof(sg);
  sg.ctx = cf->ctx;
  if (XLA_FFI_Error* e = api->XLA_FFI_Stream_Get(&sg)) return e;

  int rc;
  if (is_step) {
    const XLA_FFI_Buffer* act = static_cast<const XLA_FFI_Buffer*>(cf->args.args[0]);
    if (act->dtype != XLA_FFI_DataType_F32 || act->rank != 2 || act->dims[0] != n_env)
      return make_error(api, XLA_FFI_Error_Code_INVALID_ARGUMENT, "tmjx: action must be f32 [n_env, nu]");
    rc = tmjx_step(reinterpret_cast<const TmjxModel*>(model), reinterpret_cast<const TmjxClips*>(clips), static_cast<const float*>(act->data), &s, &o,
                   n_env, unsigned(flags), sg.stream);
  } else {
    rc = tmjx_forward(reinterpret_cast<const TmjxModel*>(model), reinterpret_cast<const TmjxClips*>(clips), &s, &o, n_env, unsigned(flags), sg.stream);
  }
  return rc == TMJX_OK ? nullptr : make_error(api, XLA_FFI_Error_Code_INTERNAL, tmjx_last_error());
}

// ---- the "XLA side" for the self-test: an XLA_FFI_Api with the two entry points the adapter uses
struct TestError { std::string msg; int code; };
thread_local void* g_test_stream = nullptr;
thread_local TestError g_test_error;
XLA_FFI_Error* test_error_create(XLA_FFI_Error_Create_Args* a) {
  g_test_error.msg = a->message ? a->message : "";
  g_test_error.code = int(a->errc);
  return reinterpret_cast<XLA_FFI_Error*>(&g_test_error);
}
XLA_FFI_Error* test_stream_get(XLA_FFI_Stream_Get_Args* a) {
  a->stream = g_test_stream;
  return nullptr;
}

}  // namespace

extern "C" {

XLA_FFI_Error* tmjx_step_ffi(XLA_FFI_CallFrame* call_frame) { return run(call_frame, true); }
XLA_FFI_Error* tmjx_forward_ffi(XLA_FFI_CallFrame* call_frame) { return run(call_frame, false); }

/* 1: built against jaxlib's own c_api.h; 0: against the recalled subset (tmjx_xla_ffi_c_api_min.h) */
int tmjx_xla_ffi_available(void) { return TMJX_FFI_REAL_HEADER; }

/* Build an XLA_FFI_CallFrame for one call by hand -- what XLA's custom-call thunk does -- and run a handler through it.
 * dims: per state leaf and per output its trailing dimension (leaves are [n_env, dim]); leaf_is_int: 1 for s32 leaves.
 * `alias` = 0 breaks the donation (operand pointers differ from result pointers) to exercise the error path.
 * Returns 0 on success, else the XLA error code; the message is available through tmjx_ffi_selftest_error(). */
int tmjx_ffi_selftest(int is_step, const void* model, const void* clips, const float* action, int nu, const TmjxState* s, const TmjxOut* o,
                      const int* state_dims, const int* state_is_int, const int* out_dims, int n_env, unsigned flags, void* stream, int alias) {
  XLA_FFI_Api api;
  std::memset(&api, 0, sizeof(api));
  api.struct_size = sizeof(api);
  api.api_version.struct_size = sizeof(api.api_version);
  api.api_version.major_version = XLA_FFI_API_MAJOR;
  api.api_version.minor_version = XLA_FFI_API_MINOR;
  api.XLA_FFI_Error_Create = test_error_create;
  api.XLA_FFI_Stream_Get = test_stream_get;
  g_test_stream = stream;
  g_test_error = TestError{"", 0};

  const int n_args = kStateLeaves + (is_step ? 1 : 0), n_rets = kStateLeaves + kOutLeaves;
  XLA_FFI_Buffer arg_buf[kStateLeaves + 1], ret_buf[kStateLeaves + kOutLeaves];
  int64_t arg_dims[kStateLeaves + 1][2], ret_dims[kStateLeaves + kOutLeaves][2];
  XLA_FFI_ArgType arg_types[kStateLeaves + 1];
  XLA_FFI_RetType ret_types[kStateLeaves + kOutLeaves];
  void* arg_ptrs[kStateLeaves + 1];
  void* ret_ptrs[kStateLeaves + kOutLeaves];
  auto fill = [](XLA_FFI_Buffer& b, int64_t* dims, void* data, XLA_FFI_DataType dt, int64_t n, int64_t d) {
    std::memset(&b, 0, sizeof(b));
    b.struct_size = sizeof(b);
    b.dtype = dt; b.data = data; b.rank = 2; dims[0] = n; dims[1] = d; b.dims = dims;
  };
  void* const* sp = reinterpret_cast<void* const*>(s);
  int a0 = 0;
  if (is_step) {
    fill(arg_buf[0], arg_dims[0], const_cast<float*>(action), XLA_FFI_DataType_F32, n_env, nu);
    a0 = 1;
  }
  static char dummy[16];
  for (int i = 0; i < kStateLeaves; ++i) {
    const XLA_FFI_DataType dt = state_is_int[i] ? XLA_FFI_DataType_S32 : XLA_FFI_DataType_F32;
    fill(ret_buf[i], ret_dims[i], sp[i], dt, n_env, state_dims[i]);
    fill(arg_buf[a0 + i], arg_dims[a0 + i], alias ? sp[i] : static_cast<void*>(dummy), dt, n_env, state_dims[i]);
  }
  void* outs[kOutLeaves] = {o->obs, o->reward, o->done, o->metrics, o->cur_frame};
  for (int i = 0; i < kOutLeaves; ++i)
    fill(ret_buf[kStateLeaves + i], ret_dims[kStateLeaves + i], outs[i], i == 4 ? XLA_FFI_DataType_S32 : XLA_FFI_DataType_F32, n_env, out_dims[i]);
  for (int i = 0; i < n_args; ++i) { arg_types[i] = XLA_FFI_ArgType_BUFFER; arg_ptrs[i] = &arg_buf[i]; }
  for (int i = 0; i < n_rets; ++i) { ret_types[i] = XLA_FFI_RetType_BUFFER; ret_ptrs[i] = &ret_buf[i]; }

  // attributes, sorted by name as XLA encodes a dictionary: clips, flags, model
  int64_t v_clips = reinterpret_cast<int64_t>(clips), v_flags = int64_t(flags), v_model = reinterpret_cast<int64_t>(model);
  XLA_FFI_Scalar sc[3] = {{XLA_FFI_DataType_S64, &v_clips}, {XLA_FFI_DataType_S64, &v_flags}, {XLA_FFI_DataType_S64, &v_model}};
  XLA_FFI_ByteSpan nm[3] = {{"clips", 5}, {"flags", 5}, {"model", 5}};
  XLA_FFI_ByteSpan* nmp[3] = {&nm[0], &nm[1], &nm[2]};
  XLA_FFI_AttrType at[3] = {XLA_FFI_AttrType_SCALAR, XLA_FFI_AttrType_SCALAR, XLA_FFI_AttrType_SCALAR};
  void* attr_ptrs[3] = {&sc[0], &sc[1], &sc[2]};

  XLA_FFI_CallFrame cf;
  std::memset(&cf, 0, sizeof(cf));
  cf.struct_size = sizeof(cf);
  cf.api = &api;
  cf.stage = XLA_FFI_ExecutionStage_EXECUTE;
  cf.args.struct_size = sizeof(cf.args); cf.args.size = n_args; cf.args.types = arg_types; cf.args.args = arg_ptrs;
  cf.rets.struct_size = sizeof(cf.rets); cf.rets.size = n_rets; cf.rets.types = ret_types; cf.rets.rets = ret_ptrs;
  cf.attrs.struct_size = sizeof(cf.attrs); cf.attrs.size = 3; cf.attrs.types = at; cf.attrs.names = nmp; cf.attrs.attr = attr_ptrs;

  // 1. the metadata probe XLA issues at registration time must be answered without touching the operands
  {
    XLA_FFI_Metadata md;
    std::memset(&md, 0, sizeof(md));
    md.struct_size = sizeof(md);
    XLA_FFI_Metadata_Extension ext;
    std::memset(&ext, 0, sizeof(ext));
    ext.extension_base.struct_size = sizeof(ext);
    ext.extension_base.type = XLA_FFI_Extension_Metadata;
    ext.metadata = &md;
    XLA_FFI_CallFrame probe = cf;
    probe.extension_start = &ext.extension_base;
    if ((is_step ? tmjx_step_ffi(&probe) : tmjx_forward_ffi(&probe)) != nullptr) return -100;
    if (md.api_version.major_version != XLA_FFI_API_MAJOR || md.api_version.minor_version != XLA_FFI_API_MINOR) return -101;
  }
  // 2. the non-execute stages are no-ops
  {
    XLA_FFI_CallFrame prep = cf;
    prep.stage = XLA_FFI_ExecutionStage_PREPARE;
    if ((is_step ? tmjx_step_ffi(&prep) : tmjx_forward_ffi(&prep)) != nullptr) return -102;
  }
  // 3. execute
  XLA_FFI_Error* e = is_step ? tmjx_step_ffi(&cf) : tmjx_forward_ffi(&cf);
  return e ? (g_test_error.code ? g_test_error.code : -1) : 0;
}

const char* tmjx_ffi_selftest_error(void) { return g_test_error.msg.c_str(); }

}  // extern "C"
