/*
 * tmjx_xla_ffi.cc — XLA-FFI custom-call adapter over the C ABI (include/tmjx.h).
 *
 * The reference's env lives inside `jax.jit` / `lax.scan` / `pmap` (reference track_mjx/agent/mlp_ppo/ppo.py:333-340, 409);
 * a JAX host therefore calls the step as an XLA custom call.  This translation unit is compiled ONLY when jaxlib's
 * `xla/ffi/api/ffi.h` is on the include path (`__graft_entry__.build()` probes for it; this image has no jaxlib, so
 * here it is compile-gated and INTEGRATION.md shows the Python side).  It adds no arithmetic: it unpacks XLA buffers
 * into TmjxState / TmjxOut in the field order of include/tmjx.h and forwards to tmjx_step / tmjx_forward on XLA's
 * stream.  XLA owns every buffer; state leaves are donated and aliased to the outputs by the caller
 * (`input_output_aliases`), so operands and results of a leaf are the same device pointer.
 *
 * Operand order : action, then every TmjxState pointer member in declaration order (25 leaves).
 * Result order  : the 25 state leaves (aliased), then obs, reward, done, metrics, cur_frame.
 * Attributes    : model (int64 handle from tmjx_model_create), clips (int64 handle), flags (int64, TMJX_F_*).
 */
#if __has_include("xla/ffi/api/ffi.h")
#include <cuda_runtime.h>

#include <cstdint>

#include "../../include/tmjx.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

constexpr int kStateLeaves = sizeof(TmjxState) / sizeof(void*);
constexpr int kOutLeaves = 5;  // obs, reward, done, metrics, cur_frame (debug taps are not exposed through XLA)

ffi::Error Unpack(ffi::RemainingRets& rets, TmjxState* s, TmjxOut* o, int* n_env) {
  if (rets.size() != size_t(kStateLeaves + kOutLeaves))
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, "tmjx: expected 25 state leaves + 5 outputs as results");
  void** sp = reinterpret_cast<void**>(s);
  for (int i = 0; i < kStateLeaves; ++i) {
    auto b = rets.get<ffi::AnyBuffer>(i);
    if (!b.has_value()) return b.error();
    sp[i] = (*b)->untyped_data();
    if (i == 0) *n_env = int((*b)->dimensions()[0]);
  }
  void* out[kOutLeaves];
  for (int i = 0; i < kOutLeaves; ++i) {
    auto b = rets.get<ffi::AnyBuffer>(kStateLeaves + i);
    if (!b.has_value()) return b.error();
    out[i] = (*b)->untyped_data();
  }
  *o = TmjxOut{};
  o->obs = static_cast<float*>(out[0]);
  o->reward = static_cast<float*>(out[1]);
  o->done = static_cast<float*>(out[2]);
  o->metrics = static_cast<float*>(out[3]);
  o->cur_frame = static_cast<int32_t*>(out[4]);
  return ffi::Error::Success();
}

ffi::Error StepImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> action, ffi::RemainingArgs /*state leaves, aliased to rets*/,
                    ffi::RemainingRets rets, int64_t model, int64_t clips, int64_t flags) {
  TmjxState s;
  TmjxOut o;
  int n_env = 0;
  if (auto e = Unpack(rets, &s, &o, &n_env); e.failure()) return e;
  const int rc = tmjx_step(reinterpret_cast<const TmjxModel*>(model), reinterpret_cast<const TmjxClips*>(clips),
                           action.typed_data(), &s, &o, n_env, unsigned(flags), stream);
  return rc == TMJX_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, tmjx_last_error());
}

ffi::Error ForwardImpl(cudaStream_t stream, ffi::RemainingArgs, ffi::RemainingRets rets, int64_t model, int64_t clips,
                       int64_t flags) {
  TmjxState s;
  TmjxOut o;
  int n_env = 0;
  if (auto e = Unpack(rets, &s, &o, &n_env); e.failure()) return e;
  const int rc = tmjx_forward(reinterpret_cast<const TmjxModel*>(model), reinterpret_cast<const TmjxClips*>(clips), &s, &o,
                              n_env, unsigned(flags), stream);
  return rc == TMJX_OK ? ffi::Error::Success() : ffi::Error(ffi::ErrorCode::kInternal, tmjx_last_error());
}

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(tmjx_step_ffi, StepImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("model")
                                  .Attr<int64_t>("clips")
                                  .Attr<int64_t>("flags"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(tmjx_forward_ffi, ForwardImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .RemainingArgs()
                                  .RemainingRets()
                                  .Attr<int64_t>("model")
                                  .Attr<int64_t>("clips")
                                  .Attr<int64_t>("flags"));
#else
/* jaxlib headers not present: nothing to build (the ctypes binding in track-mjx_b200/_lib.py is the boundary here). */
extern "C" int tmjx_xla_ffi_available(void) { return 0; }
#endif
