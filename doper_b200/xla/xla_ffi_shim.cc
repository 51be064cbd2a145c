// XLA FFI shim over the doper_b200 C ABI  —  UNTESTED IN THIS REPOSITORY.
//
// BASELINE.json's north_star asks for a jax.ffi / XLA custom-call boundary wrapped in
// jax.custom_vjp.  Neither jax nor jaxlib (nor the xla/ffi/api headers) exist in the build
// image or on the GPU boxes, and there is no network, so this file has never been compiled or
// run.  It is written against the public XLA FFI API (xla/ffi/api/ffi.h as shipped by
// jax.ffi.include_dir()) and is compile-gated on that header; doper_b200/xla/jax_binding.py
// builds it on demand when `import jax` succeeds.  Everything that is tested and measured in
// this repository goes through the same C ABI (include/doper_b200.h) via the torch/ctypes host
// binding; this shim adds no compute.
//
// Handlers (all launch on XLA's stream, allocate nothing, never synchronise):
//   doper_scene_prepare_ffi   segments f32[n_scenes,S,2,2] (attr max_radius) -> scene_ws u8[W]
//   doper_rollout_fwd_ffi     scene_ws, x0, v0 f32[B,2]               -> xT, vT f32[B,2], ws u8[R]
//   doper_rollout_bwd_ffi     scene_ws, ws, xT_bar, vT_bar f32[B,2]   -> x0_bar, v0_bar f32[B,2], scratch u8[Q]
//                             (ws is an immutable XLA input: everything the backward writes goes to its own
//                             result buffers; `scratch` is sized by doper_rollout_bwd_scratch_bytes and dropped)
// Attributes: S, per_env, n_steps, ckpt_every (i32), dt, radius, mass, friction, elasticity,
// attractor_strength, attractor_x, attractor_y (f32).
#if defined(__has_include)
#if __has_include("xla/ffi/api/ffi.h")
#define DOPER_HAVE_XLA_FFI 1
#endif
#endif

#ifdef DOPER_HAVE_XLA_FFI
#include <cuda_runtime.h>

#include <string>

#include "../../include/doper_b200.h"
#include "xla/ffi/api/ffi.h"

namespace ffi = xla::ffi;

namespace {

ffi::Error Status(int code, const char* where) {
    if (code == DOPER_OK) return ffi::Error::Success();
    return ffi::Error(ffi::ErrorCode::kInvalidArgument, std::string(where) + ": " + doper_error_string(code));
}

doper_rollout_desc MakeDesc(int64_t B, int32_t S, int32_t per_env, int32_t n_steps, int32_t ckpt_every, float dt,
                            float radius, float mass, float friction, float elasticity, float ks, float ax,
                            float ay) {
    doper_rollout_desc d{};
    d.B = B; d.S = S; d.per_env = per_env; d.n_steps = n_steps; d.ckpt_every = ckpt_every; d.dt = dt;
    d.consts[0] = radius; d.consts[1] = mass; d.consts[2] = friction; d.consts[3] = elasticity; d.consts[4] = ks;
    d.attractor[0] = ax; d.attractor[1] = ay;
    d.lanes_per_ball = 0; d.flags = 0; d.hit_capacity = 0;
    return d;
}

ffi::Error ScenePrepareImpl(cudaStream_t stream, ffi::Buffer<ffi::F32> segments,
                            ffi::ResultBuffer<ffi::U8> scene_ws, float max_radius) {
    auto dims = segments.dimensions();  // [S,2,2] or [E,S,2,2]
    const bool per_env = dims.size() == 4;
    const int64_t n = per_env ? dims[0] : 1;
    const int32_t S = static_cast<int32_t>(per_env ? dims[1] : dims[0]);
    if (scene_ws->element_count() < doper_scene_workspace_bytes(n, S))
        return Status(DOPER_ERR_WORKSPACE, "doper_scene_prepare");
    return Status(doper_scene_prepare(stream, n, S, segments.typed_data(), max_radius, scene_ws->typed_data()),
                  "doper_scene_prepare");
}

ffi::Error RolloutFwdImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> scene_ws, ffi::Buffer<ffi::F32> x0,
                          ffi::Buffer<ffi::F32> v0, ffi::ResultBuffer<ffi::F32> xT,
                          ffi::ResultBuffer<ffi::F32> vT, ffi::ResultBuffer<ffi::U8> ws, int32_t S,
                          int32_t per_env, int32_t n_steps, int32_t ckpt_every, float dt, float radius,
                          float mass, float friction, float elasticity, float ks, float ax, float ay) {
    const int64_t B = x0.dimensions()[0];
    doper_rollout_desc d = MakeDesc(B, S, per_env, n_steps, ckpt_every, dt, radius, mass, friction, elasticity, ks,
                                    ax, ay);
    if (ws->element_count() < doper_rollout_workspace_bytes(&d)) return Status(DOPER_ERR_WORKSPACE, "rollout_fwd");
    return Status(doper_rollout_fwd(stream, &d, scene_ws.typed_data(), x0.typed_data(), v0.typed_data(),
                                    xT->typed_data(), vT->typed_data(), nullptr, nullptr, nullptr,
                                    ws->typed_data()),
                  "doper_rollout_fwd");
}

ffi::Error RolloutBwdImpl(cudaStream_t stream, ffi::Buffer<ffi::U8> scene_ws, ffi::Buffer<ffi::U8> ws,
                          ffi::Buffer<ffi::F32> xT_bar, ffi::Buffer<ffi::F32> vT_bar,
                          ffi::ResultBuffer<ffi::F32> x0_bar, ffi::ResultBuffer<ffi::F32> v0_bar,
                          ffi::ResultBuffer<ffi::U8> scratch, int32_t S, int32_t per_env, int32_t n_steps, int32_t ckpt_every, float dt, float radius,
                          float mass, float friction, float elasticity, float ks, float ax, float ay) {
    const int64_t B = xT_bar.dimensions()[0];
    doper_rollout_desc d = MakeDesc(B, S, per_env, n_steps, ckpt_every, dt, radius, mass, friction, elasticity, ks,
                                    ax, ay);
    if (scratch->element_count() < doper_rollout_bwd_scratch_bytes(&d)) return Status(DOPER_ERR_WORKSPACE, "rollout_bwd");
    const void* ws_in = ws.typed_data();  // const: the forward's workspace is read only
    return Status(doper_rollout_bwd(stream, &d, scene_ws.typed_data(), ws_in, scratch->typed_data(),
                                    xT_bar.typed_data(), vT_bar.typed_data(), x0_bar->typed_data(),
                                    v0_bar->typed_data()),
                  "doper_rollout_bwd");
}

#define DOPER_ROLLOUT_ATTRS()                                                                         \
    .Attr<int32_t>("S").Attr<int32_t>("per_env").Attr<int32_t>("n_steps").Attr<int32_t>("ckpt_every")  \
    .Attr<float>("dt").Attr<float>("radius").Attr<float>("mass").Attr<float>("friction")              \
    .Attr<float>("elasticity").Attr<float>("attractor_strength").Attr<float>("attractor_x")           \
    .Attr<float>("attractor_y")

}  // namespace

XLA_FFI_DEFINE_HANDLER_SYMBOL(doper_scene_prepare_ffi, ScenePrepareImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>()
                                  .Attr<float>("max_radius"));

XLA_FFI_DEFINE_HANDLER_SYMBOL(doper_rollout_fwd_ffi, RolloutFwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>() DOPER_ROLLOUT_ATTRS());

XLA_FFI_DEFINE_HANDLER_SYMBOL(doper_rollout_bwd_ffi, RolloutBwdImpl,
                              ffi::Ffi::Bind()
                                  .Ctx<ffi::PlatformStream<cudaStream_t>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::U8>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Arg<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::F32>>()
                                  .Ret<ffi::Buffer<ffi::U8>>() DOPER_ROLLOUT_ATTRS());
#else
// jaxlib's FFI headers are not available: nothing to build (see the note at the top).
extern "C" int doper_xla_ffi_shim_unavailable(void) { return 1; }
#endif
