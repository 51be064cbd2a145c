// doper_b200 CUDA translation unit (sm_100a): the kernels live in the headers below, the C-ABI
// launchers (include/doper_b200.h) in this file.
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//          -prec-div=true -prec-sqrt=true -ftz=false  (see doper_b200/build.py)
//
//   physics.cuh    per-step physics and its adjoint in the reference's fp32 operation order
//   sweep.cuh      narrow phase (filtered exact sweep), broad phase (candidate lists), grid queries
//   scene.cuh      prepared-scene layout, scene_prepare_kernel (+ uniform grid), TMA staging
//   fwd_full.cuh   full-sweep forward kernels (every pair every step)
//   fwd_broad.cuh  broad-phase forward kernels with per-ball candidate lists: rollout_fwd_cull_kernel, rollout_fwd_tpc_kernel
//   fwd_baked.cuh  rollout_fwd_baked_kernel: static per-cell candidate lists baked into the prepared scene (default plan, shared scenes)
//   fwd_pipe.cuh   rollout_fwd_pipe_kernel: producer warp (free-flight chain, lane = (ball, axis)) + consumer warp (baked look-ups,
//                  log, collisions) per 16 balls; default plan of shared scenes up to kPipeMaxBalls balls
//   bwd.cuh        rollout_bwd_kernel (sequential fp32, bit-exact), rollout_bwd_jacobian_kernel + rollout_bwd_fold_kernel (binary64, default)
//   queries.cuh    l1_loss_kernel, closest_segment_kernel, points_inside_kernel, fp32_probe_kernel
//   sensor.cuh     range-sensor ray casting
//   peer.cuh       one-shot all-reduce of the policy gradient over NVLink peer memory
//   controller.cuh the trainer's glue: controller MLP forward / backward, gradient clip + Adam (three kernels)
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#include <type_traits>

#include "../../include/doper_b200.h"
#include "physics.cuh"
#include "sweep.cuh"
#include "scene.cuh"
#include "fwd_full.cuh"
#include "fwd_broad.cuh"
#include "fwd_baked.cuh"
#include "bwd.cuh"
#include "fwd_pipe.cuh"
#include "queries.cuh"

#include "sensor.cuh"
#include "peer.cuh"
#include "controller.cuh"

namespace doper {
// doper_value_and_grad, separate-launch plan: the L1 loss (blocks [0, loss_blocks): l1_loss_kernel's blocks) and pass 1
// of the binary64 backward (the other blocks: rollout_bwd_jacobian_kernel's threads) are independent of each other —
// both only read what the forward wrote — so they share ONE launch: a launch gap and the shorter of the two kernels
// come off BASELINE configs[1]'s critical path.  Same device functions, same bits.
__global__ void __launch_bounds__(kLossThreads) loss_jacobian_kernel(const int loss_blocks, int64_t B, const float2* __restrict__ xT,
                                                                    float tx, float ty, float* __restrict__ loss_pb,
                                                                    float* __restrict__ loss_sum, float2* __restrict__ xT_bar,
                                                                    const BwdParams p) {
    if ((int)blockIdx.x < loss_blocks) {
        l1_loss_block(blockIdx.x, B, xT, tx, ty, loss_pb, loss_sum, xT_bar);
    } else {
        bwd_jacobian_part(p, (int64_t)(blockIdx.x - loss_blocks) * blockDim.x + threadIdx.x,
                          (int64_t)(gridDim.x - loss_blocks) * blockDim.x);
    }
}
}  // namespace doper

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace doper;

namespace {

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int launch_status() { return cudaPeekAtLastError() == cudaSuccess ? DOPER_OK : DOPER_ERR_LAUNCH; }

int max_optin_smem() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    return v;
}

int sm_count() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 148;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || v <= 0) return 148;
    return v;
}

Consts make_consts(const doper_rollout_desc* d) {
    Consts c;
    c.r = d->consts[0]; c.m = d->consts[1]; c.f = d->consts[2]; c.e = d->consts[3]; c.ks = d->consts[4];
    c.ax = d->attractor[0]; c.ay = d->attractor[1];
    c.dt = d->dt;
    // div_invariant(): fast path only for well-scaled divisors (then y = RN(1/b) is normal and the
    // residuals stay normal for |a| in [1e-18, 1e18]); otherwise an empty range = always a / b
    const bool ok = c.r >= 1e-6f && c.r <= 1e6f && c.m >= 1e-6f && c.m <= 1e6f;
    c.rcp_r = 1.0f / c.r;
    c.rcp_m = 1.0f / c.m;
    c.div_lo = ok ? 1e-18f : 1.0f;
    c.div_hi = ok ? 1e18f : 0.0f;
    return c;
}

// Checkpoint period.  Checkpoints only serve the replay fallback of the backward (a collided step whose
// record could not be stored: record area full or disabled), so K trades 16/K bytes per ball-step against a
// replay of at most K steps.
int resolve_K(const doper_rollout_desc* d) { return d->ckpt_every > 0 ? d->ckpt_every : 32; }

// Collision records kept by the forward for the backward (state at the start of the step + segment index,
// 32 B each; the backward's scratch holds one 4x4 binary64 Jacobian, 128 B, per record).  Default capacity:
// one record per 64 ball-steps (the seeded BASELINE batches collide on 1 step in 500 to 1,000) and never
// fewer than 65,536; a forward that runs out of room flags the remaining collisions in the hit log and
// the backward replays those from the checkpoints (slower, same results).  flags bit 19 = no records.
int64_t record_capacity(const doper_rollout_desc* d) {
    if (d->flags & ((1 << 19) | (1 << 17))) return 0;  // bit 17: the log payload is the exact index at every step
    const int64_t steps = d->B * (int64_t)d->n_steps;
    int64_t cap = d->hit_capacity > 0 ? d->hit_capacity : (steps / 64 > 65536 ? steps / 64 : 65536);
    if (cap > steps) cap = steps;
    if (cap > (int64_t)kLogSeg) cap = (int64_t)kLogSeg;
    return cap;
}

// Forward workspace: [header 256 B: int record count, hand-over counters | hand-over flags B x 4 B | checkpoints
// n_ckpt x B x 16 B | hit log n x B x 4 B | collision records cap x 32 B | hand-over entries B x 32 B].  Backward scratch (separate, caller-owned, written by doper_rollout_bwd):
// [cap x 16 doubles].
struct WsLayout { size_t flag_off, ckpt_off, log_off, rec_off, defer_off, total; int64_t cap; };

WsLayout ws_layout(const doper_rollout_desc* d) {
    const int K = resolve_K(d);
    const size_t n_ckpt = (size_t)(d->n_steps + K - 1) / K;
    WsLayout w;
    w.cap = record_capacity(d);
    // [256: hand-over flags u32 x B, zeroed with the header by the plans that use them (use_defer)] ... [hand-over entries 32 B x B]
    w.flag_off = 256;
    w.ckpt_off = (w.flag_off + (size_t)d->B * sizeof(unsigned int) + 255) / 256 * 256;
    w.log_off = (w.ckpt_off + n_ckpt * (size_t)d->B * sizeof(float4) + 255) / 256 * 256;
    w.rec_off = (w.log_off + (size_t)d->n_steps * (size_t)d->B * sizeof(int32_t) + 255) / 256 * 256;
    w.defer_off = (w.rec_off + (size_t)w.cap * sizeof(HitRecord) + 255) / 256 * 256;
    w.total = (w.defer_off + (size_t)d->B * sizeof(DeferEntry) + 255) / 256 * 256 + 256;
    return w;
}

// Chunks per ball of the binary64 backward: enough (ball, chunk) threads to fill the SMs, at most kFoldMaxChunks.
void fold_plan(const doper_rollout_desc* d, int* n_chunks, int* chunk_len) {
    int c = 1;
    while (c < kFoldMaxChunks && d->B * c < (int64_t)148 * 2048) c *= 2;
    while (c > 1 && d->n_steps / c < 8) c /= 2;
    const int len = (d->n_steps + c - 1) / c;
    *chunk_len = len;
    *n_chunks = (d->n_steps + len - 1) / len;
}

// Forward plan: time-parallel kernel (one warp per ball) unless the caller pinned a lane count,
// disabled it (flags bit 3) or the batch is large enough for one thread per ball.
constexpr int64_t kTpMaxBalls = 65536;

// Measured on B200 (tools/gpu_tune.py, profiles/): a shared small scene with a few thousand balls
// (BASELINE configs[1]) runs fastest on the register-resident lane-split kernel with 8 lanes per
// ball (0.63 ms vs 0.74 ms time-parallel); per-environment scenes, dense scenes and every other
// batch size up to 65,536 run fastest time-parallel; beyond that one thread per ball wins.
// Broad-phase (candidate list) kernels are the default plan.  They are not used when the caller
// forbids them (flags bit 15), forces the exact sweep (bit 0) or pins one of the full-sweep kernels
// (lanes_per_ball != 0 or the kernel-selection bits 2-3) without forcing them (bit 16).
// Measured on B200 (tools/tune_fwd.py, profiles/r1_plan_sweep.md): shared scene up to 8,192 balls ->
// time-parallel variant with 16 lanes (steps) per ball; 8,192-32,767 balls -> lane-split, 4 lanes;
// more -> one thread per ball; dense scenes (>= 2,048 segments) at least 2 lanes (the refresh scan
// is split over them); per-environment scenes -> lane-split, 4 lanes, scenes read in place.
constexpr int kLegacyPlanBits = 4 | 8;
bool use_baked(const doper_rollout_desc* d);

// the lane = step kernel needs the scene in shared memory; larger scenes go to the lane-split kernel,
// which reads them in place
bool scene_fits_smem(const doper_rollout_desc* d) { return (size_t)scene_layout(1, d->S).S_pad * 20 <= 200 * 1024; }

bool use_cull(const doper_rollout_desc* d) {
    if (use_baked(d)) return false;
    if ((d->flags & (1 << 15)) || (d->flags & 1)) return false;
    if (d->flags & (1 << 16)) return true;
    return d->lanes_per_ball == 0 && !(d->flags & kLegacyPlanBits);
}

// Baked broad phase (static per-cell lists, fwd_baked.cuh): shared scenes under the default plan.  Not with the
// exact index log (bit 17: the cell lists are hit lists), a forced exact sweep (bit 0), pinned lane counts /
// legacy kernels, or flags bit 20 (which keeps the per-ball candidate-list kernels).
constexpr int kBakedOffBits = 1 | kLegacyPlanBits | (1 << 15) | (1 << 16) | (1 << 17) | (1 << 18) | (1 << 20);

bool baked_staged(const doper_rollout_desc* d) {
    if (d->per_env) return false;  // every ball has its own scene and baked lists: read in place
    const int S_pad = scene_layout(1, d->S).S_pad;
    return (size_t)S_pad * 20 + bake_bytes(S_pad) + 64 <= 128 * 1024;  // 32 KB of it are the sub-cell masks
}

// Pipelined kernel (fwd_pipe.cuh: producer / consumer warp pairs over the baked lists): the default of shared
// scenes that fit shared memory up to kPipeMaxBalls balls (beyond, one thread per ball fills the machine and
// executes fewer instructions per ball-step).  flags bit 22 forbids it; bit 23 with lanes_per_ball = 8, 16 or 32
// pins it with that window length.
constexpr int64_t kPipeMaxBalls = 6144;
bool pipe_pinned(const doper_rollout_desc* d) {
    return !d->per_env && (d->flags & (1 << 23)) && !(d->flags & kBakedOffBits) &&
           (d->lanes_per_ball == 8 || d->lanes_per_ball == 16 || d->lanes_per_ball == 32);
}
bool use_pipe(const doper_rollout_desc* d) {
    if (d->per_env || d->n_steps < 1 || !baked_staged(d)) return false;
    if (pipe_pinned(d)) return true;
    return d->lanes_per_ball == 0 && !(d->flags & (kBakedOffBits | (1 << 21) | (1 << 22) | (1 << 23))) && d->B <= kPipeMaxBalls;
}
// One launch for forward + L1 loss + backward (rollout_fwd_pipe_kernel<U, true>): the pipelined plan, the default
// (binary64) backward, collision records enabled, and a staged scene large enough to hold the fold's matrices.
constexpr int64_t kFusedMaxBalls = 1024;  // the fused variant runs one block per SM (fwd_pipe.cuh): at most 148 blocks of ceil(B / 148) balls
bool fused_ok(const doper_rollout_desc* d) {
    if (!use_pipe(d) || (d->flags & 2) || record_capacity(d) <= 0 || d->n_steps < 1 || d->B > kFusedMaxBalls) return false;
    const int S_pad = scene_layout(1, d->S).S_pad;
    return (size_t)S_pad * 20 + bake_lists_bytes(S_pad) >= 16 * 16 * 16 * sizeof(double);
}

int pipe_window(const doper_rollout_desc* d) { return pipe_pinned(d) ? d->lanes_per_ball : 16; }

bool use_baked(const doper_rollout_desc* d) {
    if (use_pipe(d)) return true;  // same lists; the lane = step kernel serves the calls the pipelined one does not (trajectory output)
    if ((d->flags & (1 << 21)) && d->lanes_per_ball >= 1 && d->lanes_per_ball <= 16) return true;  // pinned G
    return d->lanes_per_ball == 0 && !(d->flags & kBakedOffBits);  // shared scenes and, read in place, per-environment scenes
}

// lanes (= speculated time steps) per ball of the baked kernel: small batches need them to fill the SMs
int baked_lanes(const doper_rollout_desc* d) {
    if ((d->flags & (1 << 21)) && !pipe_pinned(d)) return d->lanes_per_ball;
    if (d->per_env) {  // scenes and lists read in place (L2 latency): more speculated steps per ball pay longer
        if (d->B <= 2048) return 16;   // measured (tools/tune_baked.py c3): 8,192 envs 0.33 ms with 8 lanes, 0.41 with 4;
        if (d->B <= 16384) return 8;   // 65,536 envs 2.50 ms with one thread per ball, 2.70 with 2 lanes
        return 1;
    }
    if (d->B <= 1024) return 16;
    if (d->B <= 4096) return 8;
    if (d->B <= 16384) return 4;
    if (d->B <= 49152) return 2;
    return 1;
}

// Hand-over of balls in sliding contact to tail-worker warps: the baked kernels (not the pipelined one); a launch
// otherwise lasts as long as the slowest warp's share of the chain of a ball that collides on every step.  flags
// bit 25 switches it off.
bool use_defer(const doper_rollout_desc* d) {
    return use_baked(d) && !use_pipe(d) && !(d->flags & (1 << 25));
}

bool use_tpc(const doper_rollout_desc* d) {
    if (!use_cull(d)) return false;
    if (d->lanes_per_ball != 0) return (d->flags & (1 << 18)) != 0 && d->lanes_per_ball >= 2;
    if (d->flags & (1 << 18)) return true;
    return !d->per_env && d->B <= 8192 && scene_fits_smem(d);
}

int cull_lanes(const doper_rollout_desc* d) {
    if (d->lanes_per_ball != 0) return d->lanes_per_ball;
    if (d->per_env) return 4;  // read in place (see launch_fwd_cull); measured: C3 5.4 ms vs 8.2 ms staged with 8 lanes
    // lane = time step: one ball per warp while the batch is too small to fill the SMs otherwise (measured on
    // C2-like scenes: 1,024 balls 0.105 ms with 32 lanes vs 0.132 with 16; 4,096 balls 0.189 vs 0.176)
    if (use_tpc(d)) return d->B <= 2048 ? 32 : 16;
    int g = d->B < 32768 ? 4 : 1;
    if (scene_layout(1, d->S).S_pad >= kGridMinSegments) g = 2;  // dense scene (grid refresh): measured best on C4
    if (!scene_fits_smem(d) && g < 8) g = 8;  // unstaged scene: split the refresh scan (L2 reads) over more lanes
    return g;
}

bool plan_reg8(const doper_rollout_desc* d) {
    return !use_baked(d) && d->lanes_per_ball == 0 && !d->per_env && !(d->flags & 4) && scene_layout(1, d->S).S_pad <= 256 &&
           d->B >= 1536 && d->B <= 6144 && !use_cull(d);
}

bool use_tp(const doper_rollout_desc* d) {
    return !use_baked(d) && d->lanes_per_ball == 0 && !(d->flags & 8) && d->B <= kTpMaxBalls && !plan_reg8(d) && !use_cull(d);
}

// lane = time step kernels write [B][n_steps] so that a group's stores are contiguous
bool log_ball_major(const doper_rollout_desc* d) { return use_tp(d) || use_tpc(d) || use_pipe(d) || (use_baked(d) && baked_lanes(d) > 1); }

int tp_warps_per_block(const doper_rollout_desc* d) {
    if (d->per_env) return 4;
    const size_t scene = (size_t)scene_layout(1, d->S).S_pad * 20;
    if (scene <= 24 * 1024) return 4;
    if (scene <= 56 * 1024) return 8;
    if (scene <= 110 * 1024) return 16;
    return 32;
}

int validate_desc(const doper_rollout_desc* d) {
    if (!d) return DOPER_ERR_NULL_PTR;
    if (d->B <= 0 || d->S <= 0 || d->n_steps < 0 || d->ckpt_every < 0 || d->hit_capacity < 0) return DOPER_ERR_BAD_SHAPE;
    if (d->B * (int64_t)d->n_steps > ((int64_t)1 << 40)) return DOPER_ERR_UNSUPPORTED;
    if (d->ckpt_every > 1024) return DOPER_ERR_BAD_SHAPE;
    // candidate lists hold 16-bit segment indices, the hit log 29-bit ones: 65,504 segments per scene is the
    // limit of the rollout kernels (the largest BASELINE scene has 10,002)
    if (d->S > 65504) return DOPER_ERR_UNSUPPORTED;
    const int g = d->lanes_per_ball;
    if (!(g == 0 || g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32)) return DOPER_ERR_BAD_SHAPE;
    return DOPER_OK;
}

bool plan_reg8(const doper_rollout_desc* d);

// lanes per ball: enough threads to fill 148 SMs x 256 threads; per-env scenes are staged per
// ball, so spreading one ball over 8 lanes keeps the shared-memory footprint per thread sane.
int auto_lanes(const doper_rollout_desc* d) {
    if (use_cull(d)) return cull_lanes(d);
    if (d->lanes_per_ball) return d->lanes_per_ball;
    if (plan_reg8(d)) return 8;
    if (d->per_env) return 8;
    int g = 1;
    while (g < 32 && d->B * g < 148 * 256) g *= 2;
    return g;
}

size_t fwd_smem_bytes(const doper_rollout_desc* d, int G) {
    const SceneLayout L = scene_layout(1, d->S);
    const size_t n_local = d->per_env ? (size_t)(kFwdThreads / G) : 1;
    return n_local * L.S_pad * 20 + 16;
}

template <int G, int NV>
int launch_fwd_reg(cudaStream_t st, const FwdParams& p) {
    const int bpb = kRegThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    rollout_fwd_reg_kernel<G, NV><<<grid, kRegThreads, 0, st>>>(p);
    return launch_status();
}

// register-resident scene: S_pad in {32,...,256}, G in {8,16,32}, NV = S_pad / G
template <int G>
int dispatch_fwd_reg(cudaStream_t st, const FwdParams& p, int S_pad) {
    switch (S_pad / 32) {
        case 1: return launch_fwd_reg<G, 32 / G>(st, p);
        case 2: return launch_fwd_reg<G, 64 / G>(st, p);
        case 3: return launch_fwd_reg<G, 96 / G>(st, p);
        case 4: return launch_fwd_reg<G, 128 / G>(st, p);
        case 5: return launch_fwd_reg<G, 160 / G>(st, p);
        case 6: return launch_fwd_reg<G, 192 / G>(st, p);
        case 7: return launch_fwd_reg<G, 224 / G>(st, p);
        default: return launch_fwd_reg<G, 256 / G>(st, p);
    }
}

template <int G>
int launch_fwd(cudaStream_t st, const FwdParams& p, size_t smem) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(rollout_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const int bpb = kFwdThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    rollout_fwd_kernel<G><<<grid, kFwdThreads, smem, st>>>(p);
    return launch_status();
}

template <int G>
int launch_fwd_cull(cudaStream_t st, FwdParams p, size_t scene_smem) {
    const size_t lists = (size_t)CullCfg<G>::LCAP * kFwdThreads * sizeof(unsigned short);
    // a scene that does not fit shared memory (next to the lists) is read in place
    p.stage = (int)(scene_smem + lists) <= max_optin_smem() ? 1 : 0;
    // per-environment scenes with few lanes per ball: staging 128 / G private scenes per block leaves room for
    // one block per SM at best; reading them in place (each env's 4.5 KB stays in L1 / L2) measured faster
    if (p.per_env && G < 8) p.stage = 0;
    const size_t smem = (p.stage ? scene_smem : 0) + lists;
    const int bpb = kFwdThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    const bool dense = !p.per_env && scene_layout(1, p.S).S_pad >= kGridMinSegments;
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)DOPER_ERR_SMEM_LIMIT;
        kernel<<<grid, kFwdThreads, smem, st>>>(p);
        return launch_status();
    };
    if (p.stage) return dense ? go(rollout_fwd_cull_kernel<G, true, true>) : go(rollout_fwd_cull_kernel<G, true, false>);
    return dense ? go(rollout_fwd_cull_kernel<G, false, true>) : go(rollout_fwd_cull_kernel<G, false, false>);
}

template <int G>
int launch_fwd_tpc(cudaStream_t st, const FwdParams& p, size_t scene_smem) {
    // Block size.  This kernel runs small batches (a few thousand balls, fewer resident warps than the
    // SMs could hold), so what matters is how evenly the balls spread over the 148 SMs: 4,096 balls in
    // blocks of 8 are 512 blocks = 4 blocks on some SMs and 3 on others (32 vs 24 balls, and the busy
    // SMs set the kernel time); blocks of 64 threads halve the granularity (28 vs 24).
    const bool small_blocks = !p.per_env && (p.B + 128 / G - 1) / (128 / G) < 8 * (int64_t)sm_count();
    auto run = [&](auto threads_tag) {
        constexpr int THREADS = decltype(threads_tag)::value;
        const size_t smem = scene_smem + (size_t)(THREADS / G) * kTpcListCap * sizeof(unsigned short);
        if ((int)smem > max_optin_smem()) return (int)DOPER_ERR_SMEM_LIMIT;
        const int bpb = THREADS / G;
        const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
        const bool dense = !p.per_env && scene_layout(1, p.S).S_pad >= kGridMinSegments;
        auto go = [&](auto kernel) {
            if (smem > 48 * 1024 &&
                cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
                return (int)DOPER_ERR_SMEM_LIMIT;
            kernel<<<grid, THREADS, smem, st>>>(p);
            return launch_status();
        };
        return dense ? go(rollout_fwd_tpc_kernel<G, true, THREADS>) : go(rollout_fwd_tpc_kernel<G, false, THREADS>);
    };
    if (small_blocks) return run(std::integral_constant<int, 64>{});
    return run(std::integral_constant<int, 128>{});
}

// Block size of the one-wave kernels whose warps are independent latency chains (the baked forward kernels): the kernel
// lasts as long as the SM with the most resident warps, so the batch's warps are dealt out evenly — w = ceil(warps /
// (SMs x waves)) warps per block, one block per SM and wave — instead of fixed 128-thread blocks (131,072 balls = 1,024
// blocks against 6 resident per SM by shared memory: a second wave of 136 blocks; 65,536 threads = 512 blocks: 4 on
// some SMs, 3 on others).  `fits(threads)` = resident blocks per SM at that size.  DOPER_BAKED_THREADS pins the size
// (tuning: tools/tune_block.py).
template <class Fits>
int balanced_block(int64_t threads_total, int max_threads, Fits fits) {
    const char* const env = getenv("DOPER_BAKED_THREADS");
    const int pinned = env ? atoi(env) : 0;
    if (pinned >= 32 && pinned <= max_threads && pinned % 32 == 0) return pinned;
    const int64_t warps = (threads_total + 31) / 32, sms = sm_count();
    int cap = max_threads / 32;  // warps one SM holds in ONE block of this kernel
    while (cap > 1 && fits(32 * cap) < 1) --cap;
    int resident = cap * fits(32 * cap);  // warps one SM holds in blocks of the largest size
    if (resident > 64) resident = 64;
    const int64_t waves = (warps + sms * resident - 1) / (sms * resident);
    const int per_sm = (int)((warps + sms * waves - 1) / (sms * waves));  // warps per SM and wave, dealt evenly
    const int blocks = (per_sm + cap - 1) / cap;                           // ... in this many equal blocks
    const int w = (per_sm + blocks - 1) / blocks;
    return 32 * (w < 1 ? 1 : w);
}

template <int G>
int launch_fwd_baked(cudaStream_t st, const FwdParams& p, bool staged) {
    constexpr int MAXT = 256;
    constexpr int MAXT_STAGED = G == 1 ? 1024 : MAXT;  // one thread per ball, staged scene: 64 registers, up to 32 warps in one block
    const int S_pad = scene_layout(1, p.S).S_pad;
    const size_t smem = staged ? (size_t)S_pad * 20 + (bake_bytes(S_pad) - sizeof(BakeHeader)) + 16 : 0;
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)DOPER_ERR_SMEM_LIMIT;
        const int max_threads = staged ? MAXT_STAGED : MAXT;
        // tail-worker warps (hand-over of balls in sliding contact, fwd_baked.cuh): the block's last warps
        FwdParams q = p;
        const int tail = p.defer.entry ? (max_threads >= 1024 ? 4 : 1) : 0;
        q.defer.tail_warps = tail;
        q.defer.streak = kDeferStreak;
        if (const char* e = getenv("DOPER_DEFER_STREAK")) { const int v = atoi(e); if (v >= 1) q.defer.streak = v; }  // tuning
        const int threads = balanced_block(p.B * G, max_threads - 32 * tail, [&](int t) {
            int nb = 0;
            return cudaOccupancyMaxActiveBlocksPerMultiprocessor(&nb, kernel, t + 32 * tail, smem) == cudaSuccess ? nb : 0;
        });
        const int bpb = threads / G;
        const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
        kernel<<<grid, threads + 32 * tail, smem, st>>>(q);
        if (launch_status() != DOPER_OK) return (int)DOPER_ERR_LAUNCH;
        if (tail > 0) {  // (= a hand-over queue exists) whatever the workers left in it: usually nothing
            rollout_fwd_drain_kernel<<<(unsigned)sm_count(), 128, 0, st>>>(q);
            return launch_status();
        }
        return (int)DOPER_OK;
    };
    return staged ? go(rollout_fwd_baked_kernel<G, true, MAXT_STAGED>) : go(rollout_fwd_baked_kernel<G, false, MAXT>);
}

template <int U, bool FUSED, int PB>
int launch_fwd_pipe_pb(cudaStream_t st, const FwdParams& p, int used_slots) {
    const int S_pad = scene_layout(1, p.S).S_pad;
    const size_t smem = (size_t)S_pad * 20 + bake_lists_bytes(S_pad) + 16 + sizeof(PipeShared<U>);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(rollout_fwd_pipe_kernel<U, FUSED, PB>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    FwdParams q = p;
    q.pipe_balls = used_slots < 1 ? 1 : (used_slots > PB ? PB : used_slots);
    const unsigned grid = (unsigned)((p.B + q.pipe_balls - 1) / q.pipe_balls);
    rollout_fwd_pipe_kernel<U, FUSED, PB><<<grid, 32 * (1 + PB * U / 32), smem, st>>>(q);
    return launch_status();
}

// Ball slots per block.  A collision holds up the window barrier of its block, so a block should carry as few balls as
// the machine allows: (a) small batches leave block slots of the GPU empty — they use only ceil(B / slots) of a
// block's 16 ball slots (16 balls: 0.039 -> 0.033 ms forward with 2 per block; 1,024 balls 0.090 -> 0.080 with 4);
// (b) DOPER_PIPE_PB / DOPER_PIPE_BALLS pin the block shape / the used slots (tools/tune_env.py).
template <int U, bool FUSED>
int launch_fwd_pipe(cudaStream_t st, const FwdParams& p) {
    // (c) the default window with separate backward launches: blocks of 8 ball slots (5 warps, four resident per SM)
    //     while the batch fits one wave of them — BASELINE configs[1]: forward 0.1126 -> 0.1065 ms
    int pb = kPipeBalls;
    if (U <= 16 && !FUSED && p.B <= (int64_t)8 * 4 * sm_count()) pb = 8;
    if (const char* e = getenv("DOPER_PIPE_PB")) pb = atoi(e) == 8 ? 8 : kPipeBalls;
    if (!(U <= 16 && !FUSED)) pb = kPipeBalls;
    const int resident = (FUSED || U > 16) ? 1 : (pb == 8 ? 4 : 2);  // blocks per SM (launch bounds of the kernel)
    const int64_t slots = (int64_t)sm_count() * resident;
    int used = (int)((p.B + slots - 1) / slots);
    if (used < 2) used = 2;
    if (used > pb) used = pb;
    if (const char* e = getenv("DOPER_PIPE_BALLS")) { const int v = atoi(e); if (v >= 1 && v <= pb) used = v; }
    if constexpr (U <= 16 && !FUSED) {
        if (pb == 8) return launch_fwd_pipe_pb<U, FUSED, 8>(st, p, used);
    }
    return launch_fwd_pipe_pb<U, FUSED, kPipeBalls>(st, p, used);
}

template <bool FUSED>
int dispatch_fwd_pipe(cudaStream_t st, const FwdParams& p, int window) {
    switch (window) {
        case 8: return launch_fwd_pipe<8, FUSED>(st, p);
        case 32: return launch_fwd_pipe<32, FUSED>(st, p);
        default: return launch_fwd_pipe<16, FUSED>(st, p);
    }
}

template <int G>
int launch_closest(cudaStream_t st, int64_t N, int S, int per_env, const float* pts, const float* segs, const void* ws,
                   int32_t* idx, float* dist, float* seg_out) {
    const size_t smem = (size_t)scene_layout(1, S).S_pad * 20 + 16;
    const bool staged = !per_env && (int)smem <= max_optin_smem();  // else the scene(s) are read in place
    const int ppb = kQueryThreads / G;
    const unsigned grid = (unsigned)((N + ppb - 1) / ppb);
    if (staged) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(closest_segment_kernel<G, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return DOPER_ERR_SMEM_LIMIT;
        closest_segment_kernel<G, true><<<grid, kQueryThreads, smem, st>>>(
            N, S, 0, reinterpret_cast<const float2*>(pts), reinterpret_cast<const float4*>(segs),
            reinterpret_cast<const unsigned char*>(ws), idx, dist, reinterpret_cast<float4*>(seg_out));
    } else {
        closest_segment_kernel<G, false><<<grid, kQueryThreads, 0, st>>>(
            N, S, per_env, reinterpret_cast<const float2*>(pts), reinterpret_cast<const float4*>(segs),
            reinterpret_cast<const unsigned char*>(ws), idx, dist, reinterpret_cast<float4*>(seg_out));
    }
    return launch_status();
}

int closest_dispatch(cudaStream_t st, int64_t N, int S, int per_env, const float* points, const float* segments,
                     const void* scene_ws, int32_t* idx, float* dist, float* seg_out) {
    int G = 1;
    while (G < 32 && N * G < 148 * 256) G *= 2;
    switch (G) {
        case 1: return launch_closest<1>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
        case 2: return launch_closest<2>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
        case 4: return launch_closest<4>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
        case 8: return launch_closest<8>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
        case 16: return launch_closest<16>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
        default: return launch_closest<32>(st, N, S, per_env, points, segments, scene_ws, idx, dist, seg_out);
    }
}

}  // namespace

namespace {
template <bool OBS, bool POS64>
int launch_range_sensor(cudaStream_t st, int64_t N, int R, int S, int per_env, const void* pos, const float* vel,
                        const float* ray_dirs, const void* scene_ws, float distance_range, double tx, double ty,
                        void* ranges, float* points, int32_t* idx, float* obs) {
    const size_t smem = (size_t)scene_layout(1, S).S_pad * 20 + 16;
    const bool staged = !per_env && (int)smem <= max_optin_smem();  // else the scene(s) are read in place
    const int64_t n = N * (int64_t)R;
    const unsigned grid = (unsigned)((n + kSensorThreads - 1) / kSensorThreads);
    auto go = [&](auto kernel, size_t sm) {
        if (sm > 48 * 1024 && cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)sm) != cudaSuccess)
            return (int)DOPER_ERR_SMEM_LIMIT;
        kernel<<<grid, kSensorThreads, sm, st>>>(N, R, S, per_env, pos, reinterpret_cast<const float2*>(vel),
                                                 reinterpret_cast<const float2*>(ray_dirs), reinterpret_cast<const unsigned char*>(scene_ws),
                                                 distance_range, tx, ty, ranges, reinterpret_cast<float2*>(points), idx, obs);
        return launch_status();
    };
    return staged ? go(range_sensor_kernel<OBS, true, POS64>, smem) : go(range_sensor_kernel<OBS, false, POS64>, 0);
}
}  // namespace

extern "C" {

int doper_abi_version(void) { return DOPER_ABI_VERSION; }

int doper_fp_policy(void) { return DOPER_FP_POLICY; }

const char* doper_error_string(int code) {
    switch (code) {
        case DOPER_OK: return "ok";
        case DOPER_ERR_BAD_SHAPE: return "bad shape or size argument";
        case DOPER_ERR_NULL_PTR: return "required pointer is NULL";
        case DOPER_ERR_MISALIGNED: return "pointer is not aligned as documented";
        case DOPER_ERR_SMEM_LIMIT: return "scene does not fit the shared-memory plan of the kernel";
        case DOPER_ERR_LAUNCH: return "CUDA kernel launch failed";
        case DOPER_ERR_WORKSPACE: return "workspace too small";
        case DOPER_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown doper_b200 error";
    }
}

size_t doper_scene_workspace_bytes(int64_t n_scenes, int32_t S) {
    if (n_scenes <= 0 || S <= 0) return 0;
    return scene_layout(n_scenes, S).total;
}

int doper_scene_prepare(doper_stream_t stream, int64_t n_scenes, int32_t S, const float* segments,
                        float max_radius, void* scene_ws) {
    if (n_scenes <= 0 || S <= 0 || n_scenes > 0x7fffffff) return DOPER_ERR_BAD_SHAPE;
    if (!segments || !scene_ws) return DOPER_ERR_NULL_PTR;
    if (!aligned(segments, 16) || !aligned(scene_ws, 256)) return DOPER_ERR_MISALIGNED;
    scene_prepare_kernel<<<(unsigned)n_scenes, 128, 0, (cudaStream_t)stream>>>(
        n_scenes, S, reinterpret_cast<const float4*>(segments), reinterpret_cast<unsigned char*>(scene_ws));
    if (launch_status() != DOPER_OK) return DOPER_ERR_LAUNCH;
    // bake the per-cell candidate lists of every scene (max_radius <= 0 or NaN: marked absent)
    scene_bake_kernel<<<(unsigned)n_scenes, 256, 0, (cudaStream_t)stream>>>(n_scenes, S, max_radius > 0.0f ? max_radius : 0.0f,
                                                                           reinterpret_cast<unsigned char*>(scene_ws));
    return launch_status();
}

int doper_closest_segment(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                          const float* segments, const void* scene_ws, int32_t* idx, float* dist,
                          float* seg_out) {
    if (N < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !scene_ws || !idx || !dist) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16) || !aligned(scene_ws, 256) ||
        (seg_out && !aligned(seg_out, 16)))
        return DOPER_ERR_MISALIGNED;
    return closest_dispatch((cudaStream_t)stream, N, S, 0, points, segments, scene_ws, idx, dist, seg_out);
}

int doper_closest_segment_env(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                              const float* segments, const void* scene_ws, int32_t* idx, float* dist,
                              float* seg_out) {
    if (N < 0 || S <= 0 || N > 0x7fffffff) return DOPER_ERR_BAD_SHAPE;
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !scene_ws || !idx || !dist) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16) || !aligned(scene_ws, 256) ||
        (seg_out && !aligned(seg_out, 16)))
        return DOPER_ERR_MISALIGNED;
    return closest_dispatch((cudaStream_t)stream, N, S, 1, points, segments, scene_ws, idx, dist, seg_out);
}

int doper_points_inside(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                        const float* segments, uint8_t* flag) {
    if (N < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (S % 3 != 0) return DOPER_ERR_UNSUPPORTED;  // jax_geometry.py:113 reshapes to (.., polygons, 3)
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !flag) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16)) return DOPER_ERR_MISALIGNED;
    const unsigned grid = (unsigned)((N + kQueryThreads - 1) / kQueryThreads);
    points_inside_kernel<<<grid, kQueryThreads, 0, (cudaStream_t)stream>>>(
        N, S, reinterpret_cast<const float2*>(points), reinterpret_cast<const float4*>(segments), flag);
    return launch_status();
}

int doper_points_inside_env(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                            const float* segments, uint8_t* flag) {
    if (N < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (S % 3 != 0) return DOPER_ERR_UNSUPPORTED;
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !flag) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16)) return DOPER_ERR_MISALIGNED;
    const unsigned grid = (unsigned)((N + kQueryThreads - 1) / kQueryThreads);
    points_inside_env_kernel<<<grid, kQueryThreads, 0, (cudaStream_t)stream>>>(
        N, S, reinterpret_cast<const float2*>(points), reinterpret_cast<const float4*>(segments), flag);
    return launch_status();
}

int doper_ray_segment_points(doper_stream_t stream, int64_t R, int32_t S, const float* origins,
                             const float* directions, const float* segments, float* points) {
    if (R < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (R == 0) return DOPER_OK;
    if (!origins || !directions || !segments || !points) return DOPER_ERR_NULL_PTR;
    if (!aligned(origins, 8) || !aligned(directions, 8) || !aligned(segments, 16) || !aligned(points, 8))
        return DOPER_ERR_MISALIGNED;
    const int64_t n = R * (int64_t)S;
    if (n > (int64_t)0x7fffffff * 256) return DOPER_ERR_UNSUPPORTED;
    ray_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        R, S, reinterpret_cast<const float2*>(origins), reinterpret_cast<const float2*>(directions),
        reinterpret_cast<const float4*>(segments), reinterpret_cast<float2*>(points));
    return launch_status();
}

namespace {
int sensor_args_ok(int64_t N, int R, int S, const void* positions, const void* ray_directions, const void* scene_ws, size_t pos_align) {
    if (N < 0 || R <= 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (N > 0 && (!positions || !ray_directions || !scene_ws)) return DOPER_ERR_NULL_PTR;
    if (!aligned(positions, pos_align) || !aligned(ray_directions, 8) || !aligned(scene_ws, 256)) return DOPER_ERR_MISALIGNED;
    return DOPER_OK;
}
}  // namespace

int doper_range_sensor(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                       const float* ray_directions, const void* scene_ws, float distance_range, float* ranges,
                       float* points, int32_t* idx) {
    const int rc = sensor_args_ok(N, R, S, positions, ray_directions, scene_ws, 8);
    if (rc || N == 0) return rc;
    if (!ranges) return DOPER_ERR_NULL_PTR;
    if (points && !aligned(points, 8)) return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<false, false>((cudaStream_t)stream, N, R, S, 0, positions, nullptr, ray_directions, scene_ws,
                                             distance_range, 0.0, 0.0, ranges, points, idx, nullptr);
}

int doper_range_sensor_env(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                           const float* ray_directions, const void* scene_ws, float distance_range, float* ranges,
                           float* points, int32_t* idx) {
    const int rc = sensor_args_ok(N, R, S, positions, ray_directions, scene_ws, 8);
    if (rc || N == 0) return rc;
    if (!ranges) return DOPER_ERR_NULL_PTR;
    if (points && !aligned(points, 8)) return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<false, false>((cudaStream_t)stream, N, R, S, 1, positions, nullptr, ray_directions, scene_ws,
                                             distance_range, 0.0, 0.0, ranges, points, idx, nullptr);
}

int doper_range_sensor_f64(doper_stream_t stream, int64_t N, int32_t R, int32_t S, int32_t per_env, const double* positions,
                           const float* ray_directions, const void* scene_ws, float distance_range, double* ranges,
                           float* points, int32_t* idx) {
    const int rc = sensor_args_ok(N, R, S, positions, ray_directions, scene_ws, 16);
    if (rc || N == 0) return rc;
    if (!ranges) return DOPER_ERR_NULL_PTR;
    if (points && !aligned(points, 8)) return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<false, true>((cudaStream_t)stream, N, R, S, per_env ? 1 : 0, positions, nullptr, ray_directions,
                                            scene_ws, distance_range, 0.0, 0.0, ranges, points, idx, nullptr);
}

int doper_agent_observations(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                             const float* velocities, const float* ray_directions, const void* scene_ws,
                             float distance_range, const double target[2], float* observations) {
    const int rc = sensor_args_ok(N, R, S, positions, ray_directions, scene_ws, 8);
    if (rc || N == 0) return rc;
    if (!velocities || !target || !observations) return DOPER_ERR_NULL_PTR;
    if (!aligned(velocities, 8)) return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<true, false>((cudaStream_t)stream, N, R, S, 0, positions, velocities, ray_directions, scene_ws,
                                            distance_range, target[0], target[1], nullptr, nullptr, nullptr, observations);
}

int doper_agent_observations_env(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                                 const float* velocities, const float* ray_directions, const void* scene_ws,
                                 float distance_range, const double target[2], float* observations) {
    const int rc = sensor_args_ok(N, R, S, positions, ray_directions, scene_ws, 8);
    if (rc || N == 0) return rc;
    if (!velocities || !target || !observations) return DOPER_ERR_NULL_PTR;
    if (!aligned(velocities, 8)) return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<true, false>((cudaStream_t)stream, N, R, S, 1, positions, velocities, ray_directions, scene_ws,
                                            distance_range, target[0], target[1], nullptr, nullptr, nullptr, observations);
}

int32_t doper_rollout_ckpt_every(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    return resolve_K(d);
}

int32_t doper_rollout_log_ball_major(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    return log_ball_major(d) ? 1 : 0;
}

const char* doper_rollout_plan_name(const doper_rollout_desc* d, int32_t backward) {
    if (validate_desc(d) != DOPER_OK) return "invalid";
    if (backward == 3)  // a second forward launch behind the main kernel ("" = none)
        return use_defer(d) ? "rollout_fwd_drain_kernel" : "";
    if (backward == 2)  // what doper_value_and_grad launches for the backward
        return (fused_ok(d) && !(d->flags & (1 << 24))) ? "fused into the forward launch (rollout_fwd_pipe_kernel<U, true>)"
                                                        : ((d->flags & 2) ? "rollout_bwd_kernel" : "loss_jacobian_kernel+rollout_bwd_fold_kernel");
    if (backward) return (d->flags & 2) ? "rollout_bwd_kernel" : "rollout_bwd_jacobian_kernel+rollout_bwd_fold_kernel";
    if (use_pipe(d)) {
        const int u = pipe_window(d);
        return u == 8 ? "rollout_fwd_pipe_kernel<8>" : (u == 32 ? "rollout_fwd_pipe_kernel<32>" : "rollout_fwd_pipe_kernel<16>");
    }
    if (use_baked(d)) {
        static const char* const baked[] = {"rollout_fwd_baked_kernel<1>", "rollout_fwd_baked_kernel<2>", "rollout_fwd_baked_kernel<4>",
                                            "rollout_fwd_baked_kernel<8>", "rollout_fwd_baked_kernel<16>"};
        int lg = 0;
        while ((1 << lg) < baked_lanes(d)) ++lg;
        return baked[lg];
    }
    const int G = auto_lanes(d);
    static const char* const tpc[] = {"rollout_fwd_tpc_kernel<2>", "rollout_fwd_tpc_kernel<4>", "rollout_fwd_tpc_kernel<8>",
                                      "rollout_fwd_tpc_kernel<16>", "rollout_fwd_tpc_kernel<32>"};
    static const char* const cull[] = {"rollout_fwd_cull_kernel<1>", "rollout_fwd_cull_kernel<2>", "rollout_fwd_cull_kernel<4>",
                                       "rollout_fwd_cull_kernel<8>", "rollout_fwd_cull_kernel<16>", "rollout_fwd_cull_kernel<32>"};
    static const char* const split[] = {"rollout_fwd_kernel<1>", "rollout_fwd_kernel<2>", "rollout_fwd_kernel<4>",
                                        "rollout_fwd_kernel<8>", "rollout_fwd_kernel<16>", "rollout_fwd_kernel<32>"};
    int lg = 0;
    while ((1 << lg) < G) ++lg;
    if (use_tpc(d)) return tpc[lg > 0 ? lg - 1 : 0];
    if (use_cull(d)) return cull[lg];
    if (use_tp(d)) return "rollout_fwd_tp_kernel";
    if (scene_layout(1, d->S).S_pad <= 256 && G >= 8 && !(d->flags & 4)) return "rollout_fwd_reg_kernel";
    return split[lg];
}

size_t doper_rollout_workspace_bytes(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    const WsLayout w = ws_layout(d);
    return w.total;
}

}  // extern "C"

namespace {
struct FusedArgs { const float* target; float* loss_sum; float* loss_pb; float* v0_bar; float* x0_bar; void* bwd_scratch; };

int rollout_fwd_impl(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                     const float* x0, const float* v0, float* xT, float* vT, float* traj_x,
                     float* traj_v, float* traj_a, void* ws, const FusedArgs* fa);
}  // namespace

extern "C" {

int doper_rollout_fwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const float* x0, const float* v0, float* xT, float* vT, float* traj_x,
                      float* traj_v, float* traj_a, void* ws) {
    return rollout_fwd_impl(stream, d, scene_ws, x0, v0, xT, vT, traj_x, traj_v, traj_a, ws, nullptr);
}

}  // extern "C"

namespace {
int rollout_fwd_impl(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                     const float* x0, const float* v0, float* xT, float* vT, float* traj_x,
                     float* traj_v, float* traj_a, void* ws, const FusedArgs* fa) {
    const bool fused = fa != nullptr;
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!scene_ws || !x0 || !v0 || !xT || !vT) return DOPER_ERR_NULL_PTR;
    if (!aligned(scene_ws, 256) || !aligned(x0, 8) || !aligned(v0, 8) || !aligned(xT, 8) || !aligned(vT, 8) ||
        (ws && !aligned(ws, 256)) || (traj_x && !aligned(traj_x, 8)) || (traj_v && !aligned(traj_v, 8)) ||
        (traj_a && !aligned(traj_a, 8)))
        return DOPER_ERR_MISALIGNED;
    const int G = auto_lanes(d);
    const bool tp = use_tp(d);
    const int tp_wpb = tp_warps_per_block(d);
    const size_t smem = tp ? (size_t)(d->per_env ? tp_wpb : 1) * scene_layout(1, d->S).S_pad * 20 + 16
                           : fwd_smem_bytes(d, G);
    const bool unstaged_ok = (use_cull(d) && !use_tpc(d)) || use_baked(d);  // these kernels can read the scene in place
    if ((int)smem > max_optin_smem() && !unstaged_ok) return DOPER_ERR_SMEM_LIMIT;
    const int K = resolve_K(d);
    FwdParams p;
    p.hl_st = log_ball_major(d) ? 1 : d->B;  // hit log is ball-major for the time-parallel kernels
    p.hl_sb = log_ball_major(d) ? d->n_steps : 1;
    p.B = d->B; p.S = d->S; p.per_env = d->per_env ? 1 : 0; p.n_steps = d->n_steps; p.K = K;
    p.force_exact = d->flags & 1;
    p.exact_idx = (d->flags >> 17) & 1;
    p.stage = 1;
    p.c = make_consts(d);
    p.scene_ws = reinterpret_cast<const unsigned char*>(scene_ws);
    p.x0 = reinterpret_cast<const float2*>(x0); p.v0 = reinterpret_cast<const float2*>(v0);
    p.xT = reinterpret_cast<float2*>(xT); p.vT = reinterpret_cast<float2*>(vT);
    p.traj_x = reinterpret_cast<float2*>(traj_x); p.traj_v = reinterpret_cast<float2*>(traj_v);
    p.traj_a = reinterpret_cast<float2*>(traj_a);
    p.ckpt = nullptr; p.hitlog = nullptr; p.sink.rec = nullptr; p.sink.count = nullptr; p.sink.cap = 0;
    p.tgt_x = p.tgt_y = 0.0f; p.loss_sum = nullptr; p.loss_pb = nullptr; p.g_v0 = nullptr; p.g_x0 = nullptr; p.done_blocks = nullptr; p.jac = nullptr;
    p.defer.entry = nullptr; p.defer.flag = nullptr; p.defer.alloc = p.defer.head = nullptr; p.defer.cap = 0; p.defer.tail_warps = 0; p.defer.streak = kDeferStreak;
    p.pipe_balls = kPipeBalls;
    if (fused) {
        if (!ws || !fa->target || !fa->loss_sum || !fa->loss_pb || !fa->v0_bar) return DOPER_ERR_NULL_PTR;
        p.tgt_x = fa->target[0]; p.tgt_y = fa->target[1];
        p.loss_sum = fa->loss_sum; p.loss_pb = fa->loss_pb;
        p.g_v0 = reinterpret_cast<float2*>(fa->v0_bar); p.g_x0 = reinterpret_cast<float2*>(fa->x0_bar);
        p.done_blocks = reinterpret_cast<unsigned int*>(reinterpret_cast<unsigned char*>(ws) + 16);  // header, zeroed below
        p.jac = reinterpret_cast<double*>(fa->bwd_scratch);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (ws) {
        const WsLayout w = ws_layout(d);
        unsigned char* base = reinterpret_cast<unsigned char*>(ws);
        p.ckpt = reinterpret_cast<float4*>(base + w.ckpt_off);
        p.hitlog = reinterpret_cast<int32_t*>(base + w.log_off);
        p.sink.rec = reinterpret_cast<HitRecord*>(base + w.rec_off);
        p.sink.count = reinterpret_cast<int*>(base);
        p.sink.cap = (int)w.cap;
        size_t zero = 256;  // record counter, block counter, hand-over counters
        p.defer.entry = nullptr; p.defer.flag = nullptr; p.defer.alloc = p.defer.head = nullptr; p.defer.cap = 0;
        if (use_defer(d) && !(traj_x || traj_v || traj_a)) {
            p.defer.entry = reinterpret_cast<DeferEntry*>(base + w.defer_off);
            p.defer.flag = reinterpret_cast<unsigned int*>(base + w.flag_off);
            // (their own 128-B line: the tail workers poll these two words, the record counter at offset 0 is hit by
            // every collision's atomic)
            p.defer.alloc = reinterpret_cast<unsigned int*>(base + 128);
            p.defer.head = reinterpret_cast<unsigned int*>(base + 132);
            p.defer.cap = (int)(d->B < 0x7fffffff ? d->B : 0x7fffffff);
            zero = w.ckpt_off;  // + the flags
        }
        if (cudaMemsetAsync(base, 0, zero, st) != cudaSuccess) return DOPER_ERR_LAUNCH;
    }
#ifdef DOPER_PIPE_TIMERS
    if (use_pipe(d)) {
#else
    if (use_pipe(d) && !traj_x && !traj_v && !traj_a) {
#endif
        return fused ? dispatch_fwd_pipe<true>(st, p, pipe_window(d)) : dispatch_fwd_pipe<false>(st, p, pipe_window(d));
    }
    if (use_baked(d)) {
        switch (baked_lanes(d)) {
            case 1: return launch_fwd_baked<1>(st, p, baked_staged(d));
            case 2: return launch_fwd_baked<2>(st, p, baked_staged(d));
            case 4: return launch_fwd_baked<4>(st, p, baked_staged(d));
            case 8: return launch_fwd_baked<8>(st, p, baked_staged(d));
            default: return launch_fwd_baked<16>(st, p, baked_staged(d));
        }
    }
    if (tp) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(rollout_fwd_tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess)
            return DOPER_ERR_SMEM_LIMIT;
        const unsigned grid = (unsigned)((d->B + tp_wpb - 1) / tp_wpb);
        rollout_fwd_tp_kernel<<<grid, tp_wpb * 32, smem, st>>>(p);
        return launch_status();
    }
    if (use_tpc(d)) {
        switch (G) {
            case 1: return DOPER_ERR_UNSUPPORTED;
            case 2: return launch_fwd_tpc<2>(st, p, smem);
            case 4: return launch_fwd_tpc<4>(st, p, smem);
            case 8: return launch_fwd_tpc<8>(st, p, smem);
            case 16: return launch_fwd_tpc<16>(st, p, smem);
            default: return launch_fwd_tpc<32>(st, p, smem);
        }
    }
    if (use_cull(d)) {
        switch (G) {
            case 1: return launch_fwd_cull<1>(st, p, smem);
            case 2: return launch_fwd_cull<2>(st, p, smem);
            case 4: return launch_fwd_cull<4>(st, p, smem);
            case 8: return launch_fwd_cull<8>(st, p, smem);
            case 16: return launch_fwd_cull<16>(st, p, smem);
            default: return launch_fwd_cull<32>(st, p, smem);
        }
    }
    const int S_pad = scene_layout(1, d->S).S_pad;
    if (S_pad <= 256 && G >= 8 && !(d->flags & 4)) {  // flags bit 2: force the shared-memory kernel
        switch (G) {
            case 8: return dispatch_fwd_reg<8>(st, p, S_pad);
            case 16: return dispatch_fwd_reg<16>(st, p, S_pad);
            default: return dispatch_fwd_reg<32>(st, p, S_pad);
        }
    }
    switch (G) {
        case 1: return launch_fwd<1>(st, p, smem);
        case 2: return launch_fwd<2>(st, p, smem);
        case 4: return launch_fwd<4>(st, p, smem);
        case 8: return launch_fwd<8>(st, p, smem);
        case 16: return launch_fwd<16>(st, p, smem);
        default: return launch_fwd<32>(st, p, smem);
    }
}
}  // namespace

extern "C" {

size_t doper_rollout_bwd_scratch_bytes(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    return (size_t)record_capacity(d) * 16 * sizeof(double) + 256;
}

}  // extern "C"

namespace {
// loss: when non-null (doper_value_and_grad), the L1 loss rides in the launch of the backward's pass 1
struct LossArgs { const float* xT; const float* target; float* loss_per_ball; float* loss_sum; float* xT_bar_out; };
int rollout_bwd_impl(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                     const void* ws, void* bwd_scratch, const float* xT_bar, const float* vT_bar,
                     float* x0_bar, float* v0_bar, const LossArgs* loss);
}  // namespace

extern "C" {

int doper_rollout_bwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const void* ws, void* bwd_scratch, const float* xT_bar, const float* vT_bar,
                      float* x0_bar, float* v0_bar) {
    return rollout_bwd_impl(stream, d, scene_ws, ws, bwd_scratch, xT_bar, vT_bar, x0_bar, v0_bar, nullptr);
}

}  // extern "C"

namespace {
int rollout_bwd_impl(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                     const void* ws, void* bwd_scratch, const float* xT_bar, const float* vT_bar,
                     float* x0_bar, float* v0_bar, const LossArgs* loss) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!scene_ws || !ws || !xT_bar || !v0_bar) return DOPER_ERR_NULL_PTR;
    if (!aligned(scene_ws, 256) || !aligned(ws, 256) || !aligned(xT_bar, 8) || !aligned(v0_bar, 8) ||
        (vT_bar && !aligned(vT_bar, 8)) || (x0_bar && !aligned(x0_bar, 8)) || (bwd_scratch && !aligned(bwd_scratch, 256)))
        return DOPER_ERR_MISALIGNED;
    if (d->n_steps == 0) return DOPER_ERR_BAD_SHAPE;
    const int K = resolve_K(d);
    const WsLayout w = ws_layout(d);
    const bool sequential = (d->flags & 2) != 0;
    if (!sequential && w.cap > 0 && !bwd_scratch) return DOPER_ERR_NULL_PTR;
    BwdParams p;
    p.B = d->B; p.S = d->S; p.per_env = d->per_env ? 1 : 0; p.n_steps = d->n_steps; p.K = K;
    p.c = make_consts(d);
    p.scene_ws = reinterpret_cast<const unsigned char*>(scene_ws);
    const unsigned char* base = reinterpret_cast<const unsigned char*>(ws);
    p.ckpt = reinterpret_cast<const float4*>(base + w.ckpt_off);
    p.hitlog = reinterpret_cast<const int32_t*>(base + w.log_off);
    p.rec = reinterpret_cast<const HitRecord*>(base + w.rec_off);
    p.rec_count = reinterpret_cast<const int*>(base);
    p.rec_cap = (int)w.cap;
    p.jac = reinterpret_cast<double*>(bwd_scratch);
    p.hl_st = log_ball_major(d) ? 1 : d->B;
    p.hl_sb = log_ball_major(d) ? d->n_steps : 1;
    p.xT_bar = reinterpret_cast<const float2*>(xT_bar);
    p.vT_bar = reinterpret_cast<const float2*>(vT_bar);
    p.x0_bar = reinterpret_cast<float2*>(x0_bar);
    p.v0_bar = reinterpret_cast<float2*>(v0_bar);
    cudaStream_t st = (cudaStream_t)stream;
    if (sequential) {  // fp32 in the oracle's operation order (bit-exact), one thread per ball
        if (loss != nullptr) {
            rc = doper_l1_loss(stream, d->B, loss->xT, loss->target, loss->loss_per_ball, loss->loss_sum, loss->xT_bar_out);
            if (rc) return rc;
        }
        const unsigned grid = (unsigned)((d->B + kBwdThreads - 1) / kBwdThreads);
        rollout_bwd_kernel<<<grid, kBwdThreads, 0, st>>>(p);
        return launch_status();
    }
    // binary64 plan: Jacobians of the collided steps, then chunk products + fold
    if (loss != nullptr) {  // ... with the L1 loss in the same launch (loss_jacobian_kernel)
        const int loss_blocks = 1 + (int)((d->B + kLossThreads - 1) / kLossThreads);
        int64_t jac_blocks = w.cap > 0 ? (4 * w.cap + kLossThreads - 1) / kLossThreads : 0;
        const int64_t max_blocks = (int64_t)sm_count() * 4;
        if (jac_blocks > max_blocks) jac_blocks = max_blocks;
        loss_jacobian_kernel<<<(unsigned)(loss_blocks + jac_blocks), kLossThreads, 0, st>>>(
            loss_blocks, d->B, reinterpret_cast<const float2*>(loss->xT), loss->target[0], loss->target[1], loss->loss_per_ball,
            loss->loss_sum, reinterpret_cast<float2*>(loss->xT_bar_out), p);
        if (launch_status() != DOPER_OK) return DOPER_ERR_LAUNCH;
    } else if (w.cap > 0) {
        int64_t blocks = (w.cap + 127) / 128;
        const int64_t max_blocks = (int64_t)sm_count() * 8;
        if (blocks > max_blocks) blocks = max_blocks;
        rollout_bwd_jacobian_kernel<<<(unsigned)blocks, 128, 0, st>>>(p);
        if (launch_status() != DOPER_OK) return DOPER_ERR_LAUNCH;
    }
    int n_chunks, chunk_len;
    fold_plan(d, &n_chunks, &chunk_len);
    const size_t smem = (size_t)n_chunks * 16 * 32 * sizeof(double);
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(rollout_bwd_fold_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const unsigned grid = (unsigned)((d->B + 31) / 32);
    rollout_bwd_fold_kernel<<<grid, 32 * n_chunks, smem, st>>>(p, n_chunks, chunk_len);
    return launch_status();
}
}  // namespace

extern "C" {

int doper_l1_loss(doper_stream_t stream, int64_t B, const float* xT, const float target[2],
                  float* loss_per_ball, float* loss_sum, float* xT_bar) {
    if (B <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!xT || !target || !loss_sum) return DOPER_ERR_NULL_PTR;
    if (!aligned(xT, 8) || (xT_bar && !aligned(xT_bar, 8))) return DOPER_ERR_MISALIGNED;
    const unsigned grid = 1 + (unsigned)((B + kLossThreads - 1) / kLossThreads);
    l1_loss_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(
        B, reinterpret_cast<const float2*>(xT), target[0], target[1], loss_per_ball, loss_sum,
        reinterpret_cast<float2*>(xT_bar));
    return launch_status();
}

int doper_value_and_grad(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                         const float* x0, const float* v0, const float target[2], void* ws, void* bwd_scratch,
                         float* scratch, float* loss_sum, float* loss_per_ball, float* xT,
                         float* vT, float* v0_bar, float* x0_bar) {
    if (!ws || !scratch || !target) return DOPER_ERR_NULL_PTR;
    if (validate_desc(d) == DOPER_OK && fused_ok(d) && !(d->flags & (1 << 24))) {
        // ONE launch: every block of the pipelined forward also evaluates the loss terms of its balls and pulls the
        // gradient back through the log it has just written (flags bit 24 keeps the separate launches).  The
        // per-ball loss terms are needed for the total: they go to loss_per_ball, or to `scratch` when that is NULL.
        FusedArgs fa;
        fa.target = target; fa.loss_sum = loss_sum; fa.loss_pb = loss_per_ball ? loss_per_ball : scratch;
        fa.v0_bar = v0_bar; fa.x0_bar = x0_bar; fa.bwd_scratch = bwd_scratch;
        if (!loss_sum || !v0_bar) return DOPER_ERR_NULL_PTR;
        return rollout_fwd_impl(stream, d, scene_ws, x0, v0, xT, vT, nullptr, nullptr, nullptr, ws, &fa);
    }
    int rc = doper_rollout_fwd(stream, d, scene_ws, x0, v0, xT, vT, nullptr, nullptr, nullptr, ws);
    if (rc) return rc;
    // the loss (per-ball terms, total, seeds sign(x_T - target) -> scratch) shares the launch of the backward's pass 1
    if (!xT || !loss_sum || !aligned(xT, 8) || !aligned(scratch, 8)) return !xT || !loss_sum ? DOPER_ERR_NULL_PTR : DOPER_ERR_MISALIGNED;
    const LossArgs la{xT, target, loss_per_ball, loss_sum, scratch};
    return rollout_bwd_impl(stream, d, scene_ws, ws, bwd_scratch, scratch, nullptr, x0_bar, v0_bar, &la);
}

size_t doper_peer_buffer_bytes(int32_t n) { return n > 0 ? peer_buffer_bytes(n) : 0; }

int doper_peer_alloc(int32_t n, void** buffer, unsigned char handle[64]) {
    if (n <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!buffer || !handle) return DOPER_ERR_NULL_PTR;
    static_assert(sizeof(cudaIpcMemHandle_t) == 64, "CUDA IPC handle size");
    void* p = nullptr;
    if (cudaMalloc(&p, peer_buffer_bytes(n)) != cudaSuccess) return DOPER_ERR_WORKSPACE;
    if (cudaMemset(p, 0, peer_buffer_bytes(n)) != cudaSuccess) { cudaFree(p); return DOPER_ERR_LAUNCH; }
    cudaIpcMemHandle_t h;
    if (cudaIpcGetMemHandle(&h, p) != cudaSuccess) { cudaFree(p); return DOPER_ERR_UNSUPPORTED; }
    memcpy(handle, &h, 64);
    *buffer = p;
    return DOPER_OK;
}

int doper_peer_open(const unsigned char handle[64], void** buffer) {
    if (!handle || !buffer) return DOPER_ERR_NULL_PTR;
    cudaIpcMemHandle_t h;
    memcpy(&h, handle, 64);
    void* p = nullptr;
    if (cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess) != cudaSuccess) return DOPER_ERR_UNSUPPORTED;
    *buffer = p;
    return DOPER_OK;
}

int doper_peer_close(void* buffer) { return cudaIpcCloseMemHandle(buffer) == cudaSuccess ? DOPER_OK : DOPER_ERR_LAUNCH; }

int doper_peer_free(void* buffer) { return cudaFree(buffer) == cudaSuccess ? DOPER_OK : DOPER_ERR_LAUNCH; }

int doper_allreduce_oneshot(doper_stream_t stream, int32_t rank, int32_t world, void* const* buffers, int32_t n,
                            const float* in, float* out, int32_t* status, double timeout_seconds) {
    if (world < 1 || world > kPeerMaxRanks || rank < 0 || rank >= world || n <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!buffers || !in || !out) return DOPER_ERR_NULL_PTR;
    PeerPtrs P;
    for (int r = 0; r < kPeerMaxRanks; ++r) P.base[r] = r < world ? reinterpret_cast<unsigned char*>(buffers[r]) : nullptr;
    for (int r = 0; r < world; ++r)
        if (!P.base[r]) return DOPER_ERR_NULL_PTR;
    const int blocks = (int)((peer_n4(n) + kPeerThreads - 1) / kPeerThreads);
    // SM clocks: ~2 GHz; <= 0 selects the default of one minute
    const double secs = timeout_seconds > 0.0 ? (timeout_seconds < 3600.0 ? timeout_seconds : 3600.0) : 60.0;
    allreduce_oneshot_kernel<<<blocks, kPeerThreads, 0, (cudaStream_t)stream>>>(P, rank, world, n, in, out, status,
                                                                               (long long)(secs * 2.0e9));
    return launch_status();
}

namespace {
int ctl_dims(const int32_t dims[4], CtlDims* d) {
    if (!dims) return DOPER_ERR_NULL_PTR;
    d->d_in = dims[0]; d->h1 = dims[1]; d->h2 = dims[2]; d->d_out = dims[3];
    if (d->d_in < 1 || d->d_in > kCtlMaxIn || d->h1 < 1 || d->h1 > kCtlMaxHidden || d->h2 < 1 || d->h2 > kCtlMaxHidden ||
        d->d_out < 2 || d->d_out > kCtlMaxOut)
        return DOPER_ERR_UNSUPPORTED;
    return DOPER_OK;
}
}  // namespace

size_t doper_controller_scratch_bytes(const int32_t dims[4], int64_t B) {
    CtlDims d;
    if (ctl_dims(dims, &d) != DOPER_OK || B <= 0) return 0;
    return (size_t)((B + kCtlBalls - 1) / kCtlBalls) * d.n_params() * sizeof(float);
}

int doper_controller_fwd(doper_stream_t stream, const int32_t dims[4], int64_t B, const float* params, const float* observations,
                         const float* velocity_init, float mass, float* impulse, float* v0) {
    CtlDims d;
    int rc = ctl_dims(dims, &d);
    if (rc) return rc;
    if (B <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!params || !observations || (v0 && !velocity_init) || (!v0 && !impulse)) return DOPER_ERR_NULL_PTR;
    if ((v0 && (!aligned(v0, 8) || !aligned(velocity_init, 8))) || !aligned(params, 4) || !aligned(observations, 4)) return DOPER_ERR_MISALIGNED;
    const size_t smem = CtlSmem::floats(d) * sizeof(float);
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(controller_fwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const unsigned grid = (unsigned)((B + kCtlBalls - 1) / kCtlBalls);
    controller_fwd_kernel<<<grid, kCtlThreads, smem, (cudaStream_t)stream>>>(d, B, params, observations,
                                                                            reinterpret_cast<const float2*>(velocity_init), mass, impulse,
                                                                            reinterpret_cast<float2*>(v0));
    return launch_status();
}

int doper_controller_bwd(doper_stream_t stream, const int32_t dims[4], int64_t B, const float* params, const float* observations,
                         const float* v0_bar, float mass, void* scratch, float* grad) {
    CtlDims d;
    int rc = ctl_dims(dims, &d);
    if (rc) return rc;
    if (B <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!params || !observations || !v0_bar || !scratch || !grad) return DOPER_ERR_NULL_PTR;
    if (!aligned(v0_bar, 8) || !aligned(scratch, 16)) return DOPER_ERR_MISALIGNED;
    const size_t smem = (CtlSmem::floats(d) + (size_t)kCtlBalls * (d.h2 + d.h1 + d.d_out)) * sizeof(float);
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    if (smem > 48 * 1024 && cudaFuncSetAttribute(controller_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const unsigned grid = (unsigned)((B + kCtlBalls - 1) / kCtlBalls);
    cudaStream_t st = (cudaStream_t)stream;
    controller_bwd_kernel<<<grid, kCtlThreads, smem, st>>>(d, B, params, observations, reinterpret_cast<const float2*>(v0_bar), mass,
                                                           reinterpret_cast<float*>(scratch));
    if (launch_status() != DOPER_OK) return DOPER_ERR_LAUNCH;
    controller_grad_reduce_kernel<<<(d.n_params() + 255) / 256, 256, 0, st>>>(d.n_params(), (int)grid, reinterpret_cast<const float*>(scratch), grad);
    return launch_status();
}

int doper_adam_clip_step(doper_stream_t stream, int32_t n, float* params, float* grad, float* exp_avg, float* exp_avg_sq, float* step,
                         float max_norm, float lr, float beta1, float beta2, float eps, float* total_norm) {
    if (n <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!params || !grad || !exp_avg || !exp_avg_sq || !step) return DOPER_ERR_NULL_PTR;
    adam_clip_kernel<<<1, 1024, 0, (cudaStream_t)stream>>>(n, params, grad, exp_avg, exp_avg_sq, step, max_norm, lr, beta1, beta2, eps, total_norm);
    return launch_status();
}

int doper_fp32_probe(doper_stream_t stream, int32_t blocks, int32_t threads, int32_t iters,
                     int32_t use_fma, float* out) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!out) return DOPER_ERR_NULL_PTR;
    fp32_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, use_fma, out);
    return launch_status();
}

}  // extern "C"
