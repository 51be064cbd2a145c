// fwd_full.cuh: forward rollout parameters and the full-sweep kernels (every pair every step): lane-split, register-resident, warp per ball with lane = step — part of the doper_b200 CUDA translation unit (included by kernels.cu).
#pragma once
#include "scene.cuh"

namespace doper {

// ------------------------------------------------------------------------------------------
// forward rollout
// ------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 128;

struct FwdParams {
    int64_t B;
    int S, per_env, n_steps, K, force_exact;
    int exact_idx;  // log the exact closest index at every step (default: at collisions only)
    int stage;      // 1: scene staged in shared memory (TMA); 0: read from global memory / L2 (scene too large)
    Consts c;
    const unsigned char* scene_ws;
    const float2* x0;
    const float2* v0;
    float2* xT;
    float2* vT;
    float2* traj_x;
    float2* traj_v;
    float2* traj_a;
    float4* ckpt;     // [n_ckpt][B]
    HitSink sink;     // collision records (state at the START of each collided step + segment index), appended
    int32_t* hitlog;  // element (t, ball) at t * hl_st + ball * hl_sb
    int64_t hl_st, hl_sb;
    // fused forward + loss + backward launch (rollout_fwd_pipe_kernel<U, true> only; null otherwise)
    float tgt_x, tgt_y;       // loss target
    float* loss_sum;          // sum_b |x_T - target|_1 (written by the last block to finish)
    float* loss_pb;           // per-ball loss terms (always written: the total is summed from them)
    float2* g_v0;             // dL/dv0
    float2* g_x0;             // dL/dx0 (nullable)
    unsigned int* done_blocks;  // zeroed counter in the workspace header
    // hand-over of balls in sliding / resting contact to tail-worker warps (rollout_fwd_baked_kernel<1>; null: off)
    DeferQueue defer;
    int pipe_balls;  // rollout_fwd_pipe_kernel: ball slots of a block that are used (16, or fewer: see launch_fwd_pipe)
    double* jac;                // the backward's scratch: one 4x4 Jacobian per collision record (nullable)
};

template <int G>
__global__ void __launch_bounds__(kFwdThreads) rollout_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = kFwdThreads / G;  // balls per block
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        int64_t first_scene = p.per_env ? first_ball : 0;
        int n = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, first_scene, n, bar);
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the shuffles
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    sc.q = sq + (size_t)local_scene * L.S_pad;
    sc.rl = srl + (size_t)local_scene * L.S_pad;
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;

    mbar_wait(bar, 0);

    int prev_idx = -1;  // no hint at the first step
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int until_ckpt = 0;
    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += p.B; }
        }
        --until_ckpt;
        State s1;
        free_step(x, y, vx, vy, c, s1);
        const uint32_t clipb = clip_bits(vx, vy);
        int idx; float dist;
        sweep_any<G>(sc, s1.x, s1.y, lane, prev_idx, idx, dist);
        prev_idx = idx;
        const bool hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
        const int payload = (hit && writer) ? record_collision(p.sink, ball, t, x, y, vx, vy, idx) : idx;
        if (hit) s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c);
        x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        if (writer) {
            if (hl) { *hl = pack_log(payload, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += p.B; }
            if (tv) { *tv = make_float2(vx, vy); tv += p.B; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += p.B; }
        }
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Small scenes (S_pad <= 256): every lane keeps ITS OWN segments (indices lane, lane+G, ...) in
// registers for the whole rollout, so the estimate loop is straight-line code with no loads at
// all (NV = S_pad / G visits, fully unrolled, candidates recorded in one 32-bit mask). The few
// reference-arithmetic evaluations (hint, candidates, collision) read the prepared scene through
// the read-only L1 path. Same results as rollout_fwd_kernel, by the same argument.
#ifndef DOPER_REG_THREADS
#define DOPER_REG_THREADS 128
#endif
// Block size of the register-resident kernel. (32-thread blocks spread a 4,096-ball batch more
// evenly over the 148 SMs but measured the same 0.636 ms: each warp is latency-bound on its own.)
constexpr int kRegThreads = DOPER_REG_THREADS;

// Register budget per G so that a 4,096-ball batch is co-resident in ONE wave (a second wave of a
// whole-rollout kernel doubles the time): blocks/SM = 2, 4, 7 for G = 8, 16, 32.
template <int G, int NV>
__global__ void __launch_bounds__(kRegThreads, (G == 8 ? 2 : (G == 16 ? 4 : 7)) * (128 / kRegThreads))
rollout_fwd_reg_kernel(const FwdParams p) {
    constexpr int BPB = kRegThreads / G;
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    int64_t ball = (int64_t)blockIdx.x * BPB + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the warp-wide REDUX
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const int64_t scene_id = p.per_env ? ball : 0;
    const float4* __restrict__ gq = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + scene_id * L.S_pad;
    const float* __restrict__ grl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + scene_id * L.S_pad;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[scene_id];
    SceneView sc;
    sc.q = gq; sc.rl = grl; sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale;
    sc.weird = hdr.weird | p.force_exact;

    float4 q[NV];
    float rl[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        q[j] = __ldg(gq + lane + j * G);
        rl[j] = __ldg(grl + lane + j * G);
    }

    const float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int prev_idx = 0, until_ckpt = 0;
    const int64_t stride = p.B;
    if (!sc.weird) {  // hint for the first step: smallest estimate at the initial position (any index is valid)
        float best = __int_as_float(0x7f800000);
        int bj = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float a = approx_d2(x, y, q[j], rl[j]);
            if (a < best) { best = a; bj = j; }
        }
        const unsigned long long k =
            group_min_redux<G>(((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(lane + bj * G));
        const int h = (int)(unsigned)(k & 0xffffffffull);
        prev_idx = (h >= 0 && h < p.S) ? h : 0;
    }

    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += stride; }
        }
        --until_ckpt;
        State s1;
        free_step(x, y, vx, vy, c, s1);
        const uint32_t clipb = clip_bits(vx, vy);
        const float px = s1.x, py = s1.y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        unsigned long long key;
        if (!sc.weird && M <= 1.0e15f) {
            const float kappa = kFilterKappaPerM2 * M * M;
            const float thr = __fmaf_rn(approx_d2(px, py, __ldg(gq + prev_idx), __ldg(grl + prev_idx)),
                                        kFilterOneEps, kappa);
            unsigned mask = 0u;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float a = approx_d2(px, py, q[j], rl[j]);
                mask |= (a <= thr) ? (1u << j) : 0u;
            }
            key = kNoKey;
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                const int cidx = lane + j * G;
                const unsigned long long kk = make_key(cand_dist(px, py, __ldg(gq + cidx)), cidx);
                key = kk < key ? kk : key;
            }
        } else {
            key = lane_exact<G>(sc, px, py, lane);
        }
        key = group_min_redux<G>(key);
        int idx; float dist;
        unpack_key(key, idx, dist);
        prev_idx = idx;
        const bool hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
        const int payload = (hit && writer) ? record_collision(p.sink, ball, t, x, y, vx, vy, idx) : idx;
        if (hit) s1 = resolve(x, y, vx, vy, s1, __ldg(gq + idx), c);
        x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        if (writer) {
            if (hl) { *hl = pack_log(payload, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += stride; }
            if (tv) { *tv = make_float2(vx, vy); tv += stride; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += stride; }
        }
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Time-parallel forward for small and medium batches: ONE WARP PER BALL, LANE = TIME STEP.
// Collisions are rare (~1 per 1,000 ball-steps), so the warp first integrates 32 steps of free
// flight (phase 1: O(1) work per step, every lane redundantly, lane k keeps step k), then all 32
// lanes run the closest-segment sweep of their own step at once (phase 2: the scene is read from
// shared memory as warp-wide broadcasts, no cross-lane reduction, no redundant sweep work).  A
// ballot finds the first step that hit; steps before it are committed exactly as a sequential
// rollout would have produced them (free flight until a hit is the reference's own branch), the
// hit step is resolved, and the next chunk starts right after it.  A ball that keeps colliding
// (resting against a wall) switches to the lane-split sequential sweep for a while.
// Results are bit-identical to the sequential kernels: same arithmetic per step, only scheduled
// speculatively.
constexpr int kTpEarlyHit = 6;     // a hit this early in a chunk ...
constexpr int kTpSeqSteps = 64;    // ... sends the ball to sequential mode for this many steps

__global__ void __launch_bounds__(1024) rollout_fwd_tp_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wpb = blockDim.x >> 5;  // balls per block
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const int n_local = p.per_env ? wpb : 1;
    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    const int64_t first_ball = (int64_t)blockIdx.x * wpb;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        const int n = p.per_env ? (int)min((int64_t)wpb, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, p.per_env ? first_ball : 0, n, bar);
    }
    const int64_t ball = first_ball + warp;
    const bool active = ball < p.B;  // warp-uniform
    SceneView sc;
    float x = 0.f, y = 0.f, vx = 0.f, vy = 0.f;
    if (active) {
        const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];
        const int ls = p.per_env ? warp : 0;
        sc.q = sq + (size_t)ls * L.S_pad; sc.rl = srl + (size_t)ls * L.S_pad;
        sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
        const float2 xx = p.x0[ball], vv = p.v0[ball];
        x = xx.x; y = xx.y; vx = vv.x; vy = vv.y;
    }
    mbar_wait(bar, 0);
    if (!active) return;  // whole warp leaves; no block-wide sync below

    const Consts c = p.c;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    int t = 0, prev_idx = -1, seq_left = 0;
    const int n = p.n_steps, K = p.K;

    while (t < n) {
        if (seq_left > 0) {
            // ---- sequential mode: lanes split the segments of ONE step ----
            if (p.ckpt && (t % K) == 0 && lane == 0) p.ckpt[(size_t)(t / K) * p.B + ball] = make_float4(x, y, vx, vy);
            State s1;
            free_step(x, y, vx, vy, c, s1);
            const uint32_t clipb = clip_bits(vx, vy);
            int idx; float dist;
            sweep_any<32>(sc, s1.x, s1.y, lane, prev_idx, idx, dist);
            prev_idx = idx;
            const bool hit = dist <= c.r;
            const int payload = (hit && lane == 0) ? record_collision(p.sink, ball, t, x, y, vx, vy, idx) : idx;
            if (hit) { s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c); seq_left = kTpSeqSteps; }
            x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
            if (lane == 0) {
                const size_t o = (size_t)t * p.B + ball;
                if (hl) hl[(size_t)t * p.hl_st] = pack_log(payload, hit, clipb);
                if (p.traj_x) p.traj_x[o] = make_float2(x, y);
                if (p.traj_v) p.traj_v[o] = make_float2(vx, vy);
                if (p.traj_a) p.traj_a[o] = make_float2(s1.ax, s1.ay);
            }
            ++t; --seq_left;
            continue;
        }
        // ---- speculative chunk: phase 1, free flight; lane k keeps step t+k ----
        const int T = min(32, n - t);
        float sx = x, sy = y, svx = vx, svy = vy;  // state at the START of my step
        State my; my.x = x; my.y = y; my.vx = vx; my.vy = vy; my.ax = 0.f; my.ay = 0.f;
        {
            float cx = x, cy = y, cvx = vx, cvy = vy;
            const int last = min(lane, T - 1);  // lane k integrates steps 0..k, then holds
            for (int k = 0; k < T; ++k) {
                if (k <= last) {
                    sx = cx; sy = cy; svx = cvx; svy = cvy;
                    free_step(cx, cy, cvx, cvy, c, my);
                    cx = my.x; cy = my.y; cvx = my.vx; cvy = my.vy;
                }
            }
        }
        // ---- phase 2: every lane sweeps the whole scene for its own step ----
        int idx; float dist;
        if (prev_idx < 0) prev_idx = hint_by_estimate<32>(sc, x, y, lane);  // first chunk only
        {
            const float M = fmaxf(sc.scale, fmaxf(fabsf(my.x), fabsf(my.y)));
            unsigned long long key;
            if (!sc.weird && M <= 1.0e15f) key = lane_filtered<1>(sc, my.x, my.y, 0, prev_idx, kFilterKappaPerM2 * M * M);
            else key = lane_exact<1>(sc, my.x, my.y, 0);
            unpack_key(key, idx, dist);
        }
        const bool hit = (lane < T) && (dist <= c.r);  // NaN -> false (ball_sim.py:164)
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        const int first = hits ? (__ffs(hits) - 1) : T;  // first step of the chunk that collided
        const int n_commit = min(first + 1, T);          // steps t .. t+n_commit-1 are final
        State out = my;
        if (first < T) {  // resolve the collision of step t+first (all lanes, redundantly)
            const float bx = __shfl_sync(0xffffffffu, sx, first), by = __shfl_sync(0xffffffffu, sy, first);
            const float bvx = __shfl_sync(0xffffffffu, svx, first), bvy = __shfl_sync(0xffffffffu, svy, first);
            State b1;
            b1.x = __shfl_sync(0xffffffffu, my.x, first); b1.y = __shfl_sync(0xffffffffu, my.y, first);
            b1.vx = __shfl_sync(0xffffffffu, my.vx, first); b1.vy = __shfl_sync(0xffffffffu, my.vy, first);
            b1.ax = __shfl_sync(0xffffffffu, my.ax, first); b1.ay = __shfl_sync(0xffffffffu, my.ay, first);
            const int hidx = __shfl_sync(0xffffffffu, idx, first);
            const State r = resolve(bx, by, bvx, bvy, b1, sc.q[hidx], c);
            if (lane == first) out = r;
            x = r.x; y = r.y; vx = r.vx; vy = r.vy;
            if (first < kTpEarlyHit) seq_left = kTpSeqSteps;
        } else {  // no collision: the state after the last step of the chunk
            x = __shfl_sync(0xffffffffu, my.x, T - 1); y = __shfl_sync(0xffffffffu, my.y, T - 1);
            vx = __shfl_sync(0xffffffffu, my.vx, T - 1); vy = __shfl_sync(0xffffffffu, my.vy, T - 1);
        }
        prev_idx = __shfl_sync(0xffffffffu, idx, n_commit - 1);
        if (lane < n_commit) {
            const int ts = t + lane;
            const int payload = (lane == first) ? record_collision(p.sink, ball, ts, sx, sy, svx, svy, idx) : idx;
            if (hl) hl[(size_t)ts * p.hl_st] = pack_log(payload, lane == first, clip_bits(svx, svy));
            if (p.ckpt && (ts % K) == 0) p.ckpt[(size_t)(ts / K) * p.B + ball] = make_float4(sx, sy, svx, svy);
            const size_t o = (size_t)ts * p.B + ball;
            if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
            if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
            if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
        }
        t += n_commit;
    }
    if (lane == 0) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}


}  // namespace doper
