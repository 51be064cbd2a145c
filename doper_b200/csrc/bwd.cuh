// bwd.cuh: backward rollout kernels (sequential, chunk-parallel) — part of the doper_b200 CUDA translation unit (included by kernels.cu).
#pragma once
#include "scene.cuh"

namespace doper {

// ------------------------------------------------------------------------------------------
// backward rollout
// ------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 128;

struct BwdParams {
    int64_t B;
    int S, per_env, n_steps, K;
    Consts c;
    const unsigned char* scene_ws;
    const float4* ckpt;
    const float4* hitstate;  // nullable: [n_steps][B] states at the start of collided steps
    const int32_t* hitlog;  // element (t, ball) at t * hl_st + ball * hl_sb
    int64_t hl_st, hl_sb;
    const float2* xT_bar;
    const float2* vT_bar;  // nullable
    float2* x0_bar;        // nullable
    float2* v0_bar;
    // hybrid plan of the sequential kernel: a ball that has pulled back more than defer_after collided
    // steps (sliding contact: a ~2.6 us serial chain per step) appends itself to `deferred` ([0] = count,
    // then ball ids) and leaves; rollout_bwd_chunked_kernel then redoes those balls 32 chunks in parallel
    int* deferred;         // nullable
    int defer_after;
};

constexpr int kDeferCap = 2048;  // deferred balls per launch; beyond it balls simply stay sequential

// Sequential backward, one thread per ball, bit-exact against the oracle's adjoint.  A step that
// did not collide is pulled back from its log word alone (step_vjp_free: no state, no replay);
// only a K-block that contains a collision is replayed from its checkpoint (with the logged
// indices), up to its last collision, into the shared-memory stash.
constexpr int kBwdPrefetch = 16;  // log words fetched ahead per thread (independent loads in flight)

__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_kernel(const BwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* stash = reinterpret_cast<float4*>(smem);  // [K][kBwdThreads], replay fallback only
    const int tid = threadIdx.x;
    const int64_t ball = (int64_t)blockIdx.x * kBwdThreads + tid;
    if (ball >= p.B) return;
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const float4* q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) +
                      (p.per_env ? (size_t)ball * L.S_pad : 0);
    const Consts c = p.c;
    const float dk = dfric_value(c);
    float2 g = p.xT_bar[ball];
    float gx = g.x, gy = g.y, gvx = 0.0f, gvy = 0.0f;
    if (p.vT_bar) { float2 h = p.vT_bar[ball]; gvx = h.x; gvy = h.y; }
    const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;

    const int n_ckpt = (p.n_steps + p.K - 1) / p.K;
    int collided = 0;
    for (int ck = n_ckpt - 1; ck >= 0; --ck) {
        const int t0 = ck * p.K, t1 = min(p.n_steps, t0 + p.K);
        int have = -1;  // states of steps t0 .. t0 + have are in the stash
        for (int tb = t1; tb > t0;) {
            // the log words of steps [ta, tb): kBwdPrefetch independent loads, then a latency-free loop
            const int ta = max(t0, tb - kBwdPrefetch);
            // condensed into three bit masks (bit j = step ta + j), so the processing loop stays compact
            unsigned m_hit = 0u, m_cx = 0u, m_cy = 0u;
#pragma unroll
            for (int j = 0; j < kBwdPrefetch; ++j) {
                const uint32_t w = (ta + j < tb) ? (uint32_t)hl[(size_t)(ta + j) * p.hl_st] : 0u;
                m_hit |= (w >> 31) << j;
                m_cx |= ((w >> 29) & 1u) << j;
                m_cy |= ((w >> 30) & 1u) << j;
            }
            for (int t = tb - 1; t >= ta; --t) {
                const int j = t - ta;
                if (!((m_hit >> j) & 1u)) {
                    step_vjp_free_1d(((m_cx >> j) & 1u) != 0u, dk, c, gx, gvx);
                    step_vjp_free_1d(((m_cy >> j) & 1u) != 0u, dk, c, gy, gvy);
                    continue;
                }
                if (p.deferred && ++collided > p.defer_after) {
                    const int slot = atomicAdd(p.deferred, 1);
                    if (slot < kDeferCap) { p.deferred[1 + slot] = (int)ball; return; }  // redone chunk-parallel
                    collided = -0x40000000;  // list full: stay sequential
                }
                const int32_t h = hl[(size_t)t * p.hl_st];  // collided step: the index is needed (rare reload)
                float4 st;
                if (p.hitstate) {
                    st = p.hitstate[(size_t)t * p.B + ball];  // the forward kept the state of every collided step
                } else {
                    if (have < 0) {  // first (= latest) collision of the block: replay t0 .. t, bit-identical states
                        float4 s = p.ckpt[(size_t)ck * p.B + ball];
                        for (int u = t0; u <= t; ++u) {
                            stash[(size_t)(u - t0) * kBwdThreads + tid] = s;
                            if (u < t) {
                                const int32_t hu = hl[(size_t)u * p.hl_st];
                                State s1;
                                free_step(s.x, s.y, s.z, s.w, c, s1);
                                if (hu < 0) s1 = resolve(s.x, s.y, s.z, s.w, s1, __ldg(q + (hu & kLogIdx)), c);
                                s = make_float4(s1.x, s1.y, s1.vx, s1.vy);
                            }
                        }
                        have = t - t0;
                    }
                    st = stash[(size_t)(t - t0) * kBwdThreads + tid];
                }
                const float4 seg = __ldg(q + (h & kLogIdx));
                step_vjp(st.x, st.y, st.z, st.w, true, seg, c, gx, gy, gvx, gvy);
            }
            tb = ta;
        }
    }
    p.v0_bar[ball] = make_float2(gvx, gvy);
    if (p.x0_bar) p.x0_bar[ball] = make_float2(gx, gy);
}

// Chunk-parallel backward for small batches (few balls cannot fill 148 SMs with one thread per
// ball): thread (ball, chunk) pulls the four basis cotangents e_0..e_3 back through its K-step
// chunk, which yields the chunk's transposed Jacobian M_c = J_t0^T ... J_t1-1^T (4x4; the step
// adjoint is linear in the cotangent).  Free-flight steps come from the log word alone; a chunk
// with a collision is replayed from its checkpoint first.  One thread per ball then folds
// g <- M_c g from the last chunk to the first.  Same mathematics as the sequential kernel,
// different rounding order (tolerance-checked, not bit-exact).
// No block-wide synchronisation: a ball's C_pad <= 32 chunks sit in one warp, the fold runs on warp
// shuffles.  Until a chunk meets a collision its Jacobian is block diagonal (x and y decouple in
// free flight), so only the four non-zero (position, velocity) pairs are pulled back.
// With `ball_list` ([0] = count, then ids; hybrid plan of the sequential kernel) the grid covers the
// listed balls only and `chunk_len` need not be the checkpoint period (the hit-state array is required).
__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_chunked_kernel(const BwdParams p, const int C_pad,
                                                                         const int chunk_len,
                                                                         const int* __restrict__ ball_list) {
    const int tid = threadIdx.x;
    const int bpb = kBwdThreads / C_pad;
    const int chunk = tid % C_pad, slot = tid / C_pad;
    int64_t ball = (int64_t)blockIdx.x * bpb + slot;
    if (ball_list) {
        const int count = min(ball_list[0], kDeferCap);
        ball = ball < count ? (int64_t)ball_list[1 + ball] : p.B;  // p.B = no ball
    }
    const int n_ckpt = (p.n_steps + chunk_len - 1) / chunk_len;
    const bool active = ball < p.B && chunk < n_ckpt;
    const Consts c = p.c;
    const float dk = dfric_value(c);
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    float g[4][4];  // g[k] = pull-back of e_k through this thread's chunk
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r) g[k][r] = (k == r) ? 1.0f : 0.0f;
    if (active) {
        const float4* q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) +
                          (p.per_env ? (size_t)ball * L.S_pad : 0);
        const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;
        const int t0 = chunk * chunk_len, t1 = min(p.n_steps, t0 + chunk_len);
        bool coupled = false;  // a collision has mixed x and y
        for (int tb = t1; tb > t0;) {
            const int ta = max(t0, tb - kBwdPrefetch);
            unsigned m_hit = 0u, m_cx = 0u, m_cy = 0u;
#pragma unroll
            for (int j = 0; j < kBwdPrefetch; ++j) {
                const uint32_t w = (ta + j < tb) ? (uint32_t)hl[(size_t)(ta + j) * p.hl_st] : 0u;
                m_hit |= (w >> 31) << j;
                m_cx |= ((w >> 29) & 1u) << j;
                m_cy |= ((w >> 30) & 1u) << j;
            }
            if (m_hit == 0u && !coupled) {  // free flight, decoupled: 4 (position, velocity) pairs per step
                for (int j = tb - ta - 1; j >= 0; --j) {
                    const bool cx = ((m_cx >> j) & 1u) != 0u, cy = ((m_cy >> j) & 1u) != 0u;
                    step_vjp_free_1d(cx, dk, c, g[0][0], g[0][2]);
                    step_vjp_free_1d(cx, dk, c, g[2][0], g[2][2]);
                    step_vjp_free_1d(cy, dk, c, g[1][1], g[1][3]);
                    step_vjp_free_1d(cy, dk, c, g[3][1], g[3][3]);
                }
            } else {
                for (int t = tb - 1; t >= ta; --t) {
                    const int32_t h = hl[(size_t)t * p.hl_st];
                    if (h >= 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) step_vjp_free((uint32_t)h, dk, c, g[k][0], g[k][1], g[k][2], g[k][3]);
                        continue;
                    }
                    float4 st;
                    if (p.hitstate) {
                        st = p.hitstate[(size_t)t * p.B + ball];
                    } else {  // no hit-state array: replay t0 .. t from the chunk's checkpoint (<= K steps)
                        st = p.ckpt[(size_t)chunk * p.B + ball];
                        for (int u = t0; u < t; ++u) {
                            const int32_t hu = hl[(size_t)u * p.hl_st];
                            State s1;
                            free_step(st.x, st.y, st.z, st.w, c, s1);
                            if (hu < 0) s1 = resolve(st.x, st.y, st.z, st.w, s1, __ldg(q + (hu & kLogIdx)), c);
                            st = make_float4(s1.x, s1.y, s1.vx, s1.vy);
                        }
                    }
                    const float4 seg = __ldg(q + (h & kLogIdx));
                    {   // through a copy: the out-of-line call takes its array by reference (local memory), and
                        // g itself must stay in registers for the straight-line free-flight path
                        float tmp[4][4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int r2 = 0; r2 < 4; ++r2) tmp[k][r2] = g[k][r2];
                        step_vjp_hit_n<4>(st.x, st.y, st.z, st.w, seg, c, tmp);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int r2 = 0; r2 < 4; ++r2) g[k][r2] = tmp[k][r2];
                    }
                    coupled = true;
                }
            }
            tb = ta;
        }
    }
    // fold g <- M_c g from the last chunk to the first; lane `chunk` of the ball's C_pad lanes holds M_c
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ball < p.B) {
        const float2 gg = p.xT_bar[ball];
        v[0] = gg.x; v[1] = gg.y;
        if (p.vT_bar) { const float2 h = p.vT_bar[ball]; v[2] = h.x; v[3] = h.y; }
    }
    if (C_pad <= 32) {
        for (int cc = n_ckpt - 1; cc >= 0; --cc) {
            float o[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                o[r] = (v[0] * g[0][r] + v[1] * g[1][r]) + (v[2] * g[2][r] + v[3] * g[3][r]);
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = __shfl_sync(0xffffffffu, o[r], cc, C_pad);  // chunk cc's product
        }
    } else {
        // a ball's chunks span C_pad / 32 warps: the warp holding the last chunks folds first and relays
        // the cotangent to the next one through shared memory
        __shared__ float relay[kBwdThreads / 64][4];
        const int wpb_ball = C_pad / 32, wi = chunk >> 5;
        for (int w = wpb_ball - 1; w >= 0; --w) {
            if (wi == w) {
                if (w != wpb_ball - 1) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) v[r] = relay[slot][r];
                }
                for (int cc = min(n_ckpt, 32 * (w + 1)) - 1; cc >= 32 * w; --cc) {
                    float o[4];
#pragma unroll
                    for (int r = 0; r < 4; ++r)
                        o[r] = (v[0] * g[0][r] + v[1] * g[1][r]) + (v[2] * g[2][r] + v[3] * g[3][r]);
#pragma unroll
                    for (int r = 0; r < 4; ++r) v[r] = __shfl_sync(0xffffffffu, o[r], cc & 31);
                }
                if ((chunk & 31) == 0) {
#pragma unroll
                    for (int r = 0; r < 4; ++r) relay[slot][r] = v[r];
                }
            }
            __syncthreads();
        }
    }
    if (ball < p.B && chunk == 0) {
        p.v0_bar[ball] = make_float2(v[2], v[3]);
        if (p.x0_bar) p.x0_bar[ball] = make_float2(v[0], v[1]);
    }
}


}  // namespace doper
