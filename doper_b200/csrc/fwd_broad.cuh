// fwd_broad.cuh: broad-phase forward kernels (candidate lists): lane-split and lane = time step — part of the doper_b200 CUDA translation unit (included by kernels.cu).
#pragma once
#include "fwd_full.cuh"

namespace doper {

// Broad-phase forward: same organisation as rollout_fwd_kernel<G> (G lanes per ball, scene staged
// in shared memory by TMA, sequential steps), but each ball carries a CANDIDATE LIST (sweep.cuh,
// "broad phase"): the segments that can be the closest one anywhere within rho of the list's
// centre, rho sized from the ball's current speed for about kCullPeriod steps.  Per step the sweep
// touches the list only (~5-30 segments instead of S); a displacement guard triggers the refresh,
// warp-uniformly (one lane over its rho -> every ball of the warp refreshes: no divergence on the
// O(S) scan).  Results are bit-identical to the full sweep (the list provably contains the argmin
// and all its ties).  Degenerate scenes / non-finite points take the exact path as everywhere else.
#ifndef DOPER_CULL_PERIOD
#define DOPER_CULL_PERIOD 24
#endif
constexpr int kCullPeriod = DOPER_CULL_PERIOD;  // steps a list is sized for (performance only)

template <int G>
struct CullCfg {
    static constexpr int LCAP = G == 1 ? 48 : (G == 2 ? 32 : (G == 4 ? 16 : 8));  // list entries per lane
};

// STAGED = false: the scene does not fit shared memory and is read in place (global memory / L2).
// One thread per ball (G = 1) is compiled for 7 blocks per SM (72 registers): 131,072 balls, one
// GPU's share of BASELINE configs[4], are then co-resident in ONE wave (with 96 registers the grid
// needed two, the second one 38 % full).
// GRID: dense scenes (>= 1,024 segments) refresh their lists through the scene's uniform grid.
template <int G, bool STAGED, bool GRID>
__global__ void __launch_bounds__(kFwdThreads, (G == 1 && STAGED && !GRID) ? 7 : 1) rollout_fwd_cull_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = kFwdThreads / G;  // balls per block
    constexpr int LCAP = CullCfg<G>::LCAP;
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;
    constexpr bool staged = STAGED;  // else the scene is read in place (global memory / L2): any size works,
                                     // the per-step work touches the short lists only
    const size_t scene_bytes = staged ? (size_t)n_local * L.S_pad * 20 + 16 : 0;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    unsigned short* list = reinterpret_cast<unsigned short*>(smem + scene_bytes) + tid;

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (staged) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            int64_t first_scene = p.per_env ? first_ball : 0;
            int n = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
            stage_scenes(sq, srl, p.scene_ws, L, first_scene, n, bar);
        }
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the shuffles
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    if (staged) {
        sc.q = sq + (size_t)local_scene * L.S_pad;
        sc.rl = srl + (size_t)local_scene * L.S_pad;
    } else {
        const size_t sid = p.per_env ? (size_t)ball : 0;
        sc.q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + sid * L.S_pad;
        sc.rl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + sid * L.S_pad;
    }
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;

    if (staged) mbar_wait(bar, 0);

    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int until_ckpt = 0;
    // seed of the first list: smallest estimate at the start position (any index would be valid)
    int j0 = 0;
    if (__all_sync(0xffffffffu, !sc.weird)) j0 = hint_by_estimate<G>(sc, x, y, lane);
    CullState cs;
    cs.cx = 0.f; cs.cy = 0.f; cs.rho2 = -1.0f; cs.cnt = 0; cs.far2 = 0.f;  // rho2 < 0: no list yet
    const GridView grid = grid_view(p.scene_ws, L);
    const float rho_min = sc.scale * 0.001953125f, rho_max = sc.scale * 0.125f;
    const bool exact_idx = p.exact_idx != 0;

    State s1;  // free-flight candidate of the current step
    free_step(x, y, vx, vy, c, s1);
    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += p.B; }
        }
        --until_ckpt;
        // speculate: the candidate of the NEXT step, assuming this one does not collide (independent of
        // the sweep below, so its ~100-cycle dependency chain overlaps the sweep's)
        State s2;
        const bool ok2 = free_step_fast(s1.x, s1.y, s1.vx, s1.vy, c, s2);

        const float px = s1.x, py = s1.y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        const bool fast = !sc.weird && M <= 1.0e15f;  // false for NaN / inf points
        const float ddx = px - cs.cx, ddy = py - cs.cy;
        const bool stale = fast && !(ddx * ddx + ddy * ddy <= cs.rho2);
        if (__any_sync(0xffffffffu, stale)) {  // warp-uniform refresh
            // seed of an argmin list (exact index log only; a hit list has no seed)
            if (exact_idx) j0 = list_argmin<G, LCAP>(sc, fast ? px : 0.f, fast ? py : 0.f, lane, list, kFwdThreads, cs, j0);  // cnt = 0 before the first list
            if (fast) {
                const float sx = px - x, sy = py - y;
                const float dstep = sqrtf(sx * sx + sy * sy);  // displacement of this step
                const float rho = fminf(fmaxf(1.25f * (float)kCullPeriod * dstep, rho_min), rho_max);
                cull_refresh<G, LCAP, GRID>(sc, grid, px, py, rho, c.r, lane, j0, list, kFwdThreads, cs, exact_idx);
            }
        }
        // narrow phase, estimates only: far from every wall -> no collision, nothing exact needed
        float amin = fast ? lane_list_min<G, LCAP>(sc, px, py, lane, list, kFwdThreads, cs) : 0.0f;
        amin = group_min_f<G>(amin);
        const bool near = !fast || exact_idx || !(amin > cs.far2);
        bool hit = false;
        int idx = 0;
        if (__any_sync(0xffffffffu, near)) {
            unsigned long long key = kNoKey;
            if (near) {
                if (fast) key = lane_culled<G, LCAP>(sc, px, py, lane, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M),
                                                     list, kFwdThreads, cs.cnt);
                else key = lane_exact<G>(sc, px, py, lane);
            }
            key = group_min_redux<G>(key);
            if (__any_sync(0xffffffffu, near && key == kNoKey)) {  // cannot happen (the list holds the argmin); stay safe
                const unsigned long long k2 = group_min_redux<G>(lane_exact<G>(sc, px, py, lane));
                if (key == kNoKey) key = k2;
            }
            if (near) {
                float dist;
                unpack_key(key, idx, dist);
                hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
            }
        }
        const uint32_t clipb = clip_bits(vx, vy);
        int payload = idx;
        if (hit) {
            if (writer) payload = record_collision(p.sink, ball, t, x, y, vx, vy, idx);
            s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c);
            x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        } else {
            x = px; y = py; vx = s1.vx; vy = s1.vy;
        }
        if (writer) {
            if (hl) { *hl = pack_log(payload, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += p.B; }
            if (tv) { *tv = make_float2(vx, vy); tv += p.B; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += p.B; }
        }
        if (hit || !ok2) free_step(x, y, vx, vy, c, s1);  // the speculation does not apply: redo from the real state
        else s1 = s2;
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Time-parallel broad-phase forward: G lanes per ball, LANE = TIME STEP, one candidate list per
// ball.  Per round a group (1) refreshes its list when the ball has drifted half a radius from the
// list's centre, (2) integrates up to G steps of free flight (lane k keeps step k), (3) every lane
// tests ITS step against the list: the estimate alone proves "no collision" for almost every step
// (cull_far2); the few near-wall steps get the exact distance, (4) a ballot finds the first
// collided (or out-of-list) step: the steps before it commit, the collision is resolved, the next
// round starts after it.  The per-step critical path is the free-flight chain (~100 cycles)
// instead of chain + sweep + reduction.  A ball that keeps colliding shrinks its speculation
// length to one step.  Results are bit-identical to every other forward kernel.
// The list is rebuilt once the ball has drifted sqrt(kTpcRefreshAt) * 0.98 rho from its centre: the disc is
// sized for kCullPeriod + 2 G steps, so at 0.67 of it there are still ~1.3 G steps of room for the round that
// is about to be speculated (a round that runs out of the disc is wasted).
#ifndef DOPER_TPC_REFRESH_AT
#define DOPER_TPC_REFRESH_AT 0.45f
#endif
constexpr float kTpcRefreshAt = DOPER_TPC_REFRESH_AT;
constexpr int kTpcListCap = 64;
constexpr int kContactSteps = 16;  // sequential steps per round of a ball in contact mode

// Debug build only (-DDOPER_PHASE_TIMERS, tools/phase_timers.py): per-phase clock64() totals of lane 0 of
// every warp, added into the buffer passed as traj_a (which then receives no trajectory).
#ifdef DOPER_PHASE_TIMERS
#define DOPER_PHASE(k) { const long long now_ = clock64(); phase_t[phase_cur] += now_ - phase_t0; phase_t0 = now_; phase_cur = (k); }
#else
#define DOPER_PHASE(k)
#endif

template <int G, bool GRID, int THREADS>
__global__ void __launch_bounds__(THREADS) rollout_fwd_tpc_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = THREADS / G;  // balls per block
    constexpr unsigned kLow = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G, gbase = (tid & 31) - lane;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    unsigned short* lists = reinterpret_cast<unsigned short*>(smem + (size_t)n_local * L.S_pad * 20 + 16);

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        int64_t first_scene = p.per_env ? first_ball : 0;
        int ns = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, first_scene, ns, bar);
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // absent ball: reads the last ball's inputs, never advances, writes nothing
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    sc.q = sq + (size_t)local_scene * L.S_pad;
    sc.rl = srl + (size_t)local_scene * L.S_pad;
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    const float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;  // state at step t (uniform within the group)
    const Consts c = p.c;
    const int n = p.n_steps, K = p.K;
    const int kshift = (K > 0 && (K & (K - 1)) == 0) ? (31 - __clz(K)) : -1;
    const bool exact_idx = p.exact_idx != 0;
    mbar_wait(bar, 0);

    int t = active ? 0 : n;
    int T = G;       // speculation length of the next round
    int trouble = 0; // 1: refresh before the next round no matter what; 2: sweep every segment exactly next round
    int j0 = 0;
    if (__all_sync(0xffffffffu, !sc.weird)) j0 = hint_by_estimate<G>(sc, x, y, lane);
    GroupList gl;
    gl.e = lists + (size_t)slot * kTpcListCap; gl.cap = kTpcListCap; gl.cnt = 0;
    gl.cx = 0.f; gl.cy = 0.f; gl.rho2 = -1.0f; gl.far2 = 0.f; gl.safe2 = -1.0f; gl.emin = 0.f;
    const GridView grid = grid_view(p.scene_ws, L);
    const float rho_min = sc.scale * 0.001953125f, rho_max = sc.scale * 0.125f;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
#ifdef DOPER_PHASE_TIMERS
    long long phase_t[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    long long n_rounds = 0, n_refresh = 0;
    long long phase_t0 = clock64();
    int phase_cur = 7;
    float2* const traj_a_real = nullptr;
#else
    float2* const traj_a_real = p.traj_a;
#endif

    while (__any_sync(0xffffffffu, t < n)) {
        DOPER_PHASE(0)
#ifdef DOPER_PHASE_TIMERS
        ++n_rounds;
#endif
        // ---- (0) contact mode --------------------------------------------------------------------------
        // A ball that collided at the first step of its last round (resting / sliding contact) gains
        // nothing from speculation: it runs up to kContactSteps plain sequential steps here, every lane
        // of the group redundantly (identical inputs -> identical values, so the group's state stays
        // uniform without any exchange; no warp-wide primitive inside), then sits this round out.
        bool sat_out = false;
        if (t < n && T == 1 && trouble == 0 && gl.rho2 >= 0.0f && !sc.weird) {
            int miss = 0;
            int cj = -1;  // segment of the last collision: almost always the next one too (sliding contact)
            for (int it = 0; it < kContactSteps && t < n; ++it) {
                State s1;
                if (!free_step_fast(x, y, vx, vy, c, s1)) free_step(x, y, vx, vy, c, s1);
                const float px = s1.x, py = s1.y;
                const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
                const float ddx = px - gl.cx, ddy = py - gl.cy;
                if (!(M <= 1.0e15f)) break;                                              // normal path: exact sweep
                if (!(ddx * ddx + ddy * ddy <= gl.rho2)) { trouble = 1; break; }          // normal path: refresh
                int idx = 0;
                bool hit = false, decided = false;
                if (cj >= 0 && gl.cnt <= gl.cap) {
                    // fast decision: the exact distance to segment cj, and a proof that no other list entry
                    // can tie or beat it (estimate^2 > (d + 128 u M)^2 => reference distance > d)
                    const float dj = cand_dist(px, py, sc.q[cj]);
                    const float lim0 = __fmaf_rn(kCullSafePerM, M, dj) * kCullUp;
                    const float lim = (lim0 * lim0) * kCullUp;
                    bool sole = dj == dj;
                    for (int k = 0; k < gl.cnt; ++k) {
                        const int i = gl.e[k];
                        const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
                        sole = sole && (i == cj || a > lim);
                    }
                    if (sole) { idx = cj; hit = dj <= c.r; decided = true; }
                }
                if (!decided) {
                    const float amin = glist_min(sc, px, py, gl);
                    if (exact_idx || !(amin > gl.far2)) {
                        unsigned long long key = glist_exact(sc, px, py, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M),
                                                             gl.e, gl.cap, gl.cnt);
                        if (key == kNoKey) key = lane_exact<1>(sc, px, py, 0);
                        float dist;
                        unpack_key(key, idx, dist);
                        hit = dist <= c.r;
                    }
                }
                const uint32_t clipb = clip_bits(vx, vy);
                int payload = (hit || exact_idx) ? idx : 0;
                if (active && lane == 0) {
                    if (p.ckpt && (t % K) == 0) p.ckpt[(size_t)(t / K) * p.B + ball] = make_float4(x, y, vx, vy);
                    if (hit) payload = record_collision(p.sink, ball, t, x, y, vx, vy, idx);
                }
                if (hit) { s1 = resolve_inline(x, y, vx, vy, s1, sc.q[idx], c); cj = idx; }
                x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
                if (active && lane == 0) {
                    if (hl) hl[(int64_t)t * p.hl_st] = pack_log(payload, hit, clipb);
                    const size_t o = (size_t)t * p.B + ball;
                    if (p.traj_x) p.traj_x[o] = make_float2(x, y);
                    if (p.traj_v) p.traj_v[o] = make_float2(vx, vy);
                    if (traj_a_real) traj_a_real[o] = make_float2(s1.ax, s1.ay);
                }
                ++t;
                sat_out = true;
                if (hit) miss = 0;
                else if (++miss >= 2) { T = 2; break; }  // left the wall: speculate again
            }
        }
        const bool alive = t < n && !sat_out;
        DOPER_PHASE(1)
        // ---- (1) list refresh at the current position -------------------------------------------------
        {
            const float M0 = fmaxf(sc.scale, fmaxf(fabsf(x), fabsf(y)));
            const bool fast0 = !sc.weird && M0 <= 1.0e15f;
            const float ddx = x - gl.cx, ddy = y - gl.cy;
            const bool want = alive && fast0 && (trouble == 1 || !(ddx * ddx + ddy * ddy <= kTpcRefreshAt * gl.rho2));
            if (__any_sync(0xffffffffu, want)) {
#ifdef DOPER_PHASE_TIMERS
                ++n_refresh;
#endif
                if (want && exact_idx) j0 = glist_argmin(sc, x, y, gl, j0);  // seed of an argmin list
                const float dstep = sqrtf(vx * vx + vy * vy) * c.dt;
                float rho = 1.25f * (float)(kCullPeriod + 2 * G) * dstep;  // room for this many steps at the current speed
                // ... or, in open space, most of the way to the nearest wall (known from the last scan, less what
                // the ball has moved since): the hit list of such a disc is empty, the whole disc is safe, and
                // the next scan is that much later.  Any rho is valid; this only picks a cheaper one.
                if (!exact_idx) rho = fmaxf(rho, 0.7f * ((gl.emin - sqrtf(ddx * ddx + ddy * ddy)) - c.r));
                rho = fminf(fmaxf(rho, rho_min), rho_max);
                glist_refresh<G, GRID>(sc, grid, want ? x : 0.f, want ? y : 0.f, rho, c.r, lane, gbase, j0, want, gl, exact_idx);
            }
        }
        DOPER_PHASE(2)
        // ---- (2) free flight: lane k keeps step t + k ---------------------------------------------------
        const int Tq = alive ? min(T, n - t) : 0;
        // every lane integrates from the round's start state up to ITS step (redundantly: the chain is
        // serial); the division range test is one min / max pair over the whole chain
        State my; my.x = x; my.y = y; my.vx = vx; my.vy = vy; my.ax = 0.f; my.ay = 0.f;
        const int last = min(lane, Tq - 1);
        {
            float nmin = __int_as_float(0x7f800000), nmax = 0.0f;
            for (int k = 0; k <= last; ++k) free_step_acc(my.x, my.y, my.vx, my.vy, c, my, nmin, nmax);
            const bool ok = div_range_ok(nmin, nmax, c);
            if (__any_sync(0xffffffffu, !ok)) {  // a numerator left the range of the fast division: redo exactly
                my.x = x; my.y = y; my.vx = vx; my.vy = vy; my.ax = 0.f; my.ay = 0.f;
                for (int k = 0; k <= last; ++k) free_step(my.x, my.y, my.vx, my.vy, c, my);
            }
        }
        // state at the START of my step = the end state of the lane before me
        float sx = __shfl_up_sync(0xffffffffu, my.x, 1, G), sy = __shfl_up_sync(0xffffffffu, my.y, 1, G);
        float svx = __shfl_up_sync(0xffffffffu, my.vx, 1, G), svy = __shfl_up_sync(0xffffffffu, my.vy, 1, G);
        if (lane == 0) { sx = x; sy = y; svx = vx; svy = vy; }
        DOPER_PHASE(3)
        // ---- (3) my step against the list -------------------------------------------------------------------
        const bool valid = lane < Tq;
        const float px = valid ? my.x : x, py = valid ? my.y : y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        const bool fast = !sc.weird && M <= 1.0e15f && trouble != 2;  // false for NaN / inf points
        const float ddx = px - gl.cx, ddy = py - gl.cy;
        const float dd2 = ddx * ddx + ddy * ddy;
        const bool stale = fast && !(dd2 <= gl.rho2);
        const bool safe = fast && !exact_idx && dd2 <= gl.safe2;  // inside the safe disc: no segment can be within r
        const float amin = (fast && !stale && !safe) ? glist_min(sc, px, py, gl) : 0.0f;
        const bool near = !safe && (!fast || exact_idx || !(amin > gl.far2));
        int idx = 0;
        float dist = __int_as_float(0x7f800000);
        if (valid && near && !stale) {  // exact distance and index for my own step (no warp primitives inside)
            unsigned long long key = kNoKey;
            if (fast) key = glist_exact(sc, px, py, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M), gl.e, gl.cap, gl.cnt);
            if (key == kNoKey) key = lane_exact<1>(sc, px, py, 0);  // not fast; or (cannot happen) an empty candidate set
            unpack_key(key, idx, dist);
        }
        const bool hit = valid && !stale && (dist <= c.r);  // NaN -> false (ball_sim.py:164)
        DOPER_PHASE(4)
        // ---- (4) first event of the group, commit ---------------------------------------------------------
        const unsigned evb = (__ballot_sync(0xffffffffu, valid && (stale || hit)) >> gbase) & kLow;
        const int f = evb ? (__ffs(evb) - 1) : Tq;       // lanes < f: free flight, final
        const int hit_f = __shfl_sync(0xffffffffu, (int)hit, f < Tq ? f : 0, G);  // every lane executes the shuffle
        const bool ev_hit = (f < Tq) && hit_f != 0;
        State out = my;
        float nx = __shfl_sync(0xffffffffu, my.x, f > 0 ? f - 1 : 0, G), ny = __shfl_sync(0xffffffffu, my.y, f > 0 ? f - 1 : 0, G);
        float nvx = __shfl_sync(0xffffffffu, my.vx, f > 0 ? f - 1 : 0, G), nvy = __shfl_sync(0xffffffffu, my.vy, f > 0 ? f - 1 : 0, G);
        if (f == 0) { nx = x; ny = y; nvx = vx; nvy = vy; }
        if (__any_sync(0xffffffffu, ev_hit)) {  // resolve the collision of step t + f (all lanes of the group, redundantly)
            const int src = ev_hit ? f : 0;
            const float bx = __shfl_sync(0xffffffffu, sx, src, G), by = __shfl_sync(0xffffffffu, sy, src, G);
            const float bvx = __shfl_sync(0xffffffffu, svx, src, G), bvy = __shfl_sync(0xffffffffu, svy, src, G);
            State b1;
            b1.x = __shfl_sync(0xffffffffu, my.x, src, G); b1.y = __shfl_sync(0xffffffffu, my.y, src, G);
            b1.vx = __shfl_sync(0xffffffffu, my.vx, src, G); b1.vy = __shfl_sync(0xffffffffu, my.vy, src, G);
            b1.ax = __shfl_sync(0xffffffffu, my.ax, src, G); b1.ay = __shfl_sync(0xffffffffu, my.ay, src, G);
            const int hidx = __shfl_sync(0xffffffffu, idx, src, G);
            if (ev_hit) {
                const State r = resolve(bx, by, bvx, bvy, b1, sc.q[hidx], c);
                if (lane == f) out = r;
                nx = r.x; ny = r.y; nvx = r.vx; nvy = r.vy;
            }
        }
        const int n_commit = f + (ev_hit ? 1 : 0);
        DOPER_PHASE(5)
        if (active && lane < n_commit) {
            const int ts = t + lane;
            const bool h = ev_hit && lane == f;
            const int payload = h ? record_collision(p.sink, ball, ts, sx, sy, svx, svy, idx) : (exact_idx ? idx : 0);
            if (hl) hl[(int64_t)ts * p.hl_st] = pack_log(payload, h, clip_bits(svx, svy));
            if (p.ckpt) {  // K is usually a power of two: shift and mask instead of two integer divisions
                const bool at = kshift >= 0 ? (ts & (K - 1)) == 0 : (ts % K) == 0;
                if (at) p.ckpt[(size_t)(kshift >= 0 ? ts >> kshift : ts / K) * p.B + ball] = make_float4(sx, sy, svx, svy);
            }
            const size_t o = (size_t)ts * p.B + ball;
            if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
            if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
            if (traj_a_real) traj_a_real[o] = make_float2(out.ax, out.ay);
        }
        x = nx; y = ny; vx = nvx; vy = nvy;
        t += n_commit;
        DOPER_PHASE(6)
        // ---- (5) speculation length / trouble handling for the next round ---------------------------
        if (sat_out) {
            // contact mode this round: T / trouble were set there
        } else if (f < Tq && !ev_hit) {      // the list did not cover step t + f
            trouble = (f == 0) ? min(trouble + 1, 2) : 1;
        } else {
            trouble = 0;
            if (ev_hit && f == 0) {          // colliding step after step: stop speculating, and rebuild the
                if (T != 1) trouble = 1;     // list for the (now slow) ball: contact steps scan it serially
                T = 1;
            }
            else T = min(2 * T, G);
        }
    }
    if (active && lane == 0) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
#ifdef DOPER_PHASE_TIMERS
    DOPER_PHASE(7)
    if ((threadIdx.x & 31) == 0 && p.traj_a) {
        unsigned long long* phase_out = reinterpret_cast<unsigned long long*>(p.traj_a);
        for (int k = 0; k < 8; ++k) atomicAdd(phase_out + k, (unsigned long long)phase_t[k]);
        atomicAdd(phase_out + 8, (unsigned long long)n_rounds);
        atomicAdd(phase_out + 9, (unsigned long long)n_refresh);
        // per-warp record: [total cycles, rounds, refresh passes, list-refresh cycles]
        const long long wid = ((long long)blockIdx.x * THREADS + threadIdx.x) >> 5;
        long long tot = 0;
        for (int k = 0; k < 8; ++k) tot += phase_t[k];
        phase_out[16 + 4 * wid + 0] = (unsigned long long)tot;
        phase_out[16 + 4 * wid + 1] = (unsigned long long)n_rounds;
        phase_out[16 + 4 * wid + 2] = (unsigned long long)n_refresh;
        phase_out[16 + 4 * wid + 3] = (unsigned long long)phase_t[1];
    }
#endif
}


}  // namespace doper
