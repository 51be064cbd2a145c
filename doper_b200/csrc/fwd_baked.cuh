// fwd_baked.cuh: forward rollout over the BAKED broad phase (static per-cell candidate lists, scene.cuh) —
// part of the doper_b200 CUDA translation unit (included by kernels.cu).
//
// The per-ball candidate lists of fwd_broad.cuh are rebuilt every ~35 steps by a scan of the whole scene
// (28 % of the C2 forward's warp cycles, plus the displacement guards, the "trouble" states and the
// warp-uniform refresh that makes one stale ball stall 31 others in the one-thread-per-ball kernel).  For a
// shared scene none of that depends on the ball: scene_bake_kernel builds, once per scene, the hit list of
// every cell of a uniform grid, and a step only looks its point's cell up:
//     empty cell            -> no segment within r of any point of the cell: no collision, no arithmetic
//     cell with a list      -> the estimate of every list entry (9 FMAs each); all above far2 -> no collision;
//                              else the reference arithmetic for the entries the filter keeps (sweep.cuh)
//     outside the grid      -> further than r_max + h from every segment: no collision
//     NaN / inf point, degenerate scene, r > r_max -> reference arithmetic over every segment
// The proof obligations are those of the per-ball hit lists (sweep.cuh) with the cell centre as list centre
// and rho = half diagonal / 0.98; results are bit-identical to every other forward kernel (same tests).
//
// Organisation: G lanes per ball, LANE = TIME STEP as in rollout_fwd_tpc_kernel (chain of up to G free-flight
// steps, every lane tests its own step, ballot for the first collision, commit, resolve), with nothing else
// left in the round: the critical path per step is the free-flight chain plus 1/G of a cell look-up and a
// vote.  G = 1 is one thread per ball (large batches): chain, look-up, commit per step, no warp-wide
// primitive at all.  The scene and the baked arrays are staged into shared memory by 1-D TMA bulk copies
// when they fit (STAGED), else read in place (L2).
#pragma once
#include "fwd_broad.cuh"

namespace doper {

struct BakeView {
    const uint32_t* cell_start;
    const unsigned long long* fine;  // one bit per sub-cell: its disc's hit list is not empty (scene.cuh)
    const unsigned short* items;
    float fsub;   // sub-cells per cell edge (1, 2, 4 or 8)
    int fshift;
    float x0, y0, h, inv_h, fnx, fny, far2;
    float rho2chk;  // (0.98 rho)^2: a cell's list is valid for every point this close to the cell centre
    float M;        // bound on |coordinate| of the scene and of every point within rho of a cell centre
    int nx, ok;
};

// the arrays behind a BakeHeader (in shared memory when staged, else in place)
// (with_masks = false: the caller staged the lists only; the second level is then switched off)
__device__ __forceinline__ void bake_view_arrays(BakeView& bk, const unsigned char* arrays, int S_pad, bool with_masks = true) {
    bk.cell_start = reinterpret_cast<const uint32_t*>(arrays);
    bk.items = reinterpret_cast<const unsigned short*>(arrays + kBakeItemsOff);
    bk.fine = with_masks ? reinterpret_cast<const unsigned long long*>(arrays + bake_fine_off(S_pad)) : nullptr;
}

__device__ __forceinline__ void bake_view_header(BakeView& bk, const BakeHeader& bh, int weird, float r) {
    bk.x0 = bh.x0; bk.y0 = bh.y0; bk.h = bh.h; bk.inv_h = bh.inv_h; bk.fnx = bh.fnx; bk.fny = bh.fny; bk.nx = bh.nx;
    bk.far2 = bh.far2; bk.M = bh.M;
    bk.rho2chk = (kCullRhoCheck * bh.rho) * (kCullRhoCheck * bh.rho);
    bk.ok = bh.enabled && !weird && r <= bh.r_max;
    bk.fshift = bk.fine ? bh.fine_shift : 0;  // after bake_view_arrays()
    bk.fsub = (float)(1 << bk.fshift);
}

// (hit, idx) of ONE point against the baked lists; reference semantics: hit = (min_j d_ref_j <= r), idx = its
// first index (valid only when hit).  One pass over the cell's list: an entry whose estimate^2 exceeds far2
// has d_ref > r (cull_far2's proof) and can neither collide nor tie with a colliding segment, so the
// reference arithmetic runs for the other entries only (the segments within ~r of the point: usually none,
// one, or the two that share a vertex), and the lexicographic (distance, index) minimum over those IS the
// reference's argmin whenever the step collides.
// Part 1, the cell look-up: 0 = no collision possible (outside the grid, or an empty cell), 1 = the cell's list is
// items[k0 .. k1), 2 = the baked lists do not apply to this point (NaN / inf point, degenerate scene, r > r_max).
__device__ __forceinline__ int baked_cell(const BakeView& bk, float px, float py, int& k0, int& k1) {
    const float M = fmaxf(fabsf(px), fabsf(py));
    k0 = k1 = 0;
    if (!bk.ok || !(M <= 1.0e15f)) return 2;
    // Outside the grid (box of the scene grown by r_max + h): the real distance to every segment is at least
    // r_max + h - (rounding of the look-up), and at least |p| / 2 once |p| > 2 M; d_ref >= D - 38 u max(M, |p|) > r.
    const float fx = (px - bk.x0) * bk.inv_h, fy = (py - bk.y0) * bk.inv_h;
    if (!(fx >= 0.0f && fx < bk.fnx && fy >= 0.0f && fy < bk.fny)) return 0;
    // sub-cell first (fx * fsub is exact): a clear bit = no segment within r_max of any point of the sub-cell
    const int jx = (int)(fx * bk.fsub), jy = (int)(fy * bk.fsub), sm = (1 << bk.fshift) - 1;
    const int cell = (jy >> bk.fshift) * bk.nx + (jx >> bk.fshift);
    const int bit = ((jy & sm) << bk.fshift) | (jx & sm);
    // (measured: testing the bit before the cell_start loads beats issuing the three loads together, C5 forward
    // 4.41 vs 4.77 ms — the look-up is on every warp step's path, the list on 44 % of them)
    if (bk.fine && !((bk.fine[cell] >> bit) & 1ull)) return 0;
    k0 = (int)bk.cell_start[cell];
    k1 = (int)bk.cell_start[cell + 1];
    return k1 > k0 ? 1 : 0;
}

// Part 2, the narrow phase over the cell's list (kind 1) or over the whole scene (kind 2).  One pass over the list:
// an entry whose estimate^2 exceeds far2 has d_ref > r (cull_far2's proof) and can neither collide nor tie with a
// colliding segment, so the reference arithmetic runs for the other entries only (the segments within ~r of the
// point: usually none, one, or the two that share a vertex), and the lexicographic (distance, index) minimum over
// those IS the reference's argmin whenever the step collides.
// (Measured and dropped in round 2, tools/r2_ab.sh: a sqrt-free test d2 <= max{x : sqrt(x) <= r} with the division by
// |sv|^2 through Markstein corrections on the stored reciprocal — same bits, fewer dependent instructions, but the
// compiler then predicates the cheap exact path into the hot loop: C4 forward 0.99 -> 1.31 ms, C2 0.143 -> 0.139 ms.)
__device__ __forceinline__ bool baked_list_hit(const SceneView& sc, const BakeView& bk, float px, float py, float r, int kind,
                                               int k, const int k1, int& idx) {
    idx = 0;
    if (kind == 2) {  // every pair, reference arithmetic
        float dist;
        unpack_key(lane_exact<1>(sc, px, py, 0), idx, dist);
        return dist <= r;
    }
    unsigned long long key = kNoKey;
    for (; k + 1 < k1; k += 2) {  // two entries per trip: their loads and estimates overlap
        const int i0 = bk.items[k], i1 = bk.items[k + 1];
        const float4 q0 = sc.q[i0], q1 = sc.q[i1];
        const float a0 = approx_d2(px, py, q0, sc.rl[i0]), a1 = approx_d2(px, py, q1, sc.rl[i1]);
        if (!(a0 > bk.far2)) { const unsigned long long kk = make_key(cand_dist(px, py, q0), i0); key = kk < key ? kk : key; }
        if (!(a1 > bk.far2)) { const unsigned long long kk = make_key(cand_dist(px, py, q1), i1); key = kk < key ? kk : key; }
    }
    if (k < k1) {
        const int i0 = bk.items[k];
        const float4 q0 = sc.q[i0];
        if (!(approx_d2(px, py, q0, sc.rl[i0]) > bk.far2)) { const unsigned long long kk = make_key(cand_dist(px, py, q0), i0); key = kk < key ? kk : key; }
    }
    if (key == kNoKey) return false;
    float dist;
    unpack_key(key, idx, dist);
    return dist <= r;  // NaN -> false (ball_sim.py:164)
}

__device__ __forceinline__ bool baked_hit(const SceneView& sc, const BakeView& bk, float px, float py, float r, int& idx) {
    int k0, k1;
    idx = 0;
    const int kind = baked_cell(bk, px, py, k0, k1);
    if (kind == 0) return false;
    return baked_list_hit(sc, bk, px, py, r, kind, k0, k1, idx);
}

// Scene and baked views of one environment read in place (per-environment scenes: `env` = the ball; else 0).
__device__ __forceinline__ void baked_bind_in_place(const FwdParams& p, const SceneLayout& L, int64_t env, SceneView& sc, BakeView& bk) {
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[env];
    const unsigned char* gbake = p.scene_ws + L.bake_off + (size_t)env * L.bake_stride;
    const BakeHeader bh = *reinterpret_cast<const BakeHeader*>(gbake);
    sc.q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + env * L.S_pad;
    sc.rl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + env * L.S_pad;
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
    bake_view_arrays(bk, gbake + sizeof(BakeHeader), L.S_pad);
    bake_view_header(bk, bh, sc.weird, p.c.r);
}

// Hand-over of a ball in sliding / resting contact (DeferQueue, physics.cuh).  A ball that has collided on
// kDeferStreak consecutive steps and has at least kDeferMinSteps to go is pushed to the queue with the state at the
// start of its next step; its lane stops.
constexpr int kDeferStreak = 8, kDeferMinSteps = 64;

__device__ __forceinline__ void defer_push(const DeferQueue& dq, int64_t ball, int t, float x, float y, float vx, float vy) {
    const unsigned slot = atomicAdd(dq.alloc, 1u);  // < cap = B: a ball is pushed at most once
    float4* e = reinterpret_cast<float4*>(dq.entry + slot);
    e[0] = make_float4(x, y, vx, vy);
    reinterpret_cast<int4*>(e)[1] = make_int4((int)(ball & 0xffffffffll), t, 0, (int)(ball >> 32));
    __threadfence();
    *reinterpret_cast<volatile unsigned int*>(dq.flag + slot) = 1u;
}

// One ball, one thread, steps t .. n - 1 from the state (x, y, vx, vy) at the start of step t: free flight, cell
// look-up, narrow phase, collision, log / checkpoint / record / trajectory — the loop of the one-thread-per-ball
// kernel, of its tail workers and of the drain kernel (same device functions as every forward kernel: same bits).
// (Measured and dropped for the tail workers, tools/contact_chain_probe.py: a four-entries-per-trip branch-free narrow
// phase and the collision resolved in line — the ball alone 3.11 -> 3.12 ms, C5 forward 4.25 -> 4.43, C3 1.85 -> 2.04.)
__device__ __forceinline__ void baked_run_ball(const FwdParams& p, const SceneView& sc, const BakeView& bk, const Consts& c,
                                               const int64_t ball, int t, float x, float y, float vx, float vy, const bool may_defer) {
    const int n = p.n_steps, K = p.K;
    const bool want_traj = p.traj_x != nullptr || p.traj_v != nullptr || p.traj_a != nullptr;
    const int64_t hl_step = p.hl_st;
    int32_t* hlp = p.hitlog ? p.hitlog + ball * p.hl_sb + (int64_t)t * hl_step : nullptr;  // running log pointer
    int ck_i = p.ckpt ? (t + K - 1) / K : 0;
    int next_ck = p.ckpt ? ck_i * K : n + 1;  // checkpoint counter instead of shift / mask tests
    int streak = 0;
    const bool worker = !may_defer && p.defer.entry != nullptr;  // a handed-over ball: record slots by the slab
    RecordSlab slab;
    for (; t < n; ++t) {
        State my;
        if (!free_step_fast(x, y, vx, vy, c, my)) free_step(x, y, vx, vy, c, my);  // a numerator left the fast division's range
        int idx = 0;
        const bool hit = baked_hit(sc, bk, my.x, my.y, c.r, idx);
        State out = my;
        int payload = 0;
        int slot = 0;
        if (hit) {
            // The record slot is an atomic's round trip to L2 (~1,000 cycles: a third of a collided step's serial
            // chain when the warp waits for it, profiles/r2_ncu_contact_chain.txt).  Reserved here, consumed after
            // everything else of the step.  (Inlining the resolve here was measured slower: 3.14 -> 3.31 ms for
            // the ball in sliding contact alone.)
            slot = worker ? record_reserve_slab(p.sink, slab) : record_reserve(p.sink);
            out = resolve(x, y, vx, vy, my, sc.q[idx], c);
        }
        if (t == next_ck) {
            p.ckpt[(size_t)ck_i * p.B + ball] = make_float4(x, y, vx, vy);
            ++ck_i; next_ck += K;
        }
        if (hit) payload = record_commit(p.sink, slot, ball, t, x, y, vx, vy, idx);
        if (hlp) { *hlp = pack_log(payload, hit, clip_bits(vx, vy)); hlp += hl_step; }
        if (want_traj) {
            const size_t o = (size_t)t * p.B + ball;
            if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
            if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
            if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
        }
        x = out.x; y = out.y; vx = out.vx; vy = out.vy;
        if (may_defer) {
            streak = hit ? streak + 1 : 0;
            if (streak >= p.defer.streak && n - (t + 1) >= kDeferMinSteps) {
                defer_push(p.defer, ball, t + 1, x, y, vx, vy);
                return;
            }
        }
    }
    p.xT[ball] = make_float2(x, y);
    p.vT[ball] = make_float2(vx, vy);
}

// The same ball, run by a WHOLE WARP (tail workers, drain kernel): every lane carries the same state through the same
// free flight, look-up and resolve (uniform control flow, no divergence), and the narrow phase — half of a collided
// step's serial chain when one thread walks a corner's list entry by entry — gives every list entry its own lane and
// takes the (distance, index) minimum by shuffles: the same keys, the same minimum, the same bits.  Lane 0 owns the
// stores; record slots come by the slab (physics.cuh record_reserve_slab): no atomic on the chain.
__device__ __forceinline__ void baked_run_ball_warp(const FwdParams& p, const SceneView& sc, const BakeView& bk, const Consts& c,
                                                    const int64_t ball, int t, float x, float y, float vx, float vy) {
    const int n = p.n_steps, K = p.K, lane = threadIdx.x & 31;
    const bool want_traj = p.traj_x != nullptr || p.traj_v != nullptr || p.traj_a != nullptr;
    const int64_t hl_step = p.hl_st;
    int32_t* hlp = p.hitlog ? p.hitlog + ball * p.hl_sb + (int64_t)t * hl_step : nullptr;
    int ck_i = p.ckpt ? (t + K - 1) / K : 0;
    int next_ck = p.ckpt ? ck_i * K : n + 1;
    RecordSlab slab;  // lane 0: record slots by the slab (one atomic per kRecordSlab collisions)
    for (; t < n; ++t) {
        State my;
        if (!free_step_fast(x, y, vx, vy, c, my)) free_step(x, y, vx, vy, c, my);
        int idx = 0, k0, k1;
        bool hit = false;
        const int kind = baked_cell(bk, my.x, my.y, k0, k1);
        if (kind == 1) {
            unsigned long long key = kNoKey;
            for (int k = k0 + lane; k < k1; k += 32) {
                const int i = bk.items[k];
                const float4 q = sc.q[i];
                if (!(approx_d2(my.x, my.y, q, sc.rl[i]) > bk.far2)) {
                    const unsigned long long kk = make_key(cand_dist(my.x, my.y, q), i);
                    key = kk < key ? kk : key;
                }
            }
            __syncwarp();
            key = group_min<32>(key);
            if (key != kNoKey) {
                float dist;
                unpack_key(key, idx, dist);
                hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
            }
        } else if (kind == 2) {
            hit = baked_list_hit(sc, bk, my.x, my.y, c.r, kind, k0, k1, idx);
        }
        State out = my;
        if (hit) out = resolve(x, y, vx, vy, my, sc.q[idx], c);
        if (lane == 0) {
            int payload = 0;
            if (hit) payload = record_commit(p.sink, record_reserve_slab(p.sink, slab), ball, t, x, y, vx, vy, idx);
            if (hlp) *hlp = pack_log(payload, hit, clip_bits(vx, vy));
            if (t == next_ck) p.ckpt[(size_t)ck_i * p.B + ball] = make_float4(x, y, vx, vy);
            if (want_traj) {
                const size_t o = (size_t)t * p.B + ball;
                if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
                if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
                if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
            }
        }
        if (t == next_ck) { ++ck_i; next_ck += K; }
        if (hlp) hlp += hl_step;
        x = out.x; y = out.y; vx = out.vx; vy = out.vy;
    }
    if (lane == 0) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// MAXT bounds the block size (register budget); the launcher picks blockDim.x = 32 w with w = the warps ONE SM gets
// when the batch is spread evenly over the SMs (kernels.cu balanced_block): the kernel is a latency chain per warp,
// so its duration is set by the SM with the most resident warps, and a second partial wave doubles it.
template <int G, bool STAGED, int MAXT>
__global__ void __launch_bounds__(MAXT) rollout_fwd_baked_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tail_threads = p.defer.entry ? 32 * p.defer.tail_warps : 0;  // the block's last warps: tail workers
    const int BPB = ((int)blockDim.x - tail_threads) / G;  // balls per block
    constexpr unsigned kLow = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const int tid = threadIdx.x, lane = tid % G, gbase = (tid & 31) - lane;
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    int64_t ball = (int64_t)blockIdx.x * BPB + tid / G;
    const bool active = ball < p.B && tid < (int)blockDim.x - tail_threads;
    if (!active) ball = p.B - 1;  // absent ball: reads the last ball's inputs, never advances, writes nothing
    const int64_t env = p.per_env ? ball : 0;  // per-environment scenes: ball b lives in scene b (baked per scene, read in place)
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[env];
    const unsigned char* gbake = p.scene_ws + L.bake_off + (size_t)env * L.bake_stride;
    const BakeHeader bh = *reinterpret_cast<const BakeHeader*>(gbake);

    SceneView sc;
    BakeView bk;
    __shared__ int s_normal_left;  // warps of this block that still run their own balls (tail workers poll while > 0)
    if (STAGED) {
        float4* sq = reinterpret_cast<float4*>(smem);
        float* srl = reinterpret_cast<float*>(smem + (size_t)L.S_pad * sizeof(float4));
        unsigned char* sbake = smem + (size_t)L.S_pad * 20;  // 16-byte aligned: S_pad is a multiple of 32
        const uint32_t bbytes = (uint32_t)(bake_bytes(L.S_pad) - sizeof(BakeHeader));
        uint64_t* bar = reinterpret_cast<uint64_t*>(sbake + bbytes);
        if (tid == 0) { mbar_init(bar, 1); s_normal_left = ((int)blockDim.x - tail_threads) / 32; }
        __syncthreads();
        if (tid == 0) {
            const uint32_t qb = (uint32_t)L.S_pad * sizeof(float4), rb = (uint32_t)L.S_pad * sizeof(float);
            mbar_expect_tx(bar, qb + rb + bbytes);
            const unsigned char* gq = p.scene_ws + L.q_off;
            const unsigned char* gr = p.scene_ws + L.rl_off;
            const unsigned char* gb = gbake + sizeof(BakeHeader);
            for (uint32_t o = 0; o < qb; o += kTmaChunk) tma_load_1d(reinterpret_cast<unsigned char*>(sq) + o, gq + o, min(kTmaChunk, qb - o), bar);
            for (uint32_t o = 0; o < rb; o += kTmaChunk) tma_load_1d(reinterpret_cast<unsigned char*>(srl) + o, gr + o, min(kTmaChunk, rb - o), bar);
            for (uint32_t o = 0; o < bbytes; o += kTmaChunk) tma_load_1d(sbake + o, gb + o, min(kTmaChunk, bbytes - o), bar);
        }
        sc.q = sq; sc.rl = srl;
        bake_view_arrays(bk, sbake, L.S_pad);
        sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
        mbar_wait(bar, 0);
    } else {
        sc.q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + env * L.S_pad;
        sc.rl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + env * L.S_pad;
        bake_view_arrays(bk, gbake + sizeof(BakeHeader), L.S_pad);
        sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
        if (p.defer.entry) {
            if (tid == 0) s_normal_left = ((int)blockDim.x - tail_threads) / 32;
            __syncthreads();
        }
    }
    const Consts c = p.c;
    bake_view_header(bk, bh, sc.weird, c.r);

    const float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;  // state at step t (uniform within the group)
    const int n = p.n_steps, K = p.K;
    const int kshift = (K > 0 && (K & (K - 1)) == 0) ? (31 - __clz(K)) : -1;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    int t = active ? 0 : n;
    int T = G;  // speculation length of the next round (1 after a collision at the round's first step)

    // Own balls first, then the hand-over queue.  G = 1 is one thread per ball (large batches): the step loop with
    // nothing warp-wide and nothing recomputed (baked_run_ball); G > 1: the lane = step rounds below.
    // A ball in sliding contact collides on every step: a ~2 us serial chain per step (its own ~600 dependent
    // instructions plus the divergent work of the warp's other balls) that used to set the duration of the whole
    // kernel — BASELINE configs[4]: one ball of 131,072 collides on 1,972 of 2,000 steps, and the SMs were busy 49 % of
    // the kernel (profiles/r2_ncu_fwd_baked1_c5.txt).  Such a ball is handed to a TAIL WORKER (after p.defer.streak
    // collided steps — G > 1: rounds whose first step collided — in a row): the block's last warps serve the queue from
    // the start, every other warp joins them when its own balls are done, and a worker runs ONE ball.  Workers never
    // wait for other blocks (they leave when their own block's balls are done and the queue is empty), so nothing here
    // can deadlock on residency; what is still queued when the kernel ends is finished by rollout_fwd_drain_kernel.
    const DeferQueue dq = p.defer;
    bool own = tid < (int)blockDim.x - tail_threads;  // first pass of a normal warp: its own balls
    for (;;) {
        if (!own) {
            int claimed = -1;
            if ((tid & 31) == 0) {
                for (;;) {
                    const unsigned alloc = *reinterpret_cast<volatile unsigned int*>(dq.alloc);
                    const unsigned head = *reinterpret_cast<volatile unsigned int*>(dq.head);
                    if (head < alloc) {
                        if (atomicCAS(dq.head, head, head + 1u) == head) { claimed = (int)head; break; }
                        continue;
                    }
                    if (*reinterpret_cast<volatile int*>(&s_normal_left) <= 0) break;
                    __nanosleep(1000);  // ~600 polling warps on one L2 line: a pick-up a microsecond late costs nothing
                }
            }
            claimed = __shfl_sync(0xffffffffu, claimed, 0);
            if (claimed < 0) break;
            if ((tid & 31) == 0) {
                while (*reinterpret_cast<volatile unsigned int*>(dq.flag + claimed) == 0u) __nanosleep(40);
                __threadfence();
            }
            __syncwarp();
            const float4 st = __ldcg(reinterpret_cast<const float4*>(dq.entry + claimed));  // L2: written by another SM
            const int4 m = __ldcg(reinterpret_cast<const int4*>(dq.entry + claimed) + 1);
            const int64_t eball = ((int64_t)m.w << 32) | (uint32_t)m.x;
            if (!STAGED && p.per_env) baked_bind_in_place(p, L, eball, sc, bk);
            // Scene read in place (L2 / L1): the whole warp on the ball, its list entries fetched side by side
            // (C3 forward 1.96 -> 1.64 ms).  Scene in shared memory: lane 0 alone is as fast (the shuffle
            // minimum costs what the parallel entries save: C5 forward 4.38 vs 4.50 ms).
            if (!STAGED) baked_run_ball_warp(p, sc, bk, c, eball, m.y, st.x, st.y, st.z, st.w);
            else if ((tid & 31) == 0) baked_run_ball(p, sc, bk, c, eball, m.y, st.x, st.y, st.z, st.w, false);
        } else if constexpr (G == 1) {
            if (active) baked_run_ball(p, sc, bk, c, ball, 0, x, y, vx, vy, dq.entry != nullptr);
        } else {
            const bool want_traj_g = p.traj_x != nullptr || p.traj_v != nullptr || p.traj_a != nullptr;
            int streak_g = 0;       // rounds in a row whose FIRST step collided (contact: one step per round)
            bool deferred = false;  // the ball was handed over: a tail worker writes its final state

            while (__any_sync(0xffffffffu, t < n)) {
                // ---- (1) free flight: lane k keeps step t + k ---------------------------------------------------
                const int Tq = t < n ? min(T, n - t) : 0;
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
                float sx = x, sy = y, svx = vx, svy = vy;
                if (G > 1) {
                    const float ux = __shfl_up_sync(0xffffffffu, my.x, 1, G), uy = __shfl_up_sync(0xffffffffu, my.y, 1, G);
                    const float uvx = __shfl_up_sync(0xffffffffu, my.vx, 1, G), uvy = __shfl_up_sync(0xffffffffu, my.vy, 1, G);
                    if (lane != 0) { sx = ux; sy = uy; svx = uvx; svy = uvy; }
                }
                // ---- (2) my step against the baked list of my point's cell -------------------------------------
                const bool valid = lane < Tq;
                int idx = 0;
                bool hit = false;
                hit = valid && baked_hit(sc, bk, my.x, my.y, c.r, idx);
                // ---- (3) first collision of the group, commit ---------------------------------------------------
                int f = Tq, slot_g = 0x7fffffff;
                bool ev_hit = false;
                State out = my;
                float nx = my.x, ny = my.y, nvx = my.vx, nvy = my.vy;  // G == 1: the state after my only step
                if (G > 1) {
                    const unsigned evb = (__ballot_sync(0xffffffffu, hit) >> gbase) & kLow;
                    f = evb ? (__ffs(evb) - 1) : Tq;  // lanes < f: free flight, final
                    ev_hit = f < Tq;
                    const int src = f > 0 ? f - 1 : 0;
                    nx = __shfl_sync(0xffffffffu, my.x, src, G); ny = __shfl_sync(0xffffffffu, my.y, src, G);
                    nvx = __shfl_sync(0xffffffffu, my.vx, src, G); nvy = __shfl_sync(0xffffffffu, my.vy, src, G);
                    if (f == 0) { nx = x; ny = y; nvx = vx; nvy = vy; }
                    if (ev_hit && lane == f) slot_g = record_reserve(p.sink);  // the atomic's round trip overlaps the resolve
                    if (__any_sync(0xffffffffu, ev_hit)) {  // resolve the collision of step t + f (all lanes of the group, redundantly)
                        const int s = ev_hit ? f : 0;
                        const float bx = __shfl_sync(0xffffffffu, sx, s, G), by = __shfl_sync(0xffffffffu, sy, s, G);
                        const float bvx = __shfl_sync(0xffffffffu, svx, s, G), bvy = __shfl_sync(0xffffffffu, svy, s, G);
                        State b1;
                        b1.x = __shfl_sync(0xffffffffu, my.x, s, G); b1.y = __shfl_sync(0xffffffffu, my.y, s, G);
                        b1.vx = __shfl_sync(0xffffffffu, my.vx, s, G); b1.vy = __shfl_sync(0xffffffffu, my.vy, s, G);
                        b1.ax = __shfl_sync(0xffffffffu, my.ax, s, G); b1.ay = __shfl_sync(0xffffffffu, my.ay, s, G);
                        const int hidx = __shfl_sync(0xffffffffu, idx, s, G);
                        if (ev_hit) {
                            const State r = resolve(bx, by, bvx, bvy, b1, sc.q[hidx], c);
                            if (lane == f) out = r;
                            nx = r.x; ny = r.y; nvx = r.vx; nvy = r.vy;
                        }
                    }
                } else {
                    ev_hit = hit;
                    f = hit ? 0 : Tq;
                    if (hit) {
                        out = resolve(x, y, vx, vy, my, sc.q[idx], c);
                        nx = out.x; ny = out.y; nvx = out.vx; nvy = out.vy;
                    }
                }
                const int n_commit = f + (ev_hit ? 1 : 0);
                if (active && lane < n_commit) {
                    const int ts = t + lane;
                    const bool h = ev_hit && lane == f;
                    if (G == 1 && h) slot_g = record_reserve(p.sink);
                    const int payload = h ? record_commit(p.sink, slot_g, ball, ts, sx, sy, svx, svy, idx) : 0;
                    if (hl) hl[(int64_t)ts * p.hl_st] = pack_log(payload, h, clip_bits(svx, svy));
                    if (p.ckpt) {  // K is usually a power of two: shift and mask instead of two integer divisions
                        const bool at = kshift >= 0 ? (ts & (K - 1)) == 0 : (ts % K) == 0;
                        if (at) p.ckpt[(size_t)(kshift >= 0 ? ts >> kshift : ts / K) * p.B + ball] = make_float4(sx, sy, svx, svy);
                    }
                    if (want_traj_g) {
                        const size_t o = (size_t)ts * p.B + ball;
                        if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
                        if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
                        if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
                    }
                }
                if (Tq > 0) { x = nx; y = ny; vx = nvx; vy = nvy; }
                t += n_commit;
                if (dq.entry != nullptr && Tq > 0) {  // contact hand-over, as in the one-thread-per-ball loop
                    streak_g = (ev_hit && f == 0) ? streak_g + 1 : 0;
                    if (streak_g >= dq.streak && n - t >= kDeferMinSteps) {
                        if (active && lane == 0) defer_push(dq, ball, t, x, y, vx, vy);
                        deferred = true;
                        t = n;
                    }
                }
                // colliding step after step (resting / sliding contact): speculation is wasted, one step per round
                // until the ball leaves the wall
                if (G > 1) T = (ev_hit && f == 0) ? 1 : min(2 * T, G);
            }
            if (active && lane == 0 && !deferred) {
                p.xT[ball] = make_float2(x, y);
                p.vT[ball] = make_float2(vx, vy);
            }
        }
        if (dq.entry == nullptr) break;  // no queue: a thread only has its own ball
        __syncwarp();
        if (own && (tid & 31) == 0) atomicSub(&s_normal_left, 1);
        own = false;
    }
}

// What the tail workers left in the hand-over queue (balls pushed when every worker was busy or had gone): one warp
// per entry (baked_run_ball_warp), scene and lists read in place.  Launched behind every rollout_fwd_baked_kernel<1>
// that has a queue; with an empty queue it reads two counters and exits.
__global__ void __launch_bounds__(128) rollout_fwd_drain_kernel(const FwdParams p) {
    const unsigned alloc = *p.defer.alloc, head = *p.defer.head;
    const unsigned warp = (blockIdx.x * blockDim.x + threadIdx.x) >> 5, n_warps = (gridDim.x * blockDim.x) >> 5;
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    for (unsigned e = head + warp; e < alloc; e += n_warps) {
        const float4 st = *reinterpret_cast<const float4*>(p.defer.entry + e);
        const int4 m = reinterpret_cast<const int4*>(p.defer.entry + e)[1];
        const int64_t eball = ((int64_t)m.w << 32) | (uint32_t)m.x;
        SceneView sc;
        BakeView bk;
        baked_bind_in_place(p, L, p.per_env ? eball : 0, sc, bk);
        baked_run_ball_warp(p, sc, bk, p.c, eball, m.y, st.x, st.y, st.z, st.w);
    }
}

}  // namespace doper
