// fwd_pipe.cuh: rollout_fwd_pipe_kernel — the forward rollout of small and medium batches on a shared scene,
// organised around the ONE thing that cannot be parallelised: a ball's chain of free-flight steps.
// Part of the doper_b200 CUDA translation unit (included by kernels.cu).
//
// Where the time of the lane = step kernels went (ncu source view of rollout_fwd_baked_kernel<8> on BASELINE
// configs[1], profiles/r2_ncu_fwd_baked8_c2.txt): the per-step dependent chain of the Euler step is ~18 FP32
// operations (~75 cycles), but a round of G speculated steps cost ~350 cycles per chain step (49 instructions
// for both axes in one thread, executed by a warp that has the scheduler almost to itself) plus ~1,500 cycles of
// cell look-ups, votes, shuffles and log stores that sit between two rounds of the same ball.
//
// Here a block serves PB ball slots (16, or 8: half a producer warp) with one PRODUCER warp and PB U / 32 CONSUMER warps
// that never wait for each other inside a window of U steps (U = 16, PB = 8 by default: 5 warps per block):
//
//   PRODUCER warp   lane = (ball, axis): the x and the y coordinate of a ball integrated by two lanes (the free-flight
//                   step acts on each axis separately: physics.cuh axis_step_acc — the same operations, hence the
//                   same bits).  Nothing but the chain: 21 dependent instructions per step (measured 96 cycles),
//                   candidate states stored to shared memory as they appear.  It runs AHEAD: window k + 1 is
//                   integrated from the last candidate of window k, on the assumption that window k holds no collision.
//   CONSUMER warps  lane = (ball, step of the window): 32 / U balls per warp, every candidate of a window has its own
//                   lane.  While the producer is busy with window k + 1 they check the candidates of window k against
//                   the baked cell lists (fwd_baked.cuh: same look-up, same narrow phase, same results), find the
//                   ball's first collision by a ballot, write the hit log / checkpoints / collision record of the
//                   steps up to it, resolve the collision (physics.cuh resolve; while the very next step collides
//                   again the resolving lane steps the ball itself: contact mode) and post the state after it to a
//                   mailbox; the producer restarts that ball two windows later and the consumers drop the window that
//                   was integrated from the superseded state (epoch tag).
//
// One named barrier per window (all warps of the block) orders the shared-memory traffic: two candidate buffers
// ([ball][step], so that the consumers' lanes read consecutive 16-byte words) and two restart mailboxes, all indexed
// by window parity, so that nothing is read in the interval in which it is written.  A collision costs its ball the
// rest of its window and the next one (~1.5 U steps); collisions are ~0.15 % of the steps of the BASELINE batches.
// Every value comes out of the same device functions as in every other forward kernel: results are bit-identical
// (tests/test_parity_gpu.py::test_pipelined_*).  Measurements, the ablation and the variants that lost: DESIGN.md §10.
#pragma once
#include "bwd.cuh"
#include "fwd_baked.cuh"

namespace doper {

// Debug build only (-DDOPER_PIPE_TIMERS, tools/pipe_timers.py): clock64() totals per phase of lane 0 of every warp,
// added into the buffer passed as traj_a (the launcher then ignores the trajectory pointers).
#ifdef DOPER_PIPE_TIMERS
#define PIPE_T0() long long pt_acc[8] = {0, 0, 0, 0, 0, 0, 0, 0}; long long pt_last = clock64()
#define PIPE_MARK(ph) { const long long pt_now = clock64(); pt_acc[ph] += pt_now - pt_last; pt_last = pt_now; }
#define PIPE_FLUSH(base) if (lane == 0 && p.traj_a) { unsigned long long* pt_d = reinterpret_cast<unsigned long long*>(p.traj_a); \
    for (int pt_i = 0; pt_i < 8; ++pt_i) atomicAdd(pt_d + (base) + pt_i, (unsigned long long)pt_acc[pt_i]); }
#else
#define PIPE_T0()
#define PIPE_MARK(ph)
#define PIPE_FLUSH(base)
#endif

constexpr int kPipeBalls = 16;  // balls per group (= per producer warp)

template <int U>
struct PipeShared {  // one per group
    float4 slot[2][kPipeBalls][U + 1];  // [window parity][ball][0: window start, j: after the window's step j - 1] = (x, vx, y, vy);
                                        // ball-major: the consumers' lanes (ball, step) read consecutive 16-byte words
    int t_start[2][kPipeBalls];         // first step of the window
    int epoch_w[2][kPipeBalls];         // restart epoch the window was integrated in
    float4 mail_state[2][kPipeBalls];   // restart state (x, vx, y, vy) posted by a consumer while it read window parity p
    int mail_t[2][kPipeBalls];
    int mail_epoch[2][kPipeBalls];
    int fin_at[16];                     // per consumer warp: the barrier index from which on its balls are all final (0: not yet)
};

__device__ __forceinline__ void named_barrier(int id, int threads) {
    __syncwarp();  // bar.sync counts warps: every participating warp must arrive converged
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// U: window length = steps the producer integrates between two hand-overs.  Warps per group: 1 producer + U / 2
// consumers; a consumer warp owns 32 / U balls and gives every (ball, step of the window) its own lane, so that the
// consumers' latency per window does not grow with U while the producer's does: U = 16 balances the two.
// FUSED: when the block's 16 balls have reached step n it also evaluates their loss terms and pulls the loss gradient
// back through the hit log it has just written (binary64 chunk products + fold, bwd.cuh; collision Jacobians from the
// block's own records, in place), so that forward + L1 loss + backward are ONE launch: the backward of a block
// overlaps the forward of the blocks that are still running, and a small batch (the reference's 16 balls) costs one
// launch instead of four.  The last block to finish adds the per-ball loss terms in l1_loss_kernel's order.
// (The fused variant's binary64 tail takes 168 registers per thread: one block per SM.  It is therefore used for
// batches of at most 148 blocks only — kFusedMaxBalls — where a second block per SM would not be used anyway; for
// BASELINE configs[1] (256 blocks) the separate backward launches are faster: measured 0.151 ms per step against
// 0.164 fused with the registers capped for two blocks per SM, 0.177 uncapped.)
// PB: ball slots per block (16: a full producer warp; 8: half of it, for U = 16 — a collision stalls the window barrier
// of PB - 1 other balls, and the launch ends with the block that collected the most collisions).
template <int U, bool FUSED, int PB = kPipeBalls>
__global__ void __launch_bounds__(32 * (1 + PB * U / 32), (U <= 16 && !FUSED) ? (PB == 8 ? 4 : 2) : 1)
    rollout_fwd_pipe_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    static_assert(U == 8 || U == 16 || U == 32, "window length");
    static_assert(PB == kPipeBalls || (PB == 8 && U <= 16 && !FUSED), "ball slots per block");
    constexpr int CW = PB * U / 32;    // consumer warps
    constexpr int GT = 32 * (1 + CW);  // threads per group = per block
    constexpr unsigned kSeg = U == 32 ? 0xffffffffu : ((1u << U) - 1u);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const bool producer = warp == 0;
    const SceneLayout L = scene_layout(1, p.S);
    const uint32_t scene_bytes = (uint32_t)L.S_pad * 20;
    const uint32_t bbytes = (uint32_t)bake_lists_bytes(L.S_pad);  // cell lists only: the sub-cell masks do not pay here (measured)
    unsigned char* sbake = smem + scene_bytes;  // 16-byte aligned: S_pad is a multiple of 32
    uint64_t* bar = reinterpret_cast<uint64_t*>(sbake + bbytes);
    PipeShared<U>& sh = *reinterpret_cast<PipeShared<U>*>(sbake + bbytes + 16);

    const int64_t first = (int64_t)blockIdx.x * p.pipe_balls;  // (slots >= pipe_balls stay empty)
    const int n = p.n_steps;
    const Consts c = p.c;

    if (tid == 0) mbar_init(bar, 1);
    if (tid < kPipeBalls) {
        sh.mail_epoch[0][tid] = 0; sh.mail_epoch[1][tid] = 0;
        sh.fin_at[tid] = 0;
    }
    __syncthreads();
    if (tid == 0) {  // scene + baked lists -> shared memory (1-D TMA bulk copies); only the consumers wait for them
        const uint32_t qb = (uint32_t)L.S_pad * sizeof(float4), rb = (uint32_t)L.S_pad * sizeof(float);
        mbar_expect_tx(bar, qb + rb + bbytes);
        const unsigned char* gq = p.scene_ws + L.q_off;
        const unsigned char* gr = p.scene_ws + L.rl_off;
        const unsigned char* gb = p.scene_ws + L.bake_off + sizeof(BakeHeader);
        for (uint32_t o = 0; o < qb; o += kTmaChunk) tma_load_1d(smem + o, gq + o, min(kTmaChunk, qb - o), bar);
        for (uint32_t o = 0; o < rb; o += kTmaChunk) tma_load_1d(smem + qb + o, gr + o, min(kTmaChunk, rb - o), bar);
        for (uint32_t o = 0; o < bbytes; o += kTmaChunk) tma_load_1d(sbake + o, gb + o, min(kTmaChunk, bbytes - o), bar);
    }

    // every warp leaves after the first barrier B_j at which all consumer warps had declared their balls final
    auto group_done = [&](int j) {
        bool all = true;
#pragma unroll
        for (int w = 0; w < CW; ++w) { const int f = sh.fin_at[w]; all = all && f != 0 && f <= j; }
        return all;
    };

    if (producer) {
        // ---------------------------------------------------------------------------------- producer
        const int b = lane >> 1, axis = lane & 1;
        const int64_t ball = first + b;
        const bool active = b < p.pipe_balls && ball < p.B;
        const float att = axis ? c.ay : c.ax;
        float pos = 0.0f, vel = 0.0f;
        if (active) {
            const float2 xx = p.x0[ball], vv = p.v0[ball];
            pos = axis ? xx.y : xx.x; vel = axis ? vv.y : vv.x;
        }
        int t = active ? 0 : n, epoch = 0;
        PIPE_T0();
        for (int k = 0;; ++k) {
            const int buf = k & 1;
            const int me = sh.mail_epoch[buf][b];
            if (me > epoch) {  // a consumer resolved a collision of this ball two windows ago: restart after it
                const float4 s = sh.mail_state[buf][b];
                pos = axis ? s.z : s.x; vel = axis ? s.w : s.y;
                t = sh.mail_t[buf][b];
                epoch = me;
            }
            if (axis == 0) { sh.t_start[buf][b] = t; sh.epoch_w[buf][b] = epoch; }
            float2* out = reinterpret_cast<float2*>(&sh.slot[buf][b][0]) + axis;
            constexpr int kSlotStride = 2;  // float2 per slot
            const float p0 = pos, v0 = vel;
            out[0] = make_float2(pos, vel);
            PIPE_MARK(0);
            if (t < n) {
                float nmin = __int_as_float(0x7f800000), nmax = 0.0f;
#pragma unroll
                for (int j = 1; j <= U; ++j) {
                    axis_step_acc(pos, vel, att, c, nmin, nmax);
                    out[j * kSlotStride] = make_float2(pos, vel);
                }
                if (!div_range_ok(nmin, nmax, c)) {  // a numerator left the range of the fast division: redo exactly
                    pos = p0; vel = v0;
#pragma unroll 1
                    for (int j = 1; j <= U; ++j) {
                        axis_step(pos, vel, att, c);
                        out[j * kSlotStride] = make_float2(pos, vel);
                    }
                }
                t = min(t + U, n);
            }
            PIPE_MARK(1);
            named_barrier(1, GT);  // B_k: window k is complete, the consumers have finished window k - 1
            const bool done = group_done(k);
            PIPE_MARK(2);
#ifdef DOPER_PIPE_TIMERS
            pt_acc[7] += 1;
#endif
            if (done) break;
        }
        PIPE_FLUSH(0);
        if (!FUSED) return;
    } else {
        // -------------------------------------------------------------------------------------- consumers
        const int cw = warp - 1;
        constexpr int BPW = 32 / U;                  // balls per consumer warp
        const int bl = lane / U, j = lane % U;       // my ball within the warp, my step within the window
        const int b = cw * BPW + bl;                 // my ball within the group
        const int64_t ball = first + b;
        const bool active = b < p.pipe_balls && ball < p.B;
        SceneView sc;
        BakeView bk;
        {
            const SceneHeader hdr = *reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off);
            const BakeHeader bh = *reinterpret_cast<const BakeHeader*>(p.scene_ws + L.bake_off);
            sc.q = reinterpret_cast<const float4*>(smem);
            sc.rl = reinterpret_cast<const float*>(smem + (size_t)L.S_pad * sizeof(float4));
            sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
            bake_view_arrays(bk, sbake, L.S_pad, false);
            bake_view_header(bk, bh, sc.weird, c.r);
        }
        mbar_wait(bar, 0);
        const int K = p.K;
        const int kshift = (K > 0 && (K & (K - 1)) == 0) ? (31 - __clz(K)) : -1;
        int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
        int committed = active ? 0 : n;  // steps of my ball that are final (the same value in every lane of the ball)
        int epoch = 0;
        bool finished = false;
        // the collided step of this thread whose record is still to be written (see below)
        bool pr_on = false;
        int pr_slot = 0, pr_t = 0, pr_idx = 0;
        float pr_x = 0.0f, pr_y = 0.0f, pr_vx = 0.0f, pr_vy = 0.0f;
        auto flush_record = [&]() {
            if (!pr_on) return;
            const int pl = record_commit(p.sink, pr_slot, ball, pr_t, pr_x, pr_y, pr_vx, pr_vy, pr_idx);
            if (hl) hl[(int64_t)pr_t * p.hl_st] = pack_log(pl, true, clip_bits(pr_vx, pr_vy));
            pr_on = false;
        };
        PIPE_T0();
        for (int k = 0;; ++k) {
            named_barrier(1, GT);  // B_k
            const bool done = group_done(k);
            PIPE_MARK(0);
            if (done) break;
            flush_record();  // a collision of the previous window
            const int buf = k & 1;
#if defined(DOPER_PIPE_ABLATE) && DOPER_PIPE_ABLATE == 0
            if (k >= (n + U - 1) / U) { finished = true; if (lane == 0) sh.fin_at[cw] = k + 1; }  // timing experiment: barriers only
            continue;
#endif
            const int ts0 = sh.t_start[buf][b];
            const bool fresh = committed < n && sh.epoch_w[buf][b] == epoch;  // else: integrated from a superseded state
            const int nv = fresh ? min(U, n - ts0) : 0;                       // steps of the window that exist
            const float4 s = sh.slot[buf][b][j], cand = sh.slot[buf][b][j + 1];  // start state and candidate of my step
            // (1) my candidate against the baked lists
            int idx = 0, k0 = 0, k1 = 0;
            const int kind = j < nv ? baked_cell(bk, cand.x, cand.z, k0, k1) : 0;
            PIPE_MARK(1);
            bool myhit = false;
#if defined(DOPER_PIPE_ABLATE) && DOPER_PIPE_ABLATE == 1
            if (kind == 77) myhit = true;  // timing experiment: no narrow phase
#else
            if (kind != 0) myhit = baked_list_hit(sc, bk, cand.x, cand.z, c.r, kind, k0, k1, idx);
#endif
#ifdef DOPER_PIPE_TIMERS
            __syncwarp();
            { const long long pt_now = clock64(); const long long d = pt_now - pt_last; pt_acc[4] += d; pt_last = pt_now;
              if (__any_sync(0xffffffffu, kind != 0)) pt_acc[5] += 1;
              if (d > 1000) pt_acc[6] += 1; }
#endif
            // (2) the ball's first collision in the window
            const unsigned hm = (__ballot_sync(0xffffffffu, myhit) >> (bl * U)) & kSeg;
            const int f = hm ? (__ffs(hm) - 1) : U;
            const bool hit = f < nv;
            const int n_commit = hit ? f + 1 : nv;
            State r;
            r.x = r.y = r.vx = r.vy = r.ax = r.ay = 0.0f;
            int extra = 0;
            if (hit && j == f) {  // my step: resolve it, post the state after it
                State s1;
                free_step(s.x, s.z, s.y, s.w, c, s1);  // the candidate again, with its acceleration (resolve needs it)
                r = resolve(s.x, s.z, s.y, s.w, s1, sc.q[idx], c);
                // The record's slot is an atomic's round trip to L2, and whoever consumes it waits the trip out (a call
                // in between does too: scoreboards do not survive one).  So the record of a collided step stays PENDING
                // in this thread's registers — slot reserved, nothing consumed — until the thread's next collision or
                // the next window, and record + log word are written then (flush_record).
                flush_record();
                pr_slot = record_reserve(p.sink);
                pr_on = true; pr_t = ts0 + f; pr_idx = idx; pr_x = s.x; pr_y = s.z; pr_vx = s.y; pr_vy = s.w;
                // Contact mode: while the very next step collides again (resting / sliding contact, a bounce in a corner),
                // this lane steps the ball itself — a restart through the producer costs ~1.5 windows per collision.  At
                // most U steps per window, so that the other balls of the group keep moving.
                State cur = r;
                int tn = ts0 + f + 1;
                while (extra < U && tn < n) {
                    State c1;
                    free_step(cur.x, cur.y, cur.vx, cur.vy, c, c1);
                    int i2;
                    if (!baked_hit(sc, bk, c1.x, c1.y, c.r, i2)) break;  // free flight again: the producer takes over from `cur`
                    if (p.ckpt) {
                        const bool at = kshift >= 0 ? (tn & (K - 1)) == 0 : (tn % K) == 0;
                        if (at) p.ckpt[(size_t)(kshift >= 0 ? tn >> kshift : tn / K) * p.B + ball] = make_float4(cur.x, cur.y, cur.vx, cur.vy);
                    }
                    const State r2 = resolve(cur.x, cur.y, cur.vx, cur.vy, c1, sc.q[i2], c);
                    flush_record();  // the previous collided step: its slot arrived during this step's arithmetic
                    pr_slot = record_reserve(p.sink);
                    pr_on = true; pr_t = tn; pr_idx = i2; pr_x = cur.x; pr_y = cur.y; pr_vx = cur.vx; pr_vy = cur.vy;
                    cur = r2;
                    ++tn; ++extra;
                }
                if (extra > 0 && tn == n) { p.xT[ball] = make_float2(cur.x, cur.y); p.vT[ball] = make_float2(cur.vx, cur.vy); }
                sh.mail_state[buf][b] = make_float4(cur.x, cur.vx, cur.y, cur.vy);
                sh.mail_t[buf][b] = tn;
                sh.mail_epoch[buf][b] = epoch + 1;
            }
            extra = __shfl_sync(0xffffffffu, extra, bl * U + (hit ? f : 0));  // steps the contact loop committed for my ball
            if (!hit) extra = 0;
            PIPE_MARK(2);
            // (3) commit my step if it lies before the collision or is the collided one
#if defined(DOPER_PIPE_ABLATE) && DOPER_PIPE_ABLATE <= 2
            if (j < n_commit && ts0 + j + 1 == n) {  // timing experiment: no log / checkpoint stores
#else
            if (j < n_commit) {
#endif
                const int ts = ts0 + j;
                const bool h = hit && j == f;
                if (hl && !h) hl[(int64_t)ts * p.hl_st] = pack_log(0, false, clip_bits(s.y, s.w));  // (a collided step: flush_record)
                if (p.ckpt) {
                    const bool at = kshift >= 0 ? (ts & (K - 1)) == 0 : (ts % K) == 0;
                    if (at) p.ckpt[(size_t)(kshift >= 0 ? ts >> kshift : ts / K) * p.B + ball] = make_float4(s.x, s.z, s.y, s.w);
                }
                if (ts + 1 == n) {  // the rollout's last step: final state
                    if (h) { p.xT[ball] = make_float2(r.x, r.y); p.vT[ball] = make_float2(r.vx, r.vy); }
                    else { p.xT[ball] = make_float2(cand.x, cand.z); p.vT[ball] = make_float2(cand.y, cand.w); }
                }
            }
            if (fresh) committed = ts0 + n_commit + extra;
            if (hit) ++epoch;
            if (!finished) {
                finished = __all_sync(0xffffffffu, committed >= n);
                if (finished && lane == 0) sh.fin_at[cw] = k + 1;
            }
            PIPE_MARK(3);
#ifdef DOPER_PIPE_TIMERS
            pt_acc[7] += 1;
#endif
        }
        flush_record();
        PIPE_FLUSH(8);
    }  // consumers
    if (!FUSED) return;

    // ---------------------------------------------------------------------------- fused loss + backward of this block
    named_barrier(1, GT);  // every warp left its loop after the same barrier; the block's log, records and x_T are written
    constexpr int NC = 16;  // chunks per ball: thread (ball = tid & 15, chunk = tid >> 4), 256 of the block's threads
    double* fold = reinterpret_cast<double*>(smem);  // [NC][16 entries][16 balls]: over the staged scene (no longer needed)
    BwdParams bp;
    bp.B = p.B; bp.S = p.S; bp.per_env = 0; bp.n_steps = n; bp.K = p.K; bp.c = c; bp.scene_ws = p.scene_ws;
    bp.ckpt = p.ckpt; bp.rec = p.sink.rec; bp.rec_count = p.sink.count; bp.rec_cap = p.sink.cap; bp.jac = nullptr;
    bp.hitlog = p.hitlog; bp.hl_st = p.hl_st; bp.hl_sb = p.hl_sb;
    bp.xT_bar = nullptr; bp.vT_bar = nullptr; bp.x0_bar = nullptr; bp.v0_bar = nullptr;
    const ConstsD cd = make_consts_d(c);
    const int chunk_len = (n + NC - 1) / NC;
    // (a) the block's collision Jacobians, one thread per collided step (a binary64 Jacobian is a ~5 us dependent
    //     chain: in line, a chunk with several collisions would serialise them): the threads list the records their
    //     chunks refer to, then every listed record gets its own thread; matrices go to the backward's scratch
    __shared__ int jl_count;
    int* jlist = reinterpret_cast<int*>(&sh);  // the window buffers are free now
    constexpr int kJlCap = (int)(sizeof(sh.slot) / sizeof(int));
    const int bb = tid & 15, chunk = tid >> 4;
    const int64_t fb = first + bb;
    const int t0 = chunk * chunk_len, t1 = min(n, t0 + chunk_len);
    const bool mine = tid < 16 * NC && bb < p.pipe_balls && fb < p.B && t0 < t1;
    __shared__ double t16[kPowTableDoubles];
    if (tid == 0) { jl_count = 0; free_pow_table(cd, t16); }
    named_barrier(1, GT);
    if (p.jac != nullptr && mine) {
        const int32_t* hlb = p.hitlog + (size_t)fb * p.hl_sb;
        for (int t = t0; t < t1; ++t) {
            const uint32_t wv = (uint32_t)hlb[(size_t)t * p.hl_st];
            if ((wv & kLogHit) && !(wv & kLogNoRecord)) {
                const int k = atomicAdd(&jl_count, 1);
                if (k < kJlCap) jlist[k] = (int)(wv & kLogSeg);
            }
        }
    }
    named_barrier(1, GT);
    const int n_jac = jl_count;
    if (p.jac != nullptr && n_jac <= kJlCap) {
        for (int w = tid; w < 4 * n_jac; w += GT) {  // four threads per record: one column each
            const int slot = jlist[w >> 2], col = w & 3;
            const float4* rr = reinterpret_cast<const float4*>(p.sink.rec + slot);
            const int4 m = reinterpret_cast<const int4*>(rr)[1];
            const int64_t rb = ((int64_t)m.w << 32) | (uint32_t)m.x;
            double J[16];
            hit_jacobian_d(rr[0], __ldg(scene_q(bp, rb) + m.z), __ldg(scene_s1(bp, rb) + m.z), cd, J, col, col + 1);
            double* out = p.jac + (size_t)slot * 16 + col;
#pragma unroll
            for (int k = 0; k < 4; ++k) out[4 * k] = J[4 * k + col];
        }
        bp.jac = p.jac;
    }
    named_barrier(1, GT);
    // (b) chunk products
    if (tid < 16 * NC) {
        double g[4][4];
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int r = 0; r < 4; ++r) g[k][r] = (k == r) ? 1.0 : 0.0;
        if (mine) chunk_product(bp, cd, fb, t0, t1, g, t16);
#pragma unroll
        for (int k = 0; k < 4; ++k)
#pragma unroll
            for (int r = 0; r < 4; ++r) fold[((size_t)chunk * 16 + (k * 4 + r)) * 16 + bb] = g[k][r];
    }
    named_barrier(1, GT);
    if (tid < p.pipe_balls && first + tid < p.B) {
        const int64_t fb = first + tid;
        const float2 xT = p.xT[fb];
        const float ex = xT.x - p.tgt_x, ey = xT.y - p.tgt_y;
        p.loss_pb[fb] = fabsf(ex) + fabsf(ey);                      // ball_sim.py:286
        double v[4] = {(double)sgnf(ex), (double)sgnf(ey), 0.0, 0.0};  // seed: d|e| / de
        for (int cc = NC - 1; cc >= 0; --cc) {
            const double* P = fold + (size_t)cc * 16 * 16 + tid;
            double o[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                o[r] = fma(v[0], P[(0 * 4 + r) * 16], fma(v[1], P[(1 * 4 + r) * 16], fma(v[2], P[(2 * 4 + r) * 16], v[3] * P[(3 * 4 + r) * 16])));
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = o[r];
        }
        p.g_v0[fb] = make_float2((float)v[2], (float)v[3]);
        if (p.g_x0) p.g_x0[fb] = make_float2((float)v[0], (float)v[1]);
    }
    // the loss total: the last block to get here adds the per-ball terms in l1_loss_kernel's order (same bits)
    __shared__ int last_s;
    __threadfence();
    named_barrier(1, GT);
    if (tid == 0) last_s = atomicAdd(p.done_blocks, 1u) == gridDim.x - 1 ? 1 : 0;
    named_barrier(1, GT);
    if (!last_s) return;
    __threadfence();
    float* part = reinterpret_cast<float*>(smem);
    if (tid < 256) {
        float acc = 0.0f;
        for (int64_t q = tid; q < p.B; q += 256) acc += __ldcg(p.loss_pb + q);
        part[tid] = acc;
    }
    named_barrier(1, GT);
    for (int sft = 128; sft > 0; sft >>= 1) {
        if (tid < sft) part[tid] += part[tid + sft];
        named_barrier(1, GT);
    }
    if (tid == 0) *p.loss_sum = part[0];  // ball_sim.py:324
}

}  // namespace doper
