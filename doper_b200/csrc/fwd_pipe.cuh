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
// Here a ball is served by a producer warp and CW consumer warps of a block that never wait for each other inside
// a window of U steps:
//
//   PRODUCER warp   lane = (ball, axis): 16 balls per warp, the x and the y coordinate of a ball integrated by two
//                   lanes (the free-flight step acts on each axis separately: physics.cuh axis_step_acc — the
//                   same operations, hence the same bits).  Nothing but the chain: 24 dependent instructions
//                   per step, candidate states stored to shared memory as they appear.  It runs AHEAD: window
//                   k + 1 is integrated from the last candidate of window k, on the assumption that window k
//                   holds no collision.
//   CONSUMER warps  lane = (ball, step(s) of the window), 16 U candidates spread over 32 CW lanes: check the
//                   candidates of window k against the baked cell lists
//                   (fwd_baked.cuh: same look-up, same narrow phase, same results) while the producer is busy
//                   with window k + 1, finds the ball's first collision, writes the hit log / checkpoints /
//                   collision records of the steps before it, resolves the collision (physics.cuh resolve) and
//                   posts the state after it; the producer restarts that ball two windows later and the
//                   consumer drops the window that was integrated from the wrong state (epoch tag).
//
// One named barrier per pair and window orders the shared-memory traffic (two candidate buffers, two restart
// mailboxes, all indexed by window parity, so that nothing is read in the interval in which it is written).
// A collision costs its ball the rest of its window and the next one (~1.5 U steps); collisions are ~0.15 % of
// the steps of the BASELINE batches.  Every value comes out of the same device functions as in every other
// forward kernel: results are bit-identical (tests/test_parity_gpu.py runs the same cases over this kernel).
#pragma once
#include "fwd_baked.cuh"

namespace doper {

constexpr int kPipeGroups = 2;  // (1 producer + CW consumer warps) groups per block: they share the staged scene
constexpr int kPipeBalls = 16;  // balls per group

template <int U>
struct PipeShared {  // one per group
    float4 slot[2][U + 1][kPipeBalls];  // [window parity][0: window start, j: after the window's step j - 1][ball] = (x, vx, y, vy)
    int t_start[2][kPipeBalls];         // first step of the window
    int epoch_w[2][kPipeBalls];         // restart epoch the window was integrated in
    float4 mail_state[2][kPipeBalls];   // restart state (x, vx, y, vy) posted by the consumer while it read window parity p
    int mail_t[2][kPipeBalls];
    int mail_epoch[2][kPipeBalls];
    int first_hit[4][kPipeBalls];       // per consumer warp: the ball's first collided step among that warp's steps (U: none)
    int done_at;                        // the consumers' last window + 1
    int pad[3];
};

__device__ __forceinline__ void named_barrier(int id, int threads) {
    __syncwarp();  // bar.sync counts warps: every participating warp must arrive converged
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(threads) : "memory");
}

// U: window length (steps integrated between two hand-overs); CW: consumer warps per producer warp.
// A consumer lane checks SPL = U / (2 CW) consecutive steps of one ball.
template <int U, int CW>
__global__ void __launch_bounds__(32 * (1 + CW) * kPipeGroups) rollout_fwd_pipe_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    static_assert(U % (2 * CW) == 0 && CW >= 1 && CW <= 4, "every consumer lane gets a whole number of steps");
    constexpr int SPL = U / (2 * CW);        // steps per consumer lane
    constexpr int GW = 1 + CW;               // warps per group
    constexpr int GT = 32 * GW;              // threads per group
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int group = warp / GW, role = warp % GW;  // role 0: producer; 1..CW: consumer role - 1
    const SceneLayout L = scene_layout(1, p.S);
    const uint32_t scene_bytes = (uint32_t)L.S_pad * 20;
    const uint32_t bbytes = (uint32_t)(bake_bytes(L.S_pad) - sizeof(BakeHeader));
    unsigned char* sbake = smem + scene_bytes;  // 16-byte aligned: S_pad is a multiple of 32
    uint64_t* bar = reinterpret_cast<uint64_t*>(sbake + bbytes);
    PipeShared<U>& sh = reinterpret_cast<PipeShared<U>*>(sbake + bbytes + 16)[group];

    const int64_t first = ((int64_t)blockIdx.x * kPipeGroups + group) * kPipeBalls;
    const int b = lane >> 1;  // ball of this lane within the group (both roles)
    const int64_t ball = first + b;
    const bool active = ball < p.B;
    const int n = p.n_steps;
    const Consts c = p.c;

    if (tid == 0) mbar_init(bar, 1);
    if (role == 1 && lane < kPipeBalls) {
        sh.mail_epoch[0][lane] = 0; sh.mail_epoch[1][lane] = 0;
        if (lane == 0) sh.done_at = -1;
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
    if (first >= p.B) return;  // a group without balls (last block): all its warps leave together
    const int bar_all = 1 + 2 * group, bar_cons = 2 + 2 * group;

    if (role == 0) {
        // ---------------------------------------------------------------------------------- producer
        const int axis = lane & 1;
        const float att = axis ? c.ay : c.ax;
        float pos = 0.0f, vel = 0.0f;
        if (active) {
            const float2 xx = p.x0[ball], vv = p.v0[ball];
            pos = axis ? xx.y : xx.x; vel = axis ? vv.y : vv.x;
        }
        int t = active ? 0 : n, epoch = 0;
        for (int k = 0;; ++k) {
            const int buf = k & 1;
            const int me = sh.mail_epoch[buf][b];
            if (me > epoch) {  // the consumers resolved a collision of this ball two windows ago: restart after it
                const float4 s = sh.mail_state[buf][b];
                pos = axis ? s.z : s.x; vel = axis ? s.w : s.y;
                t = sh.mail_t[buf][b];
                epoch = me;
            }
            if (axis == 0) { sh.t_start[buf][b] = t; sh.epoch_w[buf][b] = epoch; }
            float2* out = reinterpret_cast<float2*>(&sh.slot[buf][0][b]) + axis;
            constexpr int kSlotStride = kPipeBalls * 2;  // float2 per slot
            const float p0 = pos, v0 = vel;
            out[0] = make_float2(pos, vel);
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
            named_barrier(bar_all, GT);  // B_k: window k is complete, the consumers have finished window k - 1
            if (sh.done_at == k) break;
        }
        return;
    }

    // -------------------------------------------------------------------------------------- consumers
    const int cw = role - 1;
    SceneView sc;
    BakeView bk;
    {
        const SceneHeader hdr = *reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off);
        const BakeHeader bh = *reinterpret_cast<const BakeHeader*>(p.scene_ws + L.bake_off);
        sc.q = reinterpret_cast<const float4*>(smem);
        sc.rl = reinterpret_cast<const float*>(smem + (size_t)L.S_pad * sizeof(float4));
        sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
        bk.cell_start = reinterpret_cast<const uint32_t*>(sbake);
        bk.items = reinterpret_cast<const unsigned short*>(sbake + (size_t)(kBakeMaxCells + 16) * sizeof(uint32_t));
        bk.x0 = bh.x0; bk.y0 = bh.y0; bk.h = bh.h; bk.inv_h = bh.inv_h; bk.fnx = bh.fnx; bk.fny = bh.fny; bk.nx = bh.nx;
        bk.far2 = bh.far2; bk.M = bh.M;
        bk.rho2chk = (kCullRhoCheck * bh.rho) * (kCullRhoCheck * bh.rho);
        bk.ok = bh.enabled && !sc.weird && c.r <= bh.r_max;
    }
    mbar_wait(bar, 0);
    const int j0 = (cw * 2 + (lane & 1)) * SPL;  // my first step within the window
    const int K = p.K;
    const int kshift = (K > 0 && (K & (K - 1)) == 0) ? (31 - __clz(K)) : -1;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    int committed = active ? 0 : n;  // steps of this ball that are final (the same value in every lane of the ball)
    int epoch = 0;
    bool finished = false;
    for (int k = 0;; ++k) {
        named_barrier(bar_all, GT);  // B_k
        if (finished) break;
        const int buf = k & 1;
        const int ts0 = sh.t_start[buf][b];
        const bool fresh = committed < n && sh.epoch_w[buf][b] == epoch;  // else: integrated from a superseded state
        const int nv = fresh ? min(U, n - ts0) : 0;                       // steps of the window that exist
        float4 st[SPL + 1];  // st[i]: state at the start of my step i (= after the step before), (x, vx, y, vy)
#pragma unroll
        for (int i = 0; i <= SPL; ++i) st[i] = sh.slot[buf][j0 + i][b];
        // (1) my candidates against the baked lists, in step order; my first collision
        int myf = U, myidx = 0, mine = 0;
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            if (j0 + i < nv && myf == U) {
                int idx;
                if (baked_hit(sc, bk, st[i + 1].x, st[i + 1].z, c.r, idx)) { myf = j0 + i; myidx = idx; mine = i; }
            }
        }
        // (2) the ball's first collision in the window: over the two lanes of the ball in this warp, then over the warps
        const int wf = min(myf, __shfl_xor_sync(0xffffffffu, myf, 1));
        int f = wf;
        if (CW > 1) {
            if ((lane & 1) == 0) sh.first_hit[cw][b] = wf;
            named_barrier(bar_cons, 32 * CW);
#pragma unroll
            for (int w = 0; w < CW; ++w) f = min(f, sh.first_hit[w][b]);
        }
        const bool hit = f < nv;
        const int n_commit = hit ? f + 1 : nv;
        const bool owner = hit && myf == f;  // the lane that checked step f resolves and logs it
        State r;
        r.x = r.y = r.vx = r.vy = r.ax = r.ay = 0.0f;
        int payload = 0;
        if (owner) {
            float4 s = st[0];
#pragma unroll
            for (int i = 1; i < SPL; ++i)
                if (mine == i) s = st[i];
            State s1;
            free_step(s.x, s.z, s.y, s.w, c, s1);  // the candidate again, with its acceleration (resolve needs it)
            r = resolve(s.x, s.z, s.y, s.w, s1, sc.q[myidx], c);
            payload = record_collision(p.sink, ball, ts0 + f, s.x, s.z, s.y, s.w, myidx);
            sh.mail_state[buf][b] = make_float4(r.x, r.vx, r.y, r.vy);
            sh.mail_t[buf][b] = ts0 + f + 1;
            sh.mail_epoch[buf][b] = epoch + 1;
        }
        // (3) commit my steps before the collision, and the collided one
#pragma unroll
        for (int i = 0; i < SPL; ++i) {
            const int j = j0 + i;
            if (j < n_commit) {
                const int ts = ts0 + j;
                const bool h = hit && j == f;
                if (hl) hl[(int64_t)ts * p.hl_st] = pack_log(h ? payload : 0, h, clip_bits(st[i].y, st[i].w));
                if (p.ckpt) {
                    const bool at = kshift >= 0 ? (ts & (K - 1)) == 0 : (ts % K) == 0;
                    if (at) p.ckpt[(size_t)(kshift >= 0 ? ts >> kshift : ts / K) * p.B + ball] = make_float4(st[i].x, st[i].z, st[i].y, st[i].w);
                }
                if (ts + 1 == n) {  // the rollout's last step: final state
                    if (h) { p.xT[ball] = make_float2(r.x, r.y); p.vT[ball] = make_float2(r.vx, r.vy); }
                    else { p.xT[ball] = make_float2(st[i + 1].x, st[i + 1].z); p.vT[ball] = make_float2(st[i + 1].y, st[i + 1].w); }
                }
            }
        }
        if (fresh) committed = ts0 + n_commit;
        if (hit) ++epoch;
        finished = __all_sync(0xffffffffu, committed >= n);
        if (finished && cw == 0 && lane == 0) sh.done_at = k + 1;
    }
}

}  // namespace doper
