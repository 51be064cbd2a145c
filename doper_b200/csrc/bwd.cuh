// bwd.cuh: backward rollout kernels — part of the doper_b200 CUDA translation unit (included by kernels.cu).
//
//   rollout_bwd_kernel           sequential, one thread per ball, fp32 in the oracle's operation order
//                                (bit-exact vs oracle/ball_sim_step.inc:step_vjp); flags bit 1
//   rollout_bwd_jacobian_kernel  default plan, pass 1: one thread per COLLISION RECORD -> the 4x4 transposed
//   rollout_bwd_fold_kernel      Jacobian of that step in binary64; pass 2: one thread per (ball, chunk of
//                                steps) multiplies the chunk's step Jacobians in binary64 (free-flight steps
//                                from the two clip bits of the log word alone, collided steps from pass 1),
//                                then the chunks of a ball are folded onto the output cotangent.
//
// Why binary64 for the default plan.  Re-associating the chain of step Jacobians (which any time-parallel
// backward must do) changes the rounding order, so an fp32 chunked product can be neither bit-identical to
// the reference-order fp32 adjoint nor provably closer to the true derivative than it.  In binary64 the
// product is the adjoint of the real-arithmetic step evaluated on the fp32 trajectory to ~1e-15 — the
// quantity oracle_grad_truth64 (oracle/ball_sim_ref.c) defines as gradient truth — so its distance from the
// reference-order fp32 adjoint IS that adjoint's own rounding error, and it is never further from truth
// than the reference's arithmetic is.  B200 issues 64 DFMA per SM per clock (half the FP32 rate): the whole
// C2 backward is ~50 M DFMA, microseconds of pipe time.
#pragma once
#include "scene.cuh"

namespace doper {

constexpr int kBwdThreads = 128;

struct BwdParams {
    int64_t B;
    int S, per_env, n_steps, K;
    Consts c;
    const unsigned char* scene_ws;
    const float4* ckpt;      // [n_ckpt][B] state at the start of step k K (replay fallback only)
    const HitRecord* rec;    // collision records appended by the forward (nullable)
    const int* rec_count;    // device counter of the forward
    int rec_cap;
    double* jac;             // [rec_cap][16] pass-1 output: row-major 4x4, new_g = Jt g (caller's bwd scratch)
    const int32_t* hitlog;   // element (t, ball) at t * hl_st + ball * hl_sb
    int64_t hl_st, hl_sb;
    const float2* xT_bar;
    const float2* vT_bar;  // nullable
    float2* x0_bar;        // nullable
    float2* v0_bar;
};

__device__ __forceinline__ const float4* scene_q(const BwdParams& p, int64_t ball) {
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    return reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + (p.per_env ? (size_t)ball * L.S_pad : 0);
}

__device__ __forceinline__ const float2* scene_s1(const BwdParams& p, int64_t ball) {
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    return reinterpret_cast<const float2*>(p.scene_ws + L.s1_off) + (p.per_env ? (size_t)ball * L.S_pad : 0);
}

// segment index of a collided step from its log word
__device__ __forceinline__ int hit_segment(const BwdParams& p, int32_t h) {
    const uint32_t w = (uint32_t)h;
    return (w & kLogNoRecord) ? (int)(w & kLogSeg) : p.rec[w & kLogSeg].idx;
}

// State at the start of step t of `ball`, replayed from the last checkpoint at or before t (bit-identical
// to the forward: same device functions, logged segment indices).  Slow path: used only for collided steps
// whose record could not be stored.
__device__ __noinline__ float4 replay_state(const BwdParams& p, int64_t ball, int t, const float4* q) {
    const int ck = t / p.K;
    float4 s = p.ckpt[(size_t)ck * p.B + ball];
    const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;
    for (int u = ck * p.K; u < t; ++u) {
        const int32_t hu = hl[(size_t)u * p.hl_st];
        State s1;
        free_step(s.x, s.y, s.z, s.w, p.c, s1);
        if (hu < 0) s1 = resolve(s.x, s.y, s.z, s.w, s1, __ldg(q + hit_segment(p, hu)), p.c);
        s = make_float4(s1.x, s1.y, s1.vx, s1.vy);
    }
    return s;
}

// ------------------------------------------------------------------------------------------
// sequential backward (oracle order, bit-exact)
// ------------------------------------------------------------------------------------------
constexpr int kBwdPrefetch = 16;  // log words fetched ahead per thread (independent loads in flight)

__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_kernel(const BwdParams p) {
    const int tid = threadIdx.x;
    const int64_t ball = (int64_t)blockIdx.x * kBwdThreads + tid;
    if (ball >= p.B) return;
    const float4* q = scene_q(p, ball);
    const Consts c = p.c;
    const float dk = dfric_value(c);
    float2 g = p.xT_bar[ball];
    float gx = g.x, gy = g.y, gvx = 0.0f, gvy = 0.0f;
    if (p.vT_bar) { float2 h = p.vT_bar[ball]; gvx = h.x; gvy = h.y; }
    const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;

    for (int tb = p.n_steps; tb > 0;) {
        // the log words of steps [ta, tb): kBwdPrefetch independent loads, then a latency-free loop
        const int ta = max(0, tb - kBwdPrefetch);
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
            const uint32_t h = (uint32_t)hl[(size_t)t * p.hl_st];  // collided step: the payload is needed (rare reload)
            float4 st;
            int idx;
            if (h & kLogNoRecord) {
                idx = (int)(h & kLogSeg);
                st = replay_state(p, ball, t, q);
            } else {
                const float4* r = reinterpret_cast<const float4*>(p.rec + (h & kLogSeg));
                st = r[0];
                idx = reinterpret_cast<const int4*>(r)[1].z;
            }
            step_vjp(st.x, st.y, st.z, st.w, true, __ldg(q + idx), c, gx, gy, gvx, gvy);
        }
        tb = ta;
    }
    p.v0_bar[ball] = make_float2(gvx, gvy);
    if (p.x0_bar) p.x0_bar[ball] = make_float2(gx, gy);
}

// ------------------------------------------------------------------------------------------
// binary64 adjoint of a collided step
// ------------------------------------------------------------------------------------------
struct ConstsD {
    double r, m, f, e, ks, g, ax, ay, dt;
    double dk;    // d fric / d v inside the clip: -(((m g) f) / r)
    double a_p;   // free flight, per axis: gp' = gp + a_p tv,  a_p = -2 ks dt / m
    double a_v;   // gv' = tv + (clip active ? a_v : 0) tv,     a_v = dk dt / m
};

__device__ __forceinline__ ConstsD make_consts_d(const Consts& c) {
    ConstsD d;
    d.r = c.r; d.m = c.m; d.f = c.f; d.e = c.e; d.ks = c.ks; d.g = (double)9.8f;  // the reference's literal, as f32
    d.ax = c.ax; d.ay = c.ay; d.dt = c.dt;
    d.dk = -(((d.m * d.g) * d.f) / d.r);
    d.a_p = -2.0 * d.ks * (d.dt / d.m);
    d.a_v = d.dk * (d.dt / d.m);
    return d;
}

__device__ __forceinline__ double clip_d(double x, double lo, double hi) {
    if (x != x) return x;
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

__device__ __forceinline__ double fric_d(double v, const ConstsD& c) {
    return ((((-clip_d(v, -1.0, 1.0)) * c.m) * c.g) * c.f) / c.r;
}

__device__ __forceinline__ double sgn_d(double a) { return a > 0.0 ? 1.0 : (a < 0.0 ? -1.0 : 0.0); }

// Jt (row-major 4x4 over (x, y, vx, vy)): cotangent of the step OUTPUT -> cotangent of the step INPUT, for a
// step that collided with the segment s0 = (q.x, q.y) -> s1, starting from the fp32 state st.  The formulas of
// oracle/ball_sim_step.inc:step_vjp in binary64, the segment direction included
// (s1 - s0 in binary64, like oracle_grad_truth64); column k = the pull-back of e_k.
// Columns [col0, col1) only: the four columns are independent given the step's forward internals, so four
// threads can share one collided step (each recomputes the internals and pulls ONE basis cotangent back).
__device__ __noinline__ void hit_jacobian_d(const float4 st, const float4 q, const float2 s1, const ConstsD c,
                                            double* __restrict__ Jt, const int col0 = 0, const int col1 = 4) {
    const double x = st.x, y = st.y, vx = st.z, vy = st.w;
    const double s0x = q.x, s0y = q.y, svx = (double)s1.x - (double)q.x, svy = (double)s1.y - (double)q.y;
    const double L = svx * svx + svy * svy;
    // free-flight candidate (ball_sim.py:192-202)
    const double ppx = -(2.0 * (x - c.ax)), ppy = -(2.0 * (y - c.ay));
    const double a1x = (c.ks * ppx + fric_d(vx, c)) / c.m, a1y = (c.ks * ppy + fric_d(vy, c)) / c.m;
    const double v1x = vx + a1x * c.dt, v1y = vy + a1y * c.dt;
    const double x1x = x + v1x * c.dt, x1y = y + v1y * c.dt;
    // resolve_collision internals (ball_sim.py:97-132)
    const double t1 = clip_d(((x1x - s0x) * svx + (x1y - s0y) * svy) / L, 0.0, 1.0);
    const double ox = (s0x + t1 * svx) - x, oy = (s0y + t1 * svy) - y;
    const double od = sqrt(ox * ox + oy * oy);
    const double nx = ox / od, ny = oy / od;
    const double s = nx * v1x + ny * v1y;
    const double den = fabs(s);
    const double toi_raw = fabs(od - c.r) / den;
    const double toi = clip_d(toi_raw, 0.0, c.dt);
    const double vsx = vx + a1x * toi, vsy = vy + a1y * toi;
    const double xsx = x + vsx * toi, xsy = y + vsy * toi;
    const double t2 = clip_d(((xsx - s0x) * svx + (xsy - s0y) * svy) / L, 0.0, 1.0);
    const double o2x = (s0x + t2 * svx) - xsx, o2y = (s0y + t2 * svy) - xsy;
    const double o2d = sqrt(o2x * o2x + o2y * o2y);
    const double n2x = o2x / o2d, n2y = o2y / o2d;
    const double k = n2x * vsx + n2y * vsy;
    const double vnx = n2x * k, vny = n2y * k;
    const double vax = (vsx - vnx) - c.e * vnx, vay = (vsy - vny) - c.e * vny;
    const double pbx = -(2.0 * (xsx - c.ax)), pby = -(2.0 * (xsy - c.ay));
    const double abx = (pbx + fric_d(vax, c)) / c.m, aby = (pby + fric_d(vay, c)) / c.m;
    const double tau = c.dt - toi;
    const double vbx = vax + abx * tau, vby = vay + aby * tau;
    const double dfax = (vax > -1.0 && vax < 1.0) ? c.dk : 0.0, dfay = (vay > -1.0 && vay < 1.0) ? c.dk : 0.0;
    const double dfx = (vx > -1.0 && vx < 1.0) ? c.dk : 0.0, dfy = (vy > -1.0 && vy < 1.0) ? c.dk : 0.0;
    const bool in1 = t1 > 0.0 && t1 < 1.0, in2 = t2 > 0.0 && t2 < 1.0, raw_in = toi_raw > 0.0 && toi_raw < c.dt;
    const double sg_od = sgn_d(od - c.r), sg_s = sgn_d(s);
    // the pull-back divides by m, o2d, L, den and od a dozen times per column: five reciprocals, taken once (a binary64
    // division is a ~100-cycle dependent sequence and the columns are a serial chain; the products differ from the
    // quotients by an ulp of binary64, nine orders of magnitude below the fp32 result's rounding)
    const double rm = 1.0 / c.m, ro2d = 1.0 / o2d, rL = 1.0 / L, rden = 1.0 / den, rod = 1.0 / od;
#pragma unroll 1
    for (int col = col0; col < col1; ++col) {
        const double gx = col == 0 ? 1.0 : 0.0, gy = col == 1 ? 1.0 : 0.0;
        const double gvx = col == 2 ? 1.0 : 0.0, gvy = col == 3 ? 1.0 : 0.0;
        double bx = 0.0, by = 0.0, bvx = 0.0, bvy = 0.0, bax = 0.0, bay = 0.0;
        const double gvbx = gvx + gx * tau, gvby = gvy + gy * tau;
        double gxsx = gx, gxsy = gy;
        const double gtau = (gx * vbx + gy * vby) + (gvbx * abx + gvby * aby);
        double gvax = gvbx, gvay = gvby;
        const double gabx = gvbx * tau, gaby = gvby * tau;
        gxsx += -2.0 * (gabx * rm); gxsy += -2.0 * (gaby * rm);
        gvax += (gabx * rm) * dfax; gvay += (gaby * rm) * dfay;
        double gtoi = -gtau;
        double gvsx = gvax, gvsy = gvay;
        const double gvnx = -(1.0 + c.e) * gvax, gvny = -(1.0 + c.e) * gvay;
        const double gk = gvnx * n2x + gvny * n2y;
        const double gn2x = gvnx * k + gk * vsx, gn2y = gvny * k + gk * vsy;
        gvsx += gk * n2x; gvsy += gk * n2y;
        const double nd = n2x * gn2x + n2y * gn2y;
        const double go2x = (gn2x - n2x * nd) * ro2d, go2y = (gn2y - n2y * nd) * ro2d;
        const double j2 = in2 ? (go2x * svx + go2y * svy) * rL : 0.0;
        gxsx += j2 * svx - go2x; gxsy += j2 * svy - go2y;
        bx += gxsx; by += gxsy;
        gvsx += gxsx * toi; gvsy += gxsy * toi;
        gtoi += gxsx * vsx + gxsy * vsy;
        bvx += gvsx; bvy += gvsy;
        bax += gvsx * toi; bay += gvsy * toi;
        gtoi += gvsx * a1x + gvsy * a1y;
        const double graw = raw_in ? gtoi : 0.0;
        const double gnum = graw * rden;
        const double gden = -(graw * toi_raw) * rden;
        const double god = gnum * sg_od;
        const double gs = gden * sg_s;
        const double gnx = gs * v1x, gny = gs * v1y;
        const double b1vx = gs * nx, b1vy = gs * ny;
        const double nd1 = nx * gnx + ny * gny;
        const double gox = (gnx - nx * nd1) * rod + god * nx, goy = (gny - ny * nd1) * rod + god * ny;
        bx -= gox; by -= goy;
        const double j1 = in1 ? (gox * svx + goy * svy) * rL : 0.0;
        const double b1x = j1 * svx, b1y = j1 * svy;
        // common tail: x1 = x + v1 dt ; v1 = v + a dt ; a = (ks pot + fric(v)) / m
        const double tvx = b1vx + b1x * c.dt, tvy = b1vy + b1y * c.dt;
        bx += b1x; by += b1y;
        bvx += tvx; bvy += tvy;
        bax += tvx * c.dt; bay += tvy * c.dt;
        const double amx = bax * rm, amy = bay * rm;
        bx += -2.0 * (c.ks * amx); by += -2.0 * (c.ks * amy);
        bvx += amx * dfx; bvy += amy * dfy;
        Jt[0 * 4 + col] = bx; Jt[1 * 4 + col] = by; Jt[2 * 4 + col] = bvx; Jt[3 * 4 + col] = bvy;
    }
}

// pass 1: four threads per collision record, one column of the transposed Jacobian each; thread `first` of `total`
__device__ __forceinline__ void bwd_jacobian_part(const BwdParams& p, const int64_t first, const int64_t total) {
    const int n = min(*p.rec_count, p.rec_cap);
    const ConstsD cd = make_consts_d(p.c);
    for (int64_t w = first; w < 4 * (int64_t)n; w += total) {
        const int i = (int)(w >> 2), col = (int)(w & 3);
        const float4* r = reinterpret_cast<const float4*>(p.rec + i);
        const float4 st = r[0];
        const int4 m = reinterpret_cast<const int4*>(r)[1];
        const int64_t ball = ((int64_t)m.w << 32) | (uint32_t)m.x;
        // a slot no collision was written to (the unused tail of a tail worker's slab: stale bytes) — nothing refers
        // to its matrix, but its indices must not reach the scene
        if ((unsigned)m.z >= (unsigned)p.S || ball < 0 || ball >= p.B) continue;
        double J[16];
        hit_jacobian_d(st, __ldg(scene_q(p, ball) + m.z), __ldg(scene_s1(p, ball) + m.z), cd, J, col, col + 1);
        double* out = p.jac + (size_t)i * 16 + col;
#pragma unroll
        for (int k = 0; k < 4; ++k) out[4 * k] = J[4 * k + col];
    }
}

__global__ void __launch_bounds__(128) rollout_bwd_jacobian_kernel(const BwdParams p) {
    bwd_jacobian_part(p, (int64_t)blockIdx.x * blockDim.x + threadIdx.x, (int64_t)gridDim.x * blockDim.x);
}

// ------------------------------------------------------------------------------------------
// pass 2: chunk products + fold
// ------------------------------------------------------------------------------------------
// Thread (lane = ball, warp = chunk): pulls the four basis cotangents e_0..e_3 back through its chunk of
// `chunk_len` steps, g[k] <- J_t^T g[k] for t descending, which leaves column k of the chunk's transposed
// Jacobian P_c in g[k].  Until the chunk meets a collision, x and y are decoupled (free flight acts on the
// (position, velocity) pair of each axis separately), so only the 8 entries that can be non-zero are updated:
// 12 DFMA per step; 24 once a collision has coupled the axes; a collided step is one 4x4 x 4x4 product with
// the matrix of pass 1.  The block then folds v <- P_c v from the last chunk to the first through shared
// memory (warp 0, one lane per ball).  With lane = ball the loads of a step-major hit log ([n_steps][B], the
// lane-split forward kernels) are coalesced; a ball-major log ([B][n_steps], lane = step forward kernels) is
// read in per-lane runs that the L1 keeps.
// g[k] <- (product of the transposed step Jacobians of steps t1-1 .. t0) e_k for the four basis cotangents: column k
// of the chunk's transposed Jacobian.  Free-flight steps from the two clip bits of the log word alone; collided steps
// from the Jacobian pass's matrix (p.jac), or computed in place when no pass ran / no record exists.
//
// Runs of free flight.  The transposed Jacobian of a free-flight step acts on the (position, velocity) cotangent pair
// of each axis as one of two constant 2x2 matrices (clip bit of that axis clear / set) — the two axes never mix in
// free flight, whether or not an earlier collision has coupled the columns — and a ball's clip bits change a few
// times per rollout.  Between two collided steps a prefetch group is therefore, per axis, a few RUNS of equal clip
// bits, and a run of L steps is the L-th power of one matrix: the powers 1, 2, 4, 8, 16 of both matrices sit in
// shared memory (free_pow_table) and a run applies the ones in L's binary expansion (the powers of one matrix
// commute) — one product per axis for the typical group of 16 steps instead of 16 x 12 DFMA, a handful when a clip
// bit flips inside the group.  Only collided steps are multiplied one by one.  (Binary64 throughout; a power differs
// from the step-by-step product by a few 1e-16 relative.)
constexpr int kPowLevels = 5;  // powers 2^0 .. 2^4 (kBwdPrefetch <= 32 steps per group: run lengths below 32)
constexpr int kPowTableDoubles = 2 * kPowLevels * 4;
__device__ __forceinline__ void free_pow_table(const ConstsD& cd, double* tp) {  // [clip bit][level][R00 R01 R10 R11]
    for (int b = 0; b < 2; ++b) {
        const double av = b ? cd.a_v : 0.0;
        // one step on a pair (a, b): tx = a dt + b; a' = a + a_p tx; b' = tx + av tx
        double r00 = fma(cd.a_p, cd.dt, 1.0), r01 = cd.a_p, r10 = (1.0 + av) * cd.dt, r11 = 1.0 + av;
        for (int lv = 0; lv < kPowLevels; ++lv) {
            double* o = tp + (b * kPowLevels + lv) * 4;
            o[0] = r00; o[1] = r01; o[2] = r10; o[3] = r11;
            const double n00 = fma(r00, r00, r01 * r10), n01 = fma(r00, r01, r01 * r11);
            const double n10 = fma(r10, r00, r11 * r10), n11 = fma(r10, r01, r11 * r11);
            r00 = n00; r01 = n01; r10 = n10; r11 = n11;
        }
    }
}

__device__ __forceinline__ void pair_pow(double& a, double& b, const double r00, const double r01, const double r10, const double r11) {
    const double na = fma(r00, a, r01 * b), nb = fma(r10, a, r11 * b);
    a = na; b = nb;
}

// Steps lo .. hi (bit positions of `mask`, no collided step among them) of ONE axis, newest first: pos / vel index the
// cotangent pair of that axis in a column (0 / 2 for x, 1 / 3 for y); uncoupled columns off the axis stay zero and
// are skipped.
template <int POS, int VEL>
__device__ __forceinline__ void axis_runs(double (&g)[4][4], const bool coupled, const unsigned mask, const int lo, const int hi,
                                          const double* __restrict__ tp) {
    for (int q = hi; q >= lo;) {
        const unsigned bit = (mask >> q) & 1u;
        const unsigned same = (bit ? mask : ~mask) << (31 - q);  // bit 31 = step q; ones = steps with the same clip bit
        int run = __clz(~same);                                   // leading ones (32 if all)
        run = min(run, q - lo + 1);
        const double* tb = tp + bit * (kPowLevels * 4);
#pragma unroll
        for (int lv = 0; lv < kPowLevels; ++lv) {
            if ((run >> lv) & 1) {
                const double r00 = tb[lv * 4 + 0], r01 = tb[lv * 4 + 1], r10 = tb[lv * 4 + 2], r11 = tb[lv * 4 + 3];
                if (coupled) {
#pragma unroll
                    for (int k = 0; k < 4; ++k) pair_pow(g[k][POS], g[k][VEL], r00, r01, r10, r11);
                } else {  // columns POS and VEL live on this axis
                    pair_pow(g[POS][POS], g[POS][VEL], r00, r01, r10, r11);
                    pair_pow(g[VEL][POS], g[VEL][VEL], r00, r01, r10, r11);
                }
            }
        }
        q -= run;
    }
}

__device__ __forceinline__ void chunk_product(const BwdParams& p, const ConstsD& cd, const int64_t ball, const int t0, const int t1,
                                              double (&g)[4][4], const double* __restrict__ tp) {
    static_assert(kBwdPrefetch < (1 << kPowLevels), "run lengths must fit the power table");
    const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;
    bool coupled = false;
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
        for (int j = tb - ta - 1; j >= 0;) {
            if (!((m_hit >> j) & 1u)) {  // free flight down to the next collided step (or the group's first step)
                const unsigned below = m_hit & ((1u << j) - 1u);
                const int lo = below ? 32 - __clz(below) : 0;
                axis_runs<0, 2>(g, coupled, m_cx, lo, j, tp);
                axis_runs<1, 3>(g, coupled, m_cy, lo, j, tp);
                j = lo - 1;
                continue;
            }
            const int t = ta + j;
            const uint32_t h = (uint32_t)hl[(size_t)t * p.hl_st];  // collided step: the payload is needed (rare reload)
            double Jl[16];
            const double* J;
            if (h & kLogNoRecord) {  // no record: replay the state from the checkpoint, Jacobian in place (slow path)
                const float4* q = scene_q(p, ball);
                const float4 st = replay_state(p, ball, t, q);
                hit_jacobian_d(st, __ldg(q + (h & kLogSeg)), __ldg(scene_s1(p, ball) + (h & kLogSeg)), cd, Jl);
                J = Jl;
            } else if (p.jac != nullptr) {
                J = p.jac + (size_t)(h & kLogSeg) * 16;
            } else {  // no Jacobian pass ran (the fused forward + backward launch): from the record, in place
                const float4* r = reinterpret_cast<const float4*>(p.rec + (h & kLogSeg));
                const int seg = reinterpret_cast<const int4*>(r)[1].z;
                hit_jacobian_d(r[0], __ldg(scene_q(p, ball) + seg), __ldg(scene_s1(p, ball) + seg), cd, Jl);
                J = Jl;
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const double a0 = g[k][0], a1 = g[k][1], a2 = g[k][2], a3 = g[k][3];
#pragma unroll
                for (int r = 0; r < 4; ++r)
                    g[k][r] = fma(J[4 * r + 0], a0, fma(J[4 * r + 1], a1, fma(J[4 * r + 2], a2, J[4 * r + 3] * a3)));
            }
            coupled = true;
            --j;
        }
        tb = ta;
    }
}

constexpr int kFoldMaxChunks = 16;  // 512 threads per block: 128 registers per thread (the 4x4 binary64 product needs ~90)

__global__ void __launch_bounds__(32 * kFoldMaxChunks) rollout_bwd_fold_kernel(const BwdParams p, const int n_chunks,
                                                                              const int chunk_len) {
    extern __shared__ __align__(16) unsigned char fold_smem[];
    double* sm = reinterpret_cast<double*>(fold_smem);  // [n_chunks][16][32]
    const int lane = threadIdx.x & 31, chunk = threadIdx.x >> 5;
    const int64_t ball = (int64_t)blockIdx.x * 32 + lane;
    const bool active = ball < p.B;
    const ConstsD cd = make_consts_d(p.c);
    double g[4][4];
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r) g[k][r] = (k == r) ? 1.0 : 0.0;
    const int t0 = chunk * chunk_len, t1 = min(p.n_steps, t0 + chunk_len);
    __shared__ double t16[kPowTableDoubles];
    if (threadIdx.x == 0) free_pow_table(cd, t16);
    __syncthreads();
    if (active && t0 < t1) chunk_product(p, cd, ball, t0, t1, g, t16);
    // fold: v <- P_c v for c = n_chunks - 1 .. 0, P_c column k = g[k] of thread (ball, c)
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r) sm[((size_t)chunk * 16 + (k * 4 + r)) * 32 + lane] = g[k][r];
    __syncthreads();
    if (chunk == 0 && active) {
        double v[4] = {0.0, 0.0, 0.0, 0.0};
        const float2 gg = p.xT_bar[ball];
        v[0] = gg.x; v[1] = gg.y;
        if (p.vT_bar) { const float2 h = p.vT_bar[ball]; v[2] = h.x; v[3] = h.y; }
        for (int cc = n_chunks - 1; cc >= 0; --cc) {
            const double* P = sm + (size_t)cc * 16 * 32 + lane;
            double o[4];
#pragma unroll
            for (int r = 0; r < 4; ++r)
                o[r] = fma(v[0], P[(0 * 4 + r) * 32], fma(v[1], P[(1 * 4 + r) * 32], fma(v[2], P[(2 * 4 + r) * 32], v[3] * P[(3 * 4 + r) * 32])));
#pragma unroll
            for (int r = 0; r < 4; ++r) v[r] = o[r];
        }
        p.v0_bar[ball] = make_float2((float)v[2], (float)v[3]);
        if (p.x0_bar) p.x0_bar[ball] = make_float2((float)v[0], (float)v[1]);
    }
}

}  // namespace doper
