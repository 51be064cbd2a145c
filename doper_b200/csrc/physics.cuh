// O(1) per-step physics of the rollout and its adjoint, in the reference's exact fp32
// operation order.
//
// This translation unit is compiled with -fmad=false, IEEE division / sqrt (-prec-div=true
// -prec-sqrt=true, no -use_fast_math, denormals kept), so every `a*b + c` below is two
// roundings exactly as written.  WHICH mul+add pairs of the reference's forward arithmetic are
// single-rounding fused multiply-adds is decided in one header shared with the CPU oracle,
// include/doper_fp_policy.h (DOPER_POS_UPDATE / VEL_UPDATE / LERP / DOT2 / FORCE_SUM / REFLECT;
// default: only get_new_state's coordinate update, ball_sim.py:76, which the reference's recorded
// output pins as contracted).  The other explicit __fmaf_rn calls are the exact
// divisions by loop invariants (Markstein) and the sweep's candidate filter (sweep.cuh), which never
// decides a result.
//
// Reference: doper/sim/ball_sim.py:15-132,171-209 ; doper/sim/jax_geometry.py:119-138.
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/doper_fp_policy.h"

namespace doper {

struct Consts {
    float r, m, f, e, ks;  // radius, mass, friction coeff, wall elasticity, attractor strength
    float ax, ay;          // attractor
    float dt;
    float rcp_r, rcp_m;    // RN(1/r), RN(1/m) for div_invariant()
    float div_lo, div_hi;  // |numerator| range in which div_invariant's fast path is exact
};

struct State {
    float x, y, vx, vy, ax, ay;
};

// jnp.clip = min(max(x, lo), hi) with NaN propagation (CUDA fminf/fmaxf would drop the NaN)
__device__ __forceinline__ float clip_nan(float x, float lo, float hi) {
    float a, b;  // the NaN-propagating min / max of sm_80+ (FMNMX.NAN): two instructions instead of four
    asm("max.NaN.f32 %0, %1, %2;" : "=f"(a) : "f"(x), "f"(lo));
    asm("min.NaN.f32 %0, %1, %2;" : "=f"(b) : "f"(a), "f"(hi));
    return b;
}

// Correctly rounded a / b for a loop-invariant divisor, without the IEEE-division sequence:
// with y = RN(1/b), q0 = RN(a y), two Markstein corrections q <- RN(q + RN(a - b q) y), where the
// residual a - b q is exact in an FMA.  After the first correction q is within 1 ulp; the second
// is then correctly rounded (Markstein's theorem needs y = RN(1/b), which the host guarantees).
// Outside [div_lo, div_hi] (tiny / huge / NaN numerators, where the residual could leave the
// normal range) the true division runs.  Checked against a / b on 2e9 random pairs on the CPU
// (0 mismatches) and by the bit-exact parity tests.
__device__ __forceinline__ float div_invariant(float a, float b, float y, const Consts& c) {
    if (fabsf(a) >= c.div_lo && fabsf(a) <= c.div_hi) {
        float q = __fmul_rn(a, y);
        q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
        q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
        return q;
    }
    return a / b;
}

// ball_sim.py:43-45 : ((((-clip(v,-1,1)) * m) * 9.8) * f) / r
__device__ __forceinline__ float fric(float v, const Consts& c) {
    return div_invariant((((-clip_nan(v, -1.0f, 1.0f)) * c.m) * 9.8f) * c.f, c.r, c.rcp_r, c);
}

// get_new_state coordinate update, ball_sim.py:76 (single rounding, see header comment)
__device__ __forceinline__ float pos_update(float v, float dt, float x) { return DOPER_POS_UPDATE(v, dt, x); }
// get_new_state velocity update v + a dt, ball_sim.py:75 (:109, :131)
__device__ __forceinline__ float vel_update(float a, float dt, float v) { return DOPER_VEL_UPDATE(a, dt, v); }
// 2-vector dot product / squared norm a0 b0 + a1 b1
__device__ __forceinline__ float dot2(float a0, float b0, float a1, float b1) { return DOPER_DOT2(a0, b0, a1, b1); }

// jax_geometry.py:132-136 on a pre-digested segment q = (s0x, s0y, svx, svy).
// sv = s1 - s0 is ball-independent, so reading it is bit-identical to recomputing it.
__device__ __forceinline__ void proj(float px, float py, const float4& q, float& cx, float& cy,
                                     float& t) {
    float L = dot2(q.z, q.z, q.w, q.w);
    float pvx = px - q.x, pvy = py - q.y;
    t = clip_nan(dot2(pvx, q.z, pvy, q.w) / L, 0.0f, 1.0f);
    cx = DOPER_LERP(q.x, t, q.z);
    cy = DOPER_LERP(q.y, t, q.w);
}

// squared distance exactly as the reference computes it before the sqrt (jax_geometry.py:151-152)
__device__ __forceinline__ float seg_dist2(float px, float py, const float4& q) {
    float cx, cy, t;
    proj(px, py, q, cx, cy, t);
    float dx = px - cx, dy = py - cy;
    return dot2(dx, dx, dy, dy);
}

// Hit-log word (one per ball-step, written by every forward kernel, read by the backward kernels):
//   bit 31     the step collided
//   bit 30/29  |vy| < 1 / |vx| < 1 for the velocity at the START of the step: the only thing the
//              adjoint of a free-flight step depends on (the friction clip's derivative,
//              ball_sim.py:43), so the backward needs no state for the steps that did not collide
//   bits 0-28  closest-segment index (consumed at collisions only)
//   bits 0-28  payload.  Step without a collision: the closest-segment index (0 unless the exact index log,
//              flags bit 17, was asked for).  Collided step: the slot of the step's COLLISION RECORD (below),
//              or, with bit 28 set (kLogNoRecord), the segment index itself: no record could be appended
//              (record area full or disabled) and the backward replays the step from its checkpoint.
constexpr uint32_t kLogHit = 0x80000000u, kLogClipY = 0x40000000u, kLogClipX = 0x20000000u, kLogIdx = 0x1fffffffu;
constexpr uint32_t kLogNoRecord = 0x10000000u, kLogSeg = 0x0fffffffu;

// Collision record: everything the backward needs about one collided step (the state at its START and the
// segment), appended by the forward kernels in arrival order (one atomic per collision; the slot goes
// into the step's log word, so the order does not matter).  32 bytes.
struct HitRecord {
    float x, y, vx, vy;
    int32_t ball_lo, t, idx, ball_hi;
};

struct HitSink {
    HitRecord* rec;  // nullable
    int* count;      // records appended so far (may exceed cap: the excess was not stored)
    int cap;
};

// Split in two so that the atomic's round trip to L2 (~0.5 - 1 us, and the branch on its result) does not sit on
// the serial chain of a collided step: record_reserve() BEFORE the collision is resolved (the slot is not consumed
// yet), record_commit() after it.  A ball in sliding contact collides on every step; its chain sets the duration of
// the whole forward kernel (profiles/r2_ncu_fwd_baked1_c5.txt: the SMs are active 49 % of the kernel on average).
__device__ __forceinline__ int record_reserve(const HitSink& s) {
    // atomicInc, not atomicAdd(count, 1): for a uniform increment the compiler (ptxas, even from inline PTX)
    // aggregates the atomic over the warp — leader election, one ATOMG, SHFL of its result to the other lanes — and
    // that shuffle consumes the result on the spot: the warp then waits out the whole round trip here.  Colliding
    // lanes per warp step are 0 or 1 anyway.  (Bound 2^31 - 1, never reached: with the all-ones bound ptxas turns the
    // INC back into an aggregated ADD.)
    int slot = 0x7fffffff;
    if (s.rec != nullptr && s.cap > 0) slot = (int)atomicInc(reinterpret_cast<unsigned int*>(s.count), 0x7fffffffu);
    return slot;
}

// A thread that runs ONE ball in sliding contact (tail workers) reserves record slots a SLAB at a time: one atomic per
// kRecordSlab collisions instead of one per collided step, whose round trip would sit on the ball's serial chain.  What
// is left of the last slab when the ball ends stays unused: the record area may hold slots no log word refers to
// (stale bytes; the Jacobian pass validates a record's indices before it touches the scene).
constexpr int kRecordSlab = 128;
struct RecordSlab { int next = 0, end = 0; };
__device__ __forceinline__ int record_reserve_slab(const HitSink& s, RecordSlab& sl) {
    if (s.rec == nullptr || s.cap <= 0) return 0x7fffffff;
    if (sl.next == sl.end) { sl.next = atomicAdd(s.count, kRecordSlab); sl.end = sl.next + kRecordSlab; }
    return sl.next++;
}

__device__ __forceinline__ int record_commit(const HitSink& s, int slot, int64_t ball, int t, float x, float y, float vx,
                                             float vy, int idx) {
    if (slot < s.cap) {  // (no record area: slot = INT_MAX)
        float4* dst = reinterpret_cast<float4*>(s.rec + slot);
        dst[0] = make_float4(x, y, vx, vy);
        reinterpret_cast<int4*>(dst)[1] = make_int4((int)(ball & 0xffffffffll), t, idx, (int)(ball >> 32));
        return slot;
    }
    return (int)((uint32_t)idx | kLogNoRecord);
}

__device__ __forceinline__ int record_collision(const HitSink& s, int64_t ball, int t, float x, float y, float vx,
                                                float vy, int idx) {
    return record_commit(s, record_reserve(s), ball, t, x, y, vx, vy, idx);
}

// Hand-over queue of rollout_fwd_baked_kernel<1> (fwd_baked.cuh): a ball whose lane collides step after step is
// taken out of its warp (whose 31 other lanes it would stall for the rest of the rollout, and whose divergent work
// lengthens its own serial chain) and finished by a tail-worker warp that runs this one ball alone.
struct DeferEntry {   // 32 B
    float x, y, vx, vy;            // state at the START of step t
    int32_t ball_lo, t, pad, ball_hi;
};
struct DeferQueue {
    DeferEntry* entry;     // [cap]; null = no hand-over
    unsigned int* flag;    // [cap] zeroed per launch; 1 = the entry is complete
    unsigned int* alloc;   // entries allocated (zeroed counter in the workspace header)
    unsigned int* head;    // entries claimed
    int cap;
    int tail_warps;        // warps per block that only serve the queue (the block's last warps)
    int streak;            // consecutive collided steps after which a ball is handed over
};

__device__ __forceinline__ uint32_t clip_bits(float vx, float vy) {
    return ((vx > -1.0f && vx < 1.0f) ? kLogClipX : 0u) | ((vy > -1.0f && vy < 1.0f) ? kLogClipY : 0u);
}

__device__ __forceinline__ int32_t pack_log(int idx, bool hit, uint32_t clipb) {
    return (int32_t)(((uint32_t)idx & kLogIdx) | (hit ? kLogHit : 0u) | clipb);
}

// ball_sim.py:192-202 : forces + semi-implicit Euler -> candidate state (x1, v1, a)
__device__ __forceinline__ void free_step(float x, float y, float vx, float vy, const Consts& c,
                                          State& o) {
    float px = -(2.0f * (x - c.ax)), py = -(2.0f * (y - c.ay));
    float fx = fric(vx, c), fy = fric(vy, c);
    float ax = div_invariant(DOPER_FORCE_SUM(c.ks, px, fx), c.m, c.rcp_m, c), ay = div_invariant(DOPER_FORCE_SUM(c.ks, py, fy), c.m, c.rcp_m, c);
    float v1x = vel_update(ax, c.dt, vx), v1y = vel_update(ay, c.dt, vy);
    o.x = pos_update(v1x, c.dt, x);
    o.y = pos_update(v1y, c.dt, y);
    o.vx = v1x; o.vy = v1y; o.ax = ax; o.ay = ay;
}

// Straight-line variant for software-pipelined loops: the Markstein path of div_invariant for all
// four divisions with no branch in between; returns false when any numerator was outside the range
// in which that path is exact (then the caller must redo the step with free_step()).  Same values
// as free_step() whenever it returns true.
__device__ __forceinline__ float div_markstein(float a, float b, float y) {
    float q = __fmul_rn(a, y);
    q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
    q = __fmaf_rn(__fmaf_rn(-b, q, a), y, q);
    return q;
}

// The range test of the four numerators of one step, folded into a running minimum / maximum of their
// magnitudes (3-input FMNMX3: four instructions per step, no predicate logic); a loop tests the pair
// ONCE after its last step.  fminf / fmaxf skip NaN numerators, which is fine: the Markstein path
// returns the same (canonical) NaN as the IEEE division.
__device__ __forceinline__ void free_step_acc(float x, float y, float vx, float vy, const Consts& c, State& o,
                                              float& nmin, float& nmax) {
    const float px = -(2.0f * (x - c.ax)), py = -(2.0f * (y - c.ay));
    const float nfx = (((-clip_nan(vx, -1.0f, 1.0f)) * c.m) * 9.8f) * c.f;
    const float nfy = (((-clip_nan(vy, -1.0f, 1.0f)) * c.m) * 9.8f) * c.f;
    const float fx = div_markstein(nfx, c.r, c.rcp_r), fy = div_markstein(nfy, c.r, c.rcp_r);
    const float nax = DOPER_FORCE_SUM(c.ks, px, fx), nay = DOPER_FORCE_SUM(c.ks, py, fy);
    const float ax = div_markstein(nax, c.m, c.rcp_m), ay = div_markstein(nay, c.m, c.rcp_m);
    const float v1x = vel_update(ax, c.dt, vx), v1y = vel_update(ay, c.dt, vy);
    o.x = pos_update(v1x, c.dt, x);
    o.y = pos_update(v1y, c.dt, y);
    o.vx = v1x; o.vy = v1y; o.ax = ax; o.ay = ay;
    nmax = fmaxf(fmaxf(nmax, fabsf(nfx)), fabsf(nfy));
    nmax = fmaxf(fmaxf(nmax, fabsf(nax)), fabsf(nay));
    nmin = fminf(fminf(nmin, fabsf(nfx)), fabsf(nfy));
    nmin = fminf(fminf(nmin, fabsf(nax)), fabsf(nay));
}

__device__ __forceinline__ bool div_range_ok(float nmin, float nmax, const Consts& c) {
    return nmin >= c.div_lo && nmax <= c.div_hi;
}

__device__ __forceinline__ bool free_step_fast(float x, float y, float vx, float vy, const Consts& c, State& o) {
    float nmin = __int_as_float(0x7f800000), nmax = 0.0f;
    free_step_acc(x, y, vx, vy, c, o, nmin, nmax);
    return div_range_ok(nmin, nmax, c);
}

// ONE AXIS of the free-flight step (ball_sim.py:192-202 acts on x and y separately: the potential and the
// friction clip are per component): (pos, vel) of one coordinate -> the candidate (pos', vel'), `att` = that
// coordinate of the attractor.  Exactly the operations of free_step_acc() / free_step() on one component, so
// two lanes that integrate the two axes of a ball produce the same bits as one lane doing both: the serial
// chain per step is then 24 instructions instead of 49 (rollout_fwd_pipe_kernel's producer warp).
__device__ __forceinline__ void axis_step_acc(float& pos, float& vel, float att, const Consts& c, float& nmin, float& nmax) {
    const float pp = -(2.0f * (pos - att));
    const float nf = (((-clip_nan(vel, -1.0f, 1.0f)) * c.m) * 9.8f) * c.f;
    const float fr = div_markstein(nf, c.r, c.rcp_r);
    const float na = DOPER_FORCE_SUM(c.ks, pp, fr);
    const float a = div_markstein(na, c.m, c.rcp_m);
    const float v1 = vel_update(a, c.dt, vel);
    pos = pos_update(v1, c.dt, pos);
    vel = v1;
    nmax = fmaxf(fmaxf(nmax, fabsf(nf)), fabsf(na));
    nmin = fminf(fminf(nmin, fabsf(nf)), fabsf(na));
}

__device__ __forceinline__ void axis_step(float& pos, float& vel, float att, const Consts& c) {
    const float pp = -(2.0f * (pos - att));
    const float fr = fric(vel, c);
    const float a = div_invariant(DOPER_FORCE_SUM(c.ks, pp, fr), c.m, c.rcp_m, c);
    const float v1 = vel_update(a, c.dt, vel);
    pos = pos_update(v1, c.dt, pos);
    vel = v1;
}

// ball_sim.py:97-132 : time-of-impact sub-step, reflection, remaining sub-step
// (out of line: rare, and keeps the hot step loop small; everything by value = in registers)
__device__ __forceinline__ State resolve_inline(float x, float y, float vx, float vy, const State& s1, const float4& seg,
                                                const Consts& c);

__device__ __noinline__ State resolve(float x, float y, float vx, float vy, State s1, float4 seg,
                                      const Consts c) {
    return resolve_inline(x, y, vx, vy, s1, seg, c);
}

// the same, inlined at the call site (contact loop of the time-parallel kernel: one collision per step)
__device__ __forceinline__ State resolve_inline(float x, float y, float vx, float vy, const State& s1, const float4& seg,
                                                const Consts& c) {
    State o;
    float c1x, c1y, t1;
    proj(s1.x, s1.y, seg, c1x, c1y, t1);
    float ox = c1x - x, oy = c1y - y;  // OLD coordinate (ball_sim.py:100)
    float od = sqrtf(dot2(ox, ox, oy, oy));
    float nx = ox / od, ny = oy / od;
    float toi = fabsf(od - c.r) / fabsf(dot2(nx, s1.vx, ny, s1.vy));
    toi = clip_nan(toi, 0.0f, c.dt);
    float vsx = vel_update(s1.ax, toi, vx), vsy = vel_update(s1.ay, toi, vy);
    float xsx = pos_update(vsx, toi, x), xsy = pos_update(vsy, toi, y);
    float c2x, c2y, t2;
    proj(xsx, xsy, seg, c2x, c2y, t2);
    float o2x = c2x - xsx, o2y = c2y - xsy;
    float o2d = sqrtf(dot2(o2x, o2x, o2y, o2y));
    float n2x = o2x / o2d, n2y = o2y / o2d;
    float k = dot2(n2x, vsx, n2y, vsy);
    float vnx = n2x * k, vny = n2y * k;
    float vpx = vsx - vnx, vpy = vsy - vny;
    float vax = DOPER_REFLECT(vpx, c.e, vnx), vay = DOPER_REFLECT(vpy, c.e, vny);
    float pbx = -(2.0f * (xsx - c.ax)), pby = -(2.0f * (xsy - c.ay));  // no attractor_strength (:124)
    float fbx = fric(vax, c), fby = fric(vay, c);
    float abx = div_invariant(pbx + fbx, c.m, c.rcp_m, c), aby = div_invariant(pby + fby, c.m, c.rcp_m, c);
    float tau = c.dt - toi;
    float vbx = vel_update(abx, tau, vax), vby = vel_update(aby, tau, vay);
    o.x = pos_update(vbx, tau, xsx);
    o.y = pos_update(vby, tau, xsy);
    o.vx = vbx; o.vy = vby; o.ax = abx; o.ay = aby;
    return o;
}

// ---------------------------------------------------------------------------------------
// adjoint.  Operation order mirrors oracle/ball_sim_ref.c:step_vjp so
// that, on a bit-identical forward history, the gradients are bit-identical too.
// ---------------------------------------------------------------------------------------
__device__ __forceinline__ float dfric(float v, const Consts& c) {
    return (v > -1.0f && v < 1.0f) ? -((((c.m) * 9.8f) * c.f) / c.r) : 0.0f;
}

__device__ __forceinline__ float sgnf(float a) { return a > 0.0f ? 1.0f : (a < 0.0f ? -1.0f : 0.0f); }

// J^T w of proj(): [0 < t < 1] sv sv^T / L
__device__ __forceinline__ void proj_vjp(const float4& q, float t, float wx, float wy, float& gx,
                                         float& gy) {
    if (t > 0.0f && t < 1.0f) {
        float L = q.z * q.z + q.w * q.w;
        float s = (wx * q.z + wy * q.w) / L;
        gx = s * q.z; gy = s * q.w;
    } else {
        gx = 0.0f; gy = 0.0f;
    }
}

// Adjoint of one step, split so that the rare collision part lives out of line (it is ~300
// instructions; inlining it into every backward loop blew the instruction cache):
//   step_vjp_tail  free-flight part, common to both branches (x1 = x + v1 dt, v1 = v + a dt, a = ...)
//   step_vjp_hit   collision part: consumes the output cotangent, returns the cotangents of
//                  (x1, v1) in `b1` and the direct contributions to (x, v, a) in `b`/`ba`.
// Operation order is exactly oracle/ball_sim_step.inc:step_vjp.
struct HitAdj {
    float bx, by, bvx, bvy, bax, bay;  // direct cotangent contributions to the step input and to a
    float b1x, b1y, b1vx, b1vy;        // cotangent of the free-flight candidate (x1, v1)
};

// Forward internals of a collided step that its adjoint needs (recomputed from the step's start
// state with the logged segment; identical arithmetic to resolve()).
struct HitFwd {
    float s1vx, s1vy, s1ax, s1ay;   // free-flight candidate (v1, a)
    float t1, t2;                   // clipped projection parameters (candidate point / sub-step point)
    float nx, ny, od, s, den, toi_raw, toi;
    float vsx, vsy, o2d, n2x, n2y, k, vax, vay, abx, aby, tau, vbx, vby;
};

__device__ __forceinline__ HitFwd hit_forward(float x, float y, float vx, float vy, const float4& seg, const Consts& c) {
    HitFwd f;
    State s1;
    free_step(x, y, vx, vy, c, s1);
    f.s1vx = s1.vx; f.s1vy = s1.vy; f.s1ax = s1.ax; f.s1ay = s1.ay;
    float c1x, c1y;
    proj(s1.x, s1.y, seg, c1x, c1y, f.t1);
    float ox = c1x - x, oy = c1y - y;
    f.od = sqrtf(dot2(ox, ox, oy, oy));
    f.nx = ox / f.od; f.ny = oy / f.od;
    f.s = dot2(f.nx, s1.vx, f.ny, s1.vy);
    float num = fabsf(f.od - c.r);
    f.den = fabsf(f.s);
    f.toi_raw = num / f.den;
    f.toi = clip_nan(f.toi_raw, 0.0f, c.dt);
    f.vsx = vel_update(s1.ax, f.toi, vx); f.vsy = vel_update(s1.ay, f.toi, vy);
    float xsx = pos_update(f.vsx, f.toi, x), xsy = pos_update(f.vsy, f.toi, y);
    float c2x, c2y;
    proj(xsx, xsy, seg, c2x, c2y, f.t2);
    float o2x = c2x - xsx, o2y = c2y - xsy;
    f.o2d = sqrtf(dot2(o2x, o2x, o2y, o2y));
    f.n2x = o2x / f.o2d; f.n2y = o2y / f.o2d;
    f.k = dot2(f.n2x, f.vsx, f.n2y, f.vsy);
    float vnx = f.n2x * f.k, vny = f.n2y * f.k;
    f.vax = DOPER_REFLECT(f.vsx - vnx, c.e, vnx); f.vay = DOPER_REFLECT(f.vsy - vny, c.e, vny);
    float pbx = -(2.0f * (xsx - c.ax)), pby = -(2.0f * (xsy - c.ay));
    f.abx = div_invariant(pbx + fric(f.vax, c), c.m, c.rcp_m, c);
    f.aby = div_invariant(pby + fric(f.vay, c), c.m, c.rcp_m, c);
    f.tau = c.dt - f.toi;
    f.vbx = vel_update(f.abx, f.tau, f.vax); f.vby = vel_update(f.aby, f.tau, f.vay);
    return f;
}

// reverse pass of the collision part for ONE output cotangent (gx, gy, gvx, gvy)
//   RCP = false: IEEE divisions in the oracle's order (bit-identical to oracle/ball_sim_step.inc on the same history)
//   RCP = true : every division by a quantity of the forward pass becomes a multiplication by its
//                reciprocal, taken once per collided step (HitRcp).  The chunk-parallel backward, which
//                re-associates the chain of step Jacobians anyway and is compared to the oracle within a
//                tolerance, uses this: a dozen divisions per cotangent were most of its collision path.
struct HitRcp {
    float m, o2d, od, den, L;
};

__device__ __forceinline__ HitRcp hit_reciprocals(const HitFwd& f, const float4& seg, const Consts& c) {
    HitRcp r;
    r.m = c.rcp_m;
    r.o2d = 1.0f / f.o2d; r.od = 1.0f / f.od; r.den = 1.0f / f.den;
    r.L = 1.0f / (seg.z * seg.z + seg.w * seg.w);
    return r;
}

template <bool RCP>
__device__ __forceinline__ float hdiv(float a, float b, float rb) { return RCP ? a * rb : a / b; }

template <bool RCP>
__device__ __forceinline__ void proj_vjp_r(const float4& q, float t, float wx, float wy, float rL, float& gx, float& gy) {
    if (!RCP) { proj_vjp(q, t, wx, wy, gx, gy); return; }
    const float s = (t > 0.0f && t < 1.0f) ? (wx * q.z + wy * q.w) * rL : 0.0f;
    gx = s * q.z; gy = s * q.w;
}

template <bool RCP = false>
__device__ __forceinline__ HitAdj hit_reverse(const HitFwd& f, const float4& seg, const Consts& c,
                                              float gx, float gy, float gvx, float gvy, const HitRcp rc = HitRcp()) {
    float bx = 0.f, by = 0.f, bvx = 0.f, bvy = 0.f, bax = 0.f, bay = 0.f;
    float b1x, b1y, b1vx, b1vy;
    float gvbx = gvx + gx * f.tau, gvby = gvy + gy * f.tau;
    float gxsx = gx, gxsy = gy;
    float gtau = (gx * f.vbx + gy * f.vby) + (gvbx * f.abx + gvby * f.aby);
    float gvax = gvbx, gvay = gvby;
    float gabx = gvbx * f.tau, gaby = gvby * f.tau;
    gxsx += -2.0f * hdiv<RCP>(gabx, c.m, rc.m); gxsy += -2.0f * hdiv<RCP>(gaby, c.m, rc.m);
    gvax += hdiv<RCP>(gabx, c.m, rc.m) * dfric(f.vax, c); gvay += hdiv<RCP>(gaby, c.m, rc.m) * dfric(f.vay, c);
    float gtoi = -gtau;
    float gvsx = gvax, gvsy = gvay;
    float gvnx = -(1.0f + c.e) * gvax, gvny = -(1.0f + c.e) * gvay;
    float gk = gvnx * f.n2x + gvny * f.n2y;
    float gn2x = gvnx * f.k + gk * f.vsx, gn2y = gvny * f.k + gk * f.vsy;
    gvsx += gk * f.n2x; gvsy += gk * f.n2y;
    float nd = f.n2x * gn2x + f.n2y * gn2y;
    float go2x = hdiv<RCP>(gn2x - f.n2x * nd, f.o2d, rc.o2d), go2y = hdiv<RCP>(gn2y - f.n2y * nd, f.o2d, rc.o2d);
    float jx, jy;
    proj_vjp_r<RCP>(seg, f.t2, go2x, go2y, rc.L, jx, jy);
    gxsx += jx - go2x; gxsy += jy - go2y;
    bx += gxsx; by += gxsy;
    gvsx += gxsx * f.toi; gvsy += gxsy * f.toi;
    gtoi += gxsx * f.vsx + gxsy * f.vsy;
    bvx += gvsx; bvy += gvsy;
    bax += gvsx * f.toi; bay += gvsy * f.toi;
    gtoi += gvsx * f.s1ax + gvsy * f.s1ay;
    float graw = (f.toi_raw > 0.0f && f.toi_raw < c.dt) ? gtoi : 0.0f;
    float gnum = hdiv<RCP>(graw, f.den, rc.den);
    float gden = hdiv<RCP>(-(graw * f.toi_raw), f.den, rc.den);
    float god = gnum * sgnf(f.od - c.r);
    float gs = gden * sgnf(f.s);
    float gnx = gs * f.s1vx, gny = gs * f.s1vy;
    b1vx = gs * f.nx; b1vy = gs * f.ny;
    float nd1 = f.nx * gnx + f.ny * gny;
    float gox = hdiv<RCP>(gnx - f.nx * nd1, f.od, rc.od) + god * f.nx, goy = hdiv<RCP>(gny - f.ny * nd1, f.od, rc.od) + god * f.ny;
    bx -= gox; by -= goy;
    proj_vjp_r<RCP>(seg, f.t1, gox, goy, rc.L, b1x, b1y);
    HitAdj r;
    r.bx = bx; r.by = by; r.bvx = bvx; r.bvy = bvy; r.bax = bax; r.bay = bay;
    r.b1x = b1x; r.b1y = b1y; r.b1vx = b1vx; r.b1vy = b1vy;
    return r;
}

__device__ __noinline__ HitAdj step_vjp_hit(float x, float y, float vx, float vy, float4 seg, const Consts c,
                                            float gx, float gy, float gvx, float gvy) {
    const HitFwd f = hit_forward(x, y, vx, vy, seg, c);
    return hit_reverse(f, seg, c, gx, gy, gvx, gvy);
}

// free-flight part of the step adjoint, common to both branches (x1 = x + v1 dt, v1 = v + a dt, a = ...):
// consumes the collision part's result (or the plain output cotangent) -> cotangent of the step input
template <bool RCP = false>
__device__ __forceinline__ void step_vjp_tail(const HitAdj& r, float vx, float vy, const Consts& c, float& gx,
                                              float& gy, float& gvx, float& gvy) {
    float bx = r.bx, by = r.by, bvx = r.bvx, bvy = r.bvy, bax = r.bax, bay = r.bay;
    float tvx = r.b1vx + r.b1x * c.dt, tvy = r.b1vy + r.b1y * c.dt;
    bx += r.b1x; by += r.b1y;
    bvx += tvx; bvy += tvy;
    bax += tvx * c.dt; bay += tvy * c.dt;
    float amx = hdiv<RCP>(bax, c.m, c.rcp_m), amy = hdiv<RCP>(bay, c.m, c.rcp_m);
    bx += -2.0f * (c.ks * amx); by += -2.0f * (c.ks * amy);
    bvx += amx * dfric(vx, c); bvy += amy * dfric(vy, c);
    gx = bx; gy = by; gvx = bvx; gvy = bvy;
}

// N output cotangents through ONE collided step: the forward internals once, N independent reverse
// passes (instruction-level parallelism for the chunk-parallel backward's four basis vectors)
template <int N>
__device__ __noinline__ void step_vjp_hit_n(float x, float y, float vx, float vy, float4 seg, const Consts c,
                                            float (&g)[N][4]) {
    const HitFwd f = hit_forward(x, y, vx, vy, seg, c);
    const HitRcp rc = hit_reciprocals(f, seg, c);
#pragma unroll
    for (int k = 0; k < N; ++k) {
        const HitAdj r = hit_reverse<true>(f, seg, c, g[k][0], g[k][1], g[k][2], g[k][3], rc);
        step_vjp_tail<true>(r, vx, vy, c, g[k][0], g[k][1], g[k][2], g[k][3]);
    }
}

// Adjoint of a step that did NOT collide, from its log word alone: the same operations, in the same
// order, as step_vjp(hit = false) below (bit-identical results), with dfric() read from the logged
// clip bits instead of the recomputed velocity.  dk = dfric's non-zero value, hoisted by the caller:
//   dk = -((((c.m) * 9.8f) * c.f) / c.r)
__device__ __forceinline__ float dfric_value(const Consts& c) { return -((((c.m) * 9.8f) * c.f) / c.r); }

// one coordinate of the free-flight adjoint: (gp, gv) = cotangent of (position, velocity)
__device__ __forceinline__ void step_vjp_free_1d(bool clip_active, float dk, const Consts& c, float& gp, float& gv) {
    float bp = 0.f, bv = 0.f, ba = 0.f;
    const float tv = gv + gp * c.dt;
    bp += gp;
    bv += tv;
    ba += tv * c.dt;
    const float am = div_invariant(ba, c.m, c.rcp_m, c);  // == ba / c.m (correctly rounded either way)
    bp += -2.0f * (c.ks * am);
    bv += am * (clip_active ? dk : 0.0f);
    gp = bp; gv = bv;
}

__device__ __forceinline__ void step_vjp_free(uint32_t w, float dk, const Consts& c, float& gx, float& gy,
                                              float& gvx, float& gvy) {
    step_vjp_free_1d((w & kLogClipX) != 0u, dk, c, gx, gvx);
    step_vjp_free_1d((w & kLogClipY) != 0u, dk, c, gy, gvy);
}

// In/out g = (gx, gy, gvx, gvy): cotangent of the step output -> cotangent of the step input.
__device__ __forceinline__ void step_vjp(float x, float y, float vx, float vy, bool hit,
                                         const float4& seg, const Consts& c, float& gx, float& gy,
                                         float& gvx, float& gvy) {
    float bx = 0.f, by = 0.f, bvx = 0.f, bvy = 0.f, bax = 0.f, bay = 0.f;
    float b1x, b1y, b1vx, b1vy;
    if (hit) {
        const HitAdj r = step_vjp_hit(x, y, vx, vy, seg, c, gx, gy, gvx, gvy);
        bx = r.bx; by = r.by; bvx = r.bvx; bvy = r.bvy; bax = r.bax; bay = r.bay;
        b1x = r.b1x; b1y = r.b1y; b1vx = r.b1vx; b1vy = r.b1vy;
    } else {
        b1x = gx; b1y = gy; b1vx = gvx; b1vy = gvy;
    }
    float tvx = b1vx + b1x * c.dt, tvy = b1vy + b1y * c.dt;
    bx += b1x; by += b1y;
    bvx += tvx; bvy += tvy;
    bax += tvx * c.dt; bay += tvy * c.dt;
    float amx = bax / c.m, amy = bay / c.m;
    bx += -2.0f * (c.ks * amx); by += -2.0f * (c.ks * amy);
    bvx += amx * dfric(vx, c); bvy += amy * dfric(vy, c);
    gx = bx; gy = by; gvx = bvx; gvy = bvy;
}

}  // namespace doper
