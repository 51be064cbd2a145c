/*
 * Plain-C restatement of the reference rollout, forward + hand-derived adjoint.
 *
 * TEST INFRASTRUCTURE ONLY (see oracle/__init__.py): the checker for the CUDA
 * path and the timed "cpu_baseline" of bench.py.  Never linked into, imported by
 * or used as a fallback for the product library.  PARITY UNPINNED by the
 * reference except for the Euler/force arithmetic (notebook cell 7); the forward
 * here is asserted bit-identical to oracle/ball_sim_np.py and the backward is
 * asserted against torch autograd (oracle/ball_sim_torch.py) in tests/.
 *
 * Follows reference doper/sim/ball_sim.py:15-132,171-253,256-338 and
 * doper/sim/jax_geometry.py:12-24,41-53,97-177 with the arithmetic rules of
 * binary32 everywhere, dot = a0*b0 + a1*b1 (no FMA — build
 * with -ffp-contract=off, no -ffast-math), IEEE division and sqrt, NaN-propagating
 * clip, argmin = first minimal index with NaN counting as minimal.
 *
 * Data layout (identical to the product C-ABI, include/doper_b200.h):
 *   segs   [S][4]  (x0,y0,x1,y1)  or, per-environment, [B][S][4]
 *   x0,v0  [B][2];  hitlog [n_steps][B] int32 = idx | (hit << 31)
 *   traj_x/v/a [n_steps][B][2]
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

/* Which mul+add pairs of the forward arithmetic are single-rounding FMAs is decided in ONE header shared with
 * the product kernels (include/doper_fp_policy.h; default: only get_new_state's coordinate update,
 * ball_sim.py:76, which the reference's recorded output pins as contracted).  Build with
 * -DDOPER_FP_POLICY=<mask> for another policy (oracle/cref.py build(policy=...)). fmaf() is exact. */
#include "../include/doper_fp_policy.h"

#define R float
#define N(name) name##_f
#define SQRT sqrtf
#define FABS fabsf
#define LIT(x) x##f
#define POS_UPDATE(v, dt, x) DOPER_POS_UPDATE(v, dt, x)
#define VEL_UPDATE(a, dt, v) DOPER_VEL_UPDATE(a, dt, v)
#define LERP(s0, t, sv) DOPER_LERP(s0, t, sv)
#define DOT2(a0, b0, a1, b1) DOPER_DOT2(a0, b0, a1, b1)
#define FORCE_SUM(ks, p, f) DOPER_FORCE_SUM(ks, p, f)
#define REFLECT(vp, e, vn) DOPER_REFLECT(vp, e, vn)
#include "ball_sim_step.inc"
#undef R
#undef N
#undef SQRT
#undef FABS
#undef LIT
#undef POS_UPDATE
#undef VEL_UPDATE
#undef LERP
#undef DOT2
#undef FORCE_SUM
#undef REFLECT

/* binary64 instantiation (gradient truth): the real-arithmetic formulas, no policy */
#define R double
#define N(name) name##_d
#define SQRT sqrt
#define FABS fabs
#define LIT(x) x
#define POS_UPDATE(v, dt, x) ((x) + (v) * (dt))
#define VEL_UPDATE(a, dt, v) ((v) + (a) * (dt))
#define LERP(s0, t, sv) ((s0) + (t) * (sv))
#define DOT2(a0, b0, a1, b1) ((a0) * (b0) + (a1) * (b1))
#define FORCE_SUM(ks, p, f) ((ks) * (p) + (f))
#define REFLECT(vp, e, vn) ((vp) - (e) * (vn))
#include "ball_sim_step.inc"
#undef R
#undef N
#undef SQRT
#undef FABS
#undef LIT
#undef POS_UPDATE
#undef VEL_UPDATE
#undef LERP
#undef DOT2
#undef FORCE_SUM
#undef REFLECT

/* the fp32 instantiation under its historical names */
typedef consts_t_f consts_t;
typedef state_t_f state_t;
#define clipf clipf_f
#define proj proj_f
#define fric fric_f
#define free_step free_step_f
#define resolve resolve_f
#define step_vjp step_vjp_f

/* jax_geometry.py:151-152 */
static inline float seg_dist(float px, float py, const float *s) {
    float cx, cy, t;
    proj(px, py, s, &cx, &cy, &t);
    float dx = px - cx, dy = py - cy;
    return sqrtf(DOPER_DOT2(dx, dx, dy, dy));
}

/* jax_geometry.py:172-174; first minimal index, NaN counts as minimal */
static inline int closest(float px, float py, const float *segs, int S, float *dist) {
    int best = 0, best_nan = 0;
    float bd = INFINITY;
    for (int i = 0; i < S; ++i) {
        float d = seg_dist(px, py, segs + 4 * (size_t)i);
        if (d != d) {
            if (!best_nan) { best = i; bd = d; best_nan = 1; }
        } else if (!best_nan && (d < bd || i == 0)) {
            best = i; bd = d;
        }
    }
    *dist = bd;
    return best;
}

/* ball_sim.py:171-209 */
static inline int32_t sim_step(state_t *st, const float *segs, int S, const float *A,
                               const consts_t *c, float dt) {
    state_t s1;
    free_step(st->x, st->y, st->vx, st->vy, A, c, dt, &s1);
    float dist;
    int idx = closest(s1.x, s1.y, segs, S, &dist);
    int hit = dist <= c->r; /* NaN -> 0 */
    if (hit) {
        state_t s2;
        resolve(st->x, st->y, st->vx, st->vy, &s1, segs + 4 * (size_t)idx, A, c, dt, &s2);
        *st = s2;
    } else {
        *st = s1;
    }
    return (int32_t)((uint32_t)idx | ((uint32_t)hit << 31));
}

static consts_t mk(const float *k) {
    consts_t c = {k[0], k[1], k[2], k[3], k[4], 9.8f};
    return c;
}

/* ---- exported ---------------------------------------------------------- */

int oracle_fp_policy(void) { return DOPER_FP_POLICY; }

void oracle_closest_segment(int64_t N, int S, const float *pts, const float *segs,
                            int32_t *idx, float *dist) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) idx[i] = closest(pts[2 * i], pts[2 * i + 1], segs, S, &dist[i]);
}

/* jax_geometry.py:97-116 with :21-23 and :51-53 */
void oracle_points_inside(int64_t N, int n_tri, const float *pts, const float *segs,
                          uint8_t *flag) {
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < N; ++i) {
        float px = pts[2 * i], py = pts[2 * i + 1];
        int any = 0;
        for (int t = 0; t < n_tri; ++t) {
            int all = 1;
            for (int e = 0; e < 3; ++e) {
                const float *s = segs + 4 * (size_t)(3 * t + e);
                float svx = s[2] - s[0], svy = s[3] - s[1];
                float nx = svy, ny = -svx;
                float nrm = sqrtf(DOPER_DOT2(nx, nx, ny, ny));
                float nhx = nx / nrm, nhy = ny / nrm;
                float pvx = px - s[0], pvy = py - s[1];
                float d = DOPER_DOT2(pvx, nhx, pvy, nhy);
                all &= (d < 0.0f); /* sign(d) < 0 ; NaN -> false */
            }
            any |= all;
        }
        flag[i] = (uint8_t)any;
    }
}

/* Forward rollout. seg_stride = 0 (shared scene) or S*4 floats per ball (per-env).
 * consts = {radius, mass, friction, elasticity, attractor_strength}. Optional outputs may be NULL. */
void oracle_rollout_fwd(int64_t B, int S, int64_t seg_stride, int n_steps, float dt,
                        const float *consts, const float *attractor, const float *segs,
                        const float *x0, const float *v0, float *xT, float *vT, float *traj_x,
                        float *traj_v, float *traj_a, int32_t *hitlog) {
    consts_t c = mk(consts);
#pragma omp parallel for schedule(static)
    for (int64_t b = 0; b < B; ++b) {
        const float *sg = segs + seg_stride * b;
        state_t st = {x0[2 * b], x0[2 * b + 1], v0[2 * b], v0[2 * b + 1], 0.0f, 0.0f};
        for (int t = 0; t < n_steps; ++t) {
            int32_t h = sim_step(&st, sg, S, attractor, &c, dt);
            size_t o = ((size_t)t * B + b);
            if (hitlog) hitlog[o] = h;
            if (traj_x) { traj_x[2 * o] = st.x; traj_x[2 * o + 1] = st.y; }
            if (traj_v) { traj_v[2 * o] = st.vx; traj_v[2 * o + 1] = st.vy; }
            if (traj_a) { traj_a[2 * o] = st.ax; traj_a[2 * o + 1] = st.ay; }
        }
        xT[2 * b] = st.x; xT[2 * b + 1] = st.y;
        vT[2 * b] = st.vx; vT[2 * b + 1] = st.vy;
    }
}

/* value_and_grad of sum_b |x_T - target|_1 w.r.t. v0 (and x0).  Stores the whole per-step
 * (x,v) history per ball (what XLA's transposed scan keeps), then walks it backwards.
 * Outputs: loss_per_ball [B], xT, vT [B][2], gv0, gx0 [B][2]; hitlog optional [n][B]. */
void oracle_value_and_grad(int64_t B, int S, int64_t seg_stride, int n_steps, float dt,
                           const float *consts, const float *attractor, const float *target,
                           const float *segs, const float *x0, const float *v0, float *loss_pb,
                           float *xT, float *vT, float *gv0, float *gx0, int32_t *hitlog) {
    consts_t c = mk(consts);
#pragma omp parallel
    {
        float *hist = (float *)malloc(sizeof(float) * 4 * (size_t)(n_steps > 0 ? n_steps : 1));
        int32_t *hl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_steps > 0 ? n_steps : 1));
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            const float *sg = segs + seg_stride * b;
            state_t st = {x0[2 * b], x0[2 * b + 1], v0[2 * b], v0[2 * b + 1], 0.0f, 0.0f};
            for (int t = 0; t < n_steps; ++t) {
                hist[4 * t] = st.x; hist[4 * t + 1] = st.y;
                hist[4 * t + 2] = st.vx; hist[4 * t + 3] = st.vy;
                hl[t] = sim_step(&st, sg, S, attractor, &c, dt);
                if (hitlog) hitlog[(size_t)t * B + b] = hl[t];
            }
            xT[2 * b] = st.x; xT[2 * b + 1] = st.y; vT[2 * b] = st.vx; vT[2 * b + 1] = st.vy;
            float ex = st.x - target[0], ey = st.y - target[1];
            loss_pb[b] = fabsf(ex) + fabsf(ey);
            float gx = ex > 0.f ? 1.f : (ex < 0.f ? -1.f : 0.f);
            float gy = ey > 0.f ? 1.f : (ey < 0.f ? -1.f : 0.f);
            float gvx = 0.f, gvy = 0.f;
            for (int t = n_steps - 1; t >= 0; --t)
                step_vjp(hist[4 * t], hist[4 * t + 1], hist[4 * t + 2], hist[4 * t + 3], hl[t], sg,
                         attractor, &c, dt, &gx, &gy, &gvx, &gvy);
            gv0[2 * b] = gvx; gv0[2 * b + 1] = gvy;
            if (gx0) { gx0[2 * b] = gx; gx0[2 * b + 1] = gy; }
        }
        free(hist);
        free(hl);
    }
}

/* Gradient "truth": fp32 forward rollout (bit-identical to oracle_value_and_grad's), then the
 * binary64 adjoint of the real-arithmetic step evaluated at each fp32 step-start state with the
 * logged (idx, hit).  Any fp32 ordering of the backward (the reference's XLA transpose, torch
 * autograd, the sequential or the chunk-parallel CUDA kernels) is this plus rounding noise. */
void oracle_grad_truth64(int64_t B, int S, int64_t seg_stride, int n_steps, float dt,
                         const float *consts, const float *attractor, const float *target,
                         const float *segs, const float *x0, const float *v0, double *gv0,
                         double *gx0) {
    consts_t c = mk(consts);
    consts_t_d cd = {c.r, c.m, c.f, c.e, c.ks, c.g};
    const double Ad[2] = {attractor[0], attractor[1]};
#pragma omp parallel
    {
        float *hist = (float *)malloc(sizeof(float) * 4 * (size_t)(n_steps > 0 ? n_steps : 1));
        int32_t *hl = (int32_t *)malloc(sizeof(int32_t) * (size_t)(n_steps > 0 ? n_steps : 1));
        double *sd = (double *)malloc(sizeof(double) * 4 * (size_t)S);
#pragma omp for schedule(static)
        for (int64_t b = 0; b < B; ++b) {
            const float *sg = segs + seg_stride * b;
            for (int i = 0; i < 4 * S; ++i) sd[i] = sg[i];
            state_t st = {x0[2 * b], x0[2 * b + 1], v0[2 * b], v0[2 * b + 1], 0.0f, 0.0f};
            for (int t = 0; t < n_steps; ++t) {
                hist[4 * t] = st.x; hist[4 * t + 1] = st.y;
                hist[4 * t + 2] = st.vx; hist[4 * t + 3] = st.vy;
                hl[t] = sim_step(&st, sg, S, attractor, &c, dt);
            }
            float ex = st.x - target[0], ey = st.y - target[1];
            double gx = ex > 0.f ? 1. : (ex < 0.f ? -1. : 0.), gy = ey > 0.f ? 1. : (ey < 0.f ? -1. : 0.);
            double gvx = 0., gvy = 0.;
            for (int t = n_steps - 1; t >= 0; --t)
                step_vjp_d(hist[4 * t], hist[4 * t + 1], hist[4 * t + 2], hist[4 * t + 3], hl[t], sd, Ad,
                           &cd, (double)dt, &gx, &gy, &gvx, &gvy);
            gv0[2 * b] = gvx; gv0[2 * b + 1] = gvy;
            if (gx0) { gx0[2 * b] = gx; gx0[2 * b + 1] = gy; }
        }
        free(hist); free(hl); free(sd);
    }
}
