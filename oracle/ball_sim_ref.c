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

/* get_new_state's coordinate update x + v'*dt (ball_sim.py:76) is the one expression the
 * reference's recorded output (notebook cell 7) shows was contracted to a single-rounding FMA by
 * the authors' XLA:CPU; see oracle/ball_sim_np.py CONTRACT_POSITION_UPDATE. fmaf() is exact. */
#ifndef ORACLE_NO_CONTRACT
#define POS_UPDATE(v, dt, x) fmaf((v), (dt), (x))
#else
#define POS_UPDATE(v, dt, x) ((x) + (v) * (dt))
#endif

typedef struct {
    float r, m, f, e, ks, g;
} consts_t;

static inline float clipf(float x, float lo, float hi) {
    if (x != x) return x; /* NaN propagates like jnp.clip */
    x = x < lo ? lo : x;
    return x > hi ? hi : x;
}

/* jax_geometry.py:132-136 */
static inline void proj(float px, float py, const float *s, float *cx, float *cy, float *tt) {
    float svx = s[2] - s[0], svy = s[3] - s[1];
    float L = svx * svx + svy * svy;
    float pvx = px - s[0], pvy = py - s[1];
    float t = clipf((pvx * svx + pvy * svy) / L, 0.0f, 1.0f);
    *cx = s[0] + t * svx;
    *cy = s[1] + t * svy;
    *tt = t;
}

/* jax_geometry.py:151-152 */
static inline float seg_dist(float px, float py, const float *s) {
    float cx, cy, t;
    proj(px, py, s, &cx, &cy, &t);
    float dx = px - cx, dy = py - cy;
    return sqrtf(dx * dx + dy * dy);
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

/* ball_sim.py:43-45 */
static inline float fric(float v, const consts_t *c) {
    return ((((-clipf(v, -1.0f, 1.0f)) * c->m) * c->g) * c->f) / c->r;
}

typedef struct { float x, y, vx, vy, ax, ay; } state_t;

/* ball_sim.py:192-202: forces + Euler. Returns candidate state (x1,v1,a). */
static inline void free_step(float x, float y, float vx, float vy, const float *A,
                             const consts_t *c, float dt, state_t *o) {
    float px = -(2.0f * (x - A[0])), py = -(2.0f * (y - A[1]));
    float fx = fric(vx, c), fy = fric(vy, c);
    float ax = (c->ks * px + fx) / c->m, ay = (c->ks * py + fy) / c->m;
    float v1x = vx + ax * dt, v1y = vy + ay * dt;
    o->x = POS_UPDATE(v1x, dt, x); o->y = POS_UPDATE(v1y, dt, y);
    o->vx = v1x; o->vy = v1y; o->ax = ax; o->ay = ay;
}

/* ball_sim.py:97-132 */
static inline void resolve(float x, float y, float vx, float vy, const state_t *s1,
                           const float *seg, const float *A, const consts_t *c, float dt,
                           state_t *o) {
    float c1x, c1y, t1;
    proj(s1->x, s1->y, seg, &c1x, &c1y, &t1);
    float ox = c1x - x, oy = c1y - y;
    float od = sqrtf(ox * ox + oy * oy);
    float nx = ox / od, ny = oy / od;
    float toi = fabsf(od - c->r) / fabsf(nx * s1->vx + ny * s1->vy);
    toi = clipf(toi, 0.0f, dt);
    float vsx = vx + s1->ax * toi, vsy = vy + s1->ay * toi;
    float xsx = POS_UPDATE(vsx, toi, x), xsy = POS_UPDATE(vsy, toi, y);
    float c2x, c2y, t2;
    proj(xsx, xsy, seg, &c2x, &c2y, &t2);
    float o2x = c2x - xsx, o2y = c2y - xsy;
    float o2d = sqrtf(o2x * o2x + o2y * o2y);
    float n2x = o2x / o2d, n2y = o2y / o2d;
    float k = n2x * vsx + n2y * vsy;
    float vnx = n2x * k, vny = n2y * k;
    float vpx = vsx - vnx, vpy = vsy - vny;
    float vax = vpx - c->e * vnx, vay = vpy - c->e * vny;
    float pbx = -(2.0f * (xsx - A[0])), pby = -(2.0f * (xsy - A[1]));
    float fbx = fric(vax, c), fby = fric(vay, c);
    float abx = (pbx + fbx) / c->m, aby = (pby + fby) / c->m;
    float tau = dt - toi;
    float vbx = vax + abx * tau, vby = vay + aby * tau;
    o->x = POS_UPDATE(vbx, tau, xsx); o->y = POS_UPDATE(vby, tau, xsy);
    o->vx = vbx; o->vy = vby; o->ax = abx; o->ay = aby;
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
                float nrm = sqrtf(nx * nx + ny * ny);
                float nhx = nx / nrm, nhy = ny / nrm;
                float pvx = px - s[0], pvy = py - s[1];
                float d = pvx * nhx + pvy * nhy;
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

/* d fric / d v  (ball_sim.py:43-45; clip passes gradient strictly inside) */
static inline float dfric(float v, const consts_t *c) {
    return (v > -1.0f && v < 1.0f) ? -((((c->m) * c->g) * c->f) / c->r) : 0.0f;
}

/* J^T w for proj(): [0<t<1] * sv sv^T / L  (symmetric) */
static inline void proj_vjp(const float *s, float t, float wx, float wy, float *gx, float *gy) {
    if (t > 0.0f && t < 1.0f) {
        float svx = s[2] - s[0], svy = s[3] - s[1];
        float L = svx * svx + svy * svy;
        float q = (wx * svx + wy * svy) / L;
        *gx = q * svx; *gy = q * svy;
    } else {
        *gx = 0.0f; *gy = 0.0f;
    }
}

/* Adjoint of one sim_step given the state at step start and the logged (idx,hit).
 * In/out: (gx,gy,gvx,gvy) = cotangent of the step OUTPUT (x',v') -> cotangent of INPUT (x,v).
 * SURVEY.md Appendix B. The branch (not select) form: no 0*inf leakage. */
static inline void step_vjp(float x, float y, float vx, float vy, int32_t h, const float *segs,
                            const float *A, const consts_t *c, float dt, float *gx, float *gy,
                            float *gvx, float *gvy) {
    state_t s1;
    free_step(x, y, vx, vy, A, c, dt, &s1);
    float bx = 0.f, by = 0.f, bvx = 0.f, bvy = 0.f; /* cotangent of step input */
    float bax = 0.f, bay = 0.f;                     /* cotangent of a */
    float b1x, b1y, b1vx, b1vy;                     /* cotangent of (x1, v1) */
    if (h < 0) {
        const float *seg = segs + 4 * (size_t)(h & 0x7fffffff);
        /* recompute resolve() intermediates */
        float c1x, c1y, t1;
        proj(s1.x, s1.y, seg, &c1x, &c1y, &t1);
        float ox = c1x - x, oy = c1y - y;
        float od = sqrtf(ox * ox + oy * oy);
        float nx = ox / od, ny = oy / od;
        float s = nx * s1.vx + ny * s1.vy;
        float num = fabsf(od - c->r), den = fabsf(s);
        float toi_raw = num / den;
        float toi = clipf(toi_raw, 0.0f, dt);
        float vsx = vx + s1.ax * toi, vsy = vy + s1.ay * toi;
        float xsx = POS_UPDATE(vsx, toi, x), xsy = POS_UPDATE(vsy, toi, y);
        float c2x, c2y, t2;
        proj(xsx, xsy, seg, &c2x, &c2y, &t2);
        float o2x = c2x - xsx, o2y = c2y - xsy;
        float o2d = sqrtf(o2x * o2x + o2y * o2y);
        float n2x = o2x / o2d, n2y = o2y / o2d;
        float k = n2x * vsx + n2y * vsy;
        float vnx = n2x * k, vny = n2y * k;
        float vax = (vsx - vnx) - c->e * vnx, vay = (vsy - vny) - c->e * vny;
        float pbx = -(2.0f * (xsx - A[0])), pby = -(2.0f * (xsy - A[1]));
        float abx = (pbx + fric(vax, c)) / c->m, aby = (pby + fric(vay, c)) / c->m;
        float tau = dt - toi;
        float vbx = vax + abx * tau, vby = vay + aby * tau;
        /* reverse */
        float gvbx = *gvx + *gx * tau, gvby = *gvy + *gy * tau; /* xb = xs + vb*tau */
        float gxsx = *gx, gxsy = *gy;
        float gtau = (*gx * vbx + *gy * vby) + (gvbx * abx + gvby * aby);
        float gvax = gvbx, gvay = gvby;
        float gabx = gvbx * tau, gaby = gvby * tau;
        gxsx += -2.0f * (gabx / c->m); gxsy += -2.0f * (gaby / c->m);
        gvax += (gabx / c->m) * dfric(vax, c); gvay += (gaby / c->m) * dfric(vay, c);
        float gtoi = -gtau;
        /* va = vs - (1+e) * n2 * (n2.vs) */
        float gvsx = gvax, gvsy = gvay;
        float gvnx = -(1.0f + c->e) * gvax, gvny = -(1.0f + c->e) * gvay;
        float gk = gvnx * n2x + gvny * n2y;
        float gn2x = gvnx * k + gk * vsx, gn2y = gvny * k + gk * vsy;
        gvsx += gk * n2x; gvsy += gk * n2y;
        float nd = n2x * gn2x + n2y * gn2y;
        float go2x = (gn2x - n2x * nd) / o2d, go2y = (gn2y - n2y * nd) / o2d;
        float jx, jy;
        proj_vjp(seg, t2, go2x, go2y, &jx, &jy);
        gxsx += jx - go2x; gxsy += jy - go2y;
        /* xs = x + vs*toi ; vs = v + a*toi */
        bx += gxsx; by += gxsy;
        gvsx += gxsx * toi; gvsy += gxsy * toi;
        gtoi += gxsx * vsx + gxsy * vsy;
        bvx += gvsx; bvy += gvsy;
        bax += gvsx * toi; bay += gvsy * toi;
        gtoi += gvsx * s1.ax + gvsy * s1.ay;
        /* toi = clip(num/den, 0, dt) */
        float graw = (toi_raw > 0.0f && toi_raw < dt) ? gtoi : 0.0f;
        float gnum = graw / den;
        float gden = -(graw * toi_raw) / den;
        float god = gnum * ((od - c->r) > 0.0f ? 1.0f : ((od - c->r) < 0.0f ? -1.0f : 0.0f));
        float gs = gden * (s > 0.0f ? 1.0f : (s < 0.0f ? -1.0f : 0.0f));
        float gnx = gs * s1.vx, gny = gs * s1.vy;
        b1vx = gs * nx; b1vy = gs * ny;
        float nd1 = nx * gnx + ny * gny;
        float gox = (gnx - nx * nd1) / od + god * nx, goy = (gny - ny * nd1) / od + god * ny;
        bx -= gox; by -= goy;
        proj_vjp(seg, t1, gox, goy, &b1x, &b1y);
    } else {
        b1x = *gx; b1y = *gy; b1vx = *gvx; b1vy = *gvy;
    }
    /* common tail: x1 = x + v1*dt ; v1 = v + a*dt ; a = (ks*pot + fric(v))/m */
    float tvx = b1vx + b1x * dt, tvy = b1vy + b1y * dt;
    bx += b1x; by += b1y;
    bvx += tvx; bvy += tvy;
    bax += tvx * dt; bay += tvy * dt;
    float amx = bax / c->m, amy = bay / c->m;
    bx += -2.0f * (c->ks * amx); by += -2.0f * (c->ks * amy);
    bvx += amx * dfric(vx, c); bvy += amy * dfric(vy, c);
    *gx = bx; *gy = by; *gvx = bvx; *gvy = bvy;
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
