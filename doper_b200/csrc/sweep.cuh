// Closest-segment sweep: argmin_j dist(p, segment j) with the reference's exact semantics
// (first minimal index on the sqrt'd fp32 distance, NaN counts as minimal;
// reference doper/sim/jax_geometry.py:141-177).
//
// A group of G lanes (G | 32, consecutive lanes of one warp) cooperates on one point: lane j
// visits segments j, j+G, ...; the per-lane winners are merged with a warp-shuffle
// lexicographic min on (distance, index).
//
// Two implementations, identical results:
//   sweep_exact     evaluates the reference arithmetic for every pair.
//   sweep_filtered  evaluates a cheap FMA-based estimate of the squared distance for every pair
//                   (9 FP32 instructions instead of ~30) and the reference arithmetic only for
//                   the few pairs whose estimate is within a proven error margin of the best
//                   one; the estimate never decides anything.
#pragma once
#include "physics.cuh"

namespace doper {

// Pre-digested scene in shared (or global) memory.
struct SceneView {
    const float4* q;  // [S_pad] (s0x, s0y, svx, svy)
    const float* rl;  // [S_pad] 1 / (svx^2 + svy^2)   (IEEE reciprocal)
    int S;            // real segments
    int S_pad;        // padded with far-away sentinels to a multiple of 32
    float scale;      // max |coordinate| over the scene
    int weird;        // 1: scene has a degenerate / non-finite segment -> exact sweep only
};

// Uniform grid over a shared scene (built by scene_prepare_kernel): makes the candidate-list refresh
// O(segments near the ball) instead of O(S).  The cell side is at least the largest segment extent,
// so a segment's bounding box touches at most 2 x 2 cells; every segment is registered in every cell
// its box overlaps.  A query for "all segments closer than R to c" visits the cells that the square
// [c - R, c + R]^2 overlaps: a segment within R of c has a point of its box inside that square, and
// the coordinate -> cell map is the same monotone function on both sides, so it is registered in a
// visited cell (a superset is returned; duplicates across cells are dropped by the list builders).
constexpr int kGridMaxCells = 4096;

struct GridHeader {  // 32 B
    float x0, y0, inv_h;
    int gx, gy, enabled, n_items, pad;
};

struct GridView {
    const int* cell_start;  // [gx * gy + 1]
    const int* items;       // segment indices, cell-major
    float x0, y0, inv_h;
    int gx, gy, enabled;
};

__host__ __device__ __forceinline__ int grid_cell_coord(float v, float v0, float inv_h, int g) {
    const float f = floorf((v - v0) * inv_h);
    return f < 0.0f ? 0 : (f >= (float)g ? g - 1 : (int)f);  // NaN -> g - 1 branch is never taken: NaN < 0 false, >= false -> (int)NaN
}

// cells overlapped by the square of half side R around (cx, cy); false when the grid is off, the query
// is not finite, or it covers so many cells that the plain scan is cheaper
__device__ __forceinline__ bool grid_range(const GridView& g, float cx, float cy, float R, int& ix0, int& ix1,
                                           int& iy0, int& iy1) {
    if (!g.enabled || !(R < 1.0e30f) || !(fabsf(cx) < 1.0e30f) || !(fabsf(cy) < 1.0e30f)) return false;
    ix0 = grid_cell_coord(cx - R, g.x0, g.inv_h, g.gx); ix1 = grid_cell_coord(cx + R, g.x0, g.inv_h, g.gx);
    iy0 = grid_cell_coord(cy - R, g.y0, g.inv_h, g.gy); iy1 = grid_cell_coord(cy + R, g.y0, g.inv_h, g.gy);
    return 2 * (ix1 - ix0 + 1) * (iy1 - iy0 + 1) <= g.gx * g.gy;
}

constexpr unsigned long long kNoKey = 0xffffffffffffffffull;

// (distance, index) -> totally ordered key; NaN sorts first, then ascending distance, then index.
__device__ __forceinline__ unsigned long long make_key(float d, int idx) {
    unsigned int hi = (d != d) ? 0u : (__float_as_uint(d) + 1u);  // d >= +0 whenever not NaN
    return ((unsigned long long)hi << 32) | (unsigned int)idx;
}

template <int G>
__device__ __forceinline__ unsigned long long group_min(unsigned long long key) {
#pragma unroll
    for (int off = G / 2; off > 0; off >>= 1) {
        unsigned long long o = __shfl_xor_sync(0xffffffffu, key, off);
        key = o < key ? o : key;
    }
    return key;
}

__device__ __forceinline__ void unpack_key(unsigned long long key, int& idx, float& d) {
    idx = (int)(unsigned int)(key & 0xffffffffull);
    unsigned int hi = (unsigned int)(key >> 32);
    d = hi == 0u ? __int_as_float(0x7fc00000) : __uint_as_float(hi - 1u);
}

// Reference arithmetic for every pair. The compare runs on d^2 and takes the sqrt only when d^2
// improves: sqrt is monotone, so d2 >= best_d2 implies sqrt(d2) >= best_d (no update, the earlier
// index keeps a tie) and d2 < best_d2 needs the sqrt'd compare to keep the earlier index when the
// two sqrt's round to the same value.
template <int G>
__device__ __noinline__ unsigned long long lane_exact(const SceneView& sc, float px, float py,
                                                         int lane) {
    float best_d2 = __int_as_float(0x7f800000), best_d = best_d2;
    int best_i = 0x7fffffff, nan_i = 0x7fffffff;
    for (int i = lane; i < sc.S; i += G) {
        float4 q = sc.q[i];
        float d2 = seg_dist2(px, py, q);
        if (d2 != d2) {  // NaN distance: argmin returns the first one, whatever else is there
            nan_i = min(nan_i, i);
        } else if (d2 < best_d2 || best_i == 0x7fffffff) {
            float d = sqrtf(d2);
            if (d < best_d || best_i == 0x7fffffff) { best_d = d; best_i = i; }
            best_d2 = d2;
        }
    }
    if (nan_i != 0x7fffffff) return make_key(__int_as_float(0x7fc00000), nan_i);
    if (best_i != 0x7fffffff) return make_key(best_d, best_i);
    return kNoKey;
}

// ---------------------------------------------------------------------------------------
// filtered sweep
// ---------------------------------------------------------------------------------------
// Estimate of the squared distance: same formula, FMA-contracted, reciprocal instead of the
// division, saturate instead of clip. 9 FP32 instructions.
__device__ __forceinline__ float approx_d2(float px, float py, const float4& q, float rl) {
    float pvx = px - q.x, pvy = py - q.y;
    float dot = __fmaf_rn(pvx, q.z, pvy * q.w);
    float t = __saturatef(dot * rl);
    float ex = __fmaf_rn(-t, q.z, pvx), ey = __fmaf_rn(-t, q.w, pvy);
    return __fmaf_rn(ex, ex, ey * ey);
}

// Error model: for a non-degenerate scene and a finite point with
// M = max(|coordinate|) over scene and point, both |sqrt(estimate) - D| and |d_ref - D| are
// bounded by 49 u M and 38 u M (u = 2^-24, D the real-arithmetic distance), so
// |sqrt(a_i) - d_ref_i| <= delta = 128 u M (measured worst case: 7.3 u M, tools/filter_margin_check.py).
// The winner i* of the reference argmin therefore satisfies, for ANY j,
//   a_i* <= (sqrt(a_j) + 2 delta)^2 <= a_j (1 + eps) + 4 delta^2 (1 + 1/eps)      (AM-GM)
// and every segment passing that test against the running minimum estimate is kept as a
// candidate; candidates get the reference arithmetic and the lexicographic (distance, index) min.
constexpr float kFilterOneEps = 1.0f + 0.0078125f;            // eps = 2^-7
constexpr float kFilterKappaPerM2 = 4.0f * 129.0f * 1.02f *   // 4 (1 + 1/eps), 2 % slack for the
                                    (7.62939453125e-6f * 7.62939453125e-6f);  // roundings; (128 u)^2

// Reference distance of ONE candidate. The division is skipped exactly when the clip decides
// alone: dot <= 0 -> t = 0 and dot >= L -> t = 1 (L > 0 finite on this path; RN is monotone).
__device__ __forceinline__ float cand_dist(float px, float py, const float4& q) {
    const float L = dot2(q.z, q.z, q.w, q.w);
    const float pvx = px - q.x, pvy = py - q.y;
    const float dot = dot2(pvx, q.z, pvy, q.w);
    float t = 0.0f;
    if (dot > 0.0f) t = (dot >= L) ? 1.0f : dot / L;
    const float cx = DOPER_LERP(q.x, t, q.z), cy = DOPER_LERP(q.y, t, q.w);
    const float dx = px - cx, dy = py - cy;
    return sqrtf(dot2(dx, dx, dy, dy));
}

// One lane's share of the filtered sweep: branch-free estimate loop that records candidates as
// bits (bit j = j-th segment this lane visited since `base`), then the reference arithmetic for
// the set bits.  The threshold is fixed for the whole sweep: thr = U(estimate of last step's
// winner); the ball moves < radius per step, so that winner is almost always still within a
// hair of the minimum and the candidate set is the true winner plus its exact ties.
template <int G>
__device__ __forceinline__ unsigned long long lane_filtered(const SceneView& sc, float px, float py,
                                                            int lane, int prev_idx, float kappa) {
    const float thr = __fmaf_rn(approx_d2(px, py, sc.q[prev_idx], sc.rl[prev_idx]), kFilterOneEps, kappa);
    unsigned long long key = kNoKey;
    unsigned mask = 0u, bit = 1u;
    int base = lane, i = lane;
    auto flush = [&]() {
        while (mask) {
            const int j = __ffs(mask) - 1;
            mask &= mask - 1u;
            const int c = base + j * G;
            const unsigned long long kk = make_key(cand_dist(px, py, sc.q[c]), c);
            key = kk < key ? kk : key;
        }
    };
    for (; i + 3 * G < sc.S_pad; i += 4 * G) {
        const float a0 = approx_d2(px, py, sc.q[i], sc.rl[i]);
        const float a1 = approx_d2(px, py, sc.q[i + G], sc.rl[i + G]);
        const float a2 = approx_d2(px, py, sc.q[i + 2 * G], sc.rl[i + 2 * G]);
        const float a3 = approx_d2(px, py, sc.q[i + 3 * G], sc.rl[i + 3 * G]);
        mask |= (a0 <= thr ? bit : 0u) | (a1 <= thr ? bit << 1 : 0u) | (a2 <= thr ? bit << 2 : 0u) |
                (a3 <= thr ? bit << 3 : 0u);
        bit <<= 4;
        if (bit == 0u) {  // 32 visits recorded: evaluate them, start a new window
            flush();
            base = i + 4 * G;
            bit = 1u;
        }
    }
    for (; i < sc.S_pad; i += G) {  // at most 3 visits: the window cannot overflow here
        const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
        mask |= (a <= thr) ? bit : 0u;
        bit <<= 1;
    }
    flush();
    return key;
}

// min over the G lanes of a group.  A full warp uses two REDUX (distance bits, then index among
// the minima); sub-warp groups use a shuffle butterfly (REDUX with per-group masks is serialised
// over the distinct masks of the warp: measured 4x slower at G = 8).
template <int G>
__device__ __forceinline__ unsigned long long group_min_redux(unsigned long long key) {
    if (G == 1) return key;
    if (G == 32) {
        const unsigned hi = (unsigned)(key >> 32), lo = (unsigned)key;
        const unsigned hmin = __reduce_min_sync(0xffffffffu, hi);
        const unsigned lmin = __reduce_min_sync(0xffffffffu, hi == hmin ? lo : 0xffffffffu);
        return ((unsigned long long)hmin << 32) | lmin;
    }
    return group_min<G>(key);
}

// No temporal hint (first step, one-shot queries): a branch-free pre-pass finds the segment with
// the smallest estimate; it only seeds the threshold, so any index would be correct.
template <int G>
__device__ __forceinline__ int hint_by_estimate(const SceneView& sc, float px, float py, int lane) {
    float best = __int_as_float(0x7f800000);
    int bi = 0;
    for (int i = lane; i < sc.S_pad; i += G) {
        const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
        if (a < best) { best = a; bi = i; }
    }
    const unsigned long long k = group_min_redux<G>(((unsigned long long)__float_as_uint(best) << 32) | (unsigned)bi);
    const int h = (int)(unsigned)(k & 0xffffffffull);
    return h < sc.S ? h : 0;
}

// ---------------------------------------------------------------------------------------
// broad phase: per-ball candidate lists
// ---------------------------------------------------------------------------------------
// A list is built around a centre c with a radius rho and stays valid for every query point p
// with |p - c| <= rho.  With e_j = sqrt(estimate_j(c)), D_j the real-arithmetic distance (which is
// 1-Lipschitz in the point), |e_j - D_j| <= 49 u M and |d_ref_j - D_j| <= 38 u M (error model of the
// filter above), the reference argmin i* at p (and every exact tie) satisfies, for ANY segment j0,
//     e_i* <= D_i*(c) + 49uM <= D_i*(p) + rho + 49uM <= D_j0(p) + rho + (49 + 76)uM
//          <= D_j0(c) + 2 rho + 125uM <= e_j0 + 2 rho + 174 u M.
// The list keeps every j with e_j <= e_j0 + 2 rho + 256 u M (j0 = last step's winner; any j0 is
// valid, a near one keeps the list short), evaluated on squares with upward slack.  Per step the
// filtered sweep then runs over the list only: it contains the argmin and all its ties with their
// original indices, so the result is bit-identical to the full sweep.  M bounds the scene and every
// point of the list's lifetime.  A lane whose share overflows its capacity sweeps all its
// segments until the next refresh (always correct).
//
// Hit lists (the default; `need_argmin` = false).  The winning index is consumed at COLLISIONS only, and a
// step collides iff some d_ref_j(p) <= r.  For |p - c| <= rho such a j has
//     e_j <= D_j(c) + 49uM <= D_j(p) + rho + 49uM <= d_ref_j(p) + rho + 87uM <= r + rho + 87 u M,
// so the list {j : e_j <= rho + r + 128 u M} holds every segment that can touch the ball anywhere in
// the disc — the argmin of a colliding step and all its exact ties included — and a segment outside
// the list has d_ref_j(p) > r + 41uM at every point of the disc.  Its radius rho + r does not depend on
// the distance to the nearest wall (the argmin list needs e_j0 + 2 rho), so in open space it is
// short or empty.  `need_argmin` (exact index log: the index of EVERY step is wanted) keeps the
// argmin list.
constexpr float kCullDeltaPerM = 256.0f * 5.9604644775390625e-08f;  // 256 u, u = 2^-24
constexpr float kCullUp = 1.0f + 9.5367431640625e-07f;              // 1 + 2^-20: covers the roundings of T
constexpr float kCullRhoCheck = 0.98f;                              // the guard compares against (0.98 rho)^2

constexpr float kCullSafePerM = 128.0f * 5.9604644775390625e-08f;   // 128 u >= 49 u + 38 u: see cull_far()

struct CullState {
    float cx, cy;  // centre of the list
    float rho2;    // (kCullRhoCheck rho)^2; the list is used only while |p - c|^2 <= rho2
    int cnt;       // entries of this lane's share; > LCAP = overflow (sweep everything)
    float far2;    // estimates above this prove "no collision this step" (see cull_refresh)
};

__device__ __forceinline__ float cull_far2(float radius, float M);

// Largest estimate a list entry may have (see above): e_j0 + 2 rho + 256 u M for an argmin list,
// rho + r + 128 u M for a hit list; rounded up.
__device__ __forceinline__ float list_reach(const SceneView& sc, float px, float py, float rho, float radius, float M,
                                            int j0, bool need_argmin) {
    if (need_argmin) {
        const float e = sqrtf(approx_d2(px, py, sc.q[j0], sc.rl[j0])) * kCullUp;
        return e + __fmaf_rn(kCullDeltaPerM, M, rho + rho);
    }
    return __fmaf_rn(kCullSafePerM, M, rho + radius) * kCullUp;
}

// list entry k of this lane lives at list[k * stride]
template <int G, int LCAP, bool GRID>
__device__ __forceinline__ void cull_refresh(const SceneView& sc, const GridView& grid, float px, float py, float rho,
                                             float radius, int lane, int j0, unsigned short* list, int stride,
                                             CullState& cs, bool need_argmin) {
    const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)) + rho);
    const float reach = list_reach(sc, px, py, rho, radius, M, j0, need_argmin);  // kept: e_j <= reach
    const float T = (reach * reach) * kCullUp;
    int cnt = 0;
    int ix0, ix1, iy0, iy1;
    // every kept segment is closer than (e + w) + 49 u M to the centre: the grid cells around it hold them all
    if (GRID && grid_range(grid, px, py, __fmaf_rn(kCullSafePerM, M, reach) * kCullUp, ix0, ix1, iy0, iy1)) {
        for (int cy = iy0; cy <= iy1; ++cy) {
            const int row = cy * grid.gx;
            const int k1 = grid.cell_start[row + ix1 + 1];
            for (int k = grid.cell_start[row + ix0]; k < k1; ++k) {  // the cells of one grid row are contiguous
                const int i = grid.items[k];
                if ((i & (G - 1)) != lane) continue;  // this lane owns the segments with i % G == lane
                if (approx_d2(px, py, sc.q[i], sc.rl[i]) <= T) {
                    bool dup = false;  // a segment is registered in up to four cells
                    for (int m = 0; m < min(cnt, LCAP); ++m) dup = dup || list[m * stride] == (unsigned short)i;
                    if (!dup) {
                        if (cnt < LCAP) list[cnt * stride] = (unsigned short)i;
                        ++cnt;
                    }
                }
            }
        }
    } else {
        for (int i = lane; i < sc.S_pad; i += G) {
            const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
            if (a <= T) {
                if (cnt < LCAP) list[cnt * stride] = (unsigned short)i;
                ++cnt;
            }
        }
    }
    cs.cnt = cnt; cs.cx = px; cs.cy = py;
    cs.rho2 = (kCullRhoCheck * rho) * (kCullRhoCheck * rho);
    cs.far2 = cull_far2(radius, M);  // M covers every point of the list's lifetime
}

// No-collision proof from estimates alone: if every list entry has estimate^2 > far2 with
//   far2 >= (radius + 128 u M)^2, then est_j > r + 87uM, so D_j > r + 38uM and d_ref_j > r for the
// argmin (which is in the list), hence min_j d_ref_j > r: the step cannot collide, and neither the
// exact distance nor the winning index is needed (the index is only consumed at collisions).
__device__ __forceinline__ float cull_far2(float radius, float M) {
    const float s = __fmaf_rn(kCullSafePerM, M, radius) * kCullUp;
    return (s * s) * kCullUp;
}

// min of a non-negative float over the G lanes of a group (every lane of the warp calls)
template <int G>
__device__ __forceinline__ float group_min_f(float v) {
    if constexpr (G == 1) {
        return v;
    } else if constexpr (G == 32) {
        return __uint_as_float(__reduce_min_sync(0xffffffffu, __float_as_uint(v)));  // v >= +0: bit order = value order
    } else {
#pragma unroll
        for (int off = G / 2; off > 0; off >>= 1) v = fminf(v, __shfl_xor_sync(0xffffffffu, v, off));
        return v;
    }
}

// smallest estimate over this lane's share of the list (all its segments when the share overflowed)
template <int G, int LCAP>
__device__ __forceinline__ float lane_list_min(const SceneView& sc, float px, float py, int lane,
                                               const unsigned short* list, int stride, const CullState& cs) {
    float amin = __int_as_float(0x7f800000);
    if (cs.cnt > LCAP) {
        for (int i = lane; i < sc.S_pad; i += G) amin = fminf(amin, approx_d2(px, py, sc.q[i], sc.rl[i]));
    } else {
        for (int k = 0; k < cs.cnt; ++k) {
            const int i = list[k * stride];
            amin = fminf(amin, approx_d2(px, py, sc.q[i], sc.rl[i]));
        }
    }
    return amin;
}

// this lane's share of the exact evaluation: candidates are the list entries whose estimate passes
// thr (thr derived from the estimate of ANY segment is valid; the caller passes the group minimum)
template <int G, int LCAP>
__device__ __noinline__ unsigned long long lane_culled(const SceneView& sc, float px, float py, int lane,
                                                       float thr, const unsigned short* list, int stride,
                                                       int cnt) {
    unsigned long long key = kNoKey;
    if (cnt > LCAP) {
        for (int i = lane; i < sc.S_pad; i += G) {
            const float4 q = sc.q[i];
            if (approx_d2(px, py, q, sc.rl[i]) <= thr) {
                const unsigned long long kk = make_key(cand_dist(px, py, q), i);
                key = kk < key ? kk : key;
            }
        }
        return key;
    }
    for (int k = 0; k < cnt; ++k) {
        const int i = list[k * stride];
        const float4 q = sc.q[i];
        if (approx_d2(px, py, q, sc.rl[i]) <= thr) {
            const unsigned long long kk = make_key(cand_dist(px, py, q), i);
            key = kk < key ? kk : key;
        }
    }
    return key;
}

// group argmin of the estimate over the current list (seed j0 of the next refresh)
template <int G, int LCAP>
__device__ __forceinline__ int list_argmin(const SceneView& sc, float px, float py, int lane,
                                           const unsigned short* list, int stride, const CullState& cs, int fallback) {
    float amin = __int_as_float(0x7f800000);
    int imin = fallback;
    if (cs.cnt <= LCAP) {
        for (int k = 0; k < cs.cnt; ++k) {
            const int i = list[k * stride];
            const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
            if (a < amin) { amin = a; imin = i; }
        }
    }
    const unsigned long long k = group_min_redux<G>(((unsigned long long)__float_as_uint(amin) << 32) | (unsigned)imin);
    const int h = (int)(unsigned)(k & 0xffffffffull);
    return (h >= 0 && h < sc.S) ? h : fallback;
}

// ---------------------------------------------------------------------------------------
// per-BALL candidate list (time-parallel broad-phase kernel: the G lanes of a group hold G
// consecutive time steps of one ball and all read the same list)
// ---------------------------------------------------------------------------------------
struct GroupList {
    unsigned short* e;  // [cap] entries in shared memory, ascending segment index
    int cap;
    int cnt;            // > cap = overflow: sweep every segment
    float cx, cy, rho2, far2;  // as CullState
    // Safe disc: with e_min = the smallest estimate over EVERY segment at the centre, a point p of the
    // list's lifetime has d_ref_j(p) >= D_j(p) - 38uM >= D_j(c) - |p - c| - 38uM >= e_min - |p - c| - 87uM
    // for all j, so |p - c| < e_min - r - 87uM proves "no collision at p" from the displacement alone —
    // the list is not even read.  safe2 = (sqrt(min estimate^2) - (r + 128 u M))^2 (the 41uM of extra
    // slack cover the roundings of the square root, of the difference and of |p - c|^2), capped by
    // rho2 (M bounds only points within rho of the centre); negative = no safe disc.
    float safe2;
    float emin;  // that smallest estimate (0 before the first full scan): sizes the NEXT disc, a heuristic only
};

// Every lane of the warp calls (warp-wide ballots); `doit` is uniform within a group.  Segment
// i0 + lane of each stride-G slice is tested by lane `lane`; passing indices are appended in
// ascending order (ballot + popc prefix inside the group).
template <int G, bool GRID>
__device__ __forceinline__ void glist_refresh(const SceneView& sc, const GridView& grid, float px, float py, float rho,
                                              float radius, int lane, int gbase, int j0, bool doit, GroupList& gl,
                                              bool need_argmin) {
    const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)) + rho);
    const float reach = list_reach(sc, px, py, rho, radius, M, j0, need_argmin);
    const float T = (reach * reach) * kCullUp;
    const unsigned low = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    int cnt = 0;
    float e2min = __int_as_float(0x7f800000);  // smallest squared estimate over all segments (full scan only)
    bool scanned_all = false;
    int ix0 = 0, ix1 = -1, iy0 = 0, iy1 = -1;
    const bool by_grid = GRID && doit && grid_range(grid, px, py, __fmaf_rn(kCullSafePerM, M, reach) * kCullUp, ix0, ix1, iy0, iy1);
    if (GRID && __all_sync(0xffffffffu, by_grid || !doit)) {
        // every refreshing group walks its own cells (row by row: the cells of a grid row are contiguous);
        // the warp iterates until the longest walk is done (ballots need every lane)
        int cy = iy0, k = 0, k1 = 0;
        bool row_open = false;
        while (true) {
            if (doit && !row_open && cy <= iy1) {
                k = grid.cell_start[cy * grid.gx + ix0]; k1 = grid.cell_start[cy * grid.gx + ix1 + 1];
                row_open = true;
            }
            const bool busy = doit && row_open;
            if (!__any_sync(0xffffffffu, busy)) break;
            bool pass = false;
            int i = 0;
            if (busy && k + lane < k1) {
                i = grid.items[k + lane];
                pass = approx_d2(px, py, sc.q[i], sc.rl[i]) <= T;
                if (pass)  // a segment is registered in up to four cells: drop what the list already holds
                    for (int m = 0; m < min(cnt, gl.cap); ++m) pass = pass && gl.e[m] != (unsigned short)i;
            }
            const unsigned gb = (__ballot_sync(0xffffffffu, pass) >> gbase) & low;
            const int pos = cnt + __popc(gb & ((1u << lane) - 1u));
            if (pass && pos < gl.cap) gl.e[pos] = (unsigned short)i;
            cnt += __popc(gb);
            __syncwarp();  // the next round's duplicate check reads what this round appended
            if (busy) {
                k += G;
                if (k >= k1) { row_open = false; ++cy; }
            }
        }
    } else {
        // four slices of G segments per round: the four estimates (shared-memory loads + a 9-deep FP chain
        // each) and the four ballots are independent, so their latencies overlap
        constexpr int U = 4;
        const unsigned below = (1u << lane) - 1u;
        for (int i0 = 0; i0 < sc.S_pad; i0 += U * G) {
            float a[U];
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int i = i0 + u * G + lane;
                a[u] = __int_as_float(0x7f800000);
                if (doit && i < sc.S_pad) a[u] = approx_d2(px, py, sc.q[i], sc.rl[i]);
            }
            unsigned gb[U];
#pragma unroll
            for (int u = 0; u < U; ++u) gb[u] = (__ballot_sync(0xffffffffu, a[u] <= T) >> gbase) & low;
#pragma unroll
            for (int u = 0; u < U; ++u) {
                const int pos = cnt + __popc(gb[u] & below);
                if (a[u] <= T && pos < gl.cap) gl.e[pos] = (unsigned short)(i0 + u * G + lane);
                cnt += __popc(gb[u]);
                e2min = fminf(e2min, a[u]);
            }
        }
        e2min = group_min_f<G>(e2min);
        scanned_all = true;
    }
    __syncwarp();
    if (doit) {
        gl.cnt = cnt; gl.cx = px; gl.cy = py;
        gl.rho2 = (kCullRhoCheck * rho) * (kCullRhoCheck * rho);
        gl.far2 = cull_far2(radius, M);
        const float emin = sqrtf(e2min);
        const float sigma = emin - __fmaf_rn(kCullSafePerM, M, radius) * kCullUp;
        gl.emin = scanned_all ? emin : 0.0f;
        gl.safe2 = (scanned_all && sigma > 0.0f) ? fminf((sigma * sigma) * (1.0f - 9.5367431640625e-07f), gl.rho2) : -1.0f;
    }
}

// smallest estimate over the WHOLE list for this lane's own point
__device__ __forceinline__ float glist_min(const SceneView& sc, float px, float py, const GroupList& gl) {
    float amin = __int_as_float(0x7f800000);
    if (gl.cnt > gl.cap) {
        for (int i = 0; i < sc.S_pad; ++i) amin = fminf(amin, approx_d2(px, py, sc.q[i], sc.rl[i]));
    } else {
        int k = 0;
        for (; k + 1 < gl.cnt; k += 2) {
            const int i0 = gl.e[k], i1 = gl.e[k + 1];
            const float a0 = approx_d2(px, py, sc.q[i0], sc.rl[i0]);
            const float a1 = approx_d2(px, py, sc.q[i1], sc.rl[i1]);
            amin = fminf(amin, fminf(a0, a1));
        }
        if (k < gl.cnt) { const int i = gl.e[k]; amin = fminf(amin, approx_d2(px, py, sc.q[i], sc.rl[i])); }
    }
    return amin;
}

// exact (distance, index) of the closest segment for this lane's own point, over the whole list
__device__ __noinline__ unsigned long long glist_exact(const SceneView& sc, float px, float py, float thr,
                                                       const unsigned short* e, int cap, int cnt) {
    unsigned long long key = kNoKey;
    const int n = cnt > cap ? sc.S_pad : cnt;
    for (int k = 0; k < n; ++k) {
        const int i = cnt > cap ? k : (int)e[k];
        const float4 q = sc.q[i];
        if (approx_d2(px, py, q, sc.rl[i]) <= thr) {
            const unsigned long long kk = make_key(cand_dist(px, py, q), i);
            key = kk < key ? kk : key;
        }
    }
    return key;
}

// list entry with the smallest estimate at (px, py): seed of the next refresh (any index is valid)
__device__ __forceinline__ int glist_argmin(const SceneView& sc, float px, float py, const GroupList& gl, int fallback) {
    float amin = __int_as_float(0x7f800000);
    int imin = fallback;
    if (gl.cnt <= gl.cap) {
        for (int k = 0; k < gl.cnt; ++k) {
            const int i = gl.e[k];
            const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
            if (a < amin) { amin = a; imin = i; }
        }
    }
    return (imin >= 0 && imin < sc.S) ? imin : fallback;
}

// Group-cooperative sweep. Every lane of the warp must call this (warp-wide sync primitives);
// lanes of one group pass the same point. `prev_idx` in [0,S) is a hint (never trusted); < 0 =
// no hint available.
template <int G>
__device__ __forceinline__ void sweep_any(const SceneView& sc, float px, float py, int lane,
                                          int prev_idx, int& idx, float& dist) {
    const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
    const bool fast = !sc.weird && M <= 1.0e15f;  // false for NaN / inf points
    if (prev_idx < 0) {                           // uniform across the warp (step counter / query)
        const int h = hint_by_estimate<G>(sc, fast ? px : 0.0f, fast ? py : 0.0f, lane);
        prev_idx = h;
    }
    unsigned long long key;
    if (fast) {
        key = lane_filtered<G>(sc, px, py, lane, prev_idx, kFilterKappaPerM2 * M * M);
    } else {
        key = lane_exact<G>(sc, px, py, lane);
    }
    key = group_min_redux<G>(key);
    unpack_key(key, idx, dist);
}

}  // namespace doper
