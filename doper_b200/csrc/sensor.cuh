// Range-sensor ray casting: the observation path that runs right before
// every rollout.  Reference: doper/sim/jax_geometry.py:180-261 (line_ray_intersection_point,
// batch_line_ray_intersection_point), doper/agents/observations.py:29-110 (SimpleRangeSensor /
// UndirectedRangeSensor.get_observation), doper/agents/range_sensor_agent.py:17-41
// (RangeSensingAgent.get_observations).
//
// Arithmetic is the reference's, op for op in fp32 (this TU is compiled with -fmad=false, IEEE
// division / sqrt):  d = dir / |dir|;  v1 = o - s0;  v2 = s1 - s0;  v3 = (-d.y, d.x);
// denom = v2 . v3;  t1 = cross(v2, v1) / denom;  t2 = (v1 . v3) / denom;  miss when denom == 0,
// t1 < 0, t2 < 0 or t2 > 1 (point = (inf, inf)), else point = o + t1 d.
// Included by kernels.cu (needs SceneLayout / stage_scenes / make_key from there).
#pragma once

namespace doper {

struct RayHit { float px, py; bool hit; };

// jax_geometry.py:180-218 for an already normalised direction d (the reference normalises per pair;
// the value does not depend on the segment) and a pre-digested segment (s0, v2 = s1 - s0)
__device__ __forceinline__ RayHit ray_segment(float ox, float oy, float dx, float dy, float s0x, float s0y,
                                              float v2x, float v2y) {
    RayHit r;
    r.px = __int_as_float(0x7f800000); r.py = r.px; r.hit = false;
    const float v1x = ox - s0x, v1y = oy - s0y;
    const float v3x = -dy, v3y = dx;
    const float denom = v2x * v3x + v2y * v3y;
    if (denom == 0.0f) return r;
    const float cr = v2x * v1y - v2y * v1x;   // jnp.cross of two 2-vectors
    const float n2 = v1x * v3x + v1y * v3y;
    // cheap, conservative rejection with an approximate reciprocal (relative error < 1e-6): a quotient
    // that is clearly negative / clearly above one stays so in IEEE arithmetic; anything close to a
    // boundary (or non-finite) takes the two exact divisions below.  Never accepts anything.
    const float inv = __fdividef(1.0f, denom);
    const float t1a = cr * inv, t2a = n2 * inv;
    if (t1a < -1.0e-30f || t2a < -1.0e-30f || t2a > 1.0001f) return r;
    const float t1 = cr / denom;
    const float t2 = n2 / denom;
    if (t1 < 0.0f || t2 < 0.0f || t2 > 1.0f) return r;
    r.px = ox + t1 * dx; r.py = oy + t1 * dy;
    r.hit = true;
    return r;
}

__device__ __forceinline__ void normalise(float dx, float dy, float& ux, float& uy) {
    const float nrm = sqrtf(dx * dx + dy * dy);  // np.linalg.norm of a 2-vector
    ux = dx / nrm; uy = dy / nrm;
}

// batch_line_ray_intersection_point: [R] rays x [S] raw segments -> points [R][S][2]
__global__ void __launch_bounds__(256) ray_points_kernel(int64_t R, int S, const float2* __restrict__ origins,
                                                        const float2* __restrict__ dirs,
                                                        const float4* __restrict__ raw, float2* __restrict__ out) {
    const int64_t e = (int64_t)blockIdx.x * 256 + threadIdx.x;
    if (e >= R * (int64_t)S) return;
    const int64_t r = e / S;
    const int s = (int)(e - r * S);
    const float2 o = origins[r], dd = dirs[r];
    float ux, uy;
    normalise(dd.x, dd.y, ux, uy);
    const float4 g = raw[s];
    const RayHit h = ray_segment(o.x, o.y, ux, uy, g.x, g.y, g.z - g.x, g.w - g.y);
    out[e] = make_float2(h.px, h.py);
}

// Fused sensor: one thread per (position, ray); the scene is staged in shared memory (TMA).
// ranges [N][R] (clamped to distance_range), optional closest points [N][R][2] and indices [N][R];
// OBS: also / instead write RangeSensingAgent's observation rows [N][7 + R].
constexpr int kSensorThreads = 128;

// STAGED / per_env as in closest_segment_kernel (position i looks at scene i).
// POS64: the positions are float64 (the reference's numpy code keeps a float64 position as it is,
// observations.py:50; jax rounds the ray origins to float32 for the intersections, :62-65; the ranges
// `norm(points - position)` are then float64, :67): intersections from the rounded origin, ranges in binary64
// against the float64 position, float64 outputs.
template <bool OBS, bool STAGED, bool POS64>
__global__ void __launch_bounds__(kSensorThreads) range_sensor_kernel(
    int64_t N, int R, int S, int per_env, const void* __restrict__ pos_v, const float2* __restrict__ vel,
    const float2* __restrict__ ray_dirs, const unsigned char* __restrict__ scene_ws, float distance_range,
    double tx, double ty, void* __restrict__ ranges_v, float2* __restrict__ points, int32_t* __restrict__ idx_out,
    float* __restrict__ obs) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneLayout L = scene_layout(per_env ? N : 1, S);
    const int tid = threadIdx.x;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)L.S_pad * 20);
    if (STAGED) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) stage_scenes(reinterpret_cast<float4*>(smem), reinterpret_cast<float*>(smem + (size_t)L.S_pad * sizeof(float4)),
                                   scene_ws, L, 0, 1, bar);
    }
    const int64_t e = (int64_t)blockIdx.x * kSensorThreads + tid;
    const bool active = e < N * (int64_t)R;
    const int64_t i = active ? e / R : 0;
    const int r = active ? (int)(e - i * R) : 0;
    double ox64 = 0.0, oy64 = 0.0;
    float2 o;
    if (POS64) {
        const double2 o64 = reinterpret_cast<const double2*>(pos_v)[i];
        ox64 = o64.x; oy64 = o64.y;
        o = make_float2((float)o64.x, (float)o64.y);
    } else {
        o = reinterpret_cast<const float2*>(pos_v)[i];
    }
    const float2 dd = ray_dirs[r];
    float ux, uy;
    normalise(dd.x, dd.y, ux, uy);
    const float4* sq = STAGED ? reinterpret_cast<const float4*>(smem)
                              : reinterpret_cast<const float4*>(scene_ws + L.q_off) + (per_env ? i : 0) * L.S_pad;
    if (STAGED) mbar_wait(bar, 0);
    unsigned long long key = kNoKey;
    double best64 = __longlong_as_double(0x7ff0000000000000ll);  // POS64: (range, first index) minimum in binary64
    int best_i = 0;
    float bx = __int_as_float(0x7f800000), by = bx;
    for (int j = 0; j < S; ++j) {
        const float4 q = sq[j];
        const RayHit h = ray_segment(o.x, o.y, ux, uy, q.x, q.y, q.z, q.w);
        if (h.hit) {
            if (POS64) {
                const double ex = (double)h.px - ox64, ey = (double)h.py - oy64;
                const double rng = sqrt(ex * ex + ey * ey);  // observations.py:67 in float64
                if (rng < best64) { best64 = rng; best_i = j; bx = h.px; by = h.py; }
            } else {
                const float ex = h.px - o.x, ey = h.py - o.y;
                const float rng = sqrtf(ex * ex + ey * ey);  // observations.py:67
                const unsigned long long k = make_key(rng, j);
                if (k < key) { key = k; bx = h.px; by = h.py; }
            }
        }
    }
    if (!active) return;
    // all rays missed: every range is +inf, argmin = 0 (observations.py:70)
    int idx = 0;
    float rng = __int_as_float(0x7f800000);
    double rng64 = best64;
    if (POS64) {
        idx = best_i;
        if (rng64 > (double)distance_range) { bx = __int_as_float(0x7f800000); by = bx; rng64 = (double)distance_range; }
    } else {
        if (key != kNoKey) unpack_key(key, idx, rng);
        if (rng > distance_range) {  // observations.py:73-74 / :77 (NaN stays NaN)
            bx = __int_as_float(0x7f800000); by = bx;
            rng = distance_range;
        }
    }
    if (ranges_v) {
        if (POS64) reinterpret_cast<double*>(ranges_v)[e] = rng64;
        else reinterpret_cast<float*>(ranges_v)[e] = rng;
    }
    if (points) points[e] = make_float2(bx, by);
    if (idx_out) idx_out[e] = idx;
    if (OBS) {
        const int W = 7 + R;
        float* row = obs + (size_t)i * W;
        row[7 + r] = rng / distance_range;  // range_sensor_agent.py:40
        if (r == 0) {
            // range_sensor_agent.py:32-35: float64 numpy arithmetic (the target is a float64 array), cast at :41
            const double ddx = (double)o.x - tx, ddy = (double)o.y - ty;
            const double dist = sqrt(ddx * ddx + ddy * ddy);
            const float2 v = vel[i];
            row[0] = (float)dist;
            row[1] = (float)(ddx / dist);
            row[2] = (float)(ddy / dist);
            row[3] = o.x; row[4] = o.y;
            row[5] = v.x; row[6] = v.y;
        }
    }
}

}  // namespace doper
