// queries.cuh: L1 loss, closest-segment / inner-check queries, FP32 issue-rate probe — part of the doper_b200 CUDA translation unit (included by kernels.cu).
#pragma once
#include "scene.cuh"

namespace doper {

// ------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------
constexpr int kLossThreads = 256;

// block 0: deterministic total (fixed strided partials + fixed tree); blocks >= 1: per-ball outputs
// (a device function of `block`: loss_jacobian_kernel, bwd.cuh, runs the same blocks next to the backward's pass 1)
__device__ __forceinline__ void l1_loss_block(const unsigned block, int64_t B, const float2* __restrict__ xT, float tx, float ty,
                                              float* __restrict__ loss_pb, float* __restrict__ loss_sum,
                                              float2* __restrict__ xT_bar) {
    if (block == 0) {
        __shared__ float part[kLossThreads];
        float acc = 0.0f;
        for (int64_t b = threadIdx.x; b < B; b += kLossThreads) {
            float2 p = xT[b];
            acc += fabsf(p.x - tx) + fabsf(p.y - ty);  // ball_sim.py:286
        }
        part[threadIdx.x] = acc;
        __syncthreads();
        for (int s = kLossThreads / 2; s > 0; s >>= 1) {
            if ((int)threadIdx.x < s) part[threadIdx.x] += part[threadIdx.x + s];
            __syncthreads();
        }
        if (threadIdx.x == 0) *loss_sum = part[0];  // ball_sim.py:324
        return;
    }
    const int64_t b = (int64_t)(block - 1) * kLossThreads + threadIdx.x;
    if (b >= B) return;
    float2 p = xT[b];
    float ex = p.x - tx, ey = p.y - ty;
    if (loss_pb) loss_pb[b] = fabsf(ex) + fabsf(ey);
    if (xT_bar) xT_bar[b] = make_float2(sgnf(ex), sgnf(ey));  // d|e|/de, 0 at e == 0 (and NaN -> 0)
}

__global__ void __launch_bounds__(kLossThreads) l1_loss_kernel(int64_t B, const float2* __restrict__ xT,
                                                              float tx, float ty,
                                                              float* __restrict__ loss_pb,
                                                              float* __restrict__ loss_sum,
                                                              float2* __restrict__ xT_bar) {
    l1_loss_block(blockIdx.x, B, xT, tx, ty, loss_pb, loss_sum, xT_bar);
}

// ------------------------------------------------------------------------------------------
// geometry queries
// ------------------------------------------------------------------------------------------
constexpr int kQueryThreads = 128;

// STAGED: the (shared) scene is copied into shared memory by TMA; else it is read in place (a scene that does not
// fit shared memory, or per-environment scenes: point i belongs to scene i, `per_env` = 1, N scenes in scene_ws).
template <int G, bool STAGED>
__global__ void __launch_bounds__(kQueryThreads) closest_segment_kernel(
    int64_t N, int S, int per_env, const float2* __restrict__ pts, const float4* __restrict__ raw,
    const unsigned char* __restrict__ scene_ws, int32_t* __restrict__ idx_out,
    float* __restrict__ dist_out, float4* __restrict__ seg_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneLayout L = scene_layout(per_env ? N : 1, S);
    const int tid = threadIdx.x, lane = tid % G;
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)L.S_pad * 20);
    if (STAGED) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) stage_scenes(reinterpret_cast<float4*>(smem), reinterpret_cast<float*>(smem + (size_t)L.S_pad * sizeof(float4)),
                                   scene_ws, L, 0, 1, bar);
    }
    int64_t i = (int64_t)blockIdx.x * (kQueryThreads / G) + tid / G;
    const bool active = i < N;
    if (!active) i = N - 1;
    const int64_t env = per_env ? i : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(scene_ws + L.hdr_off)[env];
    SceneView sc;
    if (STAGED) {
        sc.q = reinterpret_cast<const float4*>(smem);
        sc.rl = reinterpret_cast<const float*>(smem + (size_t)L.S_pad * sizeof(float4));
    } else {
        sc.q = reinterpret_cast<const float4*>(scene_ws + L.q_off) + env * L.S_pad;
        sc.rl = reinterpret_cast<const float*>(scene_ws + L.rl_off) + env * L.S_pad;
    }
    sc.S = S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird;
    const float2 p = pts[i];
    if (STAGED) mbar_wait(bar, 0);
    int idx; float dist;
    sweep_any<G>(sc, p.x, p.y, lane, -1, idx, dist);
    if (active && lane == 0) {
        idx_out[i] = idx;
        dist_out[i] = dist;
        if (seg_out) seg_out[i] = raw[env * S + idx];
    }
}

constexpr int kInsideTile = 3 * 1024;  // segments per shared-memory tile (multiple of 3)

// jax_geometry.py:12-24,41-53,97-116. The unit normal is point-independent: computed once per
// block per segment with the reference's arithmetic, then reused for every point.
__global__ void __launch_bounds__(kQueryThreads) points_inside_kernel(int64_t N, int S,
                                                                     const float2* __restrict__ pts,
                                                                     const float4* __restrict__ raw,
                                                                     uint8_t* __restrict__ flag) {
    __shared__ float4 tile[kInsideTile];  // (s0x, s0y, nhx, nhy)
    const int64_t i = (int64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    float px = 0.f, py = 0.f;
    if (i < N) { float2 p = pts[i]; px = p.x; py = p.y; }
    bool any = false;
    for (int base = 0; base < S; base += kInsideTile) {
        const int n = min(kInsideTile, S - base);
        __syncthreads();
        for (int j = threadIdx.x; j < n; j += kQueryThreads) {
            float4 s = raw[base + j];
            float svx = s.z - s.x, svy = s.w - s.y;      // :21
            float nx = svy, ny = -svx;                   // :22
            float nrm = sqrtf(dot2(nx, nx, ny, ny));      // :23
            tile[j] = make_float4(s.x, s.y, nx / nrm, ny / nrm);
        }
        __syncthreads();
        for (int t = 0; t + 2 < n; t += 3) {
            bool all = true;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                float4 q = tile[t + e];
                float pvx = px - q.x, pvy = py - q.y;    // :52
                float d = dot2(pvx, q.z, pvy, q.w);          // :53
                all = all && (d < 0.0f);                 // sign(d) < 0 ; NaN -> false
            }
            any = any || all;
        }
    }
    if (i < N) flag[i] = any ? 1 : 0;
}

// ------------------------------------------------------------------------------------------
// FP32 issue-rate probe
// ------------------------------------------------------------------------------------------
__global__ void fp32_probe_kernel(int iters, int use_fma, float* out) {
    float a[16];
#pragma unroll
    for (int k = 0; k < 16; ++k) a[k] = 1.0f + 1e-3f * (float)(threadIdx.x + k);
    const float m = 0.9999999f, cc = 1.0e-4f;
    if (use_fma) {
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __fmaf_rn(a[k], m, cc);
        }
    } else {
        for (int it = 0; it < iters; it += 2) {
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __fmul_rn(a[k], m);
#pragma unroll
            for (int k = 0; k < 16; ++k) a[k] = __fadd_rn(a[k], cc);
        }
    }
    float s = 0.0f;
#pragma unroll
    for (int k = 0; k < 16; ++k) s += a[k];
    out[(size_t)blockIdx.x * blockDim.x + threadIdx.x] = s;
}


// Per-environment scenes: point i against the triangles of scene i (raw [N][S][4]); same arithmetic, the unit
// normals computed per use (no sharing between points).
__global__ void __launch_bounds__(kQueryThreads) points_inside_env_kernel(int64_t N, int S, const float2* __restrict__ pts,
                                                                         const float4* __restrict__ raw, uint8_t* __restrict__ flag) {
    const int64_t i = (int64_t)blockIdx.x * kQueryThreads + threadIdx.x;
    if (i >= N) return;
    const float2 p = pts[i];
    const float4* seg = raw + i * (int64_t)S;
    bool any = false;
    for (int t = 0; t + 2 < S; t += 3) {
        bool all = true;
#pragma unroll
        for (int e = 0; e < 3; ++e) {
            const float4 g = __ldg(seg + t + e);
            const float svx = g.z - g.x, svy = g.w - g.y;  // :21
            const float nx = svy, ny = -svx;               // :22
            const float nrm = sqrtf(dot2(nx, nx, ny, ny)); // :23
            const float pvx = p.x - g.x, pvy = p.y - g.y;  // :52
            const float d = dot2(pvx, nx / nrm, pvy, ny / nrm);  // :53
            all = all && (d < 0.0f);
        }
        any = any || all;
    }
    flag[i] = any ? 1 : 0;
}

}  // namespace doper
