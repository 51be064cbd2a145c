// scene.cuh: prepared-scene layout, scene_prepare_kernel (+ uniform grid), TMA / mbarrier staging — part of the doper_b200 CUDA translation unit (included by kernels.cu).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#include "sweep.cuh"

namespace doper {

// ------------------------------------------------------------------------------------------
// scene workspace layout
// ------------------------------------------------------------------------------------------
struct SceneHeader {  // 16 B per scene
    float scale;      // max |coordinate|
    int weird;        // degenerate / non-finite segment present -> exact sweep only
    int S;
    int S_pad;
};

struct SceneLayout {
    size_t hdr_off, q_off, rl_off, s1_off, grid_off, total;  // grid_off = 0: no grid (per-environment scenes)
    int S_pad;
};

__host__ __device__ inline SceneLayout scene_layout(int64_t n_scenes, int S) {
    SceneLayout L;
    L.S_pad = (S + 31) / 32 * 32;
    L.hdr_off = 0;
    size_t hdr = (size_t)n_scenes * sizeof(SceneHeader);
    L.q_off = (hdr + 255) / 256 * 256;
    L.rl_off = L.q_off + (size_t)n_scenes * L.S_pad * sizeof(float4);
    L.s1_off = (L.rl_off + (size_t)n_scenes * L.S_pad * sizeof(float) + 255) / 256 * 256;
    // raw end points s1 [n_scenes][S_pad] float2: read by the binary64 backward only (its segment direction is
    // s1 - s0 in binary64, as in oracle_grad_truth64, not the fp32-rounded sv of q); never staged
    L.total = L.s1_off + (size_t)n_scenes * L.S_pad * sizeof(float2);
    L.total = (L.total + 255) / 256 * 256;
    L.grid_off = 0;
    if (n_scenes == 1) {  // [GridHeader | cell_start (kGridMaxCells + 1) | items (4 S_pad)]
        L.grid_off = L.total;
        L.total += sizeof(GridHeader) + (size_t)(kGridMaxCells + 1) * sizeof(int) + (size_t)4 * L.S_pad * sizeof(int);
        L.total = (L.total + 255) / 256 * 256;
    }
    return L;
}

constexpr float kSentinel = 1.0e18f;  // far-away padding segment: never a candidate
constexpr int kGridMinSegments = 1024;  // scenes from this size on refresh candidate lists through the grid

__device__ __forceinline__ GridView grid_view(const unsigned char* ws, const SceneLayout& L) {
    GridView g;
    g.enabled = 0; g.cell_start = nullptr; g.items = nullptr; g.x0 = g.y0 = g.inv_h = 0.f; g.gx = g.gy = 1;
    // small scenes: scanning a few hundred staged segments beats the dependent global-memory loads of a grid
    // walk (measured at S = 201: 0.48 vs 0.29 ms on C2)
    if (L.grid_off && L.S_pad >= kGridMinSegments) {
        const GridHeader h = *reinterpret_cast<const GridHeader*>(ws + L.grid_off);
        g.cell_start = reinterpret_cast<const int*>(ws + L.grid_off + sizeof(GridHeader));
        g.items = g.cell_start + (kGridMaxCells + 1);
        g.x0 = h.x0; g.y0 = h.y0; g.inv_h = h.inv_h; g.gx = h.gx; g.gy = h.gy; g.enabled = h.enabled;
    }
    return g;
}

__global__ void __launch_bounds__(128) scene_prepare_kernel(int64_t n_scenes, int S,
                                                            const float4* __restrict__ segs,
                                                            unsigned char* __restrict__ ws) {
    const SceneLayout L = scene_layout(n_scenes, S);
    const int64_t sc = blockIdx.x;
    const float4* in = segs + sc * (int64_t)S;
    float4* q = reinterpret_cast<float4*>(ws + L.q_off) + sc * (int64_t)L.S_pad;
    float* rl = reinterpret_cast<float*>(ws + L.rl_off) + sc * (int64_t)L.S_pad;
    float2* e1 = reinterpret_cast<float2*>(ws + L.s1_off) + sc * (int64_t)L.S_pad;
    float mx = 0.0f;
    int weird = 0;
    for (int i = threadIdx.x; i < L.S_pad; i += blockDim.x) {
        if (i < S) {
            float4 s = in[i];
            float svx = s.z - s.x, svy = s.w - s.y;  // jax_geometry.py:132
            float len2 = svx * svx + svy * svy;      // :133
            q[i] = make_float4(s.x, s.y, svx, svy);
            rl[i] = 1.0f / len2;
            e1[i] = make_float2(s.z, s.w);
            float m = fmaxf(fmaxf(fabsf(s.x), fabsf(s.y)), fmaxf(fabsf(s.z), fabsf(s.w)));
            mx = fmaxf(mx, m);
            // the filtered sweep's error analysis needs a normal, finite, non-degenerate segment
            bool ok = (len2 >= 1.0e-30f) && (len2 <= 1.0e30f) && (m <= 1.0e15f);
            weird |= !ok;  // also true for NaN coordinates (comparisons false)
        } else {
            q[i] = make_float4(kSentinel, kSentinel, 0.0f, 0.0f);
            rl[i] = 0.0f;
            e1[i] = make_float2(kSentinel, kSentinel);
        }
    }
    __shared__ float s_mx[128];
    __shared__ int s_w[128];
    s_mx[threadIdx.x] = mx;
    s_w[threadIdx.x] = weird;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)blockDim.x; ++i) { mx = fmaxf(mx, s_mx[i]); weird |= s_w[i]; }
        SceneHeader h;
        h.scale = mx; h.weird = weird; h.S = S; h.S_pad = L.S_pad;
        reinterpret_cast<SceneHeader*>(ws + L.hdr_off)[sc] = h;
        s_w[0] = weird;
    }
    if (L.grid_off == 0) return;
    // ---- uniform grid over the (single, shared) scene ------------------------------------------------
    __syncthreads();
    weird = s_w[0];
    __shared__ float s_red[4][128];
    __shared__ int s_cnt[kGridMaxCells];
    __shared__ int s_cur[kGridMaxCells];
    __shared__ GridHeader s_g;
    float lox = 3.0e38f, loy = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f, ext = 0.0f;
    for (int i = threadIdx.x; i < S; i += blockDim.x) {
        const float4 s = in[i];
        lox = fminf(lox, fminf(s.x, s.z)); hix = fmaxf(hix, fmaxf(s.x, s.z));
        loy = fminf(loy, fminf(s.y, s.w)); hiy = fmaxf(hiy, fmaxf(s.y, s.w));
        ext = fmaxf(ext, fmaxf(fabsf(s.z - s.x), fabsf(s.w - s.y)));
    }
    s_red[0][threadIdx.x] = lox; s_red[1][threadIdx.x] = loy; s_red[2][threadIdx.x] = -hix; s_red[3][threadIdx.x] = -hiy;
    s_mx[threadIdx.x] = ext;
    __syncthreads();
    if (threadIdx.x == 0) {
        for (int i = 1; i < (int)blockDim.x; ++i) {
            for (int k = 0; k < 4; ++k) s_red[k][0] = fminf(s_red[k][0], s_red[k][i]);
            s_mx[0] = fmaxf(s_mx[0], s_mx[i]);
        }
        GridHeader g;
        g.x0 = s_red[0][0]; g.y0 = s_red[1][0];
        const float wx = -s_red[2][0] - g.x0, wy = -s_red[3][0] - g.y0;
        float h = s_mx[0] * 1.0001f;
        g.enabled = (!weird && h > 0.0f && wx >= 0.0f && wy >= 0.0f && wx < 1.0e30f && wy < 1.0e30f) ? 1 : 0;
        g.gx = g.gy = 1;
        if (g.enabled) {
            for (int it = 0; it < 8; ++it) {  // coarsen until the cell count fits
                g.gx = (int)fminf(floorf(wx / h) + 1.0f, 1.0e6f); g.gy = (int)fminf(floorf(wy / h) + 1.0f, 1.0e6f);
                if ((long long)g.gx * g.gy <= kGridMaxCells) break;
                h *= 1.01f * sqrtf((float)g.gx * (float)g.gy / (float)kGridMaxCells);
            }
            if ((long long)g.gx * g.gy > kGridMaxCells || g.gx * g.gy < 9) g.enabled = 0;  // too coarse to help
        }
        g.inv_h = 1.0f / h;
        g.n_items = 0; g.pad = 0;
        s_g = g;
    }
    __syncthreads();
    const GridHeader g = s_g;
    GridHeader* gh = reinterpret_cast<GridHeader*>(ws + L.grid_off);
    int* cell_start = reinterpret_cast<int*>(ws + L.grid_off + sizeof(GridHeader));
    int* items = cell_start + (kGridMaxCells + 1);
    if (!g.enabled) {
        if (threadIdx.x == 0) *gh = g;
        return;
    }
    const int ncell = g.gx * g.gy;
    for (int i = threadIdx.x; i < ncell; i += blockDim.x) s_cnt[i] = 0;
    __syncthreads();
    for (int pass = 0; pass < 2; ++pass) {
        for (int i = threadIdx.x; i < S; i += blockDim.x) {
            const float4 s = in[i];
            const int ax = grid_cell_coord(fminf(s.x, s.z), g.x0, g.inv_h, g.gx), bx = grid_cell_coord(fmaxf(s.x, s.z), g.x0, g.inv_h, g.gx);
            const int ay = grid_cell_coord(fminf(s.y, s.w), g.y0, g.inv_h, g.gy), by = grid_cell_coord(fmaxf(s.y, s.w), g.y0, g.inv_h, g.gy);
            for (int cy = ay; cy <= by; ++cy)
                for (int cx = ax; cx <= bx; ++cx) {
                    const int cell = cy * g.gx + cx;
                    if (pass == 0) atomicAdd(&s_cnt[cell], 1);
                    else {
                        const int pos = atomicAdd(&s_cur[cell], 1);
                        if (pos < 4 * L.S_pad) items[pos] = i;
                    }
                }
        }
        __syncthreads();
        if (pass == 0) {
            if (threadIdx.x == 0) {
                int acc = 0;
                for (int cidx = 0; cidx < ncell; ++cidx) { s_cur[cidx] = acc; cell_start[cidx] = acc; acc += s_cnt[cidx]; }
                cell_start[ncell] = acc;
                GridHeader out = g;
                out.n_items = acc;
                if (acc > 4 * L.S_pad) out.enabled = 0;  // cannot happen (<= 2 x 2 cells per segment); stay safe
                *gh = out;
            }
            __syncthreads();
        }
    }
}

// ------------------------------------------------------------------------------------------
// TMA (1-D bulk copy) + mbarrier primitives
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void tma_load_1d(void* dst_smem, const void* src_gmem, uint32_t bytes,
                                            uint64_t* bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
            smem_u32(dst_smem)),
        "l"(__cvta_generic_to_global(src_gmem)), "r"(bytes), "r"(smem_u32(bar))
        : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    uint32_t done;
    do {
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(smem_u32(bar)), "r"(parity)
            : "memory");
    } while (!done);
}

constexpr uint32_t kTmaChunk = 32768;  // bytes per bulk copy (multiple of 16)

// One elected thread stages `n_local` consecutive prepared scenes into shared memory.
__device__ __forceinline__ void stage_scenes(float4* sq, float* srl, const unsigned char* ws,
                                             const SceneLayout& L, int64_t first_scene, int n_local,
                                             uint64_t* bar) {
    const unsigned char* gq = ws + L.q_off + (size_t)first_scene * L.S_pad * sizeof(float4);
    const unsigned char* gr = ws + L.rl_off + (size_t)first_scene * L.S_pad * sizeof(float);
    const uint32_t qb = (uint32_t)n_local * L.S_pad * sizeof(float4);
    const uint32_t rb = (uint32_t)n_local * L.S_pad * sizeof(float);
    mbar_expect_tx(bar, qb + rb);
    for (uint32_t o = 0; o < qb; o += kTmaChunk)
        tma_load_1d(reinterpret_cast<unsigned char*>(sq) + o, gq + o, min(kTmaChunk, qb - o), bar);
    for (uint32_t o = 0; o < rb; o += kTmaChunk)
        tma_load_1d(reinterpret_cast<unsigned char*>(srl) + o, gr + o, min(kTmaChunk, rb - o), bar);
}


}  // namespace doper
