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
    size_t hdr_off, q_off, rl_off, s1_off, grid_off, bake_off, total;  // grid_off = 0: none (per-environment scenes)
    size_t bake_stride;  // bytes between the baked areas of consecutive scenes
    int bake_cap;  // item capacity of the baked cell lists
    int S_pad;
};

// ------------------------------------------------------------------------------------------
// baked broad phase: static per-cell candidate lists of a shared scene
// ------------------------------------------------------------------------------------------
// A uniform grid over the scene's bounding box (+ margin); cell k holds the hit list (sweep.cuh, "hit
// lists") of the disc that circumscribes it: every segment that a ball of radius <= r_max can touch from
// ANY point of the cell.  A rollout then never scans the scene: the list of the cell a point falls in is
// valid for that point by the same proof as a per-ball list built around the cell centre with
// rho = half diagonal / 0.98, and an empty cell proves "no collision" with no arithmetic at all.  Built
// once per (scene, r_max) by scene_bake_kernel; read by rollout_fwd_baked_kernel.
constexpr int kBakeMaxCells = 4096;
constexpr int kBakeItemsPerSeg = 32;
// Second level (round 2): every cell is cut into 2^s x 2^s sub-cells (s = fine_shift <= 3) and keeps ONE BIT per
// sub-cell — "the hit list of this sub-cell's circumscribed disc is empty / is not".  A cell of the seeded maps is
// ~5 ball radii wide, so 14 % of the points a rollout visits fall in a cell with a list and practically every warp
// step walks one (some lane of 32 does); at an eighth of the cell size 1.8 % of the points do, and the list of the
// (coarse) cell is only read for those.  Same proof as for the cells themselves with rho_f = sub-cell half diagonal
// / 0.98: a clear bit proves "no collision" for every point that falls in the sub-cell.
constexpr int kBakeFineShiftMax = 3;
// byte offsets of the arrays behind the BakeHeader: [cell_start | items | sub-cell masks]; the masks come last so that
// a kernel that does not use them (the pipelined forward) stages bake_lists_bytes() only
constexpr size_t kBakeItemsOff = (size_t)(kBakeMaxCells + 16) * sizeof(uint32_t);
__host__ __device__ inline size_t bake_lists_bytes(int S_pad) {  // cell_start + items (S_pad is a multiple of 32: 16-byte aligned)
    return kBakeItemsOff + (size_t)kBakeItemsPerSeg * S_pad * sizeof(unsigned short);
}
__host__ __device__ inline size_t bake_fine_off(int S_pad) { return bake_lists_bytes(S_pad); }

struct BakeHeader {  // 64 B
    float x0, y0, h, inv_h;   // cell (ix, iy) = [x0 + ix h, x0 + (ix + 1) h) x ...
    int nx, ny, n_items, enabled;
    float r_max;   // lists are valid for ball radii <= r_max
    float M;       // bound on |coordinate| of the scene and of every point inside the grid (+ rho)
    float far2;    // cull_far2(r_max, M): estimates^2 above it prove "no collision"
    float rho;     // radius of the disc a cell list was built for
    float fnx, fny;  // (float)nx, (float)ny
    int fine_shift;  // sub-cells per cell edge = 1 << fine_shift (0: no second level, every mask is all ones)
    int pad;
};

__host__ __device__ inline size_t bake_bytes(int S_pad) {
    size_t b = sizeof(BakeHeader) + bake_fine_off(S_pad) + (size_t)kBakeMaxCells * sizeof(unsigned long long);
    return (b + 15) / 16 * 16;
}

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
    // per scene: [BakeHeader 64 B | cell_start u32 (kBakeMaxCells + 16) | items u16 (bake_cap) | sub-cell masks u64 (kBakeMaxCells)]
    L.bake_off = L.total;
    L.bake_cap = kBakeItemsPerSeg * L.S_pad;
    L.bake_stride = (bake_bytes(L.S_pad) + 255) / 256 * 256;
    L.total += (size_t)n_scenes * L.bake_stride;
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
            float len2 = dot2(svx, svx, svy, svy);    // :133
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

// One block per scene.  Segment-driven: each segment visits the cells whose centre can be within `reach` of it
// (its bounding box grown by reach, plus one cell of slack) and applies the list test of sweep.cuh
// (estimate at the cell centre <= reach, on squares with upward slack); counting sort into CSR; the items
// of a cell are then sorted ascending so that the layout is deterministic.  If the lists overflow the
// item capacity the cell size grows by 1.5x and the pass repeats (a single cell always fits).
__global__ void __launch_bounds__(256) scene_bake_kernel(int64_t n_scenes, int S, float r_max, unsigned char* __restrict__ ws) {
    const SceneLayout L = scene_layout(n_scenes, S);
    const int64_t scn = blockIdx.x;  // one block per scene
    const SceneHeader sh = reinterpret_cast<const SceneHeader*>(ws + L.hdr_off)[scn];
    const float4* q = reinterpret_cast<const float4*>(ws + L.q_off) + scn * L.S_pad;
    const float* rl = reinterpret_cast<const float*>(ws + L.rl_off) + scn * L.S_pad;
    const float2* e1 = reinterpret_cast<const float2*>(ws + L.s1_off) + scn * L.S_pad;
    unsigned char* bake = ws + L.bake_off + (size_t)scn * L.bake_stride;
    BakeHeader* bh = reinterpret_cast<BakeHeader*>(bake);
    uint32_t* cell_start = reinterpret_cast<uint32_t*>(bake + sizeof(BakeHeader));
    unsigned long long* fine = reinterpret_cast<unsigned long long*>(bake + sizeof(BakeHeader) + bake_fine_off(L.S_pad));
    unsigned short* items = reinterpret_cast<unsigned short*>(bake + sizeof(BakeHeader) + kBakeItemsOff);
    __shared__ float s_red[4][256];
    __shared__ int s_cnt[kBakeMaxCells];
    __shared__ int s_tot;
    __shared__ BakeHeader s_b;
    const int tid = threadIdx.x;
    float lox = 3.0e38f, loy = 3.0e38f, hix = -3.0e38f, hiy = -3.0e38f;
    for (int i = tid; i < S; i += blockDim.x) {
        const float4 a = q[i];
        const float2 b = e1[i];
        lox = fminf(lox, fminf(a.x, b.x)); hix = fmaxf(hix, fmaxf(a.x, b.x));
        loy = fminf(loy, fminf(a.y, b.y)); hiy = fmaxf(hiy, fmaxf(a.y, b.y));
    }
    s_red[0][tid] = lox; s_red[1][tid] = loy; s_red[2][tid] = -hix; s_red[3][tid] = -hiy;
    __syncthreads();
    if (tid == 0) {
        for (int i = 1; i < (int)blockDim.x; ++i)
            for (int k = 0; k < 4; ++k) s_red[k][0] = fminf(s_red[k][0], s_red[k][i]);
        BakeHeader b;
        b.enabled = 0; b.nx = b.ny = 1; b.n_items = 0; b.r_max = r_max; b.x0 = b.y0 = 0.f; b.h = b.inv_h = 1.f;
        b.M = sh.scale; b.far2 = 0.f; b.rho = 0.f; b.fnx = b.fny = 1.f; b.fine_shift = 0; b.pad = 0;
        const float W = -s_red[2][0] - s_red[0][0], H = -s_red[3][0] - s_red[1][0];
        const bool ok = !sh.weird && r_max > 0.0f && r_max <= 1.0e15f && W >= 0.0f && H >= 0.0f && W <= 1.0e15f && H <= 1.0e15f &&
                        S <= 65535;
        if (ok) {
            b.enabled = 1;
            // first guess: kBakeMaxCells square cells over the box + margins, never finer than 2^-10 of the
            // coordinate scale (the cell lookup's rounding must stay far inside the 2 % slack of rho)
            const float Wm = W + 2.0f * r_max, Hm = H + 2.0f * r_max;
            b.h = fmaxf(sqrtf((Wm * Hm) / (float)kBakeMaxCells) * 1.05f, fmaxf(sh.scale, r_max) * 0.0009765625f);
            b.x0 = s_red[0][0]; b.y0 = s_red[1][0];  // box corner; the margin is applied per attempt (depends on h)
        }
        s_b = b;
        s_red[0][1] = W; s_red[1][1] = H;
    }
    __syncthreads();
    if (!s_b.enabled) {
        if (tid == 0) *bh = s_b;
        return;
    }
    const float bx0 = s_b.x0, by0 = s_b.y0, W = s_red[0][1], H = s_red[1][1];
    for (int attempt = 0; attempt < 64; ++attempt) {
        __syncthreads();
        if (tid == 0) {
            BakeHeader b = s_b;
            if (attempt > 0) b.h *= 1.5f;
            // grid rectangle = box grown by mg = r_max + h on every side: a point outside the rectangle is
            // further than r_max + h - (rounding) from every segment
            const float mg = r_max + b.h;
            b.nx = (int)fminf(ceilf((W + 2.0f * mg) / b.h) + 1.0f, 1.0e6f);
            b.ny = (int)fminf(ceilf((H + 2.0f * mg) / b.h) + 1.0f, 1.0e6f);
            while ((long long)b.nx * b.ny > kBakeMaxCells) {
                b.h *= 1.25f;
                const float mg2 = r_max + b.h;
                b.nx = (int)fminf(ceilf((W + 2.0f * mg2) / b.h) + 1.0f, 1.0e6f);
                b.ny = (int)fminf(ceilf((H + 2.0f * mg2) / b.h) + 1.0f, 1.0e6f);
            }
            const float mgf = r_max + b.h;
            b.x0 = bx0 - mgf; b.y0 = by0 - mgf;
            b.inv_h = 1.0f / b.h;
            b.fnx = (float)b.nx; b.fny = (float)b.ny;
            b.rho = (b.h * 0.70710678f) * (1.0f / kCullRhoCheck) * 1.0001f;  // half diagonal / 0.98, rounded up
            const float ext = fmaxf(fmaxf(fabsf(b.x0), fabsf(b.x0 + b.h * b.fnx)), fmaxf(fabsf(b.y0), fabsf(b.y0 + b.h * b.fny)));
            b.M = fmaxf(sh.scale, ext + b.rho) * 1.0001f;
            b.far2 = cull_far2(r_max, b.M);
            s_b = b;
            s_tot = 0;
        }
        __syncthreads();
        const BakeHeader b = s_b;
        const int ncell = b.nx * b.ny;
        for (int i = tid; i < ncell; i += blockDim.x) s_cnt[i] = 0;
        __syncthreads();
        const float reach = __fmaf_rn(kCullSafePerM, b.M, b.rho + r_max) * kCullUp;  // list_reach(), hit list
        const float T = (reach * reach) * kCullUp;
        const float R = reach * 1.01f + b.h;  // a cell centre within reach (+ 49 u M) of the segment lies in its box grown by R
        for (int pass = 0; pass < 2; ++pass) {
            for (int i = tid; i < S; i += blockDim.x) {
                const float4 a = q[i];
                const float r1 = rl[i];
                const float2 e = e1[i];
                const int ax = max(0, (int)floorf((fminf(a.x, e.x) - R - b.x0) * b.inv_h) - 1);
                const int cx1 = min(b.nx - 1, (int)floorf((fmaxf(a.x, e.x) + R - b.x0) * b.inv_h) + 1);
                const int ay = max(0, (int)floorf((fminf(a.y, e.y) - R - b.y0) * b.inv_h) - 1);
                const int cy1 = min(b.ny - 1, (int)floorf((fmaxf(a.y, e.y) + R - b.y0) * b.inv_h) + 1);
                for (int cy = ay; cy <= cy1; ++cy)
                    for (int cx = ax; cx <= cx1; ++cx) {
                        const float ccx = __fmaf_rn((float)cx + 0.5f, b.h, b.x0), ccy = __fmaf_rn((float)cy + 0.5f, b.h, b.y0);
                        if (!(approx_d2(ccx, ccy, a, r1) <= T)) continue;
                        const int cell = cy * b.nx + cx;
                        if (pass == 0) {
                            atomicAdd(&s_cnt[cell], 1);
                        } else {
                            const int pos = atomicAdd(&s_cnt[cell], 1);  // cursor (re-initialised to the cell's start)
                            items[pos] = (unsigned short)i;
                        }
                    }
            }
            __syncthreads();
            if (pass == 0) {
                if (tid == 0) {
                    int acc = 0;
                    for (int c = 0; c < ncell; ++c) { const int n = s_cnt[c]; s_cnt[c] = acc; cell_start[c] = (uint32_t)acc; acc += n; }
                    for (int c = ncell; c < kBakeMaxCells + 16; ++c) cell_start[c] = (uint32_t)acc;
                    s_tot = acc;
                }
                __syncthreads();
                if (s_tot > L.bake_cap) break;  // too many items for this cell size: coarser grid
            }
        }
        __syncthreads();
        if (s_tot <= L.bake_cap) {
            // deterministic layout: ascending segment index inside every cell
            for (int c = tid; c < ncell; c += blockDim.x) {
                const int k0 = (int)cell_start[c], k1 = (int)cell_start[c + 1];
                for (int k = k0 + 1; k < k1; ++k) {
                    const unsigned short v = items[k];
                    int m = k - 1;
                    while (m >= k0 && items[m] > v) { items[m + 1] = items[m]; --m; }
                    items[m + 1] = v;
                }
            }
            // second level: one bit per sub-cell.  Sub-cells no finer than 2^-10 of the coordinate scale (as for the
            // cells: the look-up's rounding stays far inside the 2 % slack of rho_f).  A segment on the hit list of a
            // sub-cell's disc is on the list of its cell: centre to centre is at most hd - hd_f (half diagonals), so
            // its estimate at the cell centre is below r_max + rho_f + (hd - hd_f) + safe M + 2 err, and
            // rho - rho_f - (hd - hd_f) = 0.0204 (hd - hd_f) >= 1e-5 M is far above the estimate's error (err ~
            // 40 ulp of M).  Only the cell's own list is therefore tested, sub-cell by sub-cell.
            int fs = kBakeFineShiftMax;
            while (fs > 0 && b.h / (float)(1 << fs) < fmaxf(sh.scale, r_max) * 0.0009765625f) --fs;
            const int nsub = 1 << fs;
            const float hf = b.h / (float)nsub;  // exact: a power of two
            const float rho_f = (hf * 0.70710678f) * (1.0f / kCullRhoCheck) * 1.0001f;
            const float reach_f = __fmaf_rn(kCullSafePerM, b.M, rho_f + r_max) * kCullUp;
            const float T_f = (reach_f * reach_f) * kCullUp;
            for (int c = tid; c < kBakeMaxCells; c += blockDim.x) {
                unsigned long long mask = 0ull;
                if (fs == 0) {
                    mask = ~0ull;
                } else if (c < ncell) {
                    const int k0 = (int)cell_start[c], k1 = (int)cell_start[c + 1];
                    const int cx = c % b.nx, cy = c / b.nx;
                    for (int k = k0; k < k1; ++k) {
                        const int i = items[k];
                        const float4 a = q[i];
                        const float r1 = rl[i];
                        for (int sy = 0; sy < nsub; ++sy)
                            for (int sx = 0; sx < nsub; ++sx) {
                                const float ccx = __fmaf_rn((float)(cx * nsub + sx) + 0.5f, hf, b.x0);
                                const float ccy = __fmaf_rn((float)(cy * nsub + sy) + 0.5f, hf, b.y0);
                                if (approx_d2(ccx, ccy, a, r1) <= T_f) mask |= 1ull << ((sy << fs) | sx);
                            }
                    }
                }
                fine[c] = mask;
            }
            if (tid == 0) { BakeHeader out = s_b; out.n_items = s_tot; out.fine_shift = fs; *bh = out; }
            return;
        }
    }
    if (tid == 0) { BakeHeader out = s_b; out.enabled = 0; *bh = out; }  // unreachable in practice
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
