// doper_b200 CUDA kernels (sm_100a) and their C-ABI launchers (include/doper_b200.h).
//
// Compile: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -lineinfo -fmad=false
//          -prec-div=true -prec-sqrt=true -ftz=false  (see doper_b200/build.py)
//
// Kernels
//   scene_prepare_kernel   raw segments -> staging layout (one block per scene)
//   rollout_fwd_kernel<G>  whole rollout per launch: state in registers, scene staged into shared
//                          memory by 1-D TMA bulk copies (cp.async.bulk + mbarrier), G lanes per
//                          ball on the sweep, warp-shuffle lexicographic min, K-step checkpoints +
//                          per-step hit log for the backward
//   rollout_bwd_kernel     reverse pass over checkpoints: replay a K-block from the hit log (no
//                          sweep: argmin has no gradient), stash states in shared memory, apply the
//                          hand-derived step adjoint
//   l1_loss_kernel         per-ball L1 loss, its seed cotangent and a deterministic total
//   closest_segment_kernel / points_inside_kernel   the scene-sampler queries
//   fp32_probe_kernel      register-only FP32 issue-rate probe for the roofline denominator
#include <cuda_runtime.h>
#include <stdint.h>

#include "../../include/doper_b200.h"
#include "physics.cuh"
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
    size_t hdr_off, q_off, rl_off, grid_off, total;  // grid_off = 0: no grid (per-environment scenes)
    int S_pad;
};

__host__ __device__ inline SceneLayout scene_layout(int64_t n_scenes, int S) {
    SceneLayout L;
    L.S_pad = (S + 31) / 32 * 32;
    L.hdr_off = 0;
    size_t hdr = (size_t)n_scenes * sizeof(SceneHeader);
    L.q_off = (hdr + 255) / 256 * 256;
    L.rl_off = L.q_off + (size_t)n_scenes * L.S_pad * sizeof(float4);
    L.total = L.rl_off + (size_t)n_scenes * L.S_pad * sizeof(float);
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
    float mx = 0.0f;
    int weird = 0;
    for (int i = threadIdx.x; i < L.S_pad; i += blockDim.x) {
        if (i < S) {
            float4 s = in[i];
            float svx = s.z - s.x, svy = s.w - s.y;  // jax_geometry.py:132
            float len2 = svx * svx + svy * svy;      // :133
            q[i] = make_float4(s.x, s.y, svx, svy);
            rl[i] = 1.0f / len2;
            float m = fmaxf(fmaxf(fabsf(s.x), fabsf(s.y)), fmaxf(fabsf(s.z), fabsf(s.w)));
            mx = fmaxf(mx, m);
            // the filtered sweep's error analysis needs a normal, finite, non-degenerate segment
            bool ok = (len2 >= 1.0e-30f) && (len2 <= 1.0e30f) && (m <= 1.0e15f);
            weird |= !ok;  // also true for NaN coordinates (comparisons false)
        } else {
            q[i] = make_float4(kSentinel, kSentinel, 0.0f, 0.0f);
            rl[i] = 0.0f;
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

// ------------------------------------------------------------------------------------------
// forward rollout
// ------------------------------------------------------------------------------------------
constexpr int kFwdThreads = 128;

struct FwdParams {
    int64_t B;
    int S, per_env, n_steps, K, force_exact;
    int exact_idx;  // log the exact closest index at every step (default: at collisions only)
    int stage;      // 1: scene staged in shared memory (TMA); 0: read from global memory / L2 (scene too large)
    Consts c;
    const unsigned char* scene_ws;
    const float2* x0;
    const float2* v0;
    float2* xT;
    float2* vT;
    float2* traj_x;
    float2* traj_v;
    float2* traj_a;
    float4* ckpt;     // [n_ckpt][B]
    float4* hitstate; // [n_steps][B], written at collided steps only: the state at the START of the step
    int32_t* hitlog;  // element (t, ball) at t * hl_st + ball * hl_sb
    int64_t hl_st, hl_sb;
};

template <int G>
__global__ void __launch_bounds__(kFwdThreads) rollout_fwd_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = kFwdThreads / G;  // balls per block
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        int64_t first_scene = p.per_env ? first_ball : 0;
        int n = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, first_scene, n, bar);
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the shuffles
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    sc.q = sq + (size_t)local_scene * L.S_pad;
    sc.rl = srl + (size_t)local_scene * L.S_pad;
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;

    mbar_wait(bar, 0);

    int prev_idx = -1;  // no hint at the first step
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int until_ckpt = 0;
    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += p.B; }
        }
        --until_ckpt;
        State s1;
        free_step(x, y, vx, vy, c, s1);
        const uint32_t clipb = clip_bits(vx, vy);
        int idx; float dist;
        sweep_any<G>(sc, s1.x, s1.y, lane, prev_idx, idx, dist);
        prev_idx = idx;
        const bool hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
        if (hit && writer && p.hitstate) p.hitstate[(size_t)t * p.B + ball] = make_float4(x, y, vx, vy);
        if (hit) s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c);
        x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        if (writer) {
            if (hl) { *hl = pack_log(idx, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += p.B; }
            if (tv) { *tv = make_float2(vx, vy); tv += p.B; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += p.B; }
        }
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Broad-phase forward: same organisation as rollout_fwd_kernel<G> (G lanes per ball, scene staged
// in shared memory by TMA, sequential steps), but each ball carries a CANDIDATE LIST (sweep.cuh,
// "broad phase"): the segments that can be the closest one anywhere within rho of the list's
// centre, rho sized from the ball's current speed for about kCullPeriod steps.  Per step the sweep
// touches the list only (~5-30 segments instead of S); a displacement guard triggers the refresh,
// warp-uniformly (one lane over its rho -> every ball of the warp refreshes: no divergence on the
// O(S) scan).  Results are bit-identical to the full sweep (the list provably contains the argmin
// and all its ties).  Degenerate scenes / non-finite points take the exact path as everywhere else.
constexpr int kCullPeriod = 24;  // steps a list is sized for (performance only)

template <int G>
struct CullCfg {
    static constexpr int LCAP = G == 1 ? 48 : (G == 2 ? 32 : (G == 4 ? 16 : 8));  // list entries per lane
};

// STAGED = false: the scene does not fit shared memory and is read in place (global memory / L2).
// One thread per ball (G = 1) is compiled for 7 blocks per SM (72 registers): 131,072 balls, one
// GPU's share of BASELINE configs[4], are then co-resident in ONE wave (with 96 registers the grid
// needed two, the second one 38 % full).
// GRID: dense scenes (>= 1,024 segments) refresh their lists through the scene's uniform grid.
template <int G, bool STAGED, bool GRID>
__global__ void __launch_bounds__(kFwdThreads, (G == 1 && STAGED && !GRID) ? 7 : 1) rollout_fwd_cull_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = kFwdThreads / G;  // balls per block
    constexpr int LCAP = CullCfg<G>::LCAP;
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;
    constexpr bool staged = STAGED;  // else the scene is read in place (global memory / L2): any size works,
                                     // the per-step work touches the short lists only
    const size_t scene_bytes = staged ? (size_t)n_local * L.S_pad * 20 + 16 : 0;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    unsigned short* list = reinterpret_cast<unsigned short*>(smem + scene_bytes) + tid;

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (staged) {
        if (tid == 0) mbar_init(bar, 1);
        __syncthreads();
        if (tid == 0) {
            int64_t first_scene = p.per_env ? first_ball : 0;
            int n = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
            stage_scenes(sq, srl, p.scene_ws, L, first_scene, n, bar);
        }
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the shuffles
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    if (staged) {
        sc.q = sq + (size_t)local_scene * L.S_pad;
        sc.rl = srl + (size_t)local_scene * L.S_pad;
    } else {
        const size_t sid = p.per_env ? (size_t)ball : 0;
        sc.q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + sid * L.S_pad;
        sc.rl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + sid * L.S_pad;
    }
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;

    if (staged) mbar_wait(bar, 0);

    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int until_ckpt = 0;
    // seed of the first list: smallest estimate at the start position (any index would be valid)
    int j0 = 0;
    if (__all_sync(0xffffffffu, !sc.weird)) j0 = hint_by_estimate<G>(sc, x, y, lane);
    CullState cs;
    cs.cx = 0.f; cs.cy = 0.f; cs.rho2 = -1.0f; cs.cnt = 0; cs.far2 = 0.f;  // rho2 < 0: no list yet
    const GridView grid = grid_view(p.scene_ws, L);
    const float rho_min = sc.scale * 0.001953125f, rho_max = sc.scale * 0.125f;
    const bool exact_idx = p.exact_idx != 0;

    State s1;  // free-flight candidate of the current step
    free_step(x, y, vx, vy, c, s1);
    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += p.B; }
        }
        --until_ckpt;
        // speculate: the candidate of the NEXT step, assuming this one does not collide (independent of
        // the sweep below, so its ~100-cycle dependency chain overlaps the sweep's)
        State s2;
        const bool ok2 = free_step_fast(s1.x, s1.y, s1.vx, s1.vy, c, s2);

        const float px = s1.x, py = s1.y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        const bool fast = !sc.weird && M <= 1.0e15f;  // false for NaN / inf points
        const float ddx = px - cs.cx, ddy = py - cs.cy;
        const bool stale = fast && !(ddx * ddx + ddy * ddy <= cs.rho2);
        if (__any_sync(0xffffffffu, stale)) {  // warp-uniform refresh
            j0 = list_argmin<G, LCAP>(sc, fast ? px : 0.f, fast ? py : 0.f, lane, list, kFwdThreads, cs, j0);  // cnt = 0 before the first list
            if (fast) {
                const float sx = px - x, sy = py - y;
                const float dstep = sqrtf(sx * sx + sy * sy);  // displacement of this step
                const float rho = fminf(fmaxf(1.25f * (float)kCullPeriod * dstep, rho_min), rho_max);
                cull_refresh<G, LCAP, GRID>(sc, grid, px, py, rho, c.r, lane, j0, list, kFwdThreads, cs);
            }
        }
        // narrow phase, estimates only: far from every wall -> no collision, nothing exact needed
        float amin = fast ? lane_list_min<G, LCAP>(sc, px, py, lane, list, kFwdThreads, cs) : 0.0f;
        amin = group_min_f<G>(amin);
        const bool near = !fast || exact_idx || !(amin > cs.far2);
        bool hit = false;
        int idx = 0;
        if (__any_sync(0xffffffffu, near)) {
            unsigned long long key = kNoKey;
            if (near) {
                if (fast) key = lane_culled<G, LCAP>(sc, px, py, lane, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M),
                                                     list, kFwdThreads, cs.cnt);
                else key = lane_exact<G>(sc, px, py, lane);
            }
            key = group_min_redux<G>(key);
            if (__any_sync(0xffffffffu, near && key == kNoKey)) {  // cannot happen (the list holds the argmin); stay safe
                const unsigned long long k2 = group_min_redux<G>(lane_exact<G>(sc, px, py, lane));
                if (key == kNoKey) key = k2;
            }
            if (near) {
                float dist;
                unpack_key(key, idx, dist);
                hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
            }
        }
        const uint32_t clipb = clip_bits(vx, vy);
        if (hit) {
            if (writer && p.hitstate) p.hitstate[(size_t)t * p.B + ball] = make_float4(x, y, vx, vy);
            s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c);
            x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        } else {
            x = px; y = py; vx = s1.vx; vy = s1.vy;
        }
        if (writer) {
            if (hl) { *hl = pack_log(idx, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += p.B; }
            if (tv) { *tv = make_float2(vx, vy); tv += p.B; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += p.B; }
        }
        if (hit || !ok2) free_step(x, y, vx, vy, c, s1);  // the speculation does not apply: redo from the real state
        else s1 = s2;
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Time-parallel broad-phase forward: G lanes per ball, LANE = TIME STEP, one candidate list per
// ball.  Per round a group (1) refreshes its list when the ball has drifted half a radius from the
// list's centre, (2) integrates up to G steps of free flight (lane k keeps step k), (3) every lane
// tests ITS step against the list: the estimate alone proves "no collision" for almost every step
// (cull_far2); the few near-wall steps get the exact distance, (4) a ballot finds the first
// collided (or out-of-list) step: the steps before it commit, the collision is resolved, the next
// round starts after it.  The per-step critical path is the free-flight chain (~100 cycles)
// instead of chain + sweep + reduction.  A ball that keeps colliding shrinks its speculation
// length to one step.  Results are bit-identical to every other forward kernel.
constexpr int kTpcListCap = 64;
constexpr int kContactSteps = 16;  // sequential steps per round of a ball in contact mode

template <int G, bool GRID>
__global__ void __launch_bounds__(kFwdThreads) rollout_fwd_tpc_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    constexpr int BPB = kFwdThreads / G;  // balls per block
    constexpr unsigned kLow = G == 32 ? 0xffffffffu : ((1u << G) - 1u);
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G, gbase = (tid & 31) - lane;
    const int64_t n_scenes = p.per_env ? p.B : 1;
    const SceneLayout L = scene_layout(n_scenes, p.S);
    const int n_local = p.per_env ? BPB : 1;

    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    unsigned short* lists = reinterpret_cast<unsigned short*>(smem + (size_t)n_local * L.S_pad * 20 + 16);

    const int64_t first_ball = (int64_t)blockIdx.x * BPB;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        int64_t first_scene = p.per_env ? first_ball : 0;
        int ns = p.per_env ? (int)min((int64_t)BPB, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, first_scene, ns, bar);
    }

    int64_t ball = first_ball + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // absent ball: reads the last ball's inputs, never advances, writes nothing
    const int local_scene = p.per_env ? (int)(ball - first_ball) : 0;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];

    SceneView sc;
    sc.q = sq + (size_t)local_scene * L.S_pad;
    sc.rl = srl + (size_t)local_scene * L.S_pad;
    sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;

    const float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;  // state at step t (uniform within the group)
    const Consts c = p.c;
    const int n = p.n_steps, K = p.K;
    const bool exact_idx = p.exact_idx != 0;
    mbar_wait(bar, 0);

    int t = active ? 0 : n;
    int T = G;       // speculation length of the next round
    int trouble = 0; // 1: refresh before the next round no matter what; 2: sweep every segment exactly next round
    int j0 = 0;
    if (__all_sync(0xffffffffu, !sc.weird)) j0 = hint_by_estimate<G>(sc, x, y, lane);
    GroupList gl;
    gl.e = lists + (size_t)slot * kTpcListCap; gl.cap = kTpcListCap; gl.cnt = 0;
    gl.cx = 0.f; gl.cy = 0.f; gl.rho2 = -1.0f; gl.far2 = 0.f;
    const GridView grid = grid_view(p.scene_ws, L);
    const float rho_min = sc.scale * 0.001953125f, rho_max = sc.scale * 0.125f;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;

    while (__any_sync(0xffffffffu, t < n)) {
        // ---- (0) contact mode --------------------------------------------------------------------------
        // A ball that collided at the first step of its last round (resting / sliding contact) gains
        // nothing from speculation: it runs up to kContactSteps plain sequential steps here, every lane
        // of the group redundantly (identical inputs -> identical values, so the group's state stays
        // uniform without any exchange; no warp-wide primitive inside), then sits this round out.
        bool sat_out = false;
        if (t < n && T == 1 && trouble == 0 && gl.rho2 >= 0.0f && !sc.weird) {
            int miss = 0;
            int cj = -1;  // segment of the last collision: almost always the next one too (sliding contact)
            for (int it = 0; it < kContactSteps && t < n; ++it) {
                State s1;
                if (!free_step_fast(x, y, vx, vy, c, s1)) free_step(x, y, vx, vy, c, s1);
                const float px = s1.x, py = s1.y;
                const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
                const float ddx = px - gl.cx, ddy = py - gl.cy;
                if (!(M <= 1.0e15f)) break;                                              // normal path: exact sweep
                if (!(ddx * ddx + ddy * ddy <= gl.rho2)) { trouble = 1; break; }          // normal path: refresh
                int idx = 0;
                bool hit = false, decided = false;
                if (cj >= 0 && gl.cnt <= gl.cap) {
                    // fast decision: the exact distance to segment cj, and a proof that no other list entry
                    // can tie or beat it (estimate^2 > (d + 128 u M)^2 => reference distance > d)
                    const float dj = cand_dist(px, py, sc.q[cj]);
                    const float lim0 = __fmaf_rn(kCullSafePerM, M, dj) * kCullUp;
                    const float lim = (lim0 * lim0) * kCullUp;
                    bool sole = dj == dj;
                    for (int k = 0; k < gl.cnt; ++k) {
                        const int i = gl.e[k];
                        const float a = approx_d2(px, py, sc.q[i], sc.rl[i]);
                        sole = sole && (i == cj || a > lim);
                    }
                    if (sole) { idx = cj; hit = dj <= c.r; decided = true; }
                }
                if (!decided) {
                    const float amin = glist_min(sc, px, py, gl);
                    if (exact_idx || !(amin > gl.far2)) {
                        unsigned long long key = glist_exact(sc, px, py, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M),
                                                             gl.e, gl.cap, gl.cnt);
                        if (key == kNoKey) key = lane_exact<1>(sc, px, py, 0);
                        float dist;
                        unpack_key(key, idx, dist);
                        hit = dist <= c.r;
                    }
                }
                const uint32_t clipb = clip_bits(vx, vy);
                if (active && lane == 0) {
                    if (p.ckpt && (t % K) == 0) p.ckpt[(size_t)(t / K) * p.B + ball] = make_float4(x, y, vx, vy);
                    if (hit && p.hitstate) p.hitstate[(size_t)t * p.B + ball] = make_float4(x, y, vx, vy);
                }
                if (hit) { s1 = resolve_inline(x, y, vx, vy, s1, sc.q[idx], c); cj = idx; }
                x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
                if (active && lane == 0) {
                    if (hl) hl[(int64_t)t * p.hl_st] = pack_log((hit || exact_idx) ? idx : 0, hit, clipb);
                    const size_t o = (size_t)t * p.B + ball;
                    if (p.traj_x) p.traj_x[o] = make_float2(x, y);
                    if (p.traj_v) p.traj_v[o] = make_float2(vx, vy);
                    if (p.traj_a) p.traj_a[o] = make_float2(s1.ax, s1.ay);
                }
                ++t;
                sat_out = true;
                if (hit) miss = 0;
                else if (++miss >= 2) { T = 2; break; }  // left the wall: speculate again
            }
        }
        const bool alive = t < n && !sat_out;
        // ---- (1) list refresh at the current position -------------------------------------------------
        {
            const float M0 = fmaxf(sc.scale, fmaxf(fabsf(x), fabsf(y)));
            const bool fast0 = !sc.weird && M0 <= 1.0e15f;
            const float ddx = x - gl.cx, ddy = y - gl.cy;
            const bool want = alive && fast0 && (trouble == 1 || !(ddx * ddx + ddy * ddy <= 0.25f * gl.rho2));
            if (__any_sync(0xffffffffu, want)) {
                if (want) j0 = glist_argmin(sc, x, y, gl, j0);
                const float dstep = sqrtf(vx * vx + vy * vy) * c.dt;
                const float rho = fminf(fmaxf(1.25f * (float)(kCullPeriod + 2 * G) * dstep, rho_min), rho_max);
                glist_refresh<G, GRID>(sc, grid, want ? x : 0.f, want ? y : 0.f, rho, c.r, lane, gbase, j0, want, gl);
            }
        }
        // ---- (2) free flight: lane k keeps step t + k ---------------------------------------------------
        const int Tq = alive ? min(T, n - t) : 0;
        const int Tmax = __reduce_max_sync(0xffffffffu, Tq);
        float sx = x, sy = y, svx = vx, svy = vy;  // state at the START of my step
        State my; my.x = x; my.y = y; my.vx = vx; my.vy = vy; my.ax = 0.f; my.ay = 0.f;
        bool ok = true;
        {
            float cx = x, cy = y, cvx = vx, cvy = vy;
            const int last = min(lane, Tq - 1);
            for (int k = 0; k < Tmax; ++k) {
                if (k <= last) {
                    sx = cx; sy = cy; svx = cvx; svy = cvy;
                    ok = free_step_fast(cx, cy, cvx, cvy, c, my) && ok;
                    cx = my.x; cy = my.y; cvx = my.vx; cvy = my.vy;
                }
            }
        }
        if (__any_sync(0xffffffffu, !ok)) {  // a numerator left the range of the fast division: redo exactly
            float cx = x, cy = y, cvx = vx, cvy = vy;
            const int last = min(lane, Tq - 1);
            for (int k = 0; k < Tmax; ++k) {
                if (k <= last) {
                    sx = cx; sy = cy; svx = cvx; svy = cvy;
                    free_step(cx, cy, cvx, cvy, c, my);
                    cx = my.x; cy = my.y; cvx = my.vx; cvy = my.vy;
                }
            }
        }
        // ---- (3) my step against the list -------------------------------------------------------------------
        const bool valid = lane < Tq;
        const float px = valid ? my.x : x, py = valid ? my.y : y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        const bool fast = !sc.weird && M <= 1.0e15f && trouble != 2;  // false for NaN / inf points
        const float ddx = px - gl.cx, ddy = py - gl.cy;
        const bool stale = fast && !(ddx * ddx + ddy * ddy <= gl.rho2);
        const float amin = (fast && !stale) ? glist_min(sc, px, py, gl) : 0.0f;
        const bool near = !fast || exact_idx || !(amin > gl.far2);
        int idx = 0;
        float dist = __int_as_float(0x7f800000);
        if (valid && near && !stale) {  // exact distance and index for my own step (no warp primitives inside)
            unsigned long long key = kNoKey;
            if (fast) key = glist_exact(sc, px, py, __fmaf_rn(amin, kFilterOneEps, kFilterKappaPerM2 * M * M), gl.e, gl.cap, gl.cnt);
            if (key == kNoKey) key = lane_exact<1>(sc, px, py, 0);  // not fast; or (cannot happen) an empty candidate set
            unpack_key(key, idx, dist);
        }
        const bool hit = valid && !stale && (dist <= c.r);  // NaN -> false (ball_sim.py:164)
        // ---- (4) first event of the group, commit ---------------------------------------------------------
        const unsigned evb = (__ballot_sync(0xffffffffu, valid && (stale || hit)) >> gbase) & kLow;
        const int f = evb ? (__ffs(evb) - 1) : Tq;       // lanes < f: free flight, final
        const int hit_f = __shfl_sync(0xffffffffu, (int)hit, f < Tq ? f : 0, G);  // every lane executes the shuffle
        const bool ev_hit = (f < Tq) && hit_f != 0;
        State out = my;
        float nx = __shfl_sync(0xffffffffu, my.x, f > 0 ? f - 1 : 0, G), ny = __shfl_sync(0xffffffffu, my.y, f > 0 ? f - 1 : 0, G);
        float nvx = __shfl_sync(0xffffffffu, my.vx, f > 0 ? f - 1 : 0, G), nvy = __shfl_sync(0xffffffffu, my.vy, f > 0 ? f - 1 : 0, G);
        if (f == 0) { nx = x; ny = y; nvx = vx; nvy = vy; }
        if (__any_sync(0xffffffffu, ev_hit)) {  // resolve the collision of step t + f (all lanes of the group, redundantly)
            const int src = ev_hit ? f : 0;
            const float bx = __shfl_sync(0xffffffffu, sx, src, G), by = __shfl_sync(0xffffffffu, sy, src, G);
            const float bvx = __shfl_sync(0xffffffffu, svx, src, G), bvy = __shfl_sync(0xffffffffu, svy, src, G);
            State b1;
            b1.x = __shfl_sync(0xffffffffu, my.x, src, G); b1.y = __shfl_sync(0xffffffffu, my.y, src, G);
            b1.vx = __shfl_sync(0xffffffffu, my.vx, src, G); b1.vy = __shfl_sync(0xffffffffu, my.vy, src, G);
            b1.ax = __shfl_sync(0xffffffffu, my.ax, src, G); b1.ay = __shfl_sync(0xffffffffu, my.ay, src, G);
            const int hidx = __shfl_sync(0xffffffffu, idx, src, G);
            if (ev_hit) {
                const State r = resolve(bx, by, bvx, bvy, b1, sc.q[hidx], c);
                if (lane == f) out = r;
                nx = r.x; ny = r.y; nvx = r.vx; nvy = r.vy;
            }
        }
        const int n_commit = f + (ev_hit ? 1 : 0);
        if (active && lane < n_commit) {
            const int ts = t + lane;
            const bool h = ev_hit && lane == f;
            if (hl) hl[(int64_t)ts * p.hl_st] = pack_log((h || exact_idx) ? idx : 0, h, clip_bits(svx, svy));
            if (h && p.hitstate) p.hitstate[(size_t)ts * p.B + ball] = make_float4(sx, sy, svx, svy);
            if (p.ckpt && (ts % K) == 0) p.ckpt[(size_t)(ts / K) * p.B + ball] = make_float4(sx, sy, svx, svy);
            const size_t o = (size_t)ts * p.B + ball;
            if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
            if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
            if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
        }
        x = nx; y = ny; vx = nvx; vy = nvy;
        t += n_commit;
        // ---- (5) speculation length / trouble handling for the next round ---------------------------
        if (sat_out) {
            // contact mode this round: T / trouble were set there
        } else if (f < Tq && !ev_hit) {      // the list did not cover step t + f
            trouble = (f == 0) ? min(trouble + 1, 2) : 1;
        } else {
            trouble = 0;
            if (ev_hit && f == 0) {          // colliding step after step: stop speculating, and rebuild the
                if (T != 1) trouble = 1;     // list for the (now slow) ball: contact steps scan it serially
                T = 1;
            }
            else T = min(2 * T, G);
        }
    }
    if (active && lane == 0) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Small scenes (S_pad <= 256): every lane keeps ITS OWN segments (indices lane, lane+G, ...) in
// registers for the whole rollout, so the estimate loop is straight-line code with no loads at
// all (NV = S_pad / G visits, fully unrolled, candidates recorded in one 32-bit mask). The few
// reference-arithmetic evaluations (hint, candidates, collision) read the prepared scene through
// the read-only L1 path. Same results as rollout_fwd_kernel, by the same argument.
#ifndef DOPER_REG_THREADS
#define DOPER_REG_THREADS 128
#endif
// Block size of the register-resident kernel. (32-thread blocks spread a 4,096-ball batch more
// evenly over the 148 SMs but measured the same 0.636 ms: each warp is latency-bound on its own.)
constexpr int kRegThreads = DOPER_REG_THREADS;

// Register budget per G so that a 4,096-ball batch is co-resident in ONE wave (a second wave of a
// whole-rollout kernel doubles the time): blocks/SM = 2, 4, 7 for G = 8, 16, 32.
template <int G, int NV>
__global__ void __launch_bounds__(kRegThreads, (G == 8 ? 2 : (G == 16 ? 4 : 7)) * (128 / kRegThreads))
rollout_fwd_reg_kernel(const FwdParams p) {
    constexpr int BPB = kRegThreads / G;
    const int tid = threadIdx.x, lane = tid % G, slot = tid / G;
    int64_t ball = (int64_t)blockIdx.x * BPB + slot;
    const bool active = ball < p.B;
    if (!active) ball = p.B - 1;  // duplicate the last ball: keeps every lane in the warp-wide REDUX
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const int64_t scene_id = p.per_env ? ball : 0;
    const float4* __restrict__ gq = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) + scene_id * L.S_pad;
    const float* __restrict__ grl = reinterpret_cast<const float*>(p.scene_ws + L.rl_off) + scene_id * L.S_pad;
    const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[scene_id];
    SceneView sc;
    sc.q = gq; sc.rl = grl; sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale;
    sc.weird = hdr.weird | p.force_exact;

    float4 q[NV];
    float rl[NV];
#pragma unroll
    for (int j = 0; j < NV; ++j) {
        q[j] = __ldg(gq + lane + j * G);
        rl[j] = __ldg(grl + lane + j * G);
    }

    const float2 xx = p.x0[ball], vv = p.v0[ball];
    float x = xx.x, y = xx.y, vx = vv.x, vy = vv.y;
    const Consts c = p.c;
    const bool writer = active && lane == 0;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    float4* ck = p.ckpt ? p.ckpt + ball : nullptr;
    float2* tx = p.traj_x ? p.traj_x + ball : nullptr;
    float2* tv = p.traj_v ? p.traj_v + ball : nullptr;
    float2* ta = p.traj_a ? p.traj_a + ball : nullptr;
    int prev_idx = 0, until_ckpt = 0;
    const int64_t stride = p.B;
    if (!sc.weird) {  // hint for the first step: smallest estimate at the initial position (any index is valid)
        float best = __int_as_float(0x7f800000);
        int bj = 0;
#pragma unroll
        for (int j = 0; j < NV; ++j) {
            const float a = approx_d2(x, y, q[j], rl[j]);
            if (a < best) { best = a; bj = j; }
        }
        const unsigned long long k =
            group_min_redux<G>(((unsigned long long)__float_as_uint(best) << 32) | (unsigned)(lane + bj * G));
        const int h = (int)(unsigned)(k & 0xffffffffull);
        prev_idx = (h >= 0 && h < p.S) ? h : 0;
    }

    for (int t = 0; t < p.n_steps; ++t) {
        if (until_ckpt == 0) {
            until_ckpt = p.K;
            if (writer && ck) { *ck = make_float4(x, y, vx, vy); ck += stride; }
        }
        --until_ckpt;
        State s1;
        free_step(x, y, vx, vy, c, s1);
        const uint32_t clipb = clip_bits(vx, vy);
        const float px = s1.x, py = s1.y;
        const float M = fmaxf(sc.scale, fmaxf(fabsf(px), fabsf(py)));
        unsigned long long key;
        if (!sc.weird && M <= 1.0e15f) {
            const float kappa = kFilterKappaPerM2 * M * M;
            const float thr = __fmaf_rn(approx_d2(px, py, __ldg(gq + prev_idx), __ldg(grl + prev_idx)),
                                        kFilterOneEps, kappa);
            unsigned mask = 0u;
#pragma unroll
            for (int j = 0; j < NV; ++j) {
                const float a = approx_d2(px, py, q[j], rl[j]);
                mask |= (a <= thr) ? (1u << j) : 0u;
            }
            key = kNoKey;
            while (mask) {
                const int j = __ffs(mask) - 1;
                mask &= mask - 1u;
                const int cidx = lane + j * G;
                const unsigned long long kk = make_key(cand_dist(px, py, __ldg(gq + cidx)), cidx);
                key = kk < key ? kk : key;
            }
        } else {
            key = lane_exact<G>(sc, px, py, lane);
        }
        key = group_min_redux<G>(key);
        int idx; float dist;
        unpack_key(key, idx, dist);
        prev_idx = idx;
        const bool hit = dist <= c.r;  // NaN -> false (ball_sim.py:164)
        if (hit && writer && p.hitstate) p.hitstate[(size_t)t * p.B + ball] = make_float4(x, y, vx, vy);
        if (hit) s1 = resolve(x, y, vx, vy, s1, __ldg(gq + idx), c);
        x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
        if (writer) {
            if (hl) { *hl = pack_log(idx, hit, clipb); hl += p.hl_st; }
            if (tx) { *tx = make_float2(x, y); tx += stride; }
            if (tv) { *tv = make_float2(vx, vy); tv += stride; }
            if (ta) { *ta = make_float2(s1.ax, s1.ay); ta += stride; }
        }
    }
    if (writer) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// Time-parallel forward for small and medium batches: ONE WARP PER BALL, LANE = TIME STEP.
// Collisions are rare (~1 per 1,000 ball-steps), so the warp first integrates 32 steps of free
// flight (phase 1: O(1) work per step, every lane redundantly, lane k keeps step k), then all 32
// lanes run the closest-segment sweep of their own step at once (phase 2: the scene is read from
// shared memory as warp-wide broadcasts, no cross-lane reduction, no redundant sweep work).  A
// ballot finds the first step that hit; steps before it are committed exactly as a sequential
// rollout would have produced them (free flight until a hit is the reference's own branch), the
// hit step is resolved, and the next chunk starts right after it.  A ball that keeps colliding
// (resting against a wall) switches to the lane-split sequential sweep for a while.
// Results are bit-identical to the sequential kernels: same arithmetic per step, only scheduled
// speculatively.
constexpr int kTpEarlyHit = 6;     // a hit this early in a chunk ...
constexpr int kTpSeqSteps = 64;    // ... sends the ball to sequential mode for this many steps

__global__ void __launch_bounds__(1024) rollout_fwd_tp_kernel(const FwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int wpb = blockDim.x >> 5;  // balls per block
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const int n_local = p.per_env ? wpb : 1;
    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)n_local * L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)n_local * L.S_pad * 20);
    const int64_t first_ball = (int64_t)blockIdx.x * wpb;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) {
        const int n = p.per_env ? (int)min((int64_t)wpb, p.B - first_ball) : 1;
        stage_scenes(sq, srl, p.scene_ws, L, p.per_env ? first_ball : 0, n, bar);
    }
    const int64_t ball = first_ball + warp;
    const bool active = ball < p.B;  // warp-uniform
    SceneView sc;
    float x = 0.f, y = 0.f, vx = 0.f, vy = 0.f;
    if (active) {
        const SceneHeader hdr = reinterpret_cast<const SceneHeader*>(p.scene_ws + L.hdr_off)[p.per_env ? ball : 0];
        const int ls = p.per_env ? warp : 0;
        sc.q = sq + (size_t)ls * L.S_pad; sc.rl = srl + (size_t)ls * L.S_pad;
        sc.S = p.S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird | p.force_exact;
        const float2 xx = p.x0[ball], vv = p.v0[ball];
        x = xx.x; y = xx.y; vx = vv.x; vy = vv.y;
    }
    mbar_wait(bar, 0);
    if (!active) return;  // whole warp leaves; no block-wide sync below

    const Consts c = p.c;
    int32_t* hl = p.hitlog ? p.hitlog + ball * p.hl_sb : nullptr;
    int t = 0, prev_idx = -1, seq_left = 0;
    const int n = p.n_steps, K = p.K;

    while (t < n) {
        if (seq_left > 0) {
            // ---- sequential mode: lanes split the segments of ONE step ----
            if (p.ckpt && (t % K) == 0 && lane == 0) p.ckpt[(size_t)(t / K) * p.B + ball] = make_float4(x, y, vx, vy);
            State s1;
            free_step(x, y, vx, vy, c, s1);
            const uint32_t clipb = clip_bits(vx, vy);
            int idx; float dist;
            sweep_any<32>(sc, s1.x, s1.y, lane, prev_idx, idx, dist);
            prev_idx = idx;
            const bool hit = dist <= c.r;
            if (hit && lane == 0 && p.hitstate) p.hitstate[(size_t)t * p.B + ball] = make_float4(x, y, vx, vy);
            if (hit) { s1 = resolve(x, y, vx, vy, s1, sc.q[idx], c); seq_left = kTpSeqSteps; }
            x = s1.x; y = s1.y; vx = s1.vx; vy = s1.vy;
            if (lane == 0) {
                const size_t o = (size_t)t * p.B + ball;
                if (hl) hl[(size_t)t * p.hl_st] = pack_log(idx, hit, clipb);
                if (p.traj_x) p.traj_x[o] = make_float2(x, y);
                if (p.traj_v) p.traj_v[o] = make_float2(vx, vy);
                if (p.traj_a) p.traj_a[o] = make_float2(s1.ax, s1.ay);
            }
            ++t; --seq_left;
            continue;
        }
        // ---- speculative chunk: phase 1, free flight; lane k keeps step t+k ----
        const int T = min(32, n - t);
        float sx = x, sy = y, svx = vx, svy = vy;  // state at the START of my step
        State my; my.x = x; my.y = y; my.vx = vx; my.vy = vy; my.ax = 0.f; my.ay = 0.f;
        {
            float cx = x, cy = y, cvx = vx, cvy = vy;
            const int last = min(lane, T - 1);  // lane k integrates steps 0..k, then holds
            for (int k = 0; k < T; ++k) {
                if (k <= last) {
                    sx = cx; sy = cy; svx = cvx; svy = cvy;
                    free_step(cx, cy, cvx, cvy, c, my);
                    cx = my.x; cy = my.y; cvx = my.vx; cvy = my.vy;
                }
            }
        }
        // ---- phase 2: every lane sweeps the whole scene for its own step ----
        int idx; float dist;
        if (prev_idx < 0) prev_idx = hint_by_estimate<32>(sc, x, y, lane);  // first chunk only
        {
            const float M = fmaxf(sc.scale, fmaxf(fabsf(my.x), fabsf(my.y)));
            unsigned long long key;
            if (!sc.weird && M <= 1.0e15f) key = lane_filtered<1>(sc, my.x, my.y, 0, prev_idx, kFilterKappaPerM2 * M * M);
            else key = lane_exact<1>(sc, my.x, my.y, 0);
            unpack_key(key, idx, dist);
        }
        const bool hit = (lane < T) && (dist <= c.r);  // NaN -> false (ball_sim.py:164)
        const unsigned hits = __ballot_sync(0xffffffffu, hit);
        const int first = hits ? (__ffs(hits) - 1) : T;  // first step of the chunk that collided
        const int n_commit = min(first + 1, T);          // steps t .. t+n_commit-1 are final
        State out = my;
        if (first < T) {  // resolve the collision of step t+first (all lanes, redundantly)
            const float bx = __shfl_sync(0xffffffffu, sx, first), by = __shfl_sync(0xffffffffu, sy, first);
            const float bvx = __shfl_sync(0xffffffffu, svx, first), bvy = __shfl_sync(0xffffffffu, svy, first);
            State b1;
            b1.x = __shfl_sync(0xffffffffu, my.x, first); b1.y = __shfl_sync(0xffffffffu, my.y, first);
            b1.vx = __shfl_sync(0xffffffffu, my.vx, first); b1.vy = __shfl_sync(0xffffffffu, my.vy, first);
            b1.ax = __shfl_sync(0xffffffffu, my.ax, first); b1.ay = __shfl_sync(0xffffffffu, my.ay, first);
            const int hidx = __shfl_sync(0xffffffffu, idx, first);
            const State r = resolve(bx, by, bvx, bvy, b1, sc.q[hidx], c);
            if (lane == first) out = r;
            x = r.x; y = r.y; vx = r.vx; vy = r.vy;
            if (first < kTpEarlyHit) seq_left = kTpSeqSteps;
        } else {  // no collision: the state after the last step of the chunk
            x = __shfl_sync(0xffffffffu, my.x, T - 1); y = __shfl_sync(0xffffffffu, my.y, T - 1);
            vx = __shfl_sync(0xffffffffu, my.vx, T - 1); vy = __shfl_sync(0xffffffffu, my.vy, T - 1);
        }
        prev_idx = __shfl_sync(0xffffffffu, idx, n_commit - 1);
        if (lane < n_commit) {
            const int ts = t + lane;
            if (hl) hl[(size_t)ts * p.hl_st] = pack_log(idx, lane == first, clip_bits(svx, svy));
            if (lane == first && p.hitstate) p.hitstate[(size_t)ts * p.B + ball] = make_float4(sx, sy, svx, svy);
            if (p.ckpt && (ts % K) == 0) p.ckpt[(size_t)(ts / K) * p.B + ball] = make_float4(sx, sy, svx, svy);
            const size_t o = (size_t)ts * p.B + ball;
            if (p.traj_x) p.traj_x[o] = make_float2(out.x, out.y);
            if (p.traj_v) p.traj_v[o] = make_float2(out.vx, out.vy);
            if (p.traj_a) p.traj_a[o] = make_float2(out.ax, out.ay);
        }
        t += n_commit;
    }
    if (lane == 0) {
        p.xT[ball] = make_float2(x, y);
        p.vT[ball] = make_float2(vx, vy);
    }
}

// ------------------------------------------------------------------------------------------
// backward rollout
// ------------------------------------------------------------------------------------------
constexpr int kBwdThreads = 128;

struct BwdParams {
    int64_t B;
    int S, per_env, n_steps, K;
    Consts c;
    const unsigned char* scene_ws;
    const float4* ckpt;
    const float4* hitstate;  // nullable: [n_steps][B] states at the start of collided steps
    const int32_t* hitlog;  // element (t, ball) at t * hl_st + ball * hl_sb
    int64_t hl_st, hl_sb;
    const float2* xT_bar;
    const float2* vT_bar;  // nullable
    float2* x0_bar;        // nullable
    float2* v0_bar;
    // hybrid plan of the sequential kernel: a ball that has pulled back more than defer_after collided
    // steps (sliding contact: a ~2.6 us serial chain per step) appends itself to `deferred` ([0] = count,
    // then ball ids) and leaves; rollout_bwd_chunked_kernel then redoes those balls 32 chunks in parallel
    int* deferred;         // nullable
    int defer_after;
};

constexpr int kDeferCap = 2048;  // deferred balls per launch; beyond it balls simply stay sequential

// Sequential backward, one thread per ball, bit-exact against the oracle's adjoint.  A step that
// did not collide is pulled back from its log word alone (step_vjp_free: no state, no replay);
// only a K-block that contains a collision is replayed from its checkpoint (with the logged
// indices), up to its last collision, into the shared-memory stash.
constexpr int kBwdPrefetch = 16;  // log words fetched ahead per thread (independent loads in flight)

__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_kernel(const BwdParams p) {
    extern __shared__ __align__(128) unsigned char smem[];
    float4* stash = reinterpret_cast<float4*>(smem);  // [K][kBwdThreads], replay fallback only
    const int tid = threadIdx.x;
    const int64_t ball = (int64_t)blockIdx.x * kBwdThreads + tid;
    if (ball >= p.B) return;
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    const float4* q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) +
                      (p.per_env ? (size_t)ball * L.S_pad : 0);
    const Consts c = p.c;
    const float dk = dfric_value(c);
    float2 g = p.xT_bar[ball];
    float gx = g.x, gy = g.y, gvx = 0.0f, gvy = 0.0f;
    if (p.vT_bar) { float2 h = p.vT_bar[ball]; gvx = h.x; gvy = h.y; }
    const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;

    const int n_ckpt = (p.n_steps + p.K - 1) / p.K;
    int collided = 0;
    for (int ck = n_ckpt - 1; ck >= 0; --ck) {
        const int t0 = ck * p.K, t1 = min(p.n_steps, t0 + p.K);
        int have = -1;  // states of steps t0 .. t0 + have are in the stash
        for (int tb = t1; tb > t0;) {
            // the log words of steps [ta, tb): kBwdPrefetch independent loads, then a latency-free loop
            const int ta = max(t0, tb - kBwdPrefetch);
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
                if (p.deferred && ++collided > p.defer_after) {
                    const int slot = atomicAdd(p.deferred, 1);
                    if (slot < kDeferCap) { p.deferred[1 + slot] = (int)ball; return; }  // redone chunk-parallel
                    collided = -0x40000000;  // list full: stay sequential
                }
                const int32_t h = hl[(size_t)t * p.hl_st];  // collided step: the index is needed (rare reload)
                float4 st;
                if (p.hitstate) {
                    st = p.hitstate[(size_t)t * p.B + ball];  // the forward kept the state of every collided step
                } else {
                    if (have < 0) {  // first (= latest) collision of the block: replay t0 .. t, bit-identical states
                        float4 s = p.ckpt[(size_t)ck * p.B + ball];
                        for (int u = t0; u <= t; ++u) {
                            stash[(size_t)(u - t0) * kBwdThreads + tid] = s;
                            if (u < t) {
                                const int32_t hu = hl[(size_t)u * p.hl_st];
                                State s1;
                                free_step(s.x, s.y, s.z, s.w, c, s1);
                                if (hu < 0) s1 = resolve(s.x, s.y, s.z, s.w, s1, __ldg(q + (hu & kLogIdx)), c);
                                s = make_float4(s1.x, s1.y, s1.vx, s1.vy);
                            }
                        }
                        have = t - t0;
                    }
                    st = stash[(size_t)(t - t0) * kBwdThreads + tid];
                }
                const float4 seg = __ldg(q + (h & kLogIdx));
                step_vjp(st.x, st.y, st.z, st.w, true, seg, c, gx, gy, gvx, gvy);
            }
            tb = ta;
        }
    }
    p.v0_bar[ball] = make_float2(gvx, gvy);
    if (p.x0_bar) p.x0_bar[ball] = make_float2(gx, gy);
}

// Chunk-parallel backward for small batches (few balls cannot fill 148 SMs with one thread per
// ball): thread (ball, chunk) pulls the four basis cotangents e_0..e_3 back through its K-step
// chunk, which yields the chunk's transposed Jacobian M_c = J_t0^T ... J_t1-1^T (4x4; the step
// adjoint is linear in the cotangent).  Free-flight steps come from the log word alone; a chunk
// with a collision is replayed from its checkpoint first.  One thread per ball then folds
// g <- M_c g from the last chunk to the first.  Same mathematics as the sequential kernel,
// different rounding order (tolerance-checked, not bit-exact).
// No block-wide synchronisation: a ball's C_pad <= 32 chunks sit in one warp, the fold runs on warp
// shuffles.  Until a chunk meets a collision its Jacobian is block diagonal (x and y decouple in
// free flight), so only the four non-zero (position, velocity) pairs are pulled back.
// With `ball_list` ([0] = count, then ids; hybrid plan of the sequential kernel) the grid covers the
// listed balls only and `chunk_len` need not be the checkpoint period (the hit-state array is required).
__global__ void __launch_bounds__(kBwdThreads) rollout_bwd_chunked_kernel(const BwdParams p, const int C_pad,
                                                                         const int chunk_len,
                                                                         const int* __restrict__ ball_list) {
    const int tid = threadIdx.x;
    const int bpb = kBwdThreads / C_pad;
    const int chunk = tid % C_pad, slot = tid / C_pad;
    int64_t ball = (int64_t)blockIdx.x * bpb + slot;
    if (ball_list) {
        const int count = min(ball_list[0], kDeferCap);
        ball = ball < count ? (int64_t)ball_list[1 + ball] : p.B;  // p.B = no ball
    }
    const int n_ckpt = (p.n_steps + chunk_len - 1) / chunk_len;
    const bool active = ball < p.B && chunk < n_ckpt;
    const Consts c = p.c;
    const float dk = dfric_value(c);
    const SceneLayout L = scene_layout(p.per_env ? p.B : 1, p.S);
    float g[4][4];  // g[k] = pull-back of e_k through this thread's chunk
#pragma unroll
    for (int k = 0; k < 4; ++k)
#pragma unroll
        for (int r = 0; r < 4; ++r) g[k][r] = (k == r) ? 1.0f : 0.0f;
    if (active) {
        const float4* q = reinterpret_cast<const float4*>(p.scene_ws + L.q_off) +
                          (p.per_env ? (size_t)ball * L.S_pad : 0);
        const int32_t* hl = p.hitlog + (size_t)ball * p.hl_sb;
        const int t0 = chunk * chunk_len, t1 = min(p.n_steps, t0 + chunk_len);
        bool coupled = false;  // a collision has mixed x and y
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
            if (m_hit == 0u && !coupled) {  // free flight, decoupled: 4 (position, velocity) pairs per step
                for (int j = tb - ta - 1; j >= 0; --j) {
                    const bool cx = ((m_cx >> j) & 1u) != 0u, cy = ((m_cy >> j) & 1u) != 0u;
                    step_vjp_free_1d(cx, dk, c, g[0][0], g[0][2]);
                    step_vjp_free_1d(cx, dk, c, g[2][0], g[2][2]);
                    step_vjp_free_1d(cy, dk, c, g[1][1], g[1][3]);
                    step_vjp_free_1d(cy, dk, c, g[3][1], g[3][3]);
                }
            } else {
                for (int t = tb - 1; t >= ta; --t) {
                    const int32_t h = hl[(size_t)t * p.hl_st];
                    if (h >= 0) {
#pragma unroll
                        for (int k = 0; k < 4; ++k) step_vjp_free((uint32_t)h, dk, c, g[k][0], g[k][1], g[k][2], g[k][3]);
                        continue;
                    }
                    float4 st;
                    if (p.hitstate) {
                        st = p.hitstate[(size_t)t * p.B + ball];
                    } else {  // no hit-state array: replay t0 .. t from the chunk's checkpoint (<= K steps)
                        st = p.ckpt[(size_t)chunk * p.B + ball];
                        for (int u = t0; u < t; ++u) {
                            const int32_t hu = hl[(size_t)u * p.hl_st];
                            State s1;
                            free_step(st.x, st.y, st.z, st.w, c, s1);
                            if (hu < 0) s1 = resolve(st.x, st.y, st.z, st.w, s1, __ldg(q + (hu & kLogIdx)), c);
                            st = make_float4(s1.x, s1.y, s1.vx, s1.vy);
                        }
                    }
                    const float4 seg = __ldg(q + (h & kLogIdx));
                    {   // through a copy: the out-of-line call takes its array by reference (local memory), and
                        // g itself must stay in registers for the straight-line free-flight path
                        float tmp[4][4];
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int r2 = 0; r2 < 4; ++r2) tmp[k][r2] = g[k][r2];
                        step_vjp_hit_n<4>(st.x, st.y, st.z, st.w, seg, c, tmp);
#pragma unroll
                        for (int k = 0; k < 4; ++k)
#pragma unroll
                            for (int r2 = 0; r2 < 4; ++r2) g[k][r2] = tmp[k][r2];
                    }
                    coupled = true;
                }
            }
            tb = ta;
        }
    }
    // fold g <- M_c g from the last chunk to the first; lane `chunk` of the ball's C_pad lanes holds M_c
    float v[4] = {0.f, 0.f, 0.f, 0.f};
    if (ball < p.B) {
        const float2 gg = p.xT_bar[ball];
        v[0] = gg.x; v[1] = gg.y;
        if (p.vT_bar) { const float2 h = p.vT_bar[ball]; v[2] = h.x; v[3] = h.y; }
    }
    for (int cc = n_ckpt - 1; cc >= 0; --cc) {
        float o[4];
#pragma unroll
        for (int r = 0; r < 4; ++r)
            o[r] = (v[0] * g[0][r] + v[1] * g[1][r]) + (v[2] * g[2][r] + v[3] * g[3][r]);
#pragma unroll
        for (int r = 0; r < 4; ++r) v[r] = __shfl_sync(0xffffffffu, o[r], cc, C_pad);  // chunk cc's product
    }
    if (ball < p.B && chunk == 0) {
        p.v0_bar[ball] = make_float2(v[2], v[3]);
        if (p.x0_bar) p.x0_bar[ball] = make_float2(v[0], v[1]);
    }
}

// ------------------------------------------------------------------------------------------
// loss
// ------------------------------------------------------------------------------------------
constexpr int kLossThreads = 256;

// block 0: deterministic total (fixed strided partials + fixed tree); blocks >= 1: per-ball outputs
__global__ void __launch_bounds__(kLossThreads) l1_loss_kernel(int64_t B, const float2* __restrict__ xT,
                                                              float tx, float ty,
                                                              float* __restrict__ loss_pb,
                                                              float* __restrict__ loss_sum,
                                                              float2* __restrict__ xT_bar) {
    if (blockIdx.x == 0) {
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
    const int64_t b = (int64_t)(blockIdx.x - 1) * kLossThreads + threadIdx.x;
    if (b >= B) return;
    float2 p = xT[b];
    float ex = p.x - tx, ey = p.y - ty;
    if (loss_pb) loss_pb[b] = fabsf(ex) + fabsf(ey);
    if (xT_bar) xT_bar[b] = make_float2(sgnf(ex), sgnf(ey));  // d|e|/de, 0 at e == 0 (and NaN -> 0)
}

// ------------------------------------------------------------------------------------------
// geometry queries
// ------------------------------------------------------------------------------------------
constexpr int kQueryThreads = 128;

template <int G>
__global__ void __launch_bounds__(kQueryThreads) closest_segment_kernel(
    int64_t N, int S, const float2* __restrict__ pts, const float4* __restrict__ raw,
    const unsigned char* __restrict__ scene_ws, int32_t* __restrict__ idx_out,
    float* __restrict__ dist_out, float4* __restrict__ seg_out) {
    extern __shared__ __align__(128) unsigned char smem[];
    const SceneLayout L = scene_layout(1, S);
    float4* sq = reinterpret_cast<float4*>(smem);
    float* srl = reinterpret_cast<float*>(smem + (size_t)L.S_pad * sizeof(float4));
    uint64_t* bar = reinterpret_cast<uint64_t*>(smem + (size_t)L.S_pad * 20);
    const int tid = threadIdx.x, lane = tid % G;
    if (tid == 0) mbar_init(bar, 1);
    __syncthreads();
    if (tid == 0) stage_scenes(sq, srl, scene_ws, L, 0, 1, bar);
    const SceneHeader hdr = *reinterpret_cast<const SceneHeader*>(scene_ws + L.hdr_off);
    SceneView sc;
    sc.q = sq; sc.rl = srl; sc.S = S; sc.S_pad = L.S_pad; sc.scale = hdr.scale; sc.weird = hdr.weird;
    int64_t i = (int64_t)blockIdx.x * (kQueryThreads / G) + tid / G;
    const bool active = i < N;
    if (!active) i = N - 1;
    const float2 p = pts[i];
    mbar_wait(bar, 0);
    int idx; float dist;
    sweep_any<G>(sc, p.x, p.y, lane, -1, idx, dist);
    if (active && lane == 0) {
        idx_out[i] = idx;
        dist_out[i] = dist;
        if (seg_out) seg_out[i] = raw[idx];
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
            float nrm = sqrtf(nx * nx + ny * ny);        // :23
            tile[j] = make_float4(s.x, s.y, nx / nrm, ny / nrm);
        }
        __syncthreads();
        for (int t = 0; t + 2 < n; t += 3) {
            bool all = true;
#pragma unroll
            for (int e = 0; e < 3; ++e) {
                float4 q = tile[t + e];
                float pvx = px - q.x, pvy = py - q.y;    // :52
                float d = pvx * q.z + pvy * q.w;         // :53
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

}  // namespace doper

#include "sensor.cuh"

// ==========================================================================================
// C ABI
// ==========================================================================================
using namespace doper;

namespace {

inline bool aligned(const void* p, size_t a) { return (reinterpret_cast<uintptr_t>(p) % a) == 0; }

inline int launch_status() { return cudaPeekAtLastError() == cudaSuccess ? DOPER_OK : DOPER_ERR_LAUNCH; }

int max_optin_smem() {
    int dev = 0, v = 0;
    if (cudaGetDevice(&dev) != cudaSuccess) return 0;
    if (cudaDeviceGetAttribute(&v, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev) != cudaSuccess) return 0;
    return v;
}

Consts make_consts(const doper_rollout_desc* d) {
    Consts c;
    c.r = d->consts[0]; c.m = d->consts[1]; c.f = d->consts[2]; c.e = d->consts[3]; c.ks = d->consts[4];
    c.ax = d->attractor[0]; c.ay = d->attractor[1];
    c.dt = d->dt;
    // div_invariant(): fast path only for well-scaled divisors (then y = RN(1/b) is normal and the
    // residuals stay normal for |a| in [1e-18, 1e18]); otherwise an empty range = always a / b
    const bool ok = c.r >= 1e-6f && c.r <= 1e6f && c.m >= 1e-6f && c.m <= 1e6f;
    c.rcp_r = 1.0f / c.r;
    c.rcp_m = 1.0f / c.m;
    c.div_lo = ok ? 1e-18f : 1.0f;
    c.div_hi = ok ? 1e18f : 0.0f;
    return c;
}

// Backward plan. Small batches use the chunk-parallel kernel (<= 32 chunks per ball, one thread
// each); large batches have enough balls for one thread per ball and use the sequential kernel
// (bit-identical to the oracle's adjoint). flags bit 1 forces the sequential kernel.
constexpr int64_t kChunkedMaxBalls = 32768;
constexpr int kChunkedMaxK = 96;  // stash of K x 128 float4 + matrices must fit in shared memory

bool chunked_eligible(const doper_rollout_desc* d) {
    return !(d->flags & 2) && d->B <= kChunkedMaxBalls && d->n_steps >= 64;
}

int resolve_K(const doper_rollout_desc* d) {
    if (d->ckpt_every > 0) return d->ckpt_every;
    if (chunked_eligible(d)) {
        const int K = (d->n_steps + 31) / 32;
        if (K <= kChunkedMaxK) return K < 4 ? 4 : K;
    }
    return 16;  // sequential backward: 32 KB of stash per 128-thread block -> 7 blocks per SM
}

// Backward workspace: [checkpoints n_ckpt x B x 16 B | hit log n x B x 4 B | hit states n x B x 16 B].
// The hit-state array is written and read at collided steps only (no bandwidth; it spares the
// backward every replay); flags bit 19 drops it (memory: 16 B per ball-step) and the backward
// replays blocks with a collision from their checkpoint instead.
struct WsLayout { size_t log_off, state_off, defer_off, total; };

WsLayout ws_layout(const doper_rollout_desc* d) {
    const int K = resolve_K(d);
    const size_t n_ckpt = (size_t)(d->n_steps + K - 1) / K;
    WsLayout w;
    w.log_off = (n_ckpt * (size_t)d->B * sizeof(float4) + 255) / 256 * 256;
    size_t end = w.log_off + (size_t)d->n_steps * (size_t)d->B * sizeof(int32_t);
    end = (end + 255) / 256 * 256;
    w.state_off = 0;
    if (!(d->flags & (1 << 19))) {
        w.state_off = end;
        end += (size_t)d->n_steps * (size_t)d->B * sizeof(float4);
        end = (end + 255) / 256 * 256;
    }
    w.defer_off = end;  // scratch of the hybrid backward: deferred-ball list
    end += ((size_t)(kDeferCap + 1) * sizeof(int) + 255) / 256 * 256;
    w.total = end + 256;
    return w;
}

// Forward plan: time-parallel kernel (one warp per ball) unless the caller pinned a lane count,
// disabled it (flags bit 3) or the batch is large enough for one thread per ball.
constexpr int64_t kTpMaxBalls = 65536;

// Measured on B200 (tools/gpu_tune.py, profiles/): a shared small scene with a few thousand balls
// (BASELINE configs[1]) runs fastest on the register-resident lane-split kernel with 8 lanes per
// ball (0.63 ms vs 0.74 ms time-parallel); per-environment scenes, dense scenes and every other
// batch size up to 65,536 run fastest time-parallel; beyond that one thread per ball wins.
// Broad-phase (candidate list) kernels are the default plan.  They are not used when the caller
// forbids them (flags bit 15), forces the exact sweep (bit 0) or pins one of the full-sweep kernels
// (lanes_per_ball != 0 or the kernel-selection bits 2-3) without forcing them (bit 16).
// Measured on B200 (tools/tune_fwd.py, profiles/r1_plan_sweep.md): shared scene up to 8,192 balls ->
// time-parallel variant with 16 lanes (steps) per ball; 8,192-32,767 balls -> lane-split, 4 lanes;
// more -> one thread per ball; dense scenes (>= 2,048 segments) at least 2 lanes (the refresh scan
// is split over them); per-environment scenes -> lane-split, 4 lanes, scenes read in place.
constexpr int kLegacyPlanBits = 4 | 8;

// the lane = step kernel needs the scene in shared memory; larger scenes go to the lane-split kernel,
// which reads them in place
bool scene_fits_smem(const doper_rollout_desc* d) { return (size_t)scene_layout(1, d->S).S_pad * 20 <= 200 * 1024; }

bool use_cull(const doper_rollout_desc* d) {
    if ((d->flags & (1 << 15)) || (d->flags & 1)) return false;
    if (d->flags & (1 << 16)) return true;
    return d->lanes_per_ball == 0 && !(d->flags & kLegacyPlanBits);
}

bool use_tpc(const doper_rollout_desc* d) {
    if (!use_cull(d)) return false;
    if (d->lanes_per_ball != 0) return (d->flags & (1 << 18)) != 0 && d->lanes_per_ball >= 2;
    if (d->flags & (1 << 18)) return true;
    return !d->per_env && d->B <= 8192 && scene_fits_smem(d);
}

int cull_lanes(const doper_rollout_desc* d) {
    if (d->lanes_per_ball != 0) return d->lanes_per_ball;
    if (d->per_env) return 4;  // read in place (see launch_fwd_cull); measured: C3 5.4 ms vs 8.2 ms staged with 8 lanes
    if (use_tpc(d)) return 16;
    int g = d->B < 32768 ? 4 : 1;
    if (scene_layout(1, d->S).S_pad >= kGridMinSegments) g = 2;  // dense scene (grid refresh): measured best on C4
    if (!scene_fits_smem(d) && g < 8) g = 8;  // unstaged scene: split the refresh scan (L2 reads) over more lanes
    return g;
}

bool plan_reg8(const doper_rollout_desc* d) {
    return d->lanes_per_ball == 0 && !d->per_env && !(d->flags & 4) && scene_layout(1, d->S).S_pad <= 256 &&
           d->B >= 1536 && d->B <= 6144 && !use_cull(d);
}

bool use_tp(const doper_rollout_desc* d) {
    return d->lanes_per_ball == 0 && !(d->flags & 8) && d->B <= kTpMaxBalls && !plan_reg8(d) && !use_cull(d);
}

// lane = time step kernels write [B][n_steps] so that a group's stores are contiguous
bool log_ball_major(const doper_rollout_desc* d) { return use_tp(d) || use_tpc(d); }

int tp_warps_per_block(const doper_rollout_desc* d) {
    if (d->per_env) return 4;
    const size_t scene = (size_t)scene_layout(1, d->S).S_pad * 20;
    if (scene <= 24 * 1024) return 4;
    if (scene <= 56 * 1024) return 8;
    if (scene <= 110 * 1024) return 16;
    return 32;
}

bool use_chunked(const doper_rollout_desc* d, int K) {
    const int n_ckpt = (d->n_steps + K - 1) / K;
    return chunked_eligible(d) && n_ckpt <= 32 && n_ckpt >= 2 && K <= kChunkedMaxK;
}

int validate_desc(const doper_rollout_desc* d) {
    if (!d) return DOPER_ERR_NULL_PTR;
    if (d->B <= 0 || d->S <= 0 || d->n_steps < 0 || d->ckpt_every < 0) return DOPER_ERR_BAD_SHAPE;
    if (d->ckpt_every > 1024) return DOPER_ERR_BAD_SHAPE;
    const int g = d->lanes_per_ball;
    if (!(g == 0 || g == 1 || g == 2 || g == 4 || g == 8 || g == 16 || g == 32)) return DOPER_ERR_BAD_SHAPE;
    return DOPER_OK;
}

bool plan_reg8(const doper_rollout_desc* d);

// lanes per ball: enough threads to fill 148 SMs x 256 threads; per-env scenes are staged per
// ball, so spreading one ball over 8 lanes keeps the shared-memory footprint per thread sane.
int auto_lanes(const doper_rollout_desc* d) {
    if (use_cull(d)) return cull_lanes(d);
    if (d->lanes_per_ball) return d->lanes_per_ball;
    if (plan_reg8(d)) return 8;
    if (d->per_env) return 8;
    int g = 1;
    while (g < 32 && d->B * g < 148 * 256) g *= 2;
    return g;
}

size_t fwd_smem_bytes(const doper_rollout_desc* d, int G) {
    const SceneLayout L = scene_layout(1, d->S);
    const size_t n_local = d->per_env ? (size_t)(kFwdThreads / G) : 1;
    return n_local * L.S_pad * 20 + 16;
}

template <int G, int NV>
int launch_fwd_reg(cudaStream_t st, const FwdParams& p) {
    const int bpb = kRegThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    rollout_fwd_reg_kernel<G, NV><<<grid, kRegThreads, 0, st>>>(p);
    return launch_status();
}

// register-resident scene: S_pad in {32,...,256}, G in {8,16,32}, NV = S_pad / G
template <int G>
int dispatch_fwd_reg(cudaStream_t st, const FwdParams& p, int S_pad) {
    switch (S_pad / 32) {
        case 1: return launch_fwd_reg<G, 32 / G>(st, p);
        case 2: return launch_fwd_reg<G, 64 / G>(st, p);
        case 3: return launch_fwd_reg<G, 96 / G>(st, p);
        case 4: return launch_fwd_reg<G, 128 / G>(st, p);
        case 5: return launch_fwd_reg<G, 160 / G>(st, p);
        case 6: return launch_fwd_reg<G, 192 / G>(st, p);
        case 7: return launch_fwd_reg<G, 224 / G>(st, p);
        default: return launch_fwd_reg<G, 256 / G>(st, p);
    }
}

template <int G>
int launch_fwd(cudaStream_t st, const FwdParams& p, size_t smem) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(rollout_fwd_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const int bpb = kFwdThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    rollout_fwd_kernel<G><<<grid, kFwdThreads, smem, st>>>(p);
    return launch_status();
}

template <int G>
int launch_fwd_cull(cudaStream_t st, FwdParams p, size_t scene_smem) {
    const size_t lists = (size_t)CullCfg<G>::LCAP * kFwdThreads * sizeof(unsigned short);
    // a scene that does not fit shared memory (next to the lists) is read in place
    p.stage = (int)(scene_smem + lists) <= max_optin_smem() ? 1 : 0;
    // per-environment scenes with few lanes per ball: staging 128 / G private scenes per block leaves room for
    // one block per SM at best; reading them in place (each env's 4.5 KB stays in L1 / L2) measured faster
    if (p.per_env && G < 8) p.stage = 0;
    const size_t smem = (p.stage ? scene_smem : 0) + lists;
    const int bpb = kFwdThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    const bool dense = !p.per_env && scene_layout(1, p.S).S_pad >= kGridMinSegments;
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)DOPER_ERR_SMEM_LIMIT;
        kernel<<<grid, kFwdThreads, smem, st>>>(p);
        return launch_status();
    };
    if (p.stage) return dense ? go(rollout_fwd_cull_kernel<G, true, true>) : go(rollout_fwd_cull_kernel<G, true, false>);
    return dense ? go(rollout_fwd_cull_kernel<G, false, true>) : go(rollout_fwd_cull_kernel<G, false, false>);
}

template <int G>
int launch_fwd_tpc(cudaStream_t st, const FwdParams& p, size_t scene_smem) {
    const size_t smem = scene_smem + (size_t)(kFwdThreads / G) * kTpcListCap * sizeof(unsigned short);
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    const int bpb = kFwdThreads / G;
    const unsigned grid = (unsigned)((p.B + bpb - 1) / bpb);
    const bool dense = !p.per_env && scene_layout(1, p.S).S_pad >= kGridMinSegments;
    auto go = [&](auto kernel) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) != cudaSuccess)
            return (int)DOPER_ERR_SMEM_LIMIT;
        kernel<<<grid, kFwdThreads, smem, st>>>(p);
        return launch_status();
    };
    return dense ? go(rollout_fwd_tpc_kernel<G, true>) : go(rollout_fwd_tpc_kernel<G, false>);
}

template <int G>
int launch_closest(cudaStream_t st, int64_t N, int S, const float* pts, const float* segs, const void* ws,
                   int32_t* idx, float* dist, float* seg_out, size_t smem) {
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(closest_segment_kernel<G>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                             (int)smem) != cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const int ppb = kQueryThreads / G;
    const unsigned grid = (unsigned)((N + ppb - 1) / ppb);
    closest_segment_kernel<G><<<grid, kQueryThreads, smem, st>>>(
        N, S, reinterpret_cast<const float2*>(pts), reinterpret_cast<const float4*>(segs),
        reinterpret_cast<const unsigned char*>(ws), idx, dist, reinterpret_cast<float4*>(seg_out));
    return launch_status();
}

}  // namespace

namespace {
template <bool OBS>
int launch_range_sensor(cudaStream_t st, int64_t N, int R, int S, const float* pos, const float* vel,
                        const float* ray_dirs, const void* scene_ws, float distance_range, double tx, double ty,
                        float* ranges, float* points, int32_t* idx, float* obs) {
    const size_t smem = (size_t)scene_layout(1, S).S_pad * 20 + 16;
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(range_sensor_kernel<OBS>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const int64_t n = N * (int64_t)R;
    range_sensor_kernel<OBS><<<(unsigned)((n + kSensorThreads - 1) / kSensorThreads), kSensorThreads, smem, st>>>(
        N, R, S, reinterpret_cast<const float2*>(pos), reinterpret_cast<const float2*>(vel),
        reinterpret_cast<const float2*>(ray_dirs), reinterpret_cast<const unsigned char*>(scene_ws), distance_range,
        tx, ty, ranges, reinterpret_cast<float2*>(points), idx, obs);
    return launch_status();
}
}  // namespace

extern "C" {

int doper_abi_version(void) { return DOPER_ABI_VERSION; }

const char* doper_error_string(int code) {
    switch (code) {
        case DOPER_OK: return "ok";
        case DOPER_ERR_BAD_SHAPE: return "bad shape or size argument";
        case DOPER_ERR_NULL_PTR: return "required pointer is NULL";
        case DOPER_ERR_MISALIGNED: return "pointer is not aligned as documented";
        case DOPER_ERR_SMEM_LIMIT: return "scene does not fit the shared-memory plan of the kernel";
        case DOPER_ERR_LAUNCH: return "CUDA kernel launch failed";
        case DOPER_ERR_WORKSPACE: return "workspace too small";
        case DOPER_ERR_UNSUPPORTED: return "unsupported configuration";
        default: return "unknown doper_b200 error";
    }
}

size_t doper_scene_workspace_bytes(int64_t n_scenes, int32_t S) {
    if (n_scenes <= 0 || S <= 0) return 0;
    return scene_layout(n_scenes, S).total;
}

int doper_scene_prepare(doper_stream_t stream, int64_t n_scenes, int32_t S, const float* segments,
                        void* scene_ws) {
    if (n_scenes <= 0 || S <= 0 || n_scenes > 0x7fffffff) return DOPER_ERR_BAD_SHAPE;
    if (!segments || !scene_ws) return DOPER_ERR_NULL_PTR;
    if (!aligned(segments, 16) || !aligned(scene_ws, 256)) return DOPER_ERR_MISALIGNED;
    scene_prepare_kernel<<<(unsigned)n_scenes, 128, 0, (cudaStream_t)stream>>>(
        n_scenes, S, reinterpret_cast<const float4*>(segments), reinterpret_cast<unsigned char*>(scene_ws));
    return launch_status();
}

int doper_closest_segment(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                          const float* segments, const void* scene_ws, int32_t* idx, float* dist,
                          float* seg_out) {
    if (N < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !scene_ws || !idx || !dist) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16) || !aligned(scene_ws, 256) ||
        (seg_out && !aligned(seg_out, 16)))
        return DOPER_ERR_MISALIGNED;
    const SceneLayout L = scene_layout(1, S);
    const size_t smem = (size_t)L.S_pad * 20 + 16;
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    int G = 1;
    while (G < 32 && N * G < 148 * 256) G *= 2;
    cudaStream_t st = (cudaStream_t)stream;
    switch (G) {
        case 1: return launch_closest<1>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
        case 2: return launch_closest<2>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
        case 4: return launch_closest<4>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
        case 8: return launch_closest<8>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
        case 16: return launch_closest<16>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
        default: return launch_closest<32>(st, N, S, points, segments, scene_ws, idx, dist, seg_out, smem);
    }
}

int doper_points_inside(doper_stream_t stream, int64_t N, int32_t S, const float* points,
                        const float* segments, uint8_t* flag) {
    if (N < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (S % 3 != 0) return DOPER_ERR_UNSUPPORTED;  // jax_geometry.py:113 reshapes to (.., polygons, 3)
    if (N == 0) return DOPER_OK;
    if (!points || !segments || !flag) return DOPER_ERR_NULL_PTR;
    if (!aligned(points, 8) || !aligned(segments, 16)) return DOPER_ERR_MISALIGNED;
    const unsigned grid = (unsigned)((N + kQueryThreads - 1) / kQueryThreads);
    points_inside_kernel<<<grid, kQueryThreads, 0, (cudaStream_t)stream>>>(
        N, S, reinterpret_cast<const float2*>(points), reinterpret_cast<const float4*>(segments), flag);
    return launch_status();
}

int doper_ray_segment_points(doper_stream_t stream, int64_t R, int32_t S, const float* origins,
                             const float* directions, const float* segments, float* points) {
    if (R < 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (R == 0) return DOPER_OK;
    if (!origins || !directions || !segments || !points) return DOPER_ERR_NULL_PTR;
    if (!aligned(origins, 8) || !aligned(directions, 8) || !aligned(segments, 16) || !aligned(points, 8))
        return DOPER_ERR_MISALIGNED;
    const int64_t n = R * (int64_t)S;
    if (n > (int64_t)0x7fffffff * 256) return DOPER_ERR_UNSUPPORTED;
    ray_points_kernel<<<(unsigned)((n + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
        R, S, reinterpret_cast<const float2*>(origins), reinterpret_cast<const float2*>(directions),
        reinterpret_cast<const float4*>(segments), reinterpret_cast<float2*>(points));
    return launch_status();
}

int doper_range_sensor(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                       const float* ray_directions, const void* scene_ws, float distance_range, float* ranges,
                       float* points, int32_t* idx) {
    if (N < 0 || R <= 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (N == 0) return DOPER_OK;
    if (!positions || !ray_directions || !scene_ws || !ranges) return DOPER_ERR_NULL_PTR;
    if (!aligned(positions, 8) || !aligned(ray_directions, 8) || !aligned(scene_ws, 256) ||
        (points && !aligned(points, 8)))
        return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<false>((cudaStream_t)stream, N, R, S, positions, nullptr, ray_directions, scene_ws,
                                      distance_range, 0.0, 0.0, ranges, points, idx, nullptr);
}

int doper_agent_observations(doper_stream_t stream, int64_t N, int32_t R, int32_t S, const float* positions,
                             const float* velocities, const float* ray_directions, const void* scene_ws,
                             float distance_range, const double target[2], float* observations) {
    if (N < 0 || R <= 0 || S <= 0) return DOPER_ERR_BAD_SHAPE;
    if (N == 0) return DOPER_OK;
    if (!positions || !velocities || !ray_directions || !scene_ws || !target || !observations)
        return DOPER_ERR_NULL_PTR;
    if (!aligned(positions, 8) || !aligned(velocities, 8) || !aligned(ray_directions, 8) || !aligned(scene_ws, 256))
        return DOPER_ERR_MISALIGNED;
    return launch_range_sensor<true>((cudaStream_t)stream, N, R, S, positions, velocities, ray_directions, scene_ws,
                                     distance_range, target[0], target[1], nullptr, nullptr, nullptr, observations);
}

int32_t doper_rollout_ckpt_every(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    return resolve_K(d);
}

int32_t doper_rollout_log_ball_major(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    return log_ball_major(d) ? 1 : 0;
}

const char* doper_rollout_plan_name(const doper_rollout_desc* d, int32_t backward) {
    if (validate_desc(d) != DOPER_OK) return "invalid";
    if (backward) return use_chunked(d, resolve_K(d)) ? "rollout_bwd_chunked_kernel" : "rollout_bwd_kernel";
    const int G = auto_lanes(d);
    static const char* const tpc[] = {"rollout_fwd_tpc_kernel<2>", "rollout_fwd_tpc_kernel<4>", "rollout_fwd_tpc_kernel<8>",
                                      "rollout_fwd_tpc_kernel<16>", "rollout_fwd_tpc_kernel<32>"};
    static const char* const cull[] = {"rollout_fwd_cull_kernel<1>", "rollout_fwd_cull_kernel<2>", "rollout_fwd_cull_kernel<4>",
                                       "rollout_fwd_cull_kernel<8>", "rollout_fwd_cull_kernel<16>", "rollout_fwd_cull_kernel<32>"};
    static const char* const split[] = {"rollout_fwd_kernel<1>", "rollout_fwd_kernel<2>", "rollout_fwd_kernel<4>",
                                        "rollout_fwd_kernel<8>", "rollout_fwd_kernel<16>", "rollout_fwd_kernel<32>"};
    int lg = 0;
    while ((1 << lg) < G) ++lg;
    if (use_tpc(d)) return tpc[lg > 0 ? lg - 1 : 0];
    if (use_cull(d)) return cull[lg];
    if (use_tp(d)) return "rollout_fwd_tp_kernel";
    if (scene_layout(1, d->S).S_pad <= 256 && G >= 8 && !(d->flags & 4)) return "rollout_fwd_reg_kernel";
    return split[lg];
}

size_t doper_rollout_workspace_bytes(const doper_rollout_desc* d) {
    if (validate_desc(d) != DOPER_OK) return 0;
    const WsLayout w = ws_layout(d);
    return w.total;
}

int doper_rollout_fwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const float* x0, const float* v0, float* xT, float* vT, float* traj_x,
                      float* traj_v, float* traj_a, void* ws) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!scene_ws || !x0 || !v0 || !xT || !vT) return DOPER_ERR_NULL_PTR;
    if (!aligned(scene_ws, 256) || !aligned(x0, 8) || !aligned(v0, 8) || !aligned(xT, 8) || !aligned(vT, 8) ||
        (ws && !aligned(ws, 256)) || (traj_x && !aligned(traj_x, 8)) || (traj_v && !aligned(traj_v, 8)) ||
        (traj_a && !aligned(traj_a, 8)))
        return DOPER_ERR_MISALIGNED;
    const int G = auto_lanes(d);
    const bool tp = use_tp(d);
    const int tp_wpb = tp_warps_per_block(d);
    const size_t smem = tp ? (size_t)(d->per_env ? tp_wpb : 1) * scene_layout(1, d->S).S_pad * 20 + 16
                           : fwd_smem_bytes(d, G);
    const bool unstaged_ok = use_cull(d) && !use_tpc(d);  // the lane-split broad-phase kernel can read the scene in place
    if ((int)smem > max_optin_smem() && !unstaged_ok) return DOPER_ERR_SMEM_LIMIT;
    const int K = resolve_K(d);
    FwdParams p;
    p.hl_st = log_ball_major(d) ? 1 : d->B;  // hit log is ball-major for the time-parallel kernels
    p.hl_sb = log_ball_major(d) ? d->n_steps : 1;
    p.B = d->B; p.S = d->S; p.per_env = d->per_env ? 1 : 0; p.n_steps = d->n_steps; p.K = K;
    p.force_exact = d->flags & 1;
    p.exact_idx = (d->flags >> 17) & 1;
    p.stage = 1;
    p.c = make_consts(d);
    p.scene_ws = reinterpret_cast<const unsigned char*>(scene_ws);
    p.x0 = reinterpret_cast<const float2*>(x0); p.v0 = reinterpret_cast<const float2*>(v0);
    p.xT = reinterpret_cast<float2*>(xT); p.vT = reinterpret_cast<float2*>(vT);
    p.traj_x = reinterpret_cast<float2*>(traj_x); p.traj_v = reinterpret_cast<float2*>(traj_v);
    p.traj_a = reinterpret_cast<float2*>(traj_a);
    p.ckpt = nullptr; p.hitlog = nullptr; p.hitstate = nullptr;
    if (ws) {
        const WsLayout w = ws_layout(d);
        unsigned char* base = reinterpret_cast<unsigned char*>(ws);
        p.ckpt = reinterpret_cast<float4*>(base);
        p.hitlog = reinterpret_cast<int32_t*>(base + w.log_off);
        if (w.state_off) p.hitstate = reinterpret_cast<float4*>(base + w.state_off);
    }
    cudaStream_t st = (cudaStream_t)stream;
    if (tp) {
        if (smem > 48 * 1024 &&
            cudaFuncSetAttribute(rollout_fwd_tp_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
                cudaSuccess)
            return DOPER_ERR_SMEM_LIMIT;
        const unsigned grid = (unsigned)((d->B + tp_wpb - 1) / tp_wpb);
        rollout_fwd_tp_kernel<<<grid, tp_wpb * 32, smem, st>>>(p);
        return launch_status();
    }
    if (use_tpc(d)) {
        switch (G) {
            case 1: return DOPER_ERR_UNSUPPORTED;
            case 2: return launch_fwd_tpc<2>(st, p, smem);
            case 4: return launch_fwd_tpc<4>(st, p, smem);
            case 8: return launch_fwd_tpc<8>(st, p, smem);
            case 16: return launch_fwd_tpc<16>(st, p, smem);
            default: return launch_fwd_tpc<32>(st, p, smem);
        }
    }
    if (use_cull(d)) {
        switch (G) {
            case 1: return launch_fwd_cull<1>(st, p, smem);
            case 2: return launch_fwd_cull<2>(st, p, smem);
            case 4: return launch_fwd_cull<4>(st, p, smem);
            case 8: return launch_fwd_cull<8>(st, p, smem);
            case 16: return launch_fwd_cull<16>(st, p, smem);
            default: return launch_fwd_cull<32>(st, p, smem);
        }
    }
    const int S_pad = scene_layout(1, d->S).S_pad;
    if (S_pad <= 256 && G >= 8 && !(d->flags & 4)) {  // flags bit 2: force the shared-memory kernel
        switch (G) {
            case 8: return dispatch_fwd_reg<8>(st, p, S_pad);
            case 16: return dispatch_fwd_reg<16>(st, p, S_pad);
            default: return dispatch_fwd_reg<32>(st, p, S_pad);
        }
    }
    switch (G) {
        case 1: return launch_fwd<1>(st, p, smem);
        case 2: return launch_fwd<2>(st, p, smem);
        case 4: return launch_fwd<4>(st, p, smem);
        case 8: return launch_fwd<8>(st, p, smem);
        case 16: return launch_fwd<16>(st, p, smem);
        default: return launch_fwd<32>(st, p, smem);
    }
}

int doper_rollout_bwd(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                      const void* ws, const float* xT_bar, const float* vT_bar, float* x0_bar,
                      float* v0_bar) {
    int rc = validate_desc(d);
    if (rc) return rc;
    if (!scene_ws || !ws || !xT_bar || !v0_bar) return DOPER_ERR_NULL_PTR;
    if (!aligned(scene_ws, 256) || !aligned(ws, 256) || !aligned(xT_bar, 8) || !aligned(v0_bar, 8) ||
        (vT_bar && !aligned(vT_bar, 8)) || (x0_bar && !aligned(x0_bar, 8)))
        return DOPER_ERR_MISALIGNED;
    const int K = resolve_K(d);
    BwdParams p;
    p.B = d->B; p.S = d->S; p.per_env = d->per_env ? 1 : 0; p.n_steps = d->n_steps; p.K = K;
    p.c = make_consts(d);
    p.scene_ws = reinterpret_cast<const unsigned char*>(scene_ws);
    const size_t n_ckpt = (size_t)(d->n_steps + K - 1) / K;
    const WsLayout w = ws_layout(d);
    const unsigned char* base = reinterpret_cast<const unsigned char*>(ws);
    p.ckpt = reinterpret_cast<const float4*>(base);
    p.hitlog = reinterpret_cast<const int32_t*>(base + w.log_off);
    p.hitstate = w.state_off ? reinterpret_cast<const float4*>(base + w.state_off) : nullptr;
    p.hl_st = log_ball_major(d) ? 1 : d->B;
    p.hl_sb = log_ball_major(d) ? d->n_steps : 1;
    p.xT_bar = reinterpret_cast<const float2*>(xT_bar);
    p.vT_bar = reinterpret_cast<const float2*>(vT_bar);
    p.x0_bar = reinterpret_cast<float2*>(x0_bar);
    p.v0_bar = reinterpret_cast<float2*>(v0_bar);
    p.deferred = nullptr; p.defer_after = 0;
    if (d->n_steps == 0) return DOPER_ERR_BAD_SHAPE;
    if (use_chunked(d, K)) {
        int C_pad = 2;
        while (C_pad < (int)n_ckpt) C_pad *= 2;
        const int bpb = kBwdThreads / C_pad;
        const unsigned grid = (unsigned)((d->B + bpb - 1) / bpb);
        rollout_bwd_chunked_kernel<<<grid, kBwdThreads, 0, (cudaStream_t)stream>>>(p, C_pad, K, nullptr);
        return launch_status();
    }
    // hybrid: balls in sliding contact leave the sequential kernel and are redone 32 chunks in parallel
    // (needs the hit-state array; flags bit 1 keeps everything sequential = bit-exact vs the oracle order)
    const bool hybrid = p.hitstate && !(d->flags & 2) && d->n_steps >= 256;
    if (hybrid) {
        p.deferred = reinterpret_cast<int*>(const_cast<unsigned char*>(base) + w.defer_off);
        p.defer_after = 128;
        if (cudaMemsetAsync(p.deferred, 0, sizeof(int), (cudaStream_t)stream) != cudaSuccess) return DOPER_ERR_LAUNCH;
    }
    const size_t smem = p.hitstate ? 0 : (size_t)K * kBwdThreads * sizeof(float4);
    if ((int)smem > max_optin_smem()) return DOPER_ERR_SMEM_LIMIT;
    if (smem > 48 * 1024 &&
        cudaFuncSetAttribute(rollout_bwd_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem) !=
            cudaSuccess)
        return DOPER_ERR_SMEM_LIMIT;
    const unsigned grid = (unsigned)((d->B + kBwdThreads - 1) / kBwdThreads);
    rollout_bwd_kernel<<<grid, kBwdThreads, smem, (cudaStream_t)stream>>>(p);
    if (hybrid && launch_status() == DOPER_OK) {
        const int chunk_len = (d->n_steps + 31) / 32;
        rollout_bwd_chunked_kernel<<<kDeferCap / (kBwdThreads / 32), kBwdThreads, 0, (cudaStream_t)stream>>>(
            p, 32, chunk_len, p.deferred);
    }
    return launch_status();
}

int doper_l1_loss(doper_stream_t stream, int64_t B, const float* xT, const float target[2],
                  float* loss_per_ball, float* loss_sum, float* xT_bar) {
    if (B <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!xT || !target || !loss_sum) return DOPER_ERR_NULL_PTR;
    if (!aligned(xT, 8) || (xT_bar && !aligned(xT_bar, 8))) return DOPER_ERR_MISALIGNED;
    const unsigned grid = 1 + (unsigned)((B + kLossThreads - 1) / kLossThreads);
    l1_loss_kernel<<<grid, kLossThreads, 0, (cudaStream_t)stream>>>(
        B, reinterpret_cast<const float2*>(xT), target[0], target[1], loss_per_ball, loss_sum,
        reinterpret_cast<float2*>(xT_bar));
    return launch_status();
}

int doper_value_and_grad(doper_stream_t stream, const doper_rollout_desc* d, const void* scene_ws,
                         const float* x0, const float* v0, const float target[2], void* ws,
                         float* scratch, float* loss_sum, float* loss_per_ball, float* xT,
                         float* vT, float* v0_bar, float* x0_bar) {
    if (!ws || !scratch || !target) return DOPER_ERR_NULL_PTR;
    int rc = doper_rollout_fwd(stream, d, scene_ws, x0, v0, xT, vT, nullptr, nullptr, nullptr, ws);
    if (rc) return rc;
    rc = doper_l1_loss(stream, d->B, xT, target, loss_per_ball, loss_sum, scratch);
    if (rc) return rc;
    return doper_rollout_bwd(stream, d, scene_ws, ws, scratch, nullptr, x0_bar, v0_bar);
}

int doper_fp32_probe(doper_stream_t stream, int32_t blocks, int32_t threads, int32_t iters,
                     int32_t use_fma, float* out) {
    if (blocks <= 0 || threads <= 0 || threads > 1024 || iters <= 0) return DOPER_ERR_BAD_SHAPE;
    if (!out) return DOPER_ERR_NULL_PTR;
    fp32_probe_kernel<<<blocks, threads, 0, (cudaStream_t)stream>>>(iters, use_fma, out);
    return launch_status();
}

}  // extern "C"
