// peer.cuh: one-shot all-reduce over NVLink peer memory — part of the doper_b200 CUDA translation unit.
//
// The one exchange of the data-parallel rollout trainer is tiny (the controller's 6,602 fp32 gradients +
// the loss = 26 KB per step) and sits on the step's critical path, so its cost is pure latency: through
// NCCL it measured +19 us per step at 2 GPUs and +80 us at 8.  Here every rank owns one cudaMalloc'ed
// buffer (flags + two data slots), opens its peers' buffers through CUDA IPC (NVLink / NVSwitch peer
// mappings), and ONE small kernel per rank does the whole exchange:
//   1. publish: a system-scope release store of this step's sequence number into every peer's flag row
//      (the data was written into this rank's slot by earlier work on the same stream),
//   2. wait: acquire-spin until every peer's sequence number has arrived (bounded: a missing peer makes
//      the kernel give up after ~2 s and raise the status word instead of hanging the GPU),
//   3. reduce: read the slot of every rank over NVLink (L1-bypassing loads, all in flight together: one
//      round trip) and sum in rank order — the same bits on every rank.
// Two slots, alternating by step, make a trailing barrier unnecessary: a rank can be at most one step
// ahead of the slowest one (it needs everybody's flag of step s+1 before it can overwrite slot s).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace doper {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerFlagBytes = 256;  // [2 slots][16 ranks] uint32 = 128 B, padded

struct PeerPtrs {
    unsigned char* base[kPeerMaxRanks];
};

__host__ __device__ inline size_t peer_buffer_bytes(int n) {
    const size_t n_pad = ((size_t)n + 63) / 64 * 64;
    return kPeerFlagBytes + 2 * n_pad * sizeof(float);
}

__device__ __forceinline__ void st_release_sys(uint32_t* p, uint32_t v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ uint32_t ld_acquire_sys(const uint32_t* p) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
// L1-bypassing 16-byte load (peer lines must not be served from a stale L1); unlike a volatile load the
// compiler may keep several of them in flight, which is the whole point: one NVLink round trip, not `world`
__device__ __forceinline__ float4 ld_cv_f4(const float4* p) { return __ldcv(p); }

constexpr int kPeerThreads = 256;

// grid = ceil(n / 4 / kPeerThreads) blocks: every thread owns ONE float4 of the vector and reads it from
// every rank; every block waits for the flags itself (block 0 publishes this rank's)
__global__ void __launch_bounds__(kPeerThreads) allreduce_oneshot_kernel(PeerPtrs P, int rank, int world, int n,
                                                                        uint32_t seq, float* __restrict__ out,
                                                                        int* __restrict__ status) {
    const int tid = threadIdx.x;
    const int slot = (int)(seq & 1u);
    const size_t n_pad = ((size_t)n + 63) / 64 * 64;
    __shared__ int s_fail;
    if (tid == 0) s_fail = 0;
    __threadfence_system();  // this rank's slot (written before this kernel, on this stream) before its flags
    __syncthreads();
    if (tid < world) {
        if (blockIdx.x == 0)
            st_release_sys(reinterpret_cast<uint32_t*>(P.base[tid]) + slot * kPeerMaxRanks + rank, seq);
        const uint32_t* mine = reinterpret_cast<const uint32_t*>(P.base[rank]) + slot * kPeerMaxRanks + tid;
        const long long t0 = clock64();
        while (ld_acquire_sys(mine) != seq) {
            if (clock64() - t0 > 4000000000ll) { s_fail = 1; break; }  // ~2 s at 1.9 GHz: a peer never arrived
        }
    }
    __syncthreads();
    if (s_fail) {
        if (tid == 0 && status) *status = 1;
        return;
    }
    const int i4 = blockIdx.x * kPeerThreads + tid;  // float4 index
    if (4 * i4 >= n) return;
    float4 v[kPeerMaxRanks];
#pragma unroll
    for (int r = 0; r < kPeerMaxRanks; ++r)
        if (r < world) v[r] = ld_cv_f4(reinterpret_cast<const float4*>(P.base[r] + kPeerFlagBytes) + slot * (n_pad / 4) + i4);
    float4 acc = make_float4(0.f, 0.f, 0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kPeerMaxRanks; ++r)
        if (r < world) { acc.x += v[r].x; acc.y += v[r].y; acc.z += v[r].z; acc.w += v[r].w; }  // rank order: same bits everywhere
    const float a[4] = {acc.x, acc.y, acc.z, acc.w};
#pragma unroll
    for (int k = 0; k < 4; ++k)
        if (4 * i4 + k < n) out[4 * i4 + k] = a[k];
}

}  // namespace doper
