// peer.cuh: one-shot all-reduce over NVLink peer memory — part of the doper_b200 CUDA translation unit.
//
// The one exchange of the data-parallel rollout trainer is tiny (the controller's 6,602 fp32 gradients +
// the loss = 26 KB per step) and sits on the step's critical path, so its cost is pure latency: through
// NCCL it measured 25 us per call at 2 GPUs and 31 us at 8.  Here every rank owns one cudaMalloc'ed
// buffer, opens its peers' buffers through CUDA IPC (NVLink / NVSwitch peer mappings), and ONE small
// kernel per rank does the whole exchange, input copy included:
//   1. push: every thread takes one float4 of this rank's input and stores it into EVERY rank's buffer
//      (its own included) as two 16-byte words {value, seq, value, seq} — the step's sequence number
//      travels inside the same store as the data, so there is no separate flag, no fence and no second
//      NVLink trip: a word whose two sequence fields both equal this step's number is complete (each
//      8-byte half is written atomically);
//   2. poll + reduce: the same thread spins on ITS words of every source rank in its own (local) buffer
//      until they carry this step's number, then sums them in rank order — the same bits on every rank.
// The sequence number lives in the buffer header on the device (the last block to finish bumps it), so
// the launch arguments never change and the call can be captured in a CUDA graph.  Two slots alternate
// by step parity, which makes a trailing barrier unnecessary: a rank can be at most one step ahead of
// the slowest one (it needs everybody's words of step s+1 before it can start step s+2, and a peer
// sends those only after its own step s has finished reading).  A peer that never arrives makes the
// kernel give up after `timeout_cycles` SM clocks (default ~60 s), raise the status word and fill its part
// of `out` with NaN (a visible failure: the caller must not step the optimiser on a partial sum) instead of
// hanging the GPU.
//
// MEMORY-MODEL ASSUMPTION (not guaranteed by the PTX memory model; holds on the NVLink 5 / NVSwitch
// systems this was measured on, and is what the 2-rank self-check of bench.py --gpus N verifies against
// NCCL on every run): a 16-byte st.volatile.global.v4 to peer memory becomes visible to the peer's
// ld.volatile.global.v4 in units of at least 8 aligned bytes, i.e. a {value, seq} half is never torn.  The
// PTX model only guarantees single-copy atomicity per 32-bit element of a vector access; hardware splits
// vector stores at 8- / 16-byte granularity at most.  A platform that tears a half would show up as a
// wrong sum with matching sequence numbers; the fallback there is the NCCL transport (`--nccl`).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace doper {

constexpr int kPeerMaxRanks = 16;
constexpr int kPeerHeaderBytes = 256;  // [0] sequence number of the last finished call, [1] finished-block counter
constexpr int kPeerThreads = 128;

struct PeerPtrs {
    unsigned char* base[kPeerMaxRanks];
};

__host__ __device__ inline size_t peer_n4(int n) { return ((size_t)n + 3) / 4; }  // float4 groups
// header + [2 slots][kPeerMaxRanks sources][n4 groups][2 words of 16 B]
__host__ __device__ inline size_t peer_buffer_bytes(int n) {
    return kPeerHeaderBytes + 2 * (size_t)kPeerMaxRanks * peer_n4(n) * 32;
}

__device__ __forceinline__ void st_volatile_u4(uint4* p, uint4 v) {
    asm volatile("st.volatile.global.v4.u32 [%0], {%1, %2, %3, %4};" ::"l"(p), "r"(v.x), "r"(v.y), "r"(v.z), "r"(v.w)
                 : "memory");
}
__device__ __forceinline__ uint4 ld_volatile_u4(const uint4* p) {
    uint4 v;
    asm volatile("ld.volatile.global.v4.u32 {%0, %1, %2, %3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p)
                 : "memory");
    return v;
}

// grid = ceil(n4 / kPeerThreads) blocks; thread = one float4 group of the vector
__global__ void __launch_bounds__(kPeerThreads) allreduce_oneshot_kernel(PeerPtrs P, int rank, int world, int n,
                                                                        const float* __restrict__ in,
                                                                        float* __restrict__ out,
                                                                        int* __restrict__ status, long long timeout_cycles) {
    uint32_t* hdr = reinterpret_cast<uint32_t*>(P.base[rank]);
    const uint32_t seq = *reinterpret_cast<volatile uint32_t*>(hdr) + 1u;  // bumped only after every block has read it
    const size_t n4 = peer_n4(n);
    const size_t g = (size_t)blockIdx.x * kPeerThreads + threadIdx.x;
    const size_t slot_words = (size_t)kPeerMaxRanks * n4 * 2;  // 16-byte words per slot
    bool failed = false;
    if (g < n4) {
        float x[4];
#pragma unroll
        for (int k = 0; k < 4; ++k) x[k] = (4 * g + k < (size_t)n) ? in[4 * g + k] : 0.0f;
        const uint4 w0 = make_uint4(__float_as_uint(x[0]), seq, __float_as_uint(x[1]), seq);
        const uint4 w1 = make_uint4(__float_as_uint(x[2]), seq, __float_as_uint(x[3]), seq);
        const size_t mine = (seq & 1u) * slot_words + ((size_t)rank * n4 + g) * 2;
        for (int r = 0; r < world; ++r) {  // start with the next rank so the NVSwitch ports are hit evenly
            int t = rank + 1 + r;
            if (t >= world) t -= world;
            uint4* dst = reinterpret_cast<uint4*>(P.base[t] + kPeerHeaderBytes) + mine;
            st_volatile_u4(dst, w0);
            st_volatile_u4(dst + 1, w1);
        }
        const uint4* src = reinterpret_cast<const uint4*>(P.base[rank] + kPeerHeaderBytes) + (seq & 1u) * slot_words + g * 2;
        // poll: the words of every still-missing source rank are loaded together (one L2 round trip per
        // pass, not one per rank), then checked
        uint4 a[kPeerMaxRanks], b[kPeerMaxRanks];
        unsigned pending = world >= 32 ? 0xffffffffu : ((1u << world) - 1u);
        const long long t0 = clock64();
        while (pending) {
#pragma unroll
            for (int r = 0; r < kPeerMaxRanks; ++r)
                if ((pending >> r) & 1u) {
                    const uint4* q = src + (size_t)r * n4 * 2;
                    a[r] = ld_volatile_u4(q);
                    b[r] = ld_volatile_u4(q + 1);
                }
#pragma unroll
            for (int r = 0; r < kPeerMaxRanks; ++r)
                if (((pending >> r) & 1u) && a[r].y == seq && a[r].w == seq && b[r].y == seq && b[r].w == seq)
                    pending &= ~(1u << r);
            if (pending && clock64() - t0 > timeout_cycles) { failed = true; break; }  // a peer never arrived
        }
        float acc[4] = {0.f, 0.f, 0.f, 0.f};
#pragma unroll
        for (int r = 0; r < kPeerMaxRanks; ++r)  // rank order: the same bits everywhere
            if (r < world) {
                acc[0] += __uint_as_float(a[r].x); acc[1] += __uint_as_float(a[r].z);
                acc[2] += __uint_as_float(b[r].x); acc[3] += __uint_as_float(b[r].z);
            }
#pragma unroll
        for (int k = 0; k < 4; ++k)
            if (4 * g + k < (size_t)n) out[4 * g + k] = failed ? __int_as_float(0x7fc00000) : acc[k];
    }
    if (failed && status) *status = 1;
    // the last block to get here publishes the new sequence number (even after a failure: the ranks stay in step)
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned done = atomicAdd(hdr + 1, 1u);
        if (done == gridDim.x - 1) {
            hdr[1] = 0u;
            *reinterpret_cast<volatile uint32_t*>(hdr) = seq;
        }
    }
}

}  // namespace doper
