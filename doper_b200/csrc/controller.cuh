// controller.cuh: the trainer's glue around the rollout as three kernels — part of the
// doper_b200 CUDA translation unit (included by kernels.cu).
//
// Reference: doper/networks/controllers.py:12-18 (Linear(d_in, h) - ReLU - Linear(h, 2h) - ReLU - Linear(2h, d_out),
// 25-50-100-2 = 6,602 parameters in the shipped config) and doper/trainers/ball_trainer.py:57-76:
//     impulse = controller(observation);  v0 = velocity_init + impulse / mass;  ...rollout...;
//     impulse.backward(dL/dv0)  (8 actions accumulate into .grad);  clip_grad_norm_(5);  Adam step.
// In torch that is ~25 launches per action (three GEMMs with 6 KB operands, bias adds, ReLUs, their
// backward, the division) and ~10 per optimiser step, all latency bound; at 4,096 balls they cost more than the
// rollout itself (profiles/r1_trainer_iteration_graph.json).  Here:
//
//   controller_fwd_kernel   per block: the 26 KB of parameters -> shared memory (transposed: lane = neuron reads are
//                           conflict-free), 32 balls at a time through the three layers, v0 written directly.
//   controller_bwd_kernel   recomputes the activations of its 32 balls, back-propagates dL/dv0 / mass, and writes the
//                           block's partial parameter gradient [6,602] to scratch; controller_grad_reduce_kernel
//                           adds the partials in block order into the flat gradient buffer (deterministic: the
//                           same bits on every run and rank) — that buffer IS the multi-GPU exchange buffer
//                           (gradients ‖ loss), so nothing is packed or copied before doper_allreduce_oneshot.
//   adam_clip_kernel        clip_grad_norm_ (total norm, coefficient clamped to 1) + Adam (torch.optim.Adam defaults'
//                           update rule, bias correction from the device-side step counter) + zeroing of the gradient
//                           buffer for the next iteration, one block.
//
// fp32 throughout like torch; products are accumulated in index order with fused multiply-adds (torch's GEMM order is
// unspecified: parity with the torch reference is within rounding, tests/test_controller_gpu.py).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace doper {

constexpr int kCtlThreads = 256;
constexpr int kCtlBalls = 32;  // balls per block
constexpr int kCtlMaxIn = 64, kCtlMaxHidden = 128, kCtlMaxOut = 8;

struct CtlDims {
    int d_in, h1, h2, d_out;  // layer widths: d_in -> h1 -> h2 -> d_out
    __host__ __device__ int n_params() const { return d_in * h1 + h1 + h1 * h2 + h2 + h2 * d_out + d_out; }
    // offsets into the flat parameter vector, torch order: W1 [h1][d_in], b1, W2 [h2][h1], b2, W3 [d_out][h2], b3
    __host__ __device__ int oW1() const { return 0; }
    __host__ __device__ int ob1() const { return d_in * h1; }
    __host__ __device__ int oW2() const { return ob1() + h1; }
    __host__ __device__ int ob2() const { return oW2() + h1 * h2; }
    __host__ __device__ int oW3() const { return ob2() + h2; }
    __host__ __device__ int ob3() const { return oW3() + h2 * d_out; }
};

// shared-memory layout (floats): parameters transposed per layer ([in][out]) + biases, then activations
struct CtlSmem {
    float *W1t, *b1, *W2t, *b2, *W3t, *b3, *x, *a1, *a2, *y;
    __device__ CtlSmem(float* base, const CtlDims& d) {
        W1t = base; b1 = W1t + d.d_in * d.h1; W2t = b1 + d.h1; b2 = W2t + d.h1 * d.h2; W3t = b2 + d.h2; b3 = W3t + d.h2 * d.d_out;
        x = b3 + d.d_out; a1 = x + kCtlBalls * d.d_in; a2 = a1 + kCtlBalls * d.h1; y = a2 + kCtlBalls * d.h2;
    }
    static __host__ __device__ size_t floats(const CtlDims& d) {
        return (size_t)d.n_params() + (size_t)kCtlBalls * (d.d_in + d.h1 + d.h2 + d.d_out);
    }
};

__device__ __forceinline__ void ctl_load_params(const CtlSmem& s, const CtlDims& d, const float* __restrict__ params) {
    const int tid = threadIdx.x;
    for (int e = tid; e < d.d_in * d.h1; e += kCtlThreads) { const int n = e / d.d_in, k = e - n * d.d_in; s.W1t[k * d.h1 + n] = params[d.oW1() + e]; }
    for (int e = tid; e < d.h1 * d.h2; e += kCtlThreads) { const int n = e / d.h1, k = e - n * d.h1; s.W2t[k * d.h2 + n] = params[d.oW2() + e]; }
    for (int e = tid; e < d.h2 * d.d_out; e += kCtlThreads) { const int n = e / d.h2, k = e - n * d.h2; s.W3t[k * d.d_out + n] = params[d.oW3() + e]; }
    for (int e = tid; e < d.h1; e += kCtlThreads) s.b1[e] = params[d.ob1() + e];
    for (int e = tid; e < d.h2; e += kCtlThreads) s.b2[e] = params[d.ob2() + e];
    for (int e = tid; e < d.d_out; e += kCtlThreads) s.b3[e] = params[d.ob3() + e];
}

// out[o] = post(init(o) + sum_k A(o, k) * Bm(o, k)) for o in [0, total): four outputs per thread in flight (four
// independent accumulator chains: the per-output dot products are short dependent FMA chains, and a block is
// latency bound)
template <typename Init, typename A, typename Bm, typename Post>
__device__ __forceinline__ void ctl_dots(int total, int K, Init init, A a, Bm bm, Post post) {
    for (int o0 = threadIdx.x; o0 < total; o0 += 4 * kCtlThreads) {
        int o[4];
        float acc[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) { o[u] = min(o0 + u * kCtlThreads, total - 1); acc[u] = init(o[u]); }
        for (int k = 0; k < K; ++k) {
#pragma unroll
            for (int u = 0; u < 4; ++u) acc[u] = __fmaf_rn(a(o[u], k), bm(o[u], k), acc[u]);
        }
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (o0 + u * kCtlThreads < total) post(o[u], acc[u]);
    }
}

// activations of the block's balls: a1 = relu(W1 x + b1), a2 = relu(W2 a1 + b2), y = W3 a2 + b3
__device__ __forceinline__ void ctl_forward(const CtlSmem& s, const CtlDims& d, int nb) {
    ctl_dots(nb * d.h1, d.d_in, [&](int o) { return s.b1[o % d.h1]; },
             [&](int o, int k) { return s.W1t[k * d.h1 + o % d.h1]; }, [&](int o, int k) { return s.x[(o / d.h1) * d.d_in + k]; },
             [&](int o, float v) { s.a1[o] = fmaxf(v, 0.0f); });
    __syncthreads();
    ctl_dots(nb * d.h2, d.h1, [&](int o) { return s.b2[o % d.h2]; },
             [&](int o, int k) { return s.W2t[k * d.h2 + o % d.h2]; }, [&](int o, int k) { return s.a1[(o / d.h2) * d.h1 + k]; },
             [&](int o, float v) { s.a2[o] = fmaxf(v, 0.0f); });
    __syncthreads();
    ctl_dots(nb * d.d_out, d.h2, [&](int o) { return s.b3[o % d.d_out]; },
             [&](int o, int k) { return s.W3t[k * d.d_out + o % d.d_out]; }, [&](int o, int k) { return s.a2[(o / d.d_out) * d.h2 + k]; },
             [&](int o, float v) { s.y[o] = v; });
    __syncthreads();
}

// impulse = controller(obs); v0 = v_init + impulse / mass (ball_trainer.py:57-64; d_out = 2)
__global__ void __launch_bounds__(kCtlThreads) controller_fwd_kernel(CtlDims d, int64_t B, const float* __restrict__ params,
                                                                     const float* __restrict__ obs, const float2* __restrict__ v_init,
                                                                     float mass, float* __restrict__ impulse, float2* __restrict__ v0) {
    extern __shared__ __align__(16) float ctl_smem[];
    const CtlSmem s(ctl_smem, d);
    ctl_load_params(s, d, params);
    const int64_t first = (int64_t)blockIdx.x * kCtlBalls;
    const int nb = (int)min((int64_t)kCtlBalls, B - first);
    for (int e = threadIdx.x; e < nb * d.d_in; e += kCtlThreads) s.x[e] = obs[first * d.d_in + e];
    __syncthreads();
    ctl_forward(s, d, nb);
    for (int o = threadIdx.x; o < nb * d.d_out; o += kCtlThreads) {
        if (impulse) impulse[first * d.d_out + o] = s.y[o];
    }
    if (v0 && threadIdx.x < nb) {
        const int b = threadIdx.x;
        const float2 vi = v_init[first + b];
        v0[first + b] = make_float2(vi.x + s.y[b * d.d_out] / mass, vi.y + s.y[b * d.d_out + 1] / mass);
    }
}

// Block partial of dL/dparams for the block's balls, given g = dL/dv0 (dL/dimpulse = g / mass).
// partial [gridDim.x][n_params] (torch parameter order).
__global__ void __launch_bounds__(kCtlThreads) controller_bwd_kernel(CtlDims d, int64_t B, const float* __restrict__ params,
                                                                     const float* __restrict__ obs, const float2* __restrict__ g_v0,
                                                                     float mass, float* __restrict__ partial) {
    extern __shared__ __align__(16) float ctl_smem[];
    const CtlSmem s(ctl_smem, d);
    float* dz2 = ctl_smem + CtlSmem::floats(d);       // [nb][h2]
    float* dz1 = dz2 + kCtlBalls * d.h2;              // [nb][h1]
    float* dy = dz1 + kCtlBalls * d.h1;               // [nb][d_out]
    ctl_load_params(s, d, params);
    const int tid = threadIdx.x;
    const int64_t first = (int64_t)blockIdx.x * kCtlBalls;
    const int nb = (int)min((int64_t)kCtlBalls, B - first);
    for (int e = tid; e < nb * d.d_in; e += kCtlThreads) s.x[e] = obs[first * d.d_in + e];
    for (int e = tid; e < nb; e += kCtlThreads) {
        const float2 g = g_v0[first + e];
        dy[e * d.d_out] = g.x / mass;  // d(v_init + impulse / mass) / d impulse
        dy[e * d.d_out + 1] = g.y / mass;
        for (int o = 2; o < d.d_out; ++o) dy[e * d.d_out + o] = 0.0f;
    }
    __syncthreads();
    ctl_forward(s, d, nb);
    // dz2 = (W3^T dy) * [a2 > 0]
    ctl_dots(nb * d.h2, d.d_out, [&](int) { return 0.0f; },
             [&](int o, int n) { return s.W3t[(o % d.h2) * d.d_out + n]; }, [&](int o, int n) { return dy[(o / d.h2) * d.d_out + n]; },
             [&](int o, float v) { dz2[o] = s.a2[o] > 0.0f ? v : 0.0f; });
    __syncthreads();
    // dz1 = (W2^T dz2) * [a1 > 0]
    ctl_dots(nb * d.h1, d.h2, [&](int) { return 0.0f; },
             [&](int o, int n) { return s.W2t[(o % d.h1) * d.h2 + n]; }, [&](int o, int n) { return dz2[(o / d.h1) * d.h2 + n]; },
             [&](int o, float v) { dz1[o] = s.a1[o] > 0.0f ? v : 0.0f; });
    __syncthreads();
    float* out = partial + (size_t)blockIdx.x * d.n_params();
    // weight gradients: entry (n, k) = sum over the block's balls of delta[b][n] * input[b][k]
    ctl_dots(d.h1 * d.d_in, nb, [&](int) { return 0.0f; },
             [&](int e, int b) { return dz1[b * d.h1 + e / d.d_in]; }, [&](int e, int b) { return s.x[b * d.d_in + e % d.d_in]; },
             [&](int e, float v) { out[d.oW1() + e] = v; });
    ctl_dots(d.h2 * d.h1, nb, [&](int) { return 0.0f; },
             [&](int e, int b) { return dz2[b * d.h2 + e / d.h1]; }, [&](int e, int b) { return s.a1[b * d.h1 + e % d.h1]; },
             [&](int e, float v) { out[d.oW2() + e] = v; });
    ctl_dots(d.d_out * d.h2, nb, [&](int) { return 0.0f; },
             [&](int e, int b) { return dy[b * d.d_out + e / d.h2]; }, [&](int e, int b) { return s.a2[b * d.h2 + e % d.h2]; },
             [&](int e, float v) { out[d.oW3() + e] = v; });
    for (int n = tid; n < d.h1; n += kCtlThreads) { float acc = 0.0f; for (int b = 0; b < nb; ++b) acc += dz1[b * d.h1 + n]; out[d.ob1() + n] = acc; }
    for (int n = tid; n < d.h2; n += kCtlThreads) { float acc = 0.0f; for (int b = 0; b < nb; ++b) acc += dz2[b * d.h2 + n]; out[d.ob2() + n] = acc; }
    for (int n = tid; n < d.d_out; n += kCtlThreads) { float acc = 0.0f; for (int b = 0; b < nb; ++b) acc += dy[b * d.d_out + n]; out[d.ob3() + n] = acc; }
}

// grad[e] += sum over blocks (in block order) of partial[block][e]
__global__ void __launch_bounds__(256) controller_grad_reduce_kernel(int n_params, int n_blocks, const float* __restrict__ partial,
                                                                     float* __restrict__ grad) {
    const int e = blockIdx.x * 256 + threadIdx.x;
    if (e >= n_params) return;
    float acc = 0.0f;
    int b = 0;
    for (; b + 4 <= n_blocks; b += 4) {  // four loads in flight; the additions stay in block order
        const float p0 = partial[(size_t)b * n_params + e], p1 = partial[(size_t)(b + 1) * n_params + e];
        const float p2 = partial[(size_t)(b + 2) * n_params + e], p3 = partial[(size_t)(b + 3) * n_params + e];
        acc = (((acc + p0) + p1) + p2) + p3;
    }
    for (; b < n_blocks; ++b) acc += partial[(size_t)b * n_params + e];
    grad[e] += acc;
}

// torch.nn.utils.clip_grad_norm_(params, max_norm) followed by torch.optim.Adam.step() (no weight decay, no
// amsgrad) on the flat vectors, then grad <- 0.  state = {exp_avg [n], exp_avg_sq [n]}, step: device float counter
// (torch keeps the step as a float tensor too).  One block.
__global__ void __launch_bounds__(1024) adam_clip_kernel(int n, float* __restrict__ params, float* __restrict__ grad,
                                                         float* __restrict__ exp_avg, float* __restrict__ exp_avg_sq,
                                                         float* __restrict__ step, float max_norm, float lr, float beta1, float beta2,
                                                         float eps, float* __restrict__ total_norm_out) {
    __shared__ float red[32];
    __shared__ float coef_s;
    const int tid = threadIdx.x;
    float ss = 0.0f;
    for (int e = tid; e < n; e += 1024) { const float g = grad[e]; ss = __fmaf_rn(g, g, ss); }
    for (int off = 16; off > 0; off >>= 1) ss += __shfl_xor_sync(0xffffffffu, ss, off);
    if ((tid & 31) == 0) red[tid >> 5] = ss;
    __syncthreads();
    if (tid < 32) {
        float v = red[tid];
        for (int off = 16; off > 0; off >>= 1) v += __shfl_xor_sync(0xffffffffu, v, off);
        if (tid == 0) {
            const float total = sqrtf(v);
            float coef = max_norm / (total + 1.0e-6f);  // clip_grad_norm_: clip_coef = max_norm / (total_norm + 1e-6), clamped to 1
            coef = coef > 1.0f ? 1.0f : coef;
            coef_s = max_norm > 0.0f ? coef : 1.0f;
            if (total_norm_out) *total_norm_out = total;
        }
    }
    __syncthreads();
    const float coef = coef_s;
    const float t = step[0] + 1.0f;
    const float bc1 = 1.0f - powf(beta1, t), bc2 = 1.0f - powf(beta2, t);
    const float step_size = lr / bc1, bc2_sqrt = sqrtf(bc2);
    for (int e = tid; e < n; e += 1024) {
        const float g = grad[e] * coef;
        const float m = exp_avg[e] + (g - exp_avg[e]) * (1.0f - beta1);            // lerp_(grad, 1 - beta1)
        const float v = exp_avg_sq[e] * beta2 + (g * g) * (1.0f - beta2);          // mul_(beta2).addcmul_(grad, grad, 1 - beta2)
        exp_avg[e] = m; exp_avg_sq[e] = v;
        const float denom = sqrtf(v) / bc2_sqrt + eps;
        params[e] = params[e] - step_size * (m / denom);                           // addcdiv_(exp_avg, denom, value=-step_size)
        grad[e] = 0.0f;
    }
    __syncthreads();
    if (tid == 0) step[0] = t;
}

}  // namespace doper
