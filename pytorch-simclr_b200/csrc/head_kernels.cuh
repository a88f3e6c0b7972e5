// Tail of the projection head fused into the loss: the reference produces the embeddings as
//     z = BatchNorm1d(128)(Linear(2048 -> 128, bias=False)(.))            models/simclr.py:38-39
// and hands them to objective.contrastive_loss.  With the pre-BatchNorm activations u as the loss's input, the
// BatchNorm apply runs inside the prepare kernel (z is never written to memory: u -> BN -> normalise -> bf16 operand in
// registers) and, on the way back, the backward finalize kernel produces dL/dz and the per-column reductions BatchNorm's
// backward needs; bn_backward_kernel then turns dL/dz into dL/du in place and emits dL/dgamma, dL/dbeta.
//
// BatchNorm state of one call, f32 [2 views][kBnPlanes][Dpad] (the two views go through the module in two separate
// forward calls, utils/model_utils.py:113-114, so each has its own batch statistics):
//   plane 0 scale = gamma * rstd      plane 1 shift = beta - mean * scale        (z = u * scale + shift)
//   plane 2 mean                      plane 3 rstd = 1 / sqrt(var_biased + eps)
//   plane 4 var_unbiased (for the module's running_var update)
#pragma once

#include "contrastive_kernels.cuh"

namespace simclr {

constexpr int kBnPlanes = 5;
constexpr int kBnRowsPerBlock = 64;

// Batch statistics of both views.  grid = 2 * ceil(B / 64) blocks of 256 threads (8 warps, a warp per row, lane l owns the
// columns l, l + 32, ...).  Sums are taken about the view's first row (x0): var = E[(u - x0)^2] - E[u - x0]^2 without
// the cancellation of the raw moments.  Per-block partials, last block (ticket) adds them in block order: deterministic.
template <typename T>
__global__ void __launch_bounds__(256) bn_stats_kernel(const T* __restrict__ u1, const T* __restrict__ u2, int b, int d, int d_pad,
                                                       const float* __restrict__ gamma, const float* __restrict__ beta, float eps,
                                                       float* __restrict__ state, float* __restrict__ partial,
                                                       unsigned int* __restrict__ ticket) {
    constexpr int kPer = 8;                       // d_pad <= 256
    __shared__ float red[8][2][256];
    __shared__ int is_last;
    pdl_launch_dependents();
    pdl_wait();
    const int blocks_per_view = (b + kBnRowsPerBlock - 1) / kBnRowsPerBlock;
    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const T* u = view == 0 ? u1 : u2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    float x0[kPer], s1[kPer], s2[kPer];
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        const int k = lane + 32 * j;
        x0[j] = k < d ? load_as_float(u + k) : 0.f;
        s1[j] = 0.f;
        s2[j] = 0.f;
    }
    const int r_end = min(b, (blk + 1) * kBnRowsPerBlock);
    for (int r = blk * kBnRowsPerBlock + warp; r < r_end; r += 8) {
#pragma unroll
        for (int j = 0; j < kPer; ++j) {
            const int k = lane + 32 * j;
            if (k < d) {
                const float v = load_as_float(u + static_cast<size_t>(r) * d + k) - x0[j];
                s1[j] += v;
                s2[j] = fmaf(v, v, s2[j]);
            }
        }
    }
#pragma unroll
    for (int j = 0; j < kPer; ++j) {
        red[warp][0][lane + 32 * j] = s1[j];
        red[warp][1][lane + 32 * j] = s2[j];
    }
    __syncthreads();
    if (threadIdx.x < d_pad) {
        float a = 0.f, c = 0.f;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            a += red[w][0][threadIdx.x];
            c += red[w][1][threadIdx.x];
        }
        float* dst = partial + static_cast<size_t>(blockIdx.x) * 2 * d_pad;
        dst[threadIdx.x] = a;
        dst[d_pad + threadIdx.x] = c;
    }
    __threadfence();
    __syncthreads();
    if (threadIdx.x == 0) is_last = atomicAdd(ticket, 1u) == gridDim.x - 1 ? 1 : 0;
    __syncthreads();
    if (!is_last) return;
    __threadfence();
    for (int idx = threadIdx.x; idx < 2 * d_pad; idx += blockDim.x) {
        const int v = idx / d_pad, k = idx - v * d_pad;
        float* st = state + static_cast<size_t>(v) * kBnPlanes * d_pad;
        if (k >= d) {
            for (int pl = 0; pl < kBnPlanes; ++pl) st[pl * d_pad + k] = 0.f;
            continue;
        }
        float a = 0.f, c = 0.f;
        for (int q = 0; q < blocks_per_view; ++q) {
            const float* src = partial + static_cast<size_t>(v * blocks_per_view + q) * 2 * d_pad;
            a += __ldcg(src + k);
            c += __ldcg(src + d_pad + k);
        }
        const float shift0 = load_as_float((v == 0 ? u1 : u2) + k);
        const float inv_b = 1.0f / static_cast<float>(b);
        const float m1 = a * inv_b;
        const float var = fmaxf(c * inv_b - m1 * m1, 0.f);
        const float mean = shift0 + m1;
        const float rstd = rsqrtf(var + eps);
        const float scale = gamma[k] * rstd;
        st[0 * d_pad + k] = scale;
        st[1 * d_pad + k] = beta[k] - mean * scale;
        st[2 * d_pad + k] = mean;
        st[3 * d_pad + k] = rstd;
        st[4 * d_pad + k] = b > 1 ? var * static_cast<float>(b) / static_cast<float>(b - 1) : var;
    }
    if (threadIdx.x == 0) *ticket = 0u;          // left clean for the next call
}

// dL/dz (in g1 / g2, written by the backward finalize kernel) -> dL/du in place, BatchNorm training-mode backward:
//     du = scale * (dz - mean_r(dz) - xhat * mean_r(dz * xhat)),   xhat = (u - mean) * rstd
// and dgamma = sum_r dz * xhat, dbeta = sum_r dz (both views added: the module is shared).  The column sums come from the
// per-CTA partials of the backward finalize kernel, [n_part][2][Dpad] per view, added in CTA order by every block for
// itself (64 K floats from L2: cheaper than another launch, and deterministic).  eval mode: du = scale * dz.
template <typename T>
__global__ void __launch_bounds__(256) bn_backward_kernel(const T* __restrict__ u1, const T* __restrict__ u2, T* __restrict__ g1,
                                                          T* __restrict__ g2, int b, int d, int d_pad,
                                                          const float* __restrict__ state, const float* __restrict__ partial,
                                                          int parts_per_view, int training, float* __restrict__ dgamma,
                                                          float* __restrict__ dbeta) {
    __shared__ float sums[2][256];
    pdl_launch_dependents();
    pdl_wait();
    const int blocks_per_view = (b + kBnRowsPerBlock - 1) / kBnRowsPerBlock;
    const int view = blockIdx.x / blocks_per_view;
    const int blk = blockIdx.x - view * blocks_per_view;
    const T* u = view == 0 ? u1 : u2;
    T* g = view == 0 ? g1 : g2;
    const float* st = state + static_cast<size_t>(view) * kBnPlanes * d_pad;
    for (int idx = threadIdx.x; idx < 2 * d_pad; idx += blockDim.x) {
        const int q = idx / d_pad, k = idx - q * d_pad;
        float a = 0.f;
        for (int c = 0; c < parts_per_view; ++c)
            a += __ldcg(partial + (static_cast<size_t>(view * parts_per_view + c) * 2 + q) * d_pad + k);
        sums[q][k] = a;
    }
    __syncthreads();
    // the parameter gradients: one block per view adds its share (two atomics per column and call: fixed pair, commutative)
    if (blk == 0) {
        for (int k = threadIdx.x; k < d; k += blockDim.x) {
            if (dbeta != nullptr) atomicAdd(dbeta + k, sums[0][k]);
            if (dgamma != nullptr) atomicAdd(dgamma + k, sums[1][k]);
        }
    }
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const float inv_b = 1.0f / static_cast<float>(b);
    const int r_end = min(b, (blk + 1) * kBnRowsPerBlock);
    for (int r = blk * kBnRowsPerBlock + warp; r < r_end; r += 8) {
        for (int k = lane; k < d; k += 32) {
            const size_t at = static_cast<size_t>(r) * d + k;
            const float dz = load_as_float(g + at);
            float du;
            if (training) {
                const float xhat = (load_as_float(u + at) - st[2 * d_pad + k]) * st[3 * d_pad + k];
                du = st[0 * d_pad + k] * (dz - sums[0][k] * inv_b - xhat * sums[1][k] * inv_b);
            } else {
                du = st[0 * d_pad + k] * dz;
            }
            store_from_float(g + at, du);
        }
    }
}

}  // namespace simclr
