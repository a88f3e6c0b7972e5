// Row-wise prologue kernels around the tile kernels: operand preparation (stage 1) and backward preparation
// (per-column vectors, zeroing of the accumulation buffer).  O(M*d) work, HBM / latency bound.  The epilogue
// kernels are thin wrappers around the row-block finalize routines of contrastive_kernels.cuh.
#pragma once

#include "contrastive_kernels.cuh"

namespace simclr {

struct AuxParams {
    int b_loc, b_glob, row_off, bl_pad, bg_pad;
    int d, d_pad;
    int normalize;
    float k2;        // NT-Xent: log2(e)/tau, modified: 1/tau
    float inv_tau;
    float m2;
    int const_shift;
    float qscale;    // (float) b_glob
    float op_scale;  // factor applied to the normalised rows before bf16 rounding (NT-Xent: sqrt(log2(e)/tau), else 1)
};

// ---------------------------------------------------------------------------------------------
// Stage 1: x_batch{1,2} -> bf16 operand rows, inv_norm, exact positive-pair dot product.
// One warp per image (both views); a single round of warp reductions (norms and the raw dot together).
// Also zeroes `zero_words` 32-bit words at `zero_ptr` (the forward workspace header) when given.
// ---------------------------------------------------------------------------------------------
template <typename T, int kLoss>
__global__ void prepare_kernel(const T* __restrict__ x1, const T* __restrict__ x2, AuxParams a,
                               __nv_bfloat16* __restrict__ operand, float* __restrict__ inv_norm,
                               float* __restrict__ pos_dot, unsigned int* __restrict__ zero_ptr, int zero_words,
                               unsigned long long* ktrace) {
    pdl_launch_dependents();
    pdl_wait();
    ktrace_begin(ktrace, 0);
    struct End { unsigned long long* k; __device__ ~End() { ktrace_end(k, 0); } } end_guard{ktrace};
    if (zero_ptr != nullptr && blockIdx.x == 0)
        for (int i = threadIdx.x; i < zero_words; i += blockDim.x) zero_ptr[i] = 0u;
    const int warps_per_block = blockDim.x >> 5;
    const int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5);   // local image slot
    const int lane = threadIdx.x & 31;
    if (i >= a.bl_pad) return;
    const int per_lane = a.d_pad >> 5;
    __nv_bfloat16* o1 = operand + static_cast<size_t>(i) * a.d_pad;
    __nv_bfloat16* o2 = operand + static_cast<size_t>(a.bl_pad + i) * a.d_pad;
    if (i >= a.b_loc) {     // padding slot: zero operands and neutral per-row values
        for (int u = 0; u < per_lane; ++u) {
            o1[lane + 32 * u] = __float2bfloat16_rn(0.f);
            o2[lane + 32 * u] = __float2bfloat16_rn(0.f);
        }
        if (lane == 0) {
            inv_norm[i] = 0.f;
            inv_norm[a.bl_pad + i] = 0.f;
            pos_dot[i] = 0.f;
            pos_dot[a.bl_pad + i] = 0.f;
        }
        return;
    }
    float v1[kMaxDimPerLane], v2[kMaxDimPerLane];
    float n1 = 0.f, n2 = 0.f, dot = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        const int k = lane + 32 * u;
        float e1 = 0.f, e2 = 0.f;
        if (u < per_lane && k < a.d) {
            e1 = load_as_float(x1 + static_cast<size_t>(i) * a.d + k);
            e2 = load_as_float(x2 + static_cast<size_t>(i) * a.d + k);
            if constexpr (kLoss == kModified) {
                e1 = softplus_beta(e1);
                e2 = softplus_beta(e2);
                n1 += fabsf(e1);
                n2 += fabsf(e2);
            } else {
                n1 = fmaf(e1, e1, n1);
                n2 = fmaf(e2, e2, n2);
            }
            dot = fmaf(e1, e2, dot);
        }
        v1[u] = e1;
        v2[u] = e2;
    }
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {          // three independent butterflies interleaved
        n1 += __shfl_xor_sync(0xffffffffu, n1, o);
        n2 += __shfl_xor_sync(0xffffffffu, n2, o);
        dot += __shfl_xor_sync(0xffffffffu, dot, o);
    }
    float inv1 = 1.f, inv2 = 1.f, s1 = 1.f, s2 = 1.f;
    if (kLoss == kModified || a.normalize) {
        if constexpr (kLoss == kNtXent) {
            n1 = sqrtf(n1);
            n2 = sqrtf(n2);
        }
        s1 = 1.f / fmaxf(n1, kNormEps);
        s2 = 1.f / fmaxf(n2, kNormEps);
        inv1 = n1 < kNormEps ? kInvNormClamped : s1;
        inv2 = n2 < kNormEps ? kInvNormClamped : s2;
    }
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        if (u < per_lane) {
            o1[lane + 32 * u] = __float2bfloat16_rn(v1[u] * (s1 * a.op_scale));
            o2[lane + 32 * u] = __float2bfloat16_rn(v2[u] * (s2 * a.op_scale));
        }
    }
    if (lane == 0) {
        inv_norm[i] = inv1;
        inv_norm[a.bl_pad + i] = inv2;
        const float pd = dot * s1 * s2;
        pos_dot[i] = pd;
        pos_dot[a.bl_pad + i] = pd;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 3 prologue: per-column vectors for the tile kernel; zero the accumulation buffer.
// ---------------------------------------------------------------------------------------------
__global__ void backward_prepare_kernel(AuxParams a, const float* __restrict__ lse2_cols,
                                        const float* __restrict__ col_scale, float* __restrict__ colvec,
                                        float4* __restrict__ dacc4, size_t dacc_vec4, unsigned long long* ktrace) {
    pdl_launch_dependents();
    pdl_wait();
    ktrace_begin(ktrace, 2);
    const size_t tid = static_cast<size_t>(blockIdx.x) * blockDim.x + threadIdx.x;
    const size_t nthreads = static_cast<size_t>(gridDim.x) * blockDim.x;
    const int ncols = 2 * a.bg_pad;
    for (size_t c = tid; c < static_cast<size_t>(ncols); c += nthreads) {
        const int ic = static_cast<int>(c % a.bg_pad);
        const bool ok = ic < a.b_glob;
        const float scale = ok ? (col_scale ? col_scale[c] : 0.5f / static_cast<float>(a.b_glob)) : 0.f;
        const float l2 = ok ? lse2_cols[c] : 0.f;
        colvec[c] = a.const_shift ? scale * exp2f(a.m2 - l2) : scale;
        colvec[ncols + c] = l2;
    }
    const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
    for (size_t i = tid; i < dacc_vec4; i += nthreads) dacc4[i] = z;
    ktrace_end(ktrace, 2);
}

// ---------------------------------------------------------------------------------------------
// Stage 2 epilogue: merge the per-CTA partials of every row, add the exact positive term, produce
// lse2 / row_loss / stats.  One block per row block, one thread per row.
// ---------------------------------------------------------------------------------------------
template <int kLoss>
__global__ void __launch_bounds__(kBlockM) forward_finalize_kernel(const TileParams p) {
    __shared__ float red[16];
    __shared__ int flags[4];
    pdl_launch_dependents();
    pdl_wait();
    ktrace_begin(p.ktrace, 4);
    forward_finalize_rowblock<kLoss>(p, blockIdx.x, threadIdx.x, red, flags);
    ktrace_end(p.ktrace, 4);
}

// ---------------------------------------------------------------------------------------------
// Stage 3 epilogue: exact positive-pair term + backward of the row normalisation (+ softplus).
// kBwdFinBlocksPerRowBlock blocks of 16 warps per row block: one warp per row.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdFinBlocksPerRowBlock = 8;
template <int D, int kLoss>
__global__ void __launch_bounds__(512) backward_finalize_kernel(const TileParams p) {
    pdl_launch_dependents();
    pdl_wait();
    ktrace_begin(p.ktrace, 5);
    const int rb = blockIdx.x / kBwdFinBlocksPerRowBlock;
    const int sub = blockIdx.x % kBwdFinBlocksPerRowBlock;
    backward_finalize_rowblock<D, kLoss>(p, rb, sub * 16 + (threadIdx.x >> 5), threadIdx.x & 31,
                                         16 * kBwdFinBlocksPerRowBlock);
    ktrace_end(p.ktrace, 5);
}

}  // namespace simclr
