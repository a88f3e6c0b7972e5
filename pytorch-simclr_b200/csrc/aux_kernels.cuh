// Row-wise prologue / epilogue kernels around the tile kernel: operand preparation, forward finalize,
// backward preparation and backward finalize.  All O(M*d) work, HBM-bound, one warp per image pair.
#pragma once

#include "contrastive_kernels.cuh"

namespace simclr {

constexpr float kNormEps = 1e-12f;          // F.normalize eps (reference objective.py:26-27, :77-78)
constexpr float kSoftplusBeta = 0.8f;       // reference objective.py:70-71
constexpr float kSoftplusThreshold = 20.f;  // torch default threshold of F.softplus
constexpr float kInvNormClamped = 1e12f;    // marker: the norm was clamped by eps
constexpr int kMaxDimPerLane = 8;           // Dpad <= 256 -> at most 8 elements per lane

template <typename T>
SIMCLR_DEVICE float load_as_float(const T* p);
template <>
SIMCLR_DEVICE float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
SIMCLR_DEVICE float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

SIMCLR_DEVICE void store_from_float(float* p, float v) { *p = v; }
SIMCLR_DEVICE void store_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

SIMCLR_DEVICE float softplus_beta(float x) {
    const float bx = kSoftplusBeta * x;
    return bx > kSoftplusThreshold ? x : log1pf(expf(bx)) / kSoftplusBeta;
}

struct AuxParams {
    int b_loc, b_glob, row_off, bl_pad, bg_pad;
    int d, d_pad;
    int normalize;
    float k2;        // NT-Xent: log2(e)/tau, modified: 1/tau
    float inv_tau;
    float m2;
    int const_shift;
    float qscale;    // (float) b_glob
};

// ---------------------------------------------------------------------------------------------
// Stage 1: x_batch{1,2} -> bf16 operand rows, inv_norm, exact positive-pair dot product
// ---------------------------------------------------------------------------------------------
template <typename T, int kLoss>
__global__ void prepare_kernel(const T* __restrict__ x1, const T* __restrict__ x2, AuxParams a,
                               __nv_bfloat16* __restrict__ operand, float* __restrict__ inv_norm,
                               float* __restrict__ pos_dot) {
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
    float n1 = 0.f, n2 = 0.f;
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
        }
        v1[u] = e1;
        v2[u] = e2;
    }
    n1 = warp_sum(n1);
    n2 = warp_sum(n2);
    float inv1 = 1.f, inv2 = 1.f;
    if (kLoss == kModified || a.normalize) {
        if constexpr (kLoss == kNtXent) {
            n1 = sqrtf(n1);
            n2 = sqrtf(n2);
        }
        inv1 = n1 < kNormEps ? kInvNormClamped : 1.f / n1;
        inv2 = n2 < kNormEps ? kInvNormClamped : 1.f / n2;
        const float den1 = fmaxf(n1, kNormEps), den2 = fmaxf(n2, kNormEps);
#pragma unroll
        for (int u = 0; u < kMaxDimPerLane; ++u) {
            v1[u] = v1[u] / den1;
            v2[u] = v2[u] / den2;
        }
    }
    float dot = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) dot = fmaf(v1[u], v2[u], dot);
    dot = warp_sum(dot);
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        if (u < per_lane) {
            o1[lane + 32 * u] = __float2bfloat16_rn(v1[u]);
            o2[lane + 32 * u] = __float2bfloat16_rn(v2[u]);
        }
    }
    if (lane == 0) {
        inv_norm[i] = inv1;
        inv_norm[a.bl_pad + i] = inv2;
        pos_dot[i] = dot;
        pos_dot[a.bl_pad + i] = dot;
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 2 epilogue: merge the per-CTA partials of every row, add the exact positive term, produce
// lse2 / row_loss / stats.  One block per row block, one thread per row.
// ---------------------------------------------------------------------------------------------
template <int kLoss>
SIMCLR_DEVICE float exact_logit2(const AuxParams& a, float v) {
    if constexpr (kLoss == kNtXent) return v * a.k2;
    else return log2f(v) * a.k2;
}

template <int kLoss>
__global__ void __launch_bounds__(kBlockM)
forward_finalize_kernel(AuxParams a, const float* __restrict__ part, int grid_tiles, int max_segs, int n_col_tiles,
                        long long total_tiles, const float* __restrict__ pos_dot,
                        const float* __restrict__ row_weight, float* __restrict__ lse2,
                        float* __restrict__ row_loss, float* __restrict__ block_part, unsigned int* ticket,
                        float* __restrict__ stats, float* __restrict__ loss_out) {
    const int rb = blockIdx.x;
    const int tid = threadIdx.x;
    const int blocks_per_view = a.bl_pad / kBlockM;
    const int vr = rb / blocks_per_view;
    const int img = (rb - vr * blocks_per_view) * kBlockM + tid;
    const bool row_ok = img < a.b_loc;
    const int slot = rb * kBlockM + tid;

    // positive pair in exact fp32
    float v_pos = pos_dot[slot];
    if constexpr (kLoss == kModified) v_pos = fmaxf(v_pos * a.qscale, kClampMin);

    // CTAs whose contiguous tile range [T*k/G, T*(k+1)/G) overlaps this row block: owner(t) = ((t+1)*G - 1) / T
    const long long t_lo = static_cast<long long>(rb) * n_col_tiles;
    const long long t_hi = t_lo + n_col_tiles;
    const int k_first = static_cast<int>(((t_lo + 1) * grid_tiles - 1) / total_tiles);
    const int k_last = static_cast<int>((t_hi * grid_tiles - 1) / total_tiles);

    // pass 1: maxima
    float vmax = v_pos, max_prec = kNegBig, max_foll = kNegBig, pos_mma = kNegBig;
    for (int k = k_first; k <= k_last; ++k) {
        const long long c0 = (total_tiles * k) / grid_tiles;
        const int seg = rb - static_cast<int>(c0 / n_col_tiles);
        for (int wg = 0; wg < kNumSoftmaxWG; ++wg) {
            const float* src = part + ((static_cast<size_t>(k) * max_segs + seg) * kNumSoftmaxWG + wg) * (kFwdFields * kBlockM) + tid;
            vmax = fmaxf(vmax, src[1 * kBlockM]);
            max_prec = fmaxf(max_prec, src[2 * kBlockM]);
            max_foll = fmaxf(max_foll, src[3 * kBlockM]);
            pos_mma = fmaxf(pos_mma, src[4 * kBlockM]);
        }
    }
    // pass 2: rescaled sums
    const float top = exact_logit2<kLoss>(a, vmax);
    float total = exp2f(exact_logit2<kLoss>(a, v_pos) - top);
    for (int k = k_first; k <= k_last; ++k) {
        const long long c0 = (total_tiles * k) / grid_tiles;
        const int seg = rb - static_cast<int>(c0 / n_col_tiles);
        for (int wg = 0; wg < kNumSoftmaxWG; ++wg) {
            const float* src = part + ((static_cast<size_t>(k) * max_segs + seg) * kNumSoftmaxWG + wg) * (kFwdFields * kBlockM) + tid;
            const float s = src[0];
            const float m = src[1 * kBlockM];
            if (m > kNegBig) total += s * exp2f(exact_logit2<kLoss>(a, m) - top);
        }
    }
    float l2 = 0.f, loss_r = 0.f, w = 0.f, hit = 0.f;
    if (row_ok) {
        l2 = top + log2f(total);
        loss_r = (l2 - exact_logit2<kLoss>(a, v_pos)) * 0.6931471805599453f;
        w = row_weight ? row_weight[vr * a.b_loc + img] : 1.f;
        // reference objective.py:51 -- Tensor.max returns the first maximal index
        hit = (max_prec < pos_mma && max_foll <= pos_mma) ? 1.f : 0.f;
    }
    lse2[slot] = l2;
    row_loss[slot] = loss_r;

    // block reduction (fixed order -> deterministic)
    __shared__ float red[3][kBlockM / 32];
    float r0 = warp_sum(w * loss_r), r1 = warp_sum(w), r2 = warp_sum(hit);
    if ((tid & 31) == 0) {
        red[0][tid >> 5] = r0;
        red[1][tid >> 5] = r1;
        red[2][tid >> 5] = r2;
    }
    __syncthreads();
    __shared__ bool is_last;
    if (tid == 0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int i = 0; i < kBlockM / 32; ++i) {
            s0 += red[0][i];
            s1 += red[1][i];
            s2 += red[2][i];
        }
        block_part[rb * 4 + 0] = s0;
        block_part[rb * 4 + 1] = s1;
        block_part[rb * 4 + 2] = s2;
        __threadfence();
        const unsigned int prev = atomicAdd(ticket, 1u);
        is_last = (prev == gridDim.x - 1);
    }
    __syncthreads();
    if (is_last && tid < 32) {
        __threadfence();
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int i = tid; i < static_cast<int>(gridDim.x); i += 32) {
            s0 += __ldcg(block_part + i * 4 + 0);
            s1 += __ldcg(block_part + i * 4 + 1);
            s2 += __ldcg(block_part + i * 4 + 2);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (tid == 0) {
            stats[0] = s0;
            stats[1] = s1;
            stats[2] = s2;
            stats[3] = s0 / s1;
            if (loss_out) *loss_out = s0 / s1;
            *ticket = 0u;
        }
    }
}

// ---------------------------------------------------------------------------------------------
// Stage 3 prologue: per-column vectors for the tile kernel, zero the accumulation buffer
// ---------------------------------------------------------------------------------------------
__global__ void backward_prepare_kernel(AuxParams a, const float* __restrict__ lse2_cols,
                                        const float* __restrict__ col_scale, float* __restrict__ colvec,
                                        float4* __restrict__ dacc4, size_t dacc_vec4) {
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
}

// ---------------------------------------------------------------------------------------------
// Stage 3 epilogue: exact positive-pair term + backward of the row normalisation (+ softplus)
// ---------------------------------------------------------------------------------------------
template <typename T, int kLoss>
__global__ void backward_finalize_kernel(const T* __restrict__ x1, const T* __restrict__ x2, AuxParams a,
                                         const float* __restrict__ inv_norm, const float* __restrict__ pos_dot,
                                         const float* __restrict__ lse2_cols, const float* __restrict__ col_scale,
                                         const float* __restrict__ grad_out, const float* __restrict__ dacc,
                                         T* __restrict__ g1, T* __restrict__ g2) {
    const int warps_per_block = blockDim.x >> 5;
    const int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5);
    const int lane = threadIdx.x & 31;
    if (i >= a.b_loc) return;
    const int per_lane = a.d_pad >> 5;
    const int s1 = i, s2 = a.bl_pad + i;                  // local slots
    const int c1 = a.row_off + i, c2 = a.bg_pad + a.row_off + i;   // global column ids of the two rows
    const float go = grad_out ? __ldg(grad_out) : 1.f;
    const float inv1 = inv_norm[s1], inv2 = inv_norm[s2];

    float raw1[kMaxDimPerLane], raw2[kMaxDimPerLane], h1[kMaxDimPerLane], h2[kMaxDimPerLane];
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        const int k = lane + 32 * u;
        float e1 = 0.f, e2 = 0.f;
        if (u < per_lane && k < a.d) {
            e1 = load_as_float(x1 + static_cast<size_t>(i) * a.d + k);
            e2 = load_as_float(x2 + static_cast<size_t>(i) * a.d + k);
        }
        raw1[u] = e1;
        raw2[u] = e2;
        if constexpr (kLoss == kModified) {
            const bool in = (u < per_lane && k < a.d);
            e1 = in ? softplus_beta(e1) : 0.f;
            e2 = in ? softplus_beta(e2) : 0.f;
        }
        const bool scaled = (kLoss == kModified) || a.normalize;
        h1[u] = scaled ? e1 * (inv1 == kInvNormClamped ? 1.f / kNormEps : inv1) : e1;
        h2[u] = scaled ? e2 * (inv2 == kInvNormClamped ? 1.f / kNormEps : inv2) : e2;
    }
    const float sc1 = col_scale ? col_scale[c1] : 0.5f / static_cast<float>(a.b_glob);
    const float sc2 = col_scale ? col_scale[c2] : 0.5f / static_cast<float>(a.b_glob);
    const float l21 = lse2_cols[c1], l22 = lse2_cols[c2];
    const float pd = pos_dot[s1];

    float coef, outer;
    if constexpr (kLoss == kNtXent) {
        // (g_r P[r,pos] + g_pos P[pos,r] - g_r - g_pos) * zhat_pos, exact fp32
        const float y = pd * a.k2;
        coef = sc1 * (exp2f(y - l21) - 1.f) + sc2 * (exp2f(y - l22) - 1.f);
        outer = a.inv_tau * go;
    } else {
        const float qv = pd * a.qscale;
        const bool live = qv >= kClampMin;
        const float lq = log2f(fmaxf(qv, kClampMin));
        const float y = lq * (a.k2 - 1.f);
        coef = live ? (sc1 * exp2f(y - l21) + sc2 * exp2f(y - l22) - (sc1 + sc2) * exp2f(-lq)) : 0.f;
        outer = a.inv_tau * a.qscale * go;
    }

    float d1[kMaxDimPerLane], d2[kMaxDimPerLane];
    float t1 = 0.f, t2 = 0.f;
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        const int k = lane + 32 * u;
        float a1 = 0.f, a2 = 0.f;
        if (u < per_lane) {
            a1 = dacc[static_cast<size_t>(s1) * a.d_pad + k];
            a2 = dacc[static_cast<size_t>(s2) * a.d_pad + k];
        }
        d1[u] = (a1 + coef * h2[u]) * outer;
        d2[u] = (a2 + coef * h1[u]) * outer;
        t1 = fmaf(d1[u], h1[u], t1);
        t2 = fmaf(d2[u], h2[u], t2);
    }
    t1 = warp_sum(t1);
    t2 = warp_sum(t2);
#pragma unroll
    for (int u = 0; u < kMaxDimPerLane; ++u) {
        const int k = lane + 32 * u;
        if (u < per_lane && k < a.d) {
            float o1 = d1[u], o2 = d2[u];
            if constexpr (kLoss == kNtXent) {
                if (a.normalize) {
                    // d/dz of z / max(||z||, eps): projection unless the clamp was active
                    o1 = (inv1 == kInvNormClamped) ? o1 / kNormEps : (o1 - h1[u] * t1) * inv1;
                    o2 = (inv2 == kInvNormClamped) ? o2 / kNormEps : (o2 - h2[u] * t2) * inv2;
                }
            } else {
                // L1 normalisation of a positive vector, then softplus'(x) = sigmoid(beta x)
                o1 = (inv1 == kInvNormClamped) ? o1 / kNormEps : (o1 - t1) * inv1;
                o2 = (inv2 == kInvNormClamped) ? o2 / kNormEps : (o2 - t2) * inv2;
                const float b1 = kSoftplusBeta * raw1[u], b2 = kSoftplusBeta * raw2[u];
                o1 *= (b1 > kSoftplusThreshold) ? 1.f : 1.f / (1.f + expf(-b1));
                o2 *= (b2 > kSoftplusThreshold) ? 1.f : 1.f / (1.f + expf(-b2));
            }
            store_from_float(g1 + static_cast<size_t>(i) * a.d + k, o1);
            store_from_float(g2 + static_cast<size_t>(i) * a.d + k, o2);
        }
    }
}

}  // namespace simclr
