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
    int split;       // 1: also write the residual plane lo = bf16(x - float(bf16(x))) behind the hi plane (fp32-grade mode)
    // projection-head tail (head_kernels.cuh): BatchNorm state f32 [2][kBnPlanes][d_pad] or nullptr; the rows are then the
    // PRE-BatchNorm activations and z = u * scale + shift is formed in registers
    const float* bn_state;
};

// ---------------------------------------------------------------------------------------------
// Stage 1: x_batch{1,2} -> bf16 operand rows, inv_norm, exact positive-pair dot product.
// One warp per image (both views); lane l owns the d_pad/32 consecutive columns [l*per, (l+1)*per) so that a row leaves
// the warp as ONE packed store of 2*d_pad bytes; a single round of warp reductions (norms and the raw dot together).
// Also zeroes `zero_words` 32-bit words at `zero_ptr` (the forward workspace header) when given.
// Row-sharded global batch: with peers.world > 0 every row is additionally stored into ALL ranks' copies of the global
// operand matrix (view-padded, rank-major inside a view) -- the "all-gather" is these NVLink stores, fused here.
// ---------------------------------------------------------------------------------------------
template <int kWords>
SIMCLR_DEVICE void store_words(__nv_bfloat16* row, int lane, const uint32_t (&w)[4]) {
    if constexpr (kWords == 1) reinterpret_cast<uint32_t*>(row)[lane] = w[0];
    else if constexpr (kWords == 2) reinterpret_cast<uint2*>(row)[lane] = make_uint2(w[0], w[1]);
    else reinterpret_cast<uint4*>(row)[lane] = make_uint4(w[0], w[1], w[2], w[3]);
}

// the same packed store through a multicast address: the NVSwitch replicates it into every rank's copy
template <int kWords>
SIMCLR_DEVICE void store_words_multicast(__nv_bfloat16* row, int lane, const uint32_t (&w)[4]) {
    if constexpr (kWords == 1) {
        asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" ::"l"(reinterpret_cast<uint32_t*>(row) + lane),
                     "f"(__uint_as_float(w[0]))
                     : "memory");
    } else if constexpr (kWords == 2) {
        asm volatile("multimem.st.relaxed.sys.global.v2.f32 [%0], {%1, %2};" ::"l"(reinterpret_cast<uint2*>(row) + lane),
                     "f"(__uint_as_float(w[0])), "f"(__uint_as_float(w[1]))
                     : "memory");
    } else {
        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(reinterpret_cast<uint4*>(row) + lane),
                     "f"(__uint_as_float(w[0])), "f"(__uint_as_float(w[1])), "f"(__uint_as_float(w[2])),
                     "f"(__uint_as_float(w[3]))
                     : "memory");
    }
}

template <typename T, int kLoss, int kPer /* d_pad / 32: 2, 4 or 8 */>
__global__ void prepare_kernel(const T* __restrict__ x1, const T* __restrict__ x2, AuxParams a,
                               __nv_bfloat16* __restrict__ operand, float* __restrict__ inv_norm,
                               float* __restrict__ pos_dot, unsigned int* __restrict__ zero_ptr, int zero_words,
                               unsigned long long* ktrace, PeerTable peers, unsigned int* bump_epoch,
                               unsigned int* __restrict__ cand_cnt, float* __restrict__ zstash, PeerTable signal_flags) {
    pdl_launch_dependents();
    // The input rows are loaded BEFORE griddepcontrol.wait (they are the caller's tensors: no kernel of this library
    // writes them, and a foreign producer never lets this kernel start early); every store comes after it, because the
    // previous step's kernels may still be reading the buffers this kernel overwrites.
    const int warps_per_block = blockDim.x >> 5;
    const int i = blockIdx.x * warps_per_block + (threadIdx.x >> 5);   // local image slot
    const int lane = threadIdx.x & 31;
    constexpr int kWords = kPer / 2;
    float v1[kPer], v2[kPer];
#pragma unroll
    for (int u = 0; u < kPer; ++u) v1[u] = v2[u] = 0.f;
    if (i < a.b_loc) {
        const T* r1 = x1 + static_cast<size_t>(i) * a.d;
        const T* r2 = x2 + static_cast<size_t>(i) * a.d;
        bool vec = false;
        if constexpr (sizeof(T) == 4 && kPer == 4) {
            // fp32 rows of exactly 128 columns at 16-byte aligned addresses: one float4 per lane
            vec = a.d == a.d_pad && ((reinterpret_cast<uintptr_t>(x1) | reinterpret_cast<uintptr_t>(x2)) & 15u) == 0;
            if (vec) {
                const float4 p4 = __ldg(reinterpret_cast<const float4*>(r1) + lane);
                const float4 q4 = __ldg(reinterpret_cast<const float4*>(r2) + lane);
                v1[0] = p4.x; v1[1] = p4.y; v1[2] = p4.z; v1[3] = p4.w;
                v2[0] = q4.x; v2[1] = q4.y; v2[2] = q4.z; v2[3] = q4.w;
            }
        }
        if (!vec) {
#pragma unroll
            for (int u = 0; u < kPer; ++u) {
                const int k = lane * kPer + u;
                v1[u] = k < a.d ? load_as_float(r1 + k) : 0.f;
                v2[u] = k < a.d ? load_as_float(r2 + k) : 0.f;
            }
        }
    }
    pdl_wait();
    ktrace_begin(ktrace, 0);
    struct End { unsigned long long* k; __device__ ~End() { ktrace_end(k, 0); } } end_guard{ktrace};
    // Fused row-sharded step: this rank's side of the barrier behind the operand push is signalled from HERE, by the warp
    // that finishes last (bump_epoch[1] counts the warps: zero on entry, left zero), instead of by the forward tile kernel
    // once it has been launched and has reached its first column tile -- the peers see the flag a launch latency earlier.
    // Runs at every exit of the kernel (exits are warp-uniform).
    struct Signal {
        const PeerTable& flags;
        unsigned int* epoch;
        __device__ ~Signal() {
            if (flags.world == 0) return;
            __threadfence_system();                       // this thread's peer stores are performed
            __syncwarp();
            if ((threadIdx.x & 31) != 0) return;
            const unsigned int warps = gridDim.x * (blockDim.x >> 5);
            if (atomicAdd(epoch + 1, 1u) != warps - 1u) return;
            epoch[1] = 0u;
            __threadfence();
            const unsigned int target = *reinterpret_cast<volatile unsigned int*>(epoch);
            for (int r = 0; r < flags.world; ++r) {
                unsigned int* remote = static_cast<unsigned int*>(flags.ptr[r]) + flags.rank;
                asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(target) : "memory");
            }
        }
    } signal_guard{signal_flags, bump_epoch};
    if (zero_ptr != nullptr && blockIdx.x == 0)
        for (int j = threadIdx.x; j < zero_words; j += blockDim.x) zero_ptr[j] = 0u;
    // fused row-sharded step: the epoch of the barrier that the forward tile kernel executes (TileParams::sync_epoch)
    if (bump_epoch != nullptr && blockIdx.x == 0 && threadIdx.x == 0) *bump_epoch += 1u;
    if (a.bn_state != nullptr) {
        // BatchNorm apply (the state comes from a kernel of this stream: read behind the wait)
        const float* st1 = a.bn_state;
        const float* st2 = a.bn_state + 5 * a.d_pad;
#pragma unroll
        for (int u = 0; u < kPer; ++u) {
            const int k = lane * kPer + u;
            v1[u] = fmaf(v1[u], __ldg(st1 + k), __ldg(st1 + a.d_pad + k));
            v2[u] = fmaf(v2[u], __ldg(st2 + k), __ldg(st2 + a.d_pad + k));
        }
    }
    if (i >= a.bl_pad) return;
    // candidate counters of the exact accuracy count (forward workspace): zero for both views of this image slot
    if (cand_cnt != nullptr && lane == 0) {
        cand_cnt[i] = 0u;
        cand_cnt[a.bl_pad + i] = 0u;
    }
    __nv_bfloat16* o1 = operand + static_cast<size_t>(i) * a.d_pad;
    __nv_bfloat16* o2 = operand + static_cast<size_t>(a.bl_pad + i) * a.d_pad;
    uint32_t w1[4] = {0u, 0u, 0u, 0u}, w2[4] = {0u, 0u, 0u, 0u};
    const size_t lo_plane = static_cast<size_t>(2) * a.bl_pad * a.d_pad;      // elements between the hi and lo planes
    if (i >= a.b_loc) {     // padding slot: zero operands and neutral per-row values
        store_words<kWords>(o1, lane, w1);
        store_words<kWords>(o2, lane, w2);
        if (a.split) {
            store_words<kWords>(o1 + lo_plane, lane, w1);
            store_words<kWords>(o2 + lo_plane, lane, w2);
        }
        if (lane == 0) {
            inv_norm[i] = 0.f;
            inv_norm[a.bl_pad + i] = 0.f;
            pos_dot[i] = 0.f;
            pos_dot[a.bl_pad + i] = 0.f;
        }
        return;
    }
    float n1 = 0.f, n2 = 0.f, dot = 0.f;
#pragma unroll
    for (int u = 0; u < kPer; ++u) {
        float e1 = v1[u], e2 = v2[u];
        if constexpr (kLoss == kModified) {
            const bool in = lane * kPer + u < a.d;
            e1 = in ? softplus_beta(e1) : 0.f;
            e2 = in ? softplus_beta(e2) : 0.f;
            n1 += fabsf(e1);
            n2 += fabsf(e2);
        } else {
            n1 = fmaf(e1, e1, n1);
            n2 = fmaf(e2, e2, n2);
        }
        dot = fmaf(e1, e2, dot);
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
    const float f1 = s1 * a.op_scale, f2 = s2 * a.op_scale;
#pragma unroll
    for (int u = 0; u < kPer; u += 2) {
        w1[u >> 1] = pack_bf16x2(v1[u] * f1, v1[u + 1] * f1);
        w2[u >> 1] = pack_bf16x2(v2[u] * f2, v2[u + 1] * f2);
    }
    store_words<kWords>(o1, lane, w1);
    store_words<kWords>(o2, lane, w2);
    if (zstash != nullptr) {
        // exact fp32 normalised rows for the re-scoring of accuracy candidates by OTHER ranks (symmetric memory, read over
        // NVLink only for the few candidates) or after a gather: [2*bl_pad][d_pad], lane l owns kPer consecutive columns
        float* z1 = zstash + static_cast<size_t>(i) * a.d_pad + lane * kPer;
        float* z2 = zstash + static_cast<size_t>(a.bl_pad + i) * a.d_pad + lane * kPer;
#pragma unroll
        for (int u = 0; u < kPer; u += 2) {
            *reinterpret_cast<float2*>(z1 + u) = make_float2(v1[u] * s1, v1[u + 1] * s1);
            *reinterpret_cast<float2*>(z2 + u) = make_float2(v2[u] * s2, v2[u + 1] * s2);
        }
    }
    if (a.split) {
        uint32_t l1[4] = {0u, 0u, 0u, 0u}, l2[4] = {0u, 0u, 0u, 0u};
#pragma unroll
        for (int u = 0; u < kPer; u += 2) {
            const uint32_t p1 = w1[u >> 1], p2 = w2[u >> 1];
            l1[u >> 1] = pack_bf16x2(v1[u] * f1 - __uint_as_float(p1 << 16), v1[u + 1] * f1 - __uint_as_float(p1 & 0xffff0000u));
            l2[u >> 1] = pack_bf16x2(v2[u] * f2 - __uint_as_float(p2 << 16), v2[u + 1] * f2 - __uint_as_float(p2 & 0xffff0000u));
        }
        store_words<kWords>(o1 + lo_plane, lane, l1);
        store_words<kWords>(o2 + lo_plane, lane, l2);
    }
    if (peers.mc != nullptr) {
        __nv_bfloat16* g = static_cast<__nv_bfloat16*>(peers.mc);
        store_words_multicast<kWords>(g + static_cast<size_t>(a.row_off + i) * a.d_pad, lane, w1);
        store_words_multicast<kWords>(g + static_cast<size_t>(a.bg_pad + a.row_off + i) * a.d_pad, lane, w2);
    } else {
        for (int r = 0; r < peers.world; ++r) {
            __nv_bfloat16* g = static_cast<__nv_bfloat16*>(peers.ptr[r]);
            store_words<kWords>(g + static_cast<size_t>(a.row_off + i) * a.d_pad, lane, w1);
            store_words<kWords>(g + static_cast<size_t>(a.bg_pad + a.row_off + i) * a.d_pad, lane, w2);
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
// Cross-GPU barrier of the row-sharded global batch (one process per GPU; flags live in symmetric memory).
// Rank r bumps its local epoch, stores it into flags[r] of EVERY rank and waits until all `world` flags of its own
// copy have reached the epoch.  The kernels before it in the stream are complete (and their peer stores performed)
// when it runs, so everything they pushed is visible to the peers' kernels that follow their barrier.
// With `stats_all` it also adds up the per-rank loss statistics that the forward finalize kernels pushed.
// ---------------------------------------------------------------------------------------------
__global__ void __launch_bounds__(32) peer_barrier_kernel(PeerTable flags, unsigned int* __restrict__ epoch_local,
                                                          const float* __restrict__ stats_all,
                                                          float* __restrict__ stats_out, float* __restrict__ loss_out,
                                                          unsigned long long timeout_ns) {
    pdl_launch_dependents();
    pdl_wait();
    const int t = threadIdx.x;
    unsigned int target = 0;
    if (t == 0) {
        target = *epoch_local + 1u;
        *epoch_local = target;
    }
    target = __shfl_sync(0xffffffffu, target, 0);
    __threadfence_system();
    if (t < flags.world) {
        unsigned int* remote = static_cast<unsigned int*>(flags.ptr[t]) + flags.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(target) : "memory");
        peer_flag_wait(static_cast<const unsigned int*>(flags.ptr[flags.rank]) + t, target, timeout_ns, flags.rank, t);
    }
    __syncwarp();
    __threadfence_system();
    if (stats_all != nullptr && t == 0) {
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int r = 0; r < flags.world; ++r) {      // fixed order: identical result on every rank
            s0 += __ldcv(stats_all + 4 * r + 0);
            s1 += __ldcv(stats_all + 4 * r + 1);
            s2 += __ldcv(stats_all + 4 * r + 2);
        }
        stats_out[0] = s0;
        stats_out[1] = s1;
        stats_out[2] = s2;
        stats_out[3] = s0 / s1;
        if (loss_out) *loss_out = s0 / s1;
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
template <int kLoss, bool kFused>
__global__ void __launch_bounds__(kBlockM) forward_finalize_kernel(const TileParams p) {
    __shared__ float red[16];
    __shared__ int flags[4];
    pdl_launch_dependents();
    // parameter arithmetic (64-bit divisions) and the caller's row weight: before the wait, i.e. under the tile kernel
    const int rb = blockIdx.x, tid = threadIdx.x;
    const FinIndex ix = fin_index(p, rb, tid);
    float w_row = 1.f;
    if (p.row_weight != nullptr) {
        const int blocks_per_view = p.bl_pad / kBlockM;
        const int vr = rb / blocks_per_view;
        const int img = (rb - vr * blocks_per_view) * kBlockM + tid;
        if (img < p.b_loc) w_row = __ldg(p.row_weight + vr * p.b_loc + img);
    }
    pdl_wait();
    ktrace_begin(p.ktrace, 4);
    forward_finalize_rowblock<kLoss, kFused>(p, rb, tid, ix, w_row, red, flags);
    if (kFused && p.amb_done != nullptr) {
        // the list of rows for the exact re-scoring is complete when every block has counted itself: the blocks of the
        // backward finalize kernel that work it off start ahead of their grid dependency and wait for this count instead
        __threadfence();
        __syncthreads();
        if (tid == 0) atomicAdd(p.amb_done, 1u);
    }
    ktrace_end(p.ktrace, 4);
}

// ---------------------------------------------------------------------------------------------
// Stage 3 epilogue: exact positive-pair term + backward of the row normalisation (+ softplus).
// kBwdFinBlocksPerRowBlock blocks of 16 warps per row block: one warp per row.
// ---------------------------------------------------------------------------------------------
constexpr int kBwdFinBlocksPerRowBlock = 8;
template <int D, int kLoss, bool kDet>
__global__ void __launch_bounds__(512) backward_finalize_kernel(const TileParams p) {
    static_assert(16 * kBwdFinBlocksPerRowBlock == kBlockM, "one warp per row");
    pdl_launch_dependents();
    // The first kResolveBlocks blocks (fused one-GPU step, TileParams::resolve_ambiguous) re-score the rows the forward
    // finalize kernel listed; the others are the regular blocks, eight per row block.
    const int n_extra = p.resolve_ambiguous ? kResolveBlocks : 0;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (static_cast<int>(blockIdx.x) < n_extra) {
        // Nothing here depends on the backward tile kernel: the list, the candidates and the inputs come from the forward
        // finalize kernel, two launches upstream, so these blocks work while the tile kernel's last CTAs drain (they become
        // resident as soon as the first tile CTAs exit).  Programmatic launch does not order them behind that kernel,
        // though -- on a small problem they can start while it is still running: they wait until all of its blocks have
        // counted themselves (amb_done; those blocks were resident before any block of this kernel could be).  The block
        // that finishes last reduces the loss statistics (behind the grid dependency, like the regular path).
        __shared__ int last;
        if (threadIdx.x == 0) {
            unsigned int seen;
            do {
                asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(seen) : "l"(p.amb_done) : "memory");
                if (seen >= static_cast<unsigned int>(p.n_row_blocks)) break;
                __nanosleep(100);
            } while (true);
        }
        __syncthreads();
        if (__ldcg(p.amb_cnt) == 0u) return;               // nothing listed (the usual case): the regular path finishes
        resolve_ambiguous_rows<kLoss, (D <= 128 ? 1 : 2)>(p, static_cast<int>(blockIdx.x) * 16 + warp, kResolveBlocks * 16, lane);
        __threadfence();
        __syncthreads();
        if (threadIdx.x == 0) {
            last = atomicAdd(p.amb_ticket, 1u) == static_cast<unsigned int>(kResolveBlocks) - 1u ? 1 : 0;
            if (last) *p.amb_ticket = 0u;
        }
        __syncthreads();
        if (last && warp == 0) {
            __threadfence();
            pdl_wait();
            finish_forward_stats(p, lane);
        }
        return;
    }
    const int block = static_cast<int>(blockIdx.x) - n_extra;
    const int rb = block / kBwdFinBlocksPerRowBlock;
    const int sub = block % kBwdFinBlocksPerRowBlock;
    const int n_regular = p.n_row_blocks * kBwdFinBlocksPerRowBlock;
    if (p.ktrace != nullptr && threadIdx.x == 0) {
        pdl_wait();
        ktrace_begin(p.ktrace, 5);
    }
    constexpr int kPerLane = D / 32;
    float dz[kPerLane], dzx[kPerLane];
#pragma unroll
    for (int u = 0; u < kPerLane; ++u) dz[u] = dzx[u] = 0.f;
    backward_finalize_row<D, kLoss, kDet>(p, rb, sub * 16 + warp, lane, dz, dzx);      // waits for the tile kernel inside
    if (p.bn_partial != nullptr) {
        // projection-head tail: the column sums BatchNorm's backward needs (sum dz, sum dz * xhat over this CTA's 16 rows),
        // one partial per CTA -- bn_backward_kernel adds them in CTA order
        __shared__ float red[16][2][D];
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            red[warp][0][lane * kPerLane + u] = dz[u];
            red[warp][1][lane * kPerLane + u] = dzx[u];
        }
        __syncthreads();
        for (int idx = threadIdx.x; idx < 2 * D; idx += blockDim.x) {
            const int q = idx / D, k = idx - q * D;
            float a = 0.f;
#pragma unroll
            for (int w = 0; w < 16; ++w) a += red[w][q][k];
            p.bn_partial[(static_cast<size_t>(block) * 2 + q) * D + k] = a;
        }
    }
    if (p.finish_stats && block == n_regular - 1 && warp == 15) {
        pdl_wait();
        // (rows listed for the exact re-scoring: the extra blocks finish the statistics when they are through)
        if (!(p.resolve_ambiguous && __ldcg(p.amb_cnt) != 0u)) finish_forward_stats(p, lane);
    }
    ktrace_end(p.ktrace, 5);
}

}  // namespace simclr
