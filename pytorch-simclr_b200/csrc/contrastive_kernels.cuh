// NT-Xent and "modified" (probabilistic) contrastive losses for sm_100a -- the tile kernel.
//
// Reference semantics: objective.py:6-55 (contrastive_loss) and :58-98 (modified_contrastive_loss);
// closed forms in DESIGN.md section 3.
//
// Geometry ("view-padded" layout).  B images, two views.  Every per-row array and every operand matrix
// is laid out as [2][Bpad] with Bpad = ceil(B/128)*128 and zero padding, so that a 128-row block never
// mixes views.  Rows are this rank's shard (b_loc images, padded bl_pad), columns are the global batch
// (b_glob images, padded bg_pad).  Column index c in [0, 2*bg_pad): view vc = c / bg_pad, image
// ic = c % bg_pad, valid iff ic < b_glob.  For row (vr, image g):
//     NT-Xent : all columns of both views except itself (vr, g); positive = (1-vr, g)
//     modified: the columns of the other view only;              positive = (1-vr, g)
//
//   S[r,c]   = <op_r, op_c>              tcgen05.mma kind::f16, bf16 operands, fp32 accumulate in TMEM
//   forward  : online max / sum of exp2(score) over the valid negatives; the positive pair is excluded
//              here and added in exact fp32 by the finalize kernel; first-argmax bookkeeping
//   backward : W[r,c] (symmetric form) written as bf16 into TMEM over the consumed score
//              tile and used as the A operand of a second tcgen05.mma:  dacc[r,:] += W[r,:] * op[c,:]
// The 2N x 2N matrix never exists outside one 128 x 128 TMEM tile.
//
// Warp roles (640 threads, 1 CTA / SM, each CTA owns a contiguous range of (row block, column tile)):
//   warp 16     TMA producer (row-block tile once per segment, column tiles through an mbarrier ring)
//   warp 17     UMMA issuer for the score tiles (one elected lane)
//   warp 18     TMEM allocator / deallocator
//   warp 19     backward only: UMMA issuer for the gradient MMAs (a tcgen05.mma blocks its issuing thread while it
//               executes, so two issuers let the barrier waits of one overlap the MMAs of the other)
//   warps 0-15  four softmax warpgroups = two pairs.  CTA iteration `it` lives in TMEM score buffer it % NB
//               (NB = 4 forward, (512 - D) / 128 backward) and is consumed by pair it % 2, each warpgroup of the
//               pair taking 64 of the 128 columns; a pair therefore always has a second buffer being filled by
//               the tensor core while it works (the MMA runs ahead of the softmax)
#pragma once

#include <cstdint>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "sm100_ptx.cuh"

namespace simclr {

constexpr int kBlockM = 128;          // rows per row block (= UMMA M = TMEM lanes)
constexpr int kBlockN = 128;          // columns per tile   (= UMMA N of the score MMA)
constexpr int kAtomK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kAtomBytes = kBlockM * kAtomK * 2;   // one TMA box: 128 rows x 128 B = 16 KB
constexpr int kNumSoftmaxWG = 4;      // softmax warpgroups (4 warps each: one per TMEM lane quarter)
// The warp arbiter favours the highest warp id of an SM sub-partition, so the latency-critical single-thread
// roles (UMMA issue, TMA issue) sit ABOVE the 16 softmax warps; as warps 0/1 they were starved of issue slots
// and every batch of eight tcgen05.mma took >1000 cycles to issue (profiles/r01_timeline_before_after.md).
constexpr int kSoftmaxWarp0 = 0;
constexpr int kNumSoftmaxWarps = 4 * kNumSoftmaxWG;
constexpr int kProducerWarp = kNumSoftmaxWarps + 0;
constexpr int kMmaWarp = kNumSoftmaxWarps + 1;
constexpr int kAllocWarp = kNumSoftmaxWarps + 2;
constexpr int kGradWarp = kNumSoftmaxWarps + 3;   // backward: issues the gradient MMAs (the score MMAs stay on kMmaWarp)
constexpr int kNumThreads = 32 * (kNumSoftmaxWarps + 4);   // 640
constexpr int kMaxScoreBufs = 4;
constexpr int kNumPairs = kNumSoftmaxWG / 2;             // a tile is shared by a pair of warpgroups
constexpr int kMaxSlots = kMaxScoreBufs * kNumPairs;      // barrier slots for the score / W hand-off
constexpr int kTmemCols = 512;
constexpr int kFwdFields = 5;         // per-row partial: sum, run-max, max-preceding, max-following, pos(mma)
constexpr float kNegBig = -3.0e38f;   // finite stand-in for -inf
constexpr float kClampMin = 1e-4f;    // reference objective.py:87-88

enum LossKind : int { kNtXent = 0, kModified = 1 };

struct TileParams {
    int b_loc;         // images held by this rank
    int b_glob;        // images in the global batch
    int row_off;       // global index of this rank's first image
    int bl_pad;        // b_loc rounded up to 128
    int bg_pad;        // b_glob rounded up to 128
    int n_row_blocks;  // 2 * bl_pad / 128
    int n_col_tiles;   // column tiles per row block (NT-Xent: 2*bg_pad/128, modified: bg_pad/128)
    int max_segs;      // max number of row blocks one CTA touches
    long long total_tiles;
    float k2;          // NT-Xent: log2(e)/tau.  modified: 1/tau (scores are already log2)
    float m2;          // constant log2-domain shift of the one-exp backward form
    int const_shift;   // backward: 1 -> one exp per element (bounded scores), 0 -> general two-exp form
    float qscale;      // modified loss: (float) b_glob, the factor inside the clamp
    float* part;       // forward : [grid][max_segs][kNumSoftmaxWG][kFwdFields][128]
    const float* colvec;   // backward: [2 planes][2*bg_pad]; plane 0 = a_c (or g_c), plane 1 = lse2_c
    float* dacc;           // backward: [2*bl_pad][D] fp32, zero on entry, accumulated with red.global.add
    unsigned int* ticket;  // zeroed by CTA 0 for the finalize kernel's last-block reduction
    long long* trace;      // optional (debug): per-role clock64() timestamps of CTA `trace_cta`
    int trace_cta;
};

// Debug timeline: trace[(role * kTraceIters + it) * 4 + k].  Roles: 0 TMA producer, 1 MMA issuer,
// 2 + wg softmax warpgroup wg (lane 0 of its quarter-0 warp).
constexpr int kTraceIters = 64;
constexpr int kTraceRoles = 2 + 4;
SIMCLR_DEVICE void trace_event(const TileParams& p, int role, int it, int k) {
    if (p.trace != nullptr && static_cast<int>(blockIdx.x) == p.trace_cta && it < kTraceIters)
        p.trace[(role * kTraceIters + it) * 4 + k] = clock64();
}

template <int D>
struct SmemLayout {
    static constexpr int kAtoms = D / kAtomK;
    static constexpr int kTileBytes = kAtoms * kAtomBytes;          // 128 x D bf16
    static constexpr int kStages = (D <= 64) ? 8 : (D <= 128 ? 5 : 2);
    static constexpr int kColvecBytes = 2 * kBlockN * 4;             // two planes of 128 floats
    static constexpr int kOffA = 0;
    static constexpr int kOffB = kTileBytes;
    static constexpr int kOffCv = kOffB + kStages * kTileBytes;
    static constexpr int kOffBar = kOffCv + kStages * kColvecBytes;
    // barriers: a_full, a_empty, acc_full, acc_empty, b_full[S], b_empty[S], s_full[8], s_free[8], w_full[8], w_done[8]
    static constexpr int kNumBars = 4 + 2 * kStages + 4 * kMaxSlots;   // kMaxSlots = 8
    static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
    static constexpr int kBytes = kOffTmemPtr + 16;
    static constexpr int kDynamicBytes = kBytes + 1024;              // slack for manual 1024 B alignment
};

// First global column of tile j for a row block of view vr.
template <int kLoss>
SIMCLR_DEVICE int tile_col0(const TileParams& p, int vr, int j) {
    if constexpr (kLoss == kNtXent) return j * kBlockN;
    else return (1 - vr) * p.bg_pad + j * kBlockN;
}

// ---------------------------------------------------------------------------------------------
// Per-element maths.  "v" is the tracked raw value (monotone in the logit):
//   NT-Xent : v = S              logit2 = v * k2                     (k2 = log2(e)/tau)
//   modified: v = max(B*S, 1e-4) logit2 = log2(v) * k2               (k2 = 1/tau)
// ---------------------------------------------------------------------------------------------
template <int kLoss>
SIMCLR_DEVICE float raw_value(const TileParams& p, float s) {
    if constexpr (kLoss == kNtXent) return s;
    else return fmaxf(s * p.qscale, kClampMin);
}
template <int kLoss>
SIMCLR_DEVICE float logit2(const TileParams& p, float v) {
    if constexpr (kLoss == kNtXent) return v * p.k2;
    else return lg2_approx(v) * p.k2;
}

// ---------------------------------------------------------------------------------------------
// Softmax-warp helpers.  One thread owns one row (TMEM lane); a tile is consumed in four 32-column
// chunks with the TMEM load of chunk q+1 in flight while chunk q is processed.
// ---------------------------------------------------------------------------------------------
struct RowCtx {
    int vr;         // view of the row block
    int g;          // global image index of this thread's row
    int diag_col;   // global column of the row itself (NT-Xent only, -1 if none)
    int pos_col;    // global column of the positive (-1 for padding rows)
    bool row_ok;
};

struct FwdState {
    float run_max = kNegBig, sum = 0.f, max_prec = kNegBig, max_foll = kNegBig, pos_mma = kNegBig;
};

struct BwdRow {
    float row_a, row_l2;
};

SIMCLR_DEVICE float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// ---- forward: one 32-column chunk ----
template <int kLoss, bool kSpecial>
SIMCLR_DEVICE void fwd_chunk(const TileParams& p, const uint32_t (&r)[32], int cq, int vc, const RowCtx& rc,
                             bool tile_prec, FwdState& st) {
    float v[32];
    float cm = kNegBig;
    if constexpr (!kSpecial) {
#pragma unroll
        for (int i = 0; i < 32; ++i) v[i] = raw_value<kLoss>(p, __uint_as_float(r[i]));
#pragma unroll
        for (int i = 0; i < 32; i += 2) cm = fmaxf(cm, fmaxf(v[i], v[i + 1]));     // FMNMX3
        st.max_prec = fmaxf(st.max_prec, tile_prec ? cm : kNegBig);
        st.max_foll = fmaxf(st.max_foll, tile_prec ? kNegBig : cm);
    } else {
#pragma unroll
        for (int i = 0; i < 32; ++i) {
            const int c = cq + i;
            const int ic = c - vc * p.bg_pad;
            const float x = raw_value<kLoss>(p, __uint_as_float(r[i]));
            if (c == rc.pos_col) st.pos_mma = x;
            const bool valid = rc.row_ok && ic < p.b_glob && c != rc.diag_col && c != rc.pos_col;
            v[i] = valid ? x : kNegBig;
            if (valid) {
                cm = fmaxf(cm, x);
                bool prec;
                if constexpr (kLoss == kNtXent) prec = (rc.vr == 0) ? (vc == 1 && ic < rc.g) : (vc == 1 || ic < rc.g);
                else prec = ic < rc.g;
                if (prec) st.max_prec = fmaxf(st.max_prec, x); else st.max_foll = fmaxf(st.max_foll, x);
            }
        }
    }
    // online max without a data-dependent branch in the fast path: rescale the running sum by
    // exp2(old_shift - new_shift) (== 1 when the maximum did not move)
    const float new_max = fmaxf(st.run_max, cm);
    if (!kSpecial || new_max != kNegBig) {                 // special tiles may have seen nothing valid yet
        const float shift = logit2<kLoss>(p, new_max);
        const float old_shift = (st.run_max == kNegBig) ? shift : logit2<kLoss>(p, st.run_max);
        st.sum *= ex2_approx(old_shift - shift);
        st.run_max = new_max;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < 32; i += 4) {
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                if constexpr (kLoss == kNtXent) e[u] = ex2_approx(fmaf(v[i + u], p.k2, -shift));
                else e[u] = ex2_approx(fmaf(lg2_approx(v[i + u]), p.k2, -shift));
                if constexpr (kSpecial) e[u] = (v[i + u] == kNegBig) ? 0.f : e[u];
            }
            a0 += e[0];
            a1 += e[1];
            a2 += e[2];
            a3 += e[3];
        }
        st.sum += (a0 + a1) + (a2 + a3);
    }
}

// ---- backward: one 32-column chunk -> 16 packed bf16x2 words of W ----
template <int kLoss, bool kConst, bool kSpecial>
SIMCLR_DEVICE void bwd_chunk(const TileParams& p, const uint32_t (&r)[32], uint32_t cv_addr, int cq, const RowCtx& rc,
                             const BwdRow& br, uint32_t (&w)[16]) {
#pragma unroll
    for (int i = 0; i < 32; i += 4) {
        const float4 ac = lds_f4(cv_addr + i * 4);
        float4 lc = make_float4(0.f, 0.f, 0.f, 0.f);
        if constexpr (!kConst) lc = lds_f4(cv_addr + kBlockN * 4 + i * 4);
        const float acs[4] = {ac.x, ac.y, ac.z, ac.w};
        const float lcs[4] = {lc.x, lc.y, lc.z, lc.w};
        float wv[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const float sraw = __uint_as_float(r[i + u]);
            float wval;
            if constexpr (kLoss == kNtXent) {
                if constexpr (kConst) {
                    // one exp per element: W = exp2(S*k2 - m2) * (a_r + a_c)
                    wval = ex2_approx(fmaf(sraw, p.k2, -p.m2)) * (br.row_a + acs[u]);
                } else {
                    // general form: W = g_r exp2(S*k2 - lse2_r) + g_c exp2(S*k2 - lse2_c)
                    wval = br.row_a * ex2_approx(fmaf(sraw, p.k2, -br.row_l2)) +
                           acs[u] * ex2_approx(fmaf(sraw, p.k2, -lcs[u]));
                }
            } else {
                // d/dP of log(max(B P,1e-4))/tau = 1/(tau P) where live; folded: e^{A}/P = B q^{1/tau - 1}
                const float qv = sraw * p.qscale;
                const float y = lg2_approx(fmaxf(qv, kClampMin)) * (p.k2 - 1.0f);
                if constexpr (kConst) wval = ex2_approx(y - p.m2) * (br.row_a + acs[u]);
                else wval = br.row_a * ex2_approx(y - br.row_l2) + acs[u] * ex2_approx(y - lcs[u]);
                wval = (qv >= kClampMin) ? wval : 0.f;
            }
            if constexpr (kSpecial) {
                const int c = cq + i + u;
                if (c == rc.diag_col || c == rc.pos_col) wval = 0.f;
            }
            wv[u] = wval;
        }
        w[(i >> 1) + 0] = pack_bf16x2(wv[0], wv[1]);
        w[(i >> 1) + 1] = pack_bf16x2(wv[2], wv[3]);
    }
}

// Walks a CTA's contiguous tile range without per-tile 64-bit divisions.
struct TileWalker {
    int rb, j, nct, idx, n;
    SIMCLR_DEVICE TileWalker(long long t_begin, long long t_end, int nct_) : nct(nct_), idx(0) {
        rb = static_cast<int>(t_begin / nct_);
        j = static_cast<int>(t_begin - static_cast<long long>(rb) * nct_);
        n = static_cast<int>(t_end - t_begin);
    }
    SIMCLR_DEVICE bool valid() const { return idx < n; }
    SIMCLR_DEVICE bool seg_first() const { return idx == 0 || j == 0; }
    SIMCLR_DEVICE bool seg_last() const { return idx == n - 1 || j == nct - 1; }
    SIMCLR_DEVICE void next() {
        ++idx;
        if (++j == nct) {
            j = 0;
            ++rb;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// The tile kernel
// ---------------------------------------------------------------------------------------------
template <int D, int kLoss, bool kBackward>
__global__ void __launch_bounds__(kNumThreads, 1)
contrastive_tile_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                        const TileParams p) {
    using L = SmemLayout<D>;
    constexpr int S = L::kStages;
    constexpr uint32_t kIdescScore = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t kIdescGrad = make_idesc_bf16(kBlockM, D, 0, 1);
    // TMEM: NB score buffers of 128 columns, then (backward) the D-column gradient accumulator
    constexpr int NB = kBackward ? ((kTmemCols - D) / kBlockN > kMaxScoreBufs ? kMaxScoreBufs : (kTmemCols - D) / kBlockN)
                                 : kMaxScoreBufs;
    constexpr uint32_t kTmemAcc = NB * kBlockN;
    constexpr int kSlots = NB * kNumPairs;            // slot -> fixed (pair, buffer)

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* smem_a = smem + L::kOffA;
    uint8_t* smem_b = smem + L::kOffB;
    float* smem_cv = reinterpret_cast<float*>(smem + L::kOffCv);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kOffBar);
    uint64_t* a_full = bars + 0;
    uint64_t* a_empty = bars + 1;
    uint64_t* acc_full = bars + 2;
    uint64_t* acc_empty = bars + 3;
    uint64_t* b_full = bars + 4;
    uint64_t* b_empty = b_full + S;
    uint64_t* s_full = b_empty + S;               // [kSlots] score tile ready in TMEM
    uint64_t* s_free = s_full + kMaxSlots;        // [kSlots] forward only: softmax finished reading the score tile
    uint64_t* w_full = s_free + kMaxSlots;        // [kSlots] backward only: W written to TMEM
    uint64_t* w_done = w_full + kMaxSlots;        // [kSlots] backward only: gradient MMAs finished reading W
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);

    // warp index through a shuffle so that the compiler knows it is warp-uniform (role branches stay uniform and
    // tcgen05 / TMA operands can live in uniform registers instead of per-instruction ELECT/R2UR waterfall loops)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    // contiguous tile range of this CTA
    const long long t_begin = (p.total_tiles * blockIdx.x) / gridDim.x;
    const long long t_end = (p.total_tiles * (blockIdx.x + 1)) / gridDim.x;
    const int nct = p.n_col_tiles;
    const int blocks_per_view = p.bl_pad / kBlockM;

    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_rows);
        tma_prefetch_desc(&tmap_cols);
        if (blockIdx.x == 0 && p.ticket != nullptr) *p.ticket = 0u;
    }
    if (warp == kMmaWarp && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, 1);
        mbar_init(acc_full, 1);
        mbar_init(acc_empty, 128 * kNumSoftmaxWG);
        for (int i = 0; i < S; ++i) {
            mbar_init(b_full + i, 1);
            mbar_init(b_empty + i, 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(s_full + i, 1);
            mbar_init(s_free + i, 256);
            mbar_init(w_full + i, 256);
            mbar_init(w_done + i, 1);
        }
        mbar_fence_init();
    }
    if (warp == kAllocWarp) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem_base = *tmem_ptr_smem;

    if (warp == kProducerWarp) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            int seg = 0;
            for (TileWalker w(t_begin, t_end, nct); w.valid(); w.next()) {
                const int it = w.idx;
                if (w.seg_first()) {
                    if (seg > 0) mbar_wait(a_empty, (seg - 1) & 1, 100);
                    mbar_arrive_expect_tx(a_full, L::kTileBytes);
#pragma unroll
                    for (int ka = 0; ka < L::kAtoms; ++ka)
                        tma_load_2d(smem_a + ka * kAtomBytes, &tmap_rows, a_full, ka * kAtomK, w.rb * kBlockM);
                    ++seg;
                }
                const int stage = it % S;
                const int use = it / S;
                trace_event(p, 0, it, 0);
                if (use > 0) mbar_wait(b_empty + stage, (use - 1) & 1, 101);
                trace_event(p, 0, it, 1);
                const int c0 = tile_col0<kLoss>(p, w.rb / blocks_per_view, w.j);
                mbar_arrive_expect_tx(b_full + stage, L::kTileBytes + (kBackward ? L::kColvecBytes : 0));
#pragma unroll
                for (int ka = 0; ka < L::kAtoms; ++ka)
                    tma_load_2d(smem_b + stage * L::kTileBytes + ka * kAtomBytes, &tmap_cols, b_full + stage,
                                ka * kAtomK, c0);
                if constexpr (kBackward) {
                    float* cv = smem_cv + stage * 2 * kBlockN;
                    bulk_load_1d(cv, p.colvec + c0, kBlockN * 4, b_full + stage);
                    bulk_load_1d(cv + kBlockN, p.colvec + 2 * p.bg_pad + c0, kBlockN * 4, b_full + stage);
                }
            }
        }
    } else if (warp == kMmaWarp) {
        // ================================ UMMA issuer: score tiles ================================
        // S[buf] = A * B_stage^T (both operands K-major).  The whole warp walks the loop and waits on the
        // barriers; one elected lane issues the tcgen05 ops (operands stay in uniform registers).
        const uint32_t a_addr = smem_u32(smem_a);
        const uint32_t b_addr0 = smem_u32(smem_b);
        int seg_seen = 0;
        for (TileWalker w(t_begin, t_end, nct); w.valid(); w.next()) {
            const int idx = w.idx;
            const int stage = idx % S;
            const int buf = idx % NB;
            if (lane == 0) trace_event(p, 1, idx, 0);
            if (w.seg_first()) {
                mbar_wait(a_full, seg_seen & 1, 200);
                ++seg_seen;
            }
            mbar_wait(b_full + stage, (idx / S) & 1, 201);
            if (idx >= NB) {
                // the buffer's previous tenant (tile idx-NB) must be finished: read by the softmax (forward) or
                // consumed as W by the gradient MMAs (backward)
                uint64_t* freed = (kBackward ? w_done : s_free) + (idx - NB) % kSlots;
                mbar_wait(freed, ((idx - NB) / kSlots) & 1, 202);
            }
            tc_fence_after_sync();
            if (lane == 0) trace_event(p, 1, idx, 1);
            if (elect_one()) {
                const uint32_t b_addr = b_addr0 + stage * L::kTileBytes;
#pragma unroll
                for (int ka = 0; ka < L::kAtoms; ++ka) {
#pragma unroll
                    for (int kk = 0; kk < 4; ++kk) {
                        const uint32_t off = ka * kAtomBytes + kk * 32;
                        umma_ss(tmem_base + buf * kBlockN, make_smem_desc(a_addr + off, 0, 1024),
                                make_smem_desc(b_addr + off, 0, 1024), kIdescScore, (ka | kk) != 0);
                    }
                }
                // s_full is committed last so that, when the softmax sees it, the other arrivals have landed
                if constexpr (!kBackward) umma_commit(b_empty + stage);
                if (w.seg_last()) umma_commit(a_empty);
                umma_commit(s_full + idx % kSlots);
            }
            __syncwarp();
        }
    } else if (kBackward && warp == kGradWarp) {
        // ================================ UMMA issuer: gradient MMAs ================================
        // acc += W[buf] (TMEM, 128 x 128 bf16) * B_stage (MN-major: K = column index)
        const uint32_t b_addr0 = smem_u32(smem_b);
        int seg_done = 0;     // segments whose accumulator has been handed to the flush
        for (TileWalker w(t_begin, t_end, nct); w.valid(); w.next()) {
            const int idx = w.idx;
            const int stage = idx % S;
            const int buf = idx % NB;
            const bool seg_first = w.seg_first();
            if (seg_first && seg_done > 0) mbar_wait(acc_empty, (seg_done - 1) & 1, 203);
            if (lane == 0) trace_event(p, 1, idx, 2);
            mbar_wait(w_full + idx % kSlots, (idx / kSlots) & 1, 204);
            tc_fence_after_sync();
            if (lane == 0) trace_event(p, 1, idx, 3);
            if (elect_one()) {
                const uint32_t b_addr = b_addr0 + stage * L::kTileBytes;
#pragma unroll
                for (int kc = 0; kc < kBlockN / 16; ++kc) {
                    // W of column half h (K chunks 4h..4h+3) sits at columns [64h, 64h+32) of the buffer
                    umma_ts(tmem_base + kTmemAcc, tmem_base + buf * kBlockN + (kc >> 2) * 64 + (kc & 3) * 8,
                            make_smem_desc(b_addr + kc * 2048, kAtomBytes, 1024), kIdescGrad,
                            !(seg_first && kc == 0));
                }
                umma_commit(b_empty + stage);            // B tile (and its column vectors) may be overwritten
                umma_commit(w_done + idx % kSlots);      // score buffer may be overwritten
                if (w.seg_last()) umma_commit(acc_full);
            }
            __syncwarp();
            if (w.seg_last()) ++seg_done;
        }
    } else if (warp < kNumSoftmaxWarps) {
        // ================================ softmax warpgroups ================================
        const int wg = (warp - kSoftmaxWarp0) >> 2;          // 0 .. kNumSoftmaxWG-1
        const int pair = wg >> 1;                             // which tiles (it % 2)
        const int half = wg & 1;                              // which 64 columns of the tile
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may touch
        const int row_in_block = quarter * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
        const int n = static_cast<int>(t_end - t_begin);
        const uint32_t cv_base = smem_u32(smem_cv);

        int idx = 0;
        int seg = 0;
        while (idx < n) {
            // ---- segment = run of tiles sharing one row block ----
            const long long t0 = t_begin + idx;
            const int rb = static_cast<int>(t0 / nct);
            const int j0 = static_cast<int>(t0 % nct);
            const int seg_len = min(n - idx, nct - j0);

            RowCtx rc;
            rc.vr = rb / blocks_per_view;                                            // view of this row block
            const int img = (rb - rc.vr * blocks_per_view) * kBlockM + row_in_block; // local image index
            rc.row_ok = img < p.b_loc;
            rc.g = p.row_off + img;                                                  // global image index
            rc.diag_col = (kLoss == kNtXent && rc.row_ok) ? rc.vr * p.bg_pad + rc.g : -1;
            rc.pos_col = rc.row_ok ? (1 - rc.vr) * p.bg_pad + rc.g : -1;
            // warp-uniform description of this warp's 32 rows
            const int img_lo = (rb - rc.vr * blocks_per_view) * kBlockM + quarter * 32;
            const bool warp_rows_ok = img_lo + 31 < p.b_loc;
            const int g_lo = p.row_off + img_lo, g_hi = g_lo + 31;

            FwdState fs;
            BwdRow br;
            br.row_a = 0.f;
            br.row_l2 = 0.f;
            if constexpr (kBackward) {
                if (rc.row_ok) {
                    br.row_a = __ldg(p.colvec + rc.vr * p.bg_pad + rc.g);
                    br.row_l2 = __ldg(p.colvec + 2 * p.bg_pad + rc.vr * p.bg_pad + rc.g);
                }
            }

            for (int s = 0; s < seg_len; ++s) {
                const int it = idx + s;
                if ((it % kNumPairs) != pair) continue;
                const int c0 = tile_col0<kLoss>(p, rc.vr, j0 + s);
                const int vc = c0 >= p.bg_pad ? 1 : 0;            // a tile never mixes views
                const int ic0 = c0 - vc * p.bg_pad;               // image index of the tile's first column
                const int buf = it % NB;
                const int stage = it % S;
                const uint32_t tmem_tile = tmem_base + lane_addr + buf * kBlockN;

                // ---- warp-uniform classification (first-argmax rule: does the tile precede the positive in the
                // reference's column order?  NT-Xent rows see [view-2 block | view-1 block] (objective.py:48-49);
                // modified rows see the other view in natural order (objective.py:93)) ----
                bool tile_prec = false;
                if constexpr (!kBackward) {
                    const bool before = ic0 + kBlockN - 1 < g_lo;
                    if constexpr (kLoss == kNtXent) tile_prec = (rc.vr == 0) ? (vc == 1 && before) : (vc == 1 || before);
                    else tile_prec = before;
                }
                // One barrier per slot = it % (NB * #warpgroups): a slot always maps to the same warpgroup pair and the
                // same TMEM buffer, so every barrier has a single waiter that observes all of its phases in order
                // (a parity wait is only meaningful when the waiter is at most one phase away from the barrier).
                const int slot = it % kSlots;
                if (quarter == 0 && lane == 0) trace_event(p, 2 + wg, it, 0);
                mbar_wait(s_full + slot, (it / kSlots) & 1, 300);
                if (quarter == 0 && lane == 0) trace_event(p, 2 + wg, it, 1);
                if constexpr (kBackward) mbar_wait(b_full + stage, (it / S) & 1, 301);   // colvec visibility
                tc_fence_after_sync();

                // this warpgroup's two 32-column chunks; "special" = the chunk may contain the diagonal, the positive
                // or padded columns for one of this warp's 32 rows (warp-uniform)
#pragma unroll 1
                for (int qq = 0; qq < 2; ++qq) {
                    const int q = half * 2 + qq;
                    const int icq = ic0 + q * 32;
                    bool special = !warp_rows_ok || !(icq > g_hi || icq + 31 < g_lo);
                    if constexpr (!kBackward) special = special || (icq + 31 >= p.b_glob);
                    uint32_t r[32];
                    tmem_ld32(tmem_tile + q * 32, r);
                    tmem_ld_wait();
                    if constexpr (!kBackward) {
                        if (special) fwd_chunk<kLoss, true>(p, r, c0 + q * 32, vc, rc, tile_prec, fs);
                        else fwd_chunk<kLoss, false>(p, r, c0 + q * 32, vc, rc, tile_prec, fs);
                    } else {
                        const uint32_t cv_addr = cv_base + stage * (2 * kBlockN * 4) + q * 128;
                        uint32_t w[16];
                        if (p.const_shift) {
                            if (special) bwd_chunk<kLoss, true, true>(p, r, cv_addr, c0 + q * 32, rc, br, w);
                            else bwd_chunk<kLoss, true, false>(p, r, cv_addr, c0 + q * 32, rc, br, w);
                        } else {
                            if (special) bwd_chunk<kLoss, false, true>(p, r, cv_addr, c0 + q * 32, rc, br, w);
                            else bwd_chunk<kLoss, false, false>(p, r, cv_addr, c0 + q * 32, rc, br, w);
                        }
                        // bf16 W of chunk q goes to columns [64*half + 16*qq, +16): inside this warpgroup's own
                        // 64-column region and already consumed by this thread
                        tmem_st16(tmem_tile + half * 64 + qq * 16, w);
                    }
                }
                if constexpr (!kBackward) {
                    tc_fence_before_sync();
                    mbar_arrive(s_free + slot);
                } else {
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(w_full + slot);
                }
                if (quarter == 0 && lane == 0) trace_event(p, 2 + wg, it, 2);
            }

            // ---- end of segment ----
            if constexpr (!kBackward) {
                float* dst = p.part + ((static_cast<size_t>(blockIdx.x) * p.max_segs + seg) * kNumSoftmaxWG + wg) *
                                          (kFwdFields * kBlockM) + row_in_block;
                dst[0 * kBlockM] = fs.sum;
                dst[1 * kBlockM] = fs.run_max;
                dst[2 * kBlockM] = fs.max_prec;
                dst[3 * kBlockM] = fs.max_foll;
                dst[4 * kBlockM] = fs.pos_mma;
            } else {
                // flush the gradient accumulator: the D columns are split in 32-column chunks over the warpgroups
                mbar_wait(acc_full, seg & 1, 302);
                tc_fence_after_sync();
#pragma unroll 1
                for (int q = wg; q < D / 32; q += kNumSoftmaxWG) {
                    uint32_t r[32];
                    const int col = q * 32;
                    tmem_ld32(tmem_base + lane_addr + kTmemAcc + col, r);
                    tmem_ld_wait();
                    if (rc.row_ok) {
                        float* dst = p.dacc + static_cast<size_t>(rb * kBlockM + row_in_block) * D + col;
#pragma unroll
                        for (int i = 0; i < 32; i += 4)
                            red_add_v4(dst + i, __uint_as_float(r[i]), __uint_as_float(r[i + 1]),
                                       __uint_as_float(r[i + 2]), __uint_as_float(r[i + 3]));
                    }
                }
                tc_fence_before_sync();
                mbar_arrive(acc_empty);
            }
            idx += seg_len;
            ++seg;
        }
    }

    // ================================ teardown ================================
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem_base, kTmemCols);
    }
}

}  // namespace simclr
