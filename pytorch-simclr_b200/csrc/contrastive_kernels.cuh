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
// Warp roles (forward 640 threads, backward 768; 1 CTA / SM, each CTA owns a contiguous range of (row block, column tile)):
//   warps 0-15  four softmax warpgroups = two pairs.  CTA tile `it` lives in TMEM score buffer it % NB
//               (NB = 4 forward, (512 - D) / 128 backward) and is consumed by pair it % 2, each warpgroup of the
//               pair taking 64 of the 128 columns; a pair therefore always has a second buffer being filled by
//               the tensor core while it works (the MMA runs ahead of the softmax)
//   warp 16     TMA producer (row-block tile once per segment, column tiles through an mbarrier ring)
//   warps 17,18 UMMA issuers for the score tiles (one elected lane each; warp 18 also owns the TMEM allocation)
//   warps 19,20 backward only: UMMA issuers for the gradient MMAs
//   warp 19     forward only, spare: zeroes the backward's accumulation buffer; row-sharded fused step: pushes this CTA's
//               share of the rank's operand rows to all ranks and publishes the barrier epoch (TileParams::n_windows)
//   warps 20-23 backward only: flush warpgroup (one warp per TMEM lane quarter; warp 20 doubles as an issuer): at the
//               end of a segment it drains the gradient accumulator TMEM -> 128B-swizzled fp32 boxes in two ring stages
//               the producer hands over -> TMA reduce-add into dacc, and re-zeroes the accumulator.  The softmax
//               warps never stop at a segment boundary.
// A tcgen05.mma blocks its issuing thread while it executes and the issuer's barrier waits cost several hundred
// cycles per tile, so ONE issuer keeps the tensor pipe busy only ~35 % of the time (profiles/r01_notes.md).  Issuer k of
// a kind owns the ring stages with (stage & 1) == k -- the ring size is even and the flush borrows two stages, so a
// stage, a hand-off slot and a softmax pair always meet the same issuer and every mbarrier keeps a single waiter
// that sees all of its phases in order.
#pragma once

#include <cstdint>
#include <type_traits>
#include <cuda.h>
#include <cuda_bf16.h>
#include <cuda_runtime.h>

#include "row_math.cuh"
#include "sm100_ptx.cuh"

namespace simclr {

constexpr int kBlockM = 128;          // rows per row block (= UMMA M = TMEM lanes)
constexpr int kBlockN = 128;          // columns per tile   (= UMMA N of the score MMA)
constexpr int kAtomK = 64;            // bf16 elements per 128-byte swizzle row
constexpr int kAtomBytes = kBlockM * kAtomK * 2;   // one TMA box: 128 rows x 128 B = 16 KB
constexpr int kNumSoftmaxWG = 4;      // softmax warpgroups (4 warps each: one per TMEM lane quarter)
// The warp arbiter favours the highest warp id of an SM sub-partition, so the latency-critical single-thread
// roles (UMMA issue, TMA issue) sit ABOVE the 16 softmax warps; as warps 0/1 they were starved of issue slots
// and every batch of eight tcgen05.mma took >1000 cycles to issue (profiles/r01_notes.md).
constexpr int kSoftmaxWarp0 = 0;
constexpr int kNumSoftmaxWarps = 4 * kNumSoftmaxWG;
constexpr int kProducerWarp = kNumSoftmaxWarps + 0;
constexpr int kNumIssuers = 2;                          // per kind (score / gradient)
constexpr int kScoreWarp0 = kNumSoftmaxWarps + 1;       // warps 17, 18
constexpr int kAllocWarp = kScoreWarp0 + 1;             // warp 18 allocates TMEM before it starts issuing
constexpr int kGradWarp0 = kScoreWarp0 + kNumIssuers;   // warps 19, 20 (backward)
constexpr int kThreadsForward = 32 * (kNumSoftmaxWarps + 4);                    // 640 (warp 19 idles)
// One gradient issuer (warp 20) doubles as a flush warp.  Separate warps (25 warps, SIMCLR_BWD_SHARED_FLUSH_WARPS 0) cost
// registers instead of freeing them: warps are allocated in groups of four, so ptxas budgets 28 warps and caps the kernel
// at 72 registers (spills reach the softmax loop); sharing BOTH (23 warps) buys none either -- 80 registers per thread is
// the limit down to 22 warps.  The tile walk is specialised per warp instead (pure issuer / both / pure flusher).
#ifndef SIMCLR_BWD_SHARED_FLUSH_WARPS
#define SIMCLR_BWD_SHARED_FLUSH_WARPS 1
#endif
constexpr int kFlushWarp0 = kGradWarp0 + kNumIssuers - SIMCLR_BWD_SHARED_FLUSH_WARPS;   // warps 20..23: TMEM lane quarters 0..3
constexpr int kFlushIssueWarp = kFlushWarp0 + 2;            // a flush warp that is not a gradient issuer issues the TMA reductions
constexpr int kThreadsBackward = 32 * (kFlushWarp0 + 4);    // 768
constexpr int kFlushBar = 4;                                // named barrier of the flush warpgroup
constexpr int kMaxScoreBufs = 4;
constexpr int kNumPairs = kNumSoftmaxWG / 2;             // a tile is shared by a pair of warpgroups
constexpr int kMaxSlots = kMaxScoreBufs * kNumPairs;      // barrier slots for the score / W hand-off
constexpr int kTmemCols = 512;
#ifndef SIMCLR_TOKEN_CHUNK
#define SIMCLR_TOKEN_CHUNK 3          // chunk (0..3) before which a pair passes the ping-pong token on; 4 = after the tile
#endif
#ifndef SIMCLR_TOKEN_CHUNK_FWD
#define SIMCLR_TOKEN_CHUNK_FWD SIMCLR_TOKEN_CHUNK      // the forward kernel's own setting
#endif
constexpr int kTokenBar0 = 2;         // named barriers 2, 3: ping-pong tokens of the two softmax pairs
constexpr int kFwdFields = 5;         // per-row partial: sum, run-max, max-preceding, max-following, pos(mma)
constexpr float kNegBig = -3.0e38f;   // finite stand-in for -inf
constexpr float kClampMin = 1e-4f;    // reference objective.py:87-88
// Exact accuracy count in bf16 mode (reference objective.py:51-53 / :95-97): the tensor-core score of a negative differs
// from its exact value by at most kBandRel * (largest possible |score|), so a row's first-argmax decision taken on
// tensor-core scores is certain unless its best negative lies inside that band around the EXACT positive score.  The
// forward tile kernel records, per row, the (few) 16-column chunks whose maximum fell inside the band while the row was
// still undecided; the finalize kernel re-scores the columns of those chunks in exact fp32.
// bf16 round-to-nearest operands: relative error 2^-9 each, so |sum a_i b_i - sum a~_i b~_i| <= (2^-8 + 2^-18) sum|a_i b_i|
// <= 2^-8 (1 + 2^-10) |a||b| (Cauchy-Schwarz); fp32 accumulation of d <= 256 terms adds < 2^-15 of that.  2 % margin.
constexpr float kBandRel = 1.02f / 256.0f;
constexpr int kCandMax = 8;           // candidate ranges kept per row; a row with more falls back to the tensor-core decision
constexpr int kCandRange = 16;        // columns per candidate range (one chunk of the softmax warps' tile walk)
// NT-Xent with normalised rows: |S'| <= k2 (+ bf16 rounding) and k2 <= 40 is required for this path, so exp2(S') stays
// far inside the fp32 range: no running maximum and no shift at all ("constant shift" of zero)
constexpr float kConstShiftRaw = 0.f;

enum LossKind : int { kNtXent = 0, kModified = 1 };

// Peer (NVLink / NVSwitch) addressing of a symmetric buffer: ptr[r] is THIS process's mapping of rank r's copy.
// world == 0 switches the peer stores off.  Passed by value inside the kernel parameters.
constexpr int kMaxPeers = 16;
struct PeerTable {
    void* ptr[kMaxPeers];
    void* mc;      // multicast (NVLS) mapping of the same symmetric buffer, or nullptr: one multimem.st reaches every rank
    int world;
    int rank;
};

// One forward tile-kernel launch as the finalize kernel sees it: where its per-CTA partials are and how its tiles were
// dealt out.  The overlapped row-sharded forward runs two launches (local columns, then remote columns).
struct PartSet {
    const float* part;
    long long total_tiles;
    int n_col_tiles, max_segs, grid;
};

// One column window of a forward tile-kernel launch (see TileParams::n_windows).
struct ColWindow {
    int col_start, col_cnt, n_col_tiles, max_segs;
    long long total_tiles;
    float* part;
};

struct TileParams {
    int b_loc;         // images held by this rank
    int b_glob;        // images in the global batch
    int row_off;       // global index of this rank's first image
    int bl_pad;        // b_loc rounded up to 128
    int bg_pad;        // b_glob rounded up to 128
    int n_row_blocks;  // 2 * bl_pad / 128
    int n_col_tiles;   // column tiles per row block of THIS launch (NT-Xent: 2*col_cnt, modified: col_cnt)
    // column window of this launch: per view the col_cnt tiles starting at view-tile col_start (cyclic in the view's
    // tiles_per_view = bg_pad/128 tiles).  Full problem: col_start = 0, col_cnt = tiles_per_view.
    int col_start, col_cnt, tiles_per_view;
    int max_segs;      // max number of row blocks one CTA touches
    long long total_tiles;
    float k2;          // NT-Xent: log2(e)/tau (carried by the operands as sqrt(k2) each).  modified: 1/tau
    float acc_scale;   // backward finalize: 1/sqrt(k2) undoes the operand factor in dacc = W * operand (modified: 1)
    float m2;          // constant log2-domain shift of the one-exp backward form
    int const_shift;   // backward: 1 -> one exp per element (bounded scores), 0 -> general two-exp form
    int pow;           // modified loss: 1 / 2 when 1/tau is exactly 1 / 2 (exp(A) = (B P)^(1/tau) is then a plain power: no
                       // transcendental at all), else 0
    float qscale;      // modified loss: (float) b_glob, the factor inside the clamp
    float inv_tau;
    float tau;         // the caller's temperature
    int d;             // true embedding dimension (<= D)
    int in_bf16;       // element type of x_batch / grad: 0 = f32, 1 = bf16
    int normalize;
    unsigned int* ticket;      // forward finalize: row blocks finished so far (zero on entry, left zero on exit)
    // forward
    float* part;               // [grid][max_segs][kFwdFields][128] per-CTA partials (warpgroups merged in smem)
    const float* pos_dot;      // [2*bl_pad] exact fp32 positive-pair dot products
    const float* row_weight;   // [2*b_loc] compact order, or nullptr
    float* lse2;               // [2*bl_pad]
    float* row_loss;           // [2*bl_pad]
    float* block_part;         // [n_row_blocks][4]
    float* stats;              // [4]
    float* loss_out;           // [1] or nullptr
    // backward
    const float* colvec;       // [2 planes][2*bg_pad]; plane 0 = a_c (or g_c), plane 1 = lse2_c
    float* dacc;               // [2*bl_pad][D] fp32, zero on entry, accumulated with TMA reduce-add (the 2 - 3 CTAs that
                               // share a row block add in the order they happen to finish: last-bit run-to-run noise)
    // Deterministic mode (TileParams::deterministic): every (CTA, segment) instead STORES its accumulator into its own
    // 128-row slot of det_part [grid * max_segs * 128][D] and the backward finalize kernel adds the slots of a row block
    // in CTA order -- bit-identical gradients from run to run (the reference's cudnn.deterministic switch,
    // pretrain.py:59-61), for ~2.3x the accumulator traffic.
    const float* det_part;
    int deterministic;
    // projection-head tail fused into the loss (head_kernels.cuh): x1 / x2 are the PRE-BatchNorm activations;
    // bn_state f32 [2][5][D] (scale | shift | mean | rstd | var_unbiased per view), bn_partial f32 [finalize CTAs][2][D]
    const float* bn_state;
    float* bn_partial;
    const void* x1;            // inputs [b_loc][d]
    const void* x2;
    void* g1;                  // gradients [b_loc][d]
    void* g2;
    const float* inv_norm;     // [2*bl_pad]
    const float* col_scale;    // [2*bg_pad] w_c / sum(w) or nullptr
    const float* grad_out;     // [1] or nullptr
    // row-sharded global batch over peer memory: every rank's forward finalize pushes its rows' lse2 and its three
    // loss statistics into all ranks' global buffers (no collective call)
    PeerTable colvec_peers;    // float [2 planes][2*bg_pad] per rank: the backward's column vectors (a_c | lse2_c)
    PeerTable stats_peers;     // float [world][4] per rank
    // "priming" the backward from the forward (saves the backward-prepare kernel): the finalize kernel writes the column
    // vectors, an idle warp of the forward tile kernel zeroes the gradient accumulation buffer
    float* prime_colvec;       // local [2][2*bg_pad] or nullptr
    float4* prime_dacc;        // or nullptr
    unsigned long long prime_dacc_vec4;
    long long* trace;      // optional (debug): per-role clock64() timestamps of CTA `trace_cta`
    int trace_cta;
    int tile_grid;         // grid size of the tile kernel (the finalize kernels need it to locate partials)
    PartSet fin_set[2];    // forward finalize: the launches whose partials it merges
    int n_fin_sets;
    unsigned long long* ktrace;   // optional (debug): kernel-level %globaltimer stamps
    // Backward tile kernel: the operand matrix was complete before the kernel's predecessors started (it is written by
    // the prepare kernel, two or more programmatic launches upstream, and the tile kernel of the forward -- which
    // waited for it -- occupied every SM), so the TMA producer and the score-MMA issuers need not wait for the
    // predecessor grid: the pipeline fill overlaps the forward finalize kernel.  Only the column vectors (softmax
    // warps) and the accumulation buffer (flush warps) are behind griddepcontrol.wait.
    int early_operand;
    // Forward finalize: 1 -> write the per-row-block sums only and leave the reduction into `stats` / `loss_out` to the
    // backward finalize kernel of the same step (finish_stats), which takes the ticket / last-block tail off the path
    // between the two tile kernels.
    int defer_stats;
    int finish_stats;
    // Row-sharded batch, fused step (simclr_forward_backward_peer): the cross-GPU barrier in front of this launch is
    // executed INSIDE it -- the kernel in front has published this rank's epoch (sync_presignaled; else CTA 0 signals), the
    // TMA producer of every CTA waits before it touches what the peers pushed
    // (sync_flags.world > 0).  *sync_epoch is the value to signal / wait for; the preceding kernel of the stream bumped it
    // (bump_epoch of the prepare / forward finalize kernel).  No barrier kernel, no collective call.
    PeerTable sync_flags;
    const unsigned int* sync_epoch;
    unsigned int* bump_epoch;
    int sync_presignaled;      // the kernel in front of this one has already published the epoch to the peers: wait only
    unsigned long long peer_timeout_ns;   // watchdog of the cross-GPU waits (0 = none)
    // Exact accuracy count (bf16 mode, normalised rows): per-row candidate lists filled by the forward tile kernel and
    // consumed (and reset) by the forward finalize kernel.  nullptr: the count is decided on the tensor-core scores.
    unsigned int* cand_cnt;    // [2*bl_pad], zero on entry, left zero
    int* cand;                 // [2*bl_pad][kCandMax] global column indices
    float band;                // half-width of the uncertainty band: absolute in log2-domain logits (NT-Xent: k2 * kBandRel),
                               // relative (modified: kBandRel, all terms of its products are positive)
    // where the finalize kernel finds the exact fp32 normalised rows of the candidates' columns: the caller's inputs
    // (x1 / x2 / inv_norm above, one GPU), a gathered copy [2*bg_pad][D] (collective transport) or the ranks' own copies
    // [2*bl_pad][D] in symmetric memory (peer transport)
    const float* zrows;
    PeerTable zrows_peers;
    // Fused one-GPU step: the forward finalize kernel only LISTS the rows that need the exact re-scoring (amb_cnt /
    // amb_list, defer_accuracy); extra blocks of the backward finalize kernel work the list off (resolve_ambiguous, see
    // resolve_ambiguous_rows), marking the entries of the rows that lose their positive, and the last of them to finish
    // (amb_ticket) adds the unmarked ones to the count.  Nothing of the exact arithmetic then sits between the two tile
    // kernels or inside them.
    unsigned int* amb_cnt;     // [1] zero on entry (workspace header)
    unsigned int* amb_ticket;  // [1] zero on entry, left zero (workspace header)
    unsigned int* amb_done;    // [1] zero on entry: forward finalize blocks that have finished listing (workspace header)
    int* amb_list;             // [2*bl_pad] row slots (-1 - slot: re-scored and lost)
    int defer_accuracy;
    int resolve_ambiguous;
    // backward finalize, row-sharded fused step: add up the ranks' statistics [world][4] (fixed order) into stats / loss_out
    const float* stats_all;
    int stats_world;
    // Forward tile kernel, row-sharded fused step: ONE launch walks TWO column windows.  Window 0 = the columns this rank
    // produced itself, read from its own operand rows (tmap_rows) -- no peer needed -- while the spare warp of every CTA
    // pushes this rank's operand rows into all ranks' global operand matrices (push_src -> push_peers) and the warp that
    // finishes last publishes the epoch; window 1 = everybody else's columns, behind the wait for all ranks' flags.  The
    // NVLink transfer, its drain and the ranks' skew at the start of the step hide under the tiles of window 0.
    // n_windows == 1: win[0] repeats n_col_tiles / col_start / ... above (the backward kernel reads only those).
    int n_windows;
    int win0_local;
    ColWindow win[2];
    const void* push_src;          // this rank's operand rows [2*bl_pad][D] bf16
    PeerTable push_peers;          // every rank's global operand matrix [2*bg_pad][D] (world == 0: no push)
    unsigned int* push_ticket;     // counts the CTAs whose push is performed (zero on entry, left zero)
};

// Waits until *flag (system scope) has reached `target`.  Polls with an exponential __nanosleep back-off (the waiting
// thread shares its SM with working warps and its loads cross NVLink-coherent memory) and a %globaltimer watchdog:
// ranks of a training job legitimately skew by seconds to minutes (a checkpoint, a validation pass on rank 0, a
// dataloader stall), so the limit is minutes by default and configurable (SIMCLR_B200_PEER_TIMEOUT_S, 0 = wait
// for ever; see capi.cu peer_timeout_ns()) -- NCCL's comparable watchdog is host-side and also minutes.
SIMCLR_DEVICE void peer_flag_wait(const unsigned int* flag, unsigned int target, unsigned long long timeout_ns, int rank,
                                  int peer) {
    unsigned int seen;
    unsigned int backoff = 32, polls = 0;
    unsigned long long t0 = 0;
    do {
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flag) : "memory");
        if (static_cast<int>(seen - target) >= 0) return;
        __nanosleep(backoff);
        if (backoff < 256) {
            backoff <<= 1;             // short waits (the common case: ranks arrive within microseconds) stay responsive
        } else if (timeout_ns != 0 && (++polls & 1023u) == 0u) {
            const unsigned long long now = global_timer_ns();
            if (t0 == 0) t0 = now;
            if (now - t0 > timeout_ns) {
                printf("[simclr_b200] peer barrier timeout after %llu s: rank %d waits for rank %d (epoch %u, seen %u); "
                       "SIMCLR_B200_PEER_TIMEOUT_S sets the limit (0 = none)\n",
                       timeout_ns / 1000000000ull, rank, peer, target, seen);
                __trap();
            }
        }
    } while (true);
}

// Cross-GPU barrier executed by one thread of a kernel (see TileParams::sync_flags).  `signal`: this thread also
// publishes the epoch to every rank -- everything this GPU pushed before is complete, because the pushing kernels
// completed before the caller passed griddepcontrol.wait.  Returns when all ranks have published the epoch.
SIMCLR_DEVICE void peer_sync_thread(const PeerTable& flags, const unsigned int* epoch, bool signal,
                                    unsigned long long timeout_ns) {
    const unsigned int target = __ldcg(epoch);
    if (signal) {
        __threadfence_system();
        for (int r = 0; r < flags.world; ++r) {
            unsigned int* remote = static_cast<unsigned int*>(flags.ptr[r]) + flags.rank;
            asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(target) : "memory");
        }
    }
    const unsigned int* mine = static_cast<const unsigned int*>(flags.ptr[flags.rank]);
    for (int r = 0; r < flags.world; ++r) peer_flag_wait(mine + r, target, timeout_ns, flags.rank, r);
    // what the peers stored before their signal is read next by the TMA engine (async proxy)
    asm volatile("fence.proxy.async.global;" ::: "memory");
}

// Debug timeline: trace[(role * kTraceIters + it) * 4 + k].  Roles: 0 TMA producer, 1 MMA issuer,
// 2 + wg softmax warpgroup wg (lane 0 of its quarter-0 warp).
constexpr int kTraceIters = 64;
constexpr int kTraceRoles = 2 + 4;
// Per-role / per-CTA debug stamps of the tile kernels (tools/trace_timeline.py, tools/cta_timeline.py).  Compiled OUT of
// the product library: the predicated stamp code (constant loads, CTA-id compares) sits in the single-thread producer /
// issuer loops whose per-hop latency bounds the pipeline -- measured 67.0 -> 64.0 us per step without it
// (profiles/r01b_notes.md).  build.py builds lib/libsimclr_b200_trace.so with -DSIMCLR_TRACE=1 for the tools.
#ifndef SIMCLR_TRACE
#define SIMCLR_TRACE 0
#endif
SIMCLR_DEVICE void trace_event(const TileParams& p, int role, int it, int k) {
#if SIMCLR_TRACE
    if (p.trace != nullptr && static_cast<int>(blockIdx.x) == p.trace_cta && it < kTraceIters)
        p.trace[(role * kTraceIters + it) * 4 + k] = clock64();
#endif
}

// kPrec = 0: bf16 operands.  kPrec = 1 ("split", fp32-grade): every operand row is stored as hi = bf16(x) and
// lo = bf16(x - hi) in two planes; S = hi*hi + hi*lo + lo*hi on the tensor cores reproduces the fp32 product to ~2^-17,
// and W is split the same way for the gradient MMA.  Three times the tensor work: the correctness path of the
// fp32 contract (loss 1e-5, gradients 1e-4), not the fast path.  d <= 128 only (shared memory).
template <int D, int kPrec = 0>
struct SmemLayout {
    static constexpr int kPlanes = kPrec ? 2 : 1;
    static constexpr int kAtoms = D / kAtomK;
    static constexpr int kPlaneBytes = kAtoms * kAtomBytes;          // 128 x D bf16
    static constexpr int kTileBytes = kPlanes * kPlaneBytes;
    // even: see the issuer ownership rule
    static constexpr int kStages = kPrec ? ((D <= 64) ? 4 : 2) : ((D <= 64) ? 8 : (D <= 128 ? 4 : 2));
    static constexpr int kColvecBytes = 2 * kBlockN * 4;             // two planes of 128 floats
    static constexpr int kOffA = 0;
    static constexpr int kOffB = kTileBytes;
    static constexpr int kOffCv = kOffB + kStages * kTileBytes;
    static constexpr int kOffBar = kOffCv + kStages * kColvecBytes;
    // barriers: a_full, a_empty, acc_full, acc_empty, b_full[S], b_empty[S], s_full[8], s_free[8], w_full[8], w_done[8],
    // cv_full[S] (backward: the column vectors that travel with a column tile)
    static constexpr int kNumBars = 4 + 3 * kStages + 4 * kMaxSlots;   // kMaxSlots = 8
    static constexpr int kOffTmemPtr = kOffBar + kNumBars * 8;
    static constexpr int kOffFlags = kOffTmemPtr + 16;                 // 16 ints of CTA-wide scratch
    static constexpr int kOffMerge = kOffFlags + 64;                   // forward: [2][4 WG][5][128] floats
    static constexpr int kMergeBytes = 2 * kNumSoftmaxWG * kFwdFields * kBlockM * 4;
    // forward, bf16 mode: candidates of the exact accuracy count, double-buffered by segment parity:
    // [2] x { count u32 [128] | wrong u32 [128] | columns i32 [128][kCandMax] }
    static constexpr int kOffCand = kOffMerge + kMergeBytes;
    static constexpr int kCandWords = kBlockM * (2 + kCandMax);
    static constexpr int kCandBytes = kPrec ? 0 : 2 * kCandWords * 4;
    static constexpr int kBytes = kOffCand + kCandBytes;
    static constexpr int kDynamicBytes = kBytes + 1024;              // slack for manual 1024 B alignment
    static_assert(kDynamicBytes <= 227 * 1024, "shared memory of one CTA");
};


// First global column of tile j for a row block of view vr.
template <int kLoss>
SIMCLR_DEVICE int tile_col0(int col_start, int col_cnt, int tiles_per_view, int bg_pad, int vr, int j) {
    int v = 0;
    if constexpr (kLoss == kNtXent) {
        v = j >= col_cnt ? 1 : 0;            // NT-Xent rows see both views' windows, view 0 first
        j -= v * col_cnt;
    } else {
        v = 1 - vr;                           // modified rows see the other view only
    }
    int idx = col_start + j;
    if (idx >= tiles_per_view) idx -= tiles_per_view;
    return v * bg_pad + idx * kBlockN;
}
template <int kLoss>
SIMCLR_DEVICE int tile_col0(const TileParams& p, int vr, int j) {
    return tile_col0<kLoss>(p.col_start, p.col_cnt, p.tiles_per_view, p.bg_pad, vr, j);
}

// ---------------------------------------------------------------------------------------------
// Per-element maths.  "v" is the tracked raw value (monotone in the logit):
//   NT-Xent : v = S' = k2 * S    logit2 = v      the operands carry a factor sqrt(k2), k2 = log2(e)/tau, so the MMA
//                                                delivers log2-domain logits and exp2 needs no multiply
//   modified: v = max(B*S, 1e-4) logit2 = log2(v) * k2               (k2 = 1/tau)
// ---------------------------------------------------------------------------------------------
template <int kLoss>
SIMCLR_DEVICE float raw_value(const TileParams& p, float s) {
    if constexpr (kLoss == kNtXent) return s;
    else return fmaxf(s * p.qscale, kClampMin);
}
template <int kLoss>
SIMCLR_DEVICE float logit2(const TileParams& p, float v) {
    if constexpr (kLoss == kNtXent) return v;
    else return lg2_approx(v) * p.k2;
}

// Uncertainty band [lo, hi] of the tracked raw value around the exact positive `v_pos` (same expression in the tile
// kernel that records the candidates and in the finalize kernel that classifies the row).
// (explicitly rounded operations: no FMA contraction, so that both kernels compute the same bits)
template <int kLoss>
SIMCLR_DEVICE void cand_band(float band, float v_pos, float& lo, float& hi) {
    if constexpr (kLoss == kNtXent) {
        lo = __fsub_rn(v_pos, band);
        hi = __fadd_rn(v_pos, band);
    } else {
        lo = __fmul_rn(v_pos, __fsub_rn(1.0f, band));
        hi = __fmul_rn(v_pos, __fadd_rn(1.0f, band));
    }
}
// the exact positive in the units of the tracked raw value
template <int kLoss>
SIMCLR_DEVICE float pos_raw_value(float pos_dot, float k2, float qscale) {
    if constexpr (kLoss == kNtXent) return __fmul_rn(pos_dot, k2);
    else return fmaxf(__fmul_rn(pos_dot, qscale), kClampMin);
}

#ifndef SIMCLR_CAND_MODE
#define SIMCLR_CAND_MODE 1            // 0: candidates of the exact accuracy count are never recorded (A/B measurements)
#endif

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
    float s1 = 0.f, s2 = 0.f, s3 = 0.f;   // constant-shift path: three more independent partial sums (folded at segment end)
    f32x2 acc01 = 0ull, acc23 = 0ull;     // packed form of the same four partial sums (SIMCLR_PACKED)
    SIMCLR_DEVICE float total() const {
        float a, b, c, d;
        unpack2(acc01, a, b);
        unpack2(acc23, c, d);
        return ((sum + s1) + (s2 + s3)) + ((a + b) + (c + d));
    }
};

struct BwdRow {
    float row_a, row_l2;
};

SIMCLR_DEVICE float4 lds_f4(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr));
    return v;
}

// exp2 on the FMA pipe: Cody-Waite range reduction (round(x) lands in the low mantissa bits of x + 1.5*2^23)
// and a minimax polynomial for 2^f on [-0.5, 0.5].  Degree 3: max rel. error 7.5e-5 (mean 5e-6); degree 4: 2.7e-6
// (mean 4e-8).  The MUFU pipe retires 16 ex2 per clock and SM against 128 FMA lanes, so a quarter of the
// exponentials of a tile is moved here to take the tile off the MUFU bound.  Valid for finite -120 < x < 120.
template <int kDeg>
SIMCLR_DEVICE float ex2_poly(float x) {
    const float t = x + 12582912.0f;
    const float f = x - (t - 12582912.0f);
    float q;
    if constexpr (kDeg == 3) {
        q = fmaf(f, 0.0551716685295105f, 0.2426111251115799f);
        q = fmaf(f, q, 0.6932609677314758f);
        q = fmaf(f, q, 0.9999280571937561f);
    } else {
        q = fmaf(f, 0.009570101276040077f, 0.05591785907745361f);
        q = fmaf(f, q, 0.240247443318367f);
        q = fmaf(f, q, 0.6931217908859253f);
        q = fmaf(f, q, 0.9999992847442627f);
    }
    return __int_as_float(__float_as_int(q) + (__float_as_int(t) << 23));
}
// Two exponentials at once with packed fp32x2 instructions (degree 3): 6 packed + 2 integer instructions for the pair
// instead of 2 x 7.
SIMCLR_DEVICE void ex2_poly3_pair(float xa, float xb, float& ea, float& eb) {
    constexpr f32x2 kMagic = splat2_bits(0x4B400000u);       // 12582912.0f
    const f32x2 x = pack2(xa, xb);
    const f32x2 t = add2(x, kMagic);
    const f32x2 f = sub2(x, sub2(t, kMagic));
    f32x2 q = fma2(f, splat2_bits(0x3D61FBB0u), splat2_bits(0x3E786F0Du));   // 0.0551716685, 0.2426111251
    q = fma2(f, q, splat2_bits(0x3F31798Du));                                  // 0.6932609677
    q = fma2(f, q, splat2_bits(0x3F7FFB49u));                                  // 0.9999280572
    float qa, qb, ta, tb;
    unpack2(q, qa, qb);
    unpack2(t, ta, tb);
    ea = __int_as_float(__float_as_int(qa) + (__float_as_int(ta) << 23));
    eb = __int_as_float(__float_as_int(qb) + (__float_as_int(tb) << 23));
}
#ifndef SIMCLR_PACKED
#define SIMCLR_PACKED 1               // fp32x2 instructions in the softmax warps' inner loops (0: scalar forms)
#endif
// Packed path: how many PAIRS of the 16 exponentials of a chunk run on the FMA pipe (the rest on MUFU: 8 pipe cycles
// per warp instruction against 2 x 6 for a packed polynomial pair).
#ifndef SIMCLR_POLY_PAIRS_FWD
#define SIMCLR_POLY_PAIRS_FWD 3
#endif
#ifndef SIMCLR_POLY_PAIRS_BWD
#define SIMCLR_POLY_PAIRS_BWD 1
#endif
// pairs of the 8-column round `round` (0 / 1) of a chunk: the odd pair, if any, goes to the second round
SIMCLR_DEVICE constexpr int poly_pairs_in_round(int pairs, int round) { return round == 0 ? pairs / 2 : pairs - pairs / 2; }
// element i of a 32-column chunk goes to the FMA-pipe exponential
#ifndef SIMCLR_POLY_MASK
#define SIMCLR_POLY_MASK 3            // element i of a chunk is a polynomial lane when (i & mask) == mask (3: every fourth)
#endif
#ifndef SIMCLR_FLUSH_WAIT_READ
#define SIMCLR_FLUSH_WAIT_READ 1      // last flush of a CTA: wait for the TMA engine's reads only (0: for the adds to be performed)
#endif
#ifndef SIMCLR_FLUSH_ALTERNATE
#define SIMCLR_FLUSH_ALTERNATE 1
#endif
#ifndef SIMCLR_TOKEN_FENCE
#define SIMCLR_TOKEN_FENCE 1          // 1: a basic-block boundary pins the token hand-over in front of its chunk's arithmetic
#endif
#ifndef SIMCLR_PINGPONG
#define SIMCLR_PINGPONG 1             // softmax pairs take turns (named-barrier token) instead of running freely
#endif
SIMCLR_DEVICE constexpr bool poly_lane(int i) { return (i & SIMCLR_POLY_MASK) == SIMCLR_POLY_MASK; }

constexpr int kChunk = 16;     // columns per softmax step: two 16-register TMEM loads are kept in flight per thread

// Loop-invariant scalars of the per-element maths, hoisted out of the kernel-parameter constant bank once per CTA
// (a dependent chain of LDC / LDCU between two tiles costs the softmax warps hundreds of cycles).
struct Hot {
    float k2, m2, qscale;
    int bg_pad, b_glob;
    bool const_shift;
    int pow;           // TileParams::pow
    float qc, clampc;  // modified backward, power path: qscale * 2^-m2 and 1e-4 * 2^-m2 (pow 2);  2^-m2 in qc for pow 1
};

template <int kLoss>
SIMCLR_DEVICE float raw_value_h(const Hot& h, float s) {
    if constexpr (kLoss == kNtXent) return s;
    else return fmaxf(s * h.qscale, kClampMin);
}
template <int kLoss>
SIMCLR_DEVICE float logit2_h(const Hot& h, float v) {
    if constexpr (kLoss == kNtXent) return v;
    else return lg2_approx(v) * h.k2;
}

// ---- forward: one 16-column chunk, no masked element (warp-uniform fact) ----
// kConst (NT-Xent with bounded scores): constant log2-domain shift m2, no running maximum; `cm` collects the chunk
// maximum for the caller (the first-argmax bookkeeping is folded into the state once per tile).
template <int kLoss, bool kConst, bool kPoly = true>
SIMCLR_DEVICE void fwd_chunk_fast(const Hot& h, const uint32_t (&r)[kChunk], float& cm, FwdState& st) {
    float v[kChunk];
#pragma unroll
    for (int i = 0; i < kChunk; ++i) v[i] = raw_value_h<kLoss>(h, __uint_as_float(r[i]));
    float c2 = fmaxf(v[0], v[1]);
#pragma unroll
    for (int i = 2; i < kChunk; i += 2) c2 = fmaxf(c2, fmaxf(v[i], v[i + 1]));     // FMNMX3
    if constexpr (kConst && kLoss == kModified) {
        // power path (1/tau = 1 or 2): exp(A) = v or v * v, no transcendental; bounded sums, no running maximum
        cm = fmaxf(cm, c2);
        const bool square = h.pow == 2;
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            const f32x2 v01 = pack2(v[i + 0], v[i + 1]), v23 = pack2(v[i + 2], v[i + 3]);
            st.acc01 = add2(st.acc01, square ? mul2(v01, v01) : v01);
            st.acc23 = add2(st.acc23, square ? mul2(v23, v23) : v23);
        }
    } else if constexpr (kConst) {
        cm = fmaxf(cm, c2);
        if constexpr (SIMCLR_PACKED && kPoly && SIMCLR_POLY_MASK == 3) {
            // eight columns per round: MUFU exponentials, the last pair(s) on the FMA pipe, four packed additions
#pragma unroll
            for (int i = 0; i < kChunk; i += 8) {
                float e[8];
                const int np = poly_pairs_in_round(SIMCLR_POLY_PAIRS_FWD, i / 8);
#pragma unroll
                for (int u = 0; u < 8; ++u)
                    if (u < 8 - 2 * np) e[u] = ex2_approx(v[i + u]);
#pragma unroll
                for (int u = 0; u < 8; u += 2)      // adjacent columns: an aligned register pair
                    if (u >= 8 - 2 * np) ex2_poly3_pair(v[i + u], v[i + u + 1], e[u], e[u + 1]);
                st.acc01 = add2(st.acc01, pack2(e[0], e[1]));
                st.acc23 = add2(st.acc23, pack2(e[2], e[3]));
                st.acc01 = add2(st.acc01, pack2(e[4], e[5]));
                st.acc23 = add2(st.acc23, pack2(e[6], e[7]));
            }
        } else {
#pragma unroll
            for (int i = 0; i < kChunk; i += 4) {
                st.sum += (kPoly && poly_lane(i + 0)) ? ex2_poly<3>(v[i + 0]) : ex2_approx(v[i + 0]);
                st.s1 += (kPoly && poly_lane(i + 1)) ? ex2_poly<3>(v[i + 1]) : ex2_approx(v[i + 1]);
                st.s2 += (kPoly && poly_lane(i + 2)) ? ex2_poly<3>(v[i + 2]) : ex2_approx(v[i + 2]);
                st.s3 += (kPoly && poly_lane(i + 3)) ? ex2_poly<3>(v[i + 3]) : ex2_approx(v[i + 3]);
            }
        }
    } else {
        cm = fmaxf(cm, c2);
        // online max without a data-dependent branch: rescale the running sum by exp2(old_shift - new_shift)
        const float new_max = fmaxf(st.run_max, c2);
        const float shift = logit2_h<kLoss>(h, new_max);
        const float old_shift = (st.run_max == kNegBig) ? shift : logit2_h<kLoss>(h, st.run_max);
        st.sum *= ex2_approx(old_shift - shift);
        st.run_max = new_max;
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            a0 += ex2_approx(logit2_h<kLoss>(h, v[i + 0]) - shift);
            a1 += ex2_approx(logit2_h<kLoss>(h, v[i + 1]) - shift);
            a2 += ex2_approx(logit2_h<kLoss>(h, v[i + 2]) - shift);
            a3 += ex2_approx(logit2_h<kLoss>(h, v[i + 3]) - shift);
        }
        st.sum += (a0 + a1) + (a2 + a3);
    }
}

// ---- forward: a chunk that may hold the diagonal, the positive or padded columns for some of the warp's rows.
// Branch-free: the (at most two) masked elements of a row are found by comparing the unrolled column index with the
// row's two special positions; `split` separates the columns that precede the positive in the reference's order.
template <int kLoss, bool kConst>
SIMCLR_DEVICE void fwd_chunk_special(const Hot& h, const uint32_t (&r)[kChunk], int cq, int vc, const RowCtx& rc,
                                     FwdState& st, float& cm) {
    const int icq = cq - vc * h.bg_pad;                          // image index of the chunk's first column
    const int n_valid = rc.row_ok ? h.b_glob - icq : 0;          // elements i >= n_valid are padding (or the row is)
    const int i_diag = rc.diag_col - cq;                         // outside [0, kChunk) when not in this chunk
    const int i_pos = rc.pos_col - cq;
    int split;                                                   // elements i < split precede the positive
    if constexpr (kLoss == kNtXent) split = (rc.vr == 0) ? (vc == 1 ? rc.g - icq : 0) : (vc == 1 ? kChunk : rc.g - icq);
    else split = rc.g - icq;
    float v[kChunk];
    float cp = kNegBig, cf = kNegBig;
#pragma unroll
    for (int i = 0; i < kChunk; ++i) {
        const float x = raw_value_h<kLoss>(h, __uint_as_float(r[i]));
        st.pos_mma = (i == i_pos) ? x : st.pos_mma;
        const bool dead = (i == i_diag) | (i == i_pos) | (i >= n_valid);
        const float vi = dead ? kNegBig : x;
        v[i] = vi;
        cp = fmaxf(cp, (i < split) ? vi : kNegBig);
        cf = fmaxf(cf, (i < split) ? kNegBig : vi);
    }
    st.max_prec = fmaxf(st.max_prec, cp);
    st.max_foll = fmaxf(st.max_foll, cf);
    cm = fmaxf(cm, fmaxf(cp, cf));      // largest valid negative of the chunk (exact accuracy count)
    float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
    if constexpr (kConst && kLoss == kModified) {
        // power path: masked elements (kNegBig) contribute nothing
        const bool square = h.pow == 2;
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            float e[4];
#pragma unroll
            for (int u = 0; u < 4; ++u) {
                const float x = v[i + u];
                e[u] = (x == kNegBig) ? 0.f : (square ? x * x : x);
            }
            a0 += e[0];
            a1 += e[1];
            a2 += e[2];
            a3 += e[3];
        }
        st.sum += (a0 + a1) + (a2 + a3);
    } else if constexpr (kConst) {
        // exp2(-3e38) = 0: masked elements drop out by themselves
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            a0 += ex2_approx(v[i + 0]);
            a1 += ex2_approx(v[i + 1]);
            a2 += ex2_approx(v[i + 2]);
            a3 += ex2_approx(v[i + 3]);
        }
        st.sum += (a0 + a1) + (a2 + a3);
    } else {
        const float new_max = fmaxf(st.run_max, fmaxf(cp, cf));
        if (new_max != kNegBig) {                                // nothing valid seen so far: keep the empty state
            const float shift = logit2_h<kLoss>(h, new_max);
            const float old_shift = (st.run_max == kNegBig) ? shift : logit2_h<kLoss>(h, st.run_max);
            st.sum *= ex2_approx(old_shift - shift);
            st.run_max = new_max;
#pragma unroll
            for (int i = 0; i < kChunk; i += 4) {
                float e[4];
#pragma unroll
                for (int u = 0; u < 4; ++u) {
                    e[u] = ex2_approx(logit2_h<kLoss>(h, v[i + u]) - shift);
                    e[u] = (v[i + u] == kNegBig) ? 0.f : e[u];
                }
                a0 += e[0];
                a1 += e[1];
                a2 += e[2];
                a3 += e[3];
            }
            st.sum += (a0 + a1) + (a2 + a3);
        }
    }
}

// ---- backward: one 16-column chunk -> 8 packed bf16x2 words of W ----
// kSplit: additionally emits wlo = bf16(W - float(bf16(W))) (fp32-grade mode)
template <int kLoss, bool kConst, bool kSpecial, bool kSplit = false>
SIMCLR_DEVICE void bwd_chunk(const Hot& h, const uint32_t (&r)[kChunk], uint32_t cv_addr, int cq, const RowCtx& rc,
                             const BwdRow& br, uint32_t (&w)[kChunk / 2], uint32_t (&wlo)[kChunk / 2]) {
    const int i_diag = rc.diag_col - cq, i_pos = rc.pos_col - cq;
    if (kLoss == kModified && kConst && !kSplit && h.pow != 0) {
        // Power path of the modified loss (1/tau = 1 or 2): e^A / P = B (B P)^(1/tau - 1) is 1 or B P itself, so
        // W = t (a_r + a_c) with t = [B P >= 1e-4] * (B P or 1) * 2^-m2 -- no transcendental (warp-uniform branch).
        const f32x2 row_a2 = pack2(br.row_a, br.row_a);
        const f32x2 qc2 = pack2(h.qc, h.qc);
        const bool square = h.pow == 2;
        const float thr = square ? h.clampc : kClampMin;
        const f32x2 sc2 = square ? qc2 : pack2(h.qscale, h.qscale);
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            const float4 ac = lds_f4(cv_addr + i * 4);
            float q0, q1, q2, q3;
            unpack2(mul2(pack2(__uint_as_float(r[i + 0]), __uint_as_float(r[i + 1])), sc2), q0, q1);
            unpack2(mul2(pack2(__uint_as_float(r[i + 2]), __uint_as_float(r[i + 3])), sc2), q2, q3);
            const float t0 = (q0 >= thr) ? (square ? q0 : h.qc) : 0.f, t1 = (q1 >= thr) ? (square ? q1 : h.qc) : 0.f;
            const float t2 = (q2 >= thr) ? (square ? q2 : h.qc) : 0.f, t3 = (q3 >= thr) ? (square ? q3 : h.qc) : 0.f;
            float wv[4];
            unpack2(mul2(pack2(t0, t1), add2(pack2(ac.x, ac.y), row_a2)), wv[0], wv[1]);
            unpack2(mul2(pack2(t2, t3), add2(pack2(ac.z, ac.w), row_a2)), wv[2], wv[3]);
            if constexpr (kSpecial) {
#pragma unroll
                for (int u = 0; u < 4; ++u) wv[u] = ((i + u == i_diag) | (i + u == i_pos)) ? 0.f : wv[u];
            }
            w[(i >> 1) + 0] = pack_bf16x2(wv[0], wv[1]);
            w[(i >> 1) + 1] = pack_bf16x2(wv[2], wv[3]);
        }
        return;
    }
    if constexpr (SIMCLR_PACKED && kLoss == kNtXent && kConst && !kSplit && SIMCLR_POLY_MASK == 3) {
        // W = exp2(S') * (a_r + a_c) with packed additions / multiplications; eight columns per round
        const f32x2 row_a2 = pack2(br.row_a, br.row_a);
#pragma unroll
        for (int i = 0; i < kChunk; i += 8) {
            const float4 a0 = lds_f4(cv_addr + i * 4), a1 = lds_f4(cv_addr + i * 4 + 16);
            float e[8];
            const int np = poly_pairs_in_round(SIMCLR_POLY_PAIRS_BWD, i / 8);
#pragma unroll
            for (int u = 0; u < 8; ++u)
                if (u < 8 - 2 * np) e[u] = ex2_approx(__uint_as_float(r[i + u]));
#pragma unroll
            for (int u = 0; u < 8; u += 2)
                if (u >= 8 - 2 * np)
                    ex2_poly3_pair(__uint_as_float(r[i + u]), __uint_as_float(r[i + u + 1]), e[u], e[u + 1]);
            const f32x2 w01 = mul2(pack2(e[0], e[1]), add2(pack2(a0.x, a0.y), row_a2));
            const f32x2 w23 = mul2(pack2(e[2], e[3]), add2(pack2(a0.z, a0.w), row_a2));
            const f32x2 w45 = mul2(pack2(e[4], e[5]), add2(pack2(a1.x, a1.y), row_a2));
            const f32x2 w67 = mul2(pack2(e[6], e[7]), add2(pack2(a1.z, a1.w), row_a2));
            float wv[8];
            unpack2(w01, wv[0], wv[1]);
            unpack2(w23, wv[2], wv[3]);
            unpack2(w45, wv[4], wv[5]);
            unpack2(w67, wv[6], wv[7]);
            if constexpr (kSpecial) {
#pragma unroll
                for (int u = 0; u < 8; ++u) wv[u] = ((i + u == i_diag) | (i + u == i_pos)) ? 0.f : wv[u];
            }
#pragma unroll
            for (int u = 0; u < 8; u += 2) w[(i + u) >> 1] = pack_bf16x2(wv[u], wv[u + 1]);
        }
    } else {
#pragma unroll
    for (int i = 0; i < kChunk; i += 4) {
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
                    // one exp per element: W = exp2(S') * (a_r + a_c); every fourth one on the FMA pipe
                    const float e = (!kSplit && poly_lane(i + u)) ? ex2_poly<3>(sraw) : ex2_approx(sraw);
                    wval = e * (br.row_a + acs[u]);
                } else {
                    // general form: W = g_r exp2(S' - lse2_r) + g_c exp2(S' - lse2_c)
                    wval = br.row_a * ex2_approx(sraw - br.row_l2) + acs[u] * ex2_approx(sraw - lcs[u]);
                }
            } else {
                // d/dP of log(max(B P,1e-4))/tau = 1/(tau P) where live; folded: e^{A}/P = B q^{1/tau - 1}
                const float qv = sraw * h.qscale;
                const float y = lg2_approx(fmaxf(qv, kClampMin)) * (h.k2 - 1.0f);
                if constexpr (kConst) wval = ex2_approx(y - h.m2) * (br.row_a + acs[u]);
                else wval = br.row_a * ex2_approx(y - br.row_l2) + acs[u] * ex2_approx(y - lcs[u]);
                wval = (qv >= kClampMin) ? wval : 0.f;
            }
            if constexpr (kSpecial) wval = ((i + u == i_diag) | (i + u == i_pos)) ? 0.f : wval;
            wv[u] = wval;
        }
        w[(i >> 1) + 0] = pack_bf16x2(wv[0], wv[1]);
        w[(i >> 1) + 1] = pack_bf16x2(wv[2], wv[3]);
        if constexpr (kSplit) {
            const uint32_t p0 = w[(i >> 1) + 0], p1 = w[(i >> 1) + 1];
            wlo[(i >> 1) + 0] = pack_bf16x2(wv[0] - __uint_as_float(p0 << 16), wv[1] - __uint_as_float(p0 & 0xffff0000u));
            wlo[(i >> 1) + 1] = pack_bf16x2(wv[2] - __uint_as_float(p1 << 16), wv[3] - __uint_as_float(p1 & 0xffff0000u));
        }
    }
}
}

// ---------------------------------------------------------------------------------------------
// Fused finalize steps.  A row block's tiles are spread over a few CTAs (contiguous tile ranges); the CTA that
// completes its share last (atomic ticket per row block) finishes the block's 128 rows from L2-resident data.
// ---------------------------------------------------------------------------------------------
template <int kLoss>
SIMCLR_DEVICE float exact_logit2(const TileParams& p, float v) {
    if constexpr (kLoss == kNtXent) return v;
    else return log2f(v) * p.k2;
}

// Forward: merge the per-CTA partials of row block rb, add the exact positive term, emit lse2 / row_loss and the
// block's contribution to the loss statistics.  Called by the 128 threads of softmax warpgroup 0 (tid = row).
// Where the partials of row block rb live when ONE tile-kernel launch produced them (pure parameter arithmetic: the
// finalize kernel evaluates it before griddepcontrol.wait, i.e. while the tile kernel is still running).
struct FinIndex {
    const float* base;     // first contributing CTA's partial of this row (already offset by tid)
    size_t stride_k;       // distance between consecutive CTAs' partials
    int seg_first;         // segment index of the row block inside the first contributing CTA
    int nk;                // number of contributing CTAs (kFinFast + 1: use the generic two-pass path)
};
constexpr int kFinFast = 6;
SIMCLR_DEVICE FinIndex fin_index(const TileParams& p, int rb, int tid) {
    FinIndex ix;
    ix.base = nullptr;
    ix.stride_k = 0;
    ix.seg_first = 0;
    ix.nk = kFinFast + 1;
    if (p.n_fin_sets == 1) {
        // CTAs whose range [T*k/G, T*(k+1)/G) overlaps row block rb: owner(t) = ((t+1)*G - 1) / T.  The partial of CTA k
        // lives at part[(k * max_segs + seg_k)]: the first contributing CTA may have started in an earlier row block
        // (seg_k = rb - its first row block); every later one starts inside this row block (seg 0).
        const PartSet& ps = p.fin_set[0];
        const long long g = ps.grid;
        const long long t_lo = static_cast<long long>(rb) * ps.n_col_tiles, t_hi = t_lo + ps.n_col_tiles;
        const int k_first = static_cast<int>(((t_lo + 1) * g - 1) / ps.total_tiles);
        const int k_last = static_cast<int>((t_hi * g - 1) / ps.total_tiles);
        const long long c_first = (ps.total_tiles * k_first) / g;
        ix.seg_first = rb - static_cast<int>(c_first / ps.n_col_tiles);
        ix.stride_k = static_cast<size_t>(ps.max_segs) * (kFwdFields * kBlockM);
        ix.base = ps.part + static_cast<size_t>(k_first) * ix.stride_k + tid;
        ix.nk = k_last - k_first + 1;
    }
    return ix;
}

// ---- exact re-scoring of the candidates (forward finalize kernel, warp-cooperative) ----
// Where the exact fp32 normalised row of global column (view, image) comes from: a stash (gathered, or a peer's
// symmetric copy) or the caller's inputs, normalised on the fly exactly as the prepare kernel did.
struct ZRow {
    const float* stash;        // normalised fp32 row (d_pad elements) or nullptr
    const void* x;             // raw input row (d elements of in_dtype) when stash == nullptr
    float mul;                 // 1 / norm (L2 or L1), 1 when the loss was called with normalize = False
    const float* bn;           // BatchNorm scale | shift of the row's view (projection-head tail) or nullptr
};
template <int kLoss>
SIMCLR_DEVICE ZRow zrow_of(const TileParams& p, int view, int g_img) {
    ZRow z;
    z.x = nullptr;
    z.mul = 1.f;
    z.bn = nullptr;
    const int d_pad = p.d <= 64 ? 64 : (p.d <= 128 ? 128 : 256);
    if (p.zrows_peers.world > 0) {
        const int owner = g_img / p.b_loc, li = g_img - owner * p.b_loc;
        z.stash = static_cast<const float*>(p.zrows_peers.ptr[owner]) + static_cast<size_t>(view * p.bl_pad + li) * d_pad;
    } else if (p.zrows != nullptr) {
        z.stash = p.zrows + static_cast<size_t>(view * p.bg_pad + g_img) * d_pad;
    } else {
        z.stash = nullptr;
        const int li = g_img - p.row_off;        // one GPU: every column is a local row
        const size_t off = static_cast<size_t>(li) * p.d;
        const char* base = static_cast<const char*>(view == 0 ? p.x1 : p.x2);
        z.x = base + off * (p.in_bf16 ? 2 : 4);
        if (p.bn_state != nullptr) z.bn = p.bn_state + static_cast<size_t>(view) * 5 * d_pad;
        if (kLoss == kModified || p.normalize) {
            const float inv = __ldcg(p.inv_norm + view * p.bl_pad + li);
            z.mul = inv == kInvNormClamped ? 1.f / kNormEps : inv;
        }
    }
    return z;
}
// Lane l of the warp holds the elements [128 j + 4 l, 128 j + 4 l + 4) of a row, j < kVecs (kVecs = 1: d <= 128, 2: d <= 256),
// normalised exactly as the prepare kernel normalised them; one 16-byte (f32) / 8-byte (bf16) load per j when the row
// allows it.
template <int kLoss>
SIMCLR_DEVICE float4 zrow_vec(const TileParams& p, const ZRow& z, int lane, int j) {
    const int k0 = 128 * j + 4 * lane;
    float e[4] = {0.f, 0.f, 0.f, 0.f};
    if (k0 >= p.d) return make_float4(0.f, 0.f, 0.f, 0.f);
    if (z.stash != nullptr) return __ldcg(reinterpret_cast<const float4*>(z.stash + k0));   // d_pad floats, zero beyond d
    const bool full = k0 + 4 <= p.d;
    const uintptr_t addr = reinterpret_cast<uintptr_t>(z.x) + static_cast<uintptr_t>(k0) * (p.in_bf16 ? 2 : 4);
    if (full && !p.in_bf16 && (addr & 15u) == 0) {
        const float4 v = __ldg(reinterpret_cast<const float4*>(addr));
        e[0] = v.x; e[1] = v.y; e[2] = v.z; e[3] = v.w;
    } else if (full && p.in_bf16 && (addr & 7u) == 0) {
        const uint2 v = __ldg(reinterpret_cast<const uint2*>(addr));
        e[0] = __uint_as_float(v.x << 16); e[1] = __uint_as_float(v.x & 0xffff0000u);
        e[2] = __uint_as_float(v.y << 16); e[3] = __uint_as_float(v.y & 0xffff0000u);
    } else {
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (k0 + u < p.d) e[u] = load_elem(z.x, static_cast<size_t>(k0 + u), p.in_bf16);
    }
    if (z.bn != nullptr) {
        const int d_pad = p.d <= 64 ? 64 : (p.d <= 128 ? 128 : 256);
#pragma unroll
        for (int u = 0; u < 4; ++u)
            if (k0 + u < p.d) e[u] = fmaf(e[u], __ldg(z.bn + k0 + u), __ldg(z.bn + d_pad + k0 + u));
    }
#pragma unroll
    for (int u = 0; u < 4; ++u) {
        if constexpr (kLoss == kModified) e[u] = (k0 + u < p.d) ? softplus_beta(e[u]) : 0.f;
        e[u] *= z.mul;
    }
    return make_float4(e[0], e[1], e[2], e[3]);
}
// <z_r, z_c> by the whole warp from the lanes' vectors: fixed order of operations, then one xor butterfly -- the same
// two rows always give the same bits, whichever column is "the positive" (exact ties stay exact).
template <int kVecs>
SIMCLR_DEVICE float zrow_dot(const float4 (&a)[kVecs], const float4 (&b)[kVecs]) {
    float acc = 0.f;
#pragma unroll
    for (int j = 0; j < kVecs; ++j) {
        acc = fmaf(a[j].x, b[j].x, acc);
        acc = fmaf(a[j].y, b[j].y, acc);
        acc = fmaf(a[j].z, b[j].z, acc);
        acc = fmaf(a[j].w, b[j].w, acc);
    }
    return warp_sum(acc);
}
// the reference's fp32 logit of an exact similarity (objective.py:35-43: s / temperature; :87-90: log(clamp(B s)) / temperature)
template <int kLoss>
SIMCLR_DEVICE float reference_logit(const TileParams& p, float s) {
    if constexpr (kLoss == kNtXent) return __fdiv_rn(s, p.tau);
    else return __fdiv_rn(logf(fmaxf(s * p.qscale, kClampMin)), p.tau);
}
// Does the row (view vr, global image g) keep its positive as FIRST argmax among the columns of its recorded candidate
// chunks (list[i] = first global column of a kCandRange-column chunk)?  Whole warp.  All columns of a chunk are in
// flight at once: a chunk costs one memory latency.
// (u_begin, u_count: the columns [u_begin, u_begin + u_count) of every chunk only; whole chunks by default)
template <int kLoss, int kVecs, int kRound = kCandRange>
SIMCLR_DEVICE bool exact_first_argmax_v(const TileParams& p, int vr, int g, const int* list, int n, int lane, int u_begin = 0,
                                        int u_count = kCandRange) {
    static_assert(kCandRange % kRound == 0, "a chunk is re-scored in whole rounds");
    const ZRow row = zrow_of<kLoss>(p, vr, g), pos = zrow_of<kLoss>(p, 1 - vr, g);
    float4 zr[kVecs], zp[kVecs];
#pragma unroll
    for (int j = 0; j < kVecs; ++j) {
        zr[j] = zrow_vec<kLoss>(p, row, lane, j);
        zp[j] = zrow_vec<kLoss>(p, pos, lane, j);
    }
    const float l_pos = reference_logit<kLoss>(p, zrow_dot<kVecs>(zr, zp));
    bool ok = true;
    for (int i = 0; i < n && ok; ++i) {
        const int first = __ldcg(list + i);
        const int vc = first >= p.bg_pad ? 1 : 0;            // a chunk never mixes views
        for (int u0 = u_begin; u0 < u_begin + u_count && ok; u0 += kRound) {
            const int ic0 = first + u0 - vc * p.bg_pad;
            float4 zc[kRound][kVecs];
#pragma unroll
            for (int v = 0; v < kRound; ++v) {
                // masked columns (padding, the row itself, the positive: image g in either view) load row g instead
                const bool live = ic0 + v < p.b_glob && ic0 + v != g && !(kLoss == kModified && vc == vr);
                const ZRow col = zrow_of<kLoss>(p, vc, live ? ic0 + v : g);
#pragma unroll
                for (int j = 0; j < kVecs; ++j) zc[v][j] = zrow_vec<kLoss>(p, col, lane, j);
            }
#pragma unroll
            for (int v = 0; v < kRound; ++v) {
                const int ic = ic0 + v;
                const bool live = ic < p.b_glob && ic != g && !(kLoss == kModified && vc == vr);
                const float l_c = reference_logit<kLoss>(p, zrow_dot<kVecs>(zr, zc[v]));
                // reference column order (objective.py:48-49 / :93): does column (vc, ic) come before the positive?
                bool prec;
                if constexpr (kLoss == kNtXent) prec = (vr == 0) ? (vc == 1 && ic < g) : (vc == 1 || ic < g);
                else prec = ic < g;
                if (live) ok = ok && (prec ? (l_c < l_pos) : (l_c <= l_pos));
            }
        }
    }
    return ok;
}
template <int kLoss>
SIMCLR_DEVICE bool exact_first_argmax(const TileParams& p, int vr, int g, const int* list, int n, int lane) {
    return p.d <= 128 ? exact_first_argmax_v<kLoss, 1>(p, vr, g, list, n, lane)
                      : exact_first_argmax_v<kLoss, 2>(p, vr, g, list, n, lane);
}

// Fused one-GPU step: the rows whose accuracy decision the forward finalize kernel left open (amb_list, usually none, a
// few per step on unrelated inputs) are re-scored exactly by kResolveBlocks extra blocks of the BACKWARD FINALIZE kernel,
// one warp per candidate COLUMN (row, positive and column re-derived from the inputs: ~700 instructions behind a handful
// of dependent loads), `worker` of `n_workers` warps.  The list is walked in batches of 32 entries -- lane l holds entry
// base + l and its number of chunks -- and the batch's columns are dealt round-robin.  A column that beats the positive
// marks the row's list entry (slot -> -1 - slot: the same value whoever writes it); finish_forward_stats counts the
// unmarked entries.  (First built into an idle flush warp of the backward tile kernel: under that kernel's load one
// 16-column chunk of the modified loss took 46 us of the warp, which the CTA's first accumulator flush then waited
// for -- two listed rows cost the 2N = 8192 step +17 us with a warm L2 and +45 us with a cold one.)
constexpr int kResolveBlocks = 16;
template <int kLoss, int kVecs>
SIMCLR_DEVICE void resolve_ambiguous_rows(const TileParams& p, int worker, int n_workers, int lane) {
    const unsigned int n = __ldcg(p.amb_cnt);
    int next = worker;                         // this warp's next item, counted from the start of the current batch
    for (unsigned int base = 0; base < n; base += 32) {
        const unsigned int i = base + lane;
        // (an entry another warp has already marked still counts its chunks: every warp must see the same enumeration)
        const int entry = i < n ? __ldcg(p.amb_list + i) : 0;
        const int slot = entry >= 0 ? entry : -1 - entry;
        int n_c = 0;
        if (i < n) n_c = min(static_cast<int>(__ldcg(p.cand_cnt + slot)), kCandMax);
        int incl = n_c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
            const int t = __shfl_up_sync(0xffffffffu, incl, o);
            if (lane >= o) incl += t;
        }
        const int excl = incl - n_c;
        const int items = __shfl_sync(0xffffffffu, incl, 31) * kCandRange;
        int jj = next;
        for (; jj < items; jj += n_workers) {
            const int ci = jj / kCandRange, u = jj - ci * kCandRange;
            const unsigned int owner = __ballot_sync(0xffffffffu, excl <= ci && ci < excl + n_c);
            const int src = __ffs(owner) - 1;
            const int slot_s = __shfl_sync(0xffffffffu, slot, src);
            const int c = ci - __shfl_sync(0xffffffffu, excl, src);
            const int vr = slot_s >= p.bl_pad ? 1 : 0;
            const int g = p.row_off + slot_s - vr * p.bl_pad;
            const bool ok = exact_first_argmax_v<kLoss, kVecs, 1>(p, vr, g, p.cand + static_cast<size_t>(slot_s) * kCandMax + c, 1,
                                                                  lane, u, 1);
            if (!ok && lane == 0) p.amb_list[base + src] = -1 - slot_s;
        }
        next = jj - items;
    }
}

// kFused: the instantiation of the fused one-GPU step (defer_stats and, when candidates are recorded, defer_accuracy): it
// carries neither the exact re-scoring nor the ticket / last-block code -- this kernel sits between the two tile kernels.
template <int kLoss, bool kFused>
SIMCLR_DEVICE void forward_finalize_rowblock(const TileParams& p, int rb, int tid, const FinIndex& ix, float w_row,
                                             float* red /*smem, 16 floats*/, int* flags /*smem*/) {
    const int blocks_per_view = p.bl_pad / kBlockM;
    const int vr = rb / blocks_per_view;
    const int img = (rb - vr * blocks_per_view) * kBlockM + tid;
    const bool row_ok = img < p.b_loc;
    const int slot = rb * kBlockM + tid;
    const float pos_exact = __ldcg(p.pos_dot + slot);
    // (issued with the first round of loads: read at the end, it would cost this kernel a second memory round trip --
    // and the kernel sits between the two tile kernels)
    const unsigned int n_cand_early = (p.cand_cnt != nullptr && row_ok) ? __ldcg(p.cand_cnt + slot) : 0u;
    float v_pos = pos_exact;
    if constexpr (kLoss == kModified) v_pos = fmaxf(v_pos * p.qscale, kClampMin);
    else v_pos *= p.k2;               // exact fp32 positive logit in the log2 domain of the MMA scores
    // generic path (several launches, or many small CTAs per row block): visit every contributing partial
    auto visit = [&](auto&& f) {
        for (int s = 0; s < p.n_fin_sets; ++s) {
            const PartSet& ps = p.fin_set[s];
            const long long g = ps.grid;
            const long long t_lo = static_cast<long long>(rb) * ps.n_col_tiles, t_hi = t_lo + ps.n_col_tiles;
            const int kf = static_cast<int>(((t_lo + 1) * g - 1) / ps.total_tiles);
            const int kl = static_cast<int>((t_hi * g - 1) / ps.total_tiles);
            const long long c_first = (ps.total_tiles * kf) / g;
            const int seg_first = rb - static_cast<int>(c_first / ps.n_col_tiles);
            const size_t stride_k = static_cast<size_t>(ps.max_segs) * (kFwdFields * kBlockM);
            for (int k = kf; k <= kl; ++k)
                f(ps.part + static_cast<size_t>(k) * stride_k + (k == kf ? seg_first * (kFwdFields * kBlockM) : 0) + tid);
        }
    };
    float vmax = v_pos, max_prec = kNegBig, max_foll = kNegBig, pos_mma = kNegBig, total;
    constexpr int kFast = kFinFast;
    const int nk = ix.nk;
    const float* base = ix.base;
    const size_t stride_k = ix.stride_k;
    const int seg_first = ix.seg_first;
    if (nk <= kFast) {
        // common case: issue every load up front (one L2 round trip), then merge from registers
        float sv[kFast], mv[kFast];
#pragma unroll
        for (int i = 0; i < kFast; ++i) {
            sv[i] = 0.f;
            mv[i] = kNegBig;
            if (i < nk) {
                const float* src = base + i * stride_k + (i == 0 ? seg_first * (kFwdFields * kBlockM) : 0);
                sv[i] = __ldcg(src + 0 * kBlockM);
                mv[i] = __ldcg(src + 1 * kBlockM);
                max_prec = fmaxf(max_prec, __ldcg(src + 2 * kBlockM));
                max_foll = fmaxf(max_foll, __ldcg(src + 3 * kBlockM));
                pos_mma = fmaxf(pos_mma, __ldcg(src + 4 * kBlockM));
            }
        }
#pragma unroll
        for (int i = 0; i < kFast; ++i) vmax = fmaxf(vmax, mv[i]);
        const float top = exact_logit2<kLoss>(p, vmax);
        total = exp2f(exact_logit2<kLoss>(p, v_pos) - top);
#pragma unroll
        for (int i = 0; i < kFast; ++i)
            if (mv[i] > kNegBig) total += sv[i] * exp2f(exact_logit2<kLoss>(p, mv[i]) - top);
    } else {
        // many small CTAs per row block (tiny problems) or two launches: two passes
        visit([&](const float* src) {
            vmax = fmaxf(vmax, __ldcg(src + 1 * kBlockM));
            max_prec = fmaxf(max_prec, __ldcg(src + 2 * kBlockM));
            max_foll = fmaxf(max_foll, __ldcg(src + 3 * kBlockM));
            pos_mma = fmaxf(pos_mma, __ldcg(src + 4 * kBlockM));
        });
        const float top = exact_logit2<kLoss>(p, vmax);
        total = exp2f(exact_logit2<kLoss>(p, v_pos) - top);
        visit([&](const float* src) {
            const float mk = __ldcg(src + 1 * kBlockM);
            if (mk > kNegBig) total += __ldcg(src + 0 * kBlockM) * exp2f(exact_logit2<kLoss>(p, mk) - top);
        });
    }
    const float top = exact_logit2<kLoss>(p, vmax);
    float l2 = 0.f, loss_r = 0.f, w = 0.f, hit = 0.f;
    bool ambiguous = false;
    unsigned int n_cand = 0;
    if (row_ok) {
        l2 = top + log2f(total);
        loss_r = (l2 - exact_logit2<kLoss>(p, v_pos)) * kLn2;
        w = w_row;
        // reference objective.py:51 -- Tensor.max returns the first maximal index
        hit = (max_prec < pos_mma && max_foll <= pos_mma) ? 1.f : 0.f;
        if (p.cand_cnt != nullptr) {
            // Exact count: certain unless the best negative lies in the band around the exact positive; then the
            // recorded candidates decide (a row with more than kCandMax of them keeps the tensor-core decision)
            n_cand = n_cand_early;
            float lo, hi;
            cand_band<kLoss>(p.band, pos_raw_value<kLoss>(pos_exact, p.k2, p.qscale), lo, hi);
            const float best = fmaxf(max_prec, max_foll);
            if (best > hi) hit = 0.f;
            else if (best < lo) hit = 1.f;
            else ambiguous = n_cand <= static_cast<unsigned int>(kCandMax);
        }
    }
    if (kFused && p.cand_cnt != nullptr) {
        // fused step: list the undecided rows (rare) for the extra blocks of the backward finalize kernel; they count as
        // misses here and are added back by the backward finalize kernel
        if (ambiguous) {
            p.amb_list[atomicAdd(p.amb_cnt, 1u)] = slot;
            hit = 0.f;
        } else if (n_cand != 0u) {
            p.cand_cnt[slot] = 0u;
        }
        // ... and pull the input rows of their candidate ranges towards L2 (the inputs of a step are read once by the
        // prepare kernel: by now they may only be in HBM), so that the re-scoring warp pays L2 latencies
        unsigned int todo = __ballot_sync(0xffffffffu, ambiguous);
        const int lane = tid & 31;
        while (todo != 0u) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int n_s = static_cast<int>(__shfl_sync(0xffffffffu, n_cand, src));
            const int slot_s = __shfl_sync(0xffffffffu, slot, src);
            for (int i = 0; i < n_s; ++i) {
                const int first = __ldcg(p.cand + static_cast<size_t>(slot_s) * kCandMax + i);
                const int vc = first >= p.bg_pad ? 1 : 0;
                for (int u = lane; u < kCandRange; u += 32) {
                    const int ic = first + u - vc * p.bg_pad;
                    if (ic >= p.b_glob) continue;
                    const ZRow z = zrow_of<kLoss>(p, vc, ic);
                    const char* row_ptr = static_cast<const char*>(z.stash != nullptr ? static_cast<const void*>(z.stash) : z.x);
                    const int bytes = p.d * (z.stash != nullptr || !p.in_bf16 ? 4 : 2);
                    for (int o = 0; o < bytes; o += 128) asm volatile("prefetch.global.L2 [%0];" ::"l"(row_ptr + o));
                }
            }
        }
    } else if (!kFused && p.cand_cnt != nullptr) {
        if (n_cand != 0u) p.cand_cnt[slot] = 0u;              // left zero for the next call
        unsigned int todo = __ballot_sync(0xffffffffu, ambiguous);
        const int lane = tid & 31;
        while (todo != 0u) {
            const int src = __ffs(todo) - 1;
            todo &= todo - 1u;
            const int img_s = __shfl_sync(0xffffffffu, img, src);
            const int n_s = static_cast<int>(__shfl_sync(0xffffffffu, n_cand, src));
            const int slot_s = __shfl_sync(0xffffffffu, slot, src);
            const bool ok = exact_first_argmax<kLoss>(p, vr, p.row_off + img_s, p.cand + static_cast<size_t>(slot_s) * kCandMax,
                                                     n_s, lane);
            if (lane == src) hit = ok ? 1.f : 0.f;
        }
    }
    p.lse2[slot] = l2;
    p.row_loss[slot] = loss_r;
    if (p.prime_colvec != nullptr || p.colvec_peers.world > 0) {
        // the backward's column vectors: a_c = g_c 2^(m2 - lse2_c) (or g_c in the general form) and lse2_c, with the
        // unweighted g_c = 1/(2B).  Row-sharded batch: written straight into every rank's copy over NVLink.
        const int gslot = vr * p.bg_pad + p.row_off + img;
        const int plane = 2 * p.bg_pad;
        const float g = 0.5f / static_cast<float>(p.b_glob);
        const float a_c = row_ok ? (p.const_shift ? g * exp2f(p.m2 - l2) : g) : 0.f;
        if (p.colvec_peers.world > 0) {
            if (row_ok) {
                for (int r = 0; r < p.colvec_peers.world; ++r) {
                    float* cv = static_cast<float*>(p.colvec_peers.ptr[r]);
                    cv[gslot] = a_c;
                    cv[plane + gslot] = l2;
                }
            }
        } else {                          // single GPU: bl_pad == bg_pad, padding slots are zeroed here as well
            p.prime_colvec[gslot] = a_c;
            p.prime_colvec[plane + gslot] = l2;
        }
    }

    // block sums in a fixed order (deterministic)
    const float r0 = warp_sum(w * loss_r), r1 = warp_sum(w), r2 = warp_sum(hit);
    if ((tid & 31) == 0) {
        red[0 * 4 + (tid >> 5)] = r0;
        red[1 * 4 + (tid >> 5)] = r1;
        red[2 * 4 + (tid >> 5)] = r2;
    }
    named_bar_sync(2, kBlockM);
    if (tid == 0) {
        p.block_part[rb * 4 + 0] = (red[0] + red[1]) + (red[2] + red[3]);
        p.block_part[rb * 4 + 1] = (red[4] + red[5]) + (red[6] + red[7]);
        p.block_part[rb * 4 + 2] = (red[8] + red[9]) + (red[10] + red[11]);
        if (kFused || p.defer_stats) {
            flags[1] = 0;                 // the backward finalize kernel of this step reduces block_part (finish_stats)
        } else {
            if (p.colvec_peers.world > 0) __threadfence_system();     // this block's pushes to the peers are performed
            else __threadfence();
            const unsigned int prev = atomicAdd(p.ticket, 1u);
            flags[1] = (prev == static_cast<unsigned int>(p.n_row_blocks) - 1u) ? 1 : 0;
        }
    }
    if (kFused || p.defer_stats) return;
    named_bar_sync(2, kBlockM);
    if (flags[1] && tid < 32) {
        __threadfence();
        float s0 = 0.f, s1 = 0.f, s2 = 0.f;
        for (int i = tid; i < p.n_row_blocks; i += 32) {
            s0 += __ldcg(p.block_part + i * 4 + 0);
            s1 += __ldcg(p.block_part + i * 4 + 1);
            s2 += __ldcg(p.block_part + i * 4 + 2);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        if (tid == 0) {
            p.stats[0] = s0;
            p.stats[1] = s1;
            p.stats[2] = s2;
            // `stats` may be host-mapped memory that the caller polls on element 3 (pytorch_simclr_b200/_hoststats.py)
            __threadfence_system();
            p.stats[3] = s0 / s1;
            if (p.loss_out) *p.loss_out = s0 / s1;
            for (int r = 0; r < p.stats_peers.world; ++r) {
                float* dst = static_cast<float*>(p.stats_peers.ptr[r]) + 4 * p.stats_peers.rank;
                dst[0] = s0;
                dst[1] = s1;
                dst[2] = s2;
            }
            *p.ticket = 0u;                        // leave the workspace header clean for the next call
            if (p.bump_epoch != nullptr) {
                // the epoch the backward tile kernel's barrier waits for -- published to the peers right here (every
                // block of this kernel made its pushes visible before it took its ticket): the peers see it a launch
                // latency earlier than if the backward tile kernel signalled
                const unsigned int target = *p.bump_epoch + 1u;
                *p.bump_epoch = target;
                __threadfence_system();
                for (int r = 0; r < p.sync_flags.world; ++r) {
                    unsigned int* remote = static_cast<unsigned int*>(p.sync_flags.ptr[r]) + p.sync_flags.rank;
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(target) : "memory");
                }
            }
        }
    }
}

// Backward finalize of ONE row (a warp): the tile kernel left W * operand of the row in dacc.  Adds the exact
// positive-pair term and applies the backward of the row normalisation (NT-Xent: L2, objective.py:26-27; modified:
// softplus + L1, :70-78).  The two input rows (x_self, x_other: caller inputs, never written by this library) are
// loaded BEFORE griddepcontrol.wait, i.e. while the tile kernel drains; everything the step produced comes after it.
template <int D, int kLoss, bool kDet>
SIMCLR_DEVICE void backward_finalize_row(const TileParams& p, int rb, int r, int lane, float (&dz_out)[D / 32],
                                         float (&dzx_out)[D / 32]) {
    const int blocks_per_view = p.bl_pad / kBlockM;
    const int vr = rb / blocks_per_view;
    const int img = (rb - vr * blocks_per_view) * kBlockM + r;
    if (img >= p.b_loc) return;
    const void* x_self = vr == 0 ? p.x1 : p.x2;
    const void* x_other = vr == 0 ? p.x2 : p.x1;
    void* g_self = vr == 0 ? p.g1 : p.g2;
    constexpr int kPerLane = D / 32;
    const bool vec4 = kPerLane == 4 && p.d == D && !p.in_bf16 &&
                      ((reinterpret_cast<uintptr_t>(p.x1) | reinterpret_cast<uintptr_t>(p.x2) |
                        reinterpret_cast<uintptr_t>(p.g1) | reinterpret_cast<uintptr_t>(p.g2)) & 15u) == 0;
    // Lane l owns the kPerLane consecutive columns [l * kPerLane, (l + 1) * kPerLane).  `vec4`: fp32 rows of exactly
    // D = 128 columns at 16-byte aligned addresses move as one float4 per array and lane.
    float xs[kPerLane], xo[kPerLane], ac[kPerLane];
    if (vec4) {
        if constexpr (kPerLane == 4) {
            const float4 a4 = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x_self) + static_cast<size_t>(img) * D) + lane);
            const float4 b4 = __ldg(reinterpret_cast<const float4*>(static_cast<const float*>(x_other) + static_cast<size_t>(img) * D) + lane);
            xs[0] = a4.x; xs[1] = a4.y; xs[2] = a4.z; xs[3] = a4.w;
            xo[0] = b4.x; xo[1] = b4.y; xo[2] = b4.z; xo[3] = b4.w;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const int k = lane * kPerLane + u;
            const bool in = k < p.d;
            xs[u] = in ? load_elem(x_self, static_cast<size_t>(img) * p.d + k, p.in_bf16) : 0.f;
            xo[u] = in ? load_elem(x_other, static_cast<size_t>(img) * p.d + k, p.in_bf16) : 0.f;
        }
    }
    pdl_wait();
    float xhat[kPerLane];
    if (p.bn_state != nullptr) {
        // the inputs are pre-BatchNorm activations: z = u * scale + shift (per view), xhat = (u - mean) * rstd
        const float* st_s = p.bn_state + static_cast<size_t>(vr) * 5 * D;
        const float* st_o = p.bn_state + static_cast<size_t>(1 - vr) * 5 * D;
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const int k = lane * kPerLane + u;
            xhat[u] = (xs[u] - __ldg(st_s + 2 * D + k)) * __ldg(st_s + 3 * D + k);
            xs[u] = fmaf(xs[u], __ldg(st_s + k), __ldg(st_s + D + k));
            xo[u] = fmaf(xo[u], __ldg(st_o + k), __ldg(st_o + D + k));
        }
    }
    const int slot_self = rb * kBlockM + r;
    const int slot_other = (1 - vr) * p.bl_pad + img;
    const int c_self = vr * p.bg_pad + p.row_off + img;
    const int c_other = (1 - vr) * p.bg_pad + p.row_off + img;
    if constexpr (kDet) {
        // add the (CTA, segment) slots of this row block in CTA order: CTA k of the tile kernel owned the tiles
        // [T k / G, T (k + 1) / G); the first contributing CTA may have started in an earlier row block
        const long long grid = p.tile_grid;
        const long long t_lo = static_cast<long long>(rb) * p.n_col_tiles, t_hi = t_lo + p.n_col_tiles;
        const int k_first = static_cast<int>(((t_lo + 1) * grid - 1) / p.total_tiles);
        const int k_last = static_cast<int>((t_hi * grid - 1) / p.total_tiles);
        const int seg_first = rb - static_cast<int>(((p.total_tiles * k_first) / grid) / p.n_col_tiles);
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) ac[u] = 0.f;
        for (int k = k_first; k <= k_last; ++k) {
            const size_t prow = (static_cast<size_t>(k) * p.max_segs + (k == k_first ? seg_first : 0)) * kBlockM + r;
            const float* src = p.det_part + prow * D + lane * kPerLane;
            if constexpr (kPerLane >= 4) {
#pragma unroll
                for (int u = 0; u < kPerLane; u += 4) {
                    const float4 c4 = __ldcg(reinterpret_cast<const float4*>(src + u));
                    ac[u] += c4.x; ac[u + 1] += c4.y; ac[u + 2] += c4.z; ac[u + 3] += c4.w;
                }
            } else {
                const float2 c2 = __ldcg(reinterpret_cast<const float2*>(src));
                ac[0] += c2.x; ac[1] += c2.y;
            }
        }
    } else if (vec4) {
        if constexpr (kPerLane == 4) {
            const float4 c4 = __ldcg(reinterpret_cast<const float4*>(p.dacc + static_cast<size_t>(slot_self) * D) + lane);
            ac[0] = c4.x; ac[1] = c4.y; ac[2] = c4.z; ac[3] = c4.w;
        }
    } else {
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) ac[u] = __ldcg(p.dacc + static_cast<size_t>(slot_self) * D + lane * kPerLane + u);
    }
    const float go = p.grad_out ? __ldg(p.grad_out) : 1.f;
    const float inv_s = __ldcg(p.inv_norm + slot_self), inv_o = __ldcg(p.inv_norm + slot_other);
    const float sc_s = p.col_scale ? __ldcg(p.col_scale + c_self) : 0.5f / static_cast<float>(p.b_glob);
    const float sc_o = p.col_scale ? __ldcg(p.col_scale + c_other) : 0.5f / static_cast<float>(p.b_glob);
    const float l2_s = __ldcg(p.colvec + 2 * p.bg_pad + c_self), l2_o = __ldcg(p.colvec + 2 * p.bg_pad + c_other);
    const float pd = __ldcg(p.pos_dot + slot_self);
    float coef, outer;
    if constexpr (kLoss == kNtXent) {
        // (g_r P[r,pos] + g_pos P[pos,r] - g_r - g_pos) * zhat_pos in exact fp32
        const float y = pd * p.k2;
        coef = sc_s * (exp2f(y - l2_s) - 1.f) + sc_o * (exp2f(y - l2_o) - 1.f);
        outer = p.inv_tau * go;
    } else {
        const float qv = pd * p.qscale;
        const float lq = log2f(fmaxf(qv, kClampMin));
        const float y = lq * (p.k2 - 1.f);
        coef = (qv >= kClampMin) ? (sc_s * exp2f(y - l2_s) + sc_o * exp2f(y - l2_o) - (sc_s + sc_o) * exp2f(-lq)) : 0.f;
        outer = p.inv_tau * p.qscale * go;
    }
    const bool scaled = (kLoss == kModified) || p.normalize;
    const float mul_s = scaled ? (inv_s == kInvNormClamped ? 1.f / kNormEps : inv_s) : 1.f;
    const float mul_o = scaled ? (inv_o == kInvNormClamped ? 1.f / kNormEps : inv_o) : 1.f;
    float raw[kPerLane], hs[kPerLane], dv[kPerLane];
    float t = 0.f;
#pragma unroll
    for (int u = 0; u < kPerLane; ++u) {
        const int k = lane * kPerLane + u;
        const bool in = k < p.d;
        float es = xs[u], eo = xo[u];
        raw[u] = es;
        if constexpr (kLoss == kModified) {
            es = in ? softplus_beta(es) : 0.f;
            eo = in ? softplus_beta(eo) : 0.f;
        }
        hs[u] = es * mul_s;
        const float ho = eo * mul_o;
        dv[u] = (ac[u] * p.acc_scale + coef * ho) * outer;
        t = fmaf(dv[u], hs[u], t);
    }
    t = warp_sum(t);
    float out[kPerLane];
#pragma unroll
    for (int u = 0; u < kPerLane; ++u) {
        float o = dv[u];
        if constexpr (kLoss == kNtXent) {
            // d/dz of z / max(||z||, eps): projection unless the clamp was active
            if (p.normalize) o = (inv_s == kInvNormClamped) ? o / kNormEps : (o - hs[u] * t) * inv_s;
        } else {
            // L1 normalisation of a positive vector, then softplus'(x) = sigmoid(beta x)
            o = (inv_s == kInvNormClamped) ? o / kNormEps : (o - t) * inv_s;
            o *= softplus_beta_grad(raw[u]);
        }
        out[u] = o;
    }
    if (p.bn_state != nullptr) {
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const bool in = lane * kPerLane + u < p.d;
            dz_out[u] = in ? out[u] : 0.f;
            dzx_out[u] = in ? out[u] * xhat[u] : 0.f;
        }
    }
    if (vec4) {
        if constexpr (kPerLane == 4)
            reinterpret_cast<float4*>(static_cast<float*>(g_self) + static_cast<size_t>(img) * D)[lane] =
                make_float4(out[0], out[1], out[2], out[3]);
    } else {
#pragma unroll
        for (int u = 0; u < kPerLane; ++u) {
            const int k = lane * kPerLane + u;
            if (k < p.d) store_elem(g_self, static_cast<size_t>(img) * p.d + k, p.in_bf16, out[u]);
        }
    }
}

// Deferred loss statistics (TileParams::finish_stats): one warp adds up the per-row-block sums the forward finalize
// kernel left in block_part, in a fixed order (deterministic).
SIMCLR_DEVICE void finish_forward_stats(const TileParams& p, int lane) {
    float s0 = 0.f, s1 = 0.f, s2 = 0.f;
    if (p.stats_all != nullptr) {
        // row-sharded batch: the ranks' sums, pushed by their forward finalize kernels and ordered by the barrier inside
        // the backward tile kernel; lane 0 adds them in rank order (the same value on every rank)
        for (int r = 0; r < p.stats_world; ++r) {
            s0 += __ldcv(p.stats_all + 4 * r + 0);
            s1 += __ldcv(p.stats_all + 4 * r + 1);
            s2 += __ldcv(p.stats_all + 4 * r + 2);
        }
    } else {
        for (int i = lane; i < p.n_row_blocks; i += 32) {
            s0 += __ldcg(p.block_part + i * 4 + 0);
            s1 += __ldcg(p.block_part + i * 4 + 1);
            s2 += __ldcg(p.block_part + i * 4 + 2);
        }
        s0 = warp_sum(s0);
        s1 = warp_sum(s1);
        s2 = warp_sum(s2);
        // rows this kernel's extra blocks re-scored exactly (fused step, see resolve_ambiguous_rows): an unmarked list entry
        // is a confirmed hit; the rows' candidate counters are left zero for the next call
        if (p.amb_cnt != nullptr && p.resolve_ambiguous) {
            const unsigned int n_amb = __ldcg(p.amb_cnt);
            float confirmed = 0.f;
            for (unsigned int i = lane; i < n_amb; i += 32) {
                const int entry = __ldcg(p.amb_list + i);
                confirmed += entry >= 0 ? 1.f : 0.f;
                p.cand_cnt[entry >= 0 ? entry : -1 - entry] = 0u;
            }
            s2 += warp_sum(confirmed);
        }
    }
    if (lane == 0) {
        p.stats[0] = s0;
        p.stats[1] = s1;
        p.stats[2] = s2;
        __threadfence_system();           // host-mapped `stats`: element 3 is what the caller polls on
        p.stats[3] = s0 / s1;
        if (p.loss_out) *p.loss_out = s0 / s1;
    }
}

// Debug: per-CTA %globaltimer stamps, ktrace[64 + cta*8 + k]  (k: 0 start, 1/3 segment 0/1 tiles done,
// 2/4 segment 0/1 finalize done, 5 end).
SIMCLR_DEVICE void cta_stamp(const TileParams& p, int k) {
#if SIMCLR_TRACE
    if (p.ktrace != nullptr) p.ktrace[64 + blockIdx.x * 8 + k] = global_timer_ns();
#endif
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

// Position in a ring of kSize entries: entry index + parity of its current use (flips on every wrap).
template <int kSize>
struct RingPos {
    int idx = 0, par = 0;
    SIMCLR_DEVICE void advance() {
        if (++idx == kSize) {
            idx = 0;
            par ^= 1;
        }
    }
};

// ---------------------------------------------------------------------------------------------
// The tile kernel
// ---------------------------------------------------------------------------------------------
// Ring positions.  Every tile occupies one position of the B-stage ring; in the backward kernel every segment end
// additionally occupies TWO positions whose stages the producer hands (empty) to the softmax warps as staging space
// for the accumulator flush.  All roles walk the same position sequence, so stage index and parity never need a
// division and stage / slot / pair / issuer ownership stays aligned (position parity == tile parity).
// kDet (backward): deterministic mode, see TileParams::deterministic -- a template parameter, not a run-time switch: the
// single-thread issuer loops bound the pipeline by their per-hop latency, and even a predicated-off wait in them costs
// the default mode more than a microsecond per launch (profiles/r02_notes.md).
// kWindows (forward): two column windows in one launch, see TileParams::n_windows -- also its own instantiation.
template <int D, int kLoss, bool kBackward, bool kConst, int kPrec, bool kDet = false, bool kWindows = false>
__global__ void __launch_bounds__(kBackward ? kThreadsBackward : kThreadsForward, 1)
contrastive_tile_kernel(const __grid_constant__ CUtensorMap tmap_rows, const __grid_constant__ CUtensorMap tmap_cols,
                        const __grid_constant__ CUtensorMap tmap_dacc, const __grid_constant__ TileParams p) {
    using L = SmemLayout<D, kPrec>;
    constexpr int kPlanes = L::kPlanes;
    // split mode: the three operand-plane products that make up one fp32-grade product (lo*lo is below 2^-17)
    constexpr int kProducts = kPrec ? 3 : 1;
    constexpr int S = L::kStages;
    static_assert(S % 2 == 0, "issuer ownership needs an even ring");
    constexpr uint32_t kIdescScore = make_idesc_bf16(kBlockM, kBlockN, 0, 0);
    constexpr uint32_t kIdescGrad = make_idesc_bf16(kBlockM, D, 0, 1);
    // TMEM: NB score buffers of 128 columns, then (backward) the D-column gradient accumulator
    constexpr int NB = kBackward ? ((kTmemCols - D) / kBlockN > kMaxScoreBufs ? kMaxScoreBufs : (kTmemCols - D) / kBlockN)
                                 : kMaxScoreBufs;
    constexpr uint32_t kTmemAcc = NB * kBlockN;
    constexpr int kSlots = NB * kNumPairs;            // slot -> fixed (pair, buffer)
    constexpr int kBoxesPerStage = L::kTileBytes / kAtomBytes;   // 16 KB fp32 boxes (128 rows x 32 columns) per ring stage

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    // barriers and tiles by 32-bit shared-space address (computed once)
    const uint32_t smem_base = smem_u32(smem);
    const uint32_t sa_addr = smem_base + L::kOffA;
    const uint32_t sb_addr = smem_base + L::kOffB;
    const uint32_t bars = smem_base + L::kOffBar;
    const uint32_t a_full = bars + 0 * 8;
    const uint32_t a_empty = bars + 1 * 8;
    const uint32_t acc_full = bars + 2 * 8;
    const uint32_t acc_empty = bars + 3 * 8;
    const uint32_t b_full = bars + 4 * 8;
    const uint32_t b_empty = b_full + S * 8;
    const uint32_t s_full = b_empty + S * 8;               // [kSlots] score tile ready in TMEM
    const uint32_t s_free = s_full + kMaxSlots * 8;        // [kSlots] forward only: softmax finished reading the score tile
    const uint32_t w_full = s_free + kMaxSlots * 8;        // [kSlots] backward only: W written to TMEM
    const uint32_t w_done = w_full + kMaxSlots * 8;        // [kSlots] backward only: gradient MMAs finished reading W
    const uint32_t cv_full = w_done + kMaxSlots * 8;       // [S] backward only: column vectors of the stage's tile landed
    uint32_t* tmem_ptr_smem = reinterpret_cast<uint32_t*>(smem + L::kOffTmemPtr);
    float* smem_merge = reinterpret_cast<float*>(smem + L::kOffMerge);

    // warp index through a shuffle so that the compiler knows it is warp-uniform (role branches stay uniform and
    // tcgen05 / TMA operands can live in uniform registers instead of per-instruction ELECT/R2UR waterfall loops)
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;

    // contiguous tile range of this CTA
    const long long t_begin = (p.total_tiles * blockIdx.x) / gridDim.x;
    const long long t_end = (p.total_tiles * (blockIdx.x + 1)) / gridDim.x;
    const int nct = p.n_col_tiles;
    const int blocks_per_view = p.bl_pad / kBlockM;
    // column windows (forward only, TileParams::n_windows): every role walks window 0's tile range, then window 1's; the
    // CTA-wide tile counter (ring positions, hand-off slots, ping-pong parity) and the segment counter run on across them
    struct WinRange {
        long long t_begin, t_end;
        int nct;
    };
    constexpr bool kWin = kWindows;
    static_assert(!(kWindows && kDet), "the deterministic accumulator slots are laid out for one window");
    constexpr int n_win = kWin ? 2 : 1;
    auto win_range = [&](int wi) {
        if constexpr (!kWin) {
            return WinRange{t_begin, t_end, nct};
        } else {
            const long long tt = p.win[wi].total_tiles;
            return WinRange{(tt * blockIdx.x) / gridDim.x, (tt * (blockIdx.x + 1)) / gridDim.x, p.win[wi].n_col_tiles};
        }
    };
    auto win_col0 = [&](int wi, int vr, int j) {
        if constexpr (!kWin) return tile_col0<kLoss>(p, vr, j);
        else return tile_col0<kLoss>(p.win[wi].col_start, p.win[wi].col_cnt, p.tiles_per_view, p.bg_pad, vr, j);
    };

    pdl_launch_dependents();
    if (threadIdx.x == 0) {
        tma_prefetch_desc(&tmap_rows);
        tma_prefetch_desc(&tmap_cols);
        if constexpr (kBackward) tma_prefetch_desc(&tmap_dacc);
    }
    if (warp == kScoreWarp0 && lane == 0) {
        mbar_init(a_full, 1);
        mbar_init(a_empty, kNumIssuers);
        mbar_init(acc_full, kNumIssuers);
        mbar_init(acc_empty, 128);
        for (int i = 0; i < S; ++i) {
            mbar_init(b_full + 8 * i, 1);
            mbar_init(b_empty + 8 * i, 1);
            mbar_init(cv_full + 8 * i, 1);
        }
        for (int i = 0; i < kMaxSlots; ++i) {
            mbar_init(s_full + 8 * i, 1);
            mbar_init(s_free + 8 * i, 256);
            mbar_init(w_full + 8 * i, 256);
            mbar_init(w_done + 8 * i, 1);
        }
        mbar_fence_init();
    }
    if (warp == kAllocWarp) {
        tmem_alloc(tmem_ptr_smem, kTmemCols);
        tmem_relinquish();
    }
    if constexpr (!kBackward && kPrec == 0) {
        // shared-memory candidate lists of the exact accuracy count start empty (both segment parities)
        uint32_t* cz = reinterpret_cast<uint32_t*>(smem + L::kOffCand);
        for (int i = threadIdx.x; i < 2 * L::kCandWords; i += blockDim.x) cz[i] = 0u;
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    // Every role re-reads the TMEM base address from shared memory through an opaque load: the value is then local to the
    // role's branch instead of being live (and, in the 80-register backward build, spilled to local memory) across the
    // role split -- the issuers' per-tile path no longer starts with a local-memory load.
    auto load_tmem_base = [&]() {
        uint32_t v;
        asm volatile("ld.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_base + L::kOffTmemPtr) : "memory");
        return v;
    };
    // everything above (barrier init, TMEM allocation, descriptor prefetch) overlapped the predecessor's tail
    // (early_operand: the producer waits later, after its first operand loads; the score issuers touch no global memory)
    const bool runs_ahead = kBackward && p.early_operand != 0 &&
                            (warp == kProducerWarp || (warp >= kScoreWarp0 && warp < kScoreWarp0 + kNumIssuers));
    if (!runs_ahead) pdl_wait();
    ktrace_begin(p.ktrace, kBackward ? 3 : 1);
#if SIMCLR_TRACE
    if (threadIdx.x == 0) {
        cta_stamp(p, 0);
        if (p.ktrace != nullptr) p.ktrace[64 + blockIdx.x * 8 + 6] = clock64();
    }
#endif

    if (warp == kProducerWarp) {
        // ================================ TMA producer ================================
        if (elect_one()) {
            auto load_rows = [&](int rb_) {
                mbar_arrive_expect_tx(a_full, L::kTileBytes);
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl)
#pragma unroll
                    for (int ka = 0; ka < L::kAtoms; ++ka)
                        tma_load_2d(sa_addr + pl * L::kPlaneBytes + ka * kAtomBytes, &tmap_rows, a_full, ka * kAtomK,
                                    pl * 2 * p.bl_pad + rb_ * kBlockM);
            };
            auto load_cols = [&](int stage, int c0, bool local) {
                mbar_arrive_expect_tx(b_full + 8 * stage, L::kTileBytes);
                // local (window 0 of the row-sharded forward): the tile is this rank's own, read from its operand rows
                const CUtensorMap* map = &tmap_cols;
                int y = c0, plane_rows = 2 * p.bg_pad;
                if constexpr (kWin && !kBackward) {
                    if (local) {
                        map = &tmap_rows;
                        y = c0 >= p.bg_pad ? p.bl_pad + (c0 - p.bg_pad - p.row_off) : c0 - p.row_off;
                        plane_rows = 2 * p.bl_pad;
                    }
                }
#pragma unroll
                for (int pl = 0; pl < kPlanes; ++pl)
#pragma unroll
                    for (int ka = 0; ka < L::kAtoms; ++ka)
                        tma_load_2d(sb_addr + stage * L::kTileBytes + pl * L::kPlaneBytes + ka * kAtomBytes, map,
                                    b_full + 8 * stage, ka * kAtomK, pl * plane_rows + y);
            };
            // Early operand loads (backward, see TileParams::early_operand): the row-block tile and the column tiles of
            // the first positions of the first segment leave before griddepcontrol.wait; their column vectors follow it.
            int early = 0;
            if (kBackward && p.early_operand != 0) {
                const WinRange w0 = win_range(0);
                for (TileWalker w(w0.t_begin, w0.t_end, w0.nct); w.valid() && early < S; w.next()) {
                    if (early == 0) load_rows(w.rb);
                    load_cols(early, win_col0(0, w.rb / blocks_per_view, w.j), false);
                    ++early;
                    if (w.seg_last()) break;
                }
                pdl_wait();
            }
            // two windows: the wait for the peers sits in front of window 1 (window 0 needs nobody)
            bool synced = p.sync_flags.world == 0 || n_win == 2;
            RingPos<S> ring;
            bool wrapped = false;
            int seg = 0;
            int base = 0;
            for (int wi = 0; wi < n_win; ++wi) {
                const WinRange wr = win_range(wi);
                const bool local = kWin && !kBackward && wi == 0 && p.win0_local != 0;
                if (wi == 1 && wr.t_end > wr.t_begin)
                    peer_sync_thread(p.sync_flags, p.sync_epoch, false, p.peer_timeout_ns);
                for (TileWalker w(wr.t_begin, wr.t_end, wr.nct); w.valid(); w.next()) {
                    const int it = base + w.idx;
                    if (w.seg_first()) {
                        if (it >= early) {
                            if (seg > 0) mbar_wait(a_empty, (seg - 1) & 1, 100);
                            load_rows(w.rb);
                        }
                        ++seg;
                    }
                    trace_event(p, 0, it, 0);
                    if (wrapped) mbar_wait(b_empty + 8 * ring.idx, ring.par ^ 1, 101);   // previous use of the stage released
                    trace_event(p, 0, it, 1);
                    if (!synced) {
                        // In-kernel cross-GPU barrier.  Forward: the row-block tile (local rows) is already in flight, the
                        // column tiles are what the peers pushed.  Backward: the operands were complete before this kernel
                        // started (early loads above), the peers' column vectors are what the barrier protects.
                        peer_sync_thread(p.sync_flags, p.sync_epoch, blockIdx.x == 0 && !p.sync_presignaled, p.peer_timeout_ns);
                        synced = true;
                    }
                    const int c0 = win_col0(wi, w.rb / blocks_per_view, w.j);
                    if (it >= early) load_cols(ring.idx, c0, local);
                    if constexpr (kBackward) {
                        const uint32_t cv = smem_base + L::kOffCv + ring.idx * (2 * kBlockN * 4);
                        mbar_arrive_expect_tx(cv_full + 8 * ring.idx, L::kColvecBytes);
                        bulk_load_1d(cv, p.colvec + c0, kBlockN * 4, cv_full + 8 * ring.idx);
                        bulk_load_1d(cv + kBlockN * 4, p.colvec + 2 * p.bg_pad + c0, kBlockN * 4, cv_full + 8 * ring.idx);
                    }
                    if (ring.idx == S - 1) wrapped = true;
                    ring.advance();
                    if (kBackward && w.seg_last()) {
                        // hand two empty stages to the flush warpgroup: staging space of the accumulator flush
#pragma unroll 1
                        for (int k = 0; k < 2; ++k) {
                            if (wrapped) mbar_wait(b_empty + 8 * ring.idx, ring.par ^ 1, 102);
                            mbar_arrive(b_full + 8 * ring.idx);
                            mbar_arrive(cv_full + 8 * ring.idx);      // keeps the phase of cv_full in step with the ring position
                            if (ring.idx == S - 1) wrapped = true;
                            ring.advance();
                        }
                    }
                }
                base += static_cast<int>(wr.t_end - wr.t_begin);
            }
        }
    } else if (warp >= kScoreWarp0 && warp < kScoreWarp0 + kNumIssuers) {
        // ================================ UMMA issuers: score tiles ================================
        // S[buf] = A * B_stage^T (both operands K-major).  The whole warp walks the loop and waits on the
        // barriers; one elected lane issues the tcgen05 ops (operands stay in uniform registers).
        const int me = warp - kScoreWarp0;
        const uint32_t tmem_base = load_tmem_base();
        const uint32_t a_addr = sa_addr;
        const uint32_t b_addr0 = sb_addr;
        RingPos<S> ring;
        RingPos<kSlots> slot, freed;          // freed: slot of tile idx - NB (the buffer's previous tenant)
        int buf = 0;
        int seg_seen = 0;
        int base = 0;
        for (int wi = 0; wi < n_win; ++wi) {
        const WinRange wr = win_range(wi);
        for (TileWalker w(wr.t_begin, wr.t_end, wr.nct); w.valid(); w.next()) {
            const int idx = base + w.idx;
            if (w.seg_first()) {
                mbar_wait(a_full, seg_seen & 1, 200);      // every issuer observes every phase
                ++seg_seen;
            }
            if ((ring.idx & 1) == me) {
                if (lane == 0) trace_event(p, 1, idx, 0);
                mbar_wait(b_full + 8 * ring.idx, ring.par, 201);
                // the buffer's previous tenant (tile idx-NB) must be finished: read by the softmax (forward) or
                // consumed as W by the gradient MMAs (backward)
                if (idx >= NB) mbar_wait((kBackward ? w_done : s_free) + 8 * freed.idx, freed.par, 202);
                tc_fence_after_sync();
                if (lane == 0) trace_event(p, 1, idx, 1);
                if (elect_one()) {
                    const uint32_t b_addr = b_addr0 + ring.idx * L::kTileBytes;
#pragma unroll
                    for (int pr = 0; pr < kProducts; ++pr) {
                        // (A plane, B plane): hi*hi, hi*lo, lo*hi
                        const uint32_t a_pl = a_addr + (pr == 2 ? L::kPlaneBytes : 0);
                        const uint32_t b_pl = b_addr + (pr == 1 ? L::kPlaneBytes : 0);
#pragma unroll
                        for (int ka = 0; ka < L::kAtoms; ++ka) {
#pragma unroll
                            for (int kk = 0; kk < 4; ++kk) {
                                const uint32_t off = ka * kAtomBytes + kk * 32;
                                umma_ss(tmem_base + buf * kBlockN, make_smem_desc(a_pl + off, 0, 1024),
                                        make_smem_desc(b_pl + off, 0, 1024), kIdescScore, (pr | ka | kk) != 0);
                            }
                        }
                    }
                    if constexpr (!kBackward) umma_commit(b_empty + 8 * ring.idx);
                    umma_commit(s_full + 8 * slot.idx);
                    if constexpr (!kBackward) trace_event(p, 1, idx, 2);
                }
                __syncwarp();
            }
            if (w.seg_last()) {
                // the row-block tile may be overwritten once the score MMAs of BOTH issuers are complete
                if (elect_one()) umma_commit(a_empty);
                __syncwarp();
            }
            if (idx >= NB) freed.advance();
            ring.advance();
            slot.advance();
            if (++buf == NB) buf = 0;
            if (kBackward && w.seg_last()) {
                // The two flush positions.  With S >= 4 this issuer's next own stage is filled by a TMA the producer
                // issues after both hand-overs, so the skipped phases are complete when it next waits on these
                // stages; with S == 2 the very next wait is on such a stage and the phase has to be observed.
#pragma unroll 1
                for (int k = 0; k < 2; ++k) {
                    if (S == 2 && (ring.idx & 1) == me) mbar_wait(b_full + 8 * ring.idx, ring.par, 205);
                    ring.advance();
                }
            }
        }
        base += static_cast<int>(wr.t_end - wr.t_begin);
        }
    } else if (kBackward && warp >= kGradWarp0 && warp < kFlushWarp0 + 4) {
        // ================================ UMMA issuers: gradient MMAs / flush warpgroup ================================
        // acc += W[buf] (TMEM, 128 x 128 bf16) * B_stage (MN-major: K = column index).  The accumulator is zeroed by
        // the flush warpgroup (at start and after every flush), so every MMA accumulates and the two issuers need no
        // mutual ordering.
        const bool issuer = warp < kGradWarp0 + kNumIssuers;
        const bool flusher = warp >= kFlushWarp0;
        const uint32_t tmem_base = load_tmem_base();
        const int me = warp - kGradWarp0;
        const int quarter = warp & 3;
        const int row_in_block = quarter * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t b_addr0 = sb_addr;

        if (flusher) {
            // zero the gradient accumulator (TMEM is not cleared by the allocation); phase 0 of acc_empty
#pragma unroll 1
            for (int q = 0; q < D / 32; ++q) tmem_st32_fill(tmem_base + lane_addr + kTmemAcc + q * 32, 0u);
            tmem_st_wait();
            tc_fence_before_sync();
            mbar_arrive(acc_empty);
        }

        // The tile walk, specialised at compile time for what this warp is: a warp that only issues gets a loop without the
        // flush's register pressure (the shared loop kept its counters in local memory: LDL / STL in the issuer's per-tile
        // path, whose latency bounds the pipeline), a warp that only flushes gets one without the issue code.
        auto walk = [&](auto is_issuer, auto is_flusher) {
        constexpr bool issuer = decltype(is_issuer)::value;
        constexpr bool flusher = decltype(is_flusher)::value;
        RingPos<S> ring;
        RingPos<kSlots> slot, prev_slot;      // prev_slot: hand-off slot of tile idx - 1 (deterministic mode)
        int buf = 0;
        int seg = 0;
        int base = 0;
        for (int wi = 0; wi < n_win; ++wi) {
        const WinRange wr = win_range(wi);
        for (TileWalker w(wr.t_begin, wr.t_end, wr.nct); w.valid(); w.next()) {
            const int idx = base + w.idx;
            if constexpr (issuer) {
                // phase 0 of acc_empty = initial zeroing, phase s = flush (and re-zeroing) of segment s-1
                if (w.seg_first()) mbar_wait(acc_empty, seg & 1, 203);
                if ((ring.idx & 1) == me) {
                    if (lane == 0) trace_event(p, 1, idx, 2);
                    // Deterministic mode: the two issuers would otherwise feed the accumulator in whatever order their W
                    // tiles become ready, and the fp32 accumulation order -- hence the last bit of the gradients -- would
                    // depend on timing.  The gradient MMAs of tile idx are issued when those of tile idx - 1 (the other
                    // issuer's) have COMPLETED; the score MMAs of later tiles keep the tensor pipe busy meanwhile.  Tile
                    // idx - 1 used the previous hand-off slot; this warp is that barrier's second waiter and, like the
                    // first, observes every one of its phases (the slot ring is even: a slot always meets the same issuer).
                    if constexpr (kDet) { if (idx > 0) mbar_wait(w_done + 8 * prev_slot.idx, prev_slot.par, 206); }
                    mbar_wait(w_full + 8 * slot.idx, slot.par, 204);
                    tc_fence_after_sync();
                    if (lane == 0) trace_event(p, 1, idx, 3);
                    if (elect_one()) {
                        const uint32_t b_addr = b_addr0 + ring.idx * L::kTileBytes;
#pragma unroll
                        for (int pr = 0; pr < kProducts; ++pr) {
                            // (W plane, operand plane): hi*hi, lo*hi, hi*lo.  W_hi of column half h (K chunks 4h..4h+3)
                            // sits at columns [64h, 64h+32) of the buffer, W_lo (split mode) at [64h+32, 64h+64)
                            const uint32_t w_pl = tmem_base + buf * kBlockN + (pr == 1 ? 32 : 0);
                            const uint32_t b_pl = b_addr + (pr == 2 ? L::kPlaneBytes : 0);
#pragma unroll
                            for (int kc = 0; kc < kBlockN / 16; ++kc)
                                umma_ts(tmem_base + kTmemAcc, w_pl + (kc >> 2) * 64 + (kc & 3) * 8,
                                        make_smem_desc(b_pl + kc * 2048, kAtomBytes, 1024), kIdescGrad, 1);
                        }
                        umma_commit(b_empty + 8 * ring.idx);     // B tile (and its column vectors) may be overwritten
                        umma_commit(w_done + 8 * slot.idx);      // score buffer may be overwritten
                    }
                    __syncwarp();
                }
                if (w.seg_last()) {
                    if (elect_one()) umma_commit(acc_full);      // both issuers: count 2
                    __syncwarp();
                }
            }
            ring.advance();
            if constexpr (kDet) prev_slot = slot;
            slot.advance();
            if (++buf == NB) buf = 0;
            if (!w.seg_last()) continue;

            // ---- end of segment: ring positions ring.idx, ring.idx + 1 are the flush's staging stages ----
            if constexpr (flusher) {
                RingPos<S> st1 = ring;
                st1.advance();
                mbar_wait(acc_full, seg & 1, 302);            // every gradient MMA of the segment has completed
                mbar_wait(b_full + 8 * ring.idx, ring.par, 303);   // the two staging stages are ours
                mbar_wait(b_full + 8 * st1.idx, st1.par, 304);
                tc_fence_after_sync();
                // Two 16-column TMEM loads in flight: the load of piece h+1 overlaps the shared-memory stores of piece h.
                // (16, not 32 columns per load: 64 registers of staging pushed the loop's own counters into local memory.)
                uint32_t ra[16], rb2[16];
                tmem_ld16(tmem_base + lane_addr + kTmemAcc, ra);
#pragma unroll
                for (int h = 0; h < D / 16; ++h) {
                    uint32_t (&r)[16] = (h & 1) ? rb2 : ra;
                    uint32_t (&nxt)[16] = (h & 1) ? ra : rb2;
                    tmem_ld_wait16(r);
                    if (h + 1 < D / 16) tmem_ld16(tmem_base + lane_addr + kTmemAcc + (h + 1) * 16, nxt);
                    const int q = h >> 1;                     // 32-column box
                    if (h & 1) tmem_st32_fill(tmem_base + lane_addr + kTmemAcc + q * 32, 0u);   // zero for the next segment
                    const int stage = (q / kBoxesPerStage) == 0 ? ring.idx : st1.idx;
                    const uint32_t row_addr = sb_addr + stage * L::kTileBytes + (q % kBoxesPerStage) * kAtomBytes +
                                              row_in_block * 128;
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sts_v4(row_addr + ((((h & 1) * 4 + j) ^ (row_in_block & 7)) << 4), r[4 * j], r[4 * j + 1], r[4 * j + 2],
                               r[4 * j + 3]);
                }
                tmem_st_wait();
                tc_fence_before_sync();
                mbar_arrive(acc_empty);                       // the next segment's gradient MMAs may start
                fence_proxy_async_smem();                     // generic-proxy stores -> visible to the TMA engine
                named_bar_sync(kFlushBar, 128);
                if (warp == kFlushIssueWarp && elect_one()) {
                    // Two or three CTAs share a row block and the ones that END in it flush at the same time: odd CTAs
                    // walk the 32-column boxes backwards so that simultaneous reductions start on different lines.
#pragma unroll 1
                    for (int qq = 0; qq < D / 32; ++qq) {
                        const int q = (SIMCLR_FLUSH_ALTERNATE && (blockIdx.x & 1)) ? D / 32 - 1 - qq : qq;
                        const int stage = (q / kBoxesPerStage) == 0 ? ring.idx : st1.idx;
                        const uint32_t box = sb_addr + stage * L::kTileBytes + (q % kBoxesPerStage) * kAtomBytes;
                        if constexpr (kDet)      // tmap_dacc describes det_part: this (CTA, segment)'s private slot
                            tma_store_2d(&tmap_dacc, box, q * 32, (static_cast<int>(blockIdx.x) * p.max_segs + seg) * kBlockM);
                        else
                            tma_reduce_add_2d(&tmap_dacc, box, q * 32, w.rb * kBlockM);
                    }
                    bulk_commit_group();
                    // The staging space may be reused (or the CTA may exit) once the TMA engine has read it; the adds
                    // themselves are performed by the time the grid completes, which is what the finalize kernel's
                    // griddepcontrol.wait observes.
#if SIMCLR_FLUSH_WAIT_READ
                    bulk_wait_group_read0();
#else
                    if (wi == n_win - 1 && w.idx == w.n - 1) bulk_wait_group0();
                    else bulk_wait_group_read0();
#endif
                    mbar_arrive(b_empty + 8 * ring.idx);
                    mbar_arrive(b_empty + 8 * st1.idx);
                }
                __syncwarp();
            }
            ++seg;
            ring.advance();
            ring.advance();
        }
        base += static_cast<int>(wr.t_end - wr.t_begin);
        }
        };
        if (issuer && flusher) walk(std::true_type{}, std::true_type{});
        else if (issuer) walk(std::true_type{}, std::false_type{});
        else walk(std::false_type{}, std::true_type{});
    } else if (!kBackward && warp == kScoreWarp0 + kNumIssuers) {
        // ================================ forward: the spare warp ================================
        // Row-sharded fused step: pushes this CTA's share of the rank's operand rows into every rank's global operand
        // matrix (one multimem.st per 16 bytes through the NVSwitch, or one store per peer) while the other warps work on
        // window 0; the warp whose push is performed last publishes the epoch (TileParams::n_windows).
        if (kWin && !kBackward && p.push_peers.world > 0) {
            constexpr int kVecPerRow = D * 2 / 16;                  // 16-byte pieces of one bf16 operand row
            constexpr int kRowsPerIter = kVecPerRow >= 32 ? 1 : 32 / kVecPerRow;
            const int sub = kVecPerRow >= 32 ? 0 : lane / kVecPerRow;
            const uint4* src = static_cast<const uint4*>(p.push_src);
            for (int r = static_cast<int>(blockIdx.x) * kRowsPerIter + sub; r < 2 * p.bl_pad; r += gridDim.x * kRowsPerIter) {
                const int v = r >= p.bl_pad ? 1 : 0;
                const int img = r - v * p.bl_pad;
                if (img >= p.b_loc) continue;
                const size_t dst_row = static_cast<size_t>(v) * p.bg_pad + p.row_off + img;
                for (int q = kVecPerRow >= 32 ? lane : lane % kVecPerRow; q < kVecPerRow; q += 32) {
                    const uint4 val = __ldcg(src + static_cast<size_t>(r) * kVecPerRow + q);
                    if (p.push_peers.mc != nullptr) {
                        asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(
                                         static_cast<uint4*>(p.push_peers.mc) + dst_row * kVecPerRow + q),
                                     "f"(__uint_as_float(val.x)), "f"(__uint_as_float(val.y)), "f"(__uint_as_float(val.z)),
                                     "f"(__uint_as_float(val.w))
                                     : "memory");
                    } else {
                        for (int t = 0; t < p.push_peers.world; ++t)
                            static_cast<uint4*>(p.push_peers.ptr[t])[dst_row * kVecPerRow + q] = val;
                    }
                }
            }
            __threadfence_system();                      // this thread's peer stores are performed
            __syncwarp();
            if (lane == 0 && atomicAdd(p.push_ticket, 1u) == gridDim.x - 1) {
                *p.push_ticket = 0u;
                __threadfence();
                const unsigned int target = __ldcg(p.sync_epoch);
                for (int t = 0; t < p.sync_flags.world; ++t) {
                    unsigned int* remote = static_cast<unsigned int*>(p.sync_flags.ptr[t]) + p.sync_flags.rank;
                    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(remote), "r"(target) : "memory");
                }
            }
            __syncwarp();
        }
        if (p.prime_dacc != nullptr) {
            const float4 z = make_float4(0.f, 0.f, 0.f, 0.f);
            for (unsigned long long i = static_cast<unsigned long long>(blockIdx.x) * 32 + lane; i < p.prime_dacc_vec4;
                 i += static_cast<unsigned long long>(gridDim.x) * 32)
                p.prime_dacc[i] = z;
        }
    } else if (warp < kNumSoftmaxWarps) {
        // ================================ softmax warpgroups ================================
        const uint32_t tmem_base = load_tmem_base();
        const int wg = (warp - kSoftmaxWarp0) >> 2;          // 0 .. kNumSoftmaxWG-1
        const int pair = wg >> 1;                             // which tiles (it % 2)
        const int half = wg & 1;                              // which 64 columns of the tile
        const int quarter = warp & 3;                         // TMEM lane quarter this warp may touch
        const int row_in_block = quarter * 32 + lane;
        const uint32_t lane_addr = static_cast<uint32_t>(quarter * 32) << 16;
        const uint32_t cv_base = smem_base + L::kOffCv;
        int n = 0;                                            // tiles of this CTA, all windows
        for (int wi = 0; wi < n_win; ++wi) {
            const WinRange wr = win_range(wi);
            n += static_cast<int>(wr.t_end - wr.t_begin);
        }
        // loop invariants out of the constant bank, once
        Hot h;
        h.k2 = p.k2;
        h.m2 = p.m2;
        h.qscale = p.qscale;
        h.bg_pad = p.bg_pad;
        h.b_glob = p.b_glob;
        h.const_shift = kConst;
        h.pow = p.pow;
        {
            const float c0 = exp2f(-p.m2);
            h.qc = p.pow == 2 ? p.qscale * c0 : c0;
            h.clampc = kClampMin * c0;
        }
        const int b_loc = p.b_loc, row_off = p.row_off;
        constexpr int kTokenChunk = kBackward ? SIMCLR_TOKEN_CHUNK : SIMCLR_TOKEN_CHUNK_FWD;
        const bool token_fence = p.n_row_blocks != 0;      // always true, unknown to the compiler (see SIMCLR_TOKEN_FENCE)
        // candidates of the exact accuracy count are recorded by the bf16-mode forward kernel (the split mode's scores
        // are fp32-grade already)
        constexpr bool kRecord = !kBackward && kPrec == 0 && SIMCLR_CAND_MODE != 0;
        const bool tracing = SIMCLR_TRACE && p.trace != nullptr && static_cast<int>(blockIdx.x) == p.trace_cta && quarter == 0 && lane == 0;
        constexpr int kLogS = S == 2 ? 1 : (S == 4 ? 2 : 3);
        static_assert((1 << kLogS) == S, "ring size must be 2, 4 or 8");

        // Segments (runs of tiles sharing a row block) -> the tiles of this warpgroup's pair inside each segment.
        // CTA tile `it` sits at ring position it (+ 2 per finished segment in the backward kernel: the flush
        // positions), in hand-off slot it % kSlots and TMEM buffer it % NB; the counters below step by two tiles.
        int slot = pair, slot_par = 0;
        int buf = pair % NB;
        int seg = 0;
        int base = 0;                        // tiles of the windows already walked
        for (int wi = 0; wi < n_win; ++wi) {
        const WinRange wr = win_range(wi);
        const int n_w = static_cast<int>(wr.t_end - wr.t_begin);
        int idx0 = 0, seg_w = 0;
        int rb = static_cast<int>(wr.t_begin / wr.nct);
        int j0 = static_cast<int>(wr.t_begin - static_cast<long long>(rb) * wr.nct);
        while (idx0 < n_w) {
            const int seg_len = min(n_w - idx0, wr.nct - j0);
            // ---- segment setup ----
            RowCtx rc;
            rc.vr = rb / blocks_per_view;                                            // view of this row block
            const int img = (rb - rc.vr * blocks_per_view) * kBlockM + row_in_block; // local image index
            rc.row_ok = img < b_loc;
            rc.g = row_off + img;                                                    // global image index
            rc.diag_col = (kLoss == kNtXent && rc.row_ok) ? rc.vr * h.bg_pad + rc.g : -1;
            rc.pos_col = rc.row_ok ? (1 - rc.vr) * h.bg_pad + rc.g : -1;
            // warp-uniform description of this warp's 32 rows
            const int img_lo = (rb - rc.vr * blocks_per_view) * kBlockM + quarter * 32;
            const bool warp_rows_ok = img_lo + 31 < b_loc;
            const int g_lo = row_off + img_lo, g_hi = g_lo + 31;
            FwdState fs;
            // constant-shift forward: the "running maximum" is the fixed raw value whose logit is 0
            if (!kBackward && kConst) fs.run_max = (kLoss == kNtXent) ? kConstShiftRaw : 1.0f;
            BwdRow br;
            br.row_a = 0.f;
            br.row_l2 = 0.f;
            if constexpr (kBackward) {
                if (rc.row_ok) {
                    br.row_a = __ldcg(p.colvec + rc.vr * h.bg_pad + rc.g);
                    br.row_l2 = __ldcg(p.colvec + 2 * h.bg_pad + rc.vr * h.bg_pad + rc.g);
                }
            }
            const int pos_off = kBackward ? 2 * seg : 0;
            // exact accuracy count: the band around this row's exact positive inside which a tensor-core score cannot
            // decide (lo = +inf: not recording -- padding rows, rows already known to be wrong, feature switched off)
            float band_lo = 3.0e38f, band_hi = 3.0e38f;
            // this row's slots in the shared-memory candidate buffer of the segment's parity: count | wrong flag | list
            // (addresses are recomputed where they are needed -- all of it rare or once per segment -- instead of being
            // kept in registers across the tile loop: the forward kernel runs at 96 of its 102 registers)
            auto cand_base = [&]() { return smem_base + L::kOffCand + (seg & 1) * (L::kCandWords * 4); };
            auto cand_cnt_addr_f = [&]() { return cand_base() + row_in_block * 4; };
            auto cand_wrong_addr_f = [&]() { return cand_base() + (kBlockM + row_in_block) * 4; };
            auto cand_list_addr_f = [&]() { return cand_base() + (2 * kBlockM + row_in_block * kCandMax) * 4; };
            int pending_range = -1;          // first global column of this thread's pending candidate range, or -1
            auto publish_range = [&](int first_col) {
                uint32_t at;
                asm volatile("atom.shared.add.u32 %0, [%1], 1;" : "=r"(at) : "r"(cand_cnt_addr_f()) : "memory");
                if (at < static_cast<uint32_t>(kCandMax))
                    asm volatile("st.shared.b32 [%0], %1;" ::"r"(cand_list_addr_f() + at * 4), "r"(first_col) : "memory");
#if SIMCLR_TRACE
                if (p.ktrace != nullptr) atomicAdd(p.ktrace + 33, 1ull);                       // diagnostics: published ranges
#endif
            };
            // The row's exact positive is loaded here and first USED when the first tile of the segment has been processed
            // (band_pending): consumed at once, the load's latency (~1 us from L2 / HBM) would stall the warp at every
            // segment start.
            float pos_exact = 0.f;
            bool band_pending = false;           // warp-uniform
            if constexpr (kRecord) {
                if (p.cand_cnt != nullptr) {
                    band_pending = true;
                    if (rc.row_ok) pos_exact = __ldcg(p.pos_dot + rb * kBlockM + row_in_block);
                }
            }

#pragma unroll 1
            for (int s = ((base + idx0) ^ pair) & 1; s < seg_len; s += 2) {
                const int it = base + idx0 + s;
                const int pos = it + pos_off;
                const int stage = pos & (S - 1), stage_par = (pos >> kLogS) & 1;
                const int c0 = win_col0(wi, rc.vr, j0 + s);
                const int vc = c0 >= h.bg_pad ? 1 : 0;            // a tile never mixes views
                const int ic0 = c0 - vc * h.bg_pad;               // image index of the tile's first column
                // ---- warp-uniform classification (first-argmax rule: does the tile precede the positive in the
                // reference's column order?  NT-Xent rows see [view-2 block | view-1 block] (objective.py:48-49);
                // modified rows see the other view in natural order (objective.py:93)) ----
                bool tile_prec = false;
                if constexpr (!kBackward) {
                    const bool before = ic0 + kBlockN - 1 < g_lo;
                    if constexpr (kLoss == kNtXent) tile_prec = (rc.vr == 0) ? (vc == 1 && before) : (vc == 1 || before);
                    else tile_prec = before;
                }
                // can any chunk of this tile hold a masked element for one of this warp's rows?
                bool tile_special = !(ic0 > g_hi || ic0 + kBlockN - 1 < g_lo);
                if constexpr (!kBackward) tile_special = tile_special || !warp_rows_ok || (ic0 + kBlockN - 1 >= h.b_glob);

                // One barrier per slot = it % (NB * #pairs): a slot always maps to the same warpgroup pair and the
                // same TMEM buffer, so every barrier has a single waiter that observes all of its phases in order
                // (a parity wait is only meaningful when the waiter is at most one phase away from the barrier).
                if (tracing) trace_event(p, 2 + wg, it, 0);
                mbar_wait(s_full + 8 * slot, slot_par, 300);
                if (tracing) trace_event(p, 2 + wg, it, 1);
                if constexpr (kBackward) mbar_wait(cv_full + 8 * stage, stage_par, 301);  // column vectors landed
                tc_fence_after_sync();

                // This warpgroup's 64 columns in four 16-column chunks; the TMEM load of chunk k+1 is in flight while
                // chunk k is processed.  "special" = the chunk may contain the diagonal, the positive or padded
                // columns for one of this warp's 32 rows (warp-uniform).
                const uint32_t t0 = tmem_base + lane_addr + buf * kBlockN + half * 64;
                const int cbase = c0 + half * 64;
                const uint32_t cv_tile = cv_base + stage * (2 * kBlockN * 4) + half * 64 * 4;
                // Exact accuracy count.  `tile_max`: this thread's largest valid negative of the tile.  Hot path: one compare
                // and one vote per tile.  Behind the vote (rare): a row whose negative exceeds the band is wrong whatever
                // else happens and stops taking part (the four thread instances of a row -- two pairs x two column
                // halves -- share that verdict through a flag in shared memory: a hint, plain stores); a tile maximum
                // INSIDE the band of an undecided row makes its 16-column chunk(s) the thread's pending candidate.  Pending
                // candidates are only published at the end of the segment, and only for rows that are still undecided
                // then: with unrelated embeddings nearly every row is proven wrong one tile later and nothing is
                // published at all (an immediate TMEM rescan cost ~5000 cycles per event, cold code, with both softmax
                // pairs waiting for the warp that ran it: +3.4 us per forward launch on random inputs, profiles/r02_notes.md).
                auto record_candidates = [&](float tile_max, const float (&cmk)[4]) {
                    if (SIMCLR_CAND_MODE == 3) return;                 // A/B: set-up only, never checked
                    if (band_pending) {
                        band_pending = false;
                        if (rc.row_ok) cand_band<kLoss>(p.band, pos_raw_value<kLoss>(pos_exact, h.k2, h.qscale), band_lo, band_hi);
                    }
                    // Branch-free part of the hot path: a negative above the band proves the row wrong (predicated flag
                    // store, band_lo = +inf).  With unrelated embeddings that is what happens to nearly every row in its
                    // first tile -- behind a branch it sent every warp through cold code once per segment (~1 us per
                    // forward launch).  Rows that are not recording have band_hi = 3e38: never true for them.
                    if (tile_max > band_hi && band_lo < 3.0e38f) {
                        band_lo = 3.0e38f;
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(cand_wrong_addr_f()), "r"(1u) : "memory");
                    }
                    if (!__any_sync(0xffffffffu, tile_max >= band_lo)) return;
                    if (SIMCLR_CAND_MODE == 4) { band_lo = 3.0e38f; return; }     // A/B: the vote only
#if SIMCLR_TRACE
                    if (p.ktrace != nullptr && lane == 0) atomicAdd(p.ktrace + 32, 1ull);     // diagnostics: triggers
#endif
                    if (tile_max >= band_lo) {
                        uint32_t known_wrong;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(known_wrong) : "r"(cand_wrong_addr_f()) : "memory");
                        if (known_wrong != 0u) {
                            band_lo = 3.0e38f;
                        } else {
                            // every 16-column chunk whose maximum lies in the band (none exceeds it) is a candidate
#pragma unroll
                            for (int k = 0; k < 4; ++k) {
                                if (cmk[k] >= band_lo) {
                                    if (pending_range >= 0) publish_range(pending_range);   // a second near-tie chunk: very rare
                                    pending_range = cbase + k * kChunk;
                                }
                            }
                        }
                    }
                };
                float cmk[4] = {kNegBig, kNegBig, kNegBig, kNegBig};   // forward: largest valid negative of each 16-column chunk
                uint32_t ra[kChunk], rb2[kChunk];
                tmem_ld16(t0, ra);
                // Ping-pong: the arithmetic of tile `it` starts when the other pair has finished that of tile it-1.
                // Two warps per sub-partition already saturate the MUFU / FMA pipes on this code, so running the two
                // pairs' maths back to back costs nothing, while the per-tile bookkeeping of one pair (barrier waits,
                // address arithmetic, fences: ~800 cycles) now hides behind the other pair's maths instead of both
                // pairs idling together (they otherwise drift into lock step).
                if (SIMCLR_PINGPONG && it > 0) named_bar_sync(kTokenBar0 + pair, 32 * kNumSoftmaxWarps);
                if (tracing) trace_event(p, 2 + wg, it, 3);
                if constexpr (kBackward && kPrec != 0) {
                    // Split mode: W_hi goes where it always goes, W_lo into the upper half [64h+32, 64h+64) of this
                    // warpgroup's region -- over scores of chunks 2 and 3, so the W_lo words are held in registers until
                    // the last chunk has been loaded.
                    uint32_t lo[3][kChunk / 2];
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t (&cur)[kChunk] = (k & 1) ? rb2 : ra;
                        uint32_t (&nxt)[kChunk] = (k & 1) ? ra : rb2;
                        tmem_ld_wait16(cur);
                        if (k < 3) tmem_ld16(t0 + (k + 1) * kChunk, nxt);
                        const int cq = cbase + k * kChunk;
                        const int icq = cq - vc * h.bg_pad;
                        const bool special = tile_special && !(icq > g_hi || icq + kChunk - 1 < g_lo);
                        uint32_t wq[kChunk / 2], wl[kChunk / 2];
                        if (special) bwd_chunk<kLoss, kConst, true, true>(h, cur, cv_tile + k * kChunk * 4, cq, rc, br, wq, wl);
                        else bwd_chunk<kLoss, kConst, false, true>(h, cur, cv_tile + k * kChunk * 4, cq, rc, br, wq, wl);
                        tmem_st8(t0 + k * (kChunk / 2), wq);
                        if (k < 3) {
#pragma unroll
                            for (int i = 0; i < kChunk / 2; ++i) lo[k][i] = wl[i];
                        } else {
                            tmem_st8(t0 + 32 + 0 * (kChunk / 2), lo[0]);
                            tmem_st8(t0 + 32 + 1 * (kChunk / 2), lo[1]);
                            tmem_st8(t0 + 32 + 2 * (kChunk / 2), lo[2]);
                            tmem_st8(t0 + 32 + 3 * (kChunk / 2), wl);
                        }
                    }
                    if (SIMCLR_PINGPONG && it + 1 < n) named_bar_arrive(kTokenBar0 + (pair ^ 1), 32 * kNumSoftmaxWarps);
                } else if (!tile_special) {
                    // Common case: no masked element anywhere in the tile for this warp.  One straight-line block over
                    // the four chunks, so that the tail of chunk k overlaps the head of chunk k+1.
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t (&cur)[kChunk] = (k & 1) ? rb2 : ra;
                        uint32_t (&nxt)[kChunk] = (k & 1) ? ra : rb2;
                        tmem_ld_wait16(cur);
                        if (k < 3) tmem_ld16(t0 + (k + 1) * kChunk, nxt);
                        // hand the token over one chunk early: the other pair's first chunk fills the pipes while this
                        // pair drains its last one (covers the bar.arrive -> bar.sync wake-up latency)
                        if (SIMCLR_PINGPONG && k == kTokenChunk && it + 1 < n) named_bar_arrive_pinned(kTokenBar0 + (pair ^ 1), 32 * kNumSoftmaxWarps, cur);
                        // ptxas floats a bar.arrive (nothing depends on it) to the END of its basic block, i.e. behind the
                        // arithmetic of the chunk it was written in front of -- the other pair then gets the token a chunk
                        // late.  An always-true branch the compiler cannot see through ends the basic block right here.
                        if (SIMCLR_TOKEN_FENCE && k == kTokenChunk && !token_fence) continue;
                        if constexpr (!kBackward) {
                            fwd_chunk_fast<kLoss, kConst, kPrec == 0>(h, cur, cmk[k], fs);
                        } else if constexpr (kPrec == 0) {
                            uint32_t wq[kChunk / 2], unused[kChunk / 2];
                            bwd_chunk<kLoss, kConst, false>(h, cur, cv_tile + k * kChunk * 4, 0, rc, br, wq, unused);
                            tmem_st8(t0 + k * (kChunk / 2), wq);
                        }
                    }
                    if constexpr (!kBackward) {
                        const float cm = fmaxf(fmaxf(cmk[0], cmk[1]), fmaxf(cmk[2], cmk[3]));
                        fs.max_prec = fmaxf(fs.max_prec, tile_prec ? cm : kNegBig);
                        fs.max_foll = fmaxf(fs.max_foll, tile_prec ? kNegBig : cm);
                    }
                } else {
                    auto process = [&](const uint32_t (&r)[kChunk], int k) {
                        const int cq = cbase + k * kChunk;
                        const int icq = cq - vc * h.bg_pad;
                        bool special = !(icq > g_hi || icq + kChunk - 1 < g_lo);
                        if constexpr (!kBackward) special = special || !warp_rows_ok || (icq + kChunk - 1 >= h.b_glob);
                        if constexpr (!kBackward) {
                            if (special) {
                                fwd_chunk_special<kLoss, kConst>(h, r, cq, vc, rc, fs, cmk[k]);
                            } else {
                                // an unmasked chunk of a tile that overlaps the warp's own images lies entirely before or
                                // entirely after them: classify it by itself (the tile as a whole straddles the positive)
                                const bool before = icq + kChunk - 1 < g_lo;
                                bool prec;
                                if constexpr (kLoss == kNtXent) prec = (rc.vr == 0) ? (vc == 1 && before) : (vc == 1 || before);
                                else prec = before;
                                float cm = kNegBig;
                                fwd_chunk_fast<kLoss, kConst, kPrec == 0>(h, r, cm, fs);
                                fs.max_prec = fmaxf(fs.max_prec, prec ? cm : kNegBig);
                                fs.max_foll = fmaxf(fs.max_foll, prec ? kNegBig : cm);
                                cmk[k] = cm;
                            }
                        } else {
                            uint32_t wq[kChunk / 2], unused[kChunk / 2];
                            const uint32_t cv_addr = cv_tile + k * kChunk * 4;
                            if (special) bwd_chunk<kLoss, kConst, true>(h, r, cv_addr, cq, rc, br, wq, unused);
                            else bwd_chunk<kLoss, kConst, false>(h, r, cv_addr, cq, rc, br, wq, unused);
                            // bf16 W of chunk k goes to columns [64*half + 8*k, +8): inside this warpgroup's own
                            // 64-column region and over scores this thread has already consumed
                            tmem_st8(t0 + k * (kChunk / 2), wq);
                        }
                    };
                    // same shape as the common case (loads one chunk ahead, token passed at the same chunk)
#pragma unroll
                    for (int k = 0; k < 4; ++k) {
                        uint32_t (&cur)[kChunk] = (k & 1) ? rb2 : ra;
                        uint32_t (&nxt)[kChunk] = (k & 1) ? ra : rb2;
                        tmem_ld_wait16(cur);
                        if (k < 3) tmem_ld16(t0 + (k + 1) * kChunk, nxt);
                        if (SIMCLR_PINGPONG && k == kTokenChunk && it + 1 < n) named_bar_arrive_pinned(kTokenBar0 + (pair ^ 1), 32 * kNumSoftmaxWarps, cur);
                        process(cur, k);
                    }
                }
                if (SIMCLR_PINGPONG && !(kBackward && kPrec != 0) && kTokenChunk > 3 && it + 1 < n) named_bar_arrive(kTokenBar0 + (pair ^ 1), 32 * kNumSoftmaxWarps);
                if constexpr (!kBackward) {
                    tc_fence_before_sync();
                    mbar_arrive(s_free + 8 * slot);
                    // (behind the release of the score buffer: the check needs nothing from TMEM)
                    if constexpr (kRecord) record_candidates(fmaxf(fmaxf(cmk[0], cmk[1]), fmaxf(cmk[2], cmk[3])), cmk);
                } else {
                    tmem_st_wait();
                    tc_fence_before_sync();
                    mbar_arrive(w_full + 8 * slot);
                }
                if (tracing) trace_event(p, 2 + wg, it, 2);
                slot += 2;
                if (slot >= kSlots) {
                    slot -= kSlots;
                    slot_par ^= 1;
                }
                buf += 2;
                if (buf >= NB) buf -= NB;
            }

            // ---- end of segment ----
            if (threadIdx.x == 0 && seg < 2) cta_stamp(p, 1 + 2 * seg);
            if constexpr (kRecord) {
                // a pending candidate range of a row that is still undecided now is a real near-tie: publish it
                if (pending_range >= 0 && band_lo < 3.0e38f) {
                    uint32_t known_wrong;
                    asm volatile("ld.shared.b32 %0, [%1];" : "=r"(known_wrong) : "r"(cand_wrong_addr_f()) : "memory");
                    if (known_wrong == 0u) publish_range(pending_range);
                }
            }
            if constexpr (!kBackward) {
                // merge the four warpgroups' partial (max, sum, argmax bookkeeping) through shared memory ...
                float* mg = smem_merge + (seg & 1) * (kNumSoftmaxWG * kFwdFields * kBlockM);
                float* mine = mg + wg * (kFwdFields * kBlockM) + row_in_block;
                mine[0 * kBlockM] = fs.total();
                mine[1 * kBlockM] = fs.run_max;
                mine[2 * kBlockM] = fs.max_prec;
                mine[3 * kBlockM] = fs.max_foll;
                mine[4 * kBlockM] = fs.pos_mma;
                named_bar_sync(1, 32 * kNumSoftmaxWarps);
                if (wg == 0) {
                    float m = kNegBig, mp = kNegBig, mf = kNegBig, pm = kNegBig;
#pragma unroll
                    for (int g2 = 0; g2 < kNumSoftmaxWG; ++g2) {
                        const float* o = mg + g2 * (kFwdFields * kBlockM) + row_in_block;
                        m = fmaxf(m, o[1 * kBlockM]);
                        mp = fmaxf(mp, o[2 * kBlockM]);
                        mf = fmaxf(mf, o[3 * kBlockM]);
                        pm = fmaxf(pm, o[4 * kBlockM]);
                    }
                    float total = 0.f;
                    if (m > kNegBig) {
                        const float top = logit2<kLoss>(p, m);
#pragma unroll
                        for (int g2 = 0; g2 < kNumSoftmaxWG; ++g2) {
                            const float* o = mg + g2 * (kFwdFields * kBlockM) + row_in_block;
                            const float mo = o[1 * kBlockM];
                            if (mo > kNegBig) total += o[0] * ex2_approx(logit2<kLoss>(p, mo) - top);
                        }
                    }
                    // ... and publish one partial per (CTA, segment) for the finalize kernel
                    // (the partial's slot counts the segments of THIS window: that is how the finalize kernel finds it)
                    float* dst = (kWin ? p.win[wi].part : p.part) +
                                 (static_cast<size_t>(blockIdx.x) * (kWin ? p.win[wi].max_segs : p.max_segs) + seg_w) *
                                     (kFwdFields * kBlockM) +
                                 row_in_block;
                    __stcg(dst + 0 * kBlockM, total);
                    __stcg(dst + 1 * kBlockM, m);
                    __stcg(dst + 2 * kBlockM, mp);
                    __stcg(dst + 3 * kBlockM, mf);
                    __stcg(dst + 4 * kBlockM, pm);
                    if constexpr (kRecord) {
                        // flush this row's shared-memory candidates of the segment to the global list (one global atomic
                        // per row that has any) and leave the buffer of this parity clean for segment seg + 2
                        uint32_t n_c;
                        asm volatile("ld.shared.b32 %0, [%1];" : "=r"(n_c) : "r"(cand_cnt_addr_f()) : "memory");
                        if (n_c != 0u) {
                            const int slot_r = rb * kBlockM + row_in_block;
                            const uint32_t base_c = atomicAdd(p.cand_cnt + slot_r, n_c);    // may exceed kCandMax: fallback
                            for (uint32_t i = 0; i < n_c && i < static_cast<uint32_t>(kCandMax); ++i) {
                                uint32_t col;
                                asm volatile("ld.shared.b32 %0, [%1];" : "=r"(col) : "r"(cand_list_addr_f() + i * 4) : "memory");
                                if (base_c + i < static_cast<uint32_t>(kCandMax))
                                    p.cand[static_cast<size_t>(slot_r) * kCandMax + base_c + i] = static_cast<int>(col);
                            }
                            asm volatile("st.shared.b32 [%0], %1;" ::"r"(cand_cnt_addr_f()), "r"(0u) : "memory");
                        }
                        asm volatile("st.shared.b32 [%0], %1;" ::"r"(cand_wrong_addr_f()), "r"(0u) : "memory");
                    }
                }
            }
            if (threadIdx.x == 0 && seg < 2) cta_stamp(p, 2 + 2 * seg);
            idx0 += seg_len;
            ++seg;
            ++seg_w;
            ++rb;
            j0 = 0;
        }
        base += n_w;
        }
    }

    // ================================ teardown ================================
    tc_fence_before_sync();
    __syncthreads();
    ktrace_end(p.ktrace, kBackward ? 3 : 1);
#if SIMCLR_TRACE
    if (threadIdx.x == 0) {
        cta_stamp(p, 5);
        if (p.ktrace != nullptr) p.ktrace[64 + blockIdx.x * 8 + 7] = clock64();
    }
#endif
    if (warp == kAllocWarp) {
        tc_fence_after_sync();
        tmem_dealloc(load_tmem_base(), kTmemCols);
    }
}

}  // namespace simclr
