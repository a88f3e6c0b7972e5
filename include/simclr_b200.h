/*
 * simclr_b200 -- C ABI of the B200-native contrastive-objective hot path.
 *
 * Drop-in target: the two callables of the reference's objective.py
 *     contrastive_loss(x_batch1, x_batch2, temperature, normalize, weight)   objective.py:6-55
 *     modified_contrastive_loss(x_batch1, x_batch2, **kwargs)                objective.py:58-98
 * as called from utils/model_utils.py:30 (eval, forward only) and :115,:120 (train, forward+backward).
 * The reference has no FFI of its own (pure PyTorch); these entry points are what a ctypes binding in a
 * replacement objective.py binds.  Plain pointers and sizes only; every pointer is
 * a BORROWED device pointer that must stay valid until the stream reaches the end of the call.  Nothing
 * here synchronises the host.  All functions are re-entrant for distinct streams/workspaces.
 *
 * Layout convention ("view-padded"): B images, two views.  Per-row arrays are float[2 * Bpad] and
 * operand matrices are bf16[2 * Bpad][Dpad], Bpad = simclr_pad_rows(B) (multiple of 128), Dpad =
 * simclr_pad_dim(d) (64, 128 or 256).  Slot (v, i) lives at index v * Bpad + i; padding slots are zero.
 *
 * Return value: 0 on success; a negative SIMCLR_ERR_* for rejected arguments; a positive value is a
 * cudaError_t reported by the CUDA runtime.  There is no CPU fallback.
 */
#ifndef SIMCLR_B200_H_
#define SIMCLR_B200_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define SIMCLR_ABI_VERSION 13

/* loss kinds */
#define SIMCLR_LOSS_NTXENT 0   /* objective.py:6-55  */
#define SIMCLR_LOSS_MODIFIED 1 /* objective.py:58-98 */

/* input element types of x_batch1 / x_batch2 (and of the returned gradients) */
#define SIMCLR_DTYPE_F32 0
#define SIMCLR_DTYPE_BF16 1

/* arithmetic of the tensor-core products */
#define SIMCLR_PRECISION_BF16 0  /* bf16 operands, fp32 accumulate: loss 2e-3, gradients 1e-2 (the fast path) */
#define SIMCLR_PRECISION_SPLIT 1 /* hi + lo bf16 planes, three products each: fp32-grade, loss 1e-5, gradients 1e-4;
                                    d <= 128, single GPU or NCCL transport; operand buffers hold two planes */

/* flags of the backward / fused entry points */
#define SIMCLR_FLAG_DETERMINISTIC 1 /* bit-identical gradients from run to run: every (CTA, segment) of the backward tile kernel
                                       stores its accumulator into its own slot and the finalize kernel adds the slots of a row
                                       block in a fixed order, instead of TMA reduce-adds into one buffer in completion order
                                       (the reference's cudnn.deterministic / manual_seed switch, pretrain.py:59-61).  Needs the
                                       larger workspace of simclr_backward_workspace_bytes_flags. */

/* error codes */
#define SIMCLR_OK 0
#define SIMCLR_ERR_NULL_POINTER (-1)
#define SIMCLR_ERR_BAD_SHAPE (-2)       /* B < 1, d < 1, shard outside the global batch */
#define SIMCLR_ERR_UNSUPPORTED_DIM (-3) /* d > 256 */
#define SIMCLR_ERR_BAD_DTYPE (-4)
#define SIMCLR_ERR_WORKSPACE_TOO_SMALL (-5)
#define SIMCLR_ERR_MISALIGNED (-6)      /* operand / vector pointers must be 16-byte aligned */
#define SIMCLR_ERR_BAD_TEMPERATURE (-7)
#define SIMCLR_ERR_NOT_SM100 (-8)       /* device is not compute capability 10.x */
#define SIMCLR_ERR_DRIVER_ENTRY (-9)    /* cuTensorMapEncodeTiled unavailable */
#define SIMCLR_ERR_TENSOR_MAP (-10)
#define SIMCLR_ERR_BAD_LOSS (-11)
#define SIMCLR_ERR_BAD_PEERS (-12)

int simclr_abi_version(void);
const char* simclr_error_string(int code);

/* Bpad / Dpad helpers (pure host arithmetic). simclr_pad_dim returns 0 when d is unsupported. */
int64_t simclr_pad_rows(int64_t b);
int64_t simclr_pad_dim(int64_t d);

/* Bytes of the operand matrix of b images (0 when unsupported): 2*Bpad*Dpad bf16, twice that in split precision
 * (plane 0 = hi, plane 1 = lo, each [2*Bpad][Dpad]). */
size_t simclr_operand_bytes(int64_t b, int64_t d, int precision);

/* Bytes of scratch each stage needs.  b_local == b_global on a single GPU. */
size_t simclr_forward_workspace_bytes(int loss, int64_t b_local, int64_t b_global, int64_t d);
size_t simclr_backward_workspace_bytes(int loss, int64_t b_local, int64_t b_global, int64_t d);
size_t simclr_backward_workspace_bytes_flags(int loss, int64_t b_local, int64_t b_global, int64_t d, int flags);

/*
 * Stage 1 -- prologue (objective.py:25-30 L2 normalise, or :70-78 softplus + L1 normalise).
 * Reads x_batch1 / x_batch2 ([b_local][d], row-major contiguous, `in_dtype`), writes
 *   operand  bf16 [2*Blpad][Dpad]  round-to-nearest operands for the tensor cores (padding zeroed).  NT-Xent rows
 *                                  carry a factor sqrt(log2(e)/temperature), so that the tensor-core product of
 *                                  two rows is the logit in the log2 domain (the same temperature must be passed
 *                                  to simclr_forward / simclr_backward); the modified loss ignores `temperature`
 *   inv_norm f32  [2*Blpad]        1 / max(||z||, 1e-12)   (1 when normalize == 0)
 *   pos_dot  f32  [2*Blpad]        exact fp32 <op_r, op_pos(r)> of the positive pair
 * `forward_workspace` (may be NULL) is the workspace the following simclr_forward call will use; its counters (the
 * ticket header and the per-row candidate counters of the exact accuracy count) are zeroed here so that no separate
 * memset is needed.  simclr_forward requires them to be zero on entry and leaves them zero on exit.
 */
int simclr_prepare(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d, int in_dtype,
                   int normalize, float temperature, void* operand, float* inv_norm, float* pos_dot,
                   void* forward_workspace, void* stream);

/*
 * Stage 2 -- forward over this rank's rows against the global batch's columns
 * (objective.py:35-53 / :87-97).  `operand_cols` is the [2*Bgpad][Dpad] operand of the global batch in
 * view-padded order (== operand_rows on one GPU).  Outputs:
 *   lse2     f32 [2*Blpad]  log2-domain log-sum-exp of every local row (saved for backward)
 *   row_loss f32 [2*Blpad]  per-row loss L_r (natural log)
 *   stats    f32 [4]        { sum_r w_r L_r, sum_r w_r, #rows whose first-argmax is the positive,
 *                             stats[0] / stats[1] }   -- local to this rank
 *   loss_out f32 [1]        optional separate copy of stats[3] (so a caller can hand out a tensor it may
 *                           modify in place, as utils/model_utils.py:31,116 does); may be NULL
 * `row_weight` is the reference's `weight` argument restricted to this rank: f32 [2*b_local] in the
 * reference's compact order (view-1 rows then view-2 rows), or NULL.
 * `normalize` must repeat the value given to simclr_prepare: normalised rows bound the scores (|S| <= 1), which
 * lets the NT-Xent kernel use a constant softmax shift instead of a running maximum.
 * Accuracy count (stats[2], objective.py:51-53 / :95-97: first argmax == positive): with bf16 operands a tensor-core score
 * is only known to within 2^-8 of the largest possible score, so rows whose best negative lies within that band of the
 * exact positive are decided by re-scoring, in exact fp32 and in the reference's own logit expression, the columns of the
 * (at most 8) recorded 16-column chunks whose maximum lies inside the band -- which needs the exact rows: `x_batch1` /
 * `x_batch2` / `in_dtype` / `inv_norm` as given to / produced by simclr_prepare (one GPU, b_local == b_global; or zrows_* of
 * simclr_forward_peer).  All NULL, normalize == 0 for NT-Xent, or more than 8 such chunks for one row: that row is decided
 * on the tensor-core scores (exact ties between identical rows stay exact either way).  In the fused one-GPU step
 * (simclr_forward_backward) the re-scoring runs in spare blocks of the backward finalize kernel.
 */
int simclr_forward(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                   int64_t row_offset, int64_t d, float temperature, int normalize, const float* pos_dot,
                   const float* row_weight, float* lse2, float* row_loss, float* stats, float* loss_out, void* workspace,
                   size_t workspace_bytes, const void* x_batch1, const void* x_batch2, int in_dtype, const float* inv_norm,
                   void* stream);

/*
 * Stage 3 -- backward: gradients of  grad_out * loss  with respect to x_batch1 / x_batch2 of this rank.
 *   lse2_cols  f32 [2*Bgpad]  lse2 of the global batch (== lse2 on one GPU)
 *   col_scale  f32 [2*Bgpad]  w_c / sum(w) in view-padded order, or NULL for the unweighted 1/(2B)
 *   grad_out   f32 [1] on the device, or NULL for 1.0
 *   grad1/2    [b_local][d] in `in_dtype`
 *   primed_colvec  NULL, or the column vectors f32 [2][2*Bgpad] a forward call already produced together with a
 *              zeroed accumulation buffer in `workspace` (simclr_forward_peer with `backward_workspace`): the
 *              backward-prepare kernel is then skipped and lse2_cols may be NULL.  On one GPU the vectors sit at the
 *              start of the backward workspace; in the row-sharded batch they are the rank's symmetric colvec buffer.
 */
int simclr_backward(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t b_global,
                    int64_t row_offset, int64_t d, int in_dtype, int normalize, float temperature, int precision,
                    const void* operand_rows, const void* operand_cols, const float* inv_norm, const float* pos_dot,
                    const float* lse2_cols, const float* col_scale, const float* grad_out, void* grad1, void* grad2,
                    void* workspace, size_t workspace_bytes, const float* primed_colvec, int flags, void* stream);

/*
 * Fused training-step form of the three stages for one GPU and an unweighted loss: what the reference's
 *     loss, acc = loss_fn(z1, z2, temperature=t); loss /= accum_steps; loss.backward()      utils/model_utils.py:115-120
 * enqueues when the upstream gradient (`grad_out`, device f32 [1], or NULL for 1.0 -- the reference's 1/accum_steps) is
 * known up front.  Five launches (prepare, forward tile, forward finalize, backward tile, backward finalize).  Because
 * the library sees the whole sequence it takes two things off the path between the tile kernels that the separate
 * calls cannot: the reduction of the loss statistics moves into the backward finalize kernel, and the backward tile
 * kernel loads its operands and issues its first score MMAs while the forward finalize kernel is still running.
 *   rowvec   f32 [4][2*Bpad]   inv_norm | pos_dot | lse2 | row_loss (outputs / saved state, as in the staged calls)
 *   stats    f32 [4], loss_out f32 [1] or NULL: as simclr_forward; valid when the whole call has completed
 *   operand  bf16, simclr_operand_bytes(b, d, precision) bytes
 * Workspaces as for simclr_forward / simclr_backward with b_local == b_global == b.
 */
int simclr_forward_backward(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                            int normalize, float temperature, int precision, const float* grad_out, void* operand,
                            float* rowvec, float* stats, float* loss_out, void* grad1, void* grad2,
                            void* forward_workspace, size_t forward_workspace_bytes, void* backward_workspace,
                            size_t backward_workspace_bytes, int flags, void* stream);

/*
 * The same fused step split around its last kernel, for callers that learn the upstream gradient only after they have
 * seen the loss (torch.autograd: forward() returns, the caller scales the loss -- utils/model_utils.py:116 divides it by
 * the accumulation steps in place -- and backward() delivers grad_output):
 *   simclr_forward_backward_begin   prepare, forward tile, forward finalize (loss statistics complete here: `stats` /
 *                                   `loss_out` are valid as soon as that kernel has run, while the backward tile kernel --
 *                                   launched by the same call -- is still running), backward tile: four launches;
 *   simclr_forward_backward_finish  the backward finalize kernel alone: grad1 / grad2 = grad_out * dloss/dx.  May be
 *                                   repeated (a second backward over a retained graph): it only reads the workspace.
 * Buffers as in simclr_forward_backward; `operand`, `rowvec` and `backward_workspace` must be kept between the calls.
 */
int simclr_forward_backward_begin(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                                  int normalize, float temperature, int precision, void* operand, float* rowvec,
                                  float* stats, float* loss_out, void* forward_workspace, size_t forward_workspace_bytes,
                                  void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream);
int simclr_forward_backward_finish(int loss, const void* x_batch1, const void* x_batch2, int64_t b, int64_t d, int in_dtype,
                                   int normalize, float temperature, int precision, const float* grad_out,
                                   const void* operand, const float* rowvec, void* grad1, void* grad2,
                                   void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream);

/*
 * Projection-head tail.  The reference produces the embeddings as
 *     z = BatchNorm1d(out_dim)(Linear(encoder_dim -> out_dim, bias=False)(.))        models/simclr.py:38-39
 * in two separate forward calls, one per view (utils/model_utils.py:113-114: per-view batch statistics), and passes them
 * to objective.contrastive_loss.  These entry points take the PRE-BatchNorm activations u1 / u2 instead: the BatchNorm
 * apply happens in the registers of the prepare kernel (z is never written), and on the way back the finalize kernel's
 * dL/dz is turned into dL/du, dL/dgamma, dL/dbeta -- loss(BN(u1), BN(u2)) forward and backward with the loss's five kernels
 * plus one statistics kernel in front and one BatchNorm-backward kernel behind.
 *   bn_state  f32 [2 views][5][Dpad]: scale = gamma*rstd | shift = beta - mean*scale | mean | rstd | var_unbiased.
 *             Training: simclr_bn_stats computes it from the batch (biased variance, eps as nn.BatchNorm1d); the caller
 *             updates the module's running statistics from planes 2 and 4.  Eval: the caller fills scale / shift from the
 *             running statistics and passes bn_training = 0 (dL/du = scale * dL/dz).
 *   simclr_bn_t.partial  f32 scratch inside a workspace of simclr_bn_workspace_bytes(b, d) bytes at offset 256 (the first
 *             256 bytes are simclr_bn_stats' ticket: zero before the first call, left zero).
 *   grad_gamma / grad_beta  f32 [d], ACCUMULATED into (zero them first), or NULL.
 * One GPU, unweighted loss; everything else as in simclr_forward_backward_begin / _finish.
 */
typedef struct simclr_bn {
    const float* state;   /* bn_state */
    float* partial;       /* scratch of the backward: per-CTA column sums of dL/dz and dL/dz * xhat */
} simclr_bn_t;
size_t simclr_bn_state_floats(int64_t d);
size_t simclr_bn_workspace_bytes(int64_t b, int64_t d);
int simclr_bn_stats(const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype, const float* gamma, const float* beta,
                    float eps, float* bn_state, void* workspace, size_t workspace_bytes, void* stream);
int simclr_head_forward(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype, int normalize,
                        float temperature, int precision, const float* bn_state, void* operand, float* rowvec, float* stats,
                        float* loss_out, void* forward_workspace, size_t forward_workspace_bytes, void* stream);
int simclr_head_forward_backward_begin(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype,
                                       int normalize, float temperature, int precision, const simclr_bn_t* bn, void* operand,
                                       float* rowvec, float* stats, float* loss_out, void* forward_workspace,
                                       size_t forward_workspace_bytes, void* backward_workspace,
                                       size_t backward_workspace_bytes, int flags, void* stream);
int simclr_head_forward_backward_finish(int loss, const void* u1, const void* u2, int64_t b, int64_t d, int in_dtype,
                                        int normalize, float temperature, int precision, const float* grad_out,
                                        const simclr_bn_t* bn, int bn_training, const void* operand, const float* rowvec,
                                        void* grad_u1, void* grad_u2, float* grad_gamma, float* grad_beta,
                                        void* backward_workspace, size_t backward_workspace_bytes, int flags, void* stream);

/*
 * Row-sharded global batch over peer memory (one process per GPU of one NVLink / NVSwitch node; not in the reference,
 * whose only batch-scaling device is gradient accumulation, utils/model_utils.py:113-123).  Buffers named *_peers are
 * arrays of `world` device pointers: entry r is THIS process's mapping of rank r's copy of a symmetric allocation
 * (torch.distributed._symmetric_memory, CUDA IPC / fabric handles).  Rank r owns images [r*b_local, (r+1)*b_local).
 *
 * simclr_prepare_peer : simclr_prepare that additionally stores every operand row into all ranks' global operand
 *                       matrix bf16 [2*Bgpad][Dpad] (Bgpad = simclr_pad_rows(world*b_local); the padding rows must have
 *                       been zeroed once) -- the operand "all-gather" is these NVLink stores, fused into the kernel.
 *                       `operand_global_multicast` (may be NULL) is the NVLS multicast mapping of the same buffer:
 *                       when given, each row leaves the GPU once (multimem.st) and the NVSwitch replicates it.
 * simclr_forward_peer : simclr_forward whose finalize kernel pushes the backward's column vectors of the local rows
 *                       (a_c | lse2_c) into all ranks' colvec f32 [2][2*Bgpad] (padding zeroed once) and
 *                       {sum w L, sum w, #correct} into slot `rank` of all ranks' stats_all f32 [world][4].
 *                       `backward_workspace` (may be NULL; unweighted losses only) primes the following
 *                       simclr_backward: its accumulation buffer is zeroed by a spare warp of the forward tile kernel
 *                       and, on one GPU (world == 0), the column vectors are written to its start.
 *                       `flag_peers` / `epoch_local` (may be NULL) hand the barrier between simclr_prepare_peer and
 *                       this call to the library: the tiles of the columns this rank produced itself run first,
 *                       while the other ranks' operand rows are still crossing NVLink, then the barrier, then the
 *                       remote columns (the caller must then NOT issue that barrier itself).
 * simclr_peer_barrier : device-side barrier over flags u32 [world] in symmetric memory (zero-initialised once);
 *                       `epoch_local` is a private device counter, also zero-initialised once.  With `stats_all` it
 *                       then sums the per-rank statistics into stats_out[4] / loss_out (the GLOBAL loss, identical on
 *                       all ranks).  Must follow simclr_prepare_peer / simclr_forward_peer in the same stream, on every
 *                       rank, before the pushed data is consumed.  world == 0 / NULL peers degrade to the local calls
 *                       (these are also the entry points that take `precision`; simclr_prepare / simclr_forward are
 *                       their SIMCLR_PRECISION_BF16 single-rank forms).
 * Exact accuracy count across ranks: simclr_prepare_peer additionally writes the exact fp32 normalised rows of this rank
 * into `zrows_local` f32 [2*Blpad][Dpad] (may be NULL); simclr_forward_peer re-scores the candidates of a row from
 * `zrows_peers` (world pointers: every rank's zrows_local in symmetric memory, read over NVLink for the few rows needed)
 * or from `zrows_global` f32 [2*Bgpad][Dpad] (the ranks' zrows_local gathered by a collective); both NULL: from
 * x_batch1 / x_batch2 / inv_norm as in simclr_forward when b_local == b_global, else on the tensor-core scores.
 */
int simclr_prepare_peer(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d, int in_dtype,
                        int normalize, float temperature, int precision, void* operand, float* inv_norm, float* pos_dot,
                        void* forward_workspace, int world, int rank, void* const* operand_global_peers,
                        void* operand_global_multicast, float* zrows_local, void* stream);
int simclr_forward_peer(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                        int64_t row_offset, int64_t d, float temperature, int normalize, int precision,
                        const float* pos_dot,
                        const float* row_weight, float* lse2, float* row_loss, float* stats, float* loss_out,
                        void* workspace, size_t workspace_bytes, void* backward_workspace,
                        size_t backward_workspace_bytes, int world, int rank, void* const* colvec_peers,
                        void* const* stats_peers, void* const* flag_peers, unsigned int* epoch_local,
                        const void* x_batch1, const void* x_batch2, int in_dtype, const float* inv_norm,
                        const float* zrows_global, void* const* zrows_peers, void* stream);
int simclr_peer_barrier(int world, int rank, void* const* flag_peers, unsigned int* epoch_local, const float* stats_all,
                        float* stats_out, float* loss_out, void* stream);

/*
 * Fused row-sharded step: simclr_prepare_peer + simclr_forward_peer + simclr_backward of one rank in five launches,
 * with the two cross-GPU barriers executed INSIDE the tile kernels instead of by simclr_peer_barrier launches: the
 * prepare kernel bumps `epoch_local`, CTA 0 of the forward tile kernel publishes it to every rank's flags and the TMA
 * producer of every CTA waits for all ranks before its first column-tile load (the row-block tile of local rows is
 * already in flight); the forward finalize kernel bumps the epoch again and the backward tile kernel -- whose operand
 * loads and first score MMAs run ahead of it -- does the same before it loads the peers' column vectors.  The backward
 * finalize kernel adds up the ranks' statistics into stats_global / loss_out (the GLOBAL loss, identical on all ranks).
 * `epoch_local` is u32 [2] here (both zero-initialised once): the epoch counter and the ticket with which the prepare
 * kernel's last warp -- the moment this rank's operand push is complete -- publishes the epoch to the peers itself; the
 * forward finalize kernel's last block does the same for the second barrier, so the tile kernels only wait.
 * Tile-aligned shards (b_local % 128 == 0, at least one tile per SM in this rank's own columns): the forward tile kernel
 * is ONE launch over TWO column windows and does the operand exchange itself -- window 0, this rank's own columns, is read
 * from `operand` while the spare warp of every CTA pushes those rows into all ranks' global operand matrices (multimem.st
 * through the NVSwitch when operand_global_multicast is given) and the warp whose stores are performed last publishes
 * the epoch; the TMA producers wait for all ranks' flags only in front of window 1, everybody else's columns.  The NVLink
 * transfer, its drain and the ranks' skew at the start of the step hide under the tiles of window 0 (the prepare kernel
 * then pushes nothing).  The kernel needs all of its CTAs resident at once (grid = number of SMs, one CTA per SM -- the
 * GPU must not be shared with another context's long-running kernels); SIMCLR_B200_PEER_WINDOWS=0 restores the
 * push-in-prepare form.
 * Buffers as in the separate calls; operand / colvec / stats *_peers[rank] are this rank's own copies.  Every rank must
 * issue the same sequence of fused steps and barrier calls (they share the flags and the epoch counter).
 */
int simclr_forward_backward_peer(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t d,
                                 int in_dtype, int normalize, float temperature, const float* grad_out, void* operand,
                                 float* rowvec, float* stats_local, float* stats_global, float* loss_out, void* grad1,
                                 void* grad2, void* forward_workspace, size_t forward_workspace_bytes,
                                 void* backward_workspace, size_t backward_workspace_bytes, int world, int rank,
                                 void* const* operand_global_peers, void* operand_global_multicast,
                                 void* const* colvec_peers, void* const* stats_peers, void* const* flag_peers,
                                 unsigned int* epoch_local, void* const* zrows_peers, int flags, void* stream);

/*
 * Measurement entry points (bench.py): simclr_forward_peer on one GPU / simclr_backward restricted to a subset of their
 * kernels, so that ONE kernel of the step can be timed by itself (a CUDA graph of back-to-back launches of it between
 * two events, reading the state the last full call left).  `stage_mask` is a per-call argument -- there is no process
 * state; results are only meaningful with SIMCLR_STAGE_ALL.  Arguments as in simclr_forward_peer (world == 0) and
 * simclr_backward.
 */
#define SIMCLR_STAGE_PREPARE 1u
#define SIMCLR_STAGE_FORWARD_TILE 2u
#define SIMCLR_STAGE_FORWARD_FINALIZE 4u
#define SIMCLR_STAGE_BACKWARD_TILE 8u
#define SIMCLR_STAGE_BACKWARD_FINALIZE 16u
#define SIMCLR_STAGE_BACKWARD_PREPARE 32u
#define SIMCLR_STAGE_ALL 0xFFFFFFFFu
int simclr_forward_stages(int loss, const void* operand_rows, const void* operand_cols, int64_t b_local, int64_t b_global,
                          int64_t row_offset, int64_t d, float temperature, int normalize, int precision,
                          const float* pos_dot, const float* row_weight, float* lse2, float* row_loss, float* stats,
                          float* loss_out, void* workspace, size_t workspace_bytes, void* backward_workspace,
                          size_t backward_workspace_bytes, const void* x_batch1, const void* x_batch2, int in_dtype,
                          const float* inv_norm, void* stream, unsigned int stage_mask);
int simclr_backward_stages(int loss, const void* x_batch1, const void* x_batch2, int64_t b_local, int64_t b_global,
                           int64_t row_offset, int64_t d, int in_dtype, int normalize, float temperature, int precision,
                           const void* operand_rows, const void* operand_cols, const float* inv_norm, const float* pos_dot,
                           const float* lse2_cols, const float* col_scale, const float* grad_out, void* grad1, void* grad2,
                           void* workspace, size_t workspace_bytes, const float* primed_colvec, int flags, void* stream,
                           unsigned int stage_mask);

#ifdef __cplusplus
}
#endif
#endif /* SIMCLR_B200_H_ */
