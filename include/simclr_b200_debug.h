/*
 * simclr_b200_debug -- diagnostics of the tracing build (pytorch-simclr_b200/lib/libsimclr_b200_trace.so, compiled with
 * -DSIMCLR_TRACE=1).  The product library libsimclr_b200.so exports NONE of these symbols and carries none of the
 * probe / self-test kernels or the per-role stamps; the tools under tools/ and tests/test_primitives_gpu.py load the
 * tracing build, which additionally exports everything include/simclr_b200.h declares.
 */
#ifndef SIMCLR_B200_DEBUG_H_
#define SIMCLR_B200_DEBUG_H_

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

/* Diagnostics: per-role clock64() timeline of one CTA of the tile kernels (tools/trace_timeline.py).
 * device_buffer: int64[6 roles][64 iterations][4] or NULL to switch tracing off. */
int simclr_debug_set_trace(void* device_buffer, int cta);

/* Diagnostics: kernel-level %globaltimer timeline (tools/kernel_timeline.py).  device_buffer: uint64[8][2]
 * (min start / max end in ns per kernel id: 0 prepare, 1 forward tile, 2 backward prepare, 3 backward tile),
 * start slots initialised to ~0, end slots to 0; NULL switches it off.  Captured into graphs at capture time. */
int simclr_debug_set_kernel_trace(void* device_buffer);

/* Diagnostics: tcgen05.mma issue / execution rate probe under contention (tools/mma_rate.py).
 * out: int64[4 configs][4]; mode 0 idle, 1 tcgen05.ld, 2 MUFU, 3 FFMA, 4 all; sink: float[640] scratch. */
int simclr_debug_mma_rate(long long* out_device, int batches, int grid, int mode, float* sink, void* stream);

/* Diagnostics: throughput of the softmax warps' per-chunk arithmetic without MMA/TMA (tools/chunk_rate.py).
 * out: int64[10 variants][32 warps] cycles per 32-column chunk (CTA 0); sink: float[640] scratch. */
int simclr_debug_chunk_rate(long long* out_device, int iters, int grid, int nwarps, float k2, float* sink, void* stream);

/* Diagnostics: issue rate of single SASS opcodes with 1..16 warps per SM (tools/pipe_rate.py). out: int64[32]. */
int simclr_debug_pipe_rate(long long* out_device, int iters, int grid, int nwarps, float* sink, void* stream);

/* Diagnostics: UMMA/TMA primitive self-test (tests/test_primitives.py). out_f32 receives 3*128*128 floats. */
int simclr_selftest_umma(const void* a_bf16_128x128, const void* b_bf16_128x128, float* out_f32, void* stream);

#ifdef __cplusplus
}
#endif
#endif /* SIMCLR_B200_DEBUG_H_ */
