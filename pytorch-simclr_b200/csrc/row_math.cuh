// Row-wise scalar helpers shared by the prologue kernels and the fused finalize steps of the tile kernels.
#pragma once

#include <cuda_bf16.h>

#include "sm100_ptx.cuh"

namespace simclr {

constexpr float kNormEps = 1e-12f;          // F.normalize eps (reference objective.py:26-27, :77-78)
constexpr float kSoftplusBeta = 0.8f;       // reference objective.py:70-71
constexpr float kSoftplusThreshold = 20.f;  // torch default threshold of F.softplus
constexpr float kInvNormClamped = 1e12f;    // marker: the norm was clamped by eps
constexpr int kMaxDimPerLane = 8;           // Dpad <= 256 -> at most 8 elements per lane
constexpr float kLn2 = 0.6931471805599453f;

template <typename T>
SIMCLR_DEVICE float load_as_float(const T* p);
template <>
SIMCLR_DEVICE float load_as_float<float>(const float* p) { return __ldg(p); }
template <>
SIMCLR_DEVICE float load_as_float<__nv_bfloat16>(const __nv_bfloat16* p) { return __bfloat162float(*p); }

SIMCLR_DEVICE void store_from_float(float* p, float v) { *p = v; }
SIMCLR_DEVICE void store_from_float(__nv_bfloat16* p, float v) { *p = __float2bfloat16_rn(v); }

// element `idx` of an f32 or bf16 array chosen at run time
SIMCLR_DEVICE float load_elem(const void* base, size_t idx, int is_bf16) {
    return is_bf16 ? __bfloat162float(static_cast<const __nv_bfloat16*>(base)[idx])
                   : __ldg(static_cast<const float*>(base) + idx);
}
SIMCLR_DEVICE void store_elem(void* base, size_t idx, int is_bf16, float v) {
    if (is_bf16) static_cast<__nv_bfloat16*>(base)[idx] = __float2bfloat16_rn(v);
    else static_cast<float*>(base)[idx] = v;
}

SIMCLR_DEVICE float softplus_beta(float x) {
    const float bx = kSoftplusBeta * x;
    return bx > kSoftplusThreshold ? x : log1pf(expf(bx)) / kSoftplusBeta;
}
SIMCLR_DEVICE float softplus_beta_grad(float x) {
    const float bx = kSoftplusBeta * x;
    return bx > kSoftplusThreshold ? 1.f : 1.f / (1.f + expf(-bx));
}

SIMCLR_DEVICE void named_bar_arrive(int id, int nthreads) {
    asm volatile("bar.arrive %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}
// bar.arrive that the compiler cannot move: the 16 registers of the chunk about to be processed pass through the
// statement as in/out operands, so every use of them (the chunk's arithmetic) stays behind it, and the statement itself
// stays behind the tcgen05.wait::ld that produced them.  ptxas otherwise floats the arrive to the end of the chunk
// (observed: the ping-pong token then reaches the other softmax pair a whole chunk late, +12 us per forward kernel).
SIMCLR_DEVICE void named_bar_arrive_pinned(int id, int nthreads, uint32_t (&r)[16]) {
    asm volatile("bar.arrive %16, %17;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 : "r"(id), "r"(nthreads)
                 : "memory");
}
SIMCLR_DEVICE void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

}  // namespace simclr
