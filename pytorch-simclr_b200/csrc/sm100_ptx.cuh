// Thin inline-PTX layer for sm_100a: mbarrier, TMA (cp.async.bulk[.tensor]), tcgen05 (TMEM alloc,
// UMMA issue/commit, TMEM load/store) and the shared-memory / instruction descriptors they consume.
// Hand-written for this repository; bit layouts follow the PTX ISA "tcgen05" chapter (matrix
// descriptor, instruction descriptor for .kind::f16).
#pragma once

#include <cstdint>
#include <cuda_runtime.h>

namespace simclr {

#define SIMCLR_DEVICE __device__ __forceinline__

// ---------------------------------------------------------------------------------------------
// Error reporting from device code: a watchdog on every mbarrier wait turns a protocol bug into a
// trap (reported to the host as a launch failure) instead of a hung GPU.
// ---------------------------------------------------------------------------------------------
#ifndef SIMCLR_WATCHDOG_CYCLES
#define SIMCLR_WATCHDOG_CYCLES (4000000000ll)   // ~2-3 s at B200 clocks
#endif

SIMCLR_DEVICE uint32_t smem_u32(const void* p) {
    return static_cast<uint32_t>(__cvta_generic_to_shared(p));
}

SIMCLR_DEVICE uint32_t lane_id() {
    uint32_t l;
    asm volatile("mov.u32 %0, %%laneid;" : "=r"(l));
    return l;
}

SIMCLR_DEVICE bool elect_one() {
    uint32_t pred = 0;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "elect.sync _|p, 0xffffffff;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ---------------------------------------------------------------------------------------------
// Optional kernel-level timeline (debug): every kernel stamps min(start) / max(end) of %globaltimer (ns) into
// ktrace[2*id], ktrace[2*id+1] when a buffer has been installed with simclr_debug_set_kernel_trace.
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE unsigned long long global_timer_ns() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
SIMCLR_DEVICE void ktrace_begin(unsigned long long* ktrace, int id) {
    if (ktrace != nullptr && threadIdx.x == 0) atomicMin(ktrace + 2 * id, global_timer_ns());
}
SIMCLR_DEVICE void ktrace_end(unsigned long long* ktrace, int id) {
    if (ktrace != nullptr && threadIdx.x == 0) atomicMax(ktrace + 2 * id + 1, global_timer_ns());
}

// ---------------------------------------------------------------------------------------------
// Programmatic dependent launch.  Every kernel of this library is launched with the programmatic-stream-serialization
// attribute: it may start while its predecessor in the stream is still running, executes pdl_launch_dependents() at
// once (so that its own successor can be staged the same way) and pdl_wait() before it touches global memory -- the
// wait returns when the predecessor grid has completed and its writes are visible.  Because every kernel waits, grid
// completion stays transitive along the chain.  Hides the ~1.3 us launch gap between the six kernels of a step.
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE void pdl_launch_dependents() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }
SIMCLR_DEVICE void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// mbarrier
// ---------------------------------------------------------------------------------------------
// Barriers are addressed by their 32-bit shared-space address (`smem_u32(ptr)` once per kernel, not per call: the
// generic -> shared conversion is a chain of S2UR / ULEA / ULOP3 that otherwise sits in front of every wait).
SIMCLR_DEVICE void mbar_init(uint32_t bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}

SIMCLR_DEVICE void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}

SIMCLR_DEVICE void mbar_arrive(uint32_t bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}

SIMCLR_DEVICE void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes) : "memory");
}

SIMCLR_DEVICE bool mbar_try_wait(uint32_t bar, uint32_t parity) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity)
        : "memory");
    return done != 0;
}
// try_wait that lets the hardware park the thread for up to `ns` nanoseconds before it reports "not yet"
SIMCLR_DEVICE bool mbar_try_wait_parked(uint32_t bar, uint32_t parity, uint32_t ns) {
    uint32_t done;
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t"
        "}\n"
        : "=r"(done)
        : "r"(bar), "r"(parity), "r"(ns)
        : "memory");
    return done != 0;
}

// Blocking wait with a watchdog.  `tag` identifies the barrier in the trap message.  The spin loop is three
// instructions (a waiting warp shares its sub-partition's issue slots with the softmax warps: ten-instruction spins of
// five control warps were a quarter of all instructions the backward kernel executed); the clock is only read every
// 4096 polls.
#ifndef SIMCLR_PARK_NS
#define SIMCLR_PARK_NS 2000u
#endif
SIMCLR_DEVICE void mbar_wait(uint32_t bar, uint32_t parity, int tag) {
    if (mbar_try_wait(bar, parity)) return;
    long long t0 = 0;
    uint32_t polls = 0;
    while (!mbar_try_wait_parked(bar, parity, SIMCLR_PARK_NS)) {
        if ((++polls & 4095u) == 0u) {
            const long long now = clock64();
            if (t0 == 0) t0 = now;
            if (now - t0 > SIMCLR_WATCHDOG_CYCLES) {
                printf("[simclr_b200] mbarrier watchdog: block %d thread %d tag %d parity %u\n", (int)blockIdx.x,
                       (int)threadIdx.x, tag, parity);
                __trap();
            }
        }
    }
}

// pointer forms (self-test and probe kernels)
SIMCLR_DEVICE void mbar_init(uint64_t* bar, uint32_t count) { mbar_init(smem_u32(bar), count); }
SIMCLR_DEVICE void mbar_arrive(uint64_t* bar) { mbar_arrive(smem_u32(bar)); }
SIMCLR_DEVICE void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) { mbar_arrive_expect_tx(smem_u32(bar), bytes); }
SIMCLR_DEVICE void mbar_wait(uint64_t* bar, uint32_t parity, int tag) { mbar_wait(smem_u32(bar), parity, tag); }

// ---------------------------------------------------------------------------------------------
// Proxy fences
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE void fence_proxy_async_smem() {
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}

// ---------------------------------------------------------------------------------------------
// TMA
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE void tma_prefetch_desc(const void* tmap) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(tmap)) : "memory");
}

// 2-D tiled load: box lands at `smem_dst`, completes `bytes` on `bar`.  x = inner (contiguous) coord.
SIMCLR_DEVICE void tma_load_2d(uint32_t smem_dst, const void* tmap, uint32_t bar, int32_t x, int32_t y) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(tmap)), "r"(bar), "r"(x), "r"(y)
        : "memory");
}
SIMCLR_DEVICE void tma_load_2d(void* smem_dst, const void* tmap, uint64_t* bar, int32_t x, int32_t y) {
    tma_load_2d(smem_u32(smem_dst), tmap, smem_u32(bar), x, y);
}

// 1-D bulk copy global -> shared (16-byte aligned, size multiple of 16).
SIMCLR_DEVICE void bulk_load_1d(uint32_t smem_dst, const void* gmem_src, uint32_t bytes, uint32_t bar) {
    asm volatile(
        "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
        ::"r"(smem_dst), "l"(reinterpret_cast<uint64_t>(gmem_src)), "r"(bytes), "r"(bar)
        : "memory");
}

// 2-D tiled reduction shared -> global: every element of the box is ADDED to the tensor (element type and swizzle
// come from the tensor map).  Completion is tracked by the thread's bulk async-group.
SIMCLR_DEVICE void tma_reduce_add_2d(const void* tmap, uint32_t smem_src, int32_t x, int32_t y) {
    asm volatile("cp.reduce.async.bulk.tensor.2d.global.shared::cta.add.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(x), "r"(y)
                 : "memory");
}
// 2-D tiled store shared -> global (plain overwrite of the box)
SIMCLR_DEVICE void tma_store_2d(const void* tmap, uint32_t smem_src, int32_t x, int32_t y) {
    asm volatile("cp.async.bulk.tensor.2d.global.shared::cta.tile.bulk_group [%0, {%2, %3}], [%1];"
                 ::"l"(reinterpret_cast<uint64_t>(tmap)), "r"(smem_src), "r"(x), "r"(y)
                 : "memory");
}
SIMCLR_DEVICE void bulk_commit_group() { asm volatile("cp.async.bulk.commit_group;" ::: "memory"); }
// all groups of this thread have finished READING their shared-memory source
SIMCLR_DEVICE void bulk_wait_group_read0() { asm volatile("cp.async.bulk.wait_group.read 0;" ::: "memory"); }
// all groups of this thread are complete (writes performed)
SIMCLR_DEVICE void bulk_wait_group0() { asm volatile("cp.async.bulk.wait_group 0;" ::: "memory"); }

SIMCLR_DEVICE void sts_v4(uint32_t addr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "r"(a), "r"(b), "r"(c), "r"(d) : "memory");
}

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM management
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE void tmem_alloc(uint32_t* smem_result, uint32_t ncols) {   // whole warp, .sync.aligned
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(smem_result)),
                 "r"(ncols)
                 : "memory");
}
SIMCLR_DEVICE void tmem_relinquish() {
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
SIMCLR_DEVICE void tmem_dealloc(uint32_t taddr, uint32_t ncols) {        // whole warp
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
SIMCLR_DEVICE void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
SIMCLR_DEVICE void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// tcgen05: descriptors
// ---------------------------------------------------------------------------------------------
// Shared-memory matrix descriptor (64 bit):
//   [0,14)  start address >> 4        [16,30) leading-dim byte offset >> 4
//   [32,46) stride-dim byte offset >> 4   [46,48) version = 1 (sm_100)
//   [49,52) base offset = 0           [52]    LBO mode = 0        [61,64) swizzle: 2 = 128 B
constexpr uint64_t kDescSwizzle128 = 2ull << 61;
constexpr uint64_t kDescVersion = 1ull << 46;

SIMCLR_DEVICE uint64_t make_smem_desc(uint32_t smem_addr, uint32_t lbo_bytes, uint32_t sbo_bytes) {
    return kDescSwizzle128 | kDescVersion | (static_cast<uint64_t>((sbo_bytes >> 4) & 0x3fff) << 32) |
           (static_cast<uint64_t>((lbo_bytes >> 4) & 0x3fff) << 16) | static_cast<uint64_t>((smem_addr >> 4) & 0x3fff);
}

// Instruction descriptor for .kind::f16 with bf16 inputs and fp32 accumulation (32 bit):
//   [4,6) D format: 1 = f32      [7,10) A format: 1 = bf16     [10,13) B format: 1 = bf16
//   [15] A major (0 = K, 1 = MN) [16] B major (0 = K, 1 = MN)  [17,23) N >> 3   [24,29) M >> 4
__host__ __device__ constexpr uint32_t make_idesc_bf16(uint32_t m, uint32_t n, uint32_t a_mn_major, uint32_t b_mn_major) {
    return (1u << 4) | (1u << 7) | (1u << 10) | (a_mn_major << 15) | (b_mn_major << 16) | ((n >> 3) << 17) |
           ((m >> 4) << 24);
}

// ---------------------------------------------------------------------------------------------
// tcgen05: MMA issue (one thread), commit
// ---------------------------------------------------------------------------------------------
// D[tmem] (+)= A[smem] * B[smem]
SIMCLR_DEVICE void umma_ss(uint32_t tmem_d, uint64_t desc_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "l"(desc_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]
SIMCLR_DEVICE void umma_ts(uint32_t tmem_d, uint32_t tmem_a, uint64_t desc_b, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t"
        ".reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t"
        "}\n" ::"r"(tmem_d),
        "r"(tmem_a), "l"(desc_b), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrive on `bar` once every previously issued MMA of this thread has completed (implies
// tcgen05.fence::before_thread_sync).
SIMCLR_DEVICE void umma_commit(uint32_t bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(bar) : "memory");
}
SIMCLR_DEVICE void umma_commit(uint64_t* bar) { umma_commit(smem_u32(bar)); }

// ---------------------------------------------------------------------------------------------
// tcgen05: TMEM <-> registers.  Shape 32x32b: lane i of the warp owns TMEM lane (warp%4)*32 + i and
// receives `N` consecutive 32-bit columns.
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
        "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]),
          "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]),
          "=r"(r[24]), "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
// 16-column variants.  The wait takes the destination registers as in/out operands so that the compiler cannot move
// a use of them above the wait (the load is asynchronous: the registers are only valid after tcgen05.wait::ld).
SIMCLR_DEVICE void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]),
          "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
SIMCLR_DEVICE void tmem_ld_wait16(uint32_t (&r)[16]) {
    asm volatile("tcgen05.wait::ld.sync.aligned;"
                 : "+r"(r[0]), "+r"(r[1]), "+r"(r[2]), "+r"(r[3]), "+r"(r[4]), "+r"(r[5]), "+r"(r[6]), "+r"(r[7]),
                   "+r"(r[8]), "+r"(r[9]), "+r"(r[10]), "+r"(r[11]), "+r"(r[12]), "+r"(r[13]), "+r"(r[14]), "+r"(r[15])
                 :
                 : "memory");
}
// four columns; waits for the data (rare paths that want small code rather than a deep load pipeline)
SIMCLR_DEVICE void tmem_ld4_sync(uint32_t taddr, uint32_t (&r)[4]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x4.b32 {%0, %1, %2, %3}, [%4];\n\t"
        "tcgen05.wait::ld.sync.aligned;"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3])
        : "r"(taddr)
        : "memory");
}
SIMCLR_DEVICE void tmem_st8(uint32_t taddr, const uint32_t (&r)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr),
                 "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7])
                 : "memory");
}
SIMCLR_DEVICE void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

SIMCLR_DEVICE void tmem_st16(uint32_t taddr, const uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x16.b32 [%0], "
        "{%1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, %16};" ::"r"(taddr),
        "r"(r[0]), "r"(r[1]), "r"(r[2]), "r"(r[3]), "r"(r[4]), "r"(r[5]), "r"(r[6]), "r"(r[7]), "r"(r[8]), "r"(r[9]),
        "r"(r[10]), "r"(r[11]), "r"(r[12]), "r"(r[13]), "r"(r[14]), "r"(r[15])
        : "memory");
}
// 32 columns of one value (used to zero an accumulator)
SIMCLR_DEVICE void tmem_st32_fill(uint32_t taddr, uint32_t v) {
    asm volatile(
        "tcgen05.st.sync.aligned.32x32b.x32.b32 [%0], "
        "{%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, "
        "%1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1, %1};" ::"r"(taddr), "r"(v)
        : "memory");
}
SIMCLR_DEVICE void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ---------------------------------------------------------------------------------------------
// Math helpers
// ---------------------------------------------------------------------------------------------
SIMCLR_DEVICE float ex2_approx(float x) {
    float y;
    asm("ex2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
SIMCLR_DEVICE float lg2_approx(float x) {
    float y;
    asm("lg2.approx.ftz.f32 %0, %1;" : "=f"(y) : "f"(x));
    return y;
}
// Two fp32 -> packed bf16x2 with round-to-nearest-even; `lo` lands in bits [0,16).
SIMCLR_DEVICE uint32_t pack_bf16x2(float lo, float hi) {
    uint32_t r;
    asm("cvt.rn.bf16x2.f32 %0, %1, %2;" : "=r"(r) : "f"(hi), "f"(lo));
    return r;
}
// Packed fp32x2 arithmetic (sm_100: FADD2 / FMUL2 / FFMA2 work on an aligned register pair): one issue slot for two
// results.  The softmax warps are bound by issue slots, not by the FMA pipe, so every pair of independent fp32
// operations that can share an instruction is a slot gained.  A 64-bit value holds {lo, hi}.
using f32x2 = unsigned long long;
SIMCLR_DEVICE f32x2 pack2(float lo, float hi) {
    f32x2 r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
    return r;
}
SIMCLR_DEVICE void unpack2(f32x2 v, float& lo, float& hi) { asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(v)); }
SIMCLR_DEVICE f32x2 add2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("add.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMCLR_DEVICE f32x2 sub2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMCLR_DEVICE f32x2 mul2(f32x2 a, f32x2 b) {
    f32x2 r;
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(r) : "l"(a), "l"(b));
    return r;
}
SIMCLR_DEVICE f32x2 fma2(f32x2 a, f32x2 b, f32x2 c) {
    f32x2 r;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(a), "l"(b), "l"(c));
    return r;
}
// the same fp32 constant in both halves (ptxas folds it into the instruction as a 32-bit immediate)
SIMCLR_DEVICE constexpr f32x2 splat2_bits(unsigned int bits) {
    return (static_cast<f32x2>(bits) << 32) | static_cast<f32x2>(bits);
}

SIMCLR_DEVICE float warp_sum(float v) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    return v;
}
// Vector reduction to global memory (sm_90+): one 16-byte atomic add of four floats.
SIMCLR_DEVICE void red_add_v4(float* gptr, float a, float b, float c, float d) {
    asm volatile("red.global.add.v4.f32 [%0], {%1, %2, %3, %4};" ::"l"(gptr), "f"(a), "f"(b), "f"(c), "f"(d)
                 : "memory");
}

}  // namespace simclr
