// One-CTA self-test of the primitives the tile kernel is built from, checked on the host against numpy
// (tests/test_primitives.py):
//   out[0] = A * B^T   TMA(128B swizzle) -> smem, tcgen05.mma SS, both operands K-major
//   out[1] = A * B     same smem tile of B read as an MN-major operand (K = row index of B)
//   out[2] = A * B     A written to TMEM as packed bf16 with tcgen05.st and used as the TMEM A operand
// A and B are 128 x 128 bf16 row-major.  out is 3 x 128 x 128 fp32.
#pragma once

#include "contrastive_kernels.cuh"

namespace simclr {

constexpr int kSelftestSmemBytes = 2 * 2 * kAtomBytes + 64 + 1024;

__global__ void __launch_bounds__(128, 1)
selftest_umma_kernel(const __grid_constant__ CUtensorMap map_a, const __grid_constant__ CUtensorMap map_b,
                     float* __restrict__ out) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sa = smem;
    uint8_t* sb = smem + 2 * kAtomBytes;
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 4 * kAtomBytes);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;

    if (threadIdx.x == 0) {
        mbar_init(bars + 0, 1);   // TMA landed
        mbar_init(bars + 1, 1);   // MMAs 1+2 done
        mbar_init(bars + 2, 128); // A copied into TMEM
        mbar_init(bars + 3, 1);   // MMA 3 done
        mbar_fence_init();
    }
    if (warp == 0) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_ptr;

    if (threadIdx.x == 0) {
        mbar_arrive_expect_tx(bars + 0, 4 * kAtomBytes);
        for (int ka = 0; ka < 2; ++ka) {
            tma_load_2d(sa + ka * kAtomBytes, &map_a, bars + 0, ka * kAtomK, 0);
            tma_load_2d(sb + ka * kAtomBytes, &map_b, bars + 0, ka * kAtomK, 0);
        }
        mbar_wait(bars + 0, 0, 900);
        tc_fence_after_sync();
        const uint32_t a_addr = smem_u32(sa), b_addr = smem_u32(sb);
        constexpr uint32_t idesc_kk = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t idesc_kmn = make_idesc_bf16(128, 128, 0, 1);
        for (int ka = 0; ka < 2; ++ka)
            for (int kk = 0; kk < 4; ++kk) {
                const uint32_t off = ka * kAtomBytes + kk * 32;
                umma_ss(tmem + 0, make_smem_desc(a_addr + off, 0, 1024), make_smem_desc(b_addr + off, 0, 1024),
                        idesc_kk, (ka | kk) != 0);
            }
        for (int kc = 0; kc < 8; ++kc) {
            const uint32_t aoff = (kc >> 2) * kAtomBytes + (kc & 3) * 32;
            umma_ss(tmem + 128, make_smem_desc(a_addr + aoff, 0, 1024),
                    make_smem_desc(b_addr + kc * 2048, kAtomBytes, 1024), idesc_kmn, kc != 0);
        }
        umma_commit(bars + 1);
    }
    __syncwarp();
    mbar_wait(bars + 1, 0, 901);
    tc_fence_after_sync();

    const uint32_t lane_addr = static_cast<uint32_t>(warp * 32) << 16;
    const int row = warp * 32 + lane;
    for (int t = 0; t < 2; ++t)
        for (int q = 0; q < 4; ++q) {
            uint32_t r[32];
            tmem_ld32(tmem + lane_addr + t * 128 + q * 32, r);
            tmem_ld_wait();
            for (int i = 0; i < 32; ++i) out[(t * 128 + row) * 128 + q * 32 + i] = __uint_as_float(r[i]);
        }

    mbar_wait(bars + 0, 0, 904);   // every thread observes the TMA completion before reading the tile
    // copy A (from its swizzled smem tile) into TMEM columns [256, 320) as packed bf16: thread = row
    for (int q = 0; q < 4; ++q) {          // 32 elements = 64 bytes per step
        uint32_t w[16];
        for (int i = 0; i < 16; ++i) {
            const int k = q * 32 + 2 * i;                        // element index along K
            const int atom = k >> 6, kin = k & 63;
            const int chunk = (kin >> 3) ^ (row & 7);           // 128B swizzle: 16-byte chunk index XOR row%8
            const uint8_t* src = sa + atom * kAtomBytes + row * 128 + chunk * 16 + (kin & 7) * 2;
            w[i] = *reinterpret_cast<const uint32_t*>(src);
        }
        tmem_st16(tmem + lane_addr + 256 + q * 16, w);
    }
    tmem_st_wait();
    tc_fence_before_sync();
    mbar_arrive(bars + 2);

    if (threadIdx.x == 0) {
        mbar_wait(bars + 2, 0, 902);
        tc_fence_after_sync();
        constexpr uint32_t idesc_kmn = make_idesc_bf16(128, 128, 0, 1);
        const uint32_t b_addr = smem_u32(sb);
        for (int kc = 0; kc < 8; ++kc)
            umma_ts(tmem + 384, tmem + 256 + kc * 8, make_smem_desc(b_addr + kc * 2048, kAtomBytes, 1024), idesc_kmn,
                    kc != 0);
        umma_commit(bars + 3);
    }
    __syncwarp();
    mbar_wait(bars + 3, 0, 903);
    tc_fence_after_sync();
    for (int q = 0; q < 4; ++q) {
        uint32_t r[32];
        tmem_ld32(tmem + lane_addr + 384 + q * 32, r);
        tmem_ld_wait();
        for (int i = 0; i < 32; ++i) out[(2 * 128 + row) * 128 + q * 32 + i] = __uint_as_float(r[i]);
    }

    tc_fence_before_sync();
    __syncthreads();
    if (warp == 0) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, 512);
    }
}


// ---------------------------------------------------------------------------------------------
// tcgen05.mma rate probe (tools/mma_rate.py).  Warp 17 issues `batches` batches of 8 MMAs (+ one commit per
// batch, as the tile kernel does) on zeroed smem operands while warps 0-15 optionally generate contention:
//   mode 0: idle   1: tcgen05.ld loop   2: MUFU.EX2 loop   3: FFMA loop   4: tcgen05.ld + MUFU + FFMA
// out[cfg*4 + {0,1,2}] = {issue cycles, cycles until the last commit lands, batches}
//   cfg 0: SS N=128   cfg 1: SS N=256   cfg 2: TS N=128 (A from TMEM)   cfg 3: SS N=128 then TS N=128 alternating
// ---------------------------------------------------------------------------------------------
constexpr int kRateSmemBytes = 3 * 2 * kAtomBytes + 64 + 1024;

__global__ void __launch_bounds__(640, 1) mma_rate_kernel(long long* __restrict__ out, int batches, int mode,
                                                          float* __restrict__ sink) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + 6 * kAtomBytes);
    uint32_t* tmem_ptr = reinterpret_cast<uint32_t*>(bars + 4);
    volatile int* done_flag = reinterpret_cast<volatile int*>(bars + 6);
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    for (int i = threadIdx.x; i < 6 * kAtomBytes / 4; i += blockDim.x) reinterpret_cast<uint32_t*>(smem)[i] = 0u;
    if (threadIdx.x == 0) {
        for (int i = 0; i < 4; ++i) mbar_init(bars + i, 1);
        *done_flag = 0;
        mbar_fence_init();
    }
    if (warp == 18) {
        tmem_alloc(tmem_ptr, 512);
        tmem_relinquish();
    }
    fence_proxy_async_smem();
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = *tmem_ptr;
    if (warp == 17) {
        const uint32_t a_addr = smem_u32(smem), b_addr = smem_u32(smem + 2 * kAtomBytes);
        constexpr uint32_t id128 = make_idesc_bf16(128, 128, 0, 0);
        constexpr uint32_t id256 = make_idesc_bf16(128, 256, 0, 0);
        constexpr uint32_t id128mn = make_idesc_bf16(128, 128, 0, 1);
        int phase = 0;
        for (int cfg = 0; cfg < 4; ++cfg) {
            __syncwarp();
            const long long t0 = clock64();
            for (int b = 0; b < batches; ++b) {
                if (elect_one()) {
                    const bool ts = (cfg == 2) || (cfg == 3 && (b & 1));
                    if (!ts) {
                        const uint32_t id = cfg == 1 ? id256 : id128;
#pragma unroll
                        for (int k = 0; k < 8; ++k) {
                            const uint32_t off = (k >> 2) * kAtomBytes + (k & 3) * 32;
                            umma_ss(tmem + (b & 1) * 128, make_smem_desc(a_addr + off, 0, 1024),
                                    make_smem_desc(b_addr + off, 0, 1024), id, 1);
                        }
                    } else {
#pragma unroll
                        for (int k = 0; k < 8; ++k)
                            umma_ts(tmem + 384, tmem + 256 + k * 8, make_smem_desc(b_addr + k * 2048, kAtomBytes, 1024),
                                    id128mn, 1);
                    }
                    umma_commit(bars + 1);
                }
                __syncwarp();
            }
            const long long t1 = clock64();
            if (elect_one()) umma_commit(bars + 0);
            __syncwarp();
            mbar_wait(bars + 0, phase, 950);
            phase ^= 1;
            const long long t2 = clock64();
            if ((threadIdx.x & 31) == 0) {
                out[cfg * 4 + 0] = t1 - t0;
                out[cfg * 4 + 1] = t2 - t0;
                out[cfg * 4 + 2] = batches;
            }
        }
        __syncwarp();
        if ((threadIdx.x & 31) == 0) *done_flag = 1;
    } else if (warp < 16 && mode != 0) {
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        float acc = threadIdx.x * 1e-3f, acc2 = 1.0f;
        while (*done_flag == 0) {
            if (mode == 1 || mode == 4) {
                uint32_t r[32];
                tmem_ld32(tmem + lane_addr + 256 + (warp >> 2) * 32, r);     // columns outside the MMA accumulators
                tmem_ld_wait();
                acc += __uint_as_float(r[0]) + __uint_as_float(r[31]);
            }
            if (mode == 2 || mode == 4) {
#pragma unroll
                for (int i = 0; i < 32; ++i) acc += ex2_approx(acc2 - i);
            }
            if (mode == 3 || mode == 4) {
#pragma unroll
                for (int i = 0; i < 64; ++i) acc2 = fmaf(acc2, 0.999f, 1e-3f * i);
            }
        }
        if (acc + acc2 == 12345.678f) sink[threadIdx.x] = acc;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == 18) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, 512);
    }
}

}  // namespace simclr
