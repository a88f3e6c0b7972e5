// Throughput probes for the per-element code of the softmax warps (tools/chunk_rate.py): `nwarps` warps of every CTA
// run `iters` times the tile kernel's inner loop over a 64-column half tile (four 16-column chunks, TMEM load of the
// next chunk in flight) with one variant of the chunk arithmetic and report cycles per half tile.  No MMA, no TMA, no barriers: this is the bound the softmax side alone puts on a tile
// (a 128 x 128 tile = 4 chunks for each of the 16 warps, i.e. 16 chunk-times per SM sub-partition).
#pragma once

#include "contrastive_kernels.cuh"

namespace simclr {

enum ChunkVariant : int {
    kVarFwdProd = 0,      // fwd_chunk_fast<NT-Xent, const shift> as shipped (1/4 of the exps on the FMA pipe, degree 4)
    kVarBwdProd = 1,      // bwd_chunk<NT-Xent, const shift> as shipped (+ tcgen05.st of W)
    kVarFwdAllMufu = 2,   // FFMA + MUFU + FADD + max tracking
    kVarFwdNoMax = 3,     // FFMA + MUFU + FADD
    kVarFwdHalfPoly = 4,  // every second exp on the FMA pipe
    kVarMufuOnly = 5,     // MUFU + FADD
    kVarFfmaOnly = 6,     // 4 FFMA per element, no MUFU
    kVarBwdAllMufu = 7,   // backward chunk, every exp on the MUFU pipe
    kVarLoadOnly = 8,     // tcgen05.ld + wait only
    kVarFwdPoly3 = 9,     // as shipped but degree-3 polynomial
    kNumChunkVariants = 10
};

template <int kVariant>
SIMCLR_DEVICE void probe_chunk(const Hot& h, const uint32_t (&r)[kChunk], FwdState& st, const BwdRow& br,
                               const RowCtx& rc, uint32_t cv_addr, uint32_t tmem_w) {
    if constexpr (kVariant == kVarFwdProd) {
        { float cm = kNegBig; fwd_chunk_fast<kNtXent, true>(h, r, cm, st); st.max_prec = fmaxf(st.max_prec, cm); }
    } else if constexpr (kVariant == kVarBwdProd || kVariant == kVarBwdAllMufu) {
        uint32_t w[kChunk / 2];
        if constexpr (kVariant == kVarBwdProd) {
            uint32_t unused[kChunk / 2];
            bwd_chunk<kNtXent, true, false>(h, r, cv_addr, 0, rc, br, w, unused);
        } else {
#pragma unroll
            for (int i = 0; i < kChunk; i += 4) {
                const float4 ac = lds_f4(cv_addr + i * 4);
                const float acs[4] = {ac.x, ac.y, ac.z, ac.w};
                float wv[4];
#pragma unroll
                for (int u = 0; u < 4; ++u)
                    wv[u] = ex2_approx(fmaf(__uint_as_float(r[i + u]), h.k2, -h.m2)) * (br.row_a + acs[u]);
                w[(i >> 1) + 0] = pack_bf16x2(wv[0], wv[1]);
                w[(i >> 1) + 1] = pack_bf16x2(wv[2], wv[3]);
            }
        }
        tmem_st8(tmem_w, w);
    } else if constexpr (kVariant == kVarLoadOnly) {
        st.sum += __uint_as_float(r[0] ^ r[kChunk - 1]);
    } else {
        float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
        if constexpr (kVariant == kVarFwdAllMufu) {
            float cm = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1]));
#pragma unroll
            for (int i = 2; i < kChunk; i += 2) cm = fmaxf(cm, fmaxf(__uint_as_float(r[i]), __uint_as_float(r[i + 1])));
            st.max_prec = fmaxf(st.max_prec, cm);
        }
#pragma unroll
        for (int i = 0; i < kChunk; i += 4) {
            const float v0 = __uint_as_float(r[i]), v1 = __uint_as_float(r[i + 1]), v2 = __uint_as_float(r[i + 2]),
                        v3 = __uint_as_float(r[i + 3]);
            if constexpr (kVariant == kVarFwdAllMufu || kVariant == kVarFwdNoMax) {
                a0 += ex2_approx(fmaf(v0, h.k2, -h.m2));
                a1 += ex2_approx(fmaf(v1, h.k2, -h.m2));
                a2 += ex2_approx(fmaf(v2, h.k2, -h.m2));
                a3 += ex2_approx(fmaf(v3, h.k2, -h.m2));
            } else if constexpr (kVariant == kVarFwdHalfPoly) {
                a0 += ex2_approx(fmaf(v0, h.k2, -h.m2));
                a1 += ex2_poly<4>(fmaf(v1, h.k2, -h.m2));
                a2 += ex2_approx(fmaf(v2, h.k2, -h.m2));
                a3 += ex2_poly<4>(fmaf(v3, h.k2, -h.m2));
            } else if constexpr (kVariant == kVarFwdPoly3) {
                a0 += ex2_approx(fmaf(v0, h.k2, -h.m2));
                a1 += ex2_approx(fmaf(v1, h.k2, -h.m2));
                a2 += ex2_approx(fmaf(v2, h.k2, -h.m2));
                a3 += ex2_poly<3>(fmaf(v3, h.k2, -h.m2));
            } else if constexpr (kVariant == kVarMufuOnly) {
                a0 += ex2_approx(v0);
                a1 += ex2_approx(v1);
                a2 += ex2_approx(v2);
                a3 += ex2_approx(v3);
            } else {   // kVarFfmaOnly
                a0 = fmaf(fmaf(fmaf(fmaf(v0, h.k2, a0), h.m2, v1), h.k2, v2), h.m2, a0);
                a1 = fmaf(fmaf(fmaf(fmaf(v1, h.k2, a1), h.m2, v2), h.k2, v3), h.m2, a1);
                a2 = fmaf(fmaf(fmaf(fmaf(v2, h.k2, a2), h.m2, v3), h.k2, v0), h.m2, a2);
                a3 = fmaf(fmaf(fmaf(fmaf(v3, h.k2, a3), h.m2, v0), h.k2, v1), h.m2, a3);
            }
        }
        st.sum += (a0 + a1) + (a2 + a3);
    }
}

// out[variant * 32 + warp] = cycles per 64-column half tile (four pipelined 16-column chunks, as in the tile kernel)
// seen by `warp` of CTA 0 (clock64 around its whole loop)
template <int kVariant>
__global__ void __launch_bounds__(kThreadsForward, 1)
chunk_rate_kernel(long long* __restrict__ out, int iters, int nwarps, float k2, float* __restrict__ sink) {
    __shared__ __align__(16) float cv[2 * kBlockN];
    __shared__ uint32_t tmem_ptr;
    const int warp = __shfl_sync(0xffffffffu, threadIdx.x >> 5, 0);
    const int lane = threadIdx.x & 31;
    if (threadIdx.x < 2 * kBlockN) cv[threadIdx.x] = 1e-3f * threadIdx.x;
    if (warp == kAllocWarp) {
        tmem_alloc(&tmem_ptr, kTmemCols);
        tmem_relinquish();
    }
    tc_fence_before_sync();
    __syncthreads();
    tc_fence_after_sync();
    const uint32_t tmem = tmem_ptr;
    Hot h;
    h.k2 = k2;
    h.m2 = k2 * kConstShiftRaw;
    h.qscale = 1.f;
    h.bg_pad = 1 << 20;
    h.b_glob = 1 << 20;
    h.const_shift = true;
    if (warp < nwarps && warp < kNumSoftmaxWarps) {
        const uint32_t lane_addr = static_cast<uint32_t>((warp & 3) * 32) << 16;
        const uint32_t col = (warp >> 2) * 128;                  // each warpgroup its own 128 columns
        // bounded "scores" in [-1, 1]
        for (int q = 0; q < 4; ++q) {
            uint32_t init[16];
#pragma unroll
            for (int i = 0; i < 16; ++i) init[i] = __float_as_uint(__sinf(0.37f * (lane + 1) * (i + 1 + 16 * q)));
            tmem_st16(tmem + lane_addr + col + q * 16, init);
        }
        tmem_st_wait();
        FwdState st;
        BwdRow br;
        br.row_a = 1e-3f * lane;
        br.row_l2 = 0.f;
        RowCtx rc;
        rc.diag_col = -1;
        rc.pos_col = -1;
        rc.row_ok = true;
        rc.vr = 0;
        rc.g = 0;
        const uint32_t cv_addr = smem_u32(cv);
        const uint32_t t0 = tmem + lane_addr + col;
        __syncwarp();
        const long long t_start = clock64();
        for (int it = 0; it < iters; ++it) {
            uint32_t ra[kChunk], rb[kChunk];
            tmem_ld16(t0, ra);
#pragma unroll 1
            for (int kk = 0; kk < 2; ++kk) {
                tmem_ld_wait16(ra);
                tmem_ld16(t0 + (2 * kk + 1) * kChunk, rb);
                probe_chunk<kVariant>(h, ra, st, br, rc, cv_addr, t0 + 64 + (2 * kk) * 8);
                tmem_ld_wait16(rb);
                if (kk == 0) tmem_ld16(t0 + 2 * kChunk, ra);
                probe_chunk<kVariant>(h, rb, st, br, rc, cv_addr, t0 + 64 + (2 * kk + 1) * 8);
            }
            if constexpr (kVariant == kVarBwdProd || kVariant == kVarBwdAllMufu) tmem_st_wait();
        }
        const long long t_end = clock64();
        if (blockIdx.x == 0 && lane == 0) out[kVariant * 32 + warp] = (t_end - t_start) / iters;
        if (st.sum + st.max_prec == 12345.678f) sink[threadIdx.x] = st.sum;
    }
    tc_fence_before_sync();
    __syncthreads();
    if (warp == kAllocWarp) {
        tc_fence_after_sync();
        tmem_dealloc(tmem, kTmemCols);
    }
}

// ---------------------------------------------------------------------------------------------
// Issue-rate probe for single instructions (tools/pipe_rate.py): `nwarps` warps per CTA run a loop of 16 independent
// chains of one opcode (pinned with inline PTX); out[op] = cycles for the whole loop seen by warp 0 of CTA 0.
//   0 FFMA 3-reg   1 FFMA reg,imm,reg   2 FADD reg,reg   3 FADD reg,imm   4 FMUL reg,reg   5 MUFU.EX2
//   6 FMNMX reg,reg   7 IMAD (shl-add form)   8 LEA-style shl+add (integer)   9 FFMA 3-reg + MUFU interleaved 3:1
//   10 F2FP (cvt.rn.bf16x2.f32)   11 FMNMX3 (max of three)   12 LDS.128   13 FSETP+FSEL   14 HFMA2.BF16 (fma.rn.bf16x2)
//   15 PRMT   16 FADD2   17 FFMA2   18 FADD2 : FMNMX 1:1   19 FFMA2 : MUFU 3:1   20 FADD2 : FADD 1:1
//   21 ex2.approx.ftz.f16x2   22 ex2.approx.ftz.bf16x2   23 cvt.rn.f16x2.f32 (F2FP.F16)
// ---------------------------------------------------------------------------------------------
constexpr int kPipeOps = 24;
constexpr int kPipeChains = 16;
constexpr int kPipeUnroll = 4;      // instructions per chain per loop iteration

template <int kOp>
__global__ void __launch_bounds__(kThreadsForward, 1)
pipe_rate_kernel(long long* __restrict__ out, int iters, int nwarps, float seed, float* __restrict__ sink) {
    __shared__ __align__(16) float lds_src[kPipeChains * 4 + 128];
    if (threadIdx.x < kPipeChains * 4 + 128) lds_src[threadIdx.x] = 1e-3f * threadIdx.x;
    __syncthreads();
    const uint32_t lds_addr = smem_u32(lds_src);
    const int warp = threadIdx.x >> 5;
    if (warp >= nwarps) return;
    float a[kPipeChains];
    int ia[kPipeChains];
#pragma unroll
    for (int c = 0; c < kPipeChains; ++c) {
        a[c] = seed * (threadIdx.x + c + 1) * 1e-3f;
        ia[c] = threadIdx.x + c;
    }
    const float b = seed * 0.999f, d = seed * 1e-3f;
    const int ib = static_cast<int>(seed) + 3;
    __syncwarp();
    const long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int u = 0; u < kPipeUnroll; ++u) {
#pragma unroll
            for (int c = 0; c < kPipeChains; ++c) {
                if constexpr (kOp == 0) asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(d));
                else if constexpr (kOp == 1) asm volatile("fma.rn.f32 %0, %0, 0f3F7FBE77, %1;" : "+f"(a[c]) : "f"(d));
                else if constexpr (kOp == 2) asm volatile("add.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));
                else if constexpr (kOp == 3) asm volatile("add.f32 %0, %0, 0f3A83126F;" : "+f"(a[c]));
                else if constexpr (kOp == 4) asm volatile("mul.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(b));
                else if constexpr (kOp == 5) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
                else if constexpr (kOp == 6) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));
                else if constexpr (kOp == 7) asm volatile("mad.lo.s32 %0, %0, 8388608, %1;" : "+r"(ia[c]) : "r"(ib));
                else if constexpr (kOp == 8) asm volatile("{ .reg .b32 t; shl.b32 t, %0, 23; add.s32 %0, t, %1; }" : "+r"(ia[c]) : "r"(ib));
                else if constexpr (kOp == 9) {
                    if ((c & 3) == 3) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
                    else asm volatile("fma.rn.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(b), "f"(d));
                } else if constexpr (kOp == 10) asm volatile("{ .reg .f32 t; mov.b32 t, %0; cvt.rn.bf16x2.f32 %0, t, %1; }" : "+r"(ia[c]) : "f"(b));
                else if constexpr (kOp == 11) asm volatile("max.f32 %0, %0, %1, %2;" : "+f"(a[c]) : "f"(d), "f"(b));
                else if constexpr (kOp == 12) {
                    float x0, x1, x2, x3;
                    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(x0), "=f"(x1), "=f"(x2), "=f"(x3) : "r"(lds_addr + c * 16 + ((__float_as_uint(a[c]) & 1u) << 4)));
                    a[c] = x0 + x3;
                } else if constexpr (kOp == 13) asm volatile("{ .reg .pred q; setp.gt.f32 q, %1, 0f00000000; selp.f32 %0, %0, %2, q; }" : "+f"(a[c]) : "f"(b), "f"(d));
                else if constexpr (kOp == 14) asm volatile("fma.rn.bf16x2 %0, %0, %1, %2;" : "+r"(ia[c]) : "r"(ib), "r"(ib));
                else if constexpr (kOp == 15) asm volatile("prmt.b32 %0, %0, %1, 0x5410;" : "+r"(ia[c]) : "r"(ib));
                else if constexpr (kOp == 21) asm volatile("ex2.approx.f16x2 %0, %0;" : "+r"(ia[c]));
                else if constexpr (kOp == 22) asm volatile("ex2.approx.ftz.bf16x2 %0, %0;" : "+r"(ia[c]));
                else if constexpr (kOp == 23) asm volatile("{ .reg .f32 t; mov.b32 t, %0; cvt.rn.f16x2.f32 %0, t, %1; }" : "+r"(ia[c]) : "f"(b));
                else if constexpr (kOp >= 16) {
                    // packed chains: chain c pairs a[c] with a[c ^ 8] (8 packed chains of two floats)
                    const bool packed_slot = kOp == 16 || kOp == 17 || ((kOp == 18 || kOp == 20) && (c & 1) == 0) || (kOp == 19 && (c & 3) != 3);
                    if (packed_slot) {
                        if (c < 8) {
                            f32x2 v = pack2(a[c], a[c + 8]);
                            const f32x2 dd = pack2(d, b);
                            if constexpr (kOp == 17 || kOp == 19) asm volatile("fma.rn.f32x2 %0, %0, %1, %1;" : "+l"(v) : "l"(dd));
                            else asm volatile("add.rn.f32x2 %0, %0, %1;" : "+l"(v) : "l"(dd));
                            unpack2(v, a[c], a[c + 8]);
                        }
                    } else if (c < 8) {
                        if constexpr (kOp == 18) asm volatile("max.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));
                        else if constexpr (kOp == 19) asm volatile("ex2.approx.ftz.f32 %0, %0;" : "+f"(a[c]));
                        else asm volatile("add.f32 %0, %0, %1;" : "+f"(a[c]) : "f"(d));
                    }
                }
            }
        }
    }
    const long long t1 = clock64();
    float acc = 0.f;
    int iacc = 0;
#pragma unroll
    for (int c = 0; c < kPipeChains; ++c) {
        acc += a[c];
        iacc += ia[c];
    }
    if (acc == 12345.678f || iacc == 123456789) sink[threadIdx.x] = acc;
    if (blockIdx.x == 0 && threadIdx.x == 0) out[kOp] = t1 - t0;
}

}  // namespace simclr
