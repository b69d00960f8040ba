// tc_common.cuh -- raw-PTX building blocks for the tcgen05 (5th-gen tensor core) kernels:
// mbarrier, TMEM allocation, UMMA shared-memory / instruction descriptors, tcgen05.mma /
// commit / ld / st, proxy fences, and the 128-byte-swizzle K-major operand layout.
// sm_100a only.  No CUTLASS: everything below is the PTX the hardware guide
// (/opt/skills/guides/blackwell_cuda_programming.md) and cute/arch/mma_sm100_desc.hpp describe.
#pragma once
#include <cuda_bf16.h>
#include "common.cuh"

namespace tc {

// Grid of a "persistent" kernel whose CTAs walk the tiles with stride gridDim.x (one CTA per SM
// is resident: shared memory).  With batches in flight on several streams (inference.py) part of
// the SMs can be held by another stream's latency-bound kernel (FPS, NMS) when this one starts; a
// grid of exactly one CTA per SM would then run the late CTAs alone after the early ones have
// finished.  Long kernels are therefore cut into up to 4 CTAs per SM (>= 32 tiles each): the hardware
// block scheduler hands the later CTAs to whichever SM frees up first, for ~1 % of prologue.
inline long long persistent_grid(long long tiles, int sms) {
    if (tiles <= sms) return tiles;
    long long per_sm = tiles / ((long long)sms * 32);
    if (per_sm < 1) per_sm = 1;
    if (per_sm > 4) per_sm = 4;
    return (long long)sms * per_sm;
}

// ------------------------------------------------------------------ mbarrier
__device__ __forceinline__ void mbar_init(uint64_t *bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(pn2_smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t *bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(pn2_smem_u32(bar)) : "memory");
}
// Wait on an mbarrier phase.  try_wait carries a suspend-time hint so the hardware parks the warp
// until the phase flips instead of spinning: a polling loop of 22 warps saturates the four
// schedulers of the SM and starves the warps that have real work (measured: 1.6 G of the 1.7 G
// warp-instructions of a run were polls before this).  Bounded: after ~30 s of waiting a protocol bug
// aborts the kernel (launch error on the next sync) rather than hanging the GPU (4 s was too tight under
// compute-sanitizer's racecheck, which slows a tile down by two orders of magnitude).
__device__ __forceinline__ void mbar_wait(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = pn2_smem_u32(bar);
    uint32_t done = 0;
    // NOT unrolled: ptxas otherwise replicates the try_wait / NANOSLEEP body 64x at every call site (13 k of the
    // fused SA kernel's 15 k instructions were unrolled polls) and every role pays instruction-cache misses
    // ("no_inst" stalls after each wait) for code that never runs.
#pragma unroll 1
    for (uint32_t spin = 0; spin < 32768u; ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"(1000000u)
            : "memory");
        if (done) return;
    }
    __trap();
}
// Wait of a role that runs AHEAD of its consumer (an operand producer or the weight stream waiting for a ring stage to
// drain): the wait is long by construction and nothing downstream is blocked by a slightly late wake-up.  mbar_wait()'s
// try_wait parks the warp until ANY mbarrier of the CTA changes state, then re-polls: with ~30 barriers ticking, the
// sixteen producer warps of the fused SA kernel executed 77 M polls (8 instructions each, 41 % of all issued instructions
// of the kernel, ncu source page) and competed for issue slots with the epilogue warps that bound the kernel.  Here the
// barrier is tested once and the warp sleeps a fixed ~100 ns between tests.
__device__ __forceinline__ void mbar_wait_lazy(uint64_t *bar, uint32_t parity) {
    const uint32_t addr = pn2_smem_u32(bar);
    uint32_t done = 0;
#pragma unroll 1
    for (uint32_t spin = 0; spin < (1u << 28); ++spin) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.test_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (done) return;
        __nanosleep(100);
    }
    __trap();
}
template <bool PROF>
__device__ __forceinline__ void mbar_wait_lazy_timed(uint64_t *bar, uint32_t parity, unsigned long long &acc) {
    if (!PROF) {
        mbar_wait_lazy(bar, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait_lazy(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}

// Optional in-kernel stopwatch (tools/prof_tc.py): cycles a role spends blocked on a barrier.
// PROF is a KERNEL template parameter: the product instantiation carries no stopwatch code at all.
template <bool PROF>
__device__ __forceinline__ void mbar_wait_timed(uint64_t *bar, uint32_t parity, unsigned long long &acc) {
    if (!PROF) {
        mbar_wait(bar, parity);
        return;
    }
    const long long t0 = clock64();
    mbar_wait(bar, parity);
    acc += (unsigned long long)(clock64() - t0);
}

__device__ __forceinline__ void fence_barrier_init() { asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory"); }
// generic-proxy writes to shared memory (st.shared) -> visible to the async proxy (tcgen05.mma operand reads)
__device__ __forceinline__ void fence_proxy_async_smem() { asm volatile("fence.proxy.async.shared::cta;" ::: "memory"); }

// named barrier among a subset of warps (id 1..15; 0 is __syncthreads)
__device__ __forceinline__ void named_bar_sync(int id, int nthreads) {
    asm volatile("bar.sync %0, %1;" ::"r"(id), "r"(nthreads) : "memory");
}

// ------------------------------------------------------------------ TMEM
__device__ __forceinline__ void tmem_alloc(uint32_t *dst_smem, uint32_t ncols) {  // whole warp, ncols pow2 >= 32
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(pn2_smem_u32(dst_smem)),
                 "r"(ncols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t ncols) {  // same warp that allocated
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(ncols) : "memory");
}
__device__ __forceinline__ void tc_fence_before_sync() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after_sync() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }

// ------------------------------------------------------------------ descriptors
// K-major operand tile, 128-byte swizzle: rows of 128 B (64 bf16 along K), 8-row groups of 1 KB,
// the 16-byte chunk c of row r is stored at chunk position c ^ (r & 7).  SBO = 1024 B between
// 8-row groups, LBO unused (one swizzle atom along K), descriptor version 1 (sm_100),
// layout type 2 = SWIZZLE_128B (cute::UMMA::SmemDescriptor).
__device__ __forceinline__ uint64_t make_smem_desc_sw128(uint32_t smem_addr) {
    uint64_t d = 0;
    d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);        // start address  [0,14)
    d |= (uint64_t)1 << 16;                              // LBO (ignored)  [16,30)
    d |= (uint64_t)(1024 >> 4) << 32;                    // SBO            [32,46)
    d |= (uint64_t)1 << 46;                              // version = 1    [46,48)
    d |= (uint64_t)2 << 61;                              // SWIZZLE_128B   [61,64)
    return d;
}
// byte offset of the 16-byte chunk `c` (0..7) of row `r` inside a [rows][64 bf16] swizzled tile
__device__ __forceinline__ uint32_t sw128_offset(int r, int c) {
    return (uint32_t)(((r >> 3) << 10) + ((r & 7) << 7) + (((c ^ r) & 7) << 4));
}

// kind::f16 instruction descriptor: D = f32, A = B = bf16, both K-major, M x N (cute::UMMA::InstrDescriptor)
__host__ __device__ constexpr uint32_t make_idesc_bf16(int m, int n) {
    return (1u << 4)                       // c_format  = F32
           | (1u << 7)                     // a_format  = BF16
           | (1u << 10)                    // b_format  = BF16
           | ((uint32_t)(n >> 3) << 17)    // n_dim
           | ((uint32_t)(m >> 4) << 24);   // m_dim
}

// The same descriptor as two 32-bit halves: the high word is a constant, advancing along K inside
// the swizzle atom is a 32-bit add on the low word.  The MMA-issuing thread shares its scheduler
// with six other warps, so every instruction it does not execute is tensor-pipe time gained.
constexpr uint32_t kDescHi = (1024u >> 4) | (1u << 14) | (2u << 29);   // SBO | version 1 | SWIZZLE_128B
__device__ __forceinline__ uint32_t desc_lo(uint32_t smem_addr) { return ((smem_addr & 0x3FFFFu) >> 4) | (1u << 16); }

__device__ __forceinline__ void mma_ss_lo(uint32_t d_tmem, uint32_t a_lo, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 da, db;\n"
        "mov.b64 da, {%1, %5};\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], da, db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_lo), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}
__device__ __forceinline__ void mma_ts_lo(uint32_t d_tmem, uint32_t a_tmem, uint32_t b_lo, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        ".reg .b64 db;\n"
        "mov.b64 db, {%2, %5};\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], db, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "r"(b_lo), "r"(idesc), "r"(accumulate), "r"(kDescHi)
        : "memory");
}

// One lane of a CONVERGED warp (elect.sync).  The MMA warp runs its tile / K-block loops with all 32
// lanes so that every address, descriptor and phase bit is warp-uniform and lives in uniform
// registers; only the tcgen05.mma / commit instructions sit under this predicate.  Wrapping the
// whole loop in `if (lane == 0)` instead makes the compiler treat all of it as divergent, and it then
// feeds every UTCHMMA through an ELECT + 4x R2UR.BROADCAST + BRA.U.ANY loop (11 instructions and a
// vector->uniform round trip per MMA, ~1000 instructions per 48-MMA tile).
__device__ __forceinline__ bool elect_one() {
    uint32_t pred;
    asm volatile(
        "{\n"
        ".reg .pred P;\n"
        "elect.sync _|P, 0xffffffff;\n"
        "selp.u32 %0, 1, 0, P;\n"
        "}\n"
        : "=r"(pred));
    return pred != 0;
}

// ------------------------------------------------------------------ MMA issue (one thread)
// D[tmem] (+)= A[smem] . B[smem]^T   (both operands K-major)
__device__ __forceinline__ void mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// D[tmem] (+)= A[tmem] . B[smem]^T   (A: lane = row, 32-bit column j holds k = 2j (low half), 2j+1)
__device__ __forceinline__ void mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n"
        ".reg .pred p;\n"
        "setp.ne.b32 p, %4, 0;\n"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n"
        "}\n" ::"r"(d_tmem),
        "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// all prior tcgen05.mma of this thread complete -> arrive on the mbarrier (implies fence::before_thread_sync)
__device__ __forceinline__ void mma_commit(uint64_t *bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(pn2_smem_u32(bar))
                 : "memory");
}

// ------------------------------------------------------------------ TMEM <-> registers
// 32 lanes x 32 bit, 16 consecutive columns: thread t of warp w reads lane 32*(w%4)+t
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&v)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 {%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15}, [%16];"
        : "=r"(v[0]), "=r"(v[1]), "=r"(v[2]), "=r"(v[3]), "=r"(v[4]), "=r"(v[5]), "=r"(v[6]), "=r"(v[7]), "=r"(v[8]),
          "=r"(v[9]), "=r"(v[10]), "=r"(v[11]), "=r"(v[12]), "=r"(v[13]), "=r"(v[14]), "=r"(v[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tmem_st8(uint32_t taddr, const uint32_t (&v)[8]) {
    asm volatile("tcgen05.st.sync.aligned.32x32b.x8.b32 [%0], {%1, %2, %3, %4, %5, %6, %7, %8};" ::"r"(taddr), "r"(v[0]),
                 "r"(v[1]), "r"(v[2]), "r"(v[3]), "r"(v[4]), "r"(v[5]), "r"(v[6]), "r"(v[7])
                 : "memory");
}
__device__ __forceinline__ void tmem_st_wait() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }

// ------------------------------------------------------------------ bf16 hi/lo split
// x = hi + lo + O(2^-17 |x|): hi = bf16(x), lo = bf16(x - hi).  Packs two values per 32-bit
// word with the FIRST (lower k) value in the low half, the order the MMA reads K in.
__device__ __forceinline__ void split2(float x0, float x1, uint32_t &hi, uint32_t &lo) {
    const __nv_bfloat162 h = __floats2bfloat162_rn(x0, x1);
    const float2 hf = __bfloat1622float2(h);
    const __nv_bfloat162 l = __floats2bfloat162_rn(x0 - hf.x, x1 - hf.y);
    hi = *reinterpret_cast<const uint32_t *>(&h);
    lo = *reinterpret_cast<const uint32_t *>(&l);
}

// Four values at once for the producers (the SM is instruction-issue bound in the fused kernels, so
// the split is written for instruction count): one F2FP per pair for hi, the bf16 -> fp32 expansion
// as one shift / one mask instead of PRMT + shift per element, x - hi on the packed fp32 pipe
// (FADD2, both halves IEEE round-to-nearest, i.e. the same bits as the scalar subtraction), one
// F2FP per pair for lo.  10 instructions per float4 instead of 16; identical results to split2().
__device__ __forceinline__ void split4(const float4 &v, uint2 &hi, uint2 &lo) {
    const __nv_bfloat162 h01 = __floats2bfloat162_rn(v.x, v.y), h23 = __floats2bfloat162_rn(v.z, v.w);
    hi.x = *reinterpret_cast<const uint32_t *>(&h01);
    hi.y = *reinterpret_cast<const uint32_t *>(&h23);
    const float2 n01 = make_float2(-__uint_as_float(hi.x << 16), -__uint_as_float(hi.x & 0xffff0000u));
    const float2 n23 = make_float2(-__uint_as_float(hi.y << 16), -__uint_as_float(hi.y & 0xffff0000u));
    const float2 l01 = __fadd2_rn(make_float2(v.x, v.y), n01);
    const float2 l23 = __fadd2_rn(make_float2(v.z, v.w), n23);
    const __nv_bfloat162 q01 = __floats2bfloat162_rn(l01.x, l01.y), q23 = __floats2bfloat162_rn(l23.x, l23.y);
    lo.x = *reinterpret_cast<const uint32_t *>(&q01);
    lo.y = *reinterpret_cast<const uint32_t *>(&q23);
}

}  // namespace tc
