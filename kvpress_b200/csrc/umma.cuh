// umma.cuh — thin inline-PTX layer for Blackwell (sm_100a) tensor cores, TMEM, TMA and mbarriers.
// Written for this repo's two dense scorers (SnapKV QK^T, ExpectedAttention K·Sigma): bf16/fp16
// operands, fp32 accumulators in TMEM, both operands K-major in shared memory with the 128-byte
// swizzle that TMA (cp.async.bulk.tensor, SWIZZLE_128B) produces, cta_group::1.
#pragma once
#include <cuda.h>  // CUtensorMap (types only; the encode entry point is fetched at run time)
#include <cuda_runtime.h>
#include <stdint.h>

namespace kvp {
namespace umma {

// ---- shared-memory addresses / mbarriers ----------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
    return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_fence_init() {
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint64_t* bar, uint32_t bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)),
                 "r"(bytes)
                 : "memory");
}
// Bounded wait: a barrier that never flips (a bug) traps instead of hanging the GPU.
// try_wait carries a suspend-time hint: ptxas turns it into TRYWAIT + NANOSLEEP.SYNCS, i.e. the warp SLEEPS until the
// barrier's phase completes (or the hint expires) instead of re-issuing the poll. Without the hint the waiting roles
// (TMA producer, MMA issuer, an epilogue warpgroup ahead of its accumulator) spin through ~5 instructions per poll
// and take issue slots from the epilogue warps that share their scheduler: round 1's ncu capture of snap_stats_kernel
// shows 26.2 M executed warp instructions for ~10.5 M of useful epilogue work (profiles/r01_snapkv32k_ncu_full_details.txt).
#ifndef KVP_MBAR_HINT_NS
#define KVP_MBAR_HINT_NS 20000  // 0: plain polling (A/B)
#endif
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
    const uint32_t addr = smem_u32(bar);
    uint32_t done = 0;
    for (uint32_t spins = 0; !done; ++spins) {
#if KVP_MBAR_HINT_NS > 0
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2, %3;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity), "r"((uint32_t)KVP_MBAR_HINT_NS)  // sleep up to this long per poll, woken by the barrier
            : "memory");
        if (spins > (1u << 18)) __trap();  // > 5 s without progress
#else
        asm volatile(
            "{\n\t.reg .pred p;\n\t"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
            "selp.u32 %0, 1, 0, p;\n\t}"
            : "=r"(done)
            : "r"(addr), "r"(parity)
            : "memory");
        if (spins > (1u << 26)) __trap();
#endif
    }
}

// ---- TMA tensor-map loads (tile mode, 128B swizzle) ---------------------------------------------
__device__ __forceinline__ void prefetch_tmap(const CUtensorMap* m) {
    asm volatile("prefetch.tensormap [%0];" ::"l"(reinterpret_cast<uint64_t>(m)) : "memory");
}
__device__ __forceinline__ void tma_load_3d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tma_load_4d(void* dst, const CUtensorMap* m, uint64_t* bar, int c0,
                                            int c1, int c2, int c3) {
    asm volatile(
        "cp.async.bulk.tensor.4d.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5, %6}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar)), "r"(c0), "r"(c1), "r"(c2), "r"(c3)
        : "memory");
}

// ---- TMEM ----------------------------------------------------------------------------------------
// One full warp allocates `cols` (power of two, 32..512) columns; the base address lands in *dst.
__device__ __forceinline__ void tmem_alloc(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(
                     smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols)
                 : "memory");
}
__device__ __forceinline__ void fence_before_sync() {
    asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory");
}
__device__ __forceinline__ void fence_after_sync() {
    asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory");
}
// 32 lanes x 32 consecutive fp32 columns: thread t of the warp gets row (lane base + t).
__device__ __forceinline__ void tmem_ld32(uint32_t taddr, uint32_t (&r)[32]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15,"
        "%16,%17,%18,%19,%20,%21,%22,%23,%24,%25,%26,%27,%28,%29,%30,%31}, [%32];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15]), "=r"(r[16]), "=r"(r[17]), "=r"(r[18]), "=r"(r[19]),
          "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]), "=r"(r[25]),
          "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld16(uint32_t taddr, uint32_t (&r)[16]) {
    asm volatile(
        "tcgen05.ld.sync.aligned.32x32b.x16.b32 "
        "{%0,%1,%2,%3,%4,%5,%6,%7,%8,%9,%10,%11,%12,%13,%14,%15}, [%16];"
        : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]),
          "=r"(r[7]), "=r"(r[8]), "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]),
          "=r"(r[14]), "=r"(r[15])
        : "r"(taddr)
        : "memory");
}
__device__ __forceinline__ void tmem_ld_wait() {
    asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory");
}

// ---- MMA descriptors -----------------------------------------------------------------------------
// Shared-memory matrix descriptor of a K-major operand panel [rows x 64 elements(128 B)] written by
// TMA with SWIZZLE_128B: 8-row groups are 1024 B apart (SBO), the panel base is 1024-B aligned.
// Bit layout (cute::UMMA::SmemDescriptor): [0,14) addr>>4, [16,30) LBO>>4, [32,46) SBO>>4,
// [46,48) version=1, [61,64) layout type (2 = SWIZZLE_128B).
__device__ __forceinline__ uint64_t smem_desc_sw128(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)1 << 16;             // LBO (ignored for swizzled K-major)
    d |= (uint64_t)(1024 >> 4) << 32;   // SBO
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    d |= (uint64_t)2 << 61;             // SWIZZLE_128B
    return d;
}
// Descriptor of a K-major operand that is only 16 elements (32 B) wide, stored WITHOUT swizzle in the
// canonical "interleaved" form: core matrix = 8 rows x 16 B (rows 16 B apart, 128 B total); the
// second 16-byte K half of the same 8 rows sits LBO = 128 B further, the next 8-row group SBO = 256 B
// further. Row r, half j lives at (r/8)*256 + j*128 + (r%8)*16.
__device__ __forceinline__ uint64_t smem_desc_k16_noswizzle(uint32_t saddr) {
    uint64_t d = 0;
    d |= (uint64_t)((saddr >> 4) & 0x3FFFu);
    d |= (uint64_t)(128 >> 4) << 16;    // LBO
    d |= (uint64_t)(256 >> 4) << 32;    // SBO
    d |= (uint64_t)1 << 46;             // descriptor version (Blackwell)
    return d;                           // layout type 0 = SWIZZLE_NONE
}
__device__ __forceinline__ uint32_t k16_noswizzle_offset(int row, int half) {
    return (uint32_t)((row >> 3) * 256 + half * 128 + (row & 7) * 16);
}
// Instruction descriptor for kind::f16 (cute::UMMA::InstrDescriptor): fp32 accumulate, A and B
// K-major, `ab_format` 0 = fp16, 1 = bf16.
__host__ __device__ constexpr uint32_t instr_desc_f16(int M, int N, int ab_format) {
    return (1u << 4) | ((uint32_t)ab_format << 7) | ((uint32_t)ab_format << 10) |
           ((uint32_t)(N >> 3) << 17) | ((uint32_t)(M >> 4) << 24);
}
// D[tmem] (+)= A[smem] * B[smem]^T ; issued by ONE thread.
__device__ __forceinline__ void mma_f16_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc,
                                           uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the mbarrier when all previously issued MMAs of this thread have completed
// (implies tcgen05.fence::before_thread_sync).
__device__ __forceinline__ void mma_commit(uint64_t* bar) {
    asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(
                     smem_u32(bar))
                 : "memory");
}


// ---- CTA pairs (cta_group::2): one MMA of M = 256 over the two SMs of a cluster of two ---------------------------
// Each CTA supplies its own 128 rows of A and its own N/2 rows of B from ITS shared memory (same offsets in both
// CTAs) and receives its own 128 rows x N columns of D in ITS tensor memory: per CTA the operand traffic of an MMA
// drops from (128 + N) x 32 B to (128 + N/2) x 32 B. Only the leader (cluster rank 0) issues MMAs and commits;
// barriers the leader waits on receive their arrivals / transaction bytes from both CTAs.
constexpr uint32_t kPeerBitMask = 0xFEFFFFFFu;  // clears the CTA-rank bit of a shared::cluster address: "same offset in CTA 0"
__device__ __forceinline__ uint32_t cluster_ctarank() {
    uint32_t r;
    asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
    return r;
}
__device__ __forceinline__ void cluster_sync() {
    asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// arrive on the barrier at the same offset in the LEADER CTA's shared memory (from either CTA of the pair)
__device__ __forceinline__ void mbar_arrive_leader(uint64_t* bar) {
    asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(smem_u32(bar) & kPeerBitMask) : "memory");
}
// TMA load whose transaction bytes are counted on the LEADER's barrier (issued by both CTAs for their own share of an
// operand that only the tensor core reads; tiles that the CTA's own threads read complete on the CTA's own barrier)
__device__ __forceinline__ void tma_load_3d_pair(void* dst, const CUtensorMap* m, uint64_t* bar, int c0, int c1,
                                                 int c2) {
    asm volatile(
        "cp.async.bulk.tensor.3d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes"
        " [%0], [%1, {%3, %4, %5}], [%2];" ::"r"(smem_u32(dst)),
        "l"(reinterpret_cast<uint64_t>(m)), "r"(smem_u32(bar) & kPeerBitMask), "r"(c0), "r"(c1), "r"(c2)
        : "memory");
}
__device__ __forceinline__ void tmem_alloc_pair(uint32_t* dst_smem, uint32_t cols) {
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(dst_smem)),
                 "r"(cols)
                 : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tmem_dealloc_pair(uint32_t taddr, uint32_t cols) {
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
// D[tmem of both CTAs] (+)= A[256 rows over both CTAs] * B[N rows over both CTAs]^T ; issued by ONE thread of the leader
__device__ __forceinline__ void mma_f16_ss_pair(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                                uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
        "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
        : "memory");
}
// Arrives on the barrier at this offset in BOTH CTAs when all MMAs issued so far by this thread have completed
__device__ __forceinline__ void mma_commit_pair(uint64_t* bar) {
    const uint16_t mask = 3;
    asm volatile(
        "tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(
            smem_u32(bar)),
        "h"(mask)
        : "memory");
}

}  // namespace umma

// ---- packed fp32x2 arithmetic (sm_100: one FFMA2 issues two fp32 FMAs) -----------------------------
__device__ __forceinline__ uint64_t pack_u32x2(uint32_t lo, uint32_t hi) {
    uint64_t r;
    asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "r"(lo), "r"(hi));
    return r;
}
__device__ __forceinline__ uint64_t pack_f32x2(float2 v) {
    return pack_u32x2(__float_as_uint(v.x), __float_as_uint(v.y));
}
__device__ __forceinline__ float2 unpack_f32x2(uint64_t v) {
    uint32_t lo, hi;
    asm("mov.b64 {%0, %1}, %2;" : "=r"(lo), "=r"(hi) : "l"(v));
    return make_float2(__uint_as_float(lo), __uint_as_float(hi));
}
__device__ __forceinline__ uint64_t fma_f32x2(uint64_t a, uint64_t b, uint64_t c) {
    uint64_t d;
    asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(d) : "l"(a), "l"(b), "l"(c));
    return d;
}

namespace umma {
// Byte offset of 16-byte chunk `chunk` (0..7) of row `row` inside a SWIZZLE_128B panel whose rows
// are 128 B: what generic loads must use to read a TMA-written panel.
__device__ __forceinline__ uint32_t sw128_offset(int row, int chunk) {
    return (uint32_t)(row * 128 + ((chunk ^ (row & 7)) << 4));
}

}  // namespace umma

// ---- host: tensor-map encoding through the driver entry point (no -lcuda link dependency) ------
typedef CUresult (*kvp_encode_tiled_fn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*,
                                        const cuuint64_t*, const cuuint64_t*, const cuuint32_t*,
                                        const cuuint32_t*, CUtensorMapInterleave,
                                        CUtensorMapSwizzle, CUtensorMapL2promotion,
                                        CUtensorMapFloatOOBfill);
kvp_encode_tiled_fn get_encode_tiled();
// rank-N (3 or 4) map over 16-bit elements, innermost box of 64 elements (128 B), SWIZZLE_128B,
// out-of-bounds elements read as zero. dims/strides innermost-first; strides[0] is implicit.
cudaError_t make_tmap_16bit(CUtensorMap* map, const void* base, int rank, const uint64_t* dims,
                            const uint64_t* strides_bytes, const uint32_t* box);

}  // namespace kvp
