// knorm_chunk.cuh — scoring of one 256-position chunk of one (b, h) row for KnormPress, shared by the
// standalone score kernel (knorm.cu) and the fused persistent Knorm kernel (select_compact.cu).
#pragma once
#include "common.cuh"

namespace kvp {

constexpr int kScoreChunk = 256;  // positions per score work item (== kTile)
static_assert(kScoreChunk == kTile, "score chunks and select tiles share the key layout");

// Writes -||k_s||_2 (rounded once to the storage dtype) and its ordered key for the 256 positions of
// `chunk` into shared memory. 256 threads; a sub-warp of LPR lanes per 2*D-byte row, U independent
// 128-bit loads in flight per lane. Caller synchronises before reading skeys / sscores.
// kHint: L2 policy of the loads. 1 = evict_last (rows that the compaction stage re-reads soon: they should survive
// in the 126 MB L2), 2 = evict_first (rows that will be long gone by then: they should not push the others out),
// 0 = no hint.
template <typename T, int LPR, int kHint = 0>
__device__ __forceinline__ void knorm_score_chunk(const T* __restrict__ K, Strides3 ks, int b, int h,
                                                  int chunk, int S, int D, uint16_t* skeys,
                                                  uint16_t* sscores) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int RPW = 32 / LPR;                                   // rows per warp-wide load
    constexpr int TOK_PER_WARP = kScoreChunk / (kTileThreads / 32);  // 32
    constexpr int ITERS = TOK_PER_WARP / RPW;
    constexpr int U = (ITERS < 8) ? ITERS : 8;  // independent 16-byte loads in flight per lane
    static_assert(ITERS % U == 0, "unroll must divide the iteration count");

    const int sub = lane % LPR;   // which 16-byte piece of the row
    const int rsel = lane / LPR;  // which row of the RPW rows
    const int nvec = D >> 3;      // 16-byte pieces per row
    const T* base = K + (int64_t)b * ks.b + (int64_t)h * ks.h + (int64_t)sub * 8;
    const int s_warp = chunk * kScoreChunk + warp * TOK_PER_WARP;
    const uint64_t pol = kHint == 1 ? l2_policy_evict_last() : (kHint == 2 ? l2_policy_evict_first() : 0ull);

#pragma unroll 1
    for (int it = 0; it < ITERS; it += U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s_warp + (it + u) * RPW + rsel;
            v[u] = make_int4(0, 0, 0, 0);
            if (s < S && sub < nvec)
                v[u] = kHint != 0 ? ldg_hint(base + (int64_t)s * ks.s, pol) : ldg_plain(base + (int64_t)s * ks.s);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z,
                                   (uint32_t)v[u].w};
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = F16Traits<T>::unpack2(w[j]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
#pragma unroll
            for (int off = LPR / 2; off >= 1; off >>= 1)
                ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
            if (sub == 0) {
                const int sl = warp * TOK_PER_WARP + (it + u) * RPW + rsel;
                // -sqrt(ss) rounded once to the storage dtype (negation is exact)
                const uint16_t bits = F16Traits<T>::from_float(sqrtf(ss)) ^ 0x8000u;
                sscores[sl] = bits;
                skeys[sl] = ordered_key16(bits, F16Traits<T>::kInfBits);
            }
        }
    }
}

// ||x_s||_2 in fp32 for the 256 positions of `chunk` of one (b, h) row into shared memory (same access
// pattern as knorm_score_chunk; loads are L2-evict-first: V is read exactly once).
template <typename T, int LPR>
__device__ __forceinline__ void row_norm_chunk(const T* __restrict__ X, Strides3 xs, int b, int h, int chunk,
                                               int S, int D, float* snorm) {
    const int tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;
    constexpr int RPW = 32 / LPR;
    constexpr int TOK_PER_WARP = kScoreChunk / (kTileThreads / 32);
    constexpr int ITERS = TOK_PER_WARP / RPW;
    constexpr int U = (ITERS < 8) ? ITERS : 8;
    const int sub = lane % LPR, rsel = lane / LPR;
    const int nvec = D >> 3;
    const T* base = X + (int64_t)b * xs.b + (int64_t)h * xs.h + (int64_t)sub * 8;
    const int s_warp = chunk * kScoreChunk + warp * TOK_PER_WARP;
    const uint64_t pol = l2_policy_evict_first();
#pragma unroll 1
    for (int it = 0; it < ITERS; it += U) {
        int4 v[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int s = s_warp + (it + u) * RPW + rsel;
            v[u] = make_int4(0, 0, 0, 0);
            if (s < S && sub < nvec) v[u] = ldg_hint(base + (int64_t)s * xs.s, pol);
        }
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
            float ss = 0.f;
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = F16Traits<T>::unpack2(w[j]);
                ss = fmaf(f.x, f.x, ss);
                ss = fmaf(f.y, f.y, ss);
            }
#pragma unroll
            for (int off = LPR / 2; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
            if (sub == 0) snorm[warp * TOK_PER_WARP + (it + u) * RPW + rsel] = sqrtf(ss);
        }
    }
}

}  // namespace kvp
