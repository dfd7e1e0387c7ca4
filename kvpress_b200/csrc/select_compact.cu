// select_compact.cu — stages T (per-row top-k threshold) and G (stable compaction of K and V).
//
// Replaces scores.topk(n_kept).indices + keys.gather + values.gather
// (kvpress/presses/scorer_press.py:95-100) for 16-bit scores:
//   * the score stage left, per (b,h) row, the ordered keys and a 256-bin histogram of key>>8;
//   * refine_kernel  : every tile finds the threshold bin b1 from that histogram, histograms the
//                      low byte of the keys that fall in b1 (-> row hist_lo) and records, per
//                      tile, the suffix counts of those low bytes + the count of keys above b1;
//   * compact_kernel : every tile derives the exact 16-bit threshold T and the number of ties to
//                      take, sums the records of the tiles before it (no ordering constraint
//                      between CTAs), ranks its own 1024 positions with one block scan and copies
//                      the kept K and V rows (16-byte vectors, 8 in flight per thread) to
//                      [B,H,n_kept,D] in ascending position order. Ties at T go to the lowest
//                      positions.
// Tiles are visited in REVERSE order of the score stage so the K rows touched last (still in the
// 126 MB L2) are re-read first.
#include "common.cuh"

namespace kvp {

__global__ void __launch_bounds__(kTileThreads)
refine_kernel(int S, int n_kept, Workspace ws) {
    __shared__ uint32_t shist[256];
    __shared__ uint32_t slo[260];
    __shared__ uint32_t swarp[8];
    __shared__ int s_b1;

    const int tile = blockIdx.x, row = blockIdx.y, tid = threadIdx.x;
    const int warp = tid >> 5, lane = tid & 31;

    shist[tid] = ws.hist_hi[(size_t)row * 256 + tid];
    slo[tid] = 0;
    if (tid < 4) slo[256 + tid] = 0;
    // issue the key load before the barrier so it overlaps the histogram read
    const int s0 = tile * kTile + tid * 4;
    const uint2 packed =
        *reinterpret_cast<const uint2*>(ws.keys + (size_t)row * ws.S_pad + s0);
    __syncthreads();
    if (warp == 0) {
        int b1;
        uint32_t above;
        warp_suffix_find(shist, (uint32_t)n_kept, lane, b1, above);
        if (lane == 0) s_b1 = b1;
    }
    __syncthreads();
    const unsigned b1 = (unsigned)s_b1;

    const uint16_t k[4] = {(uint16_t)(packed.x & 0xFFFF), (uint16_t)(packed.x >> 16),
                           (uint16_t)(packed.y & 0xFFFF), (uint16_t)(packed.y >> 16)};
    uint32_t n_gt = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool valid = (s0 + i) < S;
        const unsigned hi = k[i] >> 8;
        n_gt += (valid && hi > b1) ? 1u : 0u;
        const unsigned bin = (valid && hi == b1) ? (unsigned)(k[i] & 0xFF) : 256u;
        const unsigned peers = __match_any_sync(0xFFFFFFFFu, bin);
        if (bin < 256u && lane == (__ffs(peers) - 1)) atomicAdd(&slo[bin], __popc(peers));
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) n_gt += __shfl_xor_sync(0xFFFFFFFFu, n_gt, off);
    if (lane == 0) swarp[warp] = n_gt;
    __syncthreads();

    // row-level low-byte histogram
    const uint32_t c = slo[tid];
    if (c) atomicAdd(&ws.hist_lo[(size_t)row * 256 + tid], c);

    // per-tile suffix sums sfx[j] = #candidates with low byte >= j  (block-wide suffix scan)
    uint32_t v = c;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_down_sync(0xFFFFFFFFu, v, off);
        if (lane + off < 32) v += t;
    }
    __shared__ uint32_t swsum[8];
    if (lane == 0) swsum[warp] = v;  // total of this warp's 32 bins
    __syncthreads();
    uint32_t tail = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) tail += (w > warp) ? swsum[w] : 0u;
    uint16_t* rec = ws.tile_sfx + ((size_t)row * ws.n_tiles + tile) * kSfxStride;
    rec[tid] = (uint16_t)(v + tail);
    if (tid == 0) {
        uint32_t g = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) g += swarp[w];
        rec[256] = 0;
        rec[257] = (uint16_t)g;
    }
}

// Copies `count` rows of `nvec` 16-byte vectors: dst row j <- src row list[j].
template <bool kKeepInL2>
__device__ __forceinline__ void copy_rows(const char* __restrict__ src, int64_t src_row_bytes,
                                          char* __restrict__ dst, int64_t dst_row_bytes,
                                          const int* __restrict__ list, int count, int nvec) {
    constexpr int U = 4;
    const int total = count * nvec;
    const uint64_t pol_first = l2_policy_evict_first();
    for (int base = threadIdx.x; base < total; base += kTileThreads * U) {
        int4 v[U];
        int rr[U], cc[U];
#pragma unroll
        for (int u = 0; u < U; ++u) {
            const int i = base + u * kTileThreads;
            rr[u] = -1;
            if (i < total) {
                rr[u] = i / nvec;
                cc[u] = i - rr[u] * nvec;
                const char* p = src + (int64_t)list[rr[u]] * src_row_bytes + cc[u] * 16;
                v[u] = kKeepInL2 ? ldg_plain(p) : ldg_hint(p, pol_first);
            }
        }
#pragma unroll
        for (int u = 0; u < U; ++u)
            if (rr[u] >= 0)
                stg_hint(dst + (int64_t)rr[u] * dst_row_bytes + cc[u] * 16, v[u], pol_first);
    }
}

__global__ void __launch_bounds__(kTileThreads)
compact_kernel(const char* __restrict__ K, const char* __restrict__ V, Strides3 ks, Strides3 vs,
               char* __restrict__ K_out, char* __restrict__ V_out, int32_t* __restrict__ idx_out,
               int H, int S, int D, int n_kept, Workspace ws) {
    __shared__ uint32_t shist[256];
    __shared__ uint32_t slo[256];
    __shared__ int s_list[kTile];
    __shared__ uint32_t s_scan[8];
    __shared__ uint32_t s_red[2][8];
    __shared__ uint32_t s_thr[3];  // T, n_take, lo1

    // reverse visiting order: last rows / last tiles of the score stage first
    const int tile = (int)gridDim.x - 1 - (int)blockIdx.x;
    const int row = (int)gridDim.y - 1 - (int)blockIdx.y;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = row / H, h = row % H;

    shist[tid] = ws.hist_hi[(size_t)row * 256 + tid];
    slo[tid] = ws.hist_lo[(size_t)row * 256 + tid];
    const int s0 = tile * kTile + tid * 4;
    const uint2 packed =
        *reinterpret_cast<const uint2*>(ws.keys + (size_t)row * ws.S_pad + s0);
    __syncthreads();
    if (warp == 0) {
        int b1, lo1;
        uint32_t above1, above2;
        warp_suffix_find(shist, (uint32_t)n_kept, lane, b1, above1);
        const uint32_t need1 = (uint32_t)n_kept - above1;  // >= 1
        warp_suffix_find(slo, need1, lane, lo1, above2);
        if (lane == 0) {
            s_thr[0] = ((uint32_t)b1 << 8) | (uint32_t)lo1;
            s_thr[1] = need1 - above2;  // ties (key == T) to take, >= 1
            s_thr[2] = (uint32_t)lo1;
        }
    }
    __syncthreads();
    const uint32_t T = s_thr[0], n_take = s_thr[1], lo1 = s_thr[2];

    // ---- kept / tie counts of the tiles before this one -----------------------------------
    uint32_t gt_before = 0, eq_before = 0;
    {
        const uint16_t* recs = ws.tile_sfx + (size_t)row * ws.n_tiles * kSfxStride;
        for (int t = tid; t < tile; t += kTileThreads) {
            const uint16_t* r = recs + (size_t)t * kSfxStride;
            const uint32_t ge = r[lo1], gt = r[lo1 + 1];
            gt_before += (uint32_t)r[257] + gt;
            eq_before += ge - gt;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            gt_before += __shfl_xor_sync(0xFFFFFFFFu, gt_before, off);
            eq_before += __shfl_xor_sync(0xFFFFFFFFu, eq_before, off);
        }
        if (lane == 0) {
            s_red[0][warp] = gt_before;
            s_red[1][warp] = eq_before;
        }
    }

    // ---- rank this tile's positions: packed scan, low 16 bits = #gt, high 16 bits = #eq --------
    const uint16_t k[4] = {(uint16_t)(packed.x & 0xFFFF), (uint16_t)(packed.x >> 16),
                           (uint16_t)(packed.y & 0xFFFF), (uint16_t)(packed.y >> 16)};
    uint32_t f[4];
    uint32_t tsum = 0;
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const bool valid = (s0 + i) < S;
        f[i] = valid ? ((k[i] > T) ? 1u : ((k[i] == T) ? 0x10000u : 0u)) : 0u;
        tsum += f[i];
    }
    uint32_t incl = tsum;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
        const uint32_t t = __shfl_up_sync(0xFFFFFFFFu, incl, off);
        if (lane >= off) incl += t;
    }
    if (lane == 31) s_scan[warp] = incl;
    __syncthreads();
    uint32_t wprefix = 0, total = 0;
    gt_before = 0;
    eq_before = 0;
#pragma unroll
    for (int w = 0; w < 8; ++w) {
        wprefix += (w < warp) ? s_scan[w] : 0u;
        total += s_scan[w];
        gt_before += s_red[0][w];
        eq_before += s_red[1][w];
    }
    // ties are taken in position order: this tile may take those with global tie rank < n_take
    const uint32_t tie_room = (n_take > eq_before) ? (n_take - eq_before) : 0u;
    const uint32_t out_base = gt_before + min(eq_before, n_take);
    uint32_t run = wprefix + incl - tsum;  // exclusive prefix of this thread
#pragma unroll
    for (int i = 0; i < 4; ++i) {
        const uint32_t gt_rank = run & 0xFFFFu, eq_rank = run >> 16;
        const bool keep = (f[i] == 1u) || (f[i] == 0x10000u && eq_rank < tie_room);
        if (keep) s_list[gt_rank + min(eq_rank, tie_room)] = s0 + i;
        run += f[i];
    }
    const int count = (int)((total & 0xFFFFu) + min(total >> 16, tie_room));
    __syncthreads();
    if (count == 0) return;

    const int64_t out_row0 = (int64_t)row * n_kept + out_base;
    if (idx_out != nullptr)
        for (int j = tid; j < count; j += kTileThreads) idx_out[out_row0 + j] = s_list[j];

    const int nvec = D >> 3;
    const int64_t row_bytes = (int64_t)D * 2;
    copy_rows<true>(K + ((int64_t)b * ks.b + (int64_t)h * ks.h) * 2, ks.s * 2,
                    K_out + out_row0 * row_bytes, row_bytes, s_list, count, nvec);
    copy_rows<false>(V + ((int64_t)b * vs.b + (int64_t)h * vs.h) * 2, vs.s * 2,
                     V_out + out_row0 * row_bytes, row_bytes, s_list, count, nvec);
}

cudaError_t launch_select_compact(const Dims& d, const void* K, const void* V, void* K_out,
                                  void* V_out, int32_t* idx_out, const Workspace& ws,
                                  cudaStream_t st) {
    dim3 grid(ws.n_tiles, d.R);
    refine_kernel<<<grid, kTileThreads, 0, st>>>(d.S, d.n_kept, ws);
    cudaError_t e = cudaPeekAtLastError();
    if (e != cudaSuccess) return e;
    compact_kernel<<<grid, kTileThreads, 0, st>>>(
        static_cast<const char*>(K), static_cast<const char*>(V), d.ks, d.vs,
        static_cast<char*>(K_out), static_cast<char*>(V_out), idx_out, d.H, d.S, d.D, d.n_kept,
        ws);
    return cudaPeekAtLastError();
}

// ---- StreamingLLM: the answer is two ranges, no scores needed --------------------------------
// kvpress/presses/streaming_llm_press.py:50-52 + scorer_press.py:95: ones everywhere except zeros
// on [n_sink, n_sink + n_pruned) => kept = [0, head) U [S - (n_kept - head), S),
// head = min(n_sink, n_kept) (lowest-position tie rule when n_kept < n_sink).
__global__ void __launch_bounds__(kTileThreads)
streaming_compact_kernel(const char* __restrict__ K, const char* __restrict__ V, Strides3 ks,
                         Strides3 vs, char* __restrict__ K_out, char* __restrict__ V_out,
                         int32_t* __restrict__ idx_out, int H, int S, int D, int n_kept,
                         int head) {
    __shared__ int s_list[kTile];
    const int tile = blockIdx.x, row = blockIdx.y, tid = threadIdx.x;
    const int b = row / H, h = row % H;
    const int j0 = tile * kTile;
    const int count = min(kTile, n_kept - j0);
    const int shift = S - n_kept;
    for (int j = tid; j < count; j += kTileThreads) {
        const int o = j0 + j;
        s_list[j] = (o < head) ? o : o + shift;
    }
    __syncthreads();
    const int64_t out_row0 = (int64_t)row * n_kept + j0;
    if (idx_out != nullptr)
        for (int j = tid; j < count; j += kTileThreads) idx_out[out_row0 + j] = s_list[j];
    const int nvec = D >> 3;
    const int64_t row_bytes = (int64_t)D * 2;
    copy_rows<false>(K + ((int64_t)b * ks.b + (int64_t)h * ks.h) * 2, ks.s * 2,
                     K_out + out_row0 * row_bytes, row_bytes, s_list, count, nvec);
    copy_rows<false>(V + ((int64_t)b * vs.b + (int64_t)h * vs.h) * 2, vs.s * 2,
                     V_out + out_row0 * row_bytes, row_bytes, s_list, count, nvec);
}

cudaError_t launch_streaming_compress(const Dims& d, int n_sink, const void* K, const void* V,
                                      void* K_out, void* V_out, int32_t* idx_out,
                                      cudaStream_t st) {
    const int head = n_sink < d.n_kept ? n_sink : d.n_kept;
    dim3 grid((d.n_kept + kTile - 1) / kTile, d.R);
    streaming_compact_kernel<<<grid, kTileThreads, 0, st>>>(
        static_cast<const char*>(K), static_cast<const char*>(V), d.ks, d.vs,
        static_cast<char*>(K_out), static_cast<char*>(V_out), idx_out, d.H, d.S, d.D, d.n_kept,
        head);
    return cudaPeekAtLastError();
}

template <typename T>
__global__ void streaming_score_kernel(uint16_t* __restrict__ out, int S, int lo, int hi,
                                       int64_t total) {
    const uint16_t one = F16Traits<T>::from_float(1.0f);
    for (int64_t i = blockIdx.x * (int64_t)blockDim.x + threadIdx.x; i < total;
         i += (int64_t)gridDim.x * blockDim.x) {
        const int s = (int)(i % S);
        out[i] = (s >= lo && s < hi) ? (uint16_t)0 : one;
    }
}

cudaError_t launch_streaming_score(const Dims& d, int dtype, int n_sink, void* scores_out,
                                   cudaStream_t st) {
    const int n_pruned = d.S - d.n_kept;
    const int64_t total = (int64_t)d.R * d.S;
    const int lo = n_sink, hi = (n_sink + n_pruned < d.S) ? n_sink + n_pruned : d.S;
    int blocks = (int)((total + 255) / 256);
    if (blocks > 148 * 8) blocks = 148 * 8;
    if (dtype == KVP_BF16)
        streaming_score_kernel<__nv_bfloat16>
            <<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(scores_out), d.S, lo, hi, total);
    else
        streaming_score_kernel<__half>
            <<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(scores_out), d.S, lo, hi, total);
    return cudaPeekAtLastError();
}

}  // namespace kvp
