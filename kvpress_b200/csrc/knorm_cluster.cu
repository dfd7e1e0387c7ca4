// knorm_cluster.cu — KnormPress score + top-k + compaction of a SMALL cache in ONE launch, no scratch, no memset:
// the DecodingPress compaction (reference kvpress/presses/decoding_press.py:68-111 -> scorer_press.py:76-102 with
// knorm_press.py:38 as the score), e.g. [1, 8, 2560, 128] -> [1, 8, 2048, 128] every 512 generated tokens per layer.
//
// At these sizes (tens of MB) the call is latency-bound: the multi-kernel path spends its time in launch gaps, a
// memset node, global atomics and spin-waits between CTAs. Here one thread-block CLUSTER owns one (b, h) row:
//   1. every CTA requests its whole slice of the row with cp.async up front — every K row, and every V row when both
//      fit — so the slice is one memory round trip, then scores it from shared memory (fp32 sum of squares, one
//      rounding);
//   2. the 16-bit ordered keys of the slice are pushed into the shared memory of every CTA of the cluster
//      (distributed shared memory), ONE cluster barrier;
//   3. every CTA now holds the keys of the whole row and derives the exact threshold, the tie budget and the number
//      of kept positions in front of its slice on its own (a 17-step search over the threshold's bits: register-only
//      counts + block reductions, no atomics);
//   4. it ranks its slice and writes the kept K (and V) rows from shared memory — V from global memory when it was
//      not staged — to their final places (ascending positions, ties to the lowest positions — same rule as select_compact.cu).
// No global atomics, no flags, no workspace: the only inter-CTA communication is the key exchange.
#include <cooperative_groups.h>
#include <stdlib.h>

#include "common.cuh"

namespace cg = cooperative_groups;

namespace kvp {

#ifdef KVP_CL_PROFILE
// phase timestamps (globaltimer, ns) of CTA (0, 0), thread 0: tools/cluster_profile.py
__device__ unsigned long long g_cl_prof[16];
__device__ __forceinline__ unsigned long long cl_now() {
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    return t;
}
#define CL_MARK(i) do { if (blockIdx.x == 0 && blockIdx.y == 0 && threadIdx.x == 0) g_cl_prof[i] = cl_now(); } while (0)
extern "C" void kvp_debug_cluster_profile(unsigned long long* out) {
    cudaMemcpyFromSymbol(out, g_cl_prof, 16 * sizeof(unsigned long long));
}
#else
#define CL_MARK(i) do {} while (0)
#endif

constexpr int kClThreads = 256;
constexpr int kClMaxKeysPerThread = 24;  // keys of the row a thread holds in registers: S <= 256 * 24 = 6144
constexpr int kClMaxSmem = 200 * 1024;  // K slice + keys of the row + lists must fit one CTA's shared memory

struct ClusterPlan {
    int C;        // CTAs per cluster (= per row)
    int P;        // positions per CTA (multiple of 8)
    int v_smem;   // 1: the V rows of the slice are staged in shared memory too (they fit)
    int smem;     // dynamic shared memory bytes
    bool ok;
};

static ClusterPlan cluster_plan(const Dims& d, int C) {
    ClusterPlan pl;
    pl.C = C;
    pl.P = ((d.S + C - 1) / C + 7) / 8 * 8;
    const size_t tile = (size_t)pl.P * d.D * 2;
    const size_t keys = (size_t)C * pl.P * 2;
    const size_t list = (size_t)pl.P * 4;
    pl.v_smem = (2 * tile + keys + list + 64 <= (size_t)kClMaxSmem) ? 1 : 0;
    pl.smem = (int)((pl.v_smem ? 2 : 1) * tile + keys + list + 64);
    pl.ok = pl.smem <= kClMaxSmem && d.S <= kClThreads * kClMaxKeysPerThread;
    return pl;
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src) {
    asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"((uint32_t)__cvta_generic_to_shared(smem_dst)),
                 "l"(gmem_src)
                 : "memory");
}
__device__ __forceinline__ void cp_async_wait_all() {
    asm volatile("cp.async.commit_group;\n\tcp.async.wait_group 0;" ::: "memory");
}

template <typename T, int LPR>
__global__ void __launch_bounds__(kClThreads, 1)
knorm_cluster_kernel(const T* __restrict__ K, const T* __restrict__ V, Strides3 ks, Strides3 vs,
                     char* __restrict__ K_out, char* __restrict__ V_out, int32_t* __restrict__ idx_out,
                     uint16_t* __restrict__ scores_out, int H, int S, int D, int n_kept, int P, int v_smem) {
    extern __shared__ __align__(16) unsigned char smem[];
    cg::cluster_group cluster = cg::this_cluster();
    const int C = (int)cluster.num_blocks();
    const int rank = (int)cluster.block_rank();
    const int row = blockIdx.y, b = row / H, h = row % H;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int nvec = D >> 3;

    const size_t tile_bytes = (size_t)P * D * 2;
    int4* ktile = reinterpret_cast<int4*>(smem);                                         // [P][nvec] 16-byte pieces
    int4* vtile = reinterpret_cast<int4*>(smem + tile_bytes);                            // [P][nvec], only if v_smem
    unsigned char* after = smem + (v_smem ? 2 : 1) * tile_bytes;
    uint16_t* all_keys = reinterpret_cast<uint16_t*>(after);                             // [C][P]
    int* list = reinterpret_cast<int*>(after + (size_t)C * P * 2);                       // [P] kept local positions
    __shared__ uint32_t red[2][8];
    __shared__ uint32_t red3[2][3][8];

    // ---- 1. stage the slice [start, start + P): every K row (and every V row when it fits) is requested up front
    // with cp.async, so the whole slice is ONE memory round trip and costs no registers -----------------------------
    const int start = rank * P;
    const int n_rows = max(0, min(P, S - start));
    CL_MARK(0);
    {
        const char* k_src = reinterpret_cast<const char*>(K) + ((int64_t)b * ks.b + (int64_t)h * ks.h) * 2;
        const char* v_src = reinterpret_cast<const char*>(V) + ((int64_t)b * vs.b + (int64_t)h * vs.h) * 2;
        // thread -> (16-byte piece cc, rows r0, r0 + rstep, ...): no integer division per request when D/8 divides 256
        const bool pow2 = (kClThreads % nvec) == 0;
        const int cc0 = pow2 ? tid % nvec : 0, r0 = pow2 ? tid / nvec : 0, rstep = pow2 ? kClThreads / nvec : 0;
        if (pow2) {
            const char* kp = k_src + (int64_t)start * ks.s * 2 + cc0 * 16;
            const char* vp = v_src + (int64_t)start * vs.s * 2 + cc0 * 16;
            for (int r = r0; r < n_rows; r += rstep) cp_async16(&ktile[r * nvec + cc0], kp + (int64_t)r * ks.s * 2);
            if (v_smem)
                for (int r = r0; r < n_rows; r += rstep) cp_async16(&vtile[r * nvec + cc0], vp + (int64_t)r * vs.s * 2);
        } else {
            const int total = n_rows * nvec;
            for (int i = tid; i < total; i += kClThreads) {
                const int r = i / nvec, cc = i - r * nvec;
                cp_async16(&ktile[i], k_src + (int64_t)(start + r) * ks.s * 2 + cc * 16);
                if (v_smem) cp_async16(&vtile[i], v_src + (int64_t)(start + r) * vs.s * 2 + cc * 16);
            }
        }
    }
    // all CTAs of the cluster have started (their shared memory exists) before anyone writes into it; the barrier
    // overlaps the loads in flight
    CL_MARK(1);
    cluster.sync();
    CL_MARK(2);
    cp_async_wait_all();
    __syncthreads();
    CL_MARK(3);

    // ---- score the slice from shared memory --------------------------------------------------------------------------
    uint16_t* my_keys = all_keys + (size_t)rank * P;
    {
        constexpr int RPW = 32 / LPR;
        constexpr int ROWS_PER_PASS = (kClThreads / 32) * RPW;
        const int sub = lane % LPR, rsel = lane / LPR;
        const int n_pass = (P + ROWS_PER_PASS - 1) / ROWS_PER_PASS;
        constexpr int UI = 4;  // rows in flight per sub-warp: the pass is a chain of LDS -> FMA -> shuffles otherwise
#pragma unroll 1
        for (int j0 = 0; j0 < n_pass; j0 += UI) {
            int4 v[UI];
#pragma unroll
            for (int u = 0; u < UI; ++u) {
                const int r = (j0 + u) * ROWS_PER_PASS + warp * RPW + rsel;
                v[u] = make_int4(0, 0, 0, 0);
                if (r < n_rows && sub < nvec) v[u] = ktile[(size_t)r * nvec + sub];
            }
            float ss[UI];
#pragma unroll
            for (int u = 0; u < UI; ++u) {
                const uint32_t w[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z, (uint32_t)v[u].w};
                ss[u] = 0.f;
#pragma unroll
                for (int q = 0; q < 4; ++q) {
                    const float2 f = F16Traits<T>::unpack2(w[q]);
                    ss[u] = fmaf(f.x, f.x, ss[u]);
                    ss[u] = fmaf(f.y, f.y, ss[u]);
                }
            }
#pragma unroll
            for (int off = LPR / 2; off >= 1; off >>= 1)
#pragma unroll
                for (int u = 0; u < UI; ++u) ss[u] += __shfl_xor_sync(0xFFFFFFFFu, ss[u], off);
#pragma unroll
            for (int u = 0; u < UI; ++u) {
                const int r = (j0 + u) * ROWS_PER_PASS + warp * RPW + rsel;
                if (sub == 0 && r < P) {
                    uint16_t key = 0;
                    if (r < n_rows) {
                        // -sqrt(ss) rounded once to the storage dtype (negation is exact): knorm_press.py:38
                        const uint16_t bits = F16Traits<T>::from_float(sqrtf(ss[u])) ^ 0x8000u;
                        key = ordered_key16(bits, F16Traits<T>::kInfBits);
                        if (scores_out != nullptr) scores_out[(size_t)row * S + start + r] = bits;
                    }
                    my_keys[r] = key;
                }
            }
        }
    }
    __syncthreads();
    CL_MARK(4);

    // ---- 2. all-gather of the keys through distributed shared memory ---------------------------------------------
    {
        const int n8 = P / 4;  // 8-byte pieces (4 keys) of the slice
        const uint2* src = reinterpret_cast<const uint2*>(my_keys);
        for (int peer = 1; peer < C; ++peer) {
            const int dst_rank = (rank + peer) % C;  // spread the traffic over the peers
            uint2* dst = reinterpret_cast<uint2*>(cluster.map_shared_rank(all_keys, dst_rank) + (size_t)rank * P);
            for (int i = tid; i < n8; i += kClThreads) dst[i] = src[i];
        }
    }
    CL_MARK(5);
    cluster.sync();  // release / acquire at cluster scope: every slice of all_keys is complete everywhere
    CL_MARK(6);

    // ---- 3. exact threshold of the row, computed redundantly by every CTA -----------------------------------------
    // The whole row's keys sit in shared memory (position s at all_keys[s]: slices are contiguous); every thread takes
    // the positions tid, tid + 256, ... into registers. The threshold T = the n_kept-th largest key is found by a
    // 16-step search over its bits — T |= bit whenever at least n_kept keys are >= T | bit — each step one
    // register-only count + one block reduction. No shared-memory atomics: Knorm scores of a row take ~100 distinct
    // values, so histogram bins would be hammered by every warp at once (the first version of this kernel spent most
    // of its time there).
    uint32_t myk[kClMaxKeysPerThread];
    const int n_mine = (S + kClThreads - 1) / kClThreads;  // <= kClMaxKeysPerThread (cluster_plan)
#pragma unroll
    for (int i = 0; i < kClMaxKeysPerThread; ++i) {
        const int s = i * kClThreads + tid;
        myk[i] = (i < n_mine && s < S) ? (uint32_t)all_keys[s] + 1u : 0u;  // +1: 0 marks "no position", below any key
    }
    auto block_sum3 = [&](uint32_t a, uint32_t b2, uint32_t c, int slot, uint32_t (&out)[3]) {
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            a += __shfl_xor_sync(0xFFFFFFFFu, a, off);
            b2 += __shfl_xor_sync(0xFFFFFFFFu, b2, off);
            c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
        }
        if (lane == 0) {
            red3[slot][0][warp] = a;
            red3[slot][1][warp] = b2;
            red3[slot][2][warp] = c;
        }
        __syncthreads();
        out[0] = out[1] = out[2] = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            out[0] += red3[slot][0][w];
            out[1] += red3[slot][1][w];
            out[2] += red3[slot][2][w];
        }
    };
    uint32_t T1 = 0;  // threshold in the shifted (+1) key space
#pragma unroll 1
    for (int bit = 16; bit >= 0; --bit) {  // shifted keys span 17 bits
        const uint32_t cand = T1 | (1u << bit);
        uint32_t c = 0;
#pragma unroll
        for (int i = 0; i < kClMaxKeysPerThread; ++i) c += myk[i] >= cand;
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) c += __shfl_xor_sync(0xFFFFFFFFu, c, off);
        const int slot = bit & 1;  // alternating slots: one barrier per step is enough
        if (lane == 0) red3[slot][0][warp] = c;
        __syncthreads();
        uint32_t total = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) total += red3[slot][0][w];
        if (total >= (uint32_t)n_kept) T1 = cand;
    }
    // kept (> T) positions of the row, and kept / tied positions in front of this CTA's slice
    uint32_t n_gt = 0, gt_b = 0, eq_b = 0;
#pragma unroll
    for (int i = 0; i < kClMaxKeysPerThread; ++i) {
        const bool front = (i * kClThreads + tid) < start;
        n_gt += myk[i] > T1;
        gt_b += front && myk[i] > T1;
        eq_b += front && myk[i] == T1;
    }
    uint32_t tot[3];
    __syncthreads();  // slot 0 was last read in the bit loop
    block_sum3(n_gt, gt_b, eq_b, 0, tot);
    const uint32_t T16 = T1 - 1u;                       // back to the 16-bit key space (T1 >= 1: n_kept >= 1)
    const uint32_t n_take = (uint32_t)n_kept - tot[0];  // ties (key == T) to take, lowest positions first
    const uint32_t gt_before = tot[1], eq_before = tot[2];
    __syncthreads();

    CL_MARK(7);
    // ---- 4. rank the slice (position order) and copy the kept rows ------------------------------------------------
    uint32_t taken_eq = min(eq_before, n_take);  // ties already granted to lower positions
    const uint32_t out_base = gt_before + taken_eq;
    uint32_t count = 0;                          // kept rows of this slice so far
    for (int c0 = 0; c0 < P; c0 += kClThreads) {
        const int r = c0 + tid;
        const bool valid = r < P && start + r < S;
        const uint32_t key = valid ? my_keys[r] : 0u;
        const bool is_gt = valid && key > T16;
        const bool is_eq = valid && key == T16;
        const unsigned m_gt = __ballot_sync(0xFFFFFFFFu, is_gt);
        const unsigned m_eq = __ballot_sync(0xFFFFFFFFu, is_eq);
        if (lane == 0) {
            red[0][warp] = __popc(m_gt);
            red[1][warp] = __popc(m_eq);
        }
        __syncthreads();
        uint32_t gt_rank = __popc(m_gt & ((1u << lane) - 1u));
        uint32_t eq_rank = __popc(m_eq & ((1u << lane) - 1u));
        uint32_t tot_gt = 0, tot_eq = 0;
#pragma unroll
        for (int w = 0; w < 8; ++w) {
            gt_rank += (w < warp) ? red[0][w] : 0u;
            eq_rank += (w < warp) ? red[1][w] : 0u;
            tot_gt += red[0][w];
            tot_eq += red[1][w];
        }
        const uint32_t tie_room = n_take - taken_eq;  // ties this chunk may still take
        if (is_gt || (is_eq && eq_rank < tie_room)) list[count + gt_rank + min(eq_rank, tie_room)] = r;
        const uint32_t took = min(tot_eq, tie_room);
        count += tot_gt + took;
        taken_eq += took;
        __syncthreads();
    }
    CL_MARK(8);
    if (count == 0) return;
    const int64_t out_row0 = (int64_t)row * n_kept + out_base;
    if (idx_out != nullptr)
        for (uint32_t i = tid; i < count; i += kClThreads) idx_out[out_row0 + i] = start + list[i];
    const int64_t row_bytes = (int64_t)D * 2;
    const char* v_src = reinterpret_cast<const char*>(V) + ((int64_t)b * vs.b + (int64_t)h * vs.h) * 2;
    char* k_dst = K_out + out_row0 * row_bytes;
    char* v_dst = V_out + out_row0 * row_bytes;
    const uint64_t pol_first = l2_policy_evict_first();
    const int total = (int)count * nvec;
    if (v_smem) {  // both tensors come from shared memory: pure stores
        if ((kClThreads % nvec) == 0) {
            const int cc = tid % nvec, rstep = kClThreads / nvec;
            for (int r = tid / nvec; r < (int)count; r += rstep) {
                const int64_t off = (int64_t)r * row_bytes + cc * 16;
                const int src = list[r] * nvec + cc;
                stg_hint(k_dst + off, ktile[src], pol_first);
                stg_hint(v_dst + off, vtile[src], pol_first);
            }
        } else {
            for (int i = tid; i < total; i += kClThreads) {
                const int r = i / nvec, cc = i - r * nvec;
                const int64_t off = (int64_t)r * row_bytes + cc * 16;
                stg_hint(k_dst + off, ktile[(size_t)list[r] * nvec + cc], pol_first);
                stg_hint(v_dst + off, vtile[(size_t)list[r] * nvec + cc], pol_first);
            }
        }
        CL_MARK(9);
        return;
    }
    constexpr int UC = 8;
    for (int base = tid; base < total; base += kClThreads * UC) {
        int4 vv[UC];
#pragma unroll
        for (int u = 0; u < UC; ++u) {
            const int i = base + u * kClThreads;
            if (i < total) {
                const int r = i / nvec, cc = i - r * nvec;
                vv[u] = ldg_hint(v_src + (int64_t)(start + list[r]) * vs.s * 2 + cc * 16, pol_first);
            }
        }
#pragma unroll
        for (int u = 0; u < UC; ++u) {
            const int i = base + u * kClThreads;
            if (i < total) {
                const int r = i / nvec, cc = i - r * nvec;
                const int64_t off = (int64_t)r * row_bytes + cc * 16;
                stg_hint(k_dst + off, ktile[(size_t)list[r] * nvec + cc], pol_first);
                stg_hint(v_dst + off, vv[u], pol_first);
            }
        }
    }
}

template <typename T, int LPR>
static cudaError_t launch_cluster_t(const Dims& d, const ClusterPlan& pl, const void* K, const void* V, void* K_out,
                                    void* V_out, int32_t* idx_out, void* scores_out, cudaStream_t st) {
    auto kern = knorm_cluster_kernel<T, LPR>;
    static PerDeviceOnce smem_set;  // one per <T, LPR> instantiation of this launcher
    cudaError_t e = ensure_dynamic_smem(kern, kClMaxSmem, smem_set);
    if (e != cudaSuccess) return e;
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3(pl.C, d.R, 1);
    cfg.blockDim = dim3(kClThreads, 1, 1);
    cfg.dynamicSmemBytes = (size_t)pl.smem;
    cfg.stream = st;
    cudaLaunchAttribute attr[1];
    attr[0].id = cudaLaunchAttributeClusterDimension;
    attr[0].val.clusterDim.x = pl.C;
    attr[0].val.clusterDim.y = 1;
    attr[0].val.clusterDim.z = 1;
    cfg.attrs = attr;
    cfg.numAttrs = 1;
    return cudaLaunchKernelEx(&cfg, kern, static_cast<const T*>(K), static_cast<const T*>(V), d.ks, d.vs,
                              static_cast<char*>(K_out), static_cast<char*>(V_out), idx_out,
                              static_cast<uint16_t*>(scores_out), d.H, d.S, d.D, d.n_kept, pl.P, pl.v_smem);
}

static bool choose_cluster(const Dims& d, ClusterPlan* pl) {
    static const int knob_c = [] {  // A/B knob: cluster size (portable maximum 8); 0 disables the path
        const char* v = getenv("KVP_KNORM_CLUSTER");
        return (v && *v) ? atoi(v) : 8;
    }();
    if (knob_c <= 0 || d.R > 65535) return false;
    int C = knob_c > 8 ? 8 : knob_c;
    // few positions per row: a smaller cluster keeps every CTA busy
    while (C > 1 && (d.S + C - 1) / C < 64) C >>= 1;
    *pl = cluster_plan(d, C);
    return pl->ok;
}

bool knorm_cluster_applicable(const Dims& d) {
    ClusterPlan pl;
    return choose_cluster(d, &pl);
}

// Returns cudaErrorNotSupported when the cache is too large for the cluster path (the caller takes the
// multi-kernel path instead).
cudaError_t launch_knorm_cluster(const Dims& d, int dtype, const void* K, const void* V, void* K_out, void* V_out,
                                 int32_t* idx_out, void* scores_out, cudaStream_t st) {
    ClusterPlan pl;
    if (!choose_cluster(d, &pl)) return cudaErrorNotSupported;
    const int nvec = d.D / 8;
#define KVP_CL(LPR)                                                                                             \
    return (dtype == KVP_BF16)                                                                                   \
               ? launch_cluster_t<__nv_bfloat16, LPR>(d, pl, K, V, K_out, V_out, idx_out, scores_out, st)       \
               : launch_cluster_t<__half, LPR>(d, pl, K, V, K_out, V_out, idx_out, scores_out, st)
    if (nvec <= 4) KVP_CL(4);
    if (nvec <= 8) KVP_CL(8);
    if (nvec <= 16) KVP_CL(16);
    KVP_CL(32);
#undef KVP_CL
}

}  // namespace kvp
