// expected_attention.cu — stage S for ExpectedAttentionPress (sm_100a: TMA + tcgen05 + TMEM).
//
// Reference semantics (kvpress/presses/expected_attention_press.py:136-165), per kv head h and each
// of its G = Hq/Hkv query heads g, over positions s in [n_sink, S):
//     logit_g(s) = mu_g . k_s / sqrt(d) + k_s^T Sigma_g k_s / (2 d)              (:149-151)
//     p_g        = softmax_s(logit_g)                                              (:152)
//     score(s)   = mean_g p_g(s)          [* ||v_s||_2 after adding epsilon]       (:155-160)
//   and the n_sink first positions are forced to be kept (:163).
// The reference materialises repeat_kv(K)^T twice and a [B,Hq,D,S] einsum intermediate; here
//   kernel 1 (ea_logits_kernel, persistent, one CTA per SM, warp-specialised):
//       TMA streams 128-position K tiles (SWIZZLE_128B) into a 2-stage ring; one thread issues
//       tcgen05.mma  Y = K_tile[128 x D] * Sigma_g^T  for two heads at a time (N = 2D <= 256) into one
//       of two TMEM accumulator buffers; four epilogue warps read Y back (tcgen05.ld, thread = key
//       row) and finish  logit = sum_n k_n (Y_n + 2 sqrt(d) mu_n) / (2d)  with the k row taken from
//       the same shared-memory tile, keep online softmax statistics and store fp32 logits; four
//       more warps stream the matching V rows and store ||v||. Sigma for the G heads (G*D*D*2 bytes)
//       stays resident in shared memory for the CTA's whole range.
//   kernel 2 (ea_finalize_kernel): combines the per-CTA softmax statistics, forms the final score in
//       fp32, rounds it ONCE to the cache dtype, and emits keys + histogram for the select stage.
// Covariance-free mode (use_covariance=False) is a plain streaming GEMV kernel.
#include "common.cuh"
#include "umma.cuh"

namespace kvp {

constexpr int kEaTile = 128;      // key rows per MMA tile (M)
constexpr int kEaThreads = 384;   // 12 warps: TMA, MMA, TMEM-alloc, spare, 4 epilogue, 4 V-norm
constexpr int kEaMaxParts = 160;  // upper bound on CTAs per (b,h) row (>= SM count)

struct EaScratch {
    float* logits;    // [R][G][S_pad]
    float* vnorm;     // [R][S_pad]
    float2* partial;  // [R][G][n_parts] (max, sum exp) per CTA part
};

static inline size_t align256(size_t x) { return (x + 255) / 256 * 256; }

size_t ea_scratch_bytes(const Dims& d) {
    const int G = d.Hq / d.H;
    const size_t S_pad = (size_t)((d.S + kTile - 1) / kTile) * kTile;
    const size_t n_parts = (size_t)((d.S + kScoreChunkGeneric - 1) / kScoreChunkGeneric);
    const size_t parts = n_parts > kEaMaxParts ? n_parts : kEaMaxParts;
    return align256((size_t)d.R * G * S_pad * 4) + align256((size_t)d.R * S_pad * 4) +
           align256((size_t)d.R * G * parts * sizeof(float2));
}

static EaScratch carve_ea(const Dims& d, const Workspace& ws) {
    const int G = d.Hq / d.H;
    const size_t S_pad = (size_t)ws.S_pad;
    char* p = static_cast<char*>(ws.scorer);
    EaScratch s;
    s.logits = reinterpret_cast<float*>(p);
    p += align256((size_t)d.R * G * S_pad * 4);
    s.vnorm = reinterpret_cast<float*>(p);
    p += align256((size_t)d.R * S_pad * 4);
    s.partial = reinterpret_cast<float2*>(p);
    return s;
}

// ---- shared-memory carve-up of ea_logits_kernel ---------------------------------------------------
template <int D, int G>
struct EaSmem {
    static constexpr int kPanels = D / 64;                 // 64-element (128 B) K panels
    static constexpr int kCovHeadPanel = D * 128;          // bytes of one head's [D x 64] panel
    static constexpr int kCovBytes = G * kPanels * kCovHeadPanel;
    static constexpr int kStageBytes = kPanels * kEaTile * 128;
    static constexpr int kStages = 2;
    static constexpr int kCovOff = 0;
    static constexpr int kStageOff = kCovBytes;
    static constexpr int kBiasOff = kStageOff + kStages * kStageBytes;  // float [G][D]
    static constexpr int kBarOff = kBiasOff + G * D * 4;
    static constexpr int kTotal = kBarOff + 256;
};

template <typename T, int D, int G>
__global__ void __launch_bounds__(kEaThreads, 1)
ea_logits_kernel(const __grid_constant__ CUtensorMap mapK, const __grid_constant__ CUtensorMap mapCov,
                 const T* __restrict__ V, Strides3 vs, const T* __restrict__ mu, int H, int Hq, int S,
                 int n_sink, int use_vnorm, int R, int n_tiles128, int ctas_per_row, int n_parts,
                 EaScratch sc, int S_pad) {
    using L = EaSmem<D, G>;
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    // dynamic shared memory is only guaranteed 16-B aligned: round up to 1024 B for the swizzle
    unsigned char* smem = reinterpret_cast<unsigned char*>(
        (reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* s_cov = smem + L::kCovOff;
    unsigned char* s_stage = smem + L::kStageOff;
    float* s_bias = reinterpret_cast<float*>(smem + L::kBiasOff);
    uint64_t* bars = reinterpret_cast<uint64_t*>(smem + L::kBarOff);
    uint64_t* k_full = bars;         // [2]
    uint64_t* k_empty = bars + 2;    // [2]
    uint64_t* t_full = bars + 4;     // [2]
    uint64_t* t_empty = bars + 6;    // [2]
    uint64_t* cov_full = bars + 8;   // [1]
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + 10);
    float* s_red = reinterpret_cast<float*>(bars + 12);  // [4 warps][G][2]

    constexpr int kHalves = (G + 1) / 2;          // head pairs per tile
    constexpr int kHeadsPerHalf = (G >= 2) ? 2 : 1;
    constexpr int kN = kHeadsPerHalf * D;         // MMA N
    constexpr int kBufCols = 256;
    static_assert(kN <= 256, "two heads must fit one MMA");

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    // ---- which (row, tile range) this CTA owns ----------------------------------------------------
    // rows are visited round-robin when there are more rows than CTA groups
    const int n_groups = gridDim.x / ctas_per_row;  // concurrent rows
    const int group = blockIdx.x / ctas_per_row;
    const int part = blockIdx.x % ctas_per_row;

    if (tid == 0) {
        umma::prefetch_tmap(&mapK);
        umma::prefetch_tmap(&mapCov);
        for (int i = 0; i < 2; ++i) {
            umma::mbar_init(&k_full[i], 1);
            umma::mbar_init(&k_empty[i], 1 + 4);  // MMA commit + 4 epilogue warps
            umma::mbar_init(&t_full[i], 1);
            umma::mbar_init(&t_empty[i], 4);
        }
        umma::mbar_init(cov_full, 1);
        umma::mbar_fence_init();
    }
    if (warp == 2) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;

    const int tiles_per_part = (n_tiles128 + ctas_per_row - 1) / ctas_per_row;
    const int t_begin = part * tiles_per_part;
    const int t_end = min(n_tiles128, t_begin + tiles_per_part);
    const float inv_2d = 1.0f / (2.0f * (float)D);
    const float bias_scale = 2.0f * sqrtf((float)D);

    // pipeline state carried across rows (barriers keep flipping)
    uint32_t k_it = 0;  // tiles processed by this role
    uint32_t h_it = 0;  // halves processed by this role
    uint32_t cov_it = 0;

    for (int row = group; row < R; row += n_groups) {
        const int b = row / H, h = row % H;
        const int hq0 = b * Hq + h * G;  // first query head of this kv head in [B*Hq]
        // everyone: previous row's epilogue has finished reading s_bias / s_cov
        __syncthreads();
        for (int i = tid; i < G * D; i += kEaThreads)
            s_bias[i] = bias_scale * F16Traits<T>::to_float(
                                         reinterpret_cast<const uint16_t*>(mu)[(size_t)hq0 * D + i]);
        __syncthreads();

        if (warp == 0) {
            // ===== TMA producer =====
            if (lane == 0) {
                umma::mbar_arrive_expect_tx(cov_full, L::kCovBytes);
                for (int kp = 0; kp < L::kPanels; ++kp)
                    for (int g = 0; g < G; ++g)
                        umma::tma_load_3d(s_cov + (kp * G + g) * L::kCovHeadPanel, &mapCov, cov_full,
                                          kp * 64, 0, hq0 + g);
                for (int t = t_begin; t < t_end; ++t, ++k_it) {
                    const int stage = k_it & 1;
                    umma::mbar_wait(&k_empty[stage], ((k_it >> 1) & 1) ^ 1);
                    umma::mbar_arrive_expect_tx(&k_full[stage], L::kStageBytes);
                    for (int kp = 0; kp < L::kPanels; ++kp)
                        umma::tma_load_4d(s_stage + stage * L::kStageBytes + kp * (kEaTile * 128),
                                          &mapK, &k_full[stage], kp * 64, t * kEaTile, h, b);
                }
            }
        } else if (warp == 1) {
            // ===== MMA issuer =====
            if (lane == 0) {
                const uint32_t idesc =
                    umma::instr_desc_f16(kEaTile, kN, F16Traits<T>::kMmaFormat);
                umma::mbar_wait(cov_full, cov_it & 1);
                for (int t = t_begin; t < t_end; ++t, ++k_it) {
                    const int stage = k_it & 1;
                    umma::mbar_wait(&k_full[stage], (k_it >> 1) & 1);
                    umma::fence_after_sync();
                    const uint32_t a_base = umma::smem_u32(s_stage + stage * L::kStageBytes);
#pragma unroll 1
                    for (int half = 0; half < kHalves; ++half, ++h_it) {
                        const int buf = h_it & 1;
                        umma::mbar_wait(&t_empty[buf], ((h_it >> 1) & 1) ^ 1);
                        umma::fence_after_sync();
#pragma unroll
                        for (int k = 0; k < D / 16; ++k) {
                            const int kp = k >> 2, kk = k & 3;
                            const uint64_t da =
                                umma::smem_desc_sw128(a_base + kp * (kEaTile * 128) + kk * 32);
                            const uint64_t db = umma::smem_desc_sw128(
                                umma::smem_u32(s_cov + (kp * G + half * kHeadsPerHalf) * L::kCovHeadPanel) +
                                kk * 32);
                            umma::mma_f16_ss(tmem + buf * kBufCols, da, db, idesc, k > 0);
                        }
                        umma::mma_commit(&t_full[buf]);
                    }
                    umma::mma_commit(&k_empty[stage]);
                }
            }
            ++cov_it;
        } else if (warp >= 4 && warp < 8) {
            // ===== epilogue: thread = key row of the tile =====
            const int ew = warp - 4;
            const int r = ew * 32 + lane;  // row in tile == TMEM lane
            const uint32_t lane_base = (uint32_t)(ew * 32) << 16;
            float run_m[G], run_z[G];
#pragma unroll
            for (int g = 0; g < G; ++g) {
                run_m[g] = -INFINITY;
                run_z[g] = 0.f;
            }
            for (int t = t_begin; t < t_end; ++t, ++k_it) {
                const int stage = k_it & 1;
                umma::mbar_wait(&k_full[stage], (k_it >> 1) & 1);
                const unsigned char* krow = s_stage + stage * L::kStageBytes;
                const int s = t * kEaTile + r;
                const bool valid = (s >= n_sink) && (s < S);
#pragma unroll 1
                for (int half = 0; half < kHalves; ++half, ++h_it) {
                    const int buf = h_it & 1;
                    umma::mbar_wait(&t_full[buf], (h_it >> 1) & 1);
                    umma::fence_after_sync();
                    float acc[kHeadsPerHalf];
#pragma unroll
                    for (int q = 0; q < kHeadsPerHalf; ++q) acc[q] = 0.f;
#pragma unroll 1
                    for (int c0 = 0; c0 < D; c0 += 32) {
                        uint32_t y[kHeadsPerHalf][32];
#pragma unroll
                        for (int q = 0; q < kHeadsPerHalf; ++q)
                            umma::tmem_ld32(tmem + lane_base + buf * kBufCols + q * D + c0, y[q]);
                        // the k row chunk while the TMEM loads are in flight
                        float kf[32];
                        const int kp = c0 >> 6;
#pragma unroll
                        for (int ch = 0; ch < 4; ++ch) {
                            const uint4 v = *reinterpret_cast<const uint4*>(
                                krow + kp * (kEaTile * 128) +
                                umma::sw128_offset(r, ((c0 & 63) >> 3) + ch));
                            const uint32_t w4[4] = {v.x, v.y, v.z, v.w};
#pragma unroll
                            for (int j = 0; j < 4; ++j) {
                                const float2 f = F16Traits<T>::unpack2(w4[j]);
                                kf[ch * 8 + j * 2] = f.x;
                                kf[ch * 8 + j * 2 + 1] = f.y;
                            }
                        }
                        umma::tmem_ld_wait();
#pragma unroll
                        for (int q = 0; q < kHeadsPerHalf; ++q) {
                            const float* bias = s_bias + (half * kHeadsPerHalf + q) * D + c0;
#pragma unroll
                            for (int j = 0; j < 32; j += 4) {
                                const float4 bq = *reinterpret_cast<const float4*>(bias + j);
                                acc[q] = fmaf(kf[j], __uint_as_float(y[q][j]) + bq.x, acc[q]);
                                acc[q] = fmaf(kf[j + 1], __uint_as_float(y[q][j + 1]) + bq.y, acc[q]);
                                acc[q] = fmaf(kf[j + 2], __uint_as_float(y[q][j + 2]) + bq.z, acc[q]);
                                acc[q] = fmaf(kf[j + 3], __uint_as_float(y[q][j + 3]) + bq.w, acc[q]);
                            }
                        }
                    }
                    // accumulator buffer can be overwritten by the next MMA
                    umma::fence_before_sync();
                    __syncwarp();
                    if (lane == 0) umma::mbar_arrive(&t_empty[buf]);
#pragma unroll
                    for (int q = 0; q < kHeadsPerHalf; ++q) {
                        const int g = half * kHeadsPerHalf + q;
                        if (g < G) {
                            const float lg = acc[q] * inv_2d;
                            if (valid) {
                                sc.logits[((size_t)row * G + g) * S_pad + s] = lg;
                                const float m_new = fmaxf(run_m[g], lg);
                                run_z[g] = run_z[g] * __expf(run_m[g] - m_new) + __expf(lg - m_new);
                                run_m[g] = m_new;
                            }
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) umma::mbar_arrive(&k_empty[stage]);
            }
            // ---- CTA-level (max, sum-exp) per head -> partial[row][g][part] --------------------------
#pragma unroll
            for (int g = 0; g < G; ++g) {
                float m = run_m[g], z = run_z[g];
#pragma unroll
                for (int off = 16; off >= 1; off >>= 1) {
                    const float m2 = __shfl_xor_sync(0xFFFFFFFFu, m, off);
                    const float z2 = __shfl_xor_sync(0xFFFFFFFFu, z, off);
                    const float mn = fmaxf(m, m2);
                    z = (mn == -INFINITY) ? 0.f : z * __expf(m - mn) + z2 * __expf(m2 - mn);
                    m = mn;
                }
                if (lane == 0) {
                    s_red[(ew * G + g) * 2] = m;
                    s_red[(ew * G + g) * 2 + 1] = z;
                }
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");  // the four epilogue warps only
            if (ew == 0 && lane < G) {
                float m = -INFINITY, z = 0.f;
                for (int w = 0; w < 4; ++w) {
                    const float m2 = s_red[(w * G + lane) * 2], z2 = s_red[(w * G + lane) * 2 + 1];
                    const float mn = fmaxf(m, m2);
                    z = (mn == -INFINITY) ? 0.f : z * __expf(m - mn) + z2 * __expf(m2 - mn);
                    m = mn;
                }
                sc.partial[((size_t)row * G + lane) * n_parts + part] = make_float2(m, z);
            }
            asm volatile("bar.sync 1, 128;" ::: "memory");
        } else if (warp >= 8) {
            // ===== V norms for the same tiles (plain streaming loads, 16 in flight per lane) =====
            if (use_vnorm) {
                const int vw = warp - 8;
                const int sub = lane & 15, rsel = lane >> 4;
                constexpr int nvec = D / 8;
                const T* vbase = V + (int64_t)b * vs.b + (int64_t)h * vs.h + sub * 8;
                const uint64_t pol = l2_policy_evict_first();
                for (int t = t_begin; t < t_end; ++t) {
                    // 128 rows per tile, 4 warps x 2 rows per load => 16 loads per lane
                    int4 v[16];
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const int s = t * kEaTile + u * 8 + vw * 2 + rsel;
                        v[u] = make_int4(0, 0, 0, 0);
                        if (s < S && sub < nvec) v[u] = ldg_hint(vbase + (int64_t)s * vs.s, pol);
                    }
#pragma unroll
                    for (int u = 0; u < 16; ++u) {
                        const uint32_t w4[4] = {(uint32_t)v[u].x, (uint32_t)v[u].y, (uint32_t)v[u].z,
                                                (uint32_t)v[u].w};
                        float ss = 0.f;
#pragma unroll
                        for (int j = 0; j < 4; ++j) {
                            const float2 f = F16Traits<T>::unpack2(w4[j]);
                            ss = fmaf(f.x, f.x, ss);
                            ss = fmaf(f.y, f.y, ss);
                        }
#pragma unroll
                        for (int off = 8; off >= 1; off >>= 1)
                            ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
                        const int s = t * kEaTile + u * 8 + vw * 2 + rsel;
                        if (sub == 0 && s < S) sc.vnorm[(size_t)row * S_pad + s] = sqrtf(ss);
                    }
                }
            }
        }
        if (warp == 0) ++cov_it;  // keep the producer's notion of the cov phase in step (unused there)
    }

    umma::fence_before_sync();
    __syncthreads();
    if (warp == 2) umma::tmem_dealloc(tmem, 512);
}

// ---- covariance-free logits: mu.k / sqrt(d) (+ ||v||) with plain streaming loads -------------------
template <typename T>
__global__ void __launch_bounds__(256)
ea_mu_logits_kernel(const T* __restrict__ K, const T* __restrict__ V, Strides3 ks, Strides3 vs,
                    const T* __restrict__ mu, int H, int Hq, int G, int S, int D, int n_sink,
                    int use_vnorm, int n_parts, EaScratch sc, int S_pad) {
    __shared__ float s_mu[8 * 256];
    __shared__ float s_red[8][8][2];
    const int chunk = blockIdx.x, row = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int b = row / H, h = row % H;
    const int hq0 = b * Hq + h * G;
    const float scale = rsqrtf((float)D);
    for (int i = tid; i < G * D; i += 256)
        s_mu[i] = F16Traits<T>::to_float(reinterpret_cast<const uint16_t*>(mu)[(size_t)hq0 * D + i]);
    __syncthreads();
    // 32 lanes cover one row: lane handles 16-byte pieces lane, lane+32 (D <= 256 -> at most 1)
    const int nvec = D >> 3;
    float run_m[8], run_z[8];
    for (int g = 0; g < 8; ++g) {
        run_m[g] = -INFINITY;
        run_z[g] = 0.f;
    }
    for (int i = 0; i < kScoreChunkGeneric / 8; ++i) {
        const int s = chunk * kScoreChunkGeneric + warp * (kScoreChunkGeneric / 8) + i;
        if (s >= S) break;
        float kf[8];
        float vv = 0.f;
#pragma unroll
        for (int j = 0; j < 8; ++j) kf[j] = 0.f;
        if (lane < nvec) {
            const int4 kv = ldg_plain(K + (int64_t)b * ks.b + (int64_t)h * ks.h + (int64_t)s * ks.s + lane * 8);
            const uint32_t w4[4] = {(uint32_t)kv.x, (uint32_t)kv.y, (uint32_t)kv.z, (uint32_t)kv.w};
#pragma unroll
            for (int j = 0; j < 4; ++j) {
                const float2 f = F16Traits<T>::unpack2(w4[j]);
                kf[2 * j] = f.x;
                kf[2 * j + 1] = f.y;
            }
            if (use_vnorm) {
                const int4 v4 = ldg_plain(V + (int64_t)b * vs.b + (int64_t)h * vs.h + (int64_t)s * vs.s + lane * 8);
                const uint32_t x4[4] = {(uint32_t)v4.x, (uint32_t)v4.y, (uint32_t)v4.z, (uint32_t)v4.w};
#pragma unroll
                for (int j = 0; j < 4; ++j) {
                    const float2 f = F16Traits<T>::unpack2(x4[j]);
                    vv = fmaf(f.x, f.x, vv);
                    vv = fmaf(f.y, f.y, vv);
                }
            }
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) vv += __shfl_xor_sync(0xFFFFFFFFu, vv, off);
        if (lane == 0 && use_vnorm) sc.vnorm[(size_t)row * S_pad + s] = sqrtf(vv);
        for (int g = 0; g < G; ++g) {
            float dot = 0.f;
            if (lane < nvec) {
#pragma unroll
                for (int j = 0; j < 8; ++j) dot = fmaf(kf[j], s_mu[g * D + lane * 8 + j], dot);
            }
#pragma unroll
            for (int off = 16; off >= 1; off >>= 1) dot += __shfl_xor_sync(0xFFFFFFFFu, dot, off);
            const float lg = dot * scale;
            if (s >= n_sink) {
                if (lane == 0) sc.logits[((size_t)row * G + g) * S_pad + s] = lg;
                const float mn = fmaxf(run_m[g], lg);
                run_z[g] = run_z[g] * __expf(run_m[g] - mn) + __expf(lg - mn);
                run_m[g] = mn;
            }
        }
    }
    if (lane == 0)
        for (int g = 0; g < G; ++g) {
            s_red[warp][g][0] = run_m[g];
            s_red[warp][g][1] = run_z[g];
        }
    __syncthreads();
    if (tid < G) {
        float m = -INFINITY, z = 0.f;
        for (int w = 0; w < 8; ++w) {
            const float m2 = s_red[w][tid][0], z2 = s_red[w][tid][1];
            const float mn = fmaxf(m, m2);
            z = (mn == -INFINITY) ? 0.f : z * __expf(m - mn) + z2 * __expf(m2 - mn);
            m = mn;
        }
        sc.partial[((size_t)row * G + tid) * n_parts + chunk] = make_float2(m, z);
    }
}

// ---- finalize: softmax normalisation, group mean, * ||v||, ONE rounding, keys + histogram -----------
template <typename T>
__global__ void __launch_bounds__(kTileThreads)
ea_finalize_kernel(int G, int S, int n_sink, int use_vnorm, float eps, int n_parts, EaScratch sc,
                   Workspace ws, uint16_t* __restrict__ scores_out) {
    __shared__ uint16_t skeys[kTile];
    __shared__ uint16_t sscores[kTile];
    __shared__ uint32_t shist[256];
    __shared__ float s_m[8], s_iz[8];
    __shared__ float s_max[8];
    const int tile = blockIdx.x, row = blockIdx.y, tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    shist[tid] = 0;
    if (warp < G) {
        float m = -INFINITY, z = 0.f;
        for (int p = lane; p < n_parts; p += 32) {
            const float2 pz = sc.partial[((size_t)row * G + warp) * n_parts + p];
            const float mn = fmaxf(m, pz.x);
            z = (mn == -INFINITY) ? 0.f : z * __expf(m - mn) + pz.y * __expf(pz.x - mn);
            m = mn;
        }
#pragma unroll
        for (int off = 16; off >= 1; off >>= 1) {
            const float m2 = __shfl_xor_sync(0xFFFFFFFFu, m, off);
            const float z2 = __shfl_xor_sync(0xFFFFFFFFu, z, off);
            const float mn = fmaxf(m, m2);
            z = (mn == -INFINITY) ? 0.f : z * __expf(m - mn) + z2 * __expf(m2 - mn);
            m = mn;
        }
        if (lane == 0) {
            s_m[warp] = m;
            s_iz[warp] = 1.0f / z;
        }
    }
    __syncthreads();
    const int s = tile * kTile + tid;
    float score = 0.f;
    uint16_t bits = 0, key = 0;
    float fmax_valid = -INFINITY;
    if (s < S) {
        if (s < n_sink) {
            key = kForcedKey;
        } else {
            float p = 0.f;
            for (int g = 0; g < G; ++g)
                p += __expf(sc.logits[((size_t)row * G + g) * ws.S_pad + s] - s_m[g]) * s_iz[g];
            p *= (1.0f / (float)G);
            score = use_vnorm ? (p + eps) * sc.vnorm[(size_t)row * ws.S_pad + s] : p;
            bits = F16Traits<T>::from_float(score);
            key = ordered_key16(bits, F16Traits<T>::kInfBits);
            fmax_valid = F16Traits<T>::to_float(bits);
        }
    }
    skeys[tid] = key;
    sscores[tid] = bits;
    // max over valid scores of the whole tensor (for the reference's max+1 sentinel)
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1)
        fmax_valid = fmaxf(fmax_valid, __shfl_xor_sync(0xFFFFFFFFu, fmax_valid, off));
    if (lane == 0) s_max[warp] = fmax_valid;
    __syncthreads();
    if (tid == 0) {
        float m = s_max[0];
        for (int w = 1; w < 8; ++w) m = fmaxf(m, s_max[w]);
        if (m > -INFINITY) {
            // order-preserving float -> uint so atomicMax works (counters are zero-initialised)
            const uint32_t u = __float_as_uint(m);
            const uint32_t ord = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
            atomicMax(&ws.counters[kCounterMaxSlot(ws.R)], ord);
        }
    }
    flush_chunk_keys<1>(skeys, sscores, shist, row, tile * kTile, S, ws, scores_out);
}

// scores_out[..., lo:hi] = round(max_valid_score + 1) — the value the reference pads with
// (expected_attention_press.py:163, snapkv_press.py:103: `scores.max().item() + 1`).
template <typename T>
__global__ void fill_sentinel_kernel(uint16_t* __restrict__ scores_out, int R, int S, int lo, int hi,
                                     const uint32_t* __restrict__ max_slot) {
    const uint32_t ord = *max_slot;
    const uint32_t u = (ord & 0x80000000u) ? (ord & 0x7FFFFFFFu) : ~ord;
    const float sentinel = __uint_as_float(u) + 1.0f;
    const uint16_t bits = F16Traits<T>::from_float(sentinel);
    const int width = hi - lo;
    for (int i = blockIdx.x * blockDim.x + threadIdx.x; i < R * width; i += gridDim.x * blockDim.x)
        scores_out[(size_t)(i / width) * S + lo + (i % width)] = bits;
}

cudaError_t launch_fill_sentinel(int dtype, void* scores_out, int R, int S, int lo, int hi,
                                 const Workspace& ws, cudaStream_t st) {
    if (hi <= lo) return cudaSuccess;
    const uint32_t* slot = ws.counters + kCounterMaxSlot(R);
    int blocks = (R * (hi - lo) + 255) / 256;
    if (blocks > 1024) blocks = 1024;
    if (dtype == KVP_BF16)
        fill_sentinel_kernel<__nv_bfloat16><<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(scores_out), R, S, lo, hi, slot);
    else
        fill_sentinel_kernel<__half><<<blocks, 256, 0, st>>>(static_cast<uint16_t*>(scores_out), R, S, lo, hi, slot);
    return cudaPeekAtLastError();
}

// ---- host launcher -----------------------------------------------------------------------------------
template <typename T, int D, int G>
static cudaError_t launch_ea_logits_t(const Dims& d, const void* K, const void* V, const void* mu,
                                      const void* cov, int n_sink, int use_vnorm, const Workspace& ws,
                                      const EaScratch& sc, int* n_parts_out, cudaStream_t st) {
    using L = EaSmem<D, G>;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev);
        if (sm_count <= 0) sm_count = 148;
    }
    const int n_tiles128 = (d.S + kEaTile - 1) / kEaTile;
    int ctas_per_row = sm_count / d.R;
    if (ctas_per_row < 1) ctas_per_row = 1;
    if (ctas_per_row > n_tiles128) ctas_per_row = n_tiles128;
    if (ctas_per_row > kEaMaxParts) ctas_per_row = kEaMaxParts;
    int n_groups = sm_count / ctas_per_row;
    if (n_groups > d.R) n_groups = d.R;
    const int grid = n_groups * ctas_per_row;
    *n_parts_out = ctas_per_row;

    CUtensorMap mapK, mapCov;
    {
        const uint64_t row_b = (uint64_t)d.ks.s * 2;
        const uint64_t h_b = d.H > 1 ? (uint64_t)d.ks.h * 2 : row_b * (uint64_t)d.S;
        const uint64_t b_b = d.B > 1 ? (uint64_t)d.ks.b * 2 : h_b * (uint64_t)d.H;
        const uint64_t dims[4] = {(uint64_t)D, (uint64_t)d.S, (uint64_t)d.H, (uint64_t)d.B};
        const uint64_t str[4] = {0, row_b, h_b, b_b};
        const uint32_t box[4] = {64, (uint32_t)kEaTile, 1, 1};
        cudaError_t e = make_tmap_16bit(&mapK, K, 4, dims, str, box);
        if (e != cudaSuccess) return e;
    }
    {
        const uint64_t dims[3] = {(uint64_t)D, (uint64_t)D, (uint64_t)d.B * d.Hq};
        const uint64_t str[3] = {0, (uint64_t)D * 2, (uint64_t)D * D * 2};
        const uint32_t box[3] = {64, (uint32_t)D, 1};
        cudaError_t e = make_tmap_16bit(&mapCov, cov, 3, dims, str, box);
        if (e != cudaSuccess) return e;
    }
    const int smem = L::kTotal + 1024;
    auto kern = ea_logits_kernel<T, D, G>;
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    if (e != cudaSuccess) return e;
    kern<<<grid, kEaThreads, smem, st>>>(mapK, mapCov, static_cast<const T*>(V), d.vs,
                                         static_cast<const T*>(mu), d.H, d.Hq, d.S, n_sink, use_vnorm,
                                         d.R, n_tiles128, ctas_per_row, ctas_per_row, sc, ws.S_pad);
    return cudaPeekAtLastError();
}

template <typename T>
static cudaError_t launch_ea_t(const Dims& d, int dtype, const void* K, const void* V, const void* mu,
                               const void* cov, float eps, int n_sink, int use_vnorm,
                               const Workspace& ws, void* scores_out, cudaStream_t st) {
    const int G = d.Hq / d.H;
    if (G > 8) return cudaErrorNotSupported;
    const EaScratch sc = carve_ea(d, ws);
    int n_parts = 0;
    cudaError_t e = cudaSuccess;
    if (cov != nullptr) {
        // tensor-core path: head_dim 64 or 128, up to 4 query heads per kv head resident in smem
        if (d.D == 128 && G == 1) e = launch_ea_logits_t<T, 128, 1>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else if (d.D == 128 && G == 2) e = launch_ea_logits_t<T, 128, 2>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else if (d.D == 128 && G == 4) e = launch_ea_logits_t<T, 128, 4>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else if (d.D == 64 && G == 1) e = launch_ea_logits_t<T, 64, 1>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else if (d.D == 64 && G == 2) e = launch_ea_logits_t<T, 64, 2>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else if (d.D == 64 && G == 4) e = launch_ea_logits_t<T, 64, 4>(d, K, V, mu, cov, n_sink, use_vnorm, ws, sc, &n_parts, st);
        else return cudaErrorNotSupported;
    } else {
        n_parts = (d.S + kScoreChunkGeneric - 1) / kScoreChunkGeneric;
        dim3 grid(n_parts, d.R);
        ea_mu_logits_kernel<T><<<grid, 256, 0, st>>>(static_cast<const T*>(K), static_cast<const T*>(V), d.ks, d.vs,
                                                     static_cast<const T*>(mu), d.H, d.Hq, G, d.S, d.D, n_sink,
                                                     use_vnorm, n_parts, sc, ws.S_pad);
        e = cudaPeekAtLastError();
    }
    if (e != cudaSuccess) return e;
    dim3 grid2(ws.n_tiles, d.R);
    ea_finalize_kernel<T><<<grid2, kTileThreads, 0, st>>>(G, d.S, n_sink, use_vnorm, eps, n_parts, sc, ws,
                                                          static_cast<uint16_t*>(scores_out));
    e = cudaPeekAtLastError();
    if (e != cudaSuccess) return e;
    if (scores_out != nullptr) e = launch_fill_sentinel(dtype, scores_out, d.R, d.S, 0, n_sink, ws, st);
    return e;
}

cudaError_t launch_ea_score(const Dims& d, int dtype, const void* K, const void* V, const void* mu,
                            const void* cov, float eps, int n_sink, int use_vnorm,
                            const Workspace& ws, void* scores_out, bool want_keys,
                            cudaStream_t st) {
    (void)want_keys;  // keys + histogram are always produced; the select stage is skipped by the caller
    if (dtype == KVP_BF16)
        return launch_ea_t<__nv_bfloat16>(d, dtype, K, V, mu, cov, eps, n_sink, use_vnorm, ws, scores_out, st);
    return launch_ea_t<__half>(d, dtype, K, V, mu, cov, eps, n_sink, use_vnorm, ws, scores_out, st);
}

}  // namespace kvp
