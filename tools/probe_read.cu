// probe_read.cu — HBM read-side probes for the Knorm score stage (not part of the product).
// Variants of "stream [R][S][128] bf16 rows, reduce each row" to see which load path / grid shape
// reaches the copy-measured HBM peak on B200. Build: nvcc -gencode arch=compute_100a,code=sm_100a
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdint>
#include <vector>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA %s at %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

__device__ __forceinline__ int4 ld_plain(const void* p) {
    int4 r; asm volatile("ld.global.nc.L1::no_allocate.v4.s32 {%0,%1,%2,%3}, [%4];" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p)); return r; }
__device__ __forceinline__ int4 ld_hint(const void* p, uint64_t pol) {
    int4 r; asm volatile("ld.global.nc.L1::no_allocate.L2::cache_hint.v4.s32 {%0,%1,%2,%3}, [%4], %5;" : "=r"(r.x), "=r"(r.y), "=r"(r.z), "=r"(r.w) : "l"(p), "l"(pol)); return r; }
struct int8v { int4 a, b; };
__device__ __forceinline__ int8v ld_256(const void* p) {
    int8v r; asm volatile("ld.global.nc.L1::no_allocate.v8.s32 {%0,%1,%2,%3,%4,%5,%6,%7}, [%8];"
        : "=r"(r.a.x), "=r"(r.a.y), "=r"(r.a.z), "=r"(r.a.w), "=r"(r.b.x), "=r"(r.b.y), "=r"(r.b.z), "=r"(r.b.w) : "l"(p)); return r; }

__device__ __forceinline__ float sumsq(int4 v) {
    const uint32_t w[4] = {(uint32_t)v.x, (uint32_t)v.y, (uint32_t)v.z, (uint32_t)v.w};
    float s = 0.f;
#pragma unroll
    for (int j = 0; j < 4; ++j) { float a = __uint_as_float(w[j] << 16), b = __uint_as_float(w[j] & 0xFFFF0000u); s = fmaf(a, a, s); s = fmaf(b, b, s); }
    return s;
}

// MODE 0: plain, 1: evict_last hint, 2: evict_first hint. PERSIST: atomic tile queue. TILE tokens per item.
template <int MODE, bool PERSIST, int U, int TILE>
__global__ void __launch_bounds__(256) score_v4(const char* __restrict__ K, int n_items, float* __restrict__ out, int* counter) {
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane & 15, rsel = lane >> 4;
    uint64_t pol = 0;
    if (MODE == 1) asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(pol));
    if (MODE == 2) asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(pol));
    constexpr int TPW = TILE / 8;       // tokens per warp
    constexpr int ITERS = TPW / 2;
    int item = blockIdx.x;
    while (true) {
        if (PERSIST) { if (tid == 0) s_item = atomicAdd(counter, 1); __syncthreads(); item = s_item; __syncthreads(); }
        if (item >= n_items) break;
        const char* base = K + (size_t)item * TILE * 256 + (size_t)warp * TPW * 256 + rsel * 256 + sub * 16;
        float acc = 0.f;
#pragma unroll 1
        for (int it = 0; it < ITERS; it += U) {
            int4 v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) { const char* p = base + (size_t)(it + u) * 512; v[u] = (MODE == 0) ? ld_plain(p) : ld_hint(p, pol); }
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float ss = sumsq(v[u]);
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
                acc += sqrtf(ss);
            }
        }
        if (sub == 0) out[(size_t)item * 16 + warp * 2 + rsel] = acc;
        if (!PERSIST) break;
    }
}

// 256-bit loads: 8 lanes per row, 4 rows per warp instruction
template <bool PERSIST, int U, int TILE>
__global__ void __launch_bounds__(256) score_v8(const char* __restrict__ K, int n_items, float* __restrict__ out, int* counter) {
    __shared__ int s_item;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    const int sub = lane & 7, rsel = lane >> 3;
    constexpr int TPW = TILE / 8;
    constexpr int ITERS = TPW / 4;
    int item = blockIdx.x;
    while (true) {
        if (PERSIST) { if (tid == 0) s_item = atomicAdd(counter, 1); __syncthreads(); item = s_item; __syncthreads(); }
        if (item >= n_items) break;
        const char* base = K + (size_t)item * TILE * 256 + (size_t)warp * TPW * 256 + rsel * 256 + sub * 32;
        float acc = 0.f;
#pragma unroll 1
        for (int it = 0; it < ITERS; it += U) {
            int8v v[U];
#pragma unroll
            for (int u = 0; u < U; ++u) v[u] = ld_256(base + (size_t)(it + u) * 1024);
#pragma unroll
            for (int u = 0; u < U; ++u) {
                float ss = sumsq(v[u].a) + sumsq(v[u].b);
#pragma unroll
                for (int off = 4; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
                acc += sqrtf(ss);
            }
        }
        if (sub == 0) out[(size_t)item * 32 + warp * 4 + rsel] = acc;
        if (!PERSIST) break;
    }
}

// TMA-less bulk copy pipeline: cp.async.bulk global->shared with mbarrier, 1 producer thread, ring of stages.
template <int STAGES, int ROWS>
__global__ void __launch_bounds__(288) score_bulk(const char* __restrict__ K, int n_chunks, float* __restrict__ out, int* counter) {
    extern __shared__ __align__(128) unsigned char smem[];
    uint64_t* full = reinterpret_cast<uint64_t*>(smem);
    uint64_t* empty = full + STAGES;
    int* s_chunk = reinterpret_cast<int*>(empty + STAGES);
    unsigned char* data = smem + 1024;
    constexpr int BYTES = ROWS * 256;
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) {
        for (int s = 0; s < STAGES; ++s) {
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&full[s])), "r"(1));
            asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[s])), "r"(8));
        }
        asm volatile("fence.mbarrier_init.release.cluster;");
    }
    __syncthreads();
    if (warp == 8) {  // producer warp
        if (lane == 0) {
            int stage = 0, phase = 0;
            while (true) {
                const int chunk = atomicAdd(counter, 1);
                // wait for slot to be free
                uint32_t eb = (uint32_t)__cvta_generic_to_shared(&empty[stage]);
                uint32_t ok = 0;
                while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(eb), "r"(phase ^ 1));
                s_chunk[stage] = chunk;
                uint32_t fb = (uint32_t)__cvta_generic_to_shared(&full[stage]);
                if (chunk >= n_chunks) {
                    asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(fb));
                    break;
                }
                asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(fb), "r"(BYTES));
                uint32_t dst = (uint32_t)__cvta_generic_to_shared(data + (size_t)stage * BYTES);
                asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];"
                             ::"r"(dst), "l"(K + (size_t)chunk * BYTES), "r"(BYTES), "r"(fb) : "memory");
                if (++stage == STAGES) { stage = 0; phase ^= 1; }
            }
        }
    } else {
        int stage = 0, phase = 0;
        const int sub = lane & 15, rsel = lane >> 4;
        float acc = 0.f;
        while (true) {
            uint32_t fb = (uint32_t)__cvta_generic_to_shared(&full[stage]);
            uint32_t ok = 0;
            while (!ok) asm volatile("{ .reg .pred p; mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2; selp.u32 %0, 1, 0, p; }" : "=r"(ok) : "r"(fb), "r"(phase));
            const int chunk = s_chunk[stage];
            if (chunk >= n_chunks) break;
            const unsigned char* d = data + (size_t)stage * BYTES;
            constexpr int RPWARP = ROWS / 8;
#pragma unroll 4
            for (int r = 0; r < RPWARP; r += 2) {
                int4 v = *reinterpret_cast<const int4*>(d + (size_t)(warp * RPWARP + r + rsel) * 256 + sub * 16);
                float ss = sumsq(v);
#pragma unroll
                for (int off = 8; off >= 1; off >>= 1) ss += __shfl_xor_sync(0xFFFFFFFFu, ss, off);
                acc += sqrtf(ss);
            }
            __syncwarp();
            if (lane == 0) asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"((uint32_t)__cvta_generic_to_shared(&empty[stage])));
            if (++stage == STAGES) { stage = 0; phase ^= 1; }
        }
        if (sub == 0) out[blockIdx.x * 16 + warp * 2 + rsel] = acc;
    }
}

int main() {
    const size_t S = 131072, R = 8;
    const size_t bytes = R * S * 256;
    char* K; float* out; int* counter; char* flush;
    CK(cudaMalloc(&K, bytes)); CK(cudaMalloc(&out, 64 << 20)); CK(cudaMalloc(&counter, 4)); CK(cudaMalloc(&flush, 512 << 20));
    CK(cudaMemset(K, 0x3c, bytes));
    cudaEvent_t e0, e1; CK(cudaEventCreate(&e0)); CK(cudaEventCreate(&e1));
    auto run = [&](const char* name, auto launch) {
        std::vector<float> ts;
        for (int i = 0; i < 7; ++i) {
            CK(cudaMemsetAsync(flush, i, 512 << 20)); CK(cudaMemsetAsync(counter, 0, 4));
            CK(cudaEventRecord(e0)); launch(); CK(cudaEventRecord(e1)); CK(cudaEventSynchronize(e1));
            CK(cudaGetLastError());
            float ms; CK(cudaEventElapsedTime(&ms, e0, e1)); if (i >= 2) ts.push_back(ms);
        }
        std::sort(ts.begin(), ts.end());
        printf("%-44s best %7.2f us  median %7.2f us  -> %6.0f GB/s\n", name, ts[0] * 1e3, ts[ts.size() / 2] * 1e3, bytes / (ts[0] * 1e-3) / 1e9);
    };
    const int n1024 = (int)(R * S / 1024), n512 = n1024 * 2, n256 = n1024 * 4;
    run("v4 plain tile1024 grid=items U8", [&] { score_v4<0, false, 8, 1024><<<n1024, 256>>>(K, n1024, out, counter); });
    run("v4 evict_last tile1024 grid=items U8", [&] { score_v4<1, false, 8, 1024><<<n1024, 256>>>(K, n1024, out, counter); });
    run("v4 evict_first tile1024 grid=items U8", [&] { score_v4<2, false, 8, 1024><<<n1024, 256>>>(K, n1024, out, counter); });
    run("v4 plain tile512 grid=items U8", [&] { score_v4<0, false, 8, 512><<<n512, 256>>>(K, n512, out, counter); });
    run("v4 plain tile256 grid=items U8", [&] { score_v4<0, false, 8, 256><<<n256, 256>>>(K, n256, out, counter); });
    run("v4 plain tile256 grid=items U16", [&] { score_v4<0, false, 16, 256><<<n256, 256>>>(K, n256, out, counter); });
    for (int occ : {2, 3, 4, 5, 6, 8}) {
        char nm[64]; snprintf(nm, 64, "v4 plain persist tile512 148x%d U8", occ);
        run(nm, [&] { score_v4<0, true, 8, 512><<<148 * occ, 256>>>(K, n512, out, counter); });
    }
    run("v4 plain persist tile256 148x5 U8", [&] { score_v4<0, true, 8, 256><<<148 * 5, 256>>>(K, n256, out, counter); });
    run("v4 plain persist tile1024 148x5 U16", [&] { score_v4<0, true, 16, 1024><<<148 * 5, 256>>>(K, n1024, out, counter); });
    run("v8 256-bit tile1024 grid=items U4", [&] { score_v8<false, 4, 1024><<<n1024, 256>>>(K, n1024, out, counter); });
    run("v8 256-bit tile512 grid=items U8", [&] { score_v8<false, 8, 512><<<n512, 256>>>(K, n512, out, counter); });
    run("v8 256-bit persist tile512 148x4 U8", [&] { score_v8<true, 8, 512><<<148 * 4, 256>>>(K, n512, out, counter); });
    run("v8 256-bit persist tile512 148x6 U4", [&] { score_v8<true, 4, 512><<<148 * 6, 256>>>(K, n512, out, counter); });
    {
        constexpr int ST = 6, ROWS = 128;  // 6 x 32 KB
        const int smem = 1024 + ST * ROWS * 256;
        CK(cudaFuncSetAttribute(score_bulk<ST, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int n_chunks = (int)(R * S / ROWS);
        run("bulk-copy ring 6x32KB 148x1", [&] { score_bulk<ST, ROWS><<<148, 288, smem>>>(K, n_chunks, out, counter); });
    }
    {
        constexpr int ST = 4, ROWS = 64;  // 4 x 16 KB, 3 CTAs/SM
        const int smem = 1024 + ST * ROWS * 256;
        CK(cudaFuncSetAttribute(score_bulk<ST, ROWS>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        const int n_chunks = (int)(R * S / ROWS);
        run("bulk-copy ring 4x16KB 148x3", [&] { score_bulk<ST, ROWS><<<148 * 3, 288, smem>>>(K, n_chunks, out, counter); });
    }
    // reference: cudaMemcpy D2D of half the buffer (read+write = same bytes moved as `bytes`)
    run("cudaMemcpyAsync D2D (r+w = same bytes)", [&] { CK(cudaMemcpyAsync(flush, K, bytes / 2, cudaMemcpyDeviceToDevice)); });
    return 0;
}
