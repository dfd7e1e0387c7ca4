// umma_rate.cu — how fast does tcgen05.mma (cta_group::1, SS, M=128, bf16) really run on one SM,
// alone and next to tcgen05.ld / shared-memory traffic? (probe, not part of the product)
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cuda_bf16.h>
#include "../kvpress_b200/csrc/umma.cuh"
#include "../kvpress_b200/csrc/tmap.cu"
using namespace kvp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

// MODE 0: MMA only. MODE 1: + 4 warps doing tcgen05.ld of the other buffer continuously.
// MODE 2: + 4 warps doing LDS.128 sweeps of the A tile.  N = MMA N.
template <int MODE, int N>
__global__ void __launch_bounds__(256) rate_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                   int iters, long long* cycles, float* sink) {
    extern __shared__ __align__(1024) unsigned char smem_raw[];
    unsigned char* smem = reinterpret_cast<unsigned char*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
    unsigned char* sA = smem;                    // 2 panels x 16 KB
    unsigned char* sB = smem + 32768;            // 2 panels x 32 KB
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 32768 + 65536);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
    volatile int* stop = reinterpret_cast<volatile int*>(bar_full + 3);
    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;
    if (tid == 0) { umma::mbar_init(bar_full, 1); umma::mbar_init(bar_mma, 1); umma::mbar_fence_init(); *stop = 0; }
    if (warp == 0) umma::tmem_alloc(tmem_slot, 512);
    umma::fence_before_sync(); __syncthreads(); umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        umma::mbar_arrive_expect_tx(bar_full, 32768 + 65536);
        for (int kp = 0; kp < 2; ++kp) { umma::tma_load_3d(sA + kp * 16384, &mapA, bar_full, kp * 64, 0, 0); umma::tma_load_3d(sB + kp * 32768, &mapB, bar_full, kp * 64, 0, 0); }
        umma::mbar_wait(bar_full, 0);
    }
    __syncthreads();
    if (warp == 0) {
        if (lane == 0) {
            umma::fence_after_sync();
            const uint32_t idesc = umma::instr_desc_f16(128, N, 1);
            const long long t0 = clock64();
            for (int it = 0; it < iters; ++it) {
                for (int k = 0; k < 8; ++k) {
                    const int kp = k / 4, kk = k % 4;
                    umma::mma_f16_ss(tmem, umma::smem_desc_sw128(umma::smem_u32(sA + kp * 16384) + kk * 32),
                                     umma::smem_desc_sw128(umma::smem_u32(sB + kp * 32768) + kk * 32), idesc, k > 0);
                }
            }
            umma::mma_commit(bar_mma);
            umma::mbar_wait(bar_mma, 0);
            const long long t1 = clock64();
            if (blockIdx.x == 0) cycles[0] = t1 - t0;
            *stop = 1;
        }
    } else if (warp >= 4) {
        float acc = 0.f;
        if (MODE == 1) {
            const uint32_t lane_base = (uint32_t)((warp & 3) * 32) << 16;
            while (!*stop) {
                uint32_t r[32];
                for (int c0 = 256; c0 < 512; c0 += 32) { umma::tmem_ld32(tmem + lane_base + c0, r); umma::tmem_ld_wait(); acc += __uint_as_float(r[lane & 31]); }
            }
        } else if (MODE == 2) {
            const int r = (warp & 3) * 32 + lane;
            while (!*stop) {
                for (int kp = 0; kp < 2; ++kp) for (int ch = 0; ch < 8; ++ch) {
                    const uint4 v = *reinterpret_cast<const uint4*>(sA + kp * 16384 + umma::sw128_offset(r, ch));
                    acc += __uint_as_float(v.x) + __uint_as_float(v.w);
                }
            }
        }
        if (acc == 123.456f) sink[tid] = acc;
    }
    umma::fence_before_sync(); __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 512);
}

int main() {
    const int M = 128, NB = 256, KD = 128;
    std::vector<__nv_bfloat16> hA(M * KD), hB(NB * KD);
    for (auto& v : hA) v = __float2bfloat16((rand() % 200 - 100) / 100.f);
    for (auto& v : hB) v = __float2bfloat16((rand() % 200 - 100) / 100.f);
    __nv_bfloat16 *dA, *dB; long long* dC; float* dS;
    CK(cudaMalloc(&dA, hA.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dC, 8)); CK(cudaMalloc(&dS, 4096));
    CK(cudaMemcpy(dA, hA.data(), hA.size() * 2, cudaMemcpyHostToDevice)); CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap mapA, mapB;
    { uint64_t dims[3] = {KD, M, 1}, str[3] = {0, KD * 2, (uint64_t)M * KD * 2}; uint32_t box[3] = {64, 128, 1}; CK(make_tmap_16bit(&mapA, dA, 3, dims, str, box)); }
    { uint64_t dims[3] = {KD, NB, 1}, str[3] = {0, KD * 2, (uint64_t)NB * KD * 2}; uint32_t box[3] = {64, 256, 1}; CK(make_tmap_16bit(&mapB, dB, 3, dims, str, box)); }
    const int smem = 32768 + 65536 + 256 + 1024;
    const int iters = 2000;
    auto run = [&](const char* name, auto kern, int N, int grid) {
        CK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        kern<<<grid, 256, smem>>>(mapA, mapB, iters, dC, dS);
        CK(cudaDeviceSynchronize());
        long long c; CK(cudaMemcpy(&c, dC, 8, cudaMemcpyDeviceToHost));
        const double per_instr = (double)c / (iters * 8);
        printf("%-44s grid %3d: %8.1f cycles per MMA (M=128,N=%d,K=16) -> %7.0f MAC/clk/SM\n", name, grid, per_instr, N, 128.0 * N * 16 / per_instr);
    };
    run("MMA only N=256", rate_kernel<0, 256>, 256, 1);
    run("MMA only N=256", rate_kernel<0, 256>, 256, 148);
    run("MMA only N=128", rate_kernel<0, 128>, 128, 148);
    run("MMA only N=64", rate_kernel<0, 64>, 64, 148);
    run("MMA N=256 + 4 warps tcgen05.ld", rate_kernel<1, 256>, 256, 148);
    run("MMA N=256 + 4 warps LDS.128 on A", rate_kernel<2, 256>, 256, 148);
    return 0;
}
