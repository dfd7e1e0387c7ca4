// umma_smoke.cu — validates csrc/umma.cuh conventions on a B200 (not part of the product):
// TMA SWIZZLE_128B panels -> tcgen05.mma (M=128, N=256, K=128, bf16 -> fp32 in TMEM) -> tcgen05.ld,
// plus generic reads of a swizzled panel via sw128_offset. Compares with a CPU reference.
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <cmath>
#include <cuda_bf16.h>
#include "../kvpress_b200/csrc/umma.cuh"
#include "../kvpress_b200/csrc/tmap.cu"

using namespace kvp;
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("CUDA error %s at line %d\n", cudaGetErrorString(e), __LINE__); exit(1);} } while (0)

constexpr int M = 128, N = 256, KD = 128;

__global__ void __launch_bounds__(128) smoke_kernel(const __grid_constant__ CUtensorMap mapA, const __grid_constant__ CUtensorMap mapB,
                                                    float* __restrict__ D, float* __restrict__ rowdot) {
    extern __shared__ __align__(1024) unsigned char smem[];
    unsigned char* sA = smem;                    // 2 panels x 16 KB
    unsigned char* sB = smem + 32768;            // 2 panels x 32 KB
    unsigned char* sAx = smem + 32768 + 65536;            // A-extra [128 x 16] no swizzle: 4 KB
    unsigned char* sBx = sAx + 4096;                      // B-extra [256 x 16] no swizzle: 8 KB
    uint64_t* bar_full = reinterpret_cast<uint64_t*>(smem + 32768 + 65536 + 12288);
    uint64_t* bar_mma = bar_full + 1;
    uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bar_full + 2);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) {
        umma::mbar_init(bar_full, 1);
        umma::mbar_init(bar_mma, 1);
        umma::mbar_fence_init();
    }
    // A-extra: columns 0 and 1 are 1.0, rest 0; B-extra row n: col0 = bias_hi[n], col1 = bias_lo[n]
    for (int r = tid; r < 128; r += 128) {
        uint4 v = make_uint4(0x3F803F80u, 0, 0, 0);  // two bf16 ones
        *reinterpret_cast<uint4*>(sAx + umma::k16_noswizzle_offset(r, 0)) = v;
        *reinterpret_cast<uint4*>(sAx + umma::k16_noswizzle_offset(r, 1)) = make_uint4(0, 0, 0, 0);
    }
    for (int n = tid; n < 256; n += 128) {
        const float bias = 0.37f * (float)(n - 100) + 0.001f * n;
        const __nv_bfloat16 hi = __float2bfloat16(bias);
        const __nv_bfloat16 lo = __float2bfloat16(bias - __bfloat162float(hi));
        uint4 v = make_uint4((uint32_t)__bfloat16_as_ushort(hi) | ((uint32_t)__bfloat16_as_ushort(lo) << 16), 0, 0, 0);
        *reinterpret_cast<uint4*>(sBx + umma::k16_noswizzle_offset(n, 0)) = v;
        *reinterpret_cast<uint4*>(sBx + umma::k16_noswizzle_offset(n, 1)) = make_uint4(0, 0, 0, 0);
    }
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // generic-proxy writes -> async proxy (UMMA)
    if (warp == 0) umma::tmem_alloc(tmem_slot, 256);
    umma::fence_before_sync();
    __syncthreads();
    umma::fence_after_sync();
    const uint32_t tmem = *tmem_slot;
    if (tid == 0) {
        umma::mbar_arrive_expect_tx(bar_full, 32768 + 65536);
        for (int kp = 0; kp < 2; ++kp) {
            umma::tma_load_3d(sA + kp * 16384, &mapA, bar_full, kp * 64, 0, 0);
            umma::tma_load_3d(sB + kp * 32768, &mapB, bar_full, kp * 64, 0, 0);
        }
        umma::mbar_wait(bar_full, 0);
        umma::fence_after_sync();
        const uint32_t idesc = umma::instr_desc_f16(M, N, 1);
        for (int k = 0; k < KD / 16; ++k) {
            const int kp = k / 4, kk = k % 4;
            const uint64_t da = umma::smem_desc_sw128(umma::smem_u32(sA + kp * 16384) + kk * 32);
            const uint64_t db = umma::smem_desc_sw128(umma::smem_u32(sB + kp * 32768) + kk * 32);
            umma::mma_f16_ss(tmem, da, db, idesc, k > 0);
        }
        umma::mma_f16_ss(tmem, umma::smem_desc_k16_noswizzle(umma::smem_u32(sAx)),
                         umma::smem_desc_k16_noswizzle(umma::smem_u32(sBx)), idesc, 1);
        umma::mma_commit(bar_mma);
    }
    __syncthreads();  // also makes the TMA data visible to everyone below (thread 0 waited)
    umma::mbar_wait(bar_mma, 0);
    umma::fence_after_sync();
    // epilogue: thread t = row t
    float dot = 0.f;
    for (int c0 = 0; c0 < N; c0 += 32) {
        uint32_t r[32];
        umma::tmem_ld32(tmem + ((uint32_t)(warp * 32) << 16) + c0, r);
        umma::tmem_ld_wait();
        for (int j = 0; j < 32; ++j) D[(size_t)tid * N + c0 + j] = __uint_as_float(r[j]);
    }
    // generic read-back of the swizzled A panel: rowdot[t] = sum_k A[t][k]
    for (int kp = 0; kp < 2; ++kp)
        for (int ch = 0; ch < 8; ++ch) {
            const uint4 v = *reinterpret_cast<const uint4*>(sA + kp * 16384 + umma::sw128_offset(tid, ch));
            const uint32_t w[4] = {v.x, v.y, v.z, v.w};
            for (int j = 0; j < 4; ++j) dot += __uint_as_float(w[j] << 16) + __uint_as_float(w[j] & 0xFFFF0000u);
        }
    rowdot[tid] = dot;
    umma::fence_before_sync();
    __syncthreads();
    if (warp == 0) umma::tmem_dealloc(tmem, 256);
}

int main() {
    std::vector<__nv_bfloat16> hA(M * KD), hB(N * KD);
    std::vector<float> fA(M * KD), fB(N * KD);
    srand(1);
    for (int i = 0; i < M * KD; ++i) { float v = (rand() % 2001 - 1000) / 500.f; hA[i] = __float2bfloat16(v); fA[i] = __bfloat162float(hA[i]); }
    for (int i = 0; i < N * KD; ++i) { float v = (rand() % 2001 - 1000) / 500.f; hB[i] = __float2bfloat16(v); fB[i] = __bfloat162float(hB[i]); }
    // embed A in a larger strided buffer to exercise strides: rows 320 B apart (stride 160 elements)
    const int lda = 160;
    std::vector<__nv_bfloat16> hA2((size_t)M * lda, __float2bfloat16(0.f));
    for (int m = 0; m < M; ++m) for (int k = 0; k < KD; ++k) hA2[(size_t)m * lda + k] = hA[m * KD + k];
    __nv_bfloat16 *dA, *dB; float *dD, *dR;
    CK(cudaMalloc(&dA, hA2.size() * 2)); CK(cudaMalloc(&dB, hB.size() * 2)); CK(cudaMalloc(&dD, M * N * 4)); CK(cudaMalloc(&dR, M * 4));
    CK(cudaMemcpy(dA, hA2.data(), hA2.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB, hB.data(), hB.size() * 2, cudaMemcpyHostToDevice));
    CUtensorMap mapA, mapB;
    { uint64_t dims[3] = {KD, M, 1}, str[3] = {0, (uint64_t)lda * 2, (uint64_t)M * lda * 2}; uint32_t box[3] = {64, 128, 1};
      CK(make_tmap_16bit(&mapA, dA, 3, dims, str, box)); }
    { uint64_t dims[3] = {KD, N, 1}, str[3] = {0, KD * 2, (uint64_t)N * KD * 2}; uint32_t box[3] = {64, 256, 1};
      CK(make_tmap_16bit(&mapB, dB, 3, dims, str, box)); }
    const int smem = 32768 + 65536 + 12288 + 64;
    CK(cudaFuncSetAttribute(smoke_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem + 1024));
    smoke_kernel<<<1, 128, smem + 1024>>>(mapA, mapB, dD, dR);
    CK(cudaDeviceSynchronize());
    std::vector<float> hD(M * N), hR(M);
    CK(cudaMemcpy(hD.data(), dD, M * N * 4, cudaMemcpyDeviceToHost));
    CK(cudaMemcpy(hR.data(), dR, M * 4, cudaMemcpyDeviceToHost));
    double max_err = 0, max_ref = 0; int bad = 0;
    for (int m = 0; m < M; ++m) for (int n = 0; n < N; ++n) {
        double ref = 0.37f * (float)(n - 100) + 0.001f * n; for (int k = 0; k < KD; ++k) ref += (double)fA[m * KD + k] * fB[n * KD + k];
        double err = fabs(ref - hD[m * N + n]); if (err > max_err) max_err = err; if (fabs(ref) > max_ref) max_ref = fabs(ref);
        if (err > 2e-3 && bad < 5) { printf("mismatch m=%d n=%d ref=%f got=%f\n", m, n, ref, hD[m * N + n]); ++bad; }
    }
    double max_rerr = 0;
    for (int m = 0; m < M; ++m) { double ref = 0; for (int k = 0; k < KD; ++k) ref += fA[m * KD + k]; max_rerr = fmax(max_rerr, fabs(ref - hR[m])); }
    printf("UMMA smoke: max |err| = %g (max |ref| = %g), swizzled row read-back max err = %g -> %s\n", max_err, max_ref, max_rerr,
           (max_err < 2e-3 && max_rerr < 1e-3) ? "PASS" : "FAIL");
    return (max_err < 2e-3 && max_rerr < 1e-3) ? 0 : 1;
}
