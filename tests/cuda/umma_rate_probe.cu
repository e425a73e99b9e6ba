// Micro-benchmark (test tooling, not product): tensor-pipe cost of one tcgen05.mma as a function of kind and N, M = 128,
// operands resident in shared memory (SWIZZLE_128B K-major, A with the conv kernel's SBO = 18 * 128 B patch view).
// One CTA per SM, one or two issuing warps, `reps` MMAs each, timed with clock64 around issue + commit + wait.
//   usage: umma_rate_probe
#include <cstdio>
#include <cstdlib>
#include <cuda_fp16.h>
#include "../../fisr_b200/csrc/sm100_ptx.cuh"

using namespace fisr;
#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

// mode: 0 = f16 only, 1 = f8 only, 2 = tap pattern 4 x f16 + 4 x f8
__global__ void __launch_bounds__(128, 1) rate_kernel(int n, int mode, int reps, int issuers, int alt, long long* out) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sA = smem_u32(smem), sB = sA + 2 * 42 * 1024;         // two 41.5 KB patches, then B (256 rows x 128 B)
    __shared__ __align__(8) uint64_t bar[4];
    __shared__ uint32_t tmem_slot;
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { for (int i = 0; i < 4; ++i) mbar_init(smem_u32(&bar[i]), 1); fence_mbar_init(); }
    for (int i = tid; i < (2 * 42 * 1024 + 256 * 128) / 4; i += 128) reinterpret_cast<uint32_t*>(smem)[i] = 0;
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 512);
    fence_proxy_async();
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tm = tmem_slot;
    long long t0 = 0, t1 = 0;
    if (warp < issuers && (tid & 31) == 0) {
        const uint32_t a_hi = umma_desc_hi_sw128(18 * 128), b_hi = kUmmaDescHiSw128;
        const uint32_t i16 = umma_idesc_f16(128, n), i8 = umma_idesc_f8(128, n, kF8E5M2, kF8E4M3);
        const uint32_t d0 = tm + (warp & 1) * 256;      // alt = 1: this thread alternates between two accumulators (d0, d0 + 128... only for n <= 128)
        t0 = clock64();
        for (int r = 0; r < reps; ++r) {
            const uint32_t a0 = umma_desc_lo(sA + (warp * 8 + (r % 9)) * 128) , b0 = umma_desc_lo(sB);
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 0 || mode == 2) umma_f16_lohi2(d0 + ((alt && (k & 1)) ? 128 : 0), a0 + 2 * k, a_hi, b0 + 2 * k, b_hi, i16, 1u);
            }
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                if (mode == 1 || mode == 2) umma_f8_lohi2(d0 + ((alt && (k & 1)) ? 128 : 0), a0 + 2 * k + (42 * 1024 >> 4), a_hi, b0 + 2 * k, b_hi, i8, 1u);
            }
        }
        umma_commit(smem_u32(&bar[warp]));
        mbar_wait(smem_u32(&bar[warp]), 0, nullptr, 0);
        t1 = clock64();
        out[blockIdx.x * 4 + warp] = t1 - t0;
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tm, 512);
}

int main() {
    CK(cudaSetDevice(0));
    int sms = 0;
    CK(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, 0));
    long long* d_out; CK(cudaMalloc(&d_out, sms * 4 * sizeof(long long)));
    const int smem = 1024 + 2 * 42 * 1024 + 256 * 128;
    CK(cudaFuncSetAttribute(rate_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    const int reps = 2000;
    printf("# cycles per MMA instruction (M = 128, K = 16 fp16 / 32 fp8), %d SMs busy, clock64 on the issuing SM\n", sms);
    const char* names[3] = {"f16", "f8", "4f16+4f8"};
    printf("%-10s %5s %8s %4s %14s\n", "mode", "N", "issuers", "alt", "cyc/MMA");
    for (int alt = 0; alt <= 1; ++alt)
        for (int issuers : {1, 2, 4})
            for (int mode = 0; mode < 3; ++mode)
                for (int n : {64, 128}) {
                    if (alt && issuers == 4) continue;
                    CK(cudaMemset(d_out, 0, sms * 4 * sizeof(long long)));
                    rate_kernel<<<sms, 128, smem>>>(n, mode, reps, issuers, alt, d_out);
                    CK(cudaGetLastError());
                    CK(cudaDeviceSynchronize());
                    long long h[4]; CK(cudaMemcpy(h, d_out, sizeof h, cudaMemcpyDeviceToHost));
                    long long mx = 0; for (int i = 0; i < issuers; ++i) mx = h[i] > mx ? h[i] : mx;
                    const int per = (mode == 2 ? 8 : 4) * issuers;
                    printf("%-10s %5d %8d %4d %14.1f\n", names[mode], n, issuers, alt, (double)mx / (reps * (double)per));
                }
    return 0;
}
