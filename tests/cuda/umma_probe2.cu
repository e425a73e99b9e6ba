// Bring-up probe 2 (test tooling, not product).  Two hardware questions behind the r01 conv-kernel redesign:
//   test "sbo" : may the 8-row groups of a SWIZZLE_128B K-major A operand be P*128 B apart (SBO != 1024, group starts
//                not 1024-B aligned)?  If yes, an 8-px-wide x 16-row output chunk reads its taps straight out of a
//                halo'd patch of pitch P = 10 without computing the two halo columns (MMA row efficiency 1.0).
//   test "f8"  : kind::f8f6f4 with mixed e5m2 / e4m3 operand formats accumulating into the SAME TMEM accumulator as a
//                kind::f16 MMA, 8-bit operands K-concatenated in one 128-B row ([64 B fmt X | 64 B fmt Y]).
//   usage: umma_probe2 sbo | f8
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include "../../fisr_b200/csrc/sm100_ptx.cuh"

using namespace fisr;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int C = 64, CO = 64;

__device__ __forceinline__ uint64_t desc_sbo(uint32_t addr, uint32_t sbo) { return umma_smem_desc_sw128(addr, sbo, 0); }

__device__ __forceinline__ void umma_f8(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc, uint32_t accumulate) {
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "setp.ne.b32 p, %4, 0;\n\t"
        "tcgen05.mma.cta_group::1.kind::f8f6f4 [%0], %1, %2, %3, p;\n\t}\n"
        ::"r"(tmem_d), "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
        : "memory");
}
__host__ __device__ constexpr uint32_t idesc_f8(uint32_t m, uint32_t n, uint32_t afmt, uint32_t bfmt) {
    return (1u << 4) | (afmt << 7) | (bfmt << 10) | ((n >> 3) << 17) | ((m >> 4) << 24);
}

// ------------------------------------------------------------------ test sbo: 3x3 conv, tile 8 wide x 16 tall, patch 10 x 18
constexpr int P = 10, TH = 16, TW = 8;
constexpr int A_BYTES = P * (TH + 2) * 128;            // 23040
constexpr int A_ALLOC = (A_BYTES + 1023) / 1024 * 1024;
constexpr int B_BYTES = 9 * CO * 128;

__global__ void __launch_bounds__(128, 1)
sbo_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int H, int W, int* err, int a_shift) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem + a_shift;          // a_shift = 128..896: is the TMA / UMMA swizzle a function of the absolute address?
    uint8_t* sB = smem + A_ALLOC + 1024;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const int x0 = blockIdx.x * TW, y0 = blockIdx.y * TH, n = blockIdx.z;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    const uint32_t idesc = umma_idesc_f16(128, CO);
    if (tid == 0) {
        mbar_expect_tx(bar_tma, B_BYTES + A_BYTES);
        for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(sB) + t * CO * 128, &tmB, bar_tma, 0, t * CO);
        tma_load_4d(smem_u32(sA), &tmA, bar_tma, 0, x0 - 1, y0 - 1, n);      // box {64, 10, 18, 1}
        if (mbar_wait(bar_tma, 0, err, 1)) {
            tc_fence_after();
            for (int t = 0; t < 9; ++t) {
                const int ky = t / 3, kx = t % 3;
                const uint32_t a0 = smem_u32(sA) + (ky * P + kx) * 128;
                for (int k = 0; k < 4; ++k)
                    umma_f16(tmem_base, desc_sbo(a0 + k * 32, P * 128), desc_sbo(smem_u32(sB) + t * CO * 128 + k * 32, 1024), idesc,
                             (t | k) ? 1u : 0u);
            }
            umma_commit(bar_mma);
            mbar_wait(bar_mma, 0, err, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const int m = tid, ty = m >> 3, tx = m & 7;
    const int y = y0 + ty, x = x0 + tx;
    for (int c0 = 0; c0 < CO; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (y < H && x < W) {
            float* o = out + (((size_t)n * H + y) * W + x) * CO + c0;
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

// ------------------------------------------------------------------ test f8: D[128,64] = A16 B16^T + A8 B8^T (mixed formats)
__global__ void __launch_bounds__(128, 1)
f8_kernel(const __grid_constant__ CUtensorMap tmA16, const __grid_constant__ CUtensorMap tmA8, const __grid_constant__ CUtensorMap tmB16,
          const __grid_constant__ CUtensorMap tmB8, float* out, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t sA16 = smem_u32(smem), sA8 = sA16 + 128 * 128, sB16 = sA8 + 128 * 128, sB8 = sB16 + 64 * 128;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5;
    if (tid == 0) { mbar_init(bar_tma, 1); mbar_init(bar_mma, 1); fence_mbar_init(); }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    if (tid == 0) {
        mbar_expect_tx(bar_tma, 2 * 128 * 128 + 2 * 64 * 128);
        tma_load_2d(sA16, &tmA16, bar_tma, 0, 0);
        tma_load_2d(sA8, &tmA8, bar_tma, 0, 0);
        tma_load_2d(sB16, &tmB16, bar_tma, 0, 0);
        tma_load_2d(sB8, &tmB8, bar_tma, 0, 0);
        if (mbar_wait(bar_tma, 0, err, 1)) {
            tc_fence_after();
            for (int k = 0; k < 4; ++k)
                umma_f16(tmem_base, desc_sbo(sA16 + k * 32, 1024), desc_sbo(sB16 + k * 32, 1024), umma_idesc_f16(128, 64), k ? 1u : 0u);
            // bytes [0,64): A e5m2 x B e4m3 ; bytes [64,128): A e4m3 x B e5m2   (K = 32 per instruction)
            for (int k = 0; k < 4; ++k)
                umma_f8(tmem_base, desc_sbo(sA8 + k * 32, 1024), desc_sbo(sB8 + k * 32, 1024),
                        k < 2 ? idesc_f8(128, 64, 1, 0) : idesc_f8(128, 64, 0, 1), 1u);
            umma_commit(bar_mma);
            mbar_wait(bar_mma, 0, err, 3);
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    for (int c0 = 0; c0 < 64; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        for (int j = 0; j < 32; ++j) out[tid * 64 + c0 + j] = __uint_as_float(v[j]);
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
static EncodeFn g_encode = nullptr;

static bool enc2d(CUtensorMap* tm, void* base, int rows) {   // rows x 128 B
    cuuint64_t dims[2] = {64, (cuuint64_t)rows};
    cuuint64_t strides[1] = {128};
    cuuint32_t box[2] = {64, (cuuint32_t)rows};
    cuuint32_t es[2] = {1, 1};
    return g_encode(tm, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, base, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

static float dec8(uint8_t b, bool e5m2) {
    __half_raw h = __nv_cvt_fp8_to_halfraw(b, e5m2 ? __NV_E5M2 : __NV_E4M3);
    __half hh; memcpy(&hh, &h, 2);
    return __half2float(hh);
}

static int test_sbo(int a_shift) {
    const int N = 2, H = 32, W = 24;
    std::vector<__half> hx((size_t)N * H * W * C), hw((size_t)9 * CO * C);
    srand(123);
    for (auto& v : hx) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : hw) v = __float2half((rand() % 2001 - 1000) / 4000.0f);
    __half *dx, *dw; float* dout; int* derr;
    CK(cudaMalloc(&dx, hx.size() * 2)); CK(cudaMalloc(&dw, hw.size() * 2));
    CK(cudaMalloc(&dout, (size_t)N * H * W * CO * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dx, hx.data(), hx.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dw, hw.data(), hw.size() * 2, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, (size_t)N * H * W * CO * 4)); CK(cudaMemset(derr, 0, 4));
    CUtensorMap tmA, tmB;
    {
        cuuint64_t dims[4] = {C, (cuuint64_t)W, (cuuint64_t)H, (cuuint64_t)N};
        cuuint64_t strides[3] = {C * 2, (cuuint64_t)W * C * 2, (cuuint64_t)H * W * C * 2};
        cuuint32_t box[4] = {C, P, TH + 2, 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = g_encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 2; }
    }
    {
        cuuint64_t dims[2] = {C, (cuuint64_t)9 * CO};
        cuuint64_t strides[1] = {C * 2};
        cuuint32_t box[2] = {C, CO};
        cuuint32_t es[2] = {1, 1};
        CUresult r = g_encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                              CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 2; }
    }
    dim3 grid(W / TW, H / TH, N);
    const int smem = 2048 + A_ALLOC + B_BYTES;
    CK(cudaFuncSetAttribute(sbo_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    sbo_kernel<<<grid, 128, smem>>>(tmA, tmB, dout, H, W, derr, a_shift);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    std::vector<float> ho((size_t)N * H * W * CO);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0; size_t bad = 0;
    for (int n = 0; n < N; ++n) for (int y = 0; y < H; ++y) for (int x = 0; x < W; ++x) for (int co = 0; co < CO; ++co) {
        float acc = 0;
        for (int ky = 0; ky < 3; ++ky) for (int kx = 0; kx < 3; ++kx) {
            int yy = y + ky - 1, xx = x + kx - 1;
            if (yy < 0 || yy >= H || xx < 0 || xx >= W) continue;
            const __half* px = &hx[(((size_t)n * H + yy) * W + xx) * C];
            const __half* pw = &hw[((size_t)(ky * 3 + kx) * CO + co) * C];
            for (int ci = 0; ci < C; ++ci) acc += __half2float(px[ci]) * __half2float(pw[ci]);
        }
        float g = ho[(((size_t)n * H + y) * W + x) * CO + co];
        double e = fabs((double)g - acc);
        if (!(e <= 1e-2)) ++bad;
        if (e > maxerr || std::isnan(g)) maxerr = std::isnan(g) ? 1e30 : e;
    }
    printf("PROBE2 sbo (SBO = %d B, unaligned group starts, patch base +%d B) err_flag=%d max_abs_err=%.3e bad=%zu/%zu -> %s\n", P * 128, a_shift, herr, maxerr,
           bad, ho.size(), (herr == 0 && bad == 0) ? "PASS" : "FAIL");
    return (herr == 0 && bad == 0) ? 0 : 1;
}

static int test_f8() {
    std::vector<__half> a16(128 * 64), b16(64 * 64);
    std::vector<uint8_t> a8(128 * 128), b8(64 * 128);
    srand(7);
    for (auto& v : a16) v = __float2half((rand() % 2001 - 1000) / 1000.0f);
    for (auto& v : b16) v = __float2half((rand() % 2001 - 1000) / 4000.0f);
    auto r_e4m3 = []() -> uint8_t { return (uint8_t)(((rand() & 1) << 7) | ((4 + rand() % 6) << 3) | (rand() & 7)); };
    auto r_e5m2 = []() -> uint8_t { return (uint8_t)(((rand() & 1) << 7) | ((10 + rand() % 8) << 2) | (rand() & 3)); };
    for (int r = 0; r < 128; ++r) for (int j = 0; j < 128; ++j) a8[r * 128 + j] = j < 64 ? r_e5m2() : r_e4m3();
    for (int r = 0; r < 64; ++r) for (int j = 0; j < 128; ++j) b8[r * 128 + j] = j < 64 ? r_e4m3() : r_e5m2();
    void *dA16, *dA8, *dB16, *dB8; float* dout; int* derr;
    CK(cudaMalloc(&dA16, 128 * 128)); CK(cudaMalloc(&dA8, 128 * 128)); CK(cudaMalloc(&dB16, 64 * 128)); CK(cudaMalloc(&dB8, 64 * 128));
    CK(cudaMalloc(&dout, 128 * 64 * 4)); CK(cudaMalloc(&derr, 4));
    CK(cudaMemcpy(dA16, a16.data(), 128 * 128, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dA8, a8.data(), 128 * 128, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB16, b16.data(), 64 * 128, cudaMemcpyHostToDevice));
    CK(cudaMemcpy(dB8, b8.data(), 64 * 128, cudaMemcpyHostToDevice));
    CK(cudaMemset(dout, 0xFF, 128 * 64 * 4)); CK(cudaMemset(derr, 0, 4));
    CUtensorMap tA16, tA8, tB16, tB8;
    if (!enc2d(&tA16, dA16, 128) || !enc2d(&tA8, dA8, 128) || !enc2d(&tB16, dB16, 64) || !enc2d(&tB8, dB8, 64)) { printf("encode failed\n"); return 2; }
    const int smem = 1024 + 2 * 128 * 128 + 2 * 64 * 128;
    CK(cudaFuncSetAttribute(f8_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
    f8_kernel<<<1, 128, smem>>>(tA16, tA8, tB16, tB8, dout, derr);
    CK(cudaGetLastError());
    CK(cudaDeviceSynchronize());
    int herr = 0;
    CK(cudaMemcpy(&herr, derr, 4, cudaMemcpyDeviceToHost));
    std::vector<float> ho(128 * 64);
    CK(cudaMemcpy(ho.data(), dout, ho.size() * 4, cudaMemcpyDeviceToHost));
    double maxerr = 0, maxref = 0; size_t bad = 0;
    for (int m = 0; m < 128; ++m) for (int n = 0; n < 64; ++n) {
        double acc = 0;
        for (int k = 0; k < 64; ++k) acc += (double)__half2float(a16[m * 64 + k]) * __half2float(b16[n * 64 + k]);
        for (int j = 0; j < 64; ++j) acc += (double)dec8(a8[m * 128 + j], true) * dec8(b8[n * 128 + j], false);
        for (int j = 64; j < 128; ++j) acc += (double)dec8(a8[m * 128 + j], false) * dec8(b8[n * 128 + j], true);
        const float g = ho[m * 64 + n];
        const double e = fabs(g - acc);
        maxref = fmax(maxref, fabs(acc));
        if (!(e <= 1e-3 * fmax(1.0, fabs(acc)))) ++bad;
        if (e > maxerr || std::isnan(g)) maxerr = std::isnan(g) ? 1e30 : e;
    }
    printf("PROBE2 f8 (f16 + e5m2*e4m3 + e4m3*e5m2 in one accumulator) err_flag=%d max_abs_err=%.3e (max |ref| %.2f) bad=%zu/%zu -> %s\n",
           herr, maxerr, maxref, bad, ho.size(), (herr == 0 && bad == 0) ? "PASS" : "FAIL");
    return (herr == 0 && bad == 0) ? 0 : 1;
}

int main(int argc, char** argv) {
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&g_encode, cudaEnableDefault, &qres));
    if (!g_encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }
    const char* which = argc > 1 ? argv[1] : "sbo";
    if (!strcmp(which, "sbo")) return test_sbo(argc > 2 ? atoi(argv[2]) : 0);
    if (!strcmp(which, "f8")) return test_f8();
    printf("usage: umma_probe2 sbo|f8\n");
    return 2;
}
