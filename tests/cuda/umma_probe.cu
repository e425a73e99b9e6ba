// Bring-up probe (test tooling, not product): checks the tcgen05 / TMA encodings in sm100_ptx.cuh
// on a tiny 3x3 conv [2,16,32,64] -> 64 channels and answers one design question:
// can a UMMA A-descriptor start at an arbitrary 128-B row of a TMA-swizzled patch (mode 1/2),
// or must the start be 1024-B aligned (mode 0 = one TMA box per tap)?
//   usage: umma_probe <mode 0|1|2> <variant bits: 1 = m_dim at bit 23, 2 = LBO field 0>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <vector>
#include <cmath>
#include <cuda_fp16.h>
#include "../../fisr_b200/csrc/sm100_ptx.cuh"

using namespace fisr;

#define CK(x) do { cudaError_t e_ = (x); if (e_ != cudaSuccess) { printf("CUDA error %s at %s:%d\n", cudaGetErrorString(e_), __FILE__, __LINE__); exit(2); } } while (0)

constexpr int C = 64, CO = 64;
constexpr int A_BYTES = 160 * 128;          // up to 10 rows x 16 px x 128 B
constexpr int B_BYTES = 9 * CO * 128;

template <int MODE>
__global__ void __launch_bounds__(128, 1)
probe_kernel(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, float* out, int H, int W,
             int variant, int* err) {
    extern __shared__ __align__(1024) uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    uint8_t* sA = smem;
    uint8_t* sB = smem + A_BYTES + (1024 - A_BYTES % 1024) % 1024;
    __shared__ __align__(8) uint64_t bars[2];
    __shared__ uint32_t tmem_slot;

    constexpr int TWV = (MODE == 0) ? 16 : 14;   // valid output columns per tile
    const int x0 = blockIdx.x * TWV, y0 = blockIdx.y * 8, n = blockIdx.z;
    const uint32_t bar_tma = smem_u32(&bars[0]), bar_mma = smem_u32(&bars[1]);
    const int tid = threadIdx.x, warp = tid >> 5;

    if (tid == 0) {
        mbar_init(bar_tma, 1);
        mbar_init(bar_mma, 1);
        fence_mbar_init();
    }
    if (warp == 0) tmem_alloc(smem_u32(&tmem_slot), 64);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    uint32_t idesc = umma_idesc_f16(128, CO);
    if (variant & 1) idesc = (1u << 4) | ((CO >> 3) << 17) | ((128u >> 4) << 23);
    const uint64_t lbo_fix = (variant & 2) ? ~(uint64_t(1) << 16) : ~uint64_t(0);

    if (tid == 0) {
        uint32_t ph_tma = 0, ph_mma = 0;
        bool ok = true;
        if (MODE == 0) {
            mbar_expect_tx(bar_tma, B_BYTES);
            for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(sB) + t * CO * 128, &tmB, bar_tma, 0, t * CO);
            ok = mbar_wait(bar_tma, ph_tma, err, 1); ph_tma ^= 1;
            for (int t = 0; t < 9 && ok; ++t) {
                const int ky = t / 3, kx = t % 3;
                mbar_expect_tx(bar_tma, 128 * 128);
                tma_load_4d(smem_u32(sA), &tmA, bar_tma, 0, x0 + kx - 1, y0 + ky - 1, n);
                ok = mbar_wait(bar_tma, ph_tma, err, 2); ph_tma ^= 1;
                if (!ok) break;
                tc_fence_after();
                for (int k = 0; k < 4; ++k) {
                    uint64_t ad = umma_smem_desc_sw128(smem_u32(sA) + k * 32, 1024) & lbo_fix;
                    uint64_t bd = umma_smem_desc_sw128(smem_u32(sB) + t * CO * 128 + k * 32, 1024) & lbo_fix;
                    umma_f16(tmem_base, ad, bd, idesc, (t | k) ? 1u : 0u);
                }
                umma_commit(bar_mma);
                ok = mbar_wait(bar_mma, ph_mma, err, 3); ph_mma ^= 1;
            }
        } else {
            mbar_expect_tx(bar_tma, B_BYTES + A_BYTES);
            for (int t = 0; t < 9; ++t) tma_load_2d(smem_u32(sB) + t * CO * 128, &tmB, bar_tma, 0, t * CO);
            tma_load_4d(smem_u32(sA), &tmA, bar_tma, 0, x0 - 1, y0 - 1, n);      // box {64, 16, 10, 1}
            ok = mbar_wait(bar_tma, ph_tma, err, 1); ph_tma ^= 1;
            if (ok) {
                tc_fence_after();
                for (int t = 0; t < 9; ++t) {
                    const int ky = t / 3, kx = t % 3;
                    const uint32_t a0 = smem_u32(sA) + (ky * 16 + kx) * 128;
                    const uint32_t bo = (MODE == 2) ? ((a0 >> 7) & 7) : 0;
                    for (int k = 0; k < 4; ++k) {
                        uint64_t ad = umma_smem_desc_sw128(a0 + k * 32, 1024, bo) & lbo_fix;
                        uint64_t bd = umma_smem_desc_sw128(smem_u32(sB) + t * CO * 128 + k * 32, 1024) & lbo_fix;
                        umma_f16(tmem_base, ad, bd, idesc, (t | k) ? 1u : 0u);
                    }
                }
                umma_commit(bar_mma);
                ok = mbar_wait(bar_mma, ph_mma, err, 3); ph_mma ^= 1;
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    tc_fence_after();

    // epilogue: thread = accumulator row m
    const int m = tid;
    const int ty = m / 16, tx = m % 16;
    const int y = y0 + ty, x = x0 + tx;
    for (int c0 = 0; c0 < CO; c0 += 32) {
        uint32_t v[32];
        tmem_ld_32x32(tmem_base + ((uint32_t)(warp * 32) << 16) + c0, v);
        tmem_ld_wait();
        if (tx < TWV && y < H && x < W) {
            float* o = out + (((size_t)n * H + y) * W + x) * CO + c0;
            for (int j = 0; j < 32; ++j) o[j] = __uint_as_float(v[j]);
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 0) tmem_dealloc(tmem_base, 64);
}

typedef CUresult (*EncodeFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                             const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                             CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int main(int argc, char** argv) {
    const int mode = argc > 1 ? atoi(argv[1]) : 0;
    const int variant = argc > 2 ? atoi(argv[2]) : 0;
    const int N = 2, H = 16, W = 32;
    CK(cudaSetDevice(0));
    CK(cudaFree(0));
    EncodeFn encode = nullptr;
    cudaDriverEntryPointQueryResult qres;
    CK(cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", (void**)&encode, cudaEnableDefault, &qres));
    if (!encode) { printf("no cuTensorMapEncodeTiled\n"); return 2; }

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
        cuuint32_t box[4] = {C, 16, (cuuint32_t)(mode == 0 ? 8 : 10), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        CUresult r = encode(&tmA, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, dx, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode A failed %d\n", (int)r); return 2; }
    }
    {
        cuuint64_t dims[2] = {C, (cuuint64_t)9 * CO};
        cuuint64_t strides[1] = {C * 2};
        cuuint32_t box[2] = {C, CO};
        cuuint32_t es[2] = {1, 1};
        CUresult r = encode(&tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, dw, dims, strides, box, es,
                            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_128B,
                            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) { printf("encode B failed %d\n", (int)r); return 2; }
    }
    const int twv = mode == 0 ? 16 : 14;
    dim3 grid((W + twv - 1) / twv, H / 8, N);
    const int smem = 1024 + 20480 + B_BYTES + 1024;
    if (mode == 0) {
        CK(cudaFuncSetAttribute(probe_kernel<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<0><<<grid, 128, smem>>>(tmA, tmB, dout, H, W, variant, derr);
    } else if (mode == 1) {
        CK(cudaFuncSetAttribute(probe_kernel<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<1><<<grid, 128, smem>>>(tmA, tmB, dout, H, W, variant, derr);
    } else {
        CK(cudaFuncSetAttribute(probe_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem));
        probe_kernel<2><<<grid, 128, smem>>>(tmA, tmB, dout, H, W, variant, derr);
    }
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
    printf("PROBE mode=%d variant=%d err_flag=%d max_abs_err=%.3e bad=%zu/%zu -> %s\n", mode, variant, herr, maxerr, bad,
           ho.size(), (herr == 0 && bad == 0) ? "PASS" : "FAIL");
    return (herr == 0 && bad == 0) ? 0 : 1;
}
