// Weight gradient of the 3x3 SAME conv (the backward-filter half of `optim = AdamOptimizer.minimize(total_loss)`,
// FISRnet.py:489-491, for every `Conv2d` of ops.py:7-11) as a tcgen05 GEMM whose K dimension is the pixel axis:
//
//     dW[ky,kx,ci,co] = sum over pixels p of  x[p + (ky-1, kx-1), ci] * dy[p, co]
//
// Both operands are NHWC (hi, lo) fp16 planes, i.e. channel-contiguous = "MN-major" for this GEMM.  TMA drops a
// halo'd x patch (10 x 18 pixels x 64 ch) and the matching dy tile (8 x 16 pixels x 64 ch) into 128B-swizzled
// shared memory exactly as the forward kernel does; a pixel is one 128-byte row, which is the canonical MN-major
// SWIZZLE_128B layout (8 K-rows per swizzle atom, K groups 1024 B apart).  Tap (ky,kx) of tile row ty is the x
// patch read through a descriptor that starts (ty+ky)*18 + kx rows in -- no im2col, no transposes.
//
// Split operands: A = [dy_hi ; dy_lo] stacked along M (two 64-channel chunks, LBO = plane distance), B = x_hi,
// then B = x_lo, all into the same accumulator: lanes 0-63 hold dy_hi * x, lanes 64-127 hold dy_lo * x, and the
// reduction kernel adds the halves -- the full (hi+lo)*(hi+lo) product with two M=128 MMAs per K slice.  Layers with
// many pixels skip the x_lo MMA (and never load that plane): see plan_wgrad.
//
// TMEM: 64 fp32 columns per tap, so a CTA owns taps 0-4 or 5-8 of one (Cout block, Cin block) pair and a share of
// the pixel tiles (split-K); partial sums go to a slot buffer that `wgrad_reduce_kernel` folds in a fixed order
// (deterministic), scaling by 1 / loss-scale.  Warp 0 = TMA producer, warp 1 = MMA issuer, warps 2-5 = epilogue.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "wgrad_umma.h"

namespace fisr {

namespace {

constexpr int kWgThreads = 192;
enum { WERR_EMPTY = 41, WERR_FULL = 42, WERR_ACC = 43 };

// kind::f16 instruction descriptor with both operands MN-major (bits 15, 16), fp16 in, fp32 accumulate
__host__ __device__ constexpr uint32_t wg_idesc(uint32_t m, uint32_t n) { return umma_idesc_f16(m, n) | (1u << 15) | (1u << 16); }

template <bool XLO>
__global__ void __launch_bounds__(kWgThreads, 1)
wgrad3x3_umma_kernel(const __grid_constant__ CUtensorMap tmX_hi, const __grid_constant__ CUtensorMap tmX_lo,
                     const __grid_constant__ CUtensorMap tmD_hi, const __grid_constant__ CUtensorMap tmD_lo,
                     const __grid_constant__ WgradArgs a) {
    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t s0 = smem_u32(smem);
    constexpr int kWgStages = XLO ? kWgStagesXlo : kWgStagesHi;
    constexpr int XPL = XLO ? 2 : 1;                       // x planes staged per tile
    constexpr uint32_t STAGE = XPL * kWgXPlane + 2 * kWgDPlane;
    auto sXhi = [&](int s) { return s0 + s * STAGE; };
    auto sXlo = [&](int s) { return s0 + s * STAGE + kWgXPlane; };
    auto sDhi = [&](int s) { return s0 + s * STAGE + XPL * kWgXPlane; };

    __shared__ __align__(8) uint64_t bars[2 * kWgStages + 1];
    __shared__ uint32_t tmem_slot;
    const uint32_t bar0 = smem_u32(bars);
    auto full = [&](int s) { return bar0 + 8u * s; };
    auto empty = [&](int s) { return bar0 + 8u * (kWgStages + s); };
    const uint32_t acc_full = bar0 + 8u * (2 * kWgStages);

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);

    // work item: block id = split * (OB*CB*2) + (ob * CB + cb) * 2 + tap group
    int item = blockIdx.x;
    const int tg = item & 1; item >>= 1;
    const int cb = item % a.CB; item /= a.CB;
    const int ob = item % a.OB;
    const int split = item / a.OB;
    const int t0 = tg == 0 ? 0 : 5, t1 = tg == 0 ? 5 : 9;

    // CTAs of Cin block 0 / tap group 0 also sum their dy tiles over the pixels (bias gradient) on the epilogue warps,
    // which otherwise idle until the accumulator is complete; those warps then take part in releasing a stage.
    const bool do_bias = a.bias_partial != nullptr && cb == 0 && tg == 0;
    if (tid == 0) {
        for (int s = 0; s < kWgStages; ++s) { mbar_init(full(s), 1); mbar_init(empty(s), do_bias ? 5 : 1); }
        mbar_init(acc_full, 1);
        fence_mbar_init();
        tma_prefetch_desc(&tmX_hi); tma_prefetch_desc(&tmX_lo); tma_prefetch_desc(&tmD_hi); tma_prefetch_desc(&tmD_lo);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), 512);
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    if (warp == 0) {
        if (elect_one()) {
            uint32_t st = 0, ph = 0;
            for (int tile = split; tile < a.num_tiles; tile += a.S) {
                if (!mbar_wait(empty(st), ph ^ 1, a.err, WERR_EMPTY)) break;
                const int tx = tile % a.tiles_x;
                const int r = tile / a.tiles_x;
                const int ty = r % a.tiles_y, n = r / a.tiles_y;
                const int x0 = tx * kWgTW, y0 = ty * kWgTH;
                mbar_expect_tx(full(st), XPL * kWgXBox + 2 * kWgDPlane);
                tma_load_4d(sXhi(st), &tmX_hi, full(st), a.x_coff + cb * 64, x0 - 1, y0 - 1, n);
                if (XLO) tma_load_4d(sXlo(st), &tmX_lo, full(st), a.x_coff + cb * 64, x0 - 1, y0 - 1, n);
                tma_load_4d(sDhi(st), &tmD_hi, full(st), a.dy_coff + ob * 64, x0, y0, n);
                tma_load_4d(sDhi(st) + kWgDPlane, &tmD_lo, full(st), a.dy_coff + ob * 64, x0, y0, n);
                if (++st == kWgStages) { st = 0; ph ^= 1; }
            }
        }
    } else if (warp == 1) {
        constexpr uint32_t idesc = wg_idesc(128, 64);
        const bool lead = elect_one();
        uint32_t st = 0, ph = 0, first = 0;
        bool ok = true;
        // MN-major SWIZZLE_128B descriptor: LBO carries the stride between 64-element M chunks, SBO the stride between 8-row K groups
        const uint32_t lbo_field = a.lbo_a, sbo_field = 1024u >> 4;
        const uint32_t desc_hi = sbo_field | (1u << 14) | (2u << 29);
        for (int tile = split; tile < a.num_tiles && ok; tile += a.S) {
            ok = __all_sync(0xffffffffu, mbar_wait(full(st), ph, a.err, WERR_FULL));
            if (!ok) break;
            tc_fence_after();
            const uint32_t dA = ((sDhi(st) & 0x3FFFF) >> 4) | (lbo_field << 16);
            const uint32_t dBh = ((sXhi(st) & 0x3FFFF) >> 4) | (lbo_field << 16);
            const uint32_t dBl = ((sXlo(st) & 0x3FFFF) >> 4) | (lbo_field << 16);
            if (lead) {
#pragma unroll 1
                for (int t = t0; t < t1; ++t) {
                    const int ky = t / 3, kx = t - 3 * ky;
                    const uint32_t d_tmem = tmem_base + (t - t0) * 64;
#pragma unroll
                    for (int ty = 0; ty < kWgTH; ++ty) {
                        const uint32_t ao = ty * (kWgTW * 128 / 16);                               // 16 rows of 128 B, >> 4
                        const uint32_t bo = static_cast<uint32_t>((ty + ky) * (kWgTW + 2) + kx) * 8u;
                        umma_f16_lohi(d_tmem, dA + ao, dBh + bo, desc_hi, idesc, ty == 0 ? first : 1u);
                        if (XLO) umma_f16_lohi(d_tmem, dA + ao, dBl + bo, desc_hi, idesc, 1u);
                    }
                }
                umma_commit(empty(st));
            }
            first = 1u;
            if (++st == kWgStages) { st = 0; ph ^= 1; }
        }
        if (lead && ok) umma_commit(acc_full);
    } else {
        // epilogue: TMEM lane r (= co + 64 * half) x column (tap, ci) -> partial[2*split + half][tap][ci][co]
        const int q4 = warp & 3;
        const int r = q4 * 32 + lane;
        const int half = r >> 6, co = r & 63;
        const bool have_tiles = split < a.num_tiles;
        bool ok = true;
        if (do_bias) {
            // thread e: 16-byte chunk c (8 channels) of rows rg, rg+16, ..; chunk c of row r sits at c ^ (r & 7) (128B swizzle)
            const int e = tid - 64, c = e & 7, rg = e >> 3;
            const uint32_t off = rg * 128 + ((c ^ (rg & 7)) << 4);
            float acc[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] = 0.f;
            uint32_t st = 0, ph = 0;
            for (int tile = split; tile < a.num_tiles && ok; tile += a.S) {
                ok = __all_sync(0xffffffffu, mbar_wait(full(st), ph, a.err, WERR_FULL));
                if (!ok) break;
#pragma unroll
                for (int pl = 0; pl < 2; ++pl) {
#pragma unroll
                    for (int i = 0; i < 8; ++i) {
                        uint32_t w0, w1, w2, w3;
                        asm volatile("ld.shared.v4.b32 {%0, %1, %2, %3}, [%4];" : "=r"(w0), "=r"(w1), "=r"(w2), "=r"(w3)
                                     : "r"(sDhi(st) + pl * kWgDPlane + off + i * 2048) : "memory");
                        const uint32_t wv[4] = {w0, w1, w2, w3};
#pragma unroll
                        for (int q = 0; q < 4; ++q) {
                            const float2 f = __half22float2(*reinterpret_cast<const __half2*>(&wv[q]));
                            acc[2 * q] += f.x;
                            acc[2 * q + 1] += f.y;
                        }
                    }
                }
                __syncwarp();
                if (lane == 0) mbar_arrive(empty(st));
                if (++st == kWgStages) { st = 0; ph ^= 1; }
            }
            __shared__ float bsum[16][64];
#pragma unroll
            for (int j = 0; j < 8; ++j) bsum[rg][c * 8 + j] = acc[j];
            asm volatile("bar.sync 1, 128;" ::: "memory");
            if (e < 64) {
                float v = 0.f;
#pragma unroll
                for (int g = 0; g < 16; ++g) v += bsum[g][e];
                a.bias_partial[static_cast<size_t>(split) * a.cout_pad + ob * 64 + e] = v;
            }
        }
        if (have_tiles && ok) ok = __all_sync(0xffffffffu, mbar_wait(acc_full, 0, a.err, WERR_ACC));
        tc_fence_after();
        float* dst = a.partial + (static_cast<size_t>(2 * split + half) * 9) * a.cin_pad * a.cout_pad + ob * 64 + co;
        for (int t = t0; t < t1; ++t) {
#pragma unroll 1
            for (int c0 = 0; c0 < 64; c0 += 32) {
                uint32_t v[32];
                if (have_tiles && ok) {
                    tmem_ld_32x32(tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + (t - t0) * 64 + c0, v);
                    tmem_ld_wait();
                } else {
#pragma unroll
                    for (int j = 0; j < 32; ++j) v[j] = 0u;
                }
#pragma unroll
                for (int j = 0; j < 32; ++j)
                    dst[(static_cast<size_t>(t) * a.cin_pad + cb * 64 + c0 + j) * a.cout_pad] = __uint_as_float(v[j]);
            }
        }
    }
    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, 512);
}

__global__ void wgrad_reduce_kernel(const float* __restrict__ partial, int slots, int cin_pad, int cout_pad, int cin, int cout,
                                    float scale, float* __restrict__ g, const float* __restrict__ bias_partial, int splits,
                                    float* __restrict__ gb) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(9) * cin * cout;
    if (i >= total) {
        const size_t co = i - total;          // trailing threads fold the bias-gradient partials of the splits
        if (gb && co < static_cast<size_t>(cout)) {
            float acc = 0.f;
            for (int s = 0; s < splits; ++s) acc += bias_partial[static_cast<size_t>(s) * cout_pad + co];
            gb[co] = acc * scale;
        }
        return;
    }
    const int co = i % cout;
    size_t r = i / cout;
    const int ci = r % cin;
    const int t = r / cin;
    const size_t slot_stride = static_cast<size_t>(9) * cin_pad * cout_pad;
    const float* p = partial + (static_cast<size_t>(t) * cin_pad + ci) * cout_pad + co;
    float acc = 0.f;
    for (int s = 0; s < slots; ++s) acc += p[s * slot_stride];
    g[i] = acc * scale;
}

}  // namespace

void plan_wgrad(int N, int H, int W, int CB, int OB, int num_sms, bool exact, WgradLaunch* L) {
    WgradArgs& a = L->args;
    a.N = N; a.H = H; a.W = W; a.CB = CB; a.OB = OB;
    a.tiles_x = (W + kWgTW - 1) / kWgTW;
    a.tiles_y = (H + kWgTH - 1) / kWgTH;
    a.num_tiles = N * a.tiles_x * a.tiles_y;
    const int per_split = CB * OB * 2;
    int S = num_sms / per_split;
    if (S < 1) S = 1;
    if (S > a.num_tiles) S = a.num_tiles;
    a.S = S;
    a.cin_pad = CB * 64; a.cout_pad = OB * 64;
    a.lbo_a = kWgDPlane >> 4;
    // x's lo plane: leaving it out rounds the forward activation to fp16 inside this product only, a zero-mean term of
    // relative size <= 2^-12 per summand (measured: ~1e-4 of a weight-gradient tensor, the level the tensor core's
    // truncating fp32 accumulation already sets, tools/accum_probe.py) for half the MMAs and 30 % less smem traffic.
    // dy keeps both planes (gradients span many octaves).  Small layers cost nothing: they keep the exact product.
    L->x_lo = exact || static_cast<long long>(N) * H * W < kWgXloMaxPixels;
    L->grid = per_split * S;
    L->bias_offset = static_cast<size_t>(2 * S) * 9 * a.cin_pad * a.cout_pad;
    L->partial_floats = L->bias_offset + static_cast<size_t>(S) * a.cout_pad;
}

cudaError_t wgrad3x3_init() {
    cudaError_t e = cudaFuncSetAttribute(wgrad3x3_umma_kernel<true>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemXlo);
    if (e == cudaSuccess)
        e = cudaFuncSetAttribute(wgrad3x3_umma_kernel<false>, cudaFuncAttributeMaxDynamicSharedMemorySize, kWgSmemHi);
    return e;
}

cudaError_t launch_wgrad3x3(const WgradLaunch& L, cudaStream_t stream) {
    if (L.x_lo)
        wgrad3x3_umma_kernel<true><<<L.grid, kWgThreads, kWgSmemXlo, stream>>>(L.tmX_hi, L.tmX_lo, L.tmD_hi, L.tmD_lo, L.args);
    else
        wgrad3x3_umma_kernel<false><<<L.grid, kWgThreads, kWgSmemHi, stream>>>(L.tmX_hi, L.tmX_lo, L.tmD_hi, L.tmD_lo, L.args);
    return cudaGetLastError();
}

void launch_wgrad_reduce(const WgradLaunch& L, const float* partial, int cin, int cout, float scale, float* g, float* gb,
                         cudaStream_t st) {
    const size_t total = static_cast<size_t>(9) * cin * cout + (gb ? cout : 0);
    wgrad_reduce_kernel<<<static_cast<unsigned>((total + 255) / 256), 256, 0, st>>>(
        partial, 2 * L.args.S, L.args.cin_pad, L.args.cout_pad, cin, cout, scale, g, partial + L.bias_offset, L.args.S, gb);
}

}  // namespace fisr
