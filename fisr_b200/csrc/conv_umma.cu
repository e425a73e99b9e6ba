// 3x3 stride-1 SAME convolution (+bias, +residual, ReLU, depth-to-space, virtual concat) as a
// persistent, warp-specialised tcgen05 implicit GEMM for sm_100a.
//
// Replaces the reference's `Conv2d` (ops.py:7-11) together with the element-wise ops the reference
// runs around it: `relu` (ops.py:17-18), the residual add of `res_block` (ops.py:43),
// `tf.depth_to_space` (FISRnet.py:99,105), `tf.concat` (ops.py:71, FISRnet.py:108,113,144).
//
// Data layout in HBM
//   activations : NHWC fp16, channel count padded to a multiple of 64, as a (hi, lo) plane pair
//                 (x = hi + lo, see common.cuh) -- 4 B / element like the fp32 the reference stores.
//   weights     : [plane][kb][tap][cout_pad][64] fp16 (K-major rows of 128 B), hi/lo planes.
//   residual / pre-activation outputs : NHWC fp32.
//
// One CTA tile = TH x TW output pixels x NT output channels.  Per 64-channel K block the producer
// TMA-loads ONE halo'd input patch (TH+2) x (TW+2) x 64ch (4-D box, out-of-bounds = SAME zero padding)
// into 128B-swizzled shared memory with pixel pitch P = TW+2.  GEMM row m of the tile is patch position
// m (row-major with pitch P), so the A operand of tap (ky,kx) is the same patch read through a UMMA
// descriptor whose start address is advanced by (ky*P+kx) rows of 128 B: no im2col re-read, every
// activation byte crosses L2->SMEM once per tile (plus halo).  Rows with (m % P) >= TW are computed
// and discarded.  Weights stream per (tap, plane) through a ring of [NT x 64] slots.
//
// Warp roles (192 threads): warp 0 = TMA producer, warp 1 = MMA issuer + TMEM owner, warps 2-5 =
// epilogue (TMEM -> registers -> bias/residual/ReLU/split -> HBM).  Two TMEM accumulator stages let
// the epilogue of tile i overlap the main loop of tile i+1.
#include "common.cuh"
#include "sm100_ptx.cuh"
#include "conv_umma.h"

namespace fisr {

namespace {

constexpr int kThreads = 192;
constexpr int kMaxBSlots = 12;
constexpr int kMaxAStages = 2;

// error codes written to ConvArgs::err on a barrier timeout
enum { ERR_A_EMPTY = 11, ERR_B_EMPTY = 12, ERR_A_FULL = 21, ERR_B_FULL = 22, ERR_ACC_EMPTY = 23, ERR_ACC_FULL = 31 };

struct TileCoord {
    int n, y0, x0, nb;
};
__device__ __forceinline__ TileCoord decode_tile(const ConvArgs& a, int tile) {
    TileCoord t;
    t.nb = tile % a.NB;
    int s = tile / a.NB;
    const int tx = s % a.tiles_x;
    s /= a.tiles_x;
    const int ty = s % a.tiles_y;
    t.n = s / a.tiles_y;
    t.y0 = ty * a.TH;
    t.x0 = tx * a.TW;
    return t;
}

__device__ __forceinline__ uint32_t pack_half2(__half a, __half b) {
    return static_cast<uint32_t>(__half_as_ushort(a)) | (static_cast<uint32_t>(__half_as_ushort(b)) << 16);
}

// Epilogue for CW consecutive output channels [cg, cg+CW) of one output pixel.
template <int CW, int PLANES>
__device__ __forceinline__ void epilogue_store(const ConvArgs& a, const float* __restrict__ sBias,
                                               const uint32_t (&v)[CW], int n, int y, int x, int cg) {
    const size_t pix = (static_cast<size_t>(n) * a.H + y) * a.W + x;
    float f[CW];
#pragma unroll
    for (int j = 0; j < CW; ++j) f[j] = __uint_as_float(v[j]) + sBias[cg + j];

    if (a.scalar_out) {
        // narrow head outputs (Cout = 6 / 3): channel-mapped scalar stores (FISRnet.py:107-108,113,144)
#pragma unroll
        for (int j = 0; j < CW; ++j) {
            const int ch = cg + j;
            if (ch < a.cout) {
                if (a.out_raw) a.out_raw[pix * a.raw_cs + ch + (ch < a.raw_split ? a.raw_off0 : a.raw_off1)] = f[j];
                if (a.out_act) {
                    const float g = a.act_relu ? fmaxf(f[j], 0.f) : f[j];
                    const SplitHalf s = split_f32(g);
                    __half* d = a.out_act + pix * a.act_cs + ch + (ch < a.act_split ? a.act_off0 : a.act_off1);
                    d[0] = s.hi;
                    if (PLANES == 2) d[a.act_plane] = s.lo;
                }
            }
        }
        return;
    }

    if (a.res) {
        const float4* r = reinterpret_cast<const float4*>(a.res + pix * a.res_cs + cg);
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) {
            const float4 t = __ldg(r + j);
            f[4 * j + 0] += t.x; f[4 * j + 1] += t.y; f[4 * j + 2] += t.z; f[4 * j + 3] += t.w;
        }
    }
    if (a.out_raw) {
        float4* o = reinterpret_cast<float4*>(a.out_raw + pix * a.raw_cs + a.raw_off1 + cg);
#pragma unroll
        for (int j = 0; j < CW / 4; ++j) o[j] = make_float4(f[4 * j], f[4 * j + 1], f[4 * j + 2], f[4 * j + 3]);
    }
    if (a.out_act) {
        if (a.act_relu) {
#pragma unroll
            for (int j = 0; j < CW; ++j) f[j] = fmaxf(f[j], 0.f);
        }
        size_t opix;
        int och;
        if (a.act_d2s) {   // tf.depth_to_space(x, 2): out[n, 2y+i, 2x+j, c] = in[n, y, x, (2i+j)*64 + c]
            const int g = cg >> 6;
            opix = (static_cast<size_t>(n) * (2 * a.H) + 2 * y + (g >> 1)) * (2 * a.W) + 2 * x + (g & 1);
            och = cg & 63;
        } else {
            opix = pix;
            och = a.act_off1 + cg;
        }
        __half* d = a.out_act + opix * a.act_cs + och;
#pragma unroll
        for (int j = 0; j < CW / 8; ++j) {
            uint32_t hi[4], lo[4];
#pragma unroll
            for (int q = 0; q < 4; ++q) {
                const SplitHalf s0 = split_f32(f[8 * j + 2 * q]);
                const SplitHalf s1 = split_f32(f[8 * j + 2 * q + 1]);
                hi[q] = pack_half2(s0.hi, s1.hi);
                lo[q] = pack_half2(s0.lo, s1.lo);
            }
            *reinterpret_cast<uint4*>(d + 8 * j) = make_uint4(hi[0], hi[1], hi[2], hi[3]);
            if (PLANES == 2) *reinterpret_cast<uint4*>(d + a.act_plane + 8 * j) = make_uint4(lo[0], lo[1], lo[2], lo[3]);
        }
    }
}

template <int NT, int CHUNKS, int PLANES>
__global__ void __launch_bounds__(kThreads, 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ ConvArgs a) {
    constexpr int ACC_COLS = CHUNKS * NT;                       // fp32 columns of one accumulator stage
    constexpr int TMEM_COLS = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_COLS <= 512, "TMEM columns must be a power of two <= 512");
    constexpr int B_SLOT_BYTES = NT * 128;
    constexpr int CW = (NT >= 32) ? 32 : 16;

    extern __shared__ uint8_t smem_raw[];
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~uintptr_t(1023));
    const uint32_t a_stage_bytes = PLANES * a.a_plane_bytes;
    const uint32_t sA = smem_u32(smem);
    const uint32_t sB = sA + a.a_stages * a_stage_bytes;
    float* sBias = reinterpret_cast<float*>(smem + a.a_stages * a_stage_bytes + a.b_slots * B_SLOT_BYTES);

    __shared__ __align__(8) uint64_t bars[2 * kMaxAStages + 2 * kMaxBSlots + 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (kMaxAStages + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + kMaxBSlots + s); };
    auto acc_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kMaxBSlots + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kMaxBSlots + 2 + s); };

    const int tid = threadIdx.x, warp = tid >> 5, lane = tid & 31;

    if (tid == 0) {
        for (int s = 0; s < a.a_stages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), 1); }
        for (int s = 0; s < a.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), 1); }
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), 1); mbar_init(acc_empty(s), 4); }
        fence_mbar_init();
        tma_prefetch_desc(&tmA_hi);
        if (PLANES == 2) tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS);
    for (int i = tid; i < a.NB * NT; i += kThreads) sBias[i] = a.bias[i];
    tc_fence_before();
    __syncthreads();
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;

    const int box_bytes = (a.TH + 2) * a.P * 128;
    const int cout_pad = a.NB * NT;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if (elect_one()) {
            uint32_t as = 0, aph = 0, bs = 0, bph = 0;
            bool ok = true;
            // A patches are prefetched one (tile, kb) item ahead of the weight stream.
            int pf_tile = blockIdx.x, pf_kb = 0;
            auto issue_a = [&]() -> bool {
                if (pf_tile >= a.num_tiles) return true;
                if (!mbar_wait(a_empty(as), aph ^ 1, a.err, ERR_A_EMPTY)) return false;
                const TileCoord t = decode_tile(a, pf_tile);
                mbar_expect_tx(a_full(as), PLANES * box_bytes);
                const uint32_t dst = sA + as * a_stage_bytes;
                tma_load_4d(dst, &tmA_hi, a_full(as), a.cin_off + pf_kb * 64, t.x0 - 1, t.y0 - 1, t.n);
                if (PLANES == 2)
                    tma_load_4d(dst + a.a_plane_bytes, &tmA_lo, a_full(as), a.cin_off + pf_kb * 64, t.x0 - 1, t.y0 - 1, t.n);
                if (++as == (uint32_t)a.a_stages) { as = 0; aph ^= 1; }
                if (++pf_kb == a.KB) { pf_kb = 0; pf_tile += gridDim.x; }
                return true;
            };
            ok = issue_a();
            for (int tile = blockIdx.x; tile < a.num_tiles && ok; tile += gridDim.x) {
                const int nb = tile % a.NB;
                for (int kb = 0; kb < a.KB && ok; ++kb) {
                    for (int tap = 0; tap < 9 && ok; ++tap) {
                        if (tap == 2 && a.a_stages > 1) ok = issue_a();
                        for (int pl = 0; pl < PLANES && ok; ++pl) {
                            ok = mbar_wait(b_empty(bs), bph ^ 1, a.err, ERR_B_EMPTY);
                            if (!ok) break;
                            mbar_expect_tx(b_full(bs), B_SLOT_BYTES);
                            tma_load_2d(sB + bs * B_SLOT_BYTES, &tmB, b_full(bs), 0,
                                        ((pl * a.KB + kb) * 9 + tap) * cout_pad + nb * NT);
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                    }
                    if (a.a_stages == 1 && ok) ok = issue_a();
                }
            }
        }
    } else if (warp == 1) {
        // ============================== MMA issuer ==============================
        if (elect_one()) {
            constexpr uint32_t idesc = umma_idesc_f16(128, NT);
            const uint64_t desc0 = umma_smem_desc_sw128(0, 1024);
            uint32_t as = 0, aph = 0, bs = 0, bph = 0, cs = 0, cph = 0;
            bool ok = true;
            for (int tile = blockIdx.x; tile < a.num_tiles && ok; tile += gridDim.x) {
                ok = mbar_wait(acc_empty(cs), cph ^ 1, a.err, ERR_ACC_EMPTY);
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + cs * ACC_COLS;
                for (int kb = 0; kb < a.KB && ok; ++kb) {
                    ok = mbar_wait(a_full(as), aph, a.err, ERR_A_FULL);
                    if (!ok) break;
                    const uint32_t a_hi = sA + as * a_stage_bytes;
                    const uint32_t a_lo = a_hi + a.a_plane_bytes;
                    for (int tap = 0; tap < 9 && ok; ++tap) {
                        const uint32_t a_off = ((tap / 3) * a.P + (tap % 3)) * 128;
                        for (int pl = 0; pl < PLANES && ok; ++pl) {
                            ok = mbar_wait(b_full(bs), bph, a.err, ERR_B_FULL);
                            if (!ok) break;
                            tc_fence_after();
                            const uint32_t b_addr = sB + bs * B_SLOT_BYTES;
#pragma unroll
                            for (int c = 0; c < CHUNKS; ++c) {
#pragma unroll
                                for (int k = 0; k < 4; ++k) {
                                    const uint64_t bd = desc0 | ((b_addr + k * 32) >> 4);
                                    const uint32_t ao = a_off + c * (128 * 128) + k * 32;
                                    const uint64_t ahd = desc0 | ((a_hi + ao) >> 4);
                                    if (pl == 0) {
                                        // hi * hi  (+ lo * hi in split mode)
                                        umma_f16(d_tmem + c * NT, ahd, bd, idesc, (kb | tap | k) ? 1u : 0u);
                                        if (PLANES == 2) {
                                            const uint64_t ald = desc0 | ((a_lo + ao) >> 4);
                                            umma_f16(d_tmem + c * NT, ald, bd, idesc, 1u);
                                        }
                                    } else {
                                        umma_f16(d_tmem + c * NT, ahd, bd, idesc, 1u);   // hi * lo
                                    }
                                }
                            }
                            umma_commit(b_empty(bs));
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                    }
                    umma_commit(a_empty(as));
                    if (++as == (uint32_t)a.a_stages) { as = 0; aph ^= 1; }
                }
                umma_commit(acc_full(cs));
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        }
    } else {
        // ============================== epilogue ==============================
        const int q4 = warp & 3;             // TMEM lane quarter this warp may read
        uint32_t cs = 0, cph = 0;
        bool ok = true;
        for (int tile = blockIdx.x; tile < a.num_tiles && ok; tile += gridDim.x) {
            const TileCoord t = decode_tile(a, tile);
            ok = mbar_wait(acc_full(cs), cph, a.err, ERR_ACC_FULL);
            ok = __all_sync(0xffffffffu, ok);
            if (!ok) break;
            tc_fence_after();
#pragma unroll
            for (int c = 0; c < CHUNKS; ++c) {
                const int q = c * 128 + q4 * 32 + lane;
                const int ty = q / a.P, tx = q - ty * a.P;
                const int y = t.y0 + ty, x = t.x0 + tx;
                const bool valid = (tx < a.TW) && (ty < a.TH) && (y < a.H) && (x < a.W);
                const uint32_t taddr = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + cs * ACC_COLS + c * NT;
#pragma unroll 1
                for (int c0 = 0; c0 < NT; c0 += CW) {
                    uint32_t v[CW];
                    if constexpr (CW == 32) tmem_ld_32x32(taddr + c0, v); else tmem_ld_32x16(taddr + c0, v);
                    tmem_ld_wait();
                    if (valid) epilogue_store<CW, PLANES>(a, sBias, v, t.n, y, x, t.nb * NT + c0);
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(acc_empty(cs));
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
    }

    tc_fence_before();
    __syncthreads();
    if (warp == 1) tmem_dealloc(tmem_base, TMEM_COLS);
}

template <int NT, int CHUNKS, int PLANES>
cudaError_t init_inst() {
    return cudaFuncSetAttribute(conv3x3_umma_kernel<NT, CHUNKS, PLANES>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kConvMaxSmem);
}

template <int NT, int CHUNKS, int PLANES>
cudaError_t launch_inst(const ConvLaunch& L, int num_sms, cudaStream_t stream) {
    auto kern = conv3x3_umma_kernel<NT, CHUNKS, PLANES>;
    const int grid = L.args.num_tiles < num_sms ? L.args.num_tiles : num_sms;
    kern<<<grid, kThreads, L.smem_bytes, stream>>>(L.tmA_hi, L.tmA_lo, L.tmB, L.args);
    return cudaGetLastError();
}

}  // namespace

// Opts every instantiation into the large dynamic shared-memory carve-out (once per device, outside graph capture).
cudaError_t conv3x3_init() {
    cudaError_t e = cudaSuccess;
#define FISR_INIT(NT_, CH_)                                                     \
    if (e == cudaSuccess) e = init_inst<NT_, CH_, 1>();                         \
    if (e == cudaSuccess) e = init_inst<NT_, CH_, 2>();
    FISR_INIT(16, 1) FISR_INIT(16, 2) FISR_INIT(64, 1) FISR_INIT(64, 2) FISR_INIT(128, 1) FISR_INIT(128, 2)
#undef FISR_INIT
    return e;
}

cudaError_t launch_conv3x3(const ConvLaunch& L, int num_sms, cudaStream_t stream) {
#define FISR_DISPATCH(NT_, CH_)                                                                 \
    if (L.NT == NT_ && L.chunks == CH_) {                                                       \
        return L.planes == 2 ? launch_inst<NT_, CH_, 2>(L, num_sms, stream)                     \
                             : launch_inst<NT_, CH_, 1>(L, num_sms, stream);                    \
    }
    FISR_DISPATCH(16, 1)
    FISR_DISPATCH(16, 2)
    FISR_DISPATCH(64, 1)
    FISR_DISPATCH(64, 2)
    FISR_DISPATCH(128, 1)
    FISR_DISPATCH(128, 2)
#undef FISR_DISPATCH
    return cudaErrorInvalidValue;
}

// Host-side tile geometry / shared-memory carve-up for one conv launch.
bool plan_conv_geometry(int H, int W, int n_img, int cout_pad, int planes, int num_sms, ConvLaunch* L) {
    int NT = cout_pad <= 16 ? 16 : 64;
    if (cout_pad >= 128) {
        // wide tiles halve the A re-reads when there is enough spatial parallelism to fill the GPU
        const long tiles128 = (long)n_img * ((H * (long)W + 223) / 224) * (cout_pad / 128);
        if (tiles128 >= 2L * num_sms) NT = 128;
    }
    int chunks = 2;
    {
        const long tiles2 = (long)n_img * ((H * (long)W + 223) / 224) * (cout_pad / NT);
        if (tiles2 < num_sms) chunks = 1;
    }
    const int M = chunks * 128;
    const int slot_bytes = NT * 128;
    const int min_slots = planes == 2 ? 4 : 3;
    const int fixed = 1024 /*align*/ + 2048 /*bias*/;
    double best_eff = -1;
    int bestP = 0;
    for (int P = 4; P <= 130 && P <= M; ++P) {
        const int a_plane = (((M + 2 * P + 2) * 128) + 1023) / 1024 * 1024;
        if (fixed + 2 * planes * a_plane + min_slots * slot_bytes > kConvMaxSmem) break;
        const int TW = P - 2, TH = M / P;
        if (TH < 1 || TH + 2 > 256) continue;
        const long tiles = (long)((W + TW - 1) / TW) * ((H + TH - 1) / TH);
        const double eff = (double)H * W / ((double)tiles * M);
        if (eff > best_eff + 1e-9) { best_eff = eff; bestP = P; }
    }
    if (bestP == 0) return false;
    const int P = bestP;
    ConvArgs& a = L->args;
    a.P = P; a.TW = P - 2; a.TH = M / P;
    a.tiles_x = (W + a.TW - 1) / a.TW;
    a.tiles_y = (H + a.TH - 1) / a.TH;
    a.NB = cout_pad / NT;
    a.num_tiles = n_img * a.tiles_x * a.tiles_y * a.NB;
    a.a_plane_bytes = (((M + 2 * P + 2) * 128) + 1023) / 1024 * 1024;
    a.a_stages = 2;
    int slots = (kConvMaxSmem - fixed - 2 * planes * a.a_plane_bytes) / slot_bytes;
    if (slots > kMaxBSlots) slots = kMaxBSlots;
    a.b_slots = slots;
    L->NT = NT; L->chunks = chunks; L->planes = planes;
    L->smem_bytes = fixed + 2 * planes * a.a_plane_bytes + slots * slot_bytes;
    L->efficiency = best_eff;
    return true;
}

}  // namespace fisr
