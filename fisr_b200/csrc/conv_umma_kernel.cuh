// 3x3 stride-1 SAME convolution (+bias, +residual, ReLU, depth-to-space, virtual concat) as a
// persistent, warp-specialised tcgen05 implicit GEMM for sm_100a.  Device template; instantiated per
// (planes, N tile) in conv_inst_*.cu, dispatched from conv_umma.cu.
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
// One CTA tile = TH x TW output pixels x NT output channels, made of CHUNKS chunks of 8 px x 16 rows (= the 128
// rows of one MMA).  Per 64-channel K block the producer TMA-loads ONE halo'd input patch (TH+2) x (TW+2) x 64ch
// (4-D box, out-of-bounds = SAME zero padding) into 128B-swizzled shared memory with pixel pitch P = TW+2.
// The A operand of tap (ky,kx) of a chunk is that patch read through a UMMA descriptor whose start address is
// advanced by (ky*P+kx) rows of 128 B and whose stride between 8-row groups (SBO) is P*128 B: MMA row m is
// pixel (m & 7) of image row (m >> 3) of the chunk, so no halo position is ever multiplied (row efficiency 1 up to
// image-edge tiles), there is no im2col re-read, and every activation byte crosses L2->SMEM once per tile (plus
// halo).  tests/cuda/umma_probe2.cu established on B200 that group starts need not be 1024-B aligned (the 128-B
// swizzle is a function of the absolute shared-memory address for TMA and UMMA alike).  Weights stream per tap
// through a ring of slots.
//
// PLANES = 3 ("f16f8"): activations are (fp16 hi plane, 8-bit plane [e5m2(16 lo) x64 | e5m2(hi) x64] per 64-channel
// block), weights (fp16(128 w) plane, 8-bit plane [e4m3(8 w_hi) | e5m2(128 w_lo)]).  Per tap the main product is 4
// kind::f16 MMAs and both cross terms are 4 kind::f8f6f4 MMAs (K = 32, mixed formats) K-concatenated in the same
// 128-B rows, all into ONE accumulator that holds 128 x the result: 2 MMA units per K slice instead of 3.
//
// STACK (split mode, NT <= 64): the weight slot of a tap holds [B_hi; B_lo] as one 2*NT-row operand, so
//   D[:, 0:NT] = A_hi * B_hi + A_lo * B_hi        D[:, NT:2NT] = A_hi * B_lo
// take 2 MMAs (N = 2NT, then N = NT) instead of 3 and A_hi is fetched from shared memory once for both
// products: an N = 64 MMA needs 48 cycles of shared-memory operand reads for 32 cycles of math (ncu:
// pipe_tc 80 % busy at 50 % tensor math), so the narrow layers are bound by exactly that traffic.
//
// Warp roles (352 threads; 384 for f16f8, whose warp 11 requests the activation patches while warp 0 streams the weights):
// warp 0 = TMA producer, warp 1 = MMA issuer of chunk 0 + TMEM owner, warp 10 = MMA issuer of chunk 1 (one thread can issue a
// kind::f16 MMA only every ~150 cycles, tests/cuda/umma_rate_probe.cu, so every chunk gets its own issuer), warps 2-9 =
// epilogue (TMEM -> registers -> smem transpose -> bias/residual/ReLU/split -> coalesced HBM stores).
// Two TMEM accumulator stages let the epilogue of tile i overlap the main loop of tile i+1.  The epilogue
// is CUDA-core work on every output element; 8 warps (two per TMEM lane quarter) keep the four schedulers
// issuing while the tensor core runs.
#pragma once
#include "common.cuh"
#include "conv_umma.h"
#include "sm100_ptx.cuh"

namespace fisr {
namespace convk {

constexpr int kEpiWarps = 8;
constexpr int kIssuer2Warp = 2 + kEpiWarps;            // second MMA issuer, see "MMA issuers" below
constexpr int kPatchWarp = kIssuer2Warp + 1;           // f16f8 only: activation-patch producer
__host__ __device__ constexpr int conv_threads(int planes) { return 32 * (planes >= 3 ? kPatchWarp + 1 : kIssuer2Warp + 1); }
constexpr int kMaxBSlots = 12;
constexpr int kMaxAStages = 4;                // activation buffers with their own barrier pair (f16f8: plane * 2 + stage)
constexpr int kStageBytesPerWarp = 32 * 64;   // [32 px][16 ch] fp32, 16-B groups XOR-swizzled (conflict-free both ways)
constexpr float kF8AccScale = 1.f / 128.f;    // PLANES = 3: the accumulator holds 128 x the convolution (common.cuh, F8 scales)

// epilogue variants (template flags)
// backward (dgrad) variants: EPI_MASK multiplies the accumulator by [mask > 0] (the ReLU gradient, read from the hi plane
// of the forward activation) before the residual is added; EPI_S2D stores space-to-depth (adjoint of EPI_D2S).
enum { EPI_RES = 1, EPI_RAW = 2, EPI_D2S = 4, EPI_MASK = 8, EPI_S2D = 16 };

// error codes written to ConvArgs::err on a barrier timeout
enum { ERR_A_EMPTY = 11, ERR_B_EMPTY = 12, ERR_A_FULL = 21, ERR_B_FULL = 22, ERR_ACC_EMPTY = 23, ERR_ACC_FULL = 31 };

struct TileCoord {
    int n, y0, x0, nb;
};
// pair_rank < 0: one CTA per tile.  pair_rank = 0 / 1: the work item is a PAIR of x-adjacent tiles (a.tiles_x counts pairs) and
// this CTA takes the left / right one; a right tile past the image edge is all padding (TMA zero fill, epilogue masks it).
__device__ __forceinline__ TileCoord decode_tile(const ConvArgs& a, int tile, int pair_rank = -1) {
    TileCoord t;
    t.nb = tile % a.NB;
    int s = tile / a.NB;
    const int tx = s % a.tiles_x;
    s /= a.tiles_x;
    const int ty = s % a.tiles_y;
    t.n = s / a.tiles_y;
    t.y0 = ty * a.TH;
    t.x0 = (pair_rank < 0 ? tx : 2 * tx + pair_rank) * a.TW;
    return t;
}

__device__ __forceinline__ void sts128(uint32_t addr, float x, float y, float z, float w) {
    asm volatile("st.shared.v4.f32 [%0], {%1, %2, %3, %4};" ::"r"(addr), "f"(x), "f"(y), "f"(z), "f"(w) : "memory");
}
__device__ __forceinline__ float4 lds128(uint32_t addr) {
    float4 v;
    asm volatile("ld.shared.v4.f32 {%0, %1, %2, %3}, [%4];" : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "r"(addr) : "memory");
    return v;
}

// Narrow head outputs (Cout = 6 / 3, FISRnet.py:100,106): thread = pixel, channel-mapped scalar stores into the
// 9-channel pred tensor and into channels 29.. of the next level's input (FISRnet.py:107-108,113,144).
template <int PLANES, int NCOL>
__device__ __forceinline__ void epilogue_scalar(const ConvArgs& a, const uint32_t (&v)[NCOL], int n, int y, int x) {
    const int pix = n * a.opix_n + y * a.opix_y + x * a.opix_x;
    if constexpr (PLANES == 2 && NCOL == 16) {
        // all 16 columns are real channels of an 8-channel-aligned slot: 2 x 16-byte stores per plane instead of 32 scalar ones
        if (a.cout == 16 && !a.out_raw && a.out_act && ((a.act_cs | a.act_off1) & 7) == 0 && a.act_split == 0) {
            uint32_t h[8], l[8];
#pragma unroll
            for (int j = 0; j < 8; ++j) {
                const float2 b2 = __ldg(reinterpret_cast<const float2*>(a.bias) + j);
                float g0 = __uint_as_float(v[2 * j]) + b2.x, g1 = __uint_as_float(v[2 * j + 1]) + b2.y;
                if (a.act_relu) { g0 = fmaxf(g0, 0.f); g1 = fmaxf(g1, 0.f); }
                if (a.act_slope > 0.f) { g0 = fmaxf(g0, g0 * a.act_slope); g1 = fmaxf(g1, g1 * a.act_slope); }
                split2_f32(g0, g1, h[j], l[j]);
            }
            __half* d = a.out_act + static_cast<size_t>(pix) * a.act_cs + a.act_off1;
            *reinterpret_cast<uint4*>(d) = make_uint4(h[0], h[1], h[2], h[3]);
            *reinterpret_cast<uint4*>(d + 8) = make_uint4(h[4], h[5], h[6], h[7]);
            *reinterpret_cast<uint4*>(d + a.act_plane) = make_uint4(l[0], l[1], l[2], l[3]);
            *reinterpret_cast<uint4*>(d + a.act_plane + 8) = make_uint4(l[4], l[5], l[6], l[7]);
            return;
        }
    }
#pragma unroll
    for (int ch = 0; ch < NCOL; ++ch) {
        if (ch < a.cout) {
            float f = __uint_as_float(v[ch]);
            if (PLANES >= 3) f *= kF8AccScale;
            f += __ldg(a.bias + ch);
            const int co = ch, opix = pix;
            if (a.out_raw) a.out_raw[static_cast<size_t>(opix) * a.raw_cs + co + (co < a.raw_split ? a.raw_off0 : a.raw_off1)] = f;
            if (a.out_act) {
                float g = a.act_relu ? fmaxf(f, 0.f) : f;
                if (PLANES == 2 && a.act_slope > 0.f) g = fmaxf(g, g * a.act_slope);
                const SplitHalf s = split_f32(g);
                const int c = co + (co < a.act_split ? a.act_off0 : a.act_off1);
                __half* d = a.out_act + static_cast<size_t>(opix) * a.act_cs + c;
                d[0] = s.hi;
                if (PLANES == 2) d[a.act_plane] = s.lo;
                if (PLANES >= 3) {
                    uint8_t* q = reinterpret_cast<uint8_t*>(a.out_act + a.act_plane + static_cast<size_t>(opix) * a.act_cs) +
                                 (c >> 6) * 128 + (c & 63);
                    q[0] = f8_lo_byte(g - __half2float(s.hi));
                    q[64] = f8_hi_byte(g);
                }
            }
        }
    }
}

// conv/2 head with the depth_to_space folded into its weights (fisr_api.cu, conv_ps): GEMM column sub * COUT + co of input
// pixel (y, x) is channel co of output pixel (2y + (sub >> 1), 2x + (sub & 1)).  The fp32 prediction is stored as 12-float
// records per output pixel, [FI-SR 0..2, -, SR 0..2, -, FI-SR 3..5, -], so that each head writes whole 16-byte groups
// (raw_off0 / raw_off1 = float offset of this head's first / second group).
template <int PLANES, int COUT>
__device__ __forceinline__ void epilogue_scalar_ps(const ConvArgs& a, const uint32_t (&v)[COUT == 6 ? 32 : 16], int n, int y, int x) {
#pragma unroll
    for (int sub = 0; sub < 4; ++sub) {
        const size_t opix = (static_cast<size_t>(n) * 2 * a.H + 2 * y + (sub >> 1)) * (2 * a.W) + 2 * x + (sub & 1);
        float f[COUT];
#pragma unroll
        for (int co = 0; co < COUT; ++co) {
            f[co] = __uint_as_float(v[sub * COUT + co]);
            if (PLANES >= 3) f[co] *= kF8AccScale;
            f[co] += __ldg(a.bias + sub * COUT + co);
        }
        if (a.out_raw) {
            float* rec = a.out_raw + opix * a.raw_cs;
            if (COUT == 6) {
                *reinterpret_cast<float4*>(rec + a.raw_off0) = make_float4(f[0], f[1], f[2], 0.f);
                *reinterpret_cast<float4*>(rec + a.raw_off1) = make_float4(f[3], f[4], f[5], 0.f);
            } else {
                *reinterpret_cast<float4*>(rec + a.raw_off1) = make_float4(f[0], f[1], f[2], 0.f);
            }
        }
        if (a.out_act) {
#pragma unroll
            for (int co = 0; co < COUT; ++co) {
                const float g = a.act_relu ? fmaxf(f[co], 0.f) : f[co];
                const SplitHalf s = split_f32(g);
                const int c = co + (co < a.act_split ? a.act_off0 : a.act_off1);
                __half* d = a.out_act + opix * a.act_cs + c;
                d[0] = s.hi;
                if (PLANES == 2) d[a.act_plane] = s.lo;
                if (PLANES >= 3) {
                    uint8_t* q = reinterpret_cast<uint8_t*>(a.out_act + a.act_plane + opix * a.act_cs) + (c >> 6) * 128 + (c & 63);
                    q[0] = f8_lo_byte(g - __half2float(s.hi));
                    q[64] = f8_hi_byte(g);
                }
            }
        }
    }
}

template <int NT, int CHUNKS, int PLANES, int EPI>
__global__ void __launch_bounds__(conv_threads(PLANES), 1)
conv3x3_umma_kernel(const __grid_constant__ CUtensorMap tmA_hi, const __grid_constant__ CUtensorMap tmA_lo,
                    const __grid_constant__ CUtensorMap tmB, const __grid_constant__ CUtensorMap tmO_hi,
                    const __grid_constant__ CUtensorMap tmO_8, const __grid_constant__ ConvArgs a) {
    constexpr bool STACK = (PLANES == 2) && (NT <= 64);
    constexpr bool F8 = PLANES >= 3;
    constexpr bool PAIR = PLANES == 4;                          // f16f8 on CTA pairs: cluster of 2, tcgen05 cta_group::2, M = 256
    constexpr int APL = PLANES == 1 ? 1 : 2;                    // activation planes staged in shared memory
    constexpr int DCOLS = STACK ? 2 * NT : NT;                  // accumulator columns per 128-row chunk
    constexpr int ACC_COLS = CHUNKS * DCOLS;                    // fp32 columns of one accumulator stage
    constexpr int TMEM_NEED = (2 * ACC_COLS < 32) ? 32 : 2 * ACC_COLS;      // two accumulator stages
    constexpr int TMEM_COLS = TMEM_NEED <= 32 ? 32 : TMEM_NEED <= 64 ? 64 : TMEM_NEED <= 128 ? 128 : TMEM_NEED <= 256 ? 256 : 512;   // allocations are powers of two (N = 96 tiles need 384)
    static_assert((TMEM_COLS & (TMEM_COLS - 1)) == 0 && TMEM_NEED <= 512, "TMEM columns must be a power of two <= 512");
    constexpr int PLANE_BYTES = (PAIR ? NT / 2 : NT) * 128;     // one weight plane of one tap (a CTA of a pair holds half the rows)
    constexpr int B_SLOT_BYTES = STACK ? 2 * PLANE_BYTES : PLANE_BYTES;
    constexpr int SLOTS_PER_TAP = STACK ? 1 : APL;
    constexpr bool NARROW = NT <= 32;
    // tile geometry is a function of CHUNKS alone: two 8 x 16 chunks side by side (16 x 16 tile, patch pitch 18) or one chunk
    constexpr int TWc = 8 * CHUNKS, THc = 16, Pc = TWc + 2;
    constexpr uint32_t CHUNK_OFF = 8u * 128u >> 4;                      // chunk 1 starts 8 pixels to the right of chunk 0
    constexpr uint32_t A_DESC_HI = umma_desc_hi_sw128(Pc * 128u);       // SBO = one patch row

    extern __shared__ uint8_t smem_raw[];
    // 128-B alignment is all the swizzled operands need (absolute-address swizzle): planes and slots are packed tightly
    uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 127) & ~uintptr_t(127));
    const uint32_t a_stage_bytes = APL * a.a_plane_bytes;
    const uint32_t sA = smem_u32(smem);
    const uint32_t sB = sA + a.a_stages * a_stage_bytes;
    const uint32_t sStage32 = sB + a.b_slots * B_SLOT_BYTES;      // kEpiWarps x [32 px][16 ch] fp32 transpose buffers

    __shared__ __align__(8) uint64_t bars[2 * kMaxAStages + 2 * kMaxBSlots + 4];
    __shared__ uint32_t tmem_slot;
    const uint32_t bar0 = smem_u32(bars);
    auto a_full = [&](int s) { return bar0 + 8u * s; };
    auto a_empty = [&](int s) { return bar0 + 8u * (kMaxAStages + s); };
    auto b_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + s); };
    auto b_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + kMaxBSlots + s); };
    auto acc_full = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kMaxBSlots + s); };
    auto acc_empty = [&](int s) { return bar0 + 8u * (2 * kMaxAStages + 2 * kMaxBSlots + 2 + s); };

    const int tid = threadIdx.x, lane = tid & 31;
    const int warp = __shfl_sync(0xffffffffu, tid >> 5, 0);   // broadcast: lets ptxas treat the role branches as warp-uniform

    if (tid == 0) {
        // every chunk has its own MMA issuer warp: operand buffers are released, and accumulators published, by all of them
        for (int s = 0; s < kMaxAStages; ++s) { mbar_init(a_full(s), 1); mbar_init(a_empty(s), CHUNKS); }
        for (int s = 0; s < a.b_slots; ++s) { mbar_init(b_full(s), 1); mbar_init(b_empty(s), CHUNKS); }
        // (pair: the epilogue warps of both CTAs release the leader's accumulator barrier)
        for (int s = 0; s < 2; ++s) { mbar_init(acc_full(s), CHUNKS); mbar_init(acc_empty(s), PAIR ? 2 * kEpiWarps : kEpiWarps); }
        fence_mbar_init();
        tma_prefetch_desc(&tmA_hi);
        if (APL == 2) tma_prefetch_desc(&tmA_lo);
        tma_prefetch_desc(&tmB);
    }
    if (warp == 1) { if (PAIR) tmem_alloc_pair(smem_u32(&tmem_slot), TMEM_COLS); else tmem_alloc(smem_u32(&tmem_slot), TMEM_COLS); }
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();       // pair: both CTAs' barriers exist before anything signals across
    tc_fence_after();
    const uint32_t tmem_base = tmem_slot;
    // work distribution: one item per CTA, or per CTA pair (rank 0 = leader: issues the MMAs, owns the full / accumulator barriers)
    const int prank = PAIR ? static_cast<int>(cluster_ctarank()) : -1;
    const bool leader = !PAIR || prank == 0;
    const int wid0 = PAIR ? blockIdx.x >> 1 : blockIdx.x, wstride = PAIR ? gridDim.x >> 1 : gridDim.x;

    constexpr int box_bytes = (THc + 2) * Pc * 128;
    const int cout_pad = a.NB * NT;

    if (warp == 0) {
        // ============================== TMA producer ==============================
        if constexpr (F8) {
            // f16f8: every (tile, K block) item runs as two phases -- A: the taps of the 8-bit cross terms (8-bit patch, 8-bit
            // weight plane), B: the taps of the fp16 main term.  This warp streams the weight slots; the patches come from
            // the patch-producer warp below.
            if (elect_one()) {
                uint32_t bs = 0, bph = 0;
                bool ok = true;
                for (int tile = wid0; tile < a.num_tiles && ok; tile += wstride) {
                    const int nb = tile % a.NB;
                    for (int kb = 0; kb < a.KB && ok; ++kb) {
                        const uint32_t mask = a.tapmask[kb & 7];
                        for (int sl = 0; sl < 18 && ok; ++sl) {
                            const int plane = sl < 9 ? 1 : 0, tap = sl < 9 ? sl : sl - 9;
                            if (!((mask >> tap) & 1u)) continue;      // all-zero weights (fused depth_to_space heads)
                            ok = mbar_wait(b_empty(bs), bph ^ 1, a.err, ERR_B_EMPTY);
                            if (!ok) break;
                            const int row = ((plane * a.KB + kb) * 9 + tap) * cout_pad + nb * NT;
                            if (PAIR) {      // each CTA loads its half of the tap's rows; both credit the leader's barrier
                                if (leader) mbar_expect_tx(b_full(bs), 2 * B_SLOT_BYTES);
                                tma_load_2d_pair(sB + bs * B_SLOT_BYTES, &tmB, mapa_shared(b_full(bs), 0), 0, row + prank * (NT / 2));
                            } else {
                                mbar_expect_tx(b_full(bs), B_SLOT_BYTES);
                                tma_load_2d(sB + bs * B_SLOT_BYTES, &tmB, b_full(bs), 0, row);
                            }
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                    }
                }
            }
        } else
        if (elect_one()) {
            uint32_t as = 0, aph = 0, bs = 0, bph = 0;
            bool ok = true;
            // A patches are prefetched one (tile, kb) item ahead of the weight stream.
            int pf_tile = blockIdx.x, pf_kb = 0;
            auto issue_a = [&]() -> bool {
                if (pf_tile >= a.num_tiles) return true;
                if (!mbar_wait(a_empty(as), aph ^ 1, a.err, ERR_A_EMPTY)) return false;
                const TileCoord t = decode_tile(a, pf_tile);
                mbar_expect_tx(a_full(as), APL * box_bytes);
                const uint32_t dst = sA + as * a_stage_bytes;
                tma_load_4d(dst, &tmA_hi, a_full(as), a.cin_off + pf_kb * 64, t.x0 - 1, t.y0 - 1, t.n);
                if (APL == 2)
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
                        for (int sl = 0; sl < SLOTS_PER_TAP && ok; ++sl) {
                            ok = mbar_wait(b_empty(bs), bph ^ 1, a.err, ERR_B_EMPTY);
                            if (!ok) break;
                            mbar_expect_tx(b_full(bs), B_SLOT_BYTES);
                            const uint32_t dst = sB + bs * B_SLOT_BYTES;
                            if (STACK) {
                                tma_load_2d(dst, &tmB, b_full(bs), 0, ((0 * a.KB + kb) * 9 + tap) * cout_pad + nb * NT);
                                tma_load_2d(dst + PLANE_BYTES, &tmB, b_full(bs), 0, ((1 * a.KB + kb) * 9 + tap) * cout_pad + nb * NT);
                            } else {
                                tma_load_2d(dst, &tmB, b_full(bs), 0, ((sl * a.KB + kb) * 9 + tap) * cout_pad + nb * NT);
                            }
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                    }
                    if (a.a_stages == 1 && ok) ok = issue_a();
                }
            }
        }
    } else if (F8 && warp == kPatchWarp) {
        // ============================== f16f8 patch producer ==============================
        // Each activation plane sits idle during the other phase of an item, which is when its next patch is loaded: D = 1
        // buffer per plane under the 16 KB weight slots of NT = 128 (the 83 KB this frees hold 8 weight slots instead of 3,
        // enough to cover the L2 latency of the weight stream), D = 2 where shared memory allows.  Buffers drain in the order
        // 8-bit(i), fp16(i), 8-bit(i+1), ..., so one thread with blocking waits requests every patch the moment its buffer is free.
        if (elect_one()) {
            const uint32_t D = a.a_stages;
            const int my_tiles = wid0 < a.num_tiles ? (a.num_tiles - wid0 + wstride - 1) / wstride : 0;
            bool ok = true;
            uint32_t cnt = 0;
            for (int ti = 0; ti < my_tiles && ok; ++ti) {
                const TileCoord t = decode_tile(a, wid0 + ti * wstride, prank);
                for (int kb = 0; kb < a.KB && ok; ++kb, ++cnt) {
                    const uint32_t st = cnt & (D - 1), par = ((cnt >> (D - 1)) & 1) ^ 1;
#pragma unroll
                    for (int pl = 1; pl >= 0 && ok; --pl) {           // 8-bit plane first: phase A consumes it first
                        const uint32_t buf = pl * 2 + st;
                        ok = mbar_wait(a_empty(buf), par, a.err, ERR_A_EMPTY);
                        if (!ok) break;
                        if (PAIR) {
                            if (leader) mbar_expect_tx(a_full(buf), 2 * box_bytes);
                            tma_load_4d_pair(sA + (st * 2 + pl) * a.a_plane_bytes, pl ? &tmA_lo : &tmA_hi, mapa_shared(a_full(buf), 0),
                                             a.cin_off + kb * 64, t.x0 - 1, t.y0 - 1, t.n);
                        } else {
                            mbar_expect_tx(a_full(buf), box_bytes);
                            tma_load_4d(sA + (st * 2 + pl) * a.a_plane_bytes, pl ? &tmA_lo : &tmA_hi, a_full(buf), a.cin_off + kb * 64,
                                        t.x0 - 1, t.y0 - 1, t.n);
                        }
                    }
                }
            }
        }
    } else if ((warp == 1 || (CHUNKS == 2 && warp == kIssuer2Warp)) && leader) {
        // ============================== MMA issuers ==============================
        // One issuer warp per chunk (warp 1: chunk 0, last warp: chunk 1).  A single thread needs ~70 cycles per
        // tcgen05.mma (descriptor arithmetic on the uniform datapath, ncu source view), more than the 32..64 cycles an
        // N = 64 / 128 MMA occupies the tensor pipe; two issuers, each feeding its own accumulator, keep the pipe full and
        // the per-accumulator issue order -- hence the fp32 summation order -- fixed.
        // The whole warp walks the pipeline convergently (waits, stage counters and descriptor words stay in
        // uniform registers); only the tcgen05 instructions are predicated on one elected lane.
        const int c = warp == 1 ? 0 : 1;
        constexpr uint32_t idesc = umma_idesc_f16(128, NT);
        constexpr uint32_t idesc2 = umma_idesc_f16(128, STACK ? 2 * NT : NT);
        constexpr uint32_t MM = PAIR ? 256 : 128;                                // MMA M: a CTA pair issues both CTAs' rows at once
        constexpr uint32_t idesc8a = umma_idesc_f8(MM, NT, kF8E5M2, kF8E4M3);    // bytes [0,64):   e5m2(16 x_lo) * e4m3(8 w_hi)
        constexpr uint32_t idesc8b = umma_idesc_f8(MM, NT, kF8E5M2, kF8E5M2);    // bytes [64,128): e5m2(x_hi) * e5m2(128 w_lo)
        constexpr uint32_t idesc16p = umma_idesc_f16(MM, NT);
        const bool lead = elect_one();
        constexpr uint32_t a_hi_word = A_DESC_HI;
        uint32_t as = 0, aph = 0, bs = 0, bph = 0, cs = 0, cph = 0;
        bool ok = true;
        if constexpr (F8) {
            uint32_t items = 0;                           // (tile, K block) items consumed so far = patches consumed per plane
            const uint32_t D = a.a_stages;
            const uint32_t co = c * CHUNK_OFF;
            auto commit = [&](uint32_t bar) { if (PAIR) umma_commit_pair(bar); else umma_commit(bar); };
            for (int tile = wid0; tile < a.num_tiles && ok; tile += wstride) {
                ok = __all_sync(0xffffffffu, mbar_wait(acc_empty(cs), cph ^ 1, a.err, ERR_ACC_EMPTY));
                if (!ok) break;
                tc_fence_after();
                const uint32_t d_tmem = tmem_base + cs * ACC_COLS + c * DCOLS;
                uint32_t first = 0u;                                          // accumulate flag: 0 for the first MMA of the tile
                for (int kb = 0; kb < a.KB && ok; ++kb) {
                    const int ksteps = kb == a.KB - 1 ? a.ksteps_last : 4;   // 16-channel K slices that hold real channels
                    const int k8steps = (ksteps + 1) >> 1;                   // 32-channel slices of each half of the 8-bit rows
#pragma unroll 1
                    for (int phase = 0; phase < 2 && ok; ++phase) {           // 0: 8-bit cross terms, 1: fp16 main term
                        const int pl = phase == 0 ? 1 : 0;
                        const uint32_t st = items & (D - 1), buf = pl * 2 + st;
                        ok = __all_sync(0xffffffffu, mbar_wait(a_full(buf), (items >> (D - 1)) & 1, a.err, ERR_A_FULL));
                        if (!ok) break;
                        const uint32_t a0 = umma_desc_lo(sA + (st * 2 + pl) * a.a_plane_bytes) + co;
#pragma unroll 1
                        for (int ky = 0; ky < 3 && ok; ++ky) {
#pragma unroll 1
                            for (int kx = 0; kx < 3 && ok; ++kx) {
                                if (!((a.tapmask[kb & 7] >> (ky * 3 + kx)) & 1u)) continue;
                                const uint32_t at = a0 + static_cast<uint32_t>(ky * Pc + kx) * 8u;      // rows of 128 B, >> 4
                                ok = __all_sync(0xffffffffu, mbar_wait(b_full(bs), bph, a.err, ERR_B_FULL));
                                if (!ok) break;
                                tc_fence_after();
                                const uint32_t b0 = umma_desc_lo(sB + bs * B_SLOT_BYTES);
                                if (lead) {
                                    if (phase == 0) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if ((k & 1) >= k8steps) continue;
                                            if (PAIR) umma_f8_pair(d_tmem, at + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128,
                                                                   k < 2 ? idesc8a : idesc8b, k == 0 ? first : 1u);
                                            else umma_f8_lohi2(d_tmem, at + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128,
                                                               k < 2 ? idesc8a : idesc8b, k == 0 ? first : 1u);
                                        }
                                    } else {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (k >= ksteps) break;
                                            if (PAIR) umma_f16_pair(d_tmem, at + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128, idesc16p, 1u);
                                            else umma_f16_lohi2(d_tmem, at + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128, idesc, 1u);
                                        }
                                    }
                                    commit(b_empty(bs));
                                }
                                first = 1u;
                                if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                            }
                        }
                        if (lead && ok) commit(a_empty(buf));
                    }
                    ++items;
                }
                if (lead && ok) commit(acc_full(cs));
                if (++cs == 2) { cs = 0; cph ^= 1; }
            }
        } else {
        for (int tile = blockIdx.x; tile < a.num_tiles && ok; tile += gridDim.x) {
            ok = __all_sync(0xffffffffu, mbar_wait(acc_empty(cs), cph ^ 1, a.err, ERR_ACC_EMPTY));
            if (!ok) break;
            tc_fence_after();
            const uint32_t d_tmem = tmem_base + cs * ACC_COLS;
            for (int kb = 0; kb < a.KB && ok; ++kb) {
                ok = __all_sync(0xffffffffu, mbar_wait(a_full(as), aph, a.err, ERR_A_FULL));
                if (!ok) break;
                const uint32_t a_hi0 = umma_desc_lo(sA + as * a_stage_bytes);
                const uint32_t a_lo0 = umma_desc_lo(sA + as * a_stage_bytes + a.a_plane_bytes);
                uint32_t first = kb == 0 ? 0u : 1u;          // accumulate flag of the very first MMA of the tile
                const int ksteps = kb == a.KB - 1 ? a.ksteps_last : 4;   // 16-channel K slices that hold real channels
                const int k8steps = (ksteps + 1) >> 1;                   // 32-channel slices of each half of the 8-bit rows
#pragma unroll 1
                for (int ky = 0; ky < 3 && ok; ++ky) {
#pragma unroll 1
                    for (int kx = 0; kx < 3 && ok; ++kx) {
                        const uint32_t a_off = static_cast<uint32_t>(ky * Pc + kx) * 8u;      // rows of 128 B, >> 4
                        const uint32_t ah = a_hi0 + a_off, al = a_lo0 + a_off;
                        ok = __all_sync(0xffffffffu, mbar_wait(b_full(bs), bph, a.err, ERR_B_FULL));
                        if (!ok) break;
                        tc_fence_after();
                        {
                            const uint32_t b0 = umma_desc_lo(sB + bs * B_SLOT_BYTES);
                            if (lead) {
                                {
                                    const uint32_t co = c * CHUNK_OFF;
#pragma unroll
                                    for (int k = 0; k < 4; ++k) {
                                        if (k >= ksteps) break;
                                        // A_hi * [B_hi (; B_lo)]
                                        umma_f16_lohi2(d_tmem + c * DCOLS, ah + co + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128,
                                                       idesc2, k == 0 ? first : 1u);
                                        if (PLANES == 2)   // A_lo * B_hi
                                            umma_f16_lohi2(d_tmem + c * DCOLS, al + co + k * 2, a_hi_word, b0 + k * 2,
                                                           kUmmaDescHiSw128, idesc, 1u);
                                    }
                                }
                                umma_commit(b_empty(bs));
                            }
                            first = 1u;
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                        if (APL == 2 && !STACK) {   // second weight plane: A_hi * B_lo (f16x3) or the 8-bit cross terms (f16f8)
                            ok = __all_sync(0xffffffffu, mbar_wait(b_full(bs), bph, a.err, ERR_B_FULL));
                            if (!ok) break;
                            tc_fence_after();
                            const uint32_t b0 = umma_desc_lo(sB + bs * B_SLOT_BYTES);
                            if (lead) {
                                {
                                    const uint32_t co = c * CHUNK_OFF;
                                    if (F8) {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if ((k & 1) >= k8steps) continue;
                                            umma_f8_lohi2(d_tmem + c * DCOLS, al + co + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128,
                                                          k < 2 ? idesc8a : idesc8b, 1u);
                                        }
                                    } else {
#pragma unroll
                                        for (int k = 0; k < 4; ++k) {
                                            if (k >= ksteps) break;
                                            umma_f16_lohi2(d_tmem + c * DCOLS, ah + co + k * 2, a_hi_word, b0 + k * 2, kUmmaDescHiSw128,
                                                           idesc, 1u);
                                        }
                                    }
                                }
                                umma_commit(b_empty(bs));
                            }
                            if (++bs == (uint32_t)a.b_slots) { bs = 0; bph ^= 1; }
                        }
                    }
                }
                if (lead && ok) umma_commit(a_empty(as));
                if (++as == (uint32_t)a.a_stages) { as = 0; aph ^= 1; }
            }
            if (lead && ok) umma_commit(acc_full(cs));
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
        }
    } else if (warp >= 2 && warp < 2 + kEpiWarps) {
        // ============================== epilogue ==============================
        const int ew = warp - 2;             // 0..7
        const int q4 = warp & 3;             // TMEM lane quarter this warp may read
        const int grp = ew >> 2;             // the two warps of a quarter split the chunks (or the channel range)
        uint32_t cs = 0, cph = 0;
        bool ok = true;
        // this warp's share of the tile: chunk my_c, channel range [c_begin, c_end)
        constexpr int CSPAN = NARROW ? NT : (CHUNKS == 2 ? NT : NT / 2);
        const int my_c = CHUNKS == 2 ? grp : 0;
        const int c_begin = (NARROW || CHUNKS == 2) ? 0 : grp * CSPAN;
        const bool idle = NARROW && CHUNKS == 1 && grp == 1;
        const uint32_t stg = sStage32 + ew * kStageBytesPerWarp;
        const float relu_floor = a.act_relu ? 0.f : -INFINITY;
        [[maybe_unused]] const bool leaky = a.act_slope > 0.f;
        // Transpose staging: element (row r, 16-B group j) lives at group r*4 + (j ^ ((r >> 1) & 3)).
        //   write: lane = row, group j      -> the 8 lanes of a phase hit 8 different bank groups
        //   read : lane -> row (lane >> 2) + 8*i, group lane & 3 -> 2 rows x 4 groups per phase, again all different
        const uint32_t wr_base = stg + lane * 64, wr_sw = (lane >> 1) & 3;
        const uint32_t rd_row = lane >> 2, rd_grp = lane & 3;
        // chunk origin inside the tile (chunks sit side by side); MMA row m of a chunk is pixel (m & 7) of row (m >> 3)
        const int ch_x0 = my_c * 8, ch_y0 = 0;

        const uint32_t acc_empty_dst0 = PAIR ? mapa_shared(acc_empty(0), 0) : 0u, acc_empty_dst1 = PAIR ? mapa_shared(acc_empty(1), 0) : 0u;
        for (int tile = (F8 ? wid0 : static_cast<int>(blockIdx.x)); tile < a.num_tiles && ok; tile += (F8 ? wstride : static_cast<int>(gridDim.x))) {
            const TileCoord t = decode_tile(a, tile, prank);
            const uint32_t tacc = tmem_base + (static_cast<uint32_t>(q4 * 32) << 16) + cs * ACC_COLS + my_c * DCOLS;
            if constexpr (NARROW) {
                ok = __all_sync(0xffffffffu, mbar_wait(acc_full(cs), cph, a.err, ERR_ACC_FULL));
                if (!ok) break;
                tc_fence_after();
                if (!idle) {
                    const int ty = ch_y0 + q4 * 4 + (lane >> 3), tx = ch_x0 + (lane & 7);
                    const int y = t.y0 + ty, x = t.x0 + tx;
                    const bool valid = (y < a.H) && (x < a.W);
                    uint32_t v[NT];
#pragma unroll
                    for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld_32x16(tacc + c0, *reinterpret_cast<uint32_t(*)[16]>(&v[c0]));
                    if (STACK) {
                        uint32_t v2[NT];
#pragma unroll
                        for (int c0 = 0; c0 < NT; c0 += 16) tmem_ld_32x16(tacc + NT + c0, *reinterpret_cast<uint32_t(*)[16]>(&v2[c0]));
                        tmem_ld_wait();
#pragma unroll
                        for (int j = 0; j < NT; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    } else {
                        tmem_ld_wait();
                    }
                    if (valid) {
                        if (a.ps_cout > 0) epilogue_scalar_ps<PLANES, NT == 32 ? 6 : 3>(a, v, t.n, y, x);
                        else epilogue_scalar<PLANES, NT>(a, v, t.n, y, x);
                    }
                }
            } else if (EPI == 0 && F8 && a.tma_out) {
                // Activation-only output through TMA stores.  All arithmetic happens in the TMEM layout (lane = pixel of this
                // warp's 8 x 4 pixel block, 16 consecutive channels per step); each lane drops its 32 bytes of the fp16 plane
                // and its 16 + 16 bytes of the 8-bit row into a pixel-major staging block, and one lane hands the two 1 KB
                // blocks to the TMA unit (boxes {16 ch, 8 px, 4 rows} and {16 B, 2 half-rows, 8 px, 4 rows}).  No transposed
                // shared-memory reads and no LSU store wavefronts (ncu: l1tex__data_pipe_lsu_wavefronts 78 % busy on the
                // K = 576 layers, whose epilogue was the critical path); pixels outside the image are clipped by the TMA.
                constexpr int ITERS = CSPAN / 16;
                const int cg0 = a.act_off1 + t.nb * NT + c_begin;                // first output channel (slot) of this warp's span
                const int bx = t.x0 + ch_x0, by = t.y0 + ch_y0 + q4 * 4;        // pixel block origin
                ok = __all_sync(0xffffffffu, mbar_wait(acc_full(cs), cph, a.err, ERR_ACC_FULL));
                if (!ok) break;
                tc_fence_after();
                uint32_t v[16];
                tmem_ld_32x16(tacc + c_begin, v);
                tmem_ld_wait();
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const int c0 = it * 16;
                    float f[16];
#pragma unroll
                    for (int j = 0; j < 4; ++j) {
                        const float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + t.nb * NT + c_begin + c0 + 4 * j));
                        f[4 * j] = fmaxf(fmaf(__uint_as_float(v[4 * j]), kF8AccScale, b4.x), relu_floor);
                        f[4 * j + 1] = fmaxf(fmaf(__uint_as_float(v[4 * j + 1]), kF8AccScale, b4.y), relu_floor);
                        f[4 * j + 2] = fmaxf(fmaf(__uint_as_float(v[4 * j + 2]), kF8AccScale, b4.z), relu_floor);
                        f[4 * j + 3] = fmaxf(fmaf(__uint_as_float(v[4 * j + 3]), kF8AccScale, b4.w), relu_floor);
                    }
                    if (it + 1 < ITERS) tmem_ld_32x16(tacc + c_begin + c0 + 16, v);     // next 16 columns in flight
                    uint32_t h[8], l[8];
#pragma unroll
                    for (int j = 0; j < 8; ++j) split2_f32(f[2 * j], f[2 * j + 1], h[j], l[j]);
                    if (lane == 0) bulk_wait_read0();                                 // the previous step's boxes have left the staging block
                    __syncwarp();
                    const uint32_t s16 = stg + lane * 32, s8 = stg + 1024 + lane * 32;
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s16), "r"(h[0]), "r"(h[1]), "r"(h[2]), "r"(h[3]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s16 + 16), "r"(h[4]), "r"(h[5]), "r"(h[6]), "r"(h[7]) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s8), "r"(f8_pack_lo4(l[0], l[1])), "r"(f8_pack_lo4(l[2], l[3])),
                                 "r"(f8_pack_lo4(l[4], l[5])), "r"(f8_pack_lo4(l[6], l[7])) : "memory");
                    asm volatile("st.shared.v4.b32 [%0], {%1, %2, %3, %4};" ::"r"(s8 + 16), "r"(f8_pack_hi4(h[0], h[1])), "r"(f8_pack_hi4(h[2], h[3])),
                                 "r"(f8_pack_hi4(h[4], h[5])), "r"(f8_pack_hi4(h[6], h[7])) : "memory");
                    fence_proxy_async();
                    __syncwarp();
                    if (lane == 0) {
                        const int c = cg0 + c0;
                        tma_store_4d(&tmO_hi, stg, c, bx, by, t.n);
                        tma_store_5d(&tmO_8, stg + 1024, c & 63, 2 * (c >> 6), bx, by, t.n);
                        bulk_commit();
                    }
                    if (it + 1 < ITERS) tmem_ld_wait();
                }
            } else {
                // After the transpose lane l serves pixels (l >> 2) + 8*i of this warp's 32-pixel group, channels
                // 4*(l & 3) .. +3 of each 16-channel step: 4 lanes cover 64 contiguous bytes of fp32 per pixel.
                const int cg_lane = t.nb * NT + c_begin + 4 * rd_grp;
                uint32_t o_res[4], o_raw[4], o_act[4], o_msk[4];
                uint32_t vmask = 0;
#pragma unroll
                for (int i = 0; i < 4; ++i) {
                    const int ty = ch_y0 + q4 * 4 + i, tx = ch_x0 + static_cast<int>(rd_row);
                    const int y = t.y0 + ty, x = t.x0 + tx;
                    const bool valid = (y < a.H) && (x < a.W);
                    vmask |= (valid ? 1u : 0u) << i;
                    const uint32_t pix = valid ? static_cast<uint32_t>(t.n * a.opix_n + y * a.opix_y + x * a.opix_x) : 0u;
                    if (EPI & EPI_RES) o_res[i] = pix * a.res_cs + cg_lane;
                    if (EPI & EPI_RAW) o_raw[i] = pix * a.raw_cs + a.raw_off1 + cg_lane;
                    if (EPI & EPI_MASK) o_msk[i] = pix * a.mask_cs + a.mask_off + cg_lane;
                    if (EPI & EPI_D2S) o_act[i] = valid ? static_cast<uint32_t>((t.n * 2 * a.H + 2 * y) * (2 * a.W) + 2 * x) * a.act_cs : 0u;
                    else if (EPI & EPI_S2D)   // out[n, y/2, x/2, ((y&1)*2 + (x&1))*64 + c] = in[n, y, x, c]
                        o_act[i] = valid ? static_cast<uint32_t>((t.n * (a.H >> 1) + (y >> 1)) * (a.W >> 1) + (x >> 1)) * a.act_cs +
                                               static_cast<uint32_t>(((y & 1) * 2 + (x & 1)) * 64) + cg_lane : 0u;
                    else o_act[i] = pix * a.act_cs + a.act_off1 + cg_lane;
                }
                constexpr int ITERS = CSPAN / 16;
                // residual loads run one step ahead so that their DRAM latency is never exposed
                float4 rr[4];
                if (EPI & EPI_RES) {
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        rr[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                        if ((vmask >> i) & 1) rr[i] = __ldg(reinterpret_cast<const float4*>(a.res + o_res[i]));
                    }
                }
                ok = __all_sync(0xffffffffu, mbar_wait(acc_full(cs), cph, a.err, ERR_ACC_FULL));
                if (!ok) break;
                tc_fence_after();
                // TMEM loads and bias loads run one step ahead: the LDTM of step it+1 is in flight while step it is transposed,
                // converted and stored (with two epilogue warps per scheduler nothing else would hide its latency)
                uint32_t v[16], v2[16];
                float4 b4 = __ldg(reinterpret_cast<const float4*>(a.bias + cg_lane));
                tmem_ld_32x16(tacc + c_begin, v);
                if (STACK) tmem_ld_32x16(tacc + NT + c_begin, v2);
                tmem_ld_wait();
#pragma unroll
                for (int it = 0; it < ITERS; ++it) {
                    const int c0 = it * 16;
                    float4 rn[4];
                    if ((EPI & EPI_RES) && it + 1 < ITERS) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            rn[i] = make_float4(0.f, 0.f, 0.f, 0.f);
                            if ((vmask >> i) & 1) rn[i] = __ldg(reinterpret_cast<const float4*>(a.res + o_res[i] + c0 + 16));
                        }
                    }
                    uint2 mk[4];
                    if (EPI & EPI_MASK) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) {
                            mk[i] = make_uint2(0u, 0u);
                            if ((vmask >> i) & 1) mk[i] = __ldg(reinterpret_cast<const uint2*>(a.mask + o_msk[i] + c0));
                        }
                    }
                    if (STACK) {
#pragma unroll
                        for (int j = 0; j < 16; ++j) v[j] = __float_as_uint(__uint_as_float(v[j]) + __uint_as_float(v2[j]));
                    }
#pragma unroll
                    for (int j = 0; j < 4; ++j)
                        sts128(wr_base + ((j ^ wr_sw) << 4), __uint_as_float(v[4 * j]), __uint_as_float(v[4 * j + 1]),
                               __uint_as_float(v[4 * j + 2]), __uint_as_float(v[4 * j + 3]));
                    __syncwarp();
                    float4 bn = b4;
                    if (it + 1 < ITERS) {      // the registers of v are free again: fetch the next 16 columns
                        tmem_ld_32x16(tacc + c_begin + c0 + 16, v);
                        if (STACK) tmem_ld_32x16(tacc + NT + c_begin + c0 + 16, v2);
                        bn = __ldg(reinterpret_cast<const float4*>(a.bias + cg_lane + c0 + 16));
                    }
                    float4 vals[4];      // all four shared-memory reads in flight before the first use
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const uint32_t r = rd_row + 8 * i;
                        vals[i] = lds128(stg + r * 64 + ((rd_grp ^ ((r >> 1) & 3)) << 4));
                    }
                    // fused 2x2 max-pool (closing conv of an encoder level, ops.py:52-54): post-ReLU values of this lane's 4 rows
                    [[maybe_unused]] float pf[4][4];
                    const bool pooling = (EPI == EPI_RES) && a.pool_out != nullptr;
                    // outputs inside a wider buffer (split mode): columns past the true channel count are not stored
                    const uint32_t smask = (PLANES == 2 && a.store_cout > 0 && cg_lane + c0 >= a.store_cout) ? 0u : vmask;
#pragma unroll
                    for (int i = 0; i < 4; ++i) {
                        const float4 val = vals[i];
                        if (EPI == EPI_RES) { pf[i][0] = pf[i][1] = pf[i][2] = pf[i][3] = 0.f; }
                        if ((smask >> i) & 1) {
                            float f0, f1, f2, f3;
                            if (F8) {
                                f0 = fmaf(val.x, kF8AccScale, b4.x); f1 = fmaf(val.y, kF8AccScale, b4.y);
                                f2 = fmaf(val.z, kF8AccScale, b4.z); f3 = fmaf(val.w, kF8AccScale, b4.w);
                            } else {
                                f0 = val.x + b4.x; f1 = val.y + b4.y; f2 = val.z + b4.z; f3 = val.w + b4.w;
                            }
                            if (EPI & EPI_MASK) {   // fp16 bit patterns 0x0001..0x7FFF are the positive values (post-ReLU: never negative)
                                if ((mk[i].x & 0x7FFFu) == 0u) f0 = 0.f;
                                if ((mk[i].x & 0x7FFF0000u) == 0u) f1 = 0.f;
                                if ((mk[i].y & 0x7FFFu) == 0u) f2 = 0.f;
                                if ((mk[i].y & 0x7FFF0000u) == 0u) f3 = 0.f;
                            }
                            if (EPI & EPI_RES) { f0 += rr[i].x; f1 += rr[i].y; f2 += rr[i].z; f3 += rr[i].w; }
                            if (EPI & EPI_RAW) *reinterpret_cast<float4*>(a.out_raw + o_raw[i] + c0) = make_float4(f0, f1, f2, f3);
                            f0 = fmaxf(f0, relu_floor); f1 = fmaxf(f1, relu_floor);
                            f2 = fmaxf(f2, relu_floor); f3 = fmaxf(f3, relu_floor);
                            if (PLANES == 2 && leaky) {
                                f0 = fmaxf(f0, f0 * a.act_slope); f1 = fmaxf(f1, f1 * a.act_slope);
                                f2 = fmaxf(f2, f2 * a.act_slope); f3 = fmaxf(f3, f3 * a.act_slope);
                            }
                            if (EPI == EPI_RES) { pf[i][0] = f0; pf[i][1] = f1; pf[i][2] = f2; pf[i][3] = f3; }
                            uint32_t h01, l01, h23, l23;
                            split2_f32(f0, f1, h01, l01);
                            split2_f32(f2, f3, h23, l23);
                            __half* d;
                            if (EPI & EPI_D2S) {   // tf.depth_to_space(x, 2): out[n,2y+i,2x+j,c] = in[n,y,x,(2i+j)*64+c]
                                const int cg = cg_lane + c0, g = cg >> 6;
                                d = a.out_act + o_act[i] + static_cast<uint32_t>(((g >> 1) * (2 * a.W) + (g & 1)) * a.act_cs + (cg & 63));
                            } else {
                                d = a.out_act + o_act[i] + c0;
                            }
                            *reinterpret_cast<uint2*>(d) = make_uint2(h01, h23);
                            if (PLANES == 2) *reinterpret_cast<uint2*>(d + a.act_plane) = make_uint2(l01, l23);
                            if (F8) {   // 8-bit plane: [e5m2(16 lo) x 64 | e5m2(hi) x 64] per 64-channel block of the pixel
                                uint8_t* q = f8_row_ptr(d + a.act_plane);
                                *reinterpret_cast<uint32_t*>(q) = f8_pack_lo4(l01, l23);
                                *reinterpret_cast<uint32_t*>(q + 64) = f8_pack_hi4(h01, h23);
                            }
                        }
                    }
                    if constexpr (EPI == EPI_RES) {
                        if (pooling) {
                            // rows (0,1) and (2,3) of this lane, columns x and x^1 of the lane 4 further: image sizes and tile
                            // origins are even, so a 2x2 block is inside the image as a whole or not at all
                            float m[2][4];
#pragma unroll
                            for (int c = 0; c < 4; ++c) {
                                m[0][c] = fmaxf(pf[0][c], pf[1][c]);
                                m[1][c] = fmaxf(pf[2][c], pf[3][c]);
                                m[0][c] = fmaxf(m[0][c], __shfl_xor_sync(0xffffffffu, m[0][c], 4));
                                m[1][c] = fmaxf(m[1][c], __shfl_xor_sync(0xffffffffu, m[1][c], 4));
                            }
                            const int odd = rd_row & 1;                  // even columns store row pair 0, odd columns row pair 1
                            if ((vmask >> (2 * odd)) & 1) {
                                const int py = (t.y0 + ch_y0 + q4 * 4 + 2 * odd) >> 1, px = (t.x0 + ch_x0 + static_cast<int>(rd_row)) >> 1;
                                __half* d = a.pool_out + (static_cast<size_t>(t.n * (a.H >> 1) + py) * (a.W >> 1) + px) * a.pool_cs + cg_lane + c0;
                                uint32_t h01, l01, h23, l23;
                                split2_f32(odd ? m[1][0] : m[0][0], odd ? m[1][1] : m[0][1], h01, l01);
                                split2_f32(odd ? m[1][2] : m[0][2], odd ? m[1][3] : m[0][3], h23, l23);
                                *reinterpret_cast<uint2*>(d) = make_uint2(h01, h23);
                                if (PLANES == 2) *reinterpret_cast<uint2*>(d + a.pool_plane) = make_uint2(l01, l23);
                                if (F8) {
                                    uint8_t* q = f8_row_ptr(d + a.pool_plane);
                                    *reinterpret_cast<uint32_t*>(q) = f8_pack_lo4(l01, l23);
                                    *reinterpret_cast<uint32_t*>(q + 64) = f8_pack_hi4(h01, h23);
                                }
                            }
                        }
                    }
                    __syncwarp();
                    if (it + 1 < ITERS) {
                        tmem_ld_wait();
                        b4 = bn;
                    }
                    if ((EPI & EPI_RES) && it + 1 < ITERS) {
#pragma unroll
                        for (int i = 0; i < 4; ++i) rr[i] = rn[i];
                    }
                }
            }
            tc_fence_before();
            __syncwarp();
            if (lane == 0) {
                if (PAIR) mbar_arrive_cluster(cs ? acc_empty_dst1 : acc_empty_dst0);      // the leader's barrier
                else mbar_arrive(acc_empty(cs));
            }
            if (++cs == 2) { cs = 0; cph ^= 1; }
        }
    }

    if (EPI == 0 && F8 && lane == 0 && warp >= 2 && warp < 2 + kEpiWarps) bulk_wait0();     // outstanding TMA stores of this warp
    tc_fence_before();
    if (PAIR) cluster_sync_all(); else __syncthreads();       // pair: no CTA leaves while its peer may still signal its barriers
    if (warp == 1) { if (PAIR) tmem_dealloc_pair(tmem_base, TMEM_COLS); else tmem_dealloc(tmem_base, TMEM_COLS); }
}

template <int NT, int CHUNKS, int PLANES, int EPI>
cudaError_t init_inst() {
    return cudaFuncSetAttribute(conv3x3_umma_kernel<NT, CHUNKS, PLANES, EPI>, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                kConvMaxSmem);
}

template <int NT, int CHUNKS, int PLANES, int EPI>
cudaError_t launch_inst(const ConvLaunch& L, int num_sms, cudaStream_t stream) {
    if (PLANES == 4) {       // CTA pairs: L.args.num_tiles counts pairs of tiles, one cluster of 2 per pair in flight
        const int pairs = num_sms / 2;
        const int clusters = L.args.num_tiles < pairs ? L.args.num_tiles : pairs;
        cudaLaunchConfig_t cfg = {};
        cfg.gridDim = dim3(2 * clusters);
        cfg.blockDim = dim3(conv_threads(PLANES));
        cfg.dynamicSmemBytes = L.smem_bytes;
        cfg.stream = stream;
        cudaLaunchAttribute attr[1];
        attr[0].id = cudaLaunchAttributeClusterDimension;
        attr[0].val.clusterDim.x = 2; attr[0].val.clusterDim.y = 1; attr[0].val.clusterDim.z = 1;
        cfg.attrs = attr;
        cfg.numAttrs = 1;
        return cudaLaunchKernelEx(&cfg, conv3x3_umma_kernel<NT, CHUNKS, PLANES, EPI>, L.tmA_hi, L.tmA_lo, L.tmB, L.tmO_hi, L.tmO_8, L.args);
    }
    const int grid = L.args.num_tiles < num_sms ? L.args.num_tiles : num_sms;
    conv3x3_umma_kernel<NT, CHUNKS, PLANES, EPI><<<grid, conv_threads(PLANES), L.smem_bytes, stream>>>(L.tmA_hi, L.tmA_lo, L.tmB, L.tmO_hi, L.tmO_8, L.args);
    return cudaGetLastError();
}

// One translation unit per (PLANES, NT) instantiates these (conv_inst_*.cu).
template <int NT, int PLANES>
cudaError_t init_family();
template <int NT, int PLANES>
cudaError_t launch_family(const ConvLaunch& L, int num_sms, cudaStream_t stream);

#define FISR_CONV_FAMILY(NT_, PL_, LIST)                                                                    \
    template <>                                                                                             \
    cudaError_t init_family<NT_, PL_>() {                                                                   \
        cudaError_t e = cudaSuccess;                                                                        \
        LIST(FISR_INIT_ONE, NT_, PL_)                                                                       \
        return e;                                                                                           \
    }                                                                                                       \
    template <>                                                                                             \
    cudaError_t launch_family<NT_, PL_>(const ConvLaunch& L, int num_sms, cudaStream_t stream) {            \
        LIST(FISR_LAUNCH_ONE, NT_, PL_)                                                                     \
        return cudaErrorInvalidValue;                                                                       \
    }
#define FISR_INIT_ONE(NT_, PL_, CH_, EPI_) if (e == cudaSuccess) e = init_inst<NT_, CH_, PL_, (EPI_)>();
#define FISR_LAUNCH_ONE(NT_, PL_, CH_, EPI_) \
    if (L.chunks == CH_ && L.epi == (EPI_)) return launch_inst<NT_, CH_, PL_, (EPI_)>(L, num_sms, stream);
// epilogue variants the network uses: act | act+raw | act+res | act+raw+res | act+d2s
#define FISR_FOR_EPI(M, NT_, PL_)                                                                           \
    M(NT_, PL_, 1, 0) M(NT_, PL_, 2, 0) M(NT_, PL_, 1, EPI_RAW) M(NT_, PL_, 2, EPI_RAW)                      \
    M(NT_, PL_, 1, EPI_RES) M(NT_, PL_, 2, EPI_RES) M(NT_, PL_, 1, EPI_RES | EPI_RAW) M(NT_, PL_, 2, EPI_RES | EPI_RAW) \
    M(NT_, PL_, 1, EPI_D2S) M(NT_, PL_, 2, EPI_D2S)
#define FISR_FOR_EPI_NARROW(M, NT_, PL_) M(NT_, PL_, 1, 0) M(NT_, PL_, 2, 0)      // NT = 16 / 32: scalar epilogue
// N = 96 tiles (PWC-Net): act | act+raw | act+res (the two-launch stride-2 convs accumulate through raw -> res)
#define FISR_FOR_EPI_PWC(M, NT_, PL_)                                                                       \
    M(NT_, PL_, 1, 0) M(NT_, PL_, 2, 0) M(NT_, PL_, 1, EPI_RAW) M(NT_, PL_, 2, EPI_RAW) M(NT_, PL_, 1, EPI_RES) M(NT_, PL_, 2, EPI_RES)
// dgrad variants (training, split mode only): mask | mask+res | mask+res+raw | mask+raw (| mask+s2d for the 64-wide tile)
#define FISR_FOR_EPI_BWD(M, NT_, PL_)                                                                       \
    M(NT_, PL_, 1, EPI_MASK) M(NT_, PL_, 2, EPI_MASK) M(NT_, PL_, 1, EPI_MASK | EPI_RES) M(NT_, PL_, 2, EPI_MASK | EPI_RES) \
    M(NT_, PL_, 1, EPI_MASK | EPI_RES | EPI_RAW) M(NT_, PL_, 2, EPI_MASK | EPI_RES | EPI_RAW)               \
    M(NT_, PL_, 1, EPI_MASK | EPI_RAW) M(NT_, PL_, 2, EPI_MASK | EPI_RAW)
#define FISR_FOR_EPI_TRAIN(M, NT_, PL_) FISR_FOR_EPI(M, NT_, PL_) FISR_FOR_EPI_BWD(M, NT_, PL_)
#define FISR_FOR_EPI_TRAIN64(M, NT_, PL_) \
    FISR_FOR_EPI_TRAIN(M, NT_, PL_) M(NT_, PL_, 1, EPI_MASK | EPI_S2D) M(NT_, PL_, 2, EPI_MASK | EPI_S2D)

}  // namespace convk
}  // namespace fisr
