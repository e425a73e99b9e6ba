// Host side of the tcgen05 3x3 conv: kernel-family dispatch and tile geometry / shared-memory planning.
// The device code lives in conv_umma_kernel.cuh and is instantiated in conv_inst_p{1,2}_n{16,64,128}.cu.
#include <cstdlib>
#include <cstring>
#include <string>
#include "common.cuh"
#include "conv_umma.h"
#include "sm100_ptx.cuh"

namespace fisr {

namespace convk {
constexpr int kMaxBSlots = 12;
constexpr int kStageBytes = 8 * 32 * 64;         // kEpiWarps x [32 px][16 ch] fp32 (conv_umma_kernel.cuh)
template <int NT, int PLANES>
cudaError_t init_family();
template <int NT, int PLANES>
cudaError_t launch_family(const ConvLaunch& L, int num_sms, cudaStream_t stream);
#define FISR_DECL(NT_, PL_)                                     \
    template <> cudaError_t init_family<NT_, PL_>();            \
    template <> cudaError_t launch_family<NT_, PL_>(const ConvLaunch&, int, cudaStream_t);
FISR_DECL(16, 1) FISR_DECL(64, 1) FISR_DECL(128, 1) FISR_DECL(16, 2) FISR_DECL(64, 2) FISR_DECL(96, 2) FISR_DECL(128, 2)
FISR_DECL(16, 3) FISR_DECL(32, 3) FISR_DECL(64, 3) FISR_DECL(128, 3) FISR_DECL(64, 4) FISR_DECL(128, 4)
#undef FISR_DECL
}  // namespace convk

// Opts every instantiation into the large dynamic shared-memory carve-out (once per device, outside graph capture).
cudaError_t conv3x3_init() {
    using namespace convk;
    cudaError_t e = init_family<16, 1>();
    if (e == cudaSuccess) e = init_family<64, 1>();
    if (e == cudaSuccess) e = init_family<128, 1>();
    if (e == cudaSuccess) e = init_family<16, 2>();
    if (e == cudaSuccess) e = init_family<64, 2>();
    if (e == cudaSuccess) e = init_family<96, 2>();
    if (e == cudaSuccess) e = init_family<128, 2>();
    if (e == cudaSuccess) e = init_family<16, 3>();
    if (e == cudaSuccess) e = init_family<32, 3>();
    if (e == cudaSuccess) e = init_family<64, 3>();
    if (e == cudaSuccess) e = init_family<128, 3>();
    if (e == cudaSuccess) e = init_family<64, 4>();
    if (e == cudaSuccess) e = init_family<128, 4>();
    return e;
}

cudaError_t launch_conv3x3(const ConvLaunch& L, int num_sms, cudaStream_t stream) {
    using namespace convk;
    if (L.planes == 3 && L.pair) {
        if (L.NT == 64) return launch_family<64, 4>(L, num_sms, stream);
        if (L.NT == 128) return launch_family<128, 4>(L, num_sms, stream);
    } else if (L.planes == 3) {
        if (L.NT == 16) return launch_family<16, 3>(L, num_sms, stream);
        if (L.NT == 32) return launch_family<32, 3>(L, num_sms, stream);
        if (L.NT == 64) return launch_family<64, 3>(L, num_sms, stream);
        if (L.NT == 128) return launch_family<128, 3>(L, num_sms, stream);
    } else if (L.planes == 2) {
        if (L.NT == 16) return launch_family<16, 2>(L, num_sms, stream);
        if (L.NT == 64) return launch_family<64, 2>(L, num_sms, stream);
        if (L.NT == 96) return launch_family<96, 2>(L, num_sms, stream);
        if (L.NT == 128) return launch_family<128, 2>(L, num_sms, stream);
    } else {
        if (L.NT == 16) return launch_family<16, 1>(L, num_sms, stream);
        if (L.NT == 64) return launch_family<64, 1>(L, num_sms, stream);
        if (L.NT == 128) return launch_family<128, 1>(L, num_sms, stream);
    }
    return cudaErrorInvalidValue;
}

// Host-side tile geometry / shared-memory carve-up for one conv launch.
bool plan_conv_geometry(int H, int W, int n_img, int cout_pad, int planes, int num_sms, ConvLaunch* L, int kb) {
    int NT = cout_pad <= 16 ? 16 : cout_pad <= 32 ? 32 : 64;
    // Wide N tiles halve the A re-reads and balance shared-memory operand traffic against MMA math.  The choice
    // depends on the per-image geometry only (never on the batch), so results are bit-identical however a
    // window's tiles are sharded over batches / GPUs (NT decides the fp32 summation order, see STACK).
    if (cout_pad >= 128 && (long)H * W >= 64L * 64) NT = 128;
    if (planes == 3 && cout_pad >= 128) {
        // f16f8: every accumulator column sees the same MMAs in the same order whatever NT, so the choice may look at the batch: take
        // the N tile with the cheaper schedule, waves x (time of one work item).  An N = 64 item covers half the channels of an
        // N = 128 one but runs at the shared-memory operand roof (48 instead of 32 cycles per MMA): 0.67 of its time, not 0.5.
        const long tiles = (long)n_img * ((H + 15) / 16) * ((W + 15) / 16);
        const long w128 = (tiles * (cout_pad / 128) + num_sms - 1) / num_sms, w64 = (tiles * (cout_pad / 64) + num_sms - 1) / num_sms;
        NT = (3 * w128 <= 2 * w64) ? 128 : 64;
    }
    if (planes == 2 && cout_pad == 96) NT = 96;          // PWC-Net's 96-channel layers (split mode): no padding to 128
    (void)kb;
    // Two chunks (16 x 16 tiles) unless the image is a single chunk wide.  Every chunk has its own MMA issuer warp and one
    // thread issues a tcgen05.mma only every ~70-150 cycles (tests/cuda/umma_rate_probe.cu), so a one-chunk CTA is ISSUE
    // bound at about half the rate of a two-chunk one: splitting small layers into 8-pixel-wide tiles to occupy more SMs
    // (round 1) doubled the tile count without shortening a tile (ncu, profiles/r02_ncu_small_layers.md: level-1 bottleneck
    // at 30 % tensor-active with 1 chunk against 57 % for the level-2 one with 2).
    int chunks = W > 8 ? 2 : 1;
    (void)num_sms;
    const int apl = act_planes(planes);
    const bool stack = planes == 2 && NT <= 64;
    // f16f8 layers with wide outputs can run on CTA pairs (cluster of 2, cta_group::2, M = 256): each CTA keeps half of a tap's
    // weight rows, i.e. half the L2 -> smem weight traffic and half the shared-memory B reads per MMA.  Measured on B200
    // (profiles/r01_pair_mode.txt): bit-identical results, 1-3 % faster on act-only layers, 5-9 % slower on the residual
    // layers (the slower epilogue of the two CTAs gates both), 26.73 vs 26.78 ms per forward -- so it is opt-in (FISR_PAIR=1).
    bool pair = false;
    if (const char* e = getenv("FISR_PAIR")) pair = planes == 3 && NT >= 64 && atoi(e) != 0;
    const int slot_bytes = stack ? 2 * NT * 128 : (pair ? NT / 2 : NT) * 128;
    const int fixed = 128 /*align*/ + convk::kStageBytes /*epilogue transpose*/;
    // A chunk is 8 px x 16 rows (one 128-row MMA, one image row of the patch per 8-row group); two chunks sit side by side:
    // tile 16 x 16, patch 18 x 18.  The geometry is a compile-time function of CHUNKS in the kernel.
    const int cx = chunks, cy = 1;
    double best_eff;
    {
        long tx = (W + 8 * cx - 1) / (8 * cx);
        if (pair) tx = (tx + 1) / 2 * 2;          // a pair whose right tile lies past the image edge still issues its rows
        const long tiles = tx * ((H + 15) / 16);
        best_eff = (double)H * W / ((double)tiles * 128 * chunks);
    }
    ConvArgs& a = L->args;
    a.TW = 8 * cx; a.TH = 16 * cy; a.P = a.TW + 2;
    a.tiles_x = (W + a.TW - 1) / a.TW;
    if (pair) a.tiles_x = (a.tiles_x + 1) / 2;
    a.tiles_y = (H + a.TH - 1) / a.TH;
    a.NB = cout_pad / NT;
    a.num_tiles = n_img * a.tiles_x * a.tiles_y * a.NB;
    a.a_plane_bytes = (a.TH + 2) * a.P * 128;
    // f16f8 (phase-split main loop): one buffer per activation plane for the wide tiles (a second one measured 10-17 % slower on
    // the 64-wide layers: it costs weight slots), two for the narrow conv/2 heads, which are HBM bound on their 256-channel input
    // (4.7 TB/s, 72 % of the measured copy peak) and gain 3-6 % from the deeper patch prefetch (profiles/r02_layers_*.txt)
    a.a_stages = planes == 3 ? (NT <= 32 ? 2 : 1) : 2;
    for (int i = 0; i < 8; ++i) a.tapmask[i] = 0x1FFu;
    a.ps_cout = 0;
    a.opix_n = H * W; a.opix_y = W; a.opix_x = 1;
    a.act_slope = 0.f; a.store_cout = 0;
    int slots = (kConvMaxSmem - fixed - a.a_stages * apl * a.a_plane_bytes) / slot_bytes;
    if (slots > convk::kMaxBSlots) slots = convk::kMaxBSlots;
    if (slots < 2) return false;
    a.b_slots = slots;
    L->NT = NT; L->chunks = chunks; L->planes = planes; L->pair = pair;
    L->smem_bytes = fixed + a.a_stages * apl * a.a_plane_bytes + slots * slot_bytes;
    L->efficiency = best_eff;
    return true;
}


bool build_split_conv(EncodeTiledFn encode, const SplitConvDesc& d, int num_sms, int* d_err, ConvLaunch* L, std::string* why) {
    auto bad = [&](const std::string& m) { if (why) *why = m; return false; };
    memset(static_cast<void*>(L), 0, sizeof *L);
    if (d.cout_pad != split_conv_cout_pad(d.cout)) return bad("cout_pad does not match split_conv_cout_pad(cout)");
    if ((d.in_cs | d.cin_off | d.out_cs) % 8 || d.out_off % 4 || d.in_sx % 8 || d.in_sy % 8 || d.in_sn % 8 || d.in_plane % 8 || d.out_plane % 4)
        return bad("channel strides / offsets must keep 16-byte (input) and 8-byte (output) alignment");
    if (static_cast<double>(d.out_pixels) * d.out_cs >= 4294967296.0) return bad("output exceeds 32-bit element offsets");
    if (!plan_conv_geometry(d.H, d.W, d.N, d.cout_pad, 2, num_sms, L)) return bad("no tile geometry");
    ConvArgs& a = L->args;
    a.bias = d.bias;
    a.out_act = d.out; a.act_plane = d.out_plane;
    a.act_cs = d.out_cs; a.act_off0 = a.act_off1 = d.out_off; a.act_split = 0;
    a.act_relu = d.relu ? 1 : 0; a.act_slope = d.relu ? 0.f : d.slope;
    a.scalar_out = d.cout_pad <= 16;
    a.err = d_err;
    a.N = d.N; a.H = d.H; a.W = d.W;
    a.opix_n = static_cast<int>(d.opix_n); a.opix_y = static_cast<int>(d.opix_y); a.opix_x = static_cast<int>(d.opix_x);
    a.cin_off = d.cin_off; a.KB = (d.cin + 63) / 64; a.cout = d.cout;
    a.ksteps_last = (d.cin - (a.KB - 1) * 64 + 15) / 16;
    a.store_cout = (d.cout_pad > 16 && d.cout < d.cout_pad) ? d.cout : 0;
    L->epi = 0;
    if (d.raw || d.res) {
        if (d.cout_pad <= 16 || d.cout % 4) return bad("fp32 side tensors need a wide output with a multiple of 4 channels");
        a.out_raw = d.raw; a.raw_cs = d.cout; a.raw_off0 = a.raw_off1 = 0; a.raw_split = 0;
        a.res = d.res; a.res_cs = d.cout;
        L->epi = (d.res ? 1 : 0) | (d.raw ? 2 : 0);
        if (static_cast<double>(d.out_pixels) * d.cout >= 4294967296.0) return bad("fp32 side tensor exceeds 32-bit element offsets");
    }
    L->tma_out = false;
    for (int pl = 0; pl < 2; ++pl) {
        cuuint64_t dims[4] = {(cuuint64_t)d.in_cs, (cuuint64_t)d.W, (cuuint64_t)d.H, (cuuint64_t)d.N};
        cuuint64_t strides[3] = {(cuuint64_t)d.in_sx * 2, (cuuint64_t)d.in_sy * 2, (cuuint64_t)d.in_sn * 2};
        cuuint32_t box[4] = {64, (cuuint32_t)a.P, (cuuint32_t)(a.TH + 2), 1};
        cuuint32_t es[4] = {1, 1, 1, 1};
        const CUresult r = encode(pl ? &L->tmA_lo : &L->tmA_hi, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 4, const_cast<__half*>(d.in + (pl ? d.in_plane : 0)), dims,
                                  strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                                  CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return bad("cuTensorMapEncodeTiled(input view) failed: " + std::to_string((int)r));
    }
    {
        cuuint64_t dims[2] = {64, (cuuint64_t)2 * a.KB * 9 * d.cout_pad};
        cuuint64_t strides[1] = {128};
        cuuint32_t box[2] = {64, (cuuint32_t)L->NT};
        cuuint32_t es[2] = {1, 1};
        const CUresult r = encode(&L->tmB, CU_TENSOR_MAP_DATA_TYPE_FLOAT16, 2, const_cast<__half*>(d.wp), dims, strides, box, es, CU_TENSOR_MAP_INTERLEAVE_NONE,
                                  CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
        if (r != CUDA_SUCCESS) return bad("cuTensorMapEncodeTiled(weights) failed: " + std::to_string((int)r));
    }
    return true;
}

}  // namespace fisr
