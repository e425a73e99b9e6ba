// Shared helpers for the fisr_b200 CUDA sources.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_fp8.h>
#include <cuda_runtime.h>

namespace fisr {

// Precision modes of the conv stack (DESIGN.md "Numerics").
//   F16X3: activations and weights are stored as an fp16 (hi, lo) pair, x = hi + lo (22-bit mantissa);
//          each K-slice issues 3 MMAs hi*hi + lo*hi + hi*lo with fp32 TMEM accumulation (fp32-class result).
//   F16  : hi plane only, 1 MMA per K-slice (fast mode, ~7e-4 max-abs on the 138-conv cascade).
//   F16F8: fp16 hi plane + an 8-bit plane; the main product hi*hi runs as fp16 MMAs, the two cross terms lo*hi + hi*lo as
//          fp8 MMAs (kind::f8f6f4, twice the fp16 rate) into the same accumulator: 2 MMA units per K-slice instead of 3.
//          The cross terms are ~2^-11 of the main term, so their 2..3-bit mantissas leave ~2^-14 relative error per conv
//          (tools/precision_study_fp8.py: 2.5e-5 max-abs on the 138-conv cascade, 40x inside the 1e-3 bar).
//          8-bit plane of an activation, per pixel and 64-channel block (128 B, same footprint as the fp16 lo plane):
//              bytes [0,64)  = e5m2(16 * lo[c])      bytes [64,128) = e5m2(hi[c])       (e5m2 = rounded upper byte of an fp16)
//          weights: fp16 plane = fp16(128 * w) =: wh, 8-bit plane row = [e4m3(wh / 16) x 64 | e5m2(128 w - wh) x 64];
//          every product term carries the factor 128, which the epilogue removes.
enum Precision { PREC_F16X3 = 0, PREC_F16 = 1, PREC_F16F8 = 2 };
constexpr float kF8ActLoScale = 16.f;     // activation lo part -> e5m2
constexpr float kF8WScale = 128.f;        // weights -> fp16 plane and e5m2 lo part
constexpr float kF8WHiScale = 8.f;        // weight hi part -> e4m3   (16 * 8 = 1 * 128 = 128)
// number of 2-byte planes an activation buffer holds for a given kernel PLANES parameter (1 = f16, 2 = f16x3, 3 = f16f8)
__host__ __device__ constexpr int act_planes(int planes) { return planes == 1 ? 1 : 2; }

struct SplitHalf {
    __half hi, lo;
};
__device__ __forceinline__ SplitHalf split_f32(float x) {
    SplitHalf s;
    s.hi = __float2half_rn(x);
    s.lo = __float2half_rn(x - __half2float(s.hi));
    return s;
}
// Packed split of two values: cvt.rn.f16x2.f32 (F2FP, full-rate ALU) instead of scalar F2F conversions, which run
// on the quarter-rate XU pipe and were the measured bottleneck of the conv epilogue (ncu: pipe_xu at 109 %).
__device__ __forceinline__ void split2_f32(float x0, float x1, uint32_t& hi, uint32_t& lo) {
    const __half2 h = __floats2half2_rn(x0, x1);
    const float2 b = __half22float2(h);
    const __half2 l = __floats2half2_rn(x0 - b.x, x1 - b.y);
    hi = *reinterpret_cast<const uint32_t*>(&h);
    lo = *reinterpret_cast<const uint32_t*>(&l);
}
__device__ __forceinline__ float join_f16(__half hi, __half lo) { return __half2float(hi) + __half2float(lo); }

// ---- F16F8 8-bit plane helpers
// Byte address of the 8-bit row entry that belongs to the element p1 points at in the second plane (same element index
// as in the hi plane).  Buffers are >= 128-B aligned and pixel strides are multiples of 64 channels, so the channel
// index inside the 64-channel block is ((address >> 1) & 63): block base + c for the lo byte, + 64 + c for the hi byte.
__device__ __forceinline__ uint8_t* f8_row_ptr(const __half* p1) {
    const uintptr_t a = reinterpret_cast<uintptr_t>(p1);
    return reinterpret_cast<uint8_t*>(a - ((a >> 1) & 63));
}
__device__ __forceinline__ uint8_t f8_lo_byte(float lo) {    // e5m2 of fp16(lo) * 16 (same arithmetic as f8_pack_lo4)
    const uint32_t h = __half_as_ushort(__hmul(__float2half_rn(lo), __float2half_rn(kF8ActLoScale)));
    return static_cast<uint8_t>((h + 0x7Fu + ((h >> 8) & 1u)) >> 8);
}
__device__ __forceinline__ uint8_t f8_hi_byte(float x) {     // e5m2 of the fp16 hi part (same rounding as f8_pack_hi4)
    const uint32_t h = __half_as_ushort(__float2half_rn(x));
    return static_cast<uint8_t>((h + 0x7Fu + ((h >> 8) & 1u)) >> 8);
}
// Two packed fp16 -> two e5m2 bytes (left in byte 1 and byte 3 of the result): an e5m2 number is the upper byte of the fp16
// with the same value, so the conversion is a round-to-nearest-even of the low byte, done with integer adds.  (The
// cvt.rn.satfinite.e5m2x2.f16x2 instruction runs on the quarter-rate XU pipe: ncu showed it at 179 % of its sustained peak
// in the f16f8 epilogue, which made the K = 576 layers epilogue bound.)  No carry can cross the halves for finite inputs.
__device__ __forceinline__ uint32_t f16x2_round_to_e5m2(uint32_t w) { return w + 0x007F007Fu + ((w >> 8) & 0x00010001u); }
// l01 / l23: packed fp16 lo parts of 4 consecutive channels -> 4 e5m2 bytes of 16 * lo (channel order = byte order)
__device__ __forceinline__ uint32_t f8_pack_lo4(uint32_t l01, uint32_t l23) {
    const __half2 s = __floats2half2_rn(kF8ActLoScale, kF8ActLoScale);
    const __half2 a = __hmul2(*reinterpret_cast<const __half2*>(&l01), s), b = __hmul2(*reinterpret_cast<const __half2*>(&l23), s);
    return __byte_perm(f16x2_round_to_e5m2(*reinterpret_cast<const uint32_t*>(&a)),
                       f16x2_round_to_e5m2(*reinterpret_cast<const uint32_t*>(&b)), 0x7531);
}
// h01 / h23: packed fp16 hi parts of 4 consecutive channels -> 4 e5m2 bytes of hi
__device__ __forceinline__ uint32_t f8_pack_hi4(uint32_t h01, uint32_t h23) {
    return __byte_perm(f16x2_round_to_e5m2(h01), f16x2_round_to_e5m2(h23), 0x7531);
}
// value an (hi, 8-bit row) pair stands for: hi + e5m2 byte / 16 (an e5m2 byte is the upper byte of the fp16 with the same value)
__device__ __forceinline__ float join_f8(__half hi, uint8_t lo_byte) {
    return __half2float(hi) + __half2float(__ushort_as_half(static_cast<unsigned short>(lo_byte) << 8)) * (1.f / kF8ActLoScale);
}

// Everything the 3x3 conv kernel needs besides the tensor maps (see conv_umma.cu).
struct ConvArgs {
    const float* bias;        // [cout_pad]
    const float* res;         // fp32 residual, NHWC with res_cs channels, or nullptr
    float* out_raw;           // fp32 output (pre-activation), or nullptr
    __half* out_act;          // fp16 hi plane of the activation output, or nullptr (lo plane at +act_plane)
    unsigned long long act_plane;   // elements between the hi and the lo plane of out_act
    int* err;                 // device error flag (0 = ok)
    int N, H, W;              // output (= input) geometry
    int cin_off, KB;          // first input channel in the source buffer, number of 64-channel K blocks
    int ksteps_last;          // 16-channel K slices of the last block that hold real input channels (1..4)
    int cout, NB;             // true output channels, number of N blocks (cout_pad = NB * NT)
    int P, TH, TW;            // patch pitch (= TW + 2), output rows / cols per tile (TW = 8 * CHUNKS, TH = 16): host-side copy of
                              // the kernel's compile-time geometry, used for the TMA box and the tile grid
    int tiles_x, tiles_y, num_tiles;
    int raw_cs, raw_off0, raw_off1, raw_split;   // fp32 output channel stride / channel map
    int act_cs, act_off0, act_off1, act_split;   // fp16 output channel stride / channel map
    int act_relu, act_d2s, scalar_out, res_cs;
    int a_plane_bytes, a_stages, b_slots;        // smem carve-up (host computed)
    unsigned tapmask[8];      // f16f8: bit t of tapmask[kb] = tap t of K block kb has non-zero weights (others are skipped)
    int tma_out;              // f16f8 activation-only epilogue: store through the output tensor maps (TMA) instead of the LSU
    int ps_cout;              // > 0: the output is depth_to_space'd on the fly -- GEMM column ch = (2i+j) * ps_cout + co is channel co
                              // of output pixel (2y+i, 2x+j) (the conv/2 heads evaluated at input resolution, see fisr_api.cu)
    __half* pool_out;         // fused 2x2 max-pool of the (post-ReLU) activation output: [N, H/2, W/2, pool_cs] planes, or nullptr
    unsigned long long pool_plane;
    int pool_cs;
    const __half* mask;       // dgrad: hi plane of the forward activation whose ReLU gradient gates this output, or nullptr
    int mask_cs, mask_off;
    // Output pixel (n, y, x) of the launch has pixel index n * opix_n + y * opix_y + x * opix_x in every output / residual / mask
    // buffer (H * W, W, 1 for a plain NHWC image).  Other values let a launch write a strided sub-lattice of a larger image: the
    // PWC-Net context network runs a dilation-d conv as d*d undilated convs on the polyphase sub-images (pwc_api.cu).
    int opix_n, opix_y, opix_x;
    float act_slope;          // > 0: leaky ReLU with this negative slope on the activation output (act_relu must be 0); split mode only
    int store_cout;           // > 0: the wide epilogue stores only channels < store_cout (outputs that sit inside a wider buffer)
};

}  // namespace fisr
