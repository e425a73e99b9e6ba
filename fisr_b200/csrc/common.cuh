// Shared helpers for the fisr_b200 CUDA sources.
#pragma once
#include <cstdint>
#include <cstdio>
#include <string>
#include <cuda.h>
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fisr {

// Precision modes of the conv stack (DESIGN.md "Numerics").
//   F16X3: activations and weights are stored as an fp16 (hi, lo) pair, x = hi + lo (22-bit mantissa);
//          each K-slice issues 3 MMAs hi*hi + lo*hi + hi*lo with fp32 TMEM accumulation (fp32-class result).
//   F16  : hi plane only, 1 MMA per K-slice (fast mode, ~7e-4 max-abs on the 138-conv cascade).
enum Precision { PREC_F16X3 = 0, PREC_F16 = 1 };

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
    int P, TH, TW;            // patch pitch (= TW + 2), output rows / cols per tile
    int inv_p;                // ceil(2^20 / P): q / P == (q * inv_p) >> 20 for q < 2^20 / P
    int tiles_x, tiles_y, num_tiles;
    int raw_cs, raw_off0, raw_off1, raw_split;   // fp32 output channel stride / channel map
    int act_cs, act_off0, act_off1, act_split;   // fp16 output channel stride / channel map
    int act_relu, act_d2s, scalar_out, res_cs;
    int a_plane_bytes, a_stages, b_slots;        // smem carve-up (host computed)
    const __half* mask;       // dgrad: hi plane of the forward activation whose ReLU gradient gates this output, or nullptr
    int mask_cs, mask_off;
};

}  // namespace fisr
