// Host-visible launch record of the tcgen05 3x3 conv kernel (conv_umma.cu).
#pragma once
#include <string>
#include "common.cuh"

namespace fisr {

// opt-in dynamic shared memory budget per CTA: 227 KB (232448 B) minus 512 B for the kernel's static barriers
constexpr int kConvMaxSmem = 231936;

struct ConvLaunch {
    CUtensorMap tmA_hi, tmA_lo, tmB;
    CUtensorMap tmO_hi, tmO_8;   // f16f8 activation-only epilogue: TMA-store maps of the fp16 plane and of the 8-bit rows
    bool tma_out;                // the two maps above are valid
    ConvArgs args;
    int NT, chunks, planes;
    bool pair;              // f16f8 on CTA pairs (cluster of 2, M = 256 MMAs): args.tiles_x / num_tiles count PAIRS of x-adjacent tiles
    int epi;                // epilogue variant: 1 residual in, 2 fp32 out, 4 depth-to-space (convk::EPI_*)
    int smem_bytes;
    double efficiency;      // useful fraction of the MMA rows issued (tile quantisation at the image edges)
};

// Chooses NT / chunk count / patch pitch for an H x W x n_img conv and fills the geometry fields of L->args.
bool plan_conv_geometry(int H, int W, int n_img, int cout_pad, int planes, int num_sms, ConvLaunch* L, int kb = 0);

// A split-mode (f16x3: fp16 hi / lo planes, fp32-class result) conv on strided views, for callers outside the FISRnet plan
// builder (the PWC-Net flow network, pwc_api.cu).  Input: channels [cin_off, cin_off + cin) of an fp16-plane buffer with in_cs
// channels per pixel, pixels / rows / images in_sx / in_sy / in_sn ELEMENTS apart; output: channels [out_off, out_off + cout) of
// pixel n * opix_n + y * opix_y + x * opix_x of a buffer with out_cs channels per pixel.  Weights are the planes written by
// launch_prep_weights(.., planes = 2) for KB = ceil(cin / 64) blocks and cout_pad columns; bias has cout_pad entries.
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);
struct SplitConvDesc {
    const __half* in; size_t in_plane; int in_cs;
    long long in_sx, in_sy, in_sn;
    int cin_off, cin;
    __half* out; size_t out_plane; int out_cs, out_off;
    long long opix_n, opix_y, opix_x;
    long long out_pixels;        // pixel count of the whole output buffer (32-bit offset check)
    const __half* wp; const float* bias; int cout, cout_pad;
    int N, H, W;
    int relu; float slope;       // relu = 1: ReLU; slope > 0: leaky ReLU; neither: linear
    // optional fp32 side tensors, both indexed with the output pixel index and `cout` floats per pixel (+ cout_pad of slack at the end):
    float* raw;                  // pre-activation result (incl. bias and residual) is also written here
    const float* res;            // added before the activation -- a conv split over two launches accumulates through raw -> res
};
// cout_pad a layer's weights must be packed for (independent of the image size)
inline int split_conv_cout_pad(int cout) { return cout <= 16 ? 16 : cout <= 64 ? 64 : cout <= 96 ? 96 : (cout + 127) / 128 * 128; }
bool build_split_conv(EncodeTiledFn encode, const SplitConvDesc& d, int num_sms, int* d_err, ConvLaunch* L, std::string* why);

cudaError_t conv3x3_init();
cudaError_t launch_conv3x3(const ConvLaunch& L, int num_sms, cudaStream_t stream);

}  // namespace fisr
