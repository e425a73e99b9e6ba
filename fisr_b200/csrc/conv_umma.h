// Host-visible launch record of the tcgen05 3x3 conv kernel (conv_umma.cu).
#pragma once
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

cudaError_t conv3x3_init();
cudaError_t launch_conv3x3(const ConvLaunch& L, int num_sms, cudaStream_t stream);

}  // namespace fisr
