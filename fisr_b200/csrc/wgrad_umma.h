// Host-visible launch record of the tcgen05 weight-gradient kernel (wgrad_umma.cu).
#pragma once
#include "common.cuh"

namespace fisr {

struct WgradArgs {
    float* partial;           // [slots][9][cin_pad][cout_pad] fp32 partial sums (slot = 2 * split + {hi, lo} lane half)
    int* err;                 // device error flag
    int N, H, W;
    int x_coff, dy_coff;      // first channel of block 0 inside the x / dy buffers
    int CB, OB;               // 64-channel blocks of Cin (x) and Cout (dy)
    int S;                    // pixel-range splits (every split accumulates its own partial slot pair)
    int tiles_x, tiles_y, num_tiles;
    int cin_pad, cout_pad;
    unsigned lbo_a;           // descriptor field: byte distance >> 4 between the dy hi and lo smem tiles
    int variant;              // probe knob (0 = production encoding)
};

struct WgradLaunch {
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    WgradArgs args;
    int grid;
    size_t partial_floats;    // elements the partial buffer must hold
};

constexpr int kWgTH = 8, kWgTW = 16;            // output pixels per tile: 8 rows x 16 columns
constexpr int kWgXBox = (kWgTH + 2) * (kWgTW + 2) * 128;   // bytes of one x patch plane (180 pixel rows)
constexpr int kWgXPlane = (kWgXBox + 1023) / 1024 * 1024;
constexpr int kWgDPlane = kWgTH * kWgTW * 128;  // bytes of one dy tile plane (128 pixel rows)
constexpr int kWgStages = 2;
constexpr int kWgSmem = 1024 + kWgStages * 2 * (kWgXPlane + kWgDPlane);

// Fills the geometry fields (tile grid, splits, grid size) for an N x H x W layer with CB x OB channel blocks.
void plan_wgrad(int N, int H, int W, int CB, int OB, int num_sms, WgradLaunch* L);
cudaError_t wgrad3x3_init();
cudaError_t launch_wgrad3x3(const WgradLaunch& L, cudaStream_t stream);
// g[tap][ci][co] = scale * sum over slots of partial (ci < cin, co < cout), HWIO like the parameter itself.
void launch_wgrad_reduce(const float* partial, int slots, int cin_pad, int cout_pad, int cin, int cout, float scale, float* g,
                         cudaStream_t st);
// gb[co] = scale * sum over pixels of dy[pix, coff + co] (hi + lo planes); workspace holds nblk * cout partials.
size_t bias_grad_workspace(size_t npix, int cout);
void launch_bias_grad(const __half* dy, size_t plane, int cs, int coff, size_t npix, int cout, float scale, float* workspace,
                      float* gb, cudaStream_t st);

}  // namespace fisr
