// Host-visible launch record of the tcgen05 weight-gradient kernel (wgrad_umma.cu).
#pragma once
#include "common.cuh"

namespace fisr {

struct WgradArgs {
    float* partial;           // [slots][9][cin_pad][cout_pad] fp32 partial sums (slot = 2 * split + {hi, lo} lane half)
    float* bias_partial;      // [S][cout_pad] per-split column sums of dy (bias gradient), or nullptr
    int* err;                 // device error flag
    int N, H, W;
    int x_coff, dy_coff;      // first channel of block 0 inside the x / dy buffers
    int CB, OB;               // 64-channel blocks of Cin (x) and Cout (dy)
    int S;                    // pixel-range splits (every split accumulates its own partial slot pair)
    int tiles_x, tiles_y, num_tiles;
    int cin_pad, cout_pad;
    unsigned lbo_a;           // descriptor field: byte distance >> 4 between the dy hi and lo smem tiles
};

struct WgradLaunch {
    CUtensorMap tmX_hi, tmX_lo, tmD_hi, tmD_lo;
    WgradArgs args;
    int grid;
    bool x_lo;                // stage and multiply the lo plane of x too (small layers)
    size_t bias_offset;       // bias partials live behind the weight partials in the same buffer
    size_t partial_floats;    // elements the partial buffer must hold (weights + bias)
};

constexpr int kWgTH = 8, kWgTW = 16;            // output pixels per tile: 8 rows x 16 columns
constexpr int kWgXBox = (kWgTH + 2) * (kWgTW + 2) * 128;   // bytes of one x patch plane (180 pixel rows)
constexpr int kWgXPlane = (kWgXBox + 1023) / 1024 * 1024;
constexpr int kWgDPlane = kWgTH * kWgTW * 128;  // bytes of one dy tile plane (128 pixel rows)
constexpr int kWgStagesXlo = 2, kWgStagesHi = 3;     // pipeline depth with / without the x lo plane
constexpr int kWgSmemXlo = 1024 + kWgStagesXlo * (2 * kWgXPlane + 2 * kWgDPlane);
constexpr int kWgSmemHi = 1024 + kWgStagesHi * (kWgXPlane + 2 * kWgDPlane);
constexpr long long kWgXloMaxPixels = 16384;        // layers with fewer pixels also multiply by x's lo plane

// Fills the geometry fields (tile grid, splits, grid size) for an N x H x W layer with CB x OB channel blocks.
// exact: multiply by both planes of x whatever the layer size (fisr_set_wgrad_exact).
void plan_wgrad(int N, int H, int W, int CB, int OB, int num_sms, bool exact, WgradLaunch* L);
cudaError_t wgrad3x3_init();
cudaError_t launch_wgrad3x3(const WgradLaunch& L, cudaStream_t stream);
// g[tap][ci][co] = scale * sum over slots of partial (ci < cin, co < cout), HWIO like the parameter itself;
// gb[co] = scale * sum over splits of the bias partials (gb may be nullptr).
void launch_wgrad_reduce(const WgradLaunch& L, const float* partial, int cin, int cout, float scale, float* g, float* gb,
                         cudaStream_t st);

}  // namespace fisr
