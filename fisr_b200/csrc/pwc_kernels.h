// Host launchers of the PWC-Net inference kernels (pwc_kernels.cu).
#pragma once
#include <cuda_runtime.h>

namespace fisr {
namespace pwc {

// One 3x3 convolution on NHWC fp32 tensors addressed as (pointer, channel stride, channel offset).
struct PwcConv {
    const float* in; int in_cs, in_coff, cin;       // reads channels [in_coff, in_coff + cin) of every input pixel
    float* out; int out_cs, out_coff, cout;         // writes channels [out_coff, out_coff + cout)
    const float* w;                                 // [9][cin][cout] (HWIO with the 3x3 taps flattened)
    const float* b;                                 // [cout]
    int N, Hin, Win, Hout, Wout, stride, dil, pad_y, pad_x;
    int leaky;                                      // leaky ReLU 0.1 on the output
    const float* add; int add_cs;                   // narrow kernel only: residual input [.., add_cs] added before the activation
};

cudaError_t init_kernels();
void launch_conv3x3(const PwcConv& p, cudaStream_t st);
void launch_deconv4x4s2(const float* in, int in_cs, int in_coff, int cin, const float* w, const float* b, float* out, int out_cs, int out_coff,
                        int N, int h, int wd, cudaStream_t st);
void launch_cost_volume(const float* c1, int c1_cs, int c1_coff, const float* c2, int c2_cs, int c2_coff, int C, float* out, int out_cs, int out_coff,
                        int N, int h, int w, cudaStream_t st);
void launch_dense_warp(const float* img, int cs, int coff, int C, const float* flow, int f_cs, int f_coff, float scale, float* out, int N, int h, int w,
                       cudaStream_t st);
void launch_resize_flow(const float* in, float* out, int N, int h, int w, int S, float gain, cudaStream_t st);

}  // namespace pwc
}  // namespace fisr
