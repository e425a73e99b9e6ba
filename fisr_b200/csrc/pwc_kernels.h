// Host launchers of the PWC-Net inference kernels that are not 3x3 tensor-core convs (pwc_kernels.cu).
#pragma once
#include <cuda_fp16.h>
#include <cuda_runtime.h>

namespace fisr {
namespace pwc {

// Activations live in the conv kernels' split format: NHWC, an fp16 hi plane at p and an fp16 lo plane at p + plane, value =
// hi + lo (22-bit mantissa, common.cuh), so that every stride-1 3x3 conv runs on the tcgen05 kernel without a conversion pass.
struct Planes {
    __half* p; size_t plane; int cs;       // cs channels per pixel
};

// One 3x3 convolution on the CUDA cores (stride-2 / small / irregular layers): tensors addressed as (planes, channel offset).
struct PwcConv {
    Planes in; int in_coff, cin;                    // reads channels [in_coff, in_coff + cin) of every input pixel
    Planes out; int out_coff, cout;                 // writes channels [out_coff, out_coff + cout)      (wide kernel)
    float* out_f32;                                 // narrow kernel (cout = 2, the flow predictors): fp32 [.., 2] output
    const float* w;                                 // [9][cin][cout] (HWIO with the 3x3 taps flattened)
    const float* b;                                 // [cout]
    int N, Hin, Win, Hout, Wout, stride, dil, pad_y, pad_x;
    int leaky;                                      // leaky ReLU 0.1 on the output
    const float* add; int add_cs;                   // narrow kernel only: fp32 residual input [.., add_cs] added before the activation
};

cudaError_t init_kernels();
void launch_conv3x3(const PwcConv& p, cudaStream_t st);
// conv2d_transpose 4x4 s2, 2 filters: input either planes (in.p != nullptr) or fp32 [.., in_f32_cs]; writes 2 channels at out_coff
void launch_deconv4x4s2(Planes in, int in_coff, const float* in_f32, int in_f32_cs, int cin, const float* w, const float* b, Planes out, int out_coff,
                        int N, int h, int wd, cudaStream_t st);
void launch_cost_volume(Planes c1, int c1_coff, Planes c2, int c2_coff, int C, Planes out, int out_coff, int N, int h, int w, cudaStream_t st);
void launch_dense_warp(Planes img, int coff, int C, Planes flow, int f_coff, float scale, Planes out, int N, int h, int w, cudaStream_t st);
// conv1a of the feature pyramid on the fp32 input image: w [9][3][16], b [16] -> out [N, H/2, W/2, 16]
void launch_first_conv(const float* img, const float* w, const float* b, Planes out, int N, int H, int W, cudaStream_t st);
// splits the 16 columns of the fused predict_flow / up_feat conv: fp32 flow [N,h,w,2] and (Dn.p != nullptr) the 2 up_feat channels of level l - 1
void launch_flow_upfeat_scatter(Planes F, float* flow, Planes Dn, int up_off, int N, int h, int w, cudaStream_t st);
// pre / post-processing of the reference's driver (float64 arithmetic in numpy's evaluation order)
struct PrepConst { double T[9], off[3]; double ry, rx; };          // YUV -> RGB matrix / offset (utils.py:106-115), n_in / n_out per axis
constexpr int kMaxGaussRadius = 8;
struct FinishConst { double wy[kMaxGaussRadius + 1], wx[kMaxGaussRadius + 1]; int ry, rx; double ratio_y, ratio_x, scale; };   // w[j] = weight at distance j
void launch_prepare_pair(const void* src0, const void* src1, bool yuv, float* img1, float* img2, int h, int w, int oh, int ow, int Hp, int Wp,
                         const PrepConst& k, cudaStream_t st);
void launch_finish_flow(const float* flow, int N, int Hp, int Wp, int h0, int w0, int oh, int ow, float* out, const FinishConst& k, cudaStream_t st);
// dst[.., dst_off + c] = src[.., src_off + c] for c < C (C, offsets and channel strides multiples of 8)
void launch_copy_channels(Planes src, int src_off, Planes dst, int dst_off, int C, long long npix, cudaStream_t st);
void launch_resize_flow(const float* in, float* out, int N, int h, int w, int S, float gain, cudaStream_t st);

}  // namespace pwc
}  // namespace fisr
