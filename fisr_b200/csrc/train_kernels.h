// Host launchers of the training-side kernels (train_kernels.cu).
#pragma once
#include "common.cuh"
#include "aux_kernels.h"

namespace fisr {

struct LossLambdas {      // main.py:80-85
    float recn, tm1, tm2, tmm, td, ss2;
};
struct LossScales {       // block-partial layout of the three scales (l1, l2, l3)
    long long hw[3];
    int nblk[3];
    size_t offset[3];
};

void launch_assemble_passes(const float* data, const float* flow, const float* flow2, const float* warp, const float* warp2,
                            float* out, int B, int h, int w, cudaStream_t st);
void launch_groups2ovlp(const float* pred, float* out, int B, int H, int W, cudaStream_t st);
size_t temporal_loss_workspace(int B, int h, int w, LossScales* sc);
void launch_temporal_loss(const float* const pred[3], const float* label, int B, int h, int w, const LossLambdas& lam,
                          double* workspace, float* d_out, cudaStream_t st);
// d total_loss / d pred of one scale (times `scale`), + the gradient arriving through the next level's input channels
// 29..37 (`extra`, may be empty), written as the 64-channel (hi, lo) dy operand of the two conv/2 heads.
void launch_loss_grad(const float* pred, const float* label, ActBuf extra, int B, int hs, int ws, int st, int LH, int LW,
                      float wgt, const LossLambdas& lam, float scale, ActBuf out, cudaStream_t stm);
void launch_pool_bwd(ActBuf skip, ActBuf gcat, int cs, int coff, ActBuf gpool, ActBuf gout, float* rout, int N, int H, int W,
                     int C, cudaStream_t st);
void launch_upsample_bwd(ActBuf gup, ActBuf xin, ActBuf gout, float* rout, int N, int h, int w, int C, cudaStream_t st);
void launch_prep_weights_dgrad(const float* w, __half* out, int cin, int cout, int KBo, int cin_pad, cudaStream_t st);
void launch_grad_absmax(const float* g, size_t n, unsigned* out, cudaStream_t st);
// Multi-tensor optimiser step: device tables over all 276 tensors, kMtChunk elements per block, block -> tensor by block0.
constexpr int kMtChunk = 4096;
struct MtTensor {
    float* theta; const float* g; float* m; float* v;
    unsigned long long n;
    unsigned block0;
};
struct MtPack {          // operand re-pack of one conv (split mode): forward planes or rotated-transposed dgrad planes
    const float* w; __half* out;
    int cin, cout, pad, dgrad;        // pad = cout_pad (forward) or cin_pad (dgrad)
    unsigned long long per_plane;
    unsigned block0;
};
void launch_mt_absmax(const MtTensor* d_table, int n, unsigned blocks, unsigned* out, cudaStream_t st);
void launch_mt_adam(const MtTensor* d_table, int n, unsigned blocks, float lr_t, float beta1, float beta2, float eps, cudaStream_t st);
void launch_mt_repack(const MtPack* d_table, int n, unsigned blocks, cudaStream_t st);
void launch_adam_tf1(float* theta, const float* g, float* m, float* v, size_t n, float lr_t, float beta1, float beta2, float eps,
                     cudaStream_t st);

}  // namespace fisr
