// Host launchers of the training-side kernels (train_kernels.cu).
#pragma once
#include "common.cuh"

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
void launch_adam_tf1(float* theta, const float* g, float* m, float* v, size_t n, float lr_t, float beta1, float beta2, float eps,
                     cudaStream_t st);

}  // namespace fisr
