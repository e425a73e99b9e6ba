// Training-side kernels that do not need the backward pass: window assembly of the four weight-shared passes,
// Groups2Ovlp, the multi-scale temporal loss (forward values) + train PSNR, and the TF-1.13 Adam update.
// All are HBM-bound single-pass kernels; reductions are two-stage with a fixed summation order (deterministic).
#include "common.cuh"
#include "act_io.cuh"
#include "train_kernels.h"

namespace fisr {

namespace {

inline unsigned blocks_for(size_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

// ---------------------------------------------------------------- window assembly (FISRnet.py:281-306, 392-399)
// out [4B,h,w,29]: pass p < 3 = stride-1 window p (frames ch [3p,3p+9), flows [4p,4p+8), warps [6p,6p+12));
// pass 3 = stride 2 (frames 0,2,4 + flow_ss2 + warp_ss2).
__global__ void assemble_passes_kernel(const float* __restrict__ data, const float* __restrict__ flow,
                                       const float* __restrict__ flow2, const float* __restrict__ warp,
                                       const float* __restrict__ warp2, float* __restrict__ out, int B, size_t hw) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(4) * B * hw * 29;
    if (i >= total) return;
    const int c = i % 29;
    const size_t pix = i / 29;                 // over [4B, h, w]
    const size_t q = pix % (static_cast<size_t>(B) * hw);
    const int pass = pix / (static_cast<size_t>(B) * hw);
    float v;
    if (pass < 3) {
        if (c < 9) v = data[q * 15 + 3 * pass + c];
        else if (c < 17) v = flow[q * 16 + 4 * pass + (c - 9)];
        else v = warp[q * 24 + 6 * pass + (c - 17)];
    } else {
        if (c < 9) v = data[q * 15 + (c / 3) * 6 + c % 3];
        else if (c < 17) v = flow2[q * 8 + (c - 9)];
        else v = warp2[q * 12 + (c - 17)];
    }
    out[i] = v;
}

// ---------------------------------------------------------------- Groups2Ovlp (ops.py:119-144)
// pred [3B,H,W,9] (window-major) -> out [B,7,H,W,3] = [f0, f1, (f2+f3)/2, f4, (f5+f6)/2, f7, f8]
__global__ void groups2ovlp_kernel(const float* __restrict__ pred, float* __restrict__ out, int B, size_t hw) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(B) * 7 * hw * 3;
    if (i >= total) return;
    const int c = i % 3;
    size_t r = i / 3;
    const size_t p = r % hw; r /= hw;
    const int f = r % 7;
    const int b = r / 7;
    auto P = [&](int k) { return pred[((static_cast<size_t>(k / 3) * B + b) * hw + p) * 9 + 3 * (k % 3) + c]; };
    float v;
    switch (f) {
        case 0: v = P(0); break;
        case 1: v = P(1); break;
        case 2: v = (P(2) + P(3)) / 2; break;
        case 3: v = P(4); break;
        case 4: v = (P(5) + P(6)) / 2; break;
        case 5: v = P(7); break;
        default: v = P(8); break;
    }
    out[i] = v;
}

// ---------------------------------------------------------------- temporal loss (FISRnet.py:312-486)
// One scale: pred [4B,hs,ws,9] (pass-major), label [B,2h,2w,21] sampled with stride st ("bicubic" = subsample, :263-264).
// Per pixel the 7 squared-error sums + the 7 per-frame sums of (O - GT)^2 for the PSNR; block partials
// partial[(b * nblk + blk) * 14 + k].
constexpr int kLossTerms = 14;
__global__ void temporal_loss_kernel(const float* __restrict__ pred, const float* __restrict__ label, int B, int hs, int ws,
                                     int st, int LH, int LW, double* __restrict__ partial) {
    const int b = blockIdx.y;
    const size_t hw = static_cast<size_t>(hs) * ws;
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    float acc[kLossTerms];
#pragma unroll
    for (int k = 0; k < kLossTerms; ++k) acc[k] = 0.f;
    if (p < hw) {
        const int y = p / ws, x = p % ws;
        const float* lab = label + ((static_cast<size_t>(b) * LH + static_cast<size_t>(y) * st) * LW + static_cast<size_t>(x) * st) * 21;
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            float P[9], S[3], G[7], O[7];
#pragma unroll
            for (int k = 0; k < 9; ++k) P[k] = pred[((static_cast<size_t>(k / 3) * B + b) * hw + p) * 9 + 3 * (k % 3) + c];
#pragma unroll
            for (int k = 0; k < 3; ++k) S[k] = pred[((static_cast<size_t>(3) * B + b) * hw + p) * 9 + 3 * k + c];
#pragma unroll
            for (int k = 0; k < 7; ++k) G[k] = lab[3 * k + c];
            O[0] = P[0]; O[1] = P[1]; O[2] = (P[2] + P[3]) / 2; O[3] = P[4]; O[4] = (P[5] + P[6]) / 2; O[5] = P[7]; O[6] = P[8];
            float d;
#pragma unroll
            for (int w = 0; w < 3; ++w)
#pragma unroll
                for (int f = 0; f < 3; ++f) { d = P[3 * w + f] - G[2 * w + f]; acc[0] += d * d; }             // recn  :315-328
#pragma unroll
            for (int i = 0; i < 2; ++i) {
                d = P[3 * i + 2] - P[3 * i + 3]; acc[1] += d * d;                                             // tm    :331-341
                d = (P[3 * i + 2] + P[3 * i + 3]) / 2 - G[2 * (i + 1)]; acc[2] += d * d;                       // tmm   :344-357
            }
#pragma unroll
            for (int i = 0; i < 6; ++i) { d = (O[i + 1] - O[i]) - (G[i + 1] - G[i]); acc[3] += d * d; }         // td    :360-385
#pragma unroll
            for (int f = 0; f < 3; ++f) {
                d = S[f] - G[2 * f + 1]; acc[4] += d * d;                                                     // recn2 :426-428
                d = S[f] - O[2 * f + 1]; acc[6] += d * d;                                                     // tm2   :462-477
            }
#pragma unroll
            for (int i = 0; i < 2; ++i) { d = (S[i + 1] - S[i]) - (G[2 * i + 3] - G[2 * i + 1]); acc[5] += d * d; }   // td2 :431-459
#pragma unroll
            for (int f = 0; f < 7; ++f) { d = O[f] - G[f]; acc[7 + f] += d * d; }                               // PSNR  :485
        }
    }
    __shared__ double red[kLossTerms][8];
    const int lane = threadIdx.x & 31, wid = threadIdx.x >> 5;
#pragma unroll
    for (int k = 0; k < kLossTerms; ++k) {
        double v = acc[k];
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (lane == 0) red[k][wid] = v;
    }
    __syncthreads();
    if (threadIdx.x < kLossTerms) {
        double v = 0;
        for (int w = 0; w < static_cast<int>(blockDim.x >> 5); ++w) v += red[threadIdx.x][w];
        partial[(static_cast<size_t>(b) * gridDim.x + blockIdx.x) * kLossTerms + threadIdx.x] = v;
    }
}

// Fixed-order reduction of the block partials of the three scales and the 11 scalars of FISRnet.py:651-657.
// One block of 32 warps; warp tasks: 21 (scale, term) sums over images and blocks, then B x 7 per-(image, frame) sums
// for the PSNR.  Lane-strided partial sums + a shuffle tree: the summation order is fixed (deterministic).
__global__ void loss_finalize_kernel(const double* __restrict__ partial, LossScales sc, int B, LossLambdas lam,
                                     float* __restrict__ out) {
    __shared__ double sums[3][7];
    __shared__ double psnr_db[32];
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarp = blockDim.x >> 5;
    double db_acc = 0;
    for (int task = warp; task < 21 + 7 * B; task += nwarp) {
        double v = 0;
        if (task < 21) {                  // 7 terms x 3 scales: summed over images and blocks
            const int s = task / 7, k = task % 7;
            const double* p = partial + sc.offset[s];
            const size_t n = static_cast<size_t>(B) * sc.nblk[s];
            for (size_t i = lane; i < n; i += 32) v += p[i * kLossTerms + k];
        } else {                          // train PSNR on the finest scale: per (image, frame) squared error
            const int b = (task - 21) / 7, f = (task - 21) % 7;
            const double* p = partial + sc.offset[2] + static_cast<size_t>(b) * sc.nblk[2] * kLossTerms + 7 + f;
            for (int i = lane; i < sc.nblk[2]; i += 32) v += p[static_cast<size_t>(i) * kLossTerms];
        }
#pragma unroll
        for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
        if (task < 21) { if (lane == 0) sums[task / 7][task % 7] = v; }
        else db_acc += 10.0 * log10(1.0 / (v / (static_cast<double>(sc.hw[2]) * 3)));      // MSE -> dB (tf.image.psnr, max_val 1)
    }
    if (lane == 0) psnr_db[warp] = db_acc;
    __syncthreads();
    if (threadIdx.x == 0) {
        double psnr_sum = 0;
        for (int w = 0; w < nwarp; ++w) psnr_sum += psnr_db[w];
        psnr_sum /= 7.0 * B;
        const double wgt[3] = {4.0, 2.0, 1.0};                      // l1, l2, l3  (FISRnet.py:326-328)
        double term[7] = {0, 0, 0, 0, 0, 0, 0};
        for (int s = 0; s < 3; ++s) {
            const double n1 = static_cast<double>(B) * sc.hw[s] * 3;      // elements of one [B,1,h,w,3] slice
            const double cnt[7] = {3 * n1, n1, n1, n1, 3 * n1, n1, 3 * n1};
            for (int k = 0; k < 7; ++k) term[k] += wgt[s] * sums[s][k] / cnt[k];
        }
        const double s1 = lam.recn * term[0] + lam.tm1 * term[1] + lam.tmm * term[2] + lam.td * term[3];
        const double s2 = lam.recn * term[4] + lam.td * term[5] + lam.tm2 * term[6];
        out[0] = term[0]; out[1] = term[1]; out[2] = term[2]; out[3] = term[3]; out[4] = s1;
        out[5] = term[4]; out[6] = term[5]; out[7] = term[6]; out[8] = s2;
        out[9] = s1 + lam.ss2 * s2;
        out[10] = psnr_sum;
    }
}

// ---------------------------------------------------------------- d total_loss / d pred (FISRnet.py:312-484, autodiff of)
// One scale.  Thread = (image b, pixel p): the 36 derivatives wrt the 9 window frames P and the 3 stride-2 frames S,
// times the loss scale, plus the gradient that reached this prediction through the next level's input channels
// 29..37 (FISRnet.py:113,144: pred feeds the next level un-detached).  Written as the (hi, lo) dy operand of the two
// conv/2 heads: channel j of pred lands in channel perm(j) of a 128-channel buffer, [0,6) = FI-SR's outputs
// (pred 0,1,2,6,7,8), [64,67) = SR's (pred 3,4,5)  (FISRnet.py:107-108); each head's K block starts 128-B aligned.
struct LossGradCoef { float c[7]; };   // 2 * scale * weight / count of recn, tm, tmm, td, recn2, td2, tm2
__global__ void loss_grad_kernel(const float* __restrict__ pred, const float* __restrict__ label, const __half* __restrict__ extra,
                                 size_t extra_plane, int B, int hs, int ws, int st, int LH, int LW, LossGradCoef k,
                                 __half* __restrict__ out, size_t out_plane) {
    const int b = blockIdx.y;
    const size_t hw = static_cast<size_t>(hs) * ws;
    const size_t p = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (p >= hw) return;
    const int y = p / ws, x = p % ws;
    const float* lab = label + ((static_cast<size_t>(b) * LH + static_cast<size_t>(y) * st) * LW + static_cast<size_t>(x) * st) * 21;
    float gr[12][3];                   // d loss / d (frame q, colour c), q = 9 window frames then 3 stride-2 frames
#pragma unroll
    for (int c = 0; c < 3; ++c) {
        float P[9], S[3], G[7], O[7], dP[9], dS[3], dO[7];
#pragma unroll
        for (int q = 0; q < 9; ++q) P[q] = pred[((static_cast<size_t>(q / 3) * B + b) * hw + p) * 9 + 3 * (q % 3) + c];
#pragma unroll
        for (int q = 0; q < 3; ++q) S[q] = pred[((static_cast<size_t>(3) * B + b) * hw + p) * 9 + 3 * q + c];
#pragma unroll
        for (int q = 0; q < 7; ++q) { G[q] = lab[3 * q + c]; dO[q] = 0.f; }
        O[0] = P[0]; O[1] = P[1]; O[2] = (P[2] + P[3]) / 2; O[3] = P[4]; O[4] = (P[5] + P[6]) / 2; O[5] = P[7]; O[6] = P[8];
#pragma unroll
        for (int w = 0; w < 3; ++w)
#pragma unroll
            for (int f = 0; f < 3; ++f) dP[3 * w + f] = k.c[0] * (P[3 * w + f] - G[2 * w + f]);                 // recn
#pragma unroll
        for (int i = 0; i < 2; ++i) {
            const float d1 = k.c[1] * (P[3 * i + 2] - P[3 * i + 3]);                                             // tm
            const float d2 = 0.5f * k.c[2] * ((P[3 * i + 2] + P[3 * i + 3]) / 2 - G[2 * (i + 1)]);                // tmm
            dP[3 * i + 2] += d1 + d2;
            dP[3 * i + 3] += d2 - d1;
        }
#pragma unroll
        for (int i = 0; i < 6; ++i) {                                                                            // td
            const float d = k.c[3] * ((O[i + 1] - O[i]) - (G[i + 1] - G[i]));
            dO[i + 1] += d; dO[i] -= d;
        }
#pragma unroll
        for (int f = 0; f < 3; ++f) {
            const float d = k.c[6] * (S[f] - O[2 * f + 1]);                                                      // tm2
            dS[f] = k.c[4] * (S[f] - G[2 * f + 1]) + d;                                                          // recn2
            dO[2 * f + 1] -= d;
        }
#pragma unroll
        for (int i = 0; i < 2; ++i) {                                                                            // td2
            const float d = k.c[5] * ((S[i + 1] - S[i]) - (G[2 * i + 3] - G[2 * i + 1]));
            dS[i + 1] += d; dS[i] -= d;
        }
        dP[0] += dO[0]; dP[1] += dO[1]; dP[2] += 0.5f * dO[2]; dP[3] += 0.5f * dO[2]; dP[4] += dO[3];             // Groups2Ovlp
        dP[5] += 0.5f * dO[4]; dP[6] += 0.5f * dO[4]; dP[7] += dO[5]; dP[8] += dO[6];
#pragma unroll
        for (int q = 0; q < 12; ++q) gr[q][c] = q < 9 ? dP[q] : dS[q - 9];
    }
    // one image (pass, b) at a time: pred channels j = 3f + c -> FI-SR's dy in channels 0..5 (pred 0,1,2,6,7,8; one
    // 16-byte store per plane), SR's in channels 64..66 (pred 3,4,5; one 8-byte store per plane)
#pragma unroll
    for (int pass = 0; pass < 4; ++pass) {
        const size_t img = (static_cast<size_t>(pass) * B + b) * hw + p;
        float g[9];
#pragma unroll
        for (int j = 0; j < 9; ++j) g[j] = gr[pass * 3 + j / 3][j % 3];
        if (extra) {
#pragma unroll
            for (int j = 0; j < 9; ++j) g[j] += join_f16(extra[img * 64 + 29 + j], extra[extra_plane + img * 64 + 29 + j]);
        }
        const float fi[8] = {g[0], g[1], g[2], g[6], g[7], g[8], 0.f, 0.f};
        store8<2>(out + img * 128, out_plane, fi);
        uint32_t h01, l01, h23, l23;
        split2_f32(g[3], g[4], h01, l01);
        split2_f32(g[5], 0.f, h23, l23);
        *reinterpret_cast<uint2*>(out + img * 128 + 64) = make_uint2(h01, h23);
        *reinterpret_cast<uint2*>(out + out_plane + img * 128 + 64) = make_uint2(l01, l23);
    }
}

// ---------------------------------------------------------------- max-pool backward + skip add (ops.py:54, autodiff of)
// skip = channels [coff, coff+C) of the concat buffer (post-ReLU); g_cat = gradient that arrived through the concat
// (same indexing, already gated by [skip > 0]); g_pool = gradient wrt the pooled map.  Each 2x2 window routes g_pool to
// its first maximum (TF's MaxPoolGrad order); windows whose maximum is 0 route nothing (the ReLU before the pool).
__global__ void pool_bwd_kernel(const __half* __restrict__ skip, size_t pskip, const __half* __restrict__ gcat, size_t pgcat, int cs,
                                int coff, const __half* __restrict__ gpool, size_t pgpool, __half* __restrict__ gout, size_t pgout,
                                float* __restrict__ rout, int N, int H, int W, int C) {
    const int cv = C / 8, h = H / 2, w = W / 2;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * h * w * cv;
    if (i >= total) return;
    const int c8 = (i % cv) * 8;
    size_t r = i / cv;
    const int x = r % w; r /= w;
    const int y = r % h;
    const int n = r / h;
    float gp[8], v[4][8];
    load8<2>(gpool + ((static_cast<size_t>(n) * h + y) * w + x) * C + c8, pgpool, gp);
    int arg[8];
    float m[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) { m[j] = 0.f; arg[j] = -1; }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t pix = (static_cast<size_t>(n) * H + 2 * y + (k >> 1)) * W + 2 * x + (k & 1);
        load8<2>(skip + pix * cs + coff + c8, pskip, v[k]);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (v[k][j] > m[j]) { m[j] = v[k][j]; arg[j] = k; }
    }
#pragma unroll
    for (int k = 0; k < 4; ++k) {
        const size_t pix = (static_cast<size_t>(n) * H + 2 * y + (k >> 1)) * W + 2 * x + (k & 1);
        float g[8];
        load8<2>(gcat + pix * cs + coff + c8, pgcat, g);
#pragma unroll
        for (int j = 0; j < 8; ++j)
            if (arg[j] == k) g[j] += gp[j];
        store8<2>(gout + pix * C + c8, pgout, g);
        float4* ro = reinterpret_cast<float4*>(rout + pix * C + c8);
        ro[0] = make_float4(g[0], g[1], g[2], g[3]);
        ro[1] = make_float4(g[4], g[5], g[6], g[7]);
    }
}

// ---------------------------------------------------------------- legacy bilinear x2 backward (ops.py:69, autodiff of)
// Adjoint of out[2k] = in[k], out[2k+1] = (in[k] + in[min(k+1, n-1)]) / 2 along H then W, gated by [x > 0] of the
// (post-ReLU) tensor that was upsampled.
__global__ void upsample_bwd_kernel(const __half* __restrict__ gup, size_t pgup, const __half* __restrict__ xin, size_t pxin,
                                    __half* __restrict__ gout, size_t pgout, float* __restrict__ rout, int N, int h, int w, int C) {
    const int cv = C / 8;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * h * w * cv;
    if (i >= total) return;
    const int c8 = (i % cv) * 8;
    size_t r = i / cv;
    const int x = r % w; r /= w;
    const int y = r % h;
    const int n = r / h;
    float acc[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) acc[j] = 0.f;
    const float wy[3] = {y >= 1 ? 0.5f : 0.f, 1.f, y == h - 1 ? 1.f : 0.5f};
    const float wx[3] = {x >= 1 ? 0.5f : 0.f, 1.f, x == w - 1 ? 1.f : 0.5f};
#pragma unroll
    for (int dy = 0; dy < 3; ++dy) {
        if (wy[dy] == 0.f) continue;
        const int Y = 2 * y + dy - 1;
#pragma unroll
        for (int dx = 0; dx < 3; ++dx) {
            if (wx[dx] == 0.f) continue;
            const int X = 2 * x + dx - 1;
            float g[8];
            load8<2>(gup + ((static_cast<size_t>(n) * 2 * h + Y) * 2 * w + X) * C + c8, pgup, g);
            const float wgt = wy[dy] * wx[dx];
#pragma unroll
            for (int j = 0; j < 8; ++j) acc[j] += wgt * g[j];
        }
    }
    const size_t pix = (static_cast<size_t>(n) * h + y) * w + x;
    float xv[8];
    load8<2>(xin + pix * C + c8, pxin, xv);
#pragma unroll
    for (int j = 0; j < 8; ++j)
        if (!(xv[j] > 0.f)) acc[j] = 0.f;
    store8<2>(gout + pix * C + c8, pgout, acc);
    if (rout) {
        float4* ro = reinterpret_cast<float4*>(rout + pix * C + c8);
        ro[0] = make_float4(acc[0], acc[1], acc[2], acc[3]);
        ro[1] = make_float4(acc[4], acc[5], acc[6], acc[7]);
    }
}

// ---------------------------------------------------------------- dgrad weights
// w fp32 HWIO [3,3,cin,cout] -> operand planes of the transposed, 180-degree rotated filter: the data gradient of a
// SAME 3x3 conv is the SAME 3x3 conv of dy with w'[ky,kx,co,ci] = w[2-ky,2-kx,ci,co].
// Layout [plane][kb over cout][tap][cin_pad][64] like the forward operand.
__device__ __forceinline__ void prep_dgrad_element(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout, int cin_pad,
                                                   size_t per_plane, size_t i) {
    const int cl = i & 63;
    size_t r = i >> 6;
    const int ci = r % cin_pad; r /= cin_pad;
    const int tap = r % 9;
    const int kb = r / 9;
    const int co = kb * 64 + cl;
    float v = 0.f;
    if (ci < cin && co < cout) v = w[(static_cast<size_t>(8 - tap) * cin + ci) * cout + co];
    const SplitHalf s = split_f32(v);
    out[i] = s.hi;
    out[per_plane + i] = s.lo;
}
__global__ void prep_weights_dgrad_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout, int KBo,
                                          int cin_pad) {
    const size_t per_plane = static_cast<size_t>(KBo) * 9 * cin_pad * 64;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= per_plane) return;
    prep_dgrad_element(w, out, cin, cout, cin_pad, per_plane, i);
}

// forward operand planes of the split mode (aux_kernels.cu, prep_weights_kernel with planes = 2), per element
__device__ __forceinline__ void prep_fwd_split_element(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout,
                                                       int cout_pad, size_t per_plane, size_t i) {
    const int cl = i & 63;
    size_t r = i >> 6;
    const int co = r % cout_pad; r /= cout_pad;
    const int tap = r % 9;
    const int ci = static_cast<int>(r / 9) * 64 + cl;
    float v = 0.f;
    if (ci < cin && co < cout) v = w[(static_cast<size_t>(tap) * cin + ci) * cout + co];
    const SplitHalf s = split_f32(v);
    out[i] = s.hi;
    out[per_plane + i] = s.lo;
}

// ---------------------------------------------------------------- multi-tensor optimiser step
// One launch over ALL 276 tensors (block -> tensor by binary search over MtTensor::block0, kMtChunk elements per block):
// 3 launches per step (|g| max, Adam, operand re-pack) instead of 552 + 276 + 276.
__device__ __forceinline__ int mt_find(const MtTensor* __restrict__ t, int n, unsigned blk) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid].block0 <= blk) lo = mid; else hi = mid - 1;
    }
    return lo;
}

// max |g| over every gradient tensor, as the raw bits of a non-negative float (atomicMax on ints orders them); NaN / Inf
// give bits >= 0x7F800000, which is how an overflowed loss scale is detected.
__global__ void mt_absmax_kernel(const MtTensor* __restrict__ t, int n, unsigned* __restrict__ out) {
    const MtTensor e = t[mt_find(t, n, blockIdx.x)];
    const size_t base = static_cast<size_t>(blockIdx.x - e.block0) * kMtChunk;
    unsigned m = 0;
    for (size_t i = base + threadIdx.x; i < base + kMtChunk && i < e.n; i += blockDim.x)
        m = max(m, __float_as_uint(e.g[i]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

// Adam, TF-1.13 formula (FISRnet.py:489-491): lr_t = lr sqrt(1 - b2^t) / (1 - b1^t) (host), theta -= lr_t m / (sqrt(v) + eps)
__device__ __forceinline__ void adam_element(float& theta, float g, float& m, float& v, float lr_t, float beta1, float beta2, float eps) {
    m = beta1 * m + (1.f - beta1) * g;
    v = beta2 * v + (1.f - beta2) * g * g;
    theta -= lr_t * m / (sqrtf(v) + eps);
}
__global__ void mt_adam_kernel(const MtTensor* __restrict__ t, int n, float lr_t, float beta1, float beta2, float eps) {
    const MtTensor e = t[mt_find(t, n, blockIdx.x)];
    const size_t base = static_cast<size_t>(blockIdx.x - e.block0) * kMtChunk;
    for (size_t i = base + threadIdx.x; i < base + kMtChunk && i < e.n; i += blockDim.x) {
        float th = e.theta[i], m = e.m[i], v = e.v[i];
        adam_element(th, e.g[i], m, v, lr_t, beta1, beta2, eps);
        e.theta[i] = th; e.m[i] = m; e.v[i] = v;
    }
}
__global__ void mt_repack_kernel(const MtPack* __restrict__ t, int n) {
    int lo = 0, hi = n - 1;
    while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (t[mid].block0 <= blockIdx.x) lo = mid; else hi = mid - 1;
    }
    const MtPack e = t[lo];
    const size_t base = static_cast<size_t>(blockIdx.x - e.block0) * kMtChunk;
    for (size_t i = base + threadIdx.x; i < base + kMtChunk && i < e.per_plane; i += blockDim.x) {
        if (e.dgrad) prep_dgrad_element(e.w, e.out, e.cin, e.cout, e.pad, e.per_plane, i);
        else prep_fwd_split_element(e.w, e.out, e.cin, e.cout, e.pad, e.per_plane, i);
    }
}

__global__ void grad_absmax_kernel(const float* __restrict__ g, size_t n, unsigned* __restrict__ out) {
    unsigned m = 0;
    for (size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x; i < n; i += static_cast<size_t>(gridDim.x) * blockDim.x)
        m = max(m, __float_as_uint(g[i]) & 0x7FFFFFFFu);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) m = max(m, __shfl_xor_sync(0xffffffffu, m, o));
    if ((threadIdx.x & 31) == 0 && m) atomicMax(out, m);
}

__global__ void adam_tf1_kernel(float* __restrict__ theta, const float* __restrict__ g, float* __restrict__ m,
                                float* __restrict__ v, size_t n, float lr_t, float beta1, float beta2, float eps) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= n) return;
    float th = theta[i], mi = m[i], vi = v[i];
    adam_element(th, g[i], mi, vi, lr_t, beta1, beta2, eps);
    theta[i] = th; m[i] = mi; v[i] = vi;
}

}  // namespace

void launch_assemble_passes(const float* data, const float* flow, const float* flow2, const float* warp, const float* warp2,
                            float* out, int B, int h, int w, cudaStream_t st) {
    const size_t total = static_cast<size_t>(4) * B * h * w * 29;
    assemble_passes_kernel<<<blocks_for(total, 256), 256, 0, st>>>(data, flow, flow2, warp, warp2, out, B, static_cast<size_t>(h) * w);
}

void launch_groups2ovlp(const float* pred, float* out, int B, int H, int W, cudaStream_t st) {
    const size_t total = static_cast<size_t>(B) * 7 * H * W * 3;
    groups2ovlp_kernel<<<blocks_for(total, 256), 256, 0, st>>>(pred, out, B, static_cast<size_t>(H) * W);
}

size_t temporal_loss_workspace(int B, int h, int w, LossScales* sc) {
    size_t off = 0;
    for (int s = 0; s < 3; ++s) {
        const int hs = h / 2 << s, ws = w / 2 << s;           // pred_l1 is [h/2, w/2], pred_l3 [2h, 2w]
        sc->hw[s] = static_cast<long long>(hs) * ws;
        sc->nblk[s] = static_cast<int>((sc->hw[s] + 255) / 256);
        sc->offset[s] = off;
        off += static_cast<size_t>(B) * sc->nblk[s] * kLossTerms;
    }
    return off * sizeof(double);
}

void launch_temporal_loss(const float* const pred[3], const float* label, int B, int h, int w, const LossLambdas& lam,
                          double* workspace, float* d_out, cudaStream_t st) {
    LossScales sc;
    temporal_loss_workspace(B, h, w, &sc);
    for (int s = 0; s < 3; ++s) {
        const int hs = h / 2 << s, ws = w / 2 << s;
        dim3 grid(sc.nblk[s], B);
        temporal_loss_kernel<<<grid, 256, 0, st>>>(pred[s], label, B, hs, ws, 4 >> s, 2 * h, 2 * w, workspace + sc.offset[s]);
    }
    loss_finalize_kernel<<<1, 1024, 0, st>>>(workspace, sc, B, lam, d_out);
}

void launch_loss_grad(const float* pred, const float* label, ActBuf extra, int B, int hs, int ws, int st, int LH, int LW,
                      float wgt, const LossLambdas& lam, float scale, ActBuf out, cudaStream_t stm) {
    const double n1 = static_cast<double>(B) * hs * ws * 3;
    const double f = 2.0 * scale * wgt;
    LossGradCoef k;
    k.c[0] = static_cast<float>(f * lam.recn / (3 * n1));
    k.c[1] = static_cast<float>(f * lam.tm1 / n1);
    k.c[2] = static_cast<float>(f * lam.tmm / n1);
    k.c[3] = static_cast<float>(f * lam.td / n1);
    k.c[4] = static_cast<float>(f * lam.ss2 * lam.recn / (3 * n1));
    k.c[5] = static_cast<float>(f * lam.ss2 * lam.td / n1);
    k.c[6] = static_cast<float>(f * lam.ss2 * lam.tm2 / (3 * n1));
    dim3 grid(blocks_for(static_cast<size_t>(hs) * ws, 128), B);
    loss_grad_kernel<<<grid, 128, 0, stm>>>(pred, label, extra.p, extra.plane, B, hs, ws, st, LH, LW, k, out.p, out.plane);
}

void launch_pool_bwd(ActBuf skip, ActBuf gcat, int cs, int coff, ActBuf gpool, ActBuf gout, float* rout, int N, int H, int W,
                     int C, cudaStream_t st) {
    const size_t total = static_cast<size_t>(N) * (H / 2) * (W / 2) * (C / 8);
    pool_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(skip.p, skip.plane, gcat.p, gcat.plane, cs, coff, gpool.p, gpool.plane,
                                                          gout.p, gout.plane, rout, N, H, W, C);
}

void launch_upsample_bwd(ActBuf gup, ActBuf xin, ActBuf gout, float* rout, int N, int h, int w, int C, cudaStream_t st) {
    const size_t total = static_cast<size_t>(N) * h * w * (C / 8);
    upsample_bwd_kernel<<<blocks_for(total, 256), 256, 0, st>>>(gup.p, gup.plane, xin.p, xin.plane, gout.p, gout.plane, rout, N, h,
                                                              w, C);
}

void launch_prep_weights_dgrad(const float* w, __half* out, int cin, int cout, int KBo, int cin_pad, cudaStream_t st) {
    const size_t total = static_cast<size_t>(KBo) * 9 * cin_pad * 64;
    prep_weights_dgrad_kernel<<<blocks_for(total, 256), 256, 0, st>>>(w, out, cin, cout, KBo, cin_pad);
}

void launch_grad_absmax(const float* g, size_t n, unsigned* out, cudaStream_t st) {
    unsigned blocks = blocks_for(n, 256);
    if (blocks > 592) blocks = 592;
    grad_absmax_kernel<<<blocks, 256, 0, st>>>(g, n, out);
}

void launch_adam_tf1(float* theta, const float* g, float* m, float* v, size_t n, float lr_t, float beta1, float beta2,
                     float eps, cudaStream_t st) {
    adam_tf1_kernel<<<blocks_for(n, 256), 256, 0, st>>>(theta, g, m, v, n, lr_t, beta1, beta2, eps);
}

void launch_mt_absmax(const MtTensor* d_table, int n, unsigned blocks, unsigned* out, cudaStream_t st) {
    mt_absmax_kernel<<<blocks, 256, 0, st>>>(d_table, n, out);
}
void launch_mt_adam(const MtTensor* d_table, int n, unsigned blocks, float lr_t, float beta1, float beta2, float eps, cudaStream_t st) {
    mt_adam_kernel<<<blocks, 256, 0, st>>>(d_table, n, lr_t, beta1, beta2, eps);
}
void launch_mt_repack(const MtPack* d_table, int n, unsigned blocks, cudaStream_t st) {
    mt_repack_kernel<<<blocks, 256, 0, st>>>(d_table, n);
}

}  // namespace fisr
