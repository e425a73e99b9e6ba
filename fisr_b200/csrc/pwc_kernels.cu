// PWC-Net inference kernels: the optical-flow network the reference runs in front of its flow warp
// (FISR_tfoptflow/model_pwcnet.py:1012-1593, PWC-Net-large, 6-level pyramid, flow predicted at level 2).  Everything that is
// not a stride-1 3x3 conv on the tensor cores (conv_umma_kernel.cuh, launched from pwc_api.cu): CUDA-core kernels on NHWC
// tensors in the split fp16 (hi, lo) plane format of the conv kernels, addressed as (planes, channel offset), so that the
// DenseNet-style concatenations of the flow estimator (model_pwcnet.py:1415-1437: x = concat([act, x])) are channel slices of ONE
// buffer per pyramid level and never copied.  Arithmetic is fp32; the planes hold 22 mantissa bits of each value.  TF semantics restated: 'same' padding puts the odd pad element AFTER the data (stride-2 convs on even sizes pad
// 0 before, 1 after); conv2d_transpose 4x4 s2 'same' is out[2i + k - 1] += in[i] w[k].
#include "common.cuh"
#include "pwc_kernels.h"

namespace fisr {
namespace pwc {

namespace {

__device__ __forceinline__ float lrelu(float x) { return x > 0.f ? x : 0.1f * x; }      // tf.nn.leaky_relu(alpha=0.1)
__device__ __forceinline__ unsigned long long pack2(float lo, float hi) {
    return static_cast<unsigned long long>(__float_as_uint(lo)) | (static_cast<unsigned long long>(__float_as_uint(hi)) << 32);
}

// ---- split-plane accessors: element i of a buffer is hi[i] + lo[i] with lo = hi + plane
__device__ __forceinline__ float ld1(const __half* p, size_t plane, size_t i) { return __half2float(__ldg(p + i)) + __half2float(__ldg(p + plane + i)); }
__device__ __forceinline__ float2 join2(uint32_t h, uint32_t l) {
    const float2 a = __half22float2(*reinterpret_cast<const __half2*>(&h)), b = __half22float2(*reinterpret_cast<const __half2*>(&l));
    return make_float2(a.x + b.x, a.y + b.y);
}
__device__ __forceinline__ float4 ld4(const __half* p, size_t plane, size_t i) {          // i % 4 == 0
    const uint2 h = __ldg(reinterpret_cast<const uint2*>(p + i)), l = __ldg(reinterpret_cast<const uint2*>(p + plane + i));
    const float2 a = join2(h.x, l.x), b = join2(h.y, l.y);
    return make_float4(a.x, a.y, b.x, b.y);
}
__device__ __forceinline__ void ld8(const __half* p, size_t plane, size_t i, float* v) {   // i % 8 == 0
    const uint4 h = __ldg(reinterpret_cast<const uint4*>(p + i)), l = __ldg(reinterpret_cast<const uint4*>(p + plane + i));
    const float2 a = join2(h.x, l.x), b = join2(h.y, l.y), c = join2(h.z, l.z), d = join2(h.w, l.w);
    v[0] = a.x; v[1] = a.y; v[2] = b.x; v[3] = b.y; v[4] = c.x; v[5] = c.y; v[6] = d.x; v[7] = d.y;
}
__device__ __forceinline__ void st1(__half* p, size_t plane, size_t i, float v) {
    const SplitHalf s = split_f32(v);
    p[i] = s.hi;
    p[plane + i] = s.lo;
}
__device__ __forceinline__ void st2(__half* p, size_t plane, size_t i, float v0, float v1) {          // i % 2 == 0
    uint32_t h, l;
    split2_f32(v0, v1, h, l);
    *reinterpret_cast<uint32_t*>(p + i) = h;
    *reinterpret_cast<uint32_t*>(p + plane + i) = l;
}
__device__ __forceinline__ void st4(__half* p, size_t plane, size_t i, float4 v) {          // i % 4 == 0
    uint32_t h01, l01, h23, l23;
    split2_f32(v.x, v.y, h01, l01);
    split2_f32(v.z, v.w, h23, l23);
    *reinterpret_cast<uint2*>(p + i) = make_uint2(h01, h23);
    *reinterpret_cast<uint2*>(p + plane + i) = make_uint2(l01, l23);
}

// ---------------------------------------------------------------- 3x3 conv, wide outputs (Cout a multiple of 4)
// Implicit GEMM on CUDA cores: a block owns 128 output pixels x 64 output channels, a thread 8 x 8 of them; per (tap, 16 input
// channels) the block stages the 128 x 16 activation slice (gathered at the tap's stride / dilation offset, zero outside) and the
// 16 x 64 weight slice in shared memory and every thread does 64 FMAs per 4 shared-memory vector loads.
constexpr int kPx = 128, kCo = 64, kCi = 16;

template <bool LEAKY>
__global__ void __launch_bounds__(128) conv3x3_wide_kernel(const PwcConv p) {
    __shared__ __align__(16) float As[kCi][kPx];
    __shared__ __align__(16) float Bs[kCi][kCo];
    const int t = threadIdx.x;
    const int pg = t & 15, cg = t >> 4;
    const long long npix = static_cast<long long>(p.N) * p.Hout * p.Wout;
    const long long P = static_cast<long long>(blockIdx.x) * kPx + t;          // the pixel this thread gathers
    const int co_base = blockIdx.y * kCo;
    int n = 0, oy = 0, ox = 0;
    const bool live = P < npix;
    if (live) {
        ox = static_cast<int>(P % p.Wout);
        const long long r = P / p.Wout;
        oy = static_cast<int>(r % p.Hout);
        n = static_cast<int>(r / p.Hout);
    }
    const bool vec_in = ((p.in.cs | p.in_coff) & 7) == 0;
    // accumulators as packed fp32 pairs: fma.rn.f32x2 (FFMA2, sm_100) does two FMAs per issued instruction, which matters in a
    // kernel that is issue bound (512 scalar FFMAs per 8-channel slice next to ~100 other instructions)
    unsigned long long acc2[8][4];
#pragma unroll
    for (int i = 0; i < 8; ++i)
#pragma unroll
        for (int j = 0; j < 4; ++j) acc2[i][j] = 0ull;
    const int bci = t >> 4, bco = (t & 15) * 4;                                 // this thread's weight fetch: row bci, 4 columns at bco
    const int nci = (p.cin + kCi - 1) / kCi, steps = 9 * nci;                   // (tap, 8-channel slice) items
    // software pipeline: the global loads of item s + 1 are in flight while item s is multiplied out of shared memory
    float a[kCi];
    float4 wv[kCi / 8];
    auto fetch = [&](int s) {
        const int tap = s / nci, c0 = (s - tap * nci) * kCi;
        const int iy = oy * p.stride + (tap / 3) * p.dil - p.pad_y, ix = ox * p.stride + (tap % 3) * p.dil - p.pad_x;
        const bool inside = live && iy >= 0 && iy < p.Hin && ix >= 0 && ix < p.Win;
#pragma unroll
        for (int j = 0; j < kCi; ++j) a[j] = 0.f;
        if (inside) {
            const size_t src = (static_cast<size_t>(n) * p.Hin + iy) * p.Win * p.in.cs + static_cast<size_t>(ix) * p.in.cs + p.in_coff + c0;
            if (vec_in && c0 + kCi <= p.cin) {
#pragma unroll
                for (int q = 0; q < kCi / 8; ++q) ld8(p.in.p, p.in.plane, src + 8 * q, a + 8 * q);
            } else {
#pragma unroll
                for (int j = 0; j < kCi; ++j) if (c0 + j < p.cin) a[j] = ld1(p.in.p, p.in.plane, src + j);
            }
        }
#pragma unroll
        for (int q = 0; q < kCi / 8; ++q) {
            wv[q] = make_float4(0.f, 0.f, 0.f, 0.f);
            if (c0 + bci + 8 * q < p.cin && co_base + bco < p.cout)
                wv[q] = __ldg(reinterpret_cast<const float4*>(p.w + (static_cast<size_t>(tap) * p.cin + c0 + bci + 8 * q) * p.cout + co_base + bco));
        }
    };
    fetch(0);
    for (int s = 0; s < steps; ++s) {
        __syncthreads();                                                        // previous slice fully consumed
#pragma unroll
        for (int j = 0; j < kCi; ++j) As[j][t] = a[j];
#pragma unroll
        for (int q = 0; q < kCi / 8; ++q) *reinterpret_cast<float4*>(&Bs[bci + 8 * q][bco]) = wv[q];
        __syncthreads();
        if (s + 1 < steps) fetch(s + 1);
#pragma unroll
        for (int c = 0; c < kCi; ++c) {
            const float4 a0 = *reinterpret_cast<const float4*>(&As[c][pg * 8]), a1 = *reinterpret_cast<const float4*>(&As[c][pg * 8 + 4]);
            const float4 b0 = *reinterpret_cast<const float4*>(&Bs[c][cg * 8]), b1 = *reinterpret_cast<const float4*>(&Bs[c][cg * 8 + 4]);
            const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
            const unsigned long long bp[4] = {pack2(b0.x, b0.y), pack2(b0.z, b0.w), pack2(b1.x, b1.y), pack2(b1.z, b1.w)};
#pragma unroll
            for (int i = 0; i < 8; ++i) {
                const unsigned long long ap = pack2(av[i], av[i]);
#pragma unroll
                for (int j = 0; j < 4; ++j) asm("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i][j]) : "l"(ap), "l"(bp[j]));
            }
        }
    }
    const int co0 = co_base + cg * 8;
    if (co0 >= p.cout) return;
    float bias[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) bias[j] = co0 + j < p.cout ? __ldg(p.b + co0 + j) : 0.f;
    const bool vec_out = ((p.out.cs | p.out_coff) & 3) == 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
        const long long q = static_cast<long long>(blockIdx.x) * kPx + pg * 8 + i;
        if (q >= npix) break;
        const size_t d = static_cast<size_t>(q) * p.out.cs + p.out_coff + co0;
        float v[8];
#pragma unroll
        for (int j = 0; j < 8; ++j) {
            v[j] = __uint_as_float(static_cast<unsigned>(acc2[i][j >> 1] >> ((j & 1) * 32))) + bias[j];
            if (LEAKY) v[j] = lrelu(v[j]);
        }
        if (vec_out && co0 + 8 <= p.cout) {
            st4(p.out.p, p.out.plane, d, make_float4(v[0], v[1], v[2], v[3]));
            st4(p.out.p, p.out.plane, d + 4, make_float4(v[4], v[5], v[6], v[7]));
        } else {
#pragma unroll
            for (int j = 0; j < 8; ++j) if (co0 + j < p.cout) st1(p.out.p, p.out.plane, d + j, v[j]);
        }
    }
}

// ---------------------------------------------------------------- 3x3 conv, narrow outputs (Cout <= 4: the flow predictors)
// A thread owns one output pixel; the 9 x Cin x Cout weights sit in shared memory (broadcast reads).  Optional residual input
// (refine_flow: flow + dc_conv7, model_pwcnet.py:1522).
template <int COUT>
__global__ void __launch_bounds__(128) conv3x3_narrow_kernel(const PwcConv p) {
    extern __shared__ float wsm[];                              // [9][cin][COUT]
    for (int i = threadIdx.x; i < 9 * p.cin * COUT; i += blockDim.x) wsm[i] = __ldg(p.w + i);
    __syncthreads();
    const long long npix = static_cast<long long>(p.N) * p.Hout * p.Wout;
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= npix) return;
    const int ox = static_cast<int>(P % p.Wout);
    const long long r = P / p.Wout;
    const int oy = static_cast<int>(r % p.Hout), n = static_cast<int>(r / p.Hout);
    const bool vec_in = ((p.in.cs | p.in_coff) & 3) == 0;
    float acc[COUT];
#pragma unroll
    for (int j = 0; j < COUT; ++j) acc[j] = __ldg(p.b + j);
    for (int tap = 0; tap < 9; ++tap) {
        const int iy = oy * p.stride + (tap / 3) * p.dil - p.pad_y, ix = ox * p.stride + (tap % 3) * p.dil - p.pad_x;
        if (iy < 0 || iy >= p.Hin || ix < 0 || ix >= p.Win) continue;
        const size_t src = (static_cast<size_t>(n) * p.Hin + iy) * p.Win * p.in.cs + static_cast<size_t>(ix) * p.in.cs + p.in_coff;
        const float* wt = wsm + tap * p.cin * COUT;
        int c = 0;
        if (vec_in)
            for (; c + 4 <= p.cin; c += 4) {
                const float4 v = ld4(p.in.p, p.in.plane, src + c);
#pragma unroll
                for (int j = 0; j < COUT; ++j)
                    acc[j] = fmaf(v.w, wt[(c + 3) * COUT + j], fmaf(v.z, wt[(c + 2) * COUT + j], fmaf(v.y, wt[(c + 1) * COUT + j], fmaf(v.x, wt[c * COUT + j], acc[j]))));
            }
        for (; c < p.cin; ++c) {
            const float v = ld1(p.in.p, p.in.plane, src + c);
#pragma unroll
            for (int j = 0; j < COUT; ++j) acc[j] = fmaf(v, wt[c * COUT + j], acc[j]);
        }
    }
    float* d = p.out_f32 + static_cast<size_t>(P) * COUT;
#pragma unroll
    for (int j = 0; j < COUT; ++j) {
        float v = acc[j];
        if (p.add) v += __ldg(p.add + static_cast<size_t>(P) * p.add_cs + j);
        d[j] = p.leaky ? lrelu(v) : v;
    }
}

// ---------------------------------------------------------------- conv2d_transpose 4x4, stride 2, 'same', 2 filters (model_pwcnet.py:1180-1224)
// out[n, oy, ox, co] = b[co] + sum over (ky, kx, ci) with oy = 2 iy + ky - 1, ox = 2 ix + kx - 1 of in[n, iy, ix, ci] w[ky, kx, co, ci]
// The input is either a plane buffer or a plain fp32 [.., 2] flow (up_flow reads the fp32 flow of the level above).
template <bool IN_F32>
__global__ void __launch_bounds__(128) deconv4x4s2_kernel(Planes in, int in_coff, const float* __restrict__ in_f32, int in_f32_cs, int cin,
                                                          const float* __restrict__ w, const float* __restrict__ b, Planes out, int out_coff, int N, int h, int wd) {
    const long long npix = static_cast<long long>(N) * 4 * h * wd;
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= npix) return;
    const int ox = static_cast<int>(P % (2 * wd));
    const long long r = P / (2 * wd);
    const int oy = static_cast<int>(r % (2 * h)), n = static_cast<int>(r / (2 * h));
    float a0 = __ldg(b), a1 = __ldg(b + 1);
    const int ics = IN_F32 ? in_f32_cs : in.cs;
    const bool vec = !IN_F32 && ((in.cs | in_coff | cin) & 3) == 0;
#pragma unroll
    for (int sy = 0; sy < 2; ++sy) {
        const int ky = ((oy + 1) & 1) + 2 * sy, iy = (oy + 1 - ky) >> 1;
        if (iy < 0 || iy >= h) continue;
#pragma unroll
        for (int sx = 0; sx < 2; ++sx) {
            const int kx = ((ox + 1) & 1) + 2 * sx, ix = (ox + 1 - kx) >> 1;
            if (ix < 0 || ix >= wd) continue;
            const size_t src = (static_cast<size_t>(n) * h + iy) * wd * ics + static_cast<size_t>(ix) * ics + (IN_F32 ? 0 : in_coff);
            const float* w0 = w + static_cast<size_t>((ky * 4 + kx) * 2) * cin;
            const float* w1 = w0 + cin;
            int c = 0;
            if (vec)
                for (; c + 4 <= cin; c += 4) {
                    const float4 v = ld4(in.p, in.plane, src + c);
                    const float4 u0 = __ldg(reinterpret_cast<const float4*>(w0 + c)), u1 = __ldg(reinterpret_cast<const float4*>(w1 + c));
                    a0 = fmaf(v.w, u0.w, fmaf(v.z, u0.z, fmaf(v.y, u0.y, fmaf(v.x, u0.x, a0))));
                    a1 = fmaf(v.w, u1.w, fmaf(v.z, u1.z, fmaf(v.y, u1.y, fmaf(v.x, u1.x, a1))));
                }
            for (; c < cin; ++c) {
                const float v = IN_F32 ? __ldg(in_f32 + src + c) : ld1(in.p, in.plane, src + c);
                a0 = fmaf(v, __ldg(w0 + c), a0);
                a1 = fmaf(v, __ldg(w1 + c), a1);
            }
        }
    }
    st2(out.p, out.plane, static_cast<size_t>(P) * out.cs + out_coff, a0, a1);
}

// ---------------------------------------------------------------- cost volume, search range 4 (model_pwcnet.py:1226-1277)
// out[p, (dy+4)*9 + (dx+4)] = leaky_relu(mean_c c1[p, c] * c2[p + (dy, dx), c]), zero outside the image
// A block owns 4 rows x 32 columns of pixels; per 16-channel slice it stages its c1 tile and the (4 + 8) x (32 + 8) c2 region in
// shared memory, channel-major.  A thread owns 4 consecutive pixels of a row and ONE row dy of the 9 x 9 window: per channel it
// reads its 4 c1 values and the 12 c2 values they see (4 x 128-bit shared loads) for 36 FMAs; a warp = the 32 pixel groups of the
// tile for one dy, so a quarter-warp always reads 128 contiguous bytes.
constexpr int kCvTy = 4, kCvTx = 32, kCvC = 16, kCvRy = kCvTy + 8, kCvRx = kCvTx + 8, kCvThreads = 9 * 32;

__global__ void __launch_bounds__(kCvThreads) cost_volume_kernel(Planes c1, int c1_coff, Planes c2, int c2_coff, int C, Planes out, int out_coff, int N, int h, int w) {
    __shared__ __align__(16) float s1[kCvC][kCvTy * kCvTx];
    __shared__ __align__(16) float s2[kCvC][kCvRy * kCvRx];
    const int t = threadIdx.x;
    const int x0 = blockIdx.x * kCvTx, y0 = blockIdx.y * kCvTy, n = blockIdx.z;
    const int g = t & 31, dyi = t >> 5;                // pixel group, row of the search window (dy = dyi - 4)
    const int gy = g >> 3, gx = (g & 7) * 4;
    float acc[4][9];
#pragma unroll
    for (int i = 0; i < 4; ++i)
#pragma unroll
        for (int j = 0; j < 9; ++j) acc[i][j] = 0.f;
    for (int cb = 0; cb < C; cb += kCvC) {
        __syncthreads();
        // (pixel, 8-channel group) units: consecutive threads take consecutive pixels of one group
        for (int u = t; u < 2 * kCvTy * kCvTx; u += kCvThreads) {
            const int p = u % (kCvTy * kCvTx), grp = u / (kCvTy * kCvTx);
            const int y = y0 + (p >> 5), x = x0 + (p & 31);
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (y < h && x < w) {
                const size_t base = (static_cast<size_t>(n) * h + y) * w * c1.cs + static_cast<size_t>(x) * c1.cs + c1_coff + cb + 8 * grp;
                if (cb + 8 * grp + 8 <= C) { const float4 a = ld4(c1.p, c1.plane, base), b = ld4(c1.p, c1.plane, base + 4); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
                else if (cb + 8 * grp + 4 <= C) { const float4 a = ld4(c1.p, c1.plane, base); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s1[8 * grp + j][p] = v[j];
        }
        for (int u = t; u < 2 * kCvRy * kCvRx; u += kCvThreads) {
            const int p = u % (kCvRy * kCvRx), grp = u / (kCvRy * kCvRx);
            const int y = y0 - 4 + p / kCvRx, x = x0 - 4 + p % kCvRx;
            float v[8] = {0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f, 0.f};
            if (y >= 0 && y < h && x >= 0 && x < w) {
                const size_t base = (static_cast<size_t>(n) * h + y) * w * c2.cs + static_cast<size_t>(x) * c2.cs + c2_coff + cb + 8 * grp;
                if (cb + 8 * grp + 8 <= C) { const float4 a = ld4(c2.p, c2.plane, base), b = ld4(c2.p, c2.plane, base + 4); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; v[4] = b.x; v[5] = b.y; v[6] = b.z; v[7] = b.w; }
                else if (cb + 8 * grp + 4 <= C) { const float4 a = ld4(c2.p, c2.plane, base); v[0] = a.x; v[1] = a.y; v[2] = a.z; v[3] = a.w; }
            }
#pragma unroll
            for (int j = 0; j < 8; ++j) s2[8 * grp + j][p] = v[j];
        }
        __syncthreads();
#pragma unroll 4
        for (int c = 0; c < kCvC; ++c) {
            const float4 a4 = *reinterpret_cast<const float4*>(&s1[c][gy * kCvTx + gx]);
            const float* row = &s2[c][(gy + dyi) * kCvRx + gx];
            const float4 b0 = *reinterpret_cast<const float4*>(row), b1 = *reinterpret_cast<const float4*>(row + 4), b2 = *reinterpret_cast<const float4*>(row + 8);
            const float a[4] = {a4.x, a4.y, a4.z, a4.w};
            const float b[12] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w, b2.x, b2.y, b2.z, b2.w};
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int dx = 0; dx < 9; ++dx) acc[i][dx] = fmaf(a[i], b[i + dx], acc[i][dx]);
        }
    }
    // Output through shared memory: a pixel's 81 values are contiguous in the dense buffer, but a thread holds 9 of them for 4
    // pixels -- stored directly, every warp store touches 32 different 32-byte sectors with 4 bytes each (measured: the kernel ran
    // at the L2's partial-sector write rate, 3 GB of sector traffic for 340 MB).  Two rows of the tile at a time are transposed
    // through the c2 staging buffer and written by whole warps, 64 + 17 channels of one pixel per pass.
    const float inv = 1.f / static_cast<float>(C);
    float* so = &s2[0][0];                                     // [64 pixels][81]
    const int warp = t >> 5, lane = t & 31;
    for (int half = 0; half < 2; ++half) {
        __syncthreads();                                       // c2 tile (or the previous half) fully consumed
        if ((gy >> 1) == half) {
#pragma unroll
            for (int i = 0; i < 4; ++i)
#pragma unroll
                for (int dx = 0; dx < 9; ++dx) so[((gy & 1) * kCvTx + gx + i) * 81 + dyi * 9 + dx] = lrelu(acc[i][dx] * inv);
        }
        __syncthreads();
        for (int p = warp; p < 2 * kCvTx; p += kCvThreads / 32) {
            const int y = y0 + 2 * half + (p >> 5), x = x0 + (p & 31);
            if (y >= h || x >= w) continue;
            const size_t o = ((static_cast<size_t>(n) * h + y) * w + x) * out.cs + out_coff;       // even: channel pairs are 4-byte aligned
            const float* v = so + p * 81;
            st2(out.p, out.plane, o + 2 * lane, v[2 * lane], v[2 * lane + 1]);
            if (lane < 8) st2(out.p, out.plane, o + 64 + 2 * lane, v[64 + 2 * lane], v[65 + 2 * lane]);
            else if (lane == 8) st1(out.p, out.plane, o + 80, v[80]);
        }
    }
}

// ---------------------------------------------------------------- dense_image_warp (model_pwcnet.py:1106-1178)
// out[p, c] = bilinear(img, (x + s u, y + s v)) with TF's _interpolate_bilinear clamping: floor in [0, size - 2], weight in [0, 1]
__global__ void __launch_bounds__(256) dense_warp_kernel(Planes img, int coff, int C, Planes flow, int f_coff, float scale, Planes out, int N, int h, int w) {
    const int cv = C / 4;
    const long long total = static_cast<long long>(N) * h * w * cv;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int c4 = static_cast<int>(i % cv) * 4;
    const long long P = i / cv;
    const int x = static_cast<int>(P % w);
    const long long r = P / w;
    const int y = static_cast<int>(r % h), n = static_cast<int>(r / h);
    const size_t fi = static_cast<size_t>(P) * flow.cs + f_coff;
    const float qx = static_cast<float>(x) + scale * ld1(flow.p, flow.plane, fi);
    const float qy = static_cast<float>(y) + scale * ld1(flow.p, flow.plane, fi + 1);
    const float fx0 = fminf(fmaxf(floorf(qx), 0.f), static_cast<float>(w - 2)), fy0 = fminf(fmaxf(floorf(qy), 0.f), static_cast<float>(h - 2));
    const float ax = fminf(fmaxf(qx - fx0, 0.f), 1.f), ay = fminf(fmaxf(qy - fy0, 0.f), 1.f);
    const int x0 = static_cast<int>(fx0), y0 = static_cast<int>(fy0);
    const size_t base = (static_cast<size_t>(n) * h + y0) * w * img.cs + static_cast<size_t>(x0) * img.cs + coff + c4;
    const float4 tl = ld4(img.p, img.plane, base), tr = ld4(img.p, img.plane, base + img.cs);
    const float4 bl = ld4(img.p, img.plane, base + static_cast<size_t>(w) * img.cs), br = ld4(img.p, img.plane, base + static_cast<size_t>(w) * img.cs + img.cs);
    auto mix = [&](float a, float b, float c, float d) {
        const float top = a + ax * (b - a), bot = c + ax * (d - c);
        return top + ay * (bot - top);
    };
    st4(out.p, out.plane, static_cast<size_t>(P) * out.cs + c4,
        make_float4(mix(tl.x, tr.x, bl.x, br.x), mix(tl.y, tr.y, bl.y, br.y), mix(tl.z, tr.z, bl.z, br.z), mix(tl.w, tr.w, bl.w, br.w)));
}

// ---------------------------------------------------------------- conv1a: 3 -> 16, stride 2, leaky ReLU, straight from the fp32 image
// (model_pwcnet.py:1083-1096).  'same' on an even size with stride 2 pads only after the data: taps (2oy + ky, 2ox + kx), ky, kx in 0..2.
__global__ void __launch_bounds__(128) first_conv_kernel(const float* __restrict__ img, const float* __restrict__ w, const float* __restrict__ b, Planes out,
                                                         int N, int H, int W) {
    __shared__ float ws[27 * 16 + 16];
    for (int i = threadIdx.x; i < 27 * 16 + 16; i += blockDim.x) ws[i] = i < 27 * 16 ? __ldg(w + i) : __ldg(b + i - 27 * 16);
    __syncthreads();
    const int ho = H / 2, wo = W / 2;
    const long long npix = static_cast<long long>(N) * ho * wo;
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= npix) return;
    const int ox = static_cast<int>(P % wo);
    const long long r = P / wo;
    const int oy = static_cast<int>(r % ho), n = static_cast<int>(r / ho);
    float acc[16];
#pragma unroll
    for (int j = 0; j < 16; ++j) acc[j] = ws[27 * 16 + j];
#pragma unroll
    for (int ky = 0; ky < 3; ++ky) {
        const int iy = 2 * oy + ky;
        if (iy >= H) continue;
        const float* row = img + ((static_cast<size_t>(n) * H + iy) * W + 2 * ox) * 3;
#pragma unroll
        for (int kx = 0; kx < 3; ++kx) {
            if (2 * ox + kx >= W) continue;
#pragma unroll
            for (int ci = 0; ci < 3; ++ci) {
                const float v = __ldg(row + kx * 3 + ci);
                const float* wt = ws + ((ky * 3 + kx) * 3 + ci) * 16;
#pragma unroll
                for (int j = 0; j < 16; ++j) acc[j] = fmaf(v, wt[j], acc[j]);
            }
        }
    }
    const size_t o = static_cast<size_t>(P) * out.cs;
#pragma unroll
    for (int j = 0; j < 16; j += 4) st4(out.p, out.plane, o + j, make_float4(lrelu(acc[j]), lrelu(acc[j + 1]), lrelu(acc[j + 2]), lrelu(acc[j + 3])));
}

// ---------------------------------------------------------------- outputs of the fused flow-predictor / up_feat conv (pwc_api.cu)
// F [N,h,w,16]: columns 0..1 = predict_flow (fp32 out), columns 2 + (2a + b) * 2 + co = up_feat output pixel (2y + a, 2x + b), channel co
__global__ void __launch_bounds__(256) flow_upfeat_scatter_kernel(Planes F, float* __restrict__ flow, Planes Dn, int up_off, int N, int h, int w) {
    const long long npix = static_cast<long long>(N) * h * w;
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= npix) return;
    float v[16];
    ld8(F.p, F.plane, static_cast<size_t>(P) * F.cs, v);
    ld8(F.p, F.plane, static_cast<size_t>(P) * F.cs + 8, v + 8);
    reinterpret_cast<float2*>(flow)[P] = make_float2(v[0], v[1]);
    if (!Dn.p) return;
    const int x = static_cast<int>(P % w);
    const long long r = P / w;
    const int y = static_cast<int>(r % h), n = static_cast<int>(r / h);
#pragma unroll
    for (int s = 0; s < 4; ++s) {
        const size_t op = (static_cast<size_t>(n) * 2 * h + 2 * y + (s >> 1)) * (2 * w) + 2 * x + (s & 1);
        st2(Dn.p, Dn.plane, op * Dn.cs + up_off, v[2 + 2 * s], v[3 + 2 * s]);
    }
}

// ---------------------------------------------------------------- the reference driver around the network, on the device
// FISR_for_video_pwcnet_predict_from_img_test.py:113-131 + adapt_x (model_pwcnet.py:371-409): YUV -> RGB (float64, clipped to
// 0..255), x`scale` skimage.transform.resize (order 1, mode 'reflect', pixel-centre coordinates, float64; rows first, then
// columns), truncation to uint8, /255 in float32, zero pad to multiples of 64.  Every float64 step is a separately rounded
// multiply / add in the order numpy evaluates it, so the uint8 truncation sees the same value as the host pipeline.
__device__ __forceinline__ int mirror_index(long long i, int n) {          // scipy 'mirror' / skimage 'reflect': -1 -> 1, n -> n - 2
    if (n == 1) return 0;
    const long long period = 2LL * (n - 1);
    long long m = i % period;
    if (m < 0) m += period;
    return static_cast<int>(m >= n ? period - m : m);
}
struct LerpTap { int i0, i1; double f0, f1; };                               // value = a[i0] * f0 + a[i1] * f1
__device__ __forceinline__ LerpTap lerp_tap(int o, double ratio, int n) {    // coords = (o + 0.5) * (n_in / n_out) - 0.5
    const double c = __dadd_rn(__dmul_rn(static_cast<double>(o) + 0.5, ratio), -0.5);
    const double fl = floor(c);
    const double fr = __dadd_rn(c, -fl);
    LerpTap t;
    t.i0 = mirror_index(static_cast<long long>(fl), n);
    t.i1 = mirror_index(static_cast<long long>(fl) + 1, n);
    t.f0 = __dadd_rn(1.0, -fr);
    t.f1 = fr;
    return t;
}

template <bool YUV>
__device__ __forceinline__ void load_rgb(const void* src, int w, int y, int x, const PrepConst& k, double (&rgb)[3]) {
    if constexpr (YUV) {
        const uint8_t* p = static_cast<const uint8_t*>(src) + (static_cast<size_t>(y) * w + x) * 3;
        const double Y = static_cast<double>(__ldg(p)), U = static_cast<double>(__ldg(p + 1)), V = static_cast<double>(__ldg(p + 2));
#pragma unroll
        for (int c = 0; c < 3; ++c) {        // utils.py:106-115: T[p,0] * y + T[p,1] * u + T[p,2] * v - offset[p], then clip
            double v = __dadd_rn(__dadd_rn(__dmul_rn(k.T[3 * c], Y), __dmul_rn(k.T[3 * c + 1], U)), __dmul_rn(k.T[3 * c + 2], V));
            v = __dadd_rn(v, -k.off[c]);
            rgb[c] = fmin(fmax(v, 0.0), 255.0);
        }
    } else {
        const double* p = static_cast<const double*>(src) + (static_cast<size_t>(y) * w + x) * 3;
        rgb[0] = __ldg(p); rgb[1] = __ldg(p + 1); rgb[2] = __ldg(p + 2);
    }
}

// grid.y = which frame of the pair; frame 0 lands in img1[0] and img2[1], frame 1 in img1[1] and img2[0] (both directions as one batch)
template <bool YUV>
__global__ void __launch_bounds__(256) prepare_pair_kernel(const void* __restrict__ src0, const void* __restrict__ src1, float* __restrict__ img1,
                                                           float* __restrict__ img2, int h, int w, int oh, int ow, int Hp, int Wp, PrepConst k) {
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= static_cast<long long>(Hp) * Wp) return;
    const int X = static_cast<int>(P % Wp), Y = static_cast<int>(P / Wp);
    const int which = blockIdx.y;
    float o[3] = {0.f, 0.f, 0.f};
    if (Y < oh && X < ow) {
        const void* src = which ? src1 : src0;
        const LerpTap ty = lerp_tap(Y, k.ry, h), tx = lerp_tap(X, k.rx, w);
        double p00[3], p10[3], p01[3], p11[3];
        load_rgb<YUV>(src, w, ty.i0, tx.i0, k, p00);
        load_rgb<YUV>(src, w, ty.i1, tx.i0, k, p10);
        load_rgb<YUV>(src, w, ty.i0, tx.i1, k, p01);
        load_rgb<YUV>(src, w, ty.i1, tx.i1, k, p11);
#pragma unroll
        for (int c = 0; c < 3; ++c) {
            const double r0 = __dadd_rn(__dmul_rn(p00[c], ty.f0), __dmul_rn(p10[c], ty.f1));
            const double r1 = __dadd_rn(__dmul_rn(p01[c], ty.f0), __dmul_rn(p11[c], ty.f1));
            const double v = __dadd_rn(__dmul_rn(r0, tx.f0), __dmul_rn(r1, tx.f1));
            o[c] = __fdiv_rn(static_cast<float>(static_cast<int>(v) & 255), 255.f);      // np.array(.., dtype=np.uint8), then / np.float32(255)
        }
    }
    const size_t img = static_cast<size_t>(Hp) * Wp * 3, at = static_cast<size_t>(P) * 3;
    float* d1 = img1 + (which ? img : 0) + at;
    float* d2 = img2 + (which ? 0 : img) + at;
    d1[0] = o[0]; d1[1] = o[1]; d1[2] = o[2];
    d2[0] = o[0]; d2[1] = o[1]; d2[2] = o[2];
}

// postproc_y_hat_test crop (model_pwcnet.py:449-470) + ..predict_from_img_test.py:137: skimage resize of the cropped flow with
// anti_aliasing (scipy.ndimage.gaussian_filter, mode 'mirror': rows then columns, symmetric-kernel summation order of
// NI_Correlate1D), order-1 interpolation (rows, then columns), / scale, float32.
struct FlowSrc { const float2* f; int Wp, h0, w0; };
__device__ __forceinline__ void gauss_rows(const FlowSrc& s, int y, int x, const FinishConst& k, double (&g)[2]) {      // filter along y at column x
    const float2 c = __ldg(s.f + static_cast<size_t>(y) * s.Wp + x);
    g[0] = __dmul_rn(static_cast<double>(c.x), k.wy[0]);
    g[1] = __dmul_rn(static_cast<double>(c.y), k.wy[0]);
    for (int j = k.ry; j >= 1; --j) {
        const float2 a = __ldg(s.f + static_cast<size_t>(mirror_index(y - j, s.h0)) * s.Wp + x);
        const float2 b = __ldg(s.f + static_cast<size_t>(mirror_index(y + j, s.h0)) * s.Wp + x);
        g[0] = __dadd_rn(g[0], __dmul_rn(__dadd_rn(static_cast<double>(a.x), static_cast<double>(b.x)), k.wy[j]));
        g[1] = __dadd_rn(g[1], __dmul_rn(__dadd_rn(static_cast<double>(a.y), static_cast<double>(b.y)), k.wy[j]));
    }
}
__device__ __forceinline__ void gauss_both(const FlowSrc& s, int y, int x, const FinishConst& k, double (&g)[2]) {      // rows, then columns
    double c[2];
    gauss_rows(s, y, x, k, c);
    g[0] = __dmul_rn(c[0], k.wx[0]);
    g[1] = __dmul_rn(c[1], k.wx[0]);
    for (int j = k.rx; j >= 1; --j) {
        double a[2], b[2];
        gauss_rows(s, y, mirror_index(x - j, s.w0), k, a);
        gauss_rows(s, y, mirror_index(x + j, s.w0), k, b);
        g[0] = __dadd_rn(g[0], __dmul_rn(__dadd_rn(a[0], b[0]), k.wx[j]));
        g[1] = __dadd_rn(g[1], __dmul_rn(__dadd_rn(a[1], b[1]), k.wx[j]));
    }
}
__global__ void __launch_bounds__(128) finish_flow_kernel(const float* __restrict__ flow, int N, int Hp, int Wp, int h0, int w0, int oh, int ow,
                                                          float* __restrict__ out, FinishConst k) {
    const long long P = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (P >= static_cast<long long>(N) * oh * ow) return;
    const int X = static_cast<int>(P % ow);
    const long long r = P / ow;
    const int Y = static_cast<int>(r % oh), n = static_cast<int>(r / oh);
    const FlowSrc s{reinterpret_cast<const float2*>(flow) + static_cast<size_t>(n) * Hp * Wp, Wp, h0, w0};
    const LerpTap ty = lerp_tap(Y, k.ratio_y, h0), tx = lerp_tap(X, k.ratio_x, w0);
    double g00[2], g10[2], g01[2], g11[2];
    gauss_both(s, ty.i0, tx.i0, k, g00);
    gauss_both(s, ty.i1, tx.i0, k, g10);
    gauss_both(s, ty.i0, tx.i1, k, g01);
    gauss_both(s, ty.i1, tx.i1, k, g11);
    float o[2];
#pragma unroll
    for (int c = 0; c < 2; ++c) {
        const double r0 = __dadd_rn(__dmul_rn(g00[c], ty.f0), __dmul_rn(g10[c], ty.f1));
        const double r1 = __dadd_rn(__dmul_rn(g01[c], ty.f0), __dmul_rn(g11[c], ty.f1));
        o[c] = __double2float_rn(__ddiv_rn(__dadd_rn(__dmul_rn(r0, tx.f0), __dmul_rn(r1, tx.f1)), k.scale));
    }
    reinterpret_cast<float2*>(out)[P] = make_float2(o[0], o[1]);
}

// ---------------------------------------------------------------- channel-slice copy between plane buffers (C % 8 == 0, 16-byte rows)
__global__ void __launch_bounds__(256) copy_channels_kernel(Planes src, int src_off, Planes dst, int dst_off, int C, long long npix) {
    const int g = C / 8;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= npix * g) return;
    const long long P = i / g;
    const int c = static_cast<int>(i % g) * 8;
    const size_t a = static_cast<size_t>(P) * src.cs + src_off + c, b = static_cast<size_t>(P) * dst.cs + dst_off + c;
    *reinterpret_cast<uint4*>(dst.p + b) = __ldg(reinterpret_cast<const uint4*>(src.p + a));
    *reinterpret_cast<uint4*>(dst.p + dst.plane + b) = __ldg(reinterpret_cast<const uint4*>(src.p + src.plane + a));
}

// ---------------------------------------------------------------- tf.image.resize_bilinear x S (legacy: src = dst / S), times `gain`
__global__ void __launch_bounds__(256) resize_flow_kernel(const float* __restrict__ in, float* __restrict__ out, int N, int h, int w, int S, float gain) {
    const long long total = static_cast<long long>(N) * h * S * w * S;
    const long long i = static_cast<long long>(blockIdx.x) * blockDim.x + threadIdx.x;
    if (i >= total) return;
    const int ox = static_cast<int>(i % (w * S));
    const long long r = i / (w * S);
    const int oy = static_cast<int>(r % (h * S)), n = static_cast<int>(r / (h * S));
    const float sy = static_cast<float>(oy) / S, sx = static_cast<float>(ox) / S;
    const int y0 = min(static_cast<int>(sy), h - 1), x0 = min(static_cast<int>(sx), w - 1);
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const float fy = sy - y0, fx = sx - x0;
    const float2* b = reinterpret_cast<const float2*>(in) + static_cast<size_t>(n) * h * w;
    const float2 tl = __ldg(b + static_cast<size_t>(y0) * w + x0), tr = __ldg(b + static_cast<size_t>(y0) * w + x1);
    const float2 bl = __ldg(b + static_cast<size_t>(y1) * w + x0), br = __ldg(b + static_cast<size_t>(y1) * w + x1);
    // the two-pass form of the TF kernel: rows first (top / bottom interpolated along x), then along y
    const float tx = tl.x * (1.f - fx) + tr.x * fx, bx = bl.x * (1.f - fx) + br.x * fx;
    const float ty = tl.y * (1.f - fx) + tr.y * fx, by = bl.y * (1.f - fx) + br.y * fx;
    reinterpret_cast<float2*>(out)[i] = make_float2((tx * (1.f - fy) + bx * fy) * gain, (ty * (1.f - fy) + by * fy) * gain);
}

inline unsigned blocks_for(long long total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace

cudaError_t init_kernels() {
    cudaError_t e = cudaFuncSetAttribute(conv3x3_narrow_kernel<2>, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024);
    return e;
}

void launch_conv3x3(const PwcConv& p, cudaStream_t st) {
    const long long npix = static_cast<long long>(p.N) * p.Hout * p.Wout;
    if (p.cout == 2) {
        conv3x3_narrow_kernel<2><<<blocks_for(npix, 128), 128, static_cast<size_t>(9) * p.cin * 2 * sizeof(float), st>>>(p);
        return;
    }
    dim3 grid(blocks_for(npix, kPx), (p.cout + kCo - 1) / kCo);
    if (p.leaky) conv3x3_wide_kernel<true><<<grid, 128, 0, st>>>(p);
    else conv3x3_wide_kernel<false><<<grid, 128, 0, st>>>(p);
}

void launch_deconv4x4s2(Planes in, int in_coff, const float* in_f32, int in_f32_cs, int cin, const float* w, const float* b, Planes out, int out_coff,
                        int N, int h, int wd, cudaStream_t st) {
    const unsigned blocks = blocks_for(static_cast<long long>(N) * 4 * h * wd, 128);
    if (in_f32) deconv4x4s2_kernel<true><<<blocks, 128, 0, st>>>(in, in_coff, in_f32, in_f32_cs, cin, w, b, out, out_coff, N, h, wd);
    else deconv4x4s2_kernel<false><<<blocks, 128, 0, st>>>(in, in_coff, in_f32, in_f32_cs, cin, w, b, out, out_coff, N, h, wd);
}

void launch_cost_volume(Planes c1, int c1_coff, Planes c2, int c2_coff, int C, Planes out, int out_coff, int N, int h, int w, cudaStream_t st) {
    dim3 grid((w + kCvTx - 1) / kCvTx, (h + kCvTy - 1) / kCvTy, N);
    cost_volume_kernel<<<grid, kCvThreads, 0, st>>>(c1, c1_coff, c2, c2_coff, C, out, out_coff, N, h, w);
}

// ---------------------------------------------------------------- first pyramid conv: fp32 image [N,H,W,3] -> 16 channels at half resolution
void launch_first_conv(const float* img, const float* w, const float* b, Planes out, int N, int H, int W, cudaStream_t st) {
    first_conv_kernel<<<blocks_for(static_cast<long long>(N) * (H / 2) * (W / 2), 128), 128, 0, st>>>(img, w, b, out, N, H, W);
}

void launch_prepare_pair(const void* src0, const void* src1, bool yuv, float* img1, float* img2, int h, int w, int oh, int ow, int Hp, int Wp,
                         const PrepConst& k, cudaStream_t st) {
    dim3 grid(blocks_for(static_cast<long long>(Hp) * Wp, 256), 2);
    if (yuv) prepare_pair_kernel<true><<<grid, 256, 0, st>>>(src0, src1, img1, img2, h, w, oh, ow, Hp, Wp, k);
    else prepare_pair_kernel<false><<<grid, 256, 0, st>>>(src0, src1, img1, img2, h, w, oh, ow, Hp, Wp, k);
}

void launch_finish_flow(const float* flow, int N, int Hp, int Wp, int h0, int w0, int oh, int ow, float* out, const FinishConst& k, cudaStream_t st) {
    finish_flow_kernel<<<blocks_for(static_cast<long long>(N) * oh * ow, 128), 128, 0, st>>>(flow, N, Hp, Wp, h0, w0, oh, ow, out, k);
}

void launch_copy_channels(Planes src, int src_off, Planes dst, int dst_off, int C, long long npix, cudaStream_t st) {
    copy_channels_kernel<<<blocks_for(npix * (C / 8), 256), 256, 0, st>>>(src, src_off, dst, dst_off, C, npix);
}

void launch_flow_upfeat_scatter(Planes F, float* flow, Planes Dn, int up_off, int N, int h, int w, cudaStream_t st) {
    flow_upfeat_scatter_kernel<<<blocks_for(static_cast<long long>(N) * h * w, 256), 256, 0, st>>>(F, flow, Dn, up_off, N, h, w);
}

void launch_dense_warp(Planes img, int coff, int C, Planes flow, int f_coff, float scale, Planes out, int N, int h, int w, cudaStream_t st) {
    dense_warp_kernel<<<blocks_for(static_cast<long long>(N) * h * w * (C / 4), 256), 256, 0, st>>>(img, coff, C, flow, f_coff, scale, out, N, h, w);
}

void launch_resize_flow(const float* in, float* out, int N, int h, int w, int S, float gain, cudaStream_t st) {
    resize_flow_kernel<<<blocks_for(static_cast<long long>(N) * h * S * w * S, 256), 256, 0, st>>>(in, out, N, h, w, S, gain);
}

}  // namespace pwc
}  // namespace fisr
