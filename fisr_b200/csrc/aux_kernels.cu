// HBM-bound helper kernels around the conv stack: weight packing, input packing (with the reference's
// "bicubic" = strided subsample), legacy bilinear x2 upsample, tile pack / unpack for the tiled video path, and the
// flow warp.  (The 2x2 max-pool of ops.py:54 rides in the epilogue of the conv that produces its input.)  All are one-pass, vectorised, coalesced along the channel axis.
#include "common.cuh"
#include "aux_kernels.h"
#include "act_io.cuh"

namespace fisr {

namespace {


// ---------------------------------------------------------------- weights
// w: fp32 HWIO [3,3,cin,cout] (ops.py:8) -> [plane][kb][tap][cout_pad][64] fp16, zero padded.
// gap_at / gap: input channels >= gap_at sit `gap` slots further in the activation buffer (f16f8: the 9 prediction channels of
// the level-2/3 inputs start at slot 32, so that they are written as whole 32-byte sectors, see pred_to_next_kernel).
__global__ void prep_weights_kernel(const float* __restrict__ w, __half* __restrict__ out, int cin, int cout, int KB,
                                    int cout_pad, int planes, int gap_at, int gap) {
    const size_t per_plane = static_cast<size_t>(KB) * 9 * cout_pad * 64;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= per_plane) return;
    const int cl = i & 63;
    size_t r = i >> 6;
    const int co = r % cout_pad; r /= cout_pad;
    const int tap = r % 9;
    const int kb = r / 9;
    const int slot = kb * 64 + cl;
    const int ci = slot < gap_at ? slot : (slot >= gap_at + gap ? slot - gap : cin);      // slots inside the gap carry zero weights
    float v = 0.f;
    if (ci < cin && co < cout) v = w[(static_cast<size_t>(tap) * cin + ci) * cout + co];
    if (planes == 3) {   // f16f8: fp16(128 w) plane + 8-bit rows [e4m3(wh / 16) x 64 | e5m2(128 w - wh) x 64] (common.cuh)
        const float ws = v * kF8WScale;
        const __half wh = __float2half_rn(ws);
        out[i] = wh;
        uint8_t* row = reinterpret_cast<uint8_t*>(out + per_plane) + (i >> 6) * 128;
        row[cl] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(__half2float(wh) * (kF8WHiScale / kF8WScale), __NV_SATFINITE, __NV_E4M3));
        row[64 + cl] = static_cast<uint8_t>(__nv_cvt_float_to_fp8(ws - __half2float(wh), __NV_SATFINITE, __NV_E5M2));
        return;
    }
    const SplitHalf s = split_f32(v);
    out[i] = s.hi;
    if (planes == 2) out[per_plane + i] = s.lo;
}

// conv/2 of the heads (FISRnet.py:100,106) runs on relu(depth_to_space(conv/1)) at 2R resolution with 64 input channels.
// Evaluated at R resolution on the 256-channel pre-shuffle tensor it is a 3x3 conv with 4 * cout output columns (one group per
// output sub-pixel) whose weights are a scatter of the original ones:
//   output sub-pixel (i,j), 2R tap (ky,kx):  r = i + ky - 1 -> R-tap dy = floor(r / 2), source sub-row i' = r mod 2  (same in x)
//   W'[(dy+1)*3 + (dx+1)][(2i'+j')*64 + c][(2i+j)*cout + co] = w[ky*3 + kx][c][co]
// 16 of the 36 (R-tap, source group) pairs are non-zero; SAME padding at the 2R border equals SAME padding at the R border.
__global__ void expand_ps_weights_kernel(const float* __restrict__ w, const float* __restrict__ b, float* __restrict__ wps,
                                         float* __restrict__ bps, int cout) {
    const int i = blockIdx.x * blockDim.x + threadIdx.x;
    const int total = 9 * 4 * 64 * cout;
    if (i < 4 * cout) bps[i] = b[i % cout];
    if (i >= total) return;
    const int co = i % cout;
    int r = i / cout;
    const int c = r % 64; r /= 64;
    const int sub = r % 4;
    const int tap = r / 4;
    const int ky = tap / 3, kx = tap % 3, si = sub >> 1, sj = sub & 1;
    const int ry = si + ky - 1, rx = sj + kx - 1;                 // -1 .. 2
    const int dy = (ry + 2) / 2 - 1, dx = (rx + 2) / 2 - 1;       // floor division
    const int iy = ry - 2 * dy, ix = rx - 2 * dx;
    const int tap2 = (dy + 1) * 3 + (dx + 1), cin2 = (2 * iy + ix) * 64 + c, cout4 = 4 * cout;
    wps[(static_cast<size_t>(tap2) * 256 + cin2) * cout4 + sub * cout + co] = w[(static_cast<size_t>(tap) * 64 + c) * cout + co];
}

// ---------------------------------------------------------------- input pack
// img fp32 NHWC [N,H,W,29] -> level-3 input (full res), level-2 (x[::2, ::2]) and level-1 (x[::4, ::4]) buffers,
// fp16 (hi, lo), 64 channels each; only channels [0, cin) are written (FISRnet.py:81,112,144).
template <int PLANES>
__global__ void pack_input_kernel(const float* __restrict__ img, int N, int H, int W, int cin, __half* l3, size_t p3,
                                  __half* l2, size_t p2, __half* l1, size_t p1) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * H * W * 32;
    if (i >= total) return;
    const int c = i & 31;
    if (c >= cin) return;
    const size_t pix = i >> 5;
    const int x = pix % W;
    const int y = (pix / W) % H;
    const int n = pix / (static_cast<size_t>(W) * H);
    const float v = img[pix * cin + c];
    const SplitHalf s = split_f32(v);
    const uint8_t b_lo = PLANES == 3 ? f8_lo_byte(v - __half2float(s.hi)) : 0, b_hi = PLANES == 3 ? f8_hi_byte(v) : 0;
    auto put8 = [&](__half* plane1, size_t q) {       // 8-bit row of pixel q (64 channels = one block)
        uint8_t* row = reinterpret_cast<uint8_t*>(plane1 + q * 64);
        row[c] = b_lo;
        row[64 + c] = b_hi;
    };
    l3[pix * 64 + c] = s.hi;
    if (PLANES == 2) l3[p3 + pix * 64 + c] = s.lo;
    if (PLANES == 3) put8(l3 + p3, pix);
    if (((x | y) & 1) == 0) {
        const size_t q = (static_cast<size_t>(n) * (H / 2) + y / 2) * (W / 2) + x / 2;
        l2[q * 64 + c] = s.hi;
        if (PLANES == 2) l2[p2 + q * 64 + c] = s.lo;
        if (PLANES == 3) put8(l2 + p2, q);
    }
    if (((x | y) & 3) == 0) {
        const size_t q = (static_cast<size_t>(n) * (H / 4) + y / 4) * (W / 4) + x / 4;
        l1[q * 64 + c] = s.hi;
        if (PLANES == 2) l1[p1 + q * 64 + c] = s.lo;
        if (PLANES == 3) put8(l1 + p1, q);
    }
}

// ---------------------------------------------------------------- legacy bilinear x2 (ops.py:69)
// out[2k] = in[k], out[2k+1] = (in[k] + in[min(k+1, n-1)]) / 2, along H then W (TF-1.13 resize_images).
template <int PLANES>
__global__ void upsample2_kernel(const __half* __restrict__ in, size_t pin, __half* __restrict__ out, size_t pout,
                                 int N, int h, int w, int C) {
    // thread = one input pixel x 8 channels -> its 2 x 2 output block (4 loads, 4 stores; every input is read by at most
    // 4 threads, through L1/L2)
    const int cv = C / 8;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(N) * h * w * cv;
    if (i >= total) return;
    const int c8 = (i % cv) * 8;
    size_t r = i / cv;
    const int x0 = r % w; r /= w;
    const int y0 = r % h;
    const int n = r / h;
    const int y1 = min(y0 + 1, h - 1), x1 = min(x0 + 1, w - 1);
    const __half* base = in + static_cast<size_t>(n) * h * w * C + c8;
    float a[8], b[8], c[8], d[8], o[8];
    load8<PLANES>(base + (static_cast<size_t>(y0) * w + x0) * C, pin, a);
    load8<PLANES>(base + (static_cast<size_t>(y1) * w + x0) * C, pin, b);
    load8<PLANES>(base + (static_cast<size_t>(y0) * w + x1) * C, pin, c);
    load8<PLANES>(base + (static_cast<size_t>(y1) * w + x1) * C, pin, d);
    __half* ob = out + ((static_cast<size_t>(n) * 2 * h + 2 * y0) * 2 * w + 2 * x0) * C + c8;
    const size_t row = static_cast<size_t>(2) * w * C;
    store8<PLANES>(ob, pout, a);                                                  // (2y, 2x)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.5f * (a[j] + c[j]);
    store8<PLANES>(ob + C, pout, o);                                              // (2y, 2x+1)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.5f * (a[j] + b[j]);
    store8<PLANES>(ob + row, pout, o);                                            // (2y+1, 2x)
#pragma unroll
    for (int j = 0; j < 8; ++j) o[j] = 0.5f * (0.5f * (a[j] + b[j]) + 0.5f * (c[j] + d[j]));
    store8<PLANES>(ob + row + C, pout, o);                                        // (2y+1, 2x+1)
}

// ---------------------------------------------------------------- test / debug converters
template <int PLANES>
__global__ void act_from_f32_kernel(const float* __restrict__ src, int C, __half* __restrict__ dst, size_t plane,
                                    int cs, size_t npix) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= npix * cs) return;
    const int c = i % cs;
    const size_t pix = i / cs;
    const float v = c < C ? src[pix * C + c] : 0.f;
    const SplitHalf s = split_f32(v);
    dst[i] = s.hi;
    if (PLANES == 2) dst[plane + i] = s.lo;
    if (PLANES == 3) {
        uint8_t* q = f8_row_ptr(dst + plane + i);
        q[0] = f8_lo_byte(v - __half2float(s.hi));
        q[64] = f8_hi_byte(v);
    }
}
template <int PLANES>
__global__ void act_to_f32_kernel(const __half* __restrict__ src, size_t plane, int cs, int coff, float* __restrict__ dst,
                                  int C, size_t npix) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= npix * C) return;
    const int c = i % C;
    const size_t pix = i / C;
    const size_t s = pix * cs + coff + c;
    dst[i] = PLANES == 2 ? join_f16(src[s], src[plane + s])
             : PLANES == 3 ? join_f8(src[s], *f8_row_ptr(src + plane + s)) : __half2float(src[s]);
}

// ---------------------------------------------------------------- tiled video path
// Builds the network input of T tiles straight from the frame-sized device arrays (FISRnet.py:1008-1024,1035,1044):
//   frames u8 [h,w,9] -> /255 (256-entry table = the reference's float64 division rounded to fp32), clip [0,1]
//   flow  f32 [h,w,8] -> /96 /2, clip [-1,1]          warp f32 [h,w,12] (already /255) -> clip [0,1]
// Tile t covers rows [ylo[t], ylo[t]+th) x cols [xlo[t], xlo[t]+tw) of the (cropped) frame.
template <int PLANES>
__global__ void tile_pack_kernel(const uint8_t* __restrict__ frames, const float* __restrict__ flow,
                                 const float* __restrict__ warp, int fh, int fw, TileList tiles, int th, int tw,
                                 const float* __restrict__ lut255, __half* l3, size_t p3, __half* l2, size_t p2,
                                 __half* l1, size_t p1) {
    // thread = (pixel, group of 8 channels); 4 groups write channels 0..31 (29..31 = 0) as 16-byte vectors
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(tiles.count) * th * tw * 4;
    if (i >= total) return;
    const int g = i & 3;
    const size_t pix = i >> 2;
    const int x = pix % tw;
    const int y = (pix / tw) % th;
    const int t = pix / (static_cast<size_t>(tw) * th);
    const size_t src = (static_cast<size_t>(tiles.win[t]) * fh + tiles.ylo[t] + y) * fw + tiles.xlo[t] + x;
    float v[8];
#pragma unroll
    for (int j = 0; j < 8; ++j) {
        const int c = g * 8 + j;
        float f = 0.f;
        if (c < 9) {
            f = lut255[frames[src * 9 + c]];
        } else if (c < 17) {
            f = __fdiv_rn(__fdiv_rn(flow[src * 8 + (c - 9)], 96.f), 2.f);
            f = fminf(fmaxf(f, -1.f), 1.f);
        } else if (c < 29) {
            f = fminf(fmaxf(warp[src * 12 + (c - 17)], 0.f), 1.f);
        }
        v[j] = f;
    }
    store8<PLANES>(l3 + pix * 64 + g * 8, p3, v);
    if (((x | y) & 1) == 0) {
        const size_t q = (static_cast<size_t>(t) * (th / 2) + y / 2) * (tw / 2) + x / 2;
        store8<PLANES>(l2 + q * 64 + g * 8, p2, v);
    }
    if (((x | y) & 3) == 0) {
        const size_t q = (static_cast<size_t>(t) * (th / 4) + y / 4) * (tw / 4) + x / 4;
        store8<PLANES>(l1 + q * 64 + g * 8, p1, v);
    }
}

// pred_l3 of T tiles [T,2th,2tw,9] fp32 -> trim the halo (utils.py:138-159), clip [0,1], uint8(x*255) with the
// reference's float64 truncation (FISRnet.py:1060-1064), paste into the [OH,OW,9] canvas (FISRnet.py:1056-1057).
// channel k of a prediction pixel: float k of a 9-float record, or float (k / 3) * 4 + k % 3 of a 12-float record (the layout
// the depth_to_space-folded heads write, conv_umma_kernel.cuh epilogue_scalar_ps)
__device__ __forceinline__ int pred_slot(int k, int cs) { return cs == 12 ? (k / 3) * 4 + k % 3 : k; }

template <int CS>
__global__ void tile_unpack_u8_kernel(const float* __restrict__ pred, TileList tiles, int th2, int tw2,
                                      uint8_t* __restrict__ canvas, int OH, int OW, int core_h, int core_w) {
    // thread = 4 consecutive pixels of one row: 9 float4 loads, 9 packed 32-bit stores (all offsets are multiples of 4 px)
    const int qw = core_w / 4;
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(tiles.count) * core_h * qw;
    if (i >= total) return;
    const int x = (i % qw) * 4;
    size_t r = i / qw;
    const int y = r % core_h;
    const int t = r / core_h;
    const float4* src = reinterpret_cast<const float4*>(
        pred + ((static_cast<size_t>(t) * th2 + tiles.trim_y[t] + y) * tw2 + tiles.trim_x[t] + x) * CS);
    uint32_t* dst = reinterpret_cast<uint32_t*>(
        canvas + ((static_cast<size_t>(tiles.out_img[t]) * OH + tiles.out_y[t] + y) * OW + tiles.out_x[t] + x) * 9);
    float f[36];                                     // 4 pixels x 9 channels, in output order
    if (CS == 9) {
#pragma unroll
        for (int j = 0; j < 9; ++j) {
            const float4 v = __ldg(src + j);
            f[4 * j] = v.x; f[4 * j + 1] = v.y; f[4 * j + 2] = v.z; f[4 * j + 3] = v.w;
        }
    } else {                                         // 12-float records: three 16-byte groups of 3 channels per pixel
#pragma unroll
        for (int j = 0; j < 12; ++j) {
            const float4 v = __ldg(src + j);
            const int q = j / 3, g = j % 3;
            f[q * 9 + g * 3] = v.x; f[q * 9 + g * 3 + 1] = v.y; f[q * 9 + g * 3 + 2] = v.z;
        }
    }
#pragma unroll
    for (int j = 0; j < 9; ++j) {
        uint32_t w = 0;
#pragma unroll
        for (int k = 0; k < 4; ++k) {
            const double cl = fmin(fmax(static_cast<double>(f[4 * j + k]), 0.0), 1.0);
            w |= static_cast<uint32_t>(static_cast<int>(cl * 255.0)) << (8 * k);
        }
        dst[j] = w;
    }
}
__global__ void tile_unpack_u8_scalar_kernel(const float* __restrict__ pred, int cs, TileList tiles, int th2, int tw2,
                                             uint8_t* __restrict__ canvas, int OH, int OW, int core_h, int core_w) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(tiles.count) * core_h * core_w * 9;
    if (i >= total) return;
    const int c = i % 9;
    size_t r = i / 9;
    const int x = r % core_w; r /= core_w;
    const int y = r % core_h;
    const int t = r / core_h;
    const float v = pred[((static_cast<size_t>(t) * th2 + tiles.trim_y[t] + y) * tw2 + tiles.trim_x[t] + x) * cs + pred_slot(c, cs)];
    const double cl = fmin(fmax(static_cast<double>(v), 0.0), 1.0);
    canvas[((static_cast<size_t>(tiles.out_img[t]) * OH + tiles.out_y[t] + y) * OW + tiles.out_x[t] + x) * 9 + c] =
        static_cast<uint8_t>(static_cast<int>(cl * 255.0));
}
// same, but keeps fp32 (for parity tests against the oracle's float canvas)
__global__ void tile_unpack_f32_kernel(const float* __restrict__ pred, int cs, TileList tiles, int th2, int tw2,
                                       float* __restrict__ canvas, int OH, int OW, int core_h, int core_w) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    const size_t total = static_cast<size_t>(tiles.count) * core_h * core_w * 9;
    if (i >= total) return;
    const int c = i % 9;
    size_t r = i / 9;
    const int x = r % core_w; r /= core_w;
    const int y = r % core_h;
    const int t = r / core_h;
    canvas[((static_cast<size_t>(tiles.out_img[t]) * OH + tiles.out_y[t] + y) * OW + tiles.out_x[t] + x) * 9 + c] =
        pred[((static_cast<size_t>(t) * th2 + tiles.trim_y[t] + y) * tw2 + tiles.trim_x[t] + x) * cs + pred_slot(c, cs)];
}

// 12-float prediction records of one level -> the prediction channels of the next level's f16f8 input planes (FISRnet.py:113,144:
// img_l2 = concat(bicubic(img), pred_l1), img_l3 = concat(img, pred_l2)).  They occupy slots 32..40 of the 64-channel buffer
// (kPredSlot; the weights of enc/level_0/conv/0 are packed with the matching gap), so that this pass owns whole 32-byte sectors
// -- bytes 64..127 of the fp16 row, 32..63 and 96..127 of the 8-bit row -- and the tile packer owns the others: no partial
// sector is ever written (the first version wrote 9 channels at slot 29 and ran at 0.57 TB/s on DRAM read-modify-writes).
// One thread per pixel: 3 float4 loads, 8 x 16-byte stores.
__global__ void pred_to_next_kernel(const float* __restrict__ pred, __half* __restrict__ next, size_t plane, size_t npix) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= npix) return;
    const float4* src = reinterpret_cast<const float4*>(pred + i * 12);
    const float4 g0 = __ldg(src), g1 = __ldg(src + 1), g2 = __ldg(src + 2);
    const float f[9] = {g0.x, g0.y, g0.z, g1.x, g1.y, g1.z, g2.x, g2.y, g2.z};      // pred channels 0..8 -> slots 32..40
    uint32_t h[5] = {0, 0, 0, 0, 0}, lo[3] = {0, 0, 0}, hi[3] = {0, 0, 0};
#pragma unroll
    for (int k = 0; k < 9; ++k) {
        const SplitHalf s = split_f32(f[k]);
        h[k >> 1] |= static_cast<uint32_t>(__half_as_ushort(s.hi)) << (16 * (k & 1));
        lo[k >> 2] |= static_cast<uint32_t>(f8_lo_byte(f[k] - __half2float(s.hi))) << (8 * (k & 3));
        hi[k >> 2] |= static_cast<uint32_t>(f8_hi_byte(f[k])) << (8 * (k & 3));
    }
    const uint4 z = make_uint4(0u, 0u, 0u, 0u);
    uint4* ph = reinterpret_cast<uint4*>(reinterpret_cast<uint8_t*>(next + i * 64) + 64);       // fp16 plane, slots 32..63
    ph[0] = make_uint4(h[0], h[1], h[2], h[3]);
    ph[1] = make_uint4(h[4], 0u, 0u, 0u);
    ph[2] = z;
    ph[3] = z;
    uint8_t* row = reinterpret_cast<uint8_t*>(next + plane + i * 64);                            // 8-bit row of the pixel's only block
    uint4* pl = reinterpret_cast<uint4*>(row + 32);                                              // e5m2(16 lo), slots 32..63
    pl[0] = make_uint4(lo[0], lo[1], lo[2], 0u);
    pl[1] = z;
    uint4* pq = reinterpret_cast<uint4*>(row + 96);                                              // e5m2(hi), slots 32..63
    pq[0] = make_uint4(hi[0], hi[1], hi[2], 0u);
    pq[1] = z;
}

// 12-float prediction records -> the [.., 9] tensor FISRnet.model returns
__global__ void pred_compact_kernel(const float* __restrict__ src, float* __restrict__ dst, size_t npix) {
    const size_t i = blockIdx.x * static_cast<size_t>(blockDim.x) + threadIdx.x;
    if (i >= npix * 9) return;
    const int k = i % 9;
    dst[i] = src[(i / 9) * 12 + pred_slot(k, 12)];
}

// ---------------------------------------------------------------- flow warp
// One frame pair direction of FISR_for_video_Warp_Img (FISR_for_video_warp_img_with_flo.py:112-128):
//   yuv u8 [h,w,3] --YUV2RGB (:35-45)--> rgb; dst(y,x) = bilinear(rgb, (x,y) + 0.5*flow(y,x)) with OpenCV's
//   remap arithmetic (cv2.remap INTER_LINEAR, BORDER_REPLICATE, :66): coordinates rounded to 1/32 px
//   (cvRound = round-half-even of x*32), 4 taps weighted by the fp32 table (1-fy)(1-fx).., index-clamped border;
//   --RGB2YUV (:48-57)--> float32 0..255 (not rounded).  `scale` multiplies the result (1/255 yields the
//   network's warp input directly, utils.py:51).
// Arithmetic: the reference converts in float64 (numpy) and cv2.remap interpolates the float64 image with its fp32 weight table;
// every quantity is bounded by ~600 and the result is stored as float32, so fp32 FMAs reproduce it to ~1e-4 on the 0..255
// scale (parity bar against cv2.remap: 2e-3, tests/test_gpu_window.py) without touching the FP64 pipe.  Instruction diet (the
// kernel is issue bound, not HBM bound: ~160 instructions per pixel against 23 B): colours are carried in [0,1] so that every
// clip(., 0, 255) of the reference is the free .SAT modifier of the FMA that produces the value, and bytes become floats through
// the 2^23 mantissa trick (PRMT + FADD) instead of I2F.U8 on the quarter-rate XU pipe.
// rgb / 255 of one YUV sample given as floats, T = 255 * Tinv, offset = T @ [16,128,128]  (..warp_img_with_flo.py:35-45): 7 FMAs
__device__ __forceinline__ void yuv2rgb01(float y, float u, float v, float (&rgb)[3]) {
    constexpr double t00 = 0.00456621, t02 = 0.00625893, t11 = -0.00153632, t12 = -0.00318811, t21 = 0.00791071;
    constexpr double o0 = t00 * 16 + t02 * 128, o1 = t00 * 16 + t11 * 128 + t12 * 128, o2 = t00 * 16 + t21 * 128;
    rgb[0] = __saturatef(fmaf(static_cast<float>(t02), v, fmaf(static_cast<float>(t00), y, -static_cast<float>(o0))));
    rgb[1] = __saturatef(fmaf(static_cast<float>(t12), v, fmaf(static_cast<float>(t11), u, fmaf(static_cast<float>(t00), y, -static_cast<float>(o1)))));
    rgb[2] = __saturatef(fmaf(static_cast<float>(t21), u, fmaf(static_cast<float>(t00), y, -static_cast<float>(o2))));
}

// byte k (0..3) of a 32-bit word as a float: PRMT builds the bit pattern of 2^23 + b in one instruction
template <int K>
__device__ __forceinline__ float word_byte_to_float(uint32_t w) {
    return __uint_as_float(__byte_perm(w, 0x4B000000u, 0x7440 | K)) - 8388608.f;
}

// The two horizontally adjacent taps (x0, x0 + 1) of one source row are 6 consecutive bytes at byte offset `off` of the frame.
// WORDS: fetched as the two or three aligned 32-bit words that contain them (the third only when the run starts at byte 3 of a
// word) -- a quarter of the L1 tag / sector work of six byte loads (ncu on the byte-gather version: L1/TEX 85 % busy, 88 % of the
// sectors excessive).  Words are only touched when they hold a needed byte, so nothing outside the frame is read as long as
// the frame starts on a 4-byte boundary and its size is a multiple of 4 (checked by the launcher; otherwise the byte path runs).
// When `second` is false (BORDER_REPLICATE: both taps are the same column) tap b is computed from don't-care bytes and the
// caller gives it weight zero.
template <bool WORDS>
__device__ __forceinline__ void load_tap_pair(const uint8_t* __restrict__ frame, int off, bool second, float (&a)[3], float (&b)[3]) {
    if constexpr (WORDS) {
        const uint32_t* wp = reinterpret_cast<const uint32_t*>(frame) + (off >> 2);
        const uint32_t o = static_cast<uint32_t>(off & 3);
        uint32_t w0 = __ldg(wp), w1 = 0u, w2 = 0u;
        if (second || o > 1) w1 = __ldg(wp + 1);                         // predicated loads, no branches
        if (second && o == 3) w2 = __ldg(wp + 2);
        const uint32_t lo = __funnelshift_r(w0, w1, 8 * o);              // bytes off .. off+3 : y0 u0 v0 y1
        const uint32_t hi = __funnelshift_r(w1, w2, 8 * o);              // bytes off+4 .. off+7: u1 v1 . .
        yuv2rgb01(word_byte_to_float<0>(lo), word_byte_to_float<1>(lo), word_byte_to_float<2>(lo), a);
        yuv2rgb01(word_byte_to_float<3>(lo), word_byte_to_float<0>(hi), word_byte_to_float<1>(hi), b);
    } else {
        const uint8_t* p = frame + off;
        const uint8_t* q = second ? p + 3 : p;
        yuv2rgb01(static_cast<float>(__ldg(p)), static_cast<float>(__ldg(p + 1)), static_cast<float>(__ldg(p + 2)), a);
        yuv2rgb01(static_cast<float>(__ldg(q)), static_cast<float>(__ldg(q + 1)), static_cast<float>(__ldg(q + 2)), b);
    }
}

// one warped pixel: source image `src` [h,w,3] u8 YUV (< 2 GB: 32-bit offsets), flow vector f of the destination pixel (x, y)
// -> YUV float x out_scale
template <bool WORDS>
__device__ __forceinline__ void warp_pixel(const uint8_t* __restrict__ src, float2 f, float flow_scale, int x, int y, int h, int w,
                                           float out_scale255, float (&o)[3]) {
    const float mx = f.x * flow_scale + static_cast<float>(x);      // ..warp_img_with_flo.py:64-65,123
    const float my = f.y * flow_scale + static_cast<float>(y);
    const int ix = __float2int_rn(mx * 32.f), iy = __float2int_rn(my * 32.f);      // cvRound(x * INTER_TAB_SIZE)
    const int sx = ix >> 5, sy = iy >> 5;
    float fx = static_cast<float>(ix & 31) * (1.f / 32.f);
    const float fy = static_cast<float>(iy & 31) * (1.f / 32.f);
    const int x0 = min(max(sx, 0), w - 1), x1 = min(max(sx + 1, 0), w - 1);       // BORDER_REPLICATE
    const int y0 = min(max(sy, 0), h - 1), y1 = min(max(sy + 1, 0), h - 1);
    const bool second = x1 != x0;
    const int w3 = w * 3, c0 = x0 * 3;
    float a[3], b[3], c[3], d[3];
    load_tap_pair<WORDS>(src, y0 * w3 + c0, second, a, b);
    load_tap_pair<WORDS>(src, y1 * w3 + c0, second, c, d);
    // weights of cv2's table: (1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy fx; with x1 == x0 the right taps ARE the left ones, so their
    // weight moves over (a w00 + a w01 = a (1-fy) exactly as fx (1-fy) + (1-fx)(1-fy) rounds -- the sum is exact in fp32 for
    // 5-bit fractions)
    if (!second) fx = 0.f;
    const float w00 = (1.f - fy) * (1.f - fx), w01 = (1.f - fy) * fx, w10 = fy * (1.f - fx), w11 = fy * fx;
    float rgb[3];
#pragma unroll
    for (int k = 0; k < 3; ++k) rgb[k] = fmaf(d[k], w11, fmaf(c[k], w10, fmaf(b[k], w01, a[k] * w00)));
    // RGB2YUV (..warp_img_with_flo.py:48-57) on rgb / 255: yuv / 255 = (T / 255) rgb01 + off / 255, clip = .SAT, then x 255 x out_scale
    constexpr float T[3][3] = {{static_cast<float>(65.481 / 255), static_cast<float>(128.553 / 255), static_cast<float>(24.966 / 255)},
                               {static_cast<float>(-37.797 / 255), static_cast<float>(-74.203 / 255), static_cast<float>(112.0 / 255)},
                               {static_cast<float>(112.0 / 255), static_cast<float>(-93.786 / 255), static_cast<float>(-18.214 / 255)}};
    constexpr float off[3] = {static_cast<float>(16.0 / 255), static_cast<float>(128.0 / 255), static_cast<float>(128.0 / 255)};
#pragma unroll
    for (int k = 0; k < 3; ++k)
        o[k] = __saturatef(fmaf(T[k][2], rgb[2], fmaf(T[k][1], rgb[1], fmaf(T[k][0], rgb[0], off[k])))) * out_scale255;
}

// Warp job i (blockIdx.z): destination out[i] = frame yuv[src_idx[i]] sampled along flow[i].  A warp owns 32 CONSECUTIVE
// destination pixels of kWarpRows rows: neighbouring lanes gather from the same one or two 128-byte lines, the flow load is one
// coalesced 256-byte request, and the 96 output floats of a row segment are transposed through shared memory so that they
// leave as three fully coalesced 128-byte stores.  23 B of HBM traffic per pixel (3 source + 8 flow + 12 destination).
constexpr int kWarpRows = 2;
template <bool WORDS>
__global__ void __launch_bounds__(256) warp_yuv_kernel(const uint8_t* __restrict__ yuv, const float* __restrict__ flow, const int* __restrict__ src_idx,
                                                       float flow_scale, float* __restrict__ out, int h, int w, float out_scale) {
    __shared__ float stage[8][kWarpRows][96];
    const int xb = blockIdx.x * 32, x = xb + threadIdx.x;
    const int y0 = (blockIdx.y * blockDim.y + threadIdx.y) * kWarpRows;
    if (y0 >= h) return;                         // warp-uniform
    const size_t img = static_cast<size_t>(h) * w;
    const int job = blockIdx.z;
    const uint8_t* src = yuv + (src_idx ? static_cast<size_t>(__ldg(src_idx + job)) : static_cast<size_t>(job)) * img * 3;
    out_scale *= 255.f;                          // colours are carried in [0,1] (warp_pixel)
    const bool live = x < w;
    float2 f[kWarpRows];
#pragma unroll
    for (int r = 0; r < kWarpRows; ++r)          // both rows' flow vectors in flight before the dependent gathers
        f[r] = (live && y0 + r < h) ? __ldg(reinterpret_cast<const float2*>(flow + (job * img + static_cast<size_t>(y0 + r) * w + x) * 2)) : make_float2(0.f, 0.f);
#pragma unroll
    for (int r = 0; r < kWarpRows; ++r) {
        if (live && y0 + r < h) {
            float o[3];
            warp_pixel<WORDS>(src, f[r], flow_scale, x, y0 + r, h, w, out_scale, o);
            float* st = &stage[threadIdx.y][r][threadIdx.x * 3];
            st[0] = o[0]; st[1] = o[1]; st[2] = o[2];
        }
    }
    __syncwarp();
    const int nfl = min(32, w - xb) * 3;         // floats of this row segment
#pragma unroll
    for (int r = 0; r < kWarpRows; ++r) {
        if (y0 + r >= h) break;
        float* d = out + (job * img + static_cast<size_t>(y0 + r) * w + xb) * 3;
#pragma unroll
        for (int k = 0; k < 3; ++k) {
            const int i = k * 32 + threadIdx.x;
            if (i < nfl) d[i] = stage[threadIdx.y][r][i];
        }
    }
}

inline unsigned blocks_for(size_t total, int threads) { return static_cast<unsigned>((total + threads - 1) / threads); }

}  // namespace

// ================================================================ host launchers
void launch_prep_weights(const float* w, __half* out, int cin, int cout, int KB, int cout_pad, int planes,
                         cudaStream_t st, int gap_at, int gap) {
    const size_t total = static_cast<size_t>(KB) * 9 * cout_pad * 64;
    prep_weights_kernel<<<blocks_for(total, 256), 256, 0, st>>>(w, out, cin, cout, KB, cout_pad, planes, gap_at, gap);
}

void launch_expand_ps_weights(const float* w, const float* b, float* wps, float* bps, int cout, cudaStream_t st) {
    cudaMemsetAsync(wps, 0, static_cast<size_t>(9) * 256 * 4 * cout * sizeof(float), st);
    const int total = 9 * 4 * 64 * cout;
    expand_ps_weights_kernel<<<(total + 255) / 256, 256, 0, st>>>(w, b, wps, bps, cout);
}

void launch_pack_input(const float* img, int N, int H, int W, int cin, ActBuf l3, ActBuf l2, ActBuf l1, int planes,
                       cudaStream_t st) {
    const size_t total = static_cast<size_t>(N) * H * W * 32;
    if (planes == 3)
        pack_input_kernel<3><<<blocks_for(total, 256), 256, 0, st>>>(img, N, H, W, cin, l3.p, l3.plane, l2.p, l2.plane,
                                                                     l1.p, l1.plane);
    else if (planes == 2)
        pack_input_kernel<2><<<blocks_for(total, 256), 256, 0, st>>>(img, N, H, W, cin, l3.p, l3.plane, l2.p, l2.plane,
                                                                     l1.p, l1.plane);
    else
        pack_input_kernel<1><<<blocks_for(total, 256), 256, 0, st>>>(img, N, H, W, cin, l3.p, l3.plane, l2.p, l2.plane,
                                                                     l1.p, l1.plane);
}

void launch_upsample2(ActBuf in, ActBuf out, int N, int h, int w, int C, int planes, cudaStream_t st) {
    const size_t total = static_cast<size_t>(N) * h * w * (C / 8);
    if (planes == 3)
        upsample2_kernel<3><<<blocks_for(total, 256), 256, 0, st>>>(in.p, in.plane, out.p, out.plane, N, h, w, C);
    else if (planes == 2)
        upsample2_kernel<2><<<blocks_for(total, 256), 256, 0, st>>>(in.p, in.plane, out.p, out.plane, N, h, w, C);
    else
        upsample2_kernel<1><<<blocks_for(total, 256), 256, 0, st>>>(in.p, in.plane, out.p, out.plane, N, h, w, C);
}

void launch_act_from_f32(const float* src, int C, ActBuf dst, int cs, size_t npix, int planes, cudaStream_t st) {
    if (planes == 3)
        act_from_f32_kernel<3><<<blocks_for(npix * cs, 256), 256, 0, st>>>(src, C, dst.p, dst.plane, cs, npix);
    else if (planes == 2)
        act_from_f32_kernel<2><<<blocks_for(npix * cs, 256), 256, 0, st>>>(src, C, dst.p, dst.plane, cs, npix);
    else
        act_from_f32_kernel<1><<<blocks_for(npix * cs, 256), 256, 0, st>>>(src, C, dst.p, dst.plane, cs, npix);
}

void launch_act_to_f32(ActBuf src, int cs, int coff, float* dst, int C, size_t npix, int planes, cudaStream_t st) {
    if (planes == 3)
        act_to_f32_kernel<3><<<blocks_for(npix * C, 256), 256, 0, st>>>(src.p, src.plane, cs, coff, dst, C, npix);
    else if (planes == 2)
        act_to_f32_kernel<2><<<blocks_for(npix * C, 256), 256, 0, st>>>(src.p, src.plane, cs, coff, dst, C, npix);
    else
        act_to_f32_kernel<1><<<blocks_for(npix * C, 256), 256, 0, st>>>(src.p, src.plane, cs, coff, dst, C, npix);
}

void launch_tile_pack(const uint8_t* frames, const float* flow, const float* warp, int fh, int fw, const TileList& tiles,
                      int th, int tw, const float* lut255, ActBuf l3, ActBuf l2, ActBuf l1, int planes, cudaStream_t st) {
    const size_t total = static_cast<size_t>(tiles.count) * th * tw * 4;
    if (planes == 3)
        tile_pack_kernel<3><<<blocks_for(total, 256), 256, 0, st>>>(frames, flow, warp, fh, fw, tiles, th, tw, lut255, l3.p,
                                                                    l3.plane, l2.p, l2.plane, l1.p, l1.plane);
    else if (planes == 2)
        tile_pack_kernel<2><<<blocks_for(total, 256), 256, 0, st>>>(frames, flow, warp, fh, fw, tiles, th, tw, lut255, l3.p,
                                                                    l3.plane, l2.p, l2.plane, l1.p, l1.plane);
    else
        tile_pack_kernel<1><<<blocks_for(total, 256), 256, 0, st>>>(frames, flow, warp, fh, fw, tiles, th, tw, lut255, l3.p,
                                                                    l3.plane, l2.p, l2.plane, l1.p, l1.plane);
}

void launch_pred_to_next(const float* pred, ActBuf next, size_t npix, cudaStream_t st) {
    pred_to_next_kernel<<<blocks_for(npix, 256), 256, 0, st>>>(pred, next.p, next.plane, npix);
}

void launch_pred_compact(const float* src, float* dst, size_t npix, cudaStream_t st) {
    pred_compact_kernel<<<blocks_for(npix * 9, 256), 256, 0, st>>>(src, dst, npix);
}

void launch_tile_unpack_u8(const float* pred, int cs, const TileList& tiles, int th2, int tw2, uint8_t* canvas, int OH, int OW,
                           int core_h, int core_w, cudaStream_t st) {
    // vector path needs every row segment to start on a 4-pixel boundary (36-byte groups are then 4-byte aligned)
    bool vec = (core_w % 4 == 0) && (tw2 % 4 == 0) && (OW % 4 == 0);
    for (int t = 0; t < tiles.count; ++t) vec = vec && (tiles.trim_x[t] % 4 == 0) && (tiles.out_x[t] % 4 == 0);
    if (vec) {
        const size_t total = static_cast<size_t>(tiles.count) * core_h * (core_w / 4);
        if (cs == 12) tile_unpack_u8_kernel<12><<<blocks_for(total, 256), 256, 0, st>>>(pred, tiles, th2, tw2, canvas, OH, OW, core_h, core_w);
        else tile_unpack_u8_kernel<9><<<blocks_for(total, 256), 256, 0, st>>>(pred, tiles, th2, tw2, canvas, OH, OW, core_h, core_w);
    } else {
        const size_t total = static_cast<size_t>(tiles.count) * core_h * core_w * 9;
        tile_unpack_u8_scalar_kernel<<<blocks_for(total, 256), 256, 0, st>>>(pred, cs, tiles, th2, tw2, canvas, OH, OW, core_h, core_w);
    }
}
void launch_tile_unpack_f32(const float* pred, int cs, const TileList& tiles, int th2, int tw2, float* canvas, int OH, int OW,
                            int core_h, int core_w, cudaStream_t st) {
    const size_t total = static_cast<size_t>(tiles.count) * core_h * core_w * 9;
    tile_unpack_f32_kernel<<<blocks_for(total, 256), 256, 0, st>>>(pred, cs, tiles, th2, tw2, canvas, OH, OW, core_h, core_w);
}

void launch_warp_yuv(const uint8_t* yuv, const float* flow, const int* src_idx, int jobs, float flow_scale, float* out, int h, int w,
                     float out_scale, cudaStream_t st) {
    dim3 block(32, 8), grid((w + 31) / 32, (h + 8 * kWarpRows - 1) / (8 * kWarpRows), jobs);
    // aligned word gathers need frames that start on a 4-byte boundary and are a multiple of 4 bytes long (every frame then is)
    const bool words = (reinterpret_cast<uintptr_t>(yuv) & 3) == 0 && ((static_cast<size_t>(h) * w * 3) & 3) == 0;
    if (words) warp_yuv_kernel<true><<<grid, block, 0, st>>>(yuv, flow, src_idx, flow_scale, out, h, w, out_scale);
    else warp_yuv_kernel<false><<<grid, block, 0, st>>>(yuv, flow, src_idx, flow_scale, out, h, w, out_scale);
}

}  // namespace fisr
