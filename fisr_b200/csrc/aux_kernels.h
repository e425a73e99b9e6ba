// Host launchers of the HBM-bound helper kernels (aux_kernels.cu).
#pragma once
#include "common.cuh"

namespace fisr {

// fp16 activation buffer: hi plane at p, lo plane at p + plane (elements); NHWC.
struct ActBuf {
    __half* p = nullptr;
    size_t plane = 0;
};

constexpr int kMaxTiles = 32;
// Geometry of the tiles one batched forward covers (utils.py:118-159 restated on the host, see fisr_api.cu).
struct TileList {
    int count;
    int win[kMaxTiles];                        // which window (frame triple) of the batch the tile reads
    int ylo[kMaxTiles], xlo[kMaxTiles];        // input window origin inside the cropped frame
    int trim_y[kMaxTiles], trim_x[kMaxTiles];  // halo rows / cols to drop from the x2 output
    int out_img[kMaxTiles];                    // destination image (window canvas, or unit slot)
    int out_y[kMaxTiles], out_x[kMaxTiles];    // paste origin inside the destination image
};

// gap_at / gap: input channels >= gap_at are packed `gap` slots further (zero weights in between)
void launch_prep_weights(const float* w, __half* out, int cin, int cout, int KB, int cout_pad, int planes, cudaStream_t st,
                         int gap_at = 1 << 30, int gap = 0);
// f16f8: slot of the first prediction channel in the 64-channel input buffers of levels 2 and 3 (29 real input channels, gap, 9)
constexpr int kPredSlot = 32;
// depth_to_space-folded weights / bias of a conv/2 head: w [3,3,64,cout], b [cout] -> wps [3,3,256,4*cout], bps [4*cout]
void launch_expand_ps_weights(const float* w, const float* b, float* wps, float* bps, int cout, cudaStream_t st);
void launch_pack_input(const float* img, int N, int H, int W, int cin, ActBuf l3, ActBuf l2, ActBuf l1, int planes, cudaStream_t st);
void launch_upsample2(ActBuf in, ActBuf out, int N, int h, int w, int C, int planes, cudaStream_t st);
void launch_act_from_f32(const float* src, int C, ActBuf dst, int cs, size_t npix, int planes, cudaStream_t st);
void launch_act_to_f32(ActBuf src, int cs, int coff, float* dst, int C, size_t npix, int planes, cudaStream_t st);
void launch_tile_pack(const uint8_t* frames, const float* flow, const float* warp, int fh, int fw, const TileList& tiles,
                      int th, int tw, const float* lut255, ActBuf l3, ActBuf l2, ActBuf l1, int planes, cudaStream_t st);
// canvas: images of OH x OW x 9, tile t lands in image out_img[t] at (out_y[t], out_x[t])
// pred: fp32 records of cs floats per pixel (9, or the 12-float layout of the depth_to_space-folded heads)
// f16f8: 12-float prediction records [npix] -> channels 29..37 of the next level's 64-channel input planes
void launch_pred_to_next(const float* pred, ActBuf next, size_t npix, cudaStream_t st);
void launch_pred_compact(const float* src, float* dst, size_t npix, cudaStream_t st);
void launch_tile_unpack_u8(const float* pred, int cs, const TileList& tiles, int th2, int tw2, uint8_t* canvas, int OH, int OW,
                           int core_h, int core_w, cudaStream_t st);
void launch_tile_unpack_f32(const float* pred, int cs, const TileList& tiles, int th2, int tw2, float* canvas, int OH, int OW,
                            int core_h, int core_w, cudaStream_t st);
// jobs warps in one launch: out[i] = frame yuv[src_idx[i]] (or yuv[i] when src_idx is null) sampled along flow[i]
void launch_warp_yuv(const uint8_t* yuv, const float* flow, const int* src_idx, int jobs, float flow_scale, float* out, int h, int w,
                     float out_scale, cudaStream_t st);

}  // namespace fisr
