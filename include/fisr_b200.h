/* fisr_b200 -- C ABI of the B200-native FISRnet hot path.
 *
 * The reference (JihyongOh/FISR @ 34d9305) is pure Python on TensorFlow 1.13 and has no FFI of its own; the
 * operator surface that `main.py` drives is the Python class `FISRnet` (FISRnet.py:14-17).  This library is what a
 * ctypes binding underneath that class calls (fisr_b200/_lib.py; INTEGRATION.md shows the stub).  Each entry point
 * names the reference code it stands in for.
 *
 * Conventions: every function returns 0 on success or a negative FISR_E_* code; fisr_last_error() gives the text.
 * No exceptions cross the boundary.  Pointers named d_* are CUDA device pointers on the context's device, h_* are
 * host pointers (pinned memory makes the host entry points faster but is not required).  The caller owns every
 * buffer it passes; the context owns its workspace.  A context is bound to one GPU and is not thread-safe (the
 * reference drives one tf.Session from one Python thread, main.py:137-139).  All tensors are NHWC, float32 unless
 * the name says u8.
 */
#ifndef FISR_B200_H
#define FISR_B200_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

typedef struct fisr_ctx fisr_ctx;

enum {
    FISR_OK = 0,
    FISR_E_INVALID = -1,    /* bad argument (shape not a multiple of 32, unknown parameter name, ...) */
    FISR_E_CUDA = -2,       /* CUDA runtime / driver error */
    FISR_E_KERNEL = -3,     /* a kernel reported a pipeline time-out through its error flag */
    FISR_E_NOMEM = -4,
    FISR_E_OVERFLOW = -5    /* non-finite gradient: the loss scale overflowed the fp16 gradient planes (fisr_train_backward) */
};

/* Arithmetic of the conv stack (DESIGN.md "Numerics").  The reference computes in fp32 (cuDNN on Pascal). */
enum {
    FISR_PREC_F16X3 = 0,    /* fp16 (hi,lo) split operands, 3 tcgen05 MMAs per K-slice, fp32 accumulate: fp32-class */
    FISR_PREC_F16 = 1,      /* single fp16 operands, 1 MMA per K-slice: fast mode, ~7e-4 max-abs on the cascade   */
    FISR_PREC_F16F8 = 2     /* fp16 main term + both cross terms as fp8 (e5m2 x e4m3) MMAs at twice the rate: 2 MMA units
                               per K-slice, ~3e-5 max-abs on the cascade; inference only (training needs F16X3)     */
};

/* ---- lifetime --------------------------------------------------------------------------------------------- */
/* Stands in for `tf.Session(...)` + `FISRnet(sess, args)` (main.py:137-143, FISRnet.py:17). */
int fisr_create(int device, fisr_ctx** out);
void fisr_destroy(fisr_ctx* ctx);
/* Text of the last error on this context (ctx may be NULL for a failed fisr_create). Never NULL. */
const char* fisr_last_error(const fisr_ctx* ctx);
int fisr_set_precision(fisr_ctx* ctx, int precision);
int fisr_get_precision(const fisr_ctx* ctx);

/* ---- parameters: the 138 (w, b) pairs created by tf.get_variable in ops.py:8-9 under scope "FISRnet/" ---- */
/* Names are the TF variable names without the ":0" suffix, e.g. "FISRnet/level_1/enc/level_0/conv/0/w";
 * w is HWIO [3,3,Cin,Cout], b is [Cout]; creation order of FISRnet.py:78-171. */
int fisr_num_params(void);
const char* fisr_param_name(int index);
int fisr_param_shape(int index, int dims[4]);               /* returns rank (4 for w, 1 for b) */
int fisr_set_param(fisr_ctx* ctx, const char* name, const float* h_data, size_t count);   /* saver.restore, FISRnet.py:1109 */
int fisr_get_param(fisr_ctx* ctx, const char* name, float* h_data, size_t count);         /* saver.save,    FISRnet.py:1098 */

/* ---- FISRnet.model (FISRnet.py:73-173) -------------------------------------------------------------------- */
/* img [N,H,W,29] -> pred_l1 [N,H/2,W/2,9], pred_l2 [N,H,W,9], pred_l3 [N,2H,2W,9]; H, W multiples of 32.
 * Output pointers may be NULL to skip that copy.  `stream` is a cudaStream_t (NULL = the context's stream);
 * the call is asynchronous with respect to the host on that stream. */
int fisr_forward(fisr_ctx* ctx, const float* d_img, int N, int H, int W, float* d_pred_l1, float* d_pred_l2,
                 float* d_pred_l3, void* stream);
/* Same through host buffers: the `sess.run(test_Pred, feed_dict=...)` of FISRnet.py:1048 (H2D, forward, D2H, sync). */
int fisr_forward_host(fisr_ctx* ctx, const float* h_img, int N, int H, int W, float* h_pred_l1, float* h_pred_l2,
                      float* h_pred_l3);

/* ---- one sliding window of FISR_for_video / test (FISRnet.py:994-1065, 798-883) ---------------------------- */
/* frames u8 [H,W,9] (3 consecutive YUV frames), flow f32 [H,W,8] in LR pixels, warp f32 [H,W,12] already /255
 * (utils.py:51).  Crops to h = H - H % (32*pH), w likewise (:1006-1007), normalises and clips (:1011-1021), runs the
 * pH x pW tile grid with the 32-pixel halo (utils.py:118-159), trims, pastes, clips to [0,1] and writes
 * uint8(x*255) into canvas u8 [2h,2w,9] (:1060-1064).  Only tiles [tile_first, tile_first+tile_count) of the
 * row-major grid are computed and pasted, so ranks of a multi-GPU job can shard one window; pass 0, pH*pW for all. */
int fisr_window_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int H, int W,
                       int pH, int pW, int tile_first, int tile_count, uint8_t* d_canvas, void* stream);
/* Batched form over B windows: frames u8 [B,H,W,9], flow [B,H,W,8], warp [B,H,W,12].  A unit is one (window, tile)
 * pair, id = window * (pH*pW) + tile; h_units lists the units this call computes (the shard of this rank).  Units of
 * equal tile size run as ONE batched forward.  layout 0: d_out is [B,2h,2w,9], tiles pasted into their frames;
 * layout 1: d_out is [n_units, 2h/pH, 2w/pW, 9], unit i of the list in slot i -- the contiguous send buffer of the
 * one all-gather the multi-GPU path needs (the loops of FISRnet.py:994,1028 carry no dependence). */
int fisr_units_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int B, int H, int W,
                      int pH, int pW, const int* h_units, int n_units, int layout, uint8_t* d_out, void* stream);
int fisr_window_host(fisr_ctx* ctx, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H, int W,
                     int pH, int pW, uint8_t* h_canvas);
/* Pipelined form of fisr_window_host for clips: submit() enqueues H2D (copy stream), the tiles (context stream) and
 * D2H (copy stream) for one window and returns at once; wait() blocks until that window's canvas is in h_canvas.
 * Two slots (0, 1): submitting window k+1 before waiting for window k overlaps its copies with window k's kernels.
 * Host buffers must stay valid (and should be pinned) until wait() returns. */
int fisr_window_submit(fisr_ctx* ctx, int slot, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H,
                       int W, int pH, int pW, uint8_t* h_canvas);
int fisr_window_wait(fisr_ctx* ctx, int slot);
/* float canvas [2h,2w,9] before clipping: `test_Pred_full` of FISRnet.py:844,880 -- what FISRnet.test() clips and scores
 * (PSNR on the float prediction, FISRnet.py:883-887) before it truncates to uint8 for the PNGs (:901). */
int fisr_window_host_f32(fisr_ctx* ctx, const uint8_t* h_frames, const float* h_flow, const float* h_warp, int H, int W,
                         int pH, int pW, float* h_canvas);
int fisr_window_device_f32(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const float* d_warp, int H,
                           int W, int pH, int pW, float* d_canvas, void* stream);

/* ---- flow warp (FISR_tfoptflow/FISR_for_video_warp_img_with_flo.py:61-67,112-128) -------------------------- */
/* yuv u8 [h,w,3] --YUV2RGB--> cv2.remap(INTER_LINEAR, BORDER_REPLICATE) at (x,y) + flow_scale*flow --RGB2YUV-->
 * out f32 [h,w,3] = value * out_scale (the reference stores 0..255, i.e. out_scale 1; 1/255 feeds the network). */
int fisr_warp_device(fisr_ctx* ctx, const uint8_t* d_yuv, const float* d_flow, float flow_scale, float* d_out, int h,
                     int w, float out_scale, void* stream);
/* All warps of a clip in ONE launch (the loop of ..warp_img_with_flo.py:112-128): job i writes d_out[i] [h,w,3] = frame
 * d_frames[d_src_index[i]] (u8 [F,h,w,3]; NULL index = frame i) sampled along d_flow[i] [h,w,2].  For pair k of the reference
 * loop: job 2k = frame k+1 along flow(k -> k+1), job 2k+1 = frame k along flow(k+1 -> k). */
int fisr_warp_batch_device(fisr_ctx* ctx, const uint8_t* d_frames, const float* d_flow, const int* d_src_index, int jobs,
                           float flow_scale, float* d_out, int h, int w, float out_scale, void* stream);
int fisr_warp_host(fisr_ctx* ctx, const uint8_t* h_yuv, const float* h_flow, float flow_scale, float* h_out, int h,
                   int w, float out_scale);

/* ---- training step: `sess.run([self.optim, ...])` of FISRnet.py:651 -------------------------------------------- */
/* Groups2Ovlp (ops.py:119-144): d_pred [3B,H,W,9] = pred of windows 0,1,2 (window-major) -> d_out [B,7,H,W,3]. */
int fisr_groups2ovlp(fisr_ctx* ctx, const float* d_pred, int B, int H, int W, float* d_out, void* stream);
/* Multi-scale temporal loss + train PSNR (FISRnet.py:312-486).  d_pred_l* are the three outputs of FISRnet.model on the
 * 4B-batch [window 0 | window 1 | window 2 | stride 2] (pass-major) of an LR size h x w; d_label [B,2h,2w,21].
 * lambdas = {recn, tm1, tm2, tmm, td, ss2} (NULL = main.py:80-85 defaults).  h_out[11] = recnLoss, tmLoss, tmmLoss, tdLoss,
 * totalLoss_s1, recnLoss_ss2, tdLoss_ss2, tmLoss_ss2, totalLoss_ss2, total_loss, train_PSNR (FISRnet.py:651-657).
 * Synchronous (reads the scalars back). */
int fisr_temporal_loss(fisr_ctx* ctx, const float* d_pred_l1, const float* d_pred_l2, const float* d_pred_l3,
                       const float* d_label, int B, int h, int w, const float* lambdas, float* h_out, void* stream);
/* The forward half of `sess.run([optim, ...])` (FISRnet.py:651): assembles the four weight-shared passes from
 * data [B,h,w,15], flow [..,16], flow_ss2 [..,8], warp [..,24], warp_ss2 [..,12] (FISRnet.py:281-306,392-399), runs them as
 * ONE batched forward and evaluates the loss scalars. */
int fisr_train_forward(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                       const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float* h_out,
                       void* stream);
/* tf.train.AdamOptimizer(lr) update (FISRnet.py:489-491, TF-1.13 formula: lr_t = lr*sqrt(1-b2^t)/(1-b1^t),
 * theta -= lr_t*m/(sqrt(v)+eps)) of all 276 tensors from device gradients listed in creation order (w, b, w, b, ...);
 * keeps m, v and the step counter in the context and re-packs the operand planes.  TF defaults: 0.9, 0.999, 1e-8. */
int fisr_adam_step(fisr_ctx* ctx, const float* const* d_grads, int n_grads, float lr, float beta1, float beta2, float eps);
/* Forward + loss + backward of one step: d total_loss / d (every w, b) by hand-written dgrad (the forward tcgen05 kernel
 * on rotated-transposed operand planes, ReLU gate / residual / space-to-depth fused in the epilogue) and wgrad kernels,
 * what `AdamOptimizer.minimize` derives from the graph of FISRnet.py:281-484.  Gradients stay in the context
 * (fisr_get_grad, fisr_adam_apply); h_out[11] as fisr_temporal_loss.  Needs FISR_PREC_F16X3.  Synchronous. */
int fisr_train_backward(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                        const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float* h_out,
                        void* stream);
/* Gradient of the last fisr_train_backward for one variable (same names / shapes as fisr_get_param). */
int fisr_get_grad(fisr_ctx* ctx, const char* name, float* h_data, size_t count);
/* Adam update (same formula as fisr_adam_step) from the context's own gradients. */
int fisr_adam_apply(fisr_ctx* ctx, float lr, float beta1, float beta2, float eps);
/* fisr_train_backward + fisr_adam_apply(lr, 0.9, 0.999, 1e-8): one `sess.run(optim)` (FISRnet.py:489-491, 651). */
int fisr_train_step(fisr_ctx* ctx, const float* d_data, const float* d_flow, const float* d_flow_ss2, const float* d_warp,
                    const float* d_warp_ss2, const float* d_label, int B, int h, int w, const float* lambdas, float lr,
                    float* h_out, void* stream);
/* Gradients travel through the network as fp16 (hi, lo) planes multiplied by a power-of-two loss scale (default:
 * 2^floor(log2(B*2h*2w*3)), divided out by the weight-gradient reduction).  0 restores the default; a non-finite gradient
 * makes fisr_train_backward return FISR_E_OVERFLOW.  fisr_train_step scales dynamically: on overflow it skips the update,
 * lowers the scale (to the default, then by 16 per retry, up to 5 retries) and repeats the step; the lowered scale stays in force. */
int fisr_set_loss_scale(fisr_ctx* ctx, float scale);
float fisr_get_loss_scale(fisr_ctx* ctx, int B, int h, int w);
/* Weight gradients of layers with >= 16384 pixels multiply dy (hi, lo) by the hi plane of the forward activation only
 * (half the MMAs; adds ~1e-4 relative rounding noise per gradient tensor).  exact = 1 uses both planes everywhere. */
int fisr_set_wgrad_exact(fisr_ctx* ctx, int exact);
long long fisr_adam_steps(const fisr_ctx* ctx);
/* Frees the moments (they restart at zero) and sets the step counter. */
int fisr_adam_reset(fisr_ctx* ctx, long long step);
/* Optimizer state of a training checkpoint: tf.train.Saver stores "<var>/Adam" (m), "<var>/Adam_1" (v) and the beta powers
 * next to the weights (FISRnet.py:1092-1099), so a resumed run continues seamlessly.  which = 0 -> m, 1 -> v; names / shapes as
 * fisr_get_param.  fisr_adam_set_steps sets t (beta1_power = beta1^t) without touching the moments. */
int fisr_get_adam_slot(fisr_ctx* ctx, const char* name, int which, float* h_data, size_t count);
int fisr_set_adam_slot(fisr_ctx* ctx, const char* name, int which, const float* h_data, size_t count);
int fisr_adam_set_steps(fisr_ctx* ctx, long long step);

/* ---- multi-GPU: frames exchanged over NVLink peer memory (the reference is single-GPU, main.py:19-20) -------- */
/* One process per GPU.  Every rank keeps its output frames [B,2h,2w,9] u8 in a buffer from fisr_ipc_alloc, hands the 64-byte
 * handle (a cudaIpcMemHandle_t) to its peers over any channel (torch.distributed), and maps theirs with fisr_ipc_open.  A rank
 * computes its (window, tile) units straight into its own frames (fisr_units_device, layout 0) and pushes each finished tile
 * rectangle into the same place of every peer's frames with fisr_copy2d_async: cudaMemcpy2DAsync on peer-mapped pointers runs on
 * the copy engines over NVLink 5 / NVSwitch -- an all-gather in frame layout with no SM-resident collective kernel competing with
 * the persistent conv grids and no re-assembly pass.  (fisr_b200/sharding.py: PeerFrames; NCCL all-gather remains the fallback.) */
int fisr_ipc_alloc(fisr_ctx* ctx, size_t bytes, void** d_ptr, unsigned char* handle64);
int fisr_ipc_open(fisr_ctx* ctx, const unsigned char* handle64, void** d_ptr);
int fisr_ipc_close(fisr_ctx* ctx, void* d_ptr);
int fisr_ipc_free(fisr_ctx* ctx, void* d_ptr);
/* dst / src: device pointers (local or peer-mapped), pitches and width in bytes; asynchronous on `stream`. */
int fisr_copy2d_async(fisr_ctx* ctx, void* d_dst, size_t dst_pitch, const void* d_src, size_t src_pitch, size_t width_bytes,
                      size_t rows, void* stream);

/* ---- introspection / test hooks ---------------------------------------------------------------------------- */
/* Single 3x3 SAME conv through the production kernel (ops.py:7-11 plus the fused epilogue):
 * y = conv(x, w) + b (+ res); raw = y; act = relu ? max(y,0) : y, optionally depth_to_space(2) (FISRnet.py:99).
 * x [N,H,W,Cin], w HWIO, res / raw [N,H,W,Cout], act [N,H,W,Cout] or [N,2H,2W,Cout/4]; device pointers, any may be
 * NULL except x, w, b (d2s excludes res and raw: the network never combines them).  Synchronous. */
int fisr_conv3x3(fisr_ctx* ctx, const float* d_x, const float* d_w, const float* d_b, const float* d_res, int N, int H,
                 int W, int Cin, int Cout, int relu, int d2s, float* d_raw, float* d_act);
/* Data gradient of one 3x3 SAME conv through the production kernel: dx = conv3x3(dy, rot180(w)^T) * [mask > 0] + res.
 * dy [N,H,W,Cout], w HWIO [3,3,Cin,Cout] (the FORWARD filter), mask / res / raw / act [N,H,W,Cin]; with s2d the act
 * output is space_to_depth(2): [N,H/2,W/2,4*Cin] (adjoint of the depth_to_space epilogue).  Synchronous. */
int fisr_dgrad3x3(fisr_ctx* ctx, const float* d_dy, const float* d_w, const float* d_mask, const float* d_res, int N, int H,
                  int W, int Cin, int Cout, int s2d, float* d_raw, float* d_act);
/* Weight / bias gradient of one 3x3 SAME conv through the production wgrad kernel (the backward-filter op TF derives
 * for ops.py:10): gw[ky,kx,ci,co] = scale * sum_p x[p+(ky-1,kx-1),ci] * dy[p,co], gb[co] = scale * sum_p dy[p,co].
 * x [N,H,W,Cin], dy [N,H,W,Cout], gw HWIO [3,3,Cin,Cout], gb [Cout] (may be NULL); device pointers.  Synchronous. */
int fisr_wgrad3x3(fisr_ctx* ctx, const float* d_x, const float* d_dy, int N, int H, int W, int Cin, int Cout, float scale,
                  float* d_gw, float* d_gb);
/* After a forward: copies the pre-activation output of the named conv (e.g. ".../enc/level_0/conv/0") as
 * float32 NHWC to host, when the plan materialises it; returns FISR_E_INVALID otherwise. */
int fisr_debug_conv_output(fisr_ctx* ctx, const char* conv_name, float* h_dst, size_t count);
/* Per-launch device times of the plan for (N,H,W): runs it `reps` times op by op with CUDA events between launches
 * (no graph) and returns the op count; with ms == NULL only returns the count.  flops / bytes are the algorithmic
 * figures of each launch (2*9*Cin*Cout*h*w*N; operands + outputs once), kinds[i] = 0 conv (+ its N tile), 1000
 * upsample, 2000 prediction -> next level input; names is max_ops strings of name_stride bytes.  Feeds bench.py's roofline block. */
int fisr_profile_ops(fisr_ctx* ctx, int N, int H, int W, int reps, int max_ops, float* ms, double* flops, double* bytes,
                     int* kinds, char* names, int name_stride);
/* Per-op device times of the backward pass for batch B (4B passes) at LR size h x w, like fisr_profile_ops; call
 * fisr_train_backward once first.  Names end in [dgrad] / [wgrad]. */
int fisr_profile_train(fisr_ctx* ctx, int B, int h, int w, int reps, int max_ops, float* ms, double* flops, char* names,
                       int name_stride);
/* Kernel launches issued by this context since creation (the `gpu_launches` evidence bench.py reports). */
long long fisr_launch_count(const fisr_ctx* ctx);
/* Conv FLOPs (2*9*Cin*Cout*h*w*N) and mean MMA row efficiency of the plan for (N,H,W). */
int fisr_plan_info(fisr_ctx* ctx, int N, int H, int W, double* flops, double* mma_efficiency, int* num_launches,
                   size_t* workspace_bytes);

/* ---- PWC-Net inference: the flow estimator in front of the warp ------------------------------------- */
/* `ModelPWCNet(mode='test')` + `predict_from_img_pairs` (FISR_tfoptflow/model_pwcnet.py:204-237,957-1006) configured as
 * FISR_for_video_pwcnet_predict_from_img_test.py:96-110 sets it: PWC-Net-large, dense + residual connections, 6 pyramid levels,
 * flow predicted at level 2, search range 4.  Parameters carry the TensorFlow variable names of tfoptflow's checkpoints,
 * "pwcnet/featpyr/conv1a/kernel" ... "pwcnet/upsample/up_feat3/bias" (182 tensors, 14,079,050 values); conv kernels are HWIO
 * [3,3,Cin,Cout], conv2d_transpose kernels [4,4,2,Cin].  PARITY UNPINNED: eight modules of the reference's PWC-Net copy and its
 * checkpoint are not in the tree (model_pwcnet.py:21-28); the CUDA path is checked against oracle/pwcnet_oracle.py.
 * Every 3x3 conv with >= 16 outputs on an image of >= 4 x 4 pixels runs on the tcgen05 kernel in split mode (fp16 hi / lo planes,
 * fp32-class): dilated layers as polyphase launches, stride-2 layers as two row-phase launches, the flow predictor fused with the
 * next level's up_feat transposed conv.  FISR_PWC_UMMA=0..3 in the environment of fisr_pwc_create selects how much of that is
 * used (0: CUDA-core kernels only ... 3: everything, default) for A/B measurements. */
typedef struct fisr_pwc fisr_pwc;
int fisr_pwc_create(int device, fisr_pwc** out);
void fisr_pwc_destroy(fisr_pwc* pwc);
const char* fisr_pwc_last_error(const fisr_pwc* pwc);
int fisr_pwc_num_params(void);
const char* fisr_pwc_param_name(int index);
int fisr_pwc_param_shape(int index, int dims[4]);
int fisr_pwc_set_param(fisr_pwc* pwc, const char* name, const float* h_data, size_t count);
/* `nn()` (model_pwcnet.py:1525-1593) on N image pairs: d_img1, d_img2 f32 [N,H,W,3] in 0..1 (adapt_x: /255, zero-padded so that
 * H, W are multiples of 64) -> d_flow f32 [N,H,W,2] = flow from image 1 to image 2 in pixels, (u, v) per pixel.  Asynchronous. */
int fisr_pwc_forward(fisr_pwc* pwc, const float* d_img1, const float* d_img2, int N, int H, int W, float* d_flow, void* stream);
/* The driver around the network for ONE frame pair (FISR_for_video_pwcnet_predict_from_img_test.py:113-131, adapt_x
 * model_pwcnet.py:371-409), on the device: [YUV -> RGB,] x`scale` skimage.transform.resize (order 1, 'reflect'), uint8 truncation,
 * /255, zero pad to multiples of 64, written as the batch of both directions: d_img1 = [frame 0, frame 1], d_img2 = [frame 1,
 * frame 0], each f32 [2,Hp,Wp,3] with Hp, Wp = h * scale, w * scale rounded up to 64.  src_kind 0: uint8 YUV frames [h,w,3]
 * (h_yuv2rgb = 12 doubles, the 3x3 matrix T then the offset of utils.py:106-115, as the host computes them); src_kind 1: float64
 * RGB frames [h,w,3] in 0..255 (h_yuv2rgb ignored).  Arithmetic is float64 in numpy's evaluation order.  Asynchronous. */
int fisr_pwc_prepare_pair(fisr_pwc* pwc, const void* d_frame0, const void* d_frame1, int src_kind, const double* h_yuv2rgb, int h, int w,
                          int scale, float* d_img1, float* d_img2, void* stream);
/* postproc_y_hat_test (model_pwcnet.py:449-470) + ..predict_from_img_test.py:137: d_flow f32 [N,Hp,Wp,2] cropped to h0 x w0,
 * Gaussian pre-filter (scipy.ndimage.gaussian_filter, mode 'mirror'; h_wy / h_wx = the ry + 1 / rx + 1 kernel weights at distance
 * 0, 1, .. as scipy computes them; radius 0 = no filter), order-1 resize to h_out x w_out, / scale -> d_out f32 [N,h_out,w_out,2]. */
int fisr_pwc_finish_flow(fisr_pwc* pwc, const float* d_flow, int N, int Hp, int Wp, int h0, int w0, int h_out, int w_out, const double* h_wy,
                         int ry, const double* h_wx, int rx, double scale, float* d_out, void* stream);
/* refined flow of pyramid level lvl (2..6) of the last forward, [N,H/2^lvl,W/2^lvl,2] (test hook) */
int fisr_pwc_debug_flow(fisr_pwc* pwc, int lvl, float* h_dst, size_t count);
long long fisr_pwc_launch_count(const fisr_pwc* pwc);

#ifdef __cplusplus
}
#endif
#endif /* FISR_B200_H */
