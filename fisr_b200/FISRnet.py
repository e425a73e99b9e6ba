"""``FISRnet`` -- the operator surface ``main.py`` drives (reference FISRnet.py:14-17), on the B200 path.

Same constructor, attribute names, methods and argument meaning as the reference class, so that
``main.py``'s ``test`` and ``FISR_for_video`` phases run unchanged after swapping the import
(INTEGRATION.md).  ``sess`` was a ``tf.Session``; here it is ``None``, a CUDA device index or a
``fisr_b200.Engine``.  All arithmetic runs in libfisr_b200.so; this file only moves data.

Documented deviations from the reference:
  * ``FISR_for_video`` sorts the frame list (the reference's unsorted ``glob`` at FISRnet.py:953 returns a
    file-system-dependent order; the shipped scene1 outputs correspond to the sorted order, like ``test()`` :762).
  * the tile loop runs as one batched forward instead of growing a TF graph per tile (FISRnet.py:1039-1041).
  * checkpoints are TensorFlow V2 bundles (``FISRnet-<step>.index`` / ``.data-00000-of-00001``) written by
    ``fisr_b200.tf_checkpoint`` without TensorFlow, under the variable names the reference's Saver uses: the 276 weights,
    the Adam slots ``<var>/Adam`` / ``<var>/Adam_1``, ``beta1_power`` / ``beta2_power`` and the global step, so a resumed
    run continues seamlessly and the files go back to the reference.  ``.npz`` checkpoints of round 1 still load.
  * ``train`` prints what the reference prints; the TensorBoard summaries (FISRnet.py:533-578) are not written.
  * the printed "Estimated Inference Time" is per window from CUDA-synchronised wall time.
"""
from __future__ import annotations

import glob
import math
import os
import re
import time
from datetime import datetime

import numpy as np
from PIL import Image

from . import utils
from .engine import Engine, param_inventory
from .init import xavier_params
from .utils import check_folder, merge_seq_dim, read_flo_file_5dim, read_mat_file_warp



class _FrameWriter:
    """PNG output off the critical path: colour conversion + PNG compression of a 4K frame take ~1 s on one core against 25 ms of GPU
    work per window, so they run on a thread pool (PIL and numpy release the GIL) while the next windows are computed.  The files
    are the ones the sequential loop of the reference writes (FISRnet.py:888-904, 1066-1077); at most ``max_pending`` frames wait."""

    def __init__(self, workers=None, max_pending=12):
        import threading
        from concurrent.futures import ThreadPoolExecutor
        self.pool = ThreadPoolExecutor(max_workers=workers or min(8, os.cpu_count() or 1))
        self.slots = threading.Semaphore(max_pending)
        self.futures = []

    def _job(self, yuv, rgb_path, yuv_path):
        try:
            Image.fromarray(utils.YUV2RGB_matlab(yuv).astype('uint8')).save(rgb_path)
            if yuv_path:
                Image.fromarray(yuv.astype('uint8')).save(yuv_path)
        finally:
            self.slots.release()

    def save(self, yuv, rgb_path, yuv_path=None):
        """``yuv``: uint8 [H,W,3]; writes its RGB conversion to ``rgb_path`` and, if given, the YUV frame itself to ``yuv_path``."""
        self.slots.acquire()
        self.futures.append(self.pool.submit(self._job, yuv, rgb_path, yuv_path))

    def close(self):
        self.pool.shutdown(wait=True)
        for f in self.futures:
            f.result()                      # re-raise the first failure
        self.futures = []


class FISRnet(object):
    model_name = "FISRnet"

    def __init__(self, sess, args):
        self.sess = sess
        self.args = args
        for name in ("checkpoint_dir", "test_img_dir", "text_dir", "log_dir", "train_data_path", "train_flow_data_path",
                     "train_flow_ss2_data_path", "train_warped_data_path", "train_wapred_ss2_data_path", "train_label_path",
                     "test_data_path", "test_flow_data_path", "test_warped_data_path", "test_label_path", "exp_num",
                     "scale_factor", "epoch", "init_lr", "freq_display", "lr_type", "lr_stair_decay_points",
                     "lr_decreasing_factor", "lr_linear_decay_point", "batch_size", "val_batch_size", "val_data_size",
                     "n_train_img_showed", "recn_lambda", "tm1_lambda", "tm2_lambda", "tmm_lambda", "td_lambda",
                     "ss2_lambda", "test_patch", "test_input_size", "FISR_test_patch", "frame_folder_path",
                     "FISR_input_size", "frame_num"):                       # FISRnet.py:20-66
            setattr(self, name, getattr(args, name, None))
        if isinstance(sess, Engine):
            self.engine = sess
        else:
            self.engine = Engine(int(sess) if isinstance(sess, int) else 0,
                                 precision=getattr(args, "precision", "f16x3"))
        self._initialized = False
        print('Model arguments, [{:s}]'.format((str(datetime.now())[:-7])))
        for arg in vars(args):
            print('# {} : {}'.format(arg, getattr(args, arg)))

    # ------------------------------------------------------------------ variables
    def _ensure_variables(self):
        """``tf.global_variables_initializer().run()`` (FISRnet.py:757,948): Xavier weights, zero biases."""
        if not self._initialized:
            self.engine.set_params(xavier_params(seed=0, bias_std=0.0))
            self._initialized = True

    # ------------------------------------------------------------------ FISRnet.model (FISRnet.py:73-173)
    def model(self, img, sf, reuse=False, scope="model"):
        """img [N,H,W,29] (numpy array or CUDA torch tensor) -> (pred_l1, pred_l2, pred_l3) in the same container."""
        if int(sf) != 2:
            raise ValueError("FISRnet is a x2 network (main.py:29 default scale_factor=2)")
        self._ensure_variables()
        if isinstance(img, np.ndarray):
            return self.engine.forward_host(img)
        return self.engine.forward(img)

    # ------------------------------------------------------------------ training (FISRnet.py:175-743)
    def build_model(self):
        """Reads the training set and fixes the schedule (FISRnet.py:175-248).  The graph itself -- four weight-shared
        forwards, multi-scale temporal loss, backward, Adam (FISRnet.py:250-491) -- is ``fisr_train_step`` in the library."""
        print(" Start to read 4K data.")
        data, label = utils.read_mat_file(self.train_data_path, self.train_label_path, 'LR_data', 'HR_data')  # [B,N_seq,H,W,C]
        print(" Successfully load.")
        data, label = merge_seq_dim(data), merge_seq_dim(label)          # [B,h,w,15], [B,2h,2w,21]
        self.data_sz, self.label_sz = data.shape, label.shape
        print(" Start to read flow data.")
        flow = merge_seq_dim(read_flo_file_5dim(self.train_flow_data_path)) / self.data_sz[1] / 2          # FISRnet.py:195-197
        flow_ss2 = merge_seq_dim(read_flo_file_5dim(self.train_flow_ss2_data_path)) / self.data_sz[1] / 2  # :199-202
        print(" Successfully load.")
        print(" Start to read warped data.")
        warp = merge_seq_dim(read_mat_file_warp(self.train_warped_data_path, 'pred'))
        warp_ss2 = merge_seq_dim(read_mat_file_warp(self.train_wapred_ss2_data_path, 'pred'))
        print(" Successfully load.")
        v = self.val_data_size                                           # split val / train, FISRnet.py:213-227
        as32 = lambda a: np.ascontiguousarray(a, dtype=np.float32)
        self.data_val, self.label_val, self.flow_val, self.warp_val = (as32(a[-v:]) for a in (data, label, flow, warp))
        self.flow_ss2_val, self.warp_ss2_val = as32(flow_ss2[-v:]), as32(warp_ss2[-v:])
        self.data, self.label, self.flow, self.flow_ss2, self.warp, self.warp_ss2 = (
            as32(a[:-v]) for a in (data, label, flow, flow_ss2, warp, warp_ss2))
        self.train_iter = math.floor((self.data_sz[0] - v) / self.batch_size)
        self.val_iter = math.floor(v / self.val_batch_size)
        self.global_step = 0
        if self.lr_type == "stair_decay":                                # tf.train.piecewise_constant, FISRnet.py:233-240
            self.epoch_lr_to_be_decayed_boundaries = [y * self.train_iter for y in self.lr_stair_decay_points]
            self.epoch_lr_to_be_decayed_value = [self.init_lr * (self.lr_decreasing_factor ** y)
                                                 for y in range(len(self.lr_stair_decay_points) + 1)]
            print("lr_type: stair_decay")
        elif self.lr_type == "linear_decay":
            print("lr_type: linear_decay")
        else:
            print("lr_type: no decay")
        self.lambdas = {"recn": self.recn_lambda, "tm1": self.tm1_lambda, "tm2": self.tm2_lambda, "tmm": self.tmm_lambda,
                        "td": self.td_lambda, "ss2": self.ss2_lambda}    # main.py:80-85
        self._built = True

    def _lr(self, epoch):
        """Learning rate of the step about to run: piecewise constant in global_step (x <= boundary keeps the earlier value,
        like tf.train.piecewise_constant), linear decay in the epoch (FISRnet.py:631-633), or constant."""
        if self.lr_type == "stair_decay":
            k = sum(1 for b in self.epoch_lr_to_be_decayed_boundaries if self.global_step > b)
            return self.epoch_lr_to_be_decayed_value[k]
        if self.lr_type == "linear_decay" and epoch >= self.lr_linear_decay_point:
            return self.init_lr * (self.epoch - epoch) / (self.epoch - self.lr_linear_decay_point)
        return self.init_lr

    def validate(self, data, label, flow, warp):
        """Validation graph of FISRnet.py:492-531 on one batch: three stride-1 windows through the shared network,
        ``Groups2Ovlp``, then L2 and ``tf.image.psnr`` (per image, then mean) against the 7 GT frames."""
        import torch
        dev = "cuda:%d" % self.engine.device
        d, f, w = (torch.as_tensor(a, dtype=torch.float32, device=dev) for a in (data, flow, warp))
        gt = torch.as_tensor(label, dtype=torch.float32, device=dev)
        B, H2, W2, _ = gt.shape
        # window i: frames i..i+2, flows 4i..4i+8, warps 6i..6i+12 (ops.py:90-116), the three windows as one batch
        x = torch.cat([torch.cat((d[..., 3 * i:3 * i + 9], f[..., 4 * i:4 * i + 8], w[..., 6 * i:6 * i + 12]), dim=3)
                       for i in range(3)], dim=0)
        pred = self.engine.forward(x, want=(False, False, True))[2]      # [3B, 2h, 2w, 9], windows 0, 1, 2 stacked
        seq = self.engine.groups2ovlp(pred)                              # [B, 7, 2h, 2w, 3]
        gt = gt.reshape(B, H2, W2, 7, 3).permute(0, 3, 1, 2, 4)           # tf_split_seq_dim
        err = (seq - gt) ** 2
        recn = float(err.mean())                                         # L2_loss, ops.py:30-32
        mse = err.mean(dim=(2, 3, 4))                                    # tf.image.psnr: last three dims of [B, 7, H, W, 3]
        return recn, float((-10.0 * torch.log10(mse)).mean())

    def train(self):
        if not getattr(self, "_built", False):
            self.build_model()
        self._ensure_variables()                                         # tf.global_variables_initializer().run()
        self.engine.set_precision("f16x3")                               # the backward kernels need the split operand planes
        could_load, checkpoint_counter = self.load(self.checkpoint_dir)  # FISRnet.py:593-604
        if could_load:
            start_epoch = int(checkpoint_counter / self.train_iter)
            counter = checkpoint_counter
            self.global_step = checkpoint_counter                        # lr schedule (tf.train.piecewise_constant)
            if not self._adam_restored:
                # a checkpoint without optimizer slots: zero moments need the bias correction of step 0 (with t ~ 1e5 the
                # first updates would be ~3 lr sign(g), perturbing a trained model)
                self.engine.adam_reset(0)
            print(" [*] Load SUCCESS")
        else:
            start_epoch, counter = 0, 1
            self.engine.adam_reset(0)
            print(" [!] Load failed...")
        import torch
        dev = "cuda:%d" % self.engine.device
        names = self.engine.LOSS_NAMES
        start_time = time.time()
        for epoch in range(start_epoch, self.epoch):
            hist = {k: [] for k in names}
            rand_idx = np.random.permutation(self.data_sz[0] - self.val_data_size)      # FISRnet.py:622
            lr = self._lr(epoch)
            for idx in range(self.train_iter):
                sel = rand_idx[self.batch_size * idx:self.batch_size * (idx + 1)]
                lr = self._lr(epoch)
                batch = [torch.from_numpy(a[sel]).to(dev, non_blocking=True)            # the feed_dict of FISRnet.py:634-639
                         for a in (self.data, self.flow, self.flow_ss2, self.warp, self.warp_ss2, self.label)]
                s = self.engine.train_step(*batch, lr, self.lambdas)
                self.global_step += 1
                if np.mod(idx, self.freq_display) == 0:
                    print("Epoch: [%3d], [%4d/%4d]-th batch, time: %4.2f(min.), "
                          "train_PSNR: %.3f, recnLoss: %.6f, tmLoss: %.6f, tmmLoss: %.6f, tdLoss: %.6f, "
                          "totalLoss_s1: %.6f,recnLoss_ss2: %.6f,"
                          "tdLoss_ss2: %.6f, tmLoss_ss2: %.6f, totalLoss_ss2: %.6f, total_loss: %.6f"
                          % (epoch, idx, self.train_iter, (time.time() - start_time) / 60, s["train_PSNR"], s["recnLoss"],
                             s["tmLoss"], s["tmmLoss"], s["tdLoss"], s["totalLoss_s1"], s["recnLoss_ss2"], s["tdLoss_ss2"],
                             s["tmLoss_ss2"], s["totalLoss_ss2"], s["total_loss"]))
                counter += 1
                for k in names:
                    hist[k].append(s[k])
            m = {k: float(np.mean(hist[k])) if hist[k] else float("nan") for k in names}
            print("# (average) Epoch: [%4d], LR: %1.10f, time: %4.2f(minutes), "
                  "train_PSNR: %.3f, recnLoss: %.6f, tmLoss: %.6f, tmmLoss: %.6f, tdLoss: %.6f, "
                  "totalLoss_s1: %.6f,recnLoss_ss2: %.6f,"
                  "tdLoss_ss2: %.6f, tmLoss_ss2: %.6f, totalLoss_ss2: %.6f, total_loss: %.6f"
                  % (epoch, lr, (time.time() - start_time) / 60, m["train_PSNR"], m["recnLoss"], m["tmLoss"], m["tmmLoss"],
                     m["tdLoss"], m["totalLoss_s1"], m["recnLoss_ss2"], m["tdLoss_ss2"], m["tmLoss_ss2"], m["totalLoss_ss2"],
                     m["total_loss"]))
            val_recn, val_psnr = [], []                                  # FISRnet.py:706-722
            for val_idx in range(self.val_iter):
                sl = slice(self.val_batch_size * val_idx, self.val_batch_size * (val_idx + 1))
                r, p_ = self.validate(self.data_val[sl], self.label_val[sl], self.flow_val[sl], self.warp_val[sl])
                val_recn.append(r)
                val_psnr.append(p_)
            print("######### Validation (average),Epoch: [%4d/%4d]-th epoch, time: %4.2f(min.), val_PSNR: %.3f[dB], "
                  "recnLoss: %.6f #########"
                  % (epoch, self.epoch, (time.time() - start_time) / 60,
                     float(np.mean(val_psnr)) if val_psnr else float("nan"), float(np.mean(val_recn)) if val_recn else float("nan")))
            self.save_checkpoint(self.checkpoint_dir, self.global_step)  # FISRnet.py:737
        self.save_checkpoint(self.checkpoint_dir, self.global_step)      # FISRnet.py:743

    # ------------------------------------------------------------------ test (FISRnet.py:746-935)
    def test(self):
        self._ensure_variables()
        _, _ = self.load(self.checkpoint_dir)
        test_data_path = sorted(glob.glob(os.path.join(self.test_data_path, '*.png')))
        test_label_path = sorted(glob.glob(os.path.join(self.test_label_path, '*.png')))
        print(" Start to read flow data (test).")
        flow = merge_seq_dim(read_flo_file_5dim(self.test_flow_data_path))
        print(" Start to read warped data (test).")
        warp = merge_seq_dim(read_mat_file_warp(self.test_warped_data_path, 'pred'))
        num_patch = self.test_patch
        test_img_dir = check_folder(os.path.join(self.test_img_dir, self.model_dir))
        n_in_seq, n_test_in_seq = 3, 5
        n_GT_seq, n_test_label_seq = n_in_seq * 2 - 3, 2 * n_test_in_seq - 3
        psnr_fisr, psnr_sr, ssim_fisr, ssim_sr, inf_time = [], [], [], [], []
        writer = _FrameWriter()
        start_time = time.time()
        H, W = self.test_input_size
        h = H - np.remainder(H, 32 * num_patch[0])
        w = W - np.remainder(W, 32 * num_patch[1])
        for scene_i in range(int(len(test_data_path) / n_test_in_seq)):
            for sample_i in range(n_test_in_seq - n_in_seq + 1):
                img = np.concatenate([np.array(Image.open(test_data_path[scene_i * n_test_in_seq + sample_i + s]))
                                      for s in range(n_in_seq)], axis=2)
                label = np.concatenate([np.array(Image.open(test_label_path[scene_i * n_test_label_seq + sample_i * 2 + s]))
                                        for s in range(n_GT_seq)], axis=2)
                label = np.clip(np.array(label[:h * 2, :w * 2, :], dtype=np.double) / 255., 0, 1)
                flow_sample = flow[scene_i, :, :, 4 * sample_i:4 * sample_i + 8]
                warp_sample = warp[scene_i, :, :, 6 * sample_i:6 * sample_i + 12]
                # frames, flow and warp cropped to (h, w) like the reference (FISRnet.py:826-840): files stored at the
                # cropped size work too
                t0 = time.time()
                full = self.engine.window_host_f32(img[:h, :w], flow_sample[:h, :w], warp_sample[:h, :w],
                                                   tuple(int(v) for v in num_patch))          # test_Pred_full, float
                inf_time.append(time.time() - t0)
                # the reference scores the CLIPPED FLOAT prediction (FISRnet.py:883-887) and truncates to uint8 only for the
                # PNGs and SSIM (:888,901): scoring the uint8 canvas would cost ~(1/255)^2/3 of MSE (-1.2 dB at 48 dB)
                test_pred = np.clip(full.astype(np.double), 0, 1)
                pred = np.uint8(test_pred * 255)
                test_PSNR = [utils._compute_psnr(test_pred[:, :, 3 * s:3 * (s + 1)], label[:, :, 3 * s:3 * (s + 1)], 1.)
                             for s in range(n_GT_seq)]
                print(" <Test> [%4d/%4d]-th image, scene: %2d-%d, time: %4.4f(minutes), test_PSNR: fr1 (FI-SR) %.8f[dB], "
                      "fr2 (SR) %.8f[dB], fr3 (FI-SR) %.8f[dB]  "
                      % (scene_i * 3 + sample_i, len(test_data_path) / n_test_in_seq * 3, scene_i, sample_i,
                         (time.time() - start_time) / 60, test_PSNR[0], test_PSNR[1], test_PSNR[2]))
                label_u8 = (label * 255).astype('uint8')                                       # FISRnet.py:890-891
                test_SSIM = [utils.compare_ssim(pred[:, :, 3 * s:3 * (s + 1)], label_u8[:, :, 3 * s:3 * (s + 1)])
                             for s in range(n_GT_seq)]
                print(" --------------------------------------------------------------- test_SSIM: fr1 (FI-SR) %.8f, "
                      "fr2 (SR) %.8f, fr3 (FI-SR) %.8f  " % (test_SSIM[0], test_SSIM[1], test_SSIM[2]))
                for s in range(n_GT_seq):
                    if s == 2 and sample_i < n_test_in_seq - n_in_seq:
                        continue            # the next sample's frame 0 has the same file name and is written later (FISRnet.py:901-904)
                    fr_name = os.path.basename(test_label_path[scene_i * n_test_label_seq + sample_i * 2 + s])[3:]
                    writer.save(pred[:, :, s * 3:(s + 1) * 3], os.path.join(test_img_dir, 'pred_{}'.format(fr_name)))
                psnr_fisr.append(test_PSNR[0])
                psnr_sr.append(test_PSNR[1])
                ssim_fisr.append(test_SSIM[0])
                ssim_sr.append(test_SSIM[1])
                if sample_i == 2:
                    psnr_fisr.append(test_PSNR[2])
                    ssim_fisr.append(test_SSIM[2])
        print("######### Test (average) test_PSNR: FISR %.8f[dB], SR %.8f[dB]  #########"
              % (np.mean(psnr_fisr), np.mean(psnr_sr)))
        print("######### Test (average) test_SSIM: FISR %.8f, SR %.8f #########" % (np.mean(ssim_fisr), np.mean(ssim_sr)))
        writer.close()
        print("######### Estimated Inference Time (per window = three 4K frames): %.8f[s]  #########" % np.mean(inf_time))

    # ------------------------------------------------------------------ FISR_for_video (FISRnet.py:937-1084)
    def FISR_for_video(self, flow_file_name, warp_file_name):
        self._ensure_variables()
        _, _ = self.load(self.checkpoint_dir)
        test_data_path = sorted(glob.glob(os.path.join(self.frame_folder_path, '*.png')))      # YUV
        num_fr = self.frame_num
        FISR_img_dir = check_folder(os.path.join(self.frame_folder_path, 'FISR_frames'))
        print(" Start to read flow data (FISR test).")
        flow = read_flo_file_5dim(flow_file_name)                                               # [N-1, 2, h, w, 2]
        flow = merge_seq_dim(np.concatenate((flow[0:num_fr - 2], flow[1:num_fr - 1]), axis=1))  # FISRnet.py:965-967
        print(" Start to read warped data (FISR test).")
        warp = read_mat_file_warp(warp_file_name, 'pred')                                       # [N-1, 2, h, w, 3]
        warp = merge_seq_dim(np.concatenate((warp[0:num_fr - 2], warp[1:num_fr - 1]), axis=1))  # FISRnet.py:972-975
        num_patch = self.FISR_test_patch
        check_folder(os.path.join(self.test_img_dir, self.model_dir))
        H, W = self.FISR_input_size
        h = H - np.remainder(H, 32 * num_patch[0])                                              # FISRnet.py:1006-1007
        w = W - np.remainder(W, 32 * num_patch[1])
        inf_time = []
        start_time = time.time()
        digits = math.ceil(math.log10(2 * (num_fr - 1)))
        def windows():
            # every PNG is decoded once (consecutive windows share two of their three frames), a few frames ahead, on worker threads
            from concurrent.futures import ThreadPoolExecutor
            with ThreadPoolExecutor(max_workers=2) as ex:
                decoded = {}
                def frame(i):
                    if i not in decoded:
                        decoded[i] = ex.submit(lambda p: np.array(Image.open(p)), test_data_path[i])
                    return decoded[i]
                for fr in range(num_fr - 2):
                    for ahead in range(5):
                        if fr + ahead < num_fr:
                            frame(fr + ahead)
                    img = np.concatenate([frame(fr + s).result() for s in range(3)], axis=2)
                    decoded.pop(fr - 1, None)
                    yield img[:h, :w], flow[fr, :h, :w], warp[fr, :h, :w]      # FISRnet.py:1008-1021

        # two windows in flight: the copies of window k+1 overlap the kernels of window k (fisr_window_submit / _wait)
        writer = _FrameWriter()
        t_prev = time.time()
        for fr, pred in enumerate(self.engine.video_windows(windows(), tuple(int(v) for v in num_patch))):   # YUV uint8 [2h,2w,9]
            inf_time.append(time.time() - t_prev)
            t_prev = time.time()
            for seq_i in range(3):                                                              # FISRnet.py:1066-1077
                if seq_i == 2 and fr + 1 < num_fr - 2:
                    continue                # the next window's frame 0 has the same file name and is written later
                name = str(fr * 2 + seq_i).zfill(digits)
                writer.save(pred[:, :, seq_i * 3:(seq_i + 1) * 3], FISR_img_dir + '/pred_{}.png'.format(name),
                            FISR_img_dir + '/pred_YUV_{}.png'.format(name))
            print(" <FISR processing> [%4d/%4d]-th input multiple data sample (stride1), time: %4.4f(minutes)  "
                  % (fr + 1, num_fr - 2, (time.time() - start_time) / 60))
        writer.close()
        print("######### Estimated Inference Time (per window = three 4K frames): %.8f[s]  #########" % np.mean(inf_time))

    # ------------------------------------------------------------------ checkpoints (FISRnet.py:1086-1115)
    @property
    def model_dir(self):
        return "{}_exp{}".format(self.model_name, self.exp_num)

    def save_checkpoint(self, checkpoint_dir, step):
        """``self.saver.save(sess, <dir>/FISRnet, global_step=step)`` (FISRnet.py:1092-1099): a TensorFlow V2 bundle with the
        weights and, once training has started, the optimizer state (Adam slots, beta powers, global step)."""
        from .tf_checkpoint import save_fisrnet_checkpoint
        checkpoint_dir = os.path.join(checkpoint_dir, self.model_dir)
        os.makedirs(checkpoint_dir, exist_ok=True)
        name = "{}-{}".format(self.model_name, int(step))
        adam = self.engine.get_adam_state() if self.engine.adam_steps > 0 else None
        save_fisrnet_checkpoint(os.path.join(checkpoint_dir, name), self.engine.get_params(), adam, int(step))
        with open(os.path.join(checkpoint_dir, "checkpoint"), "w") as f:
            f.write('model_checkpoint_path: "{}"\nall_model_checkpoint_paths: "{}"\n'.format(name, name))

    def load(self, checkpoint_dir):
        print(" [*] Reading checkpoints...")
        checkpoint_dir = os.path.join(checkpoint_dir, self.model_dir)
        state = os.path.join(checkpoint_dir, "checkpoint")
        self._adam_restored = False
        ckpt_name = None
        if os.path.exists(state):
            m = re.search(r'model_checkpoint_path:\s*"([^"]+)"', open(state).read())
            if m and (os.path.exists(os.path.join(checkpoint_dir, os.path.basename(m.group(1)))) or
                      os.path.exists(os.path.join(checkpoint_dir, os.path.basename(m.group(1)) + ".index"))):
                ckpt_name = os.path.basename(m.group(1))
        if ckpt_name and ckpt_name.endswith(".npz"):
            data = np.load(os.path.join(checkpoint_dir, ckpt_name))
            params = {k.replace('.', '/'): data[k] for k in data.files}
            missing = [k for k in param_inventory() if k not in params]
            if missing:
                raise KeyError("checkpoint {} lacks {} variables, e.g. {}".format(ckpt_name, len(missing), missing[0]))
            self.engine.set_params(params)
            self._initialized = True
            counter = int(next(re.finditer(r"(\d+)(?!.*\d)", ckpt_name)).group(0))
            print(" [*] Success to read {}".format(ckpt_name))
            return True, counter
        if ckpt_name and os.path.exists(os.path.join(checkpoint_dir, ckpt_name + ".index")):
            # a tf.train.Saver V2 bundle: written by the reference itself (e.g. the released FISRnet-122000) or by save_checkpoint
            from .tf_checkpoint import fisrnet_adam_state, fisrnet_weights
            prefix = os.path.join(checkpoint_dir, ckpt_name)
            self.engine.set_params(fisrnet_weights(prefix))
            self._initialized = True
            adam = fisrnet_adam_state(prefix)
            if adam is not None:                      # training checkpoint: m, v and beta1_power^t restored like Saver.restore
                self.engine.set_adam_state(*adam)
                self._adam_restored = True
            counter = int(next(re.finditer(r"(\d+)(?!.*\d)", ckpt_name)).group(0))       # FISRnet.py:1110
            print(" [*] Success to read {}".format(ckpt_name))
            return True, counter
        print(" [*] Failed to find a checkpoint")           # like the reference, not an error (FISRnet.py:1113-1115)
        return False, 0
