"""``FISR_for_video`` pre-processing entry points (reference ``FISR_tfoptflow/FISR_for_video_*.py``) on the B200 path."""
from __future__ import annotations

import glob
import os

import numpy as np
from PIL import Image

from .engine import Engine
from .utils import read_flo_file_5dim


def FISR_for_video_Warp_Img(args, flow_file_name, engine: Engine = None):
    """``FISR_for_video_Warp_Img`` (FISR_for_video_warp_img_with_flo.py:97-151): half-flow backward warp of every
    adjacent frame pair with the CUDA warp kernel (YUV->RGB, cv2.remap arithmetic, RGB->YUV fused; ONE launch for the clip),
    written like the reference as the MATLAB v7.3 file ``<folder>/<name>_ss1_fr<N>_warp.mat`` (dataset ``pred``: the
    [N-1, 2, h, w, 3] float32 array, 0..255, stored transposed as ``hdf5storage`` does; fisr_b200/hdf5_min.py).  Frames are taken
    in sorted order (see FISRnet.FISR_for_video)."""
    engine = engine or Engine(0)
    num_fr = args.frame_num
    data_list = sorted(glob.glob(os.path.join(args.frame_folder_path, '*.png')))
    h, w = args.FISR_input_size[0], args.FISR_input_size[1]
    import torch
    flow = read_flo_file_5dim(flow_file_name)                                   # [N-1, 2, h, w, 2]
    frames = np.stack([np.ascontiguousarray(np.array(Image.open(data_list[fr]))[:h, :w], dtype=np.uint8) for fr in range(num_fr)])
    # job 2*fr = frame fr+1 sampled along 0.5 * flow(fr -> fr+1)  (:121-124); job 2*fr+1 = frame fr along 0.5 * flow(fr+1 -> fr)  (:125-128)
    src = [fr + 1 - (j & 1) for fr in range(num_fr - 1) for j in range(2)]
    dev = "cuda:%d" % engine.device
    out = engine.warp_batch(torch.from_numpy(frames).to(dev),
                            torch.from_numpy(np.ascontiguousarray(flow[:num_fr - 1, :, :h, :w], dtype=np.float32).reshape(-1, h, w, 2)).to(dev),
                            src, 0.5, 1.0)                                      # ONE launch for the whole clip
    pred = out.cpu().numpy().reshape(num_fr - 1, 2, h, w, 3)
    for fr in range(num_fr - 1):
        print("Processing for warping imgs [%5d/%5d]" % (fr + 1, num_fr))
    folder = args.frame_folder_path.rstrip('/')
    warp_file_name = folder + '/' + folder.split('/')[-1] + '_ss{}_fr{}_warp.mat'.format(1, num_fr)
    from .utils import write_mat_file_warp
    write_mat_file_warp(warp_file_name, pred)                                  # hdf5storage.write(..., matlab_compatible=True), :131-137
    print('[*] Warp file saved!')
    return warp_file_name


PWCNET_CKPT = './models/pwcnet-lg-6-2-multisteps-chairsthingsmix/pwcnet.ckpt-595000'      # ..predict_from_img_test.py:31


def FISR_for_video_Compute_Flow(args, pwcnet=None):
    """``FISR_for_video_Compute_Flow`` (FISR_for_video_pwcnet_predict_from_img_test.py:84-147): bidirectional PWC-Net flow of
    every adjacent frame pair -> ``<folder>/<name>_test_ss1_fr<N>.flo`` ([N-1, 2, h, w, 2]).  The network runs in
    libfisr_b200.so (``fisr_b200.pwcnet.PWCNet``); its weights come from a tfoptflow checkpoint (``args.pwcnet_ckpt_path`` or the
    reference's default path), read without TensorFlow.  Neither the checkpoint nor eight modules of the reference's PWC-Net copy
    are in the reference tree, so this row is parity-unpinned; an existing flow file is used as is."""
    from .pwcnet import PWCNet
    from .utils import write_flo_file_5dim
    folder = args.frame_folder_path.rstrip('/')
    path = folder + '/' + folder.split('/')[-1] + '_test_ss{}_fr{}.flo'.format(1, args.frame_num)
    if os.path.exists(path) and pwcnet is None:
        return path
    own = pwcnet is None
    if own:
        ckpt = getattr(args, 'pwcnet_ckpt_path', None) or PWCNET_CKPT
        if not os.path.exists(ckpt + '.index'):
            raise FileNotFoundError("PWC-Net checkpoint %s.index not found (the reference downloads it separately, README.md:102); "
                                    "pass args.pwcnet_ckpt_path or supply the flow file %s" % (ckpt, path))
        pwcnet = PWCNet(getattr(args, 'gpu', 0) or 0)
        pwcnet.load_checkpoint(ckpt)
    data_list = sorted(glob.glob(os.path.join(args.frame_folder_path, '*.png')))
    h, w = args.FISR_input_size[0], args.FISR_input_size[1]
    num_fr = args.frame_num
    pred = np.zeros((num_fr - 1, 2, h, w, 2), dtype=np.float32)

    def frames():
        # PWC-Net works on RGB: the YUV frames are converted first (:113-120; utils.YUV2RGB_matlab's arithmetic, on the device).
        # Every PNG is decoded once, a few frames ahead of the GPU, on worker threads (PIL releases the GIL while decoding).
        from concurrent.futures import ThreadPoolExecutor
        load = lambda p: np.array(Image.open(p), dtype=np.uint8)[:h, :w]
        with ThreadPoolExecutor(max_workers=2) as ex:
            ahead = [ex.submit(load, p) for p in data_list[:min(4, num_fr)]]
            for i in range(num_fr):
                if i + 4 < num_fr:
                    ahead.append(ex.submit(load, data_list[i + 4]))
                yield ahead.pop(0).result()

    for fr, flow in enumerate(pwcnet.flow_sequence_yuv(frames(), scale=2)):
        pred[fr] = flow
        print("Processing for computing flows [%5d/%5d]" % (fr + 1, num_fr))
    write_flo_file_5dim(pred, path)
    print('[*] Flow file saved!')
    if own:
        pwcnet.close()
    return path
