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
    adjacent frame pair with the CUDA warp kernel (YUV->RGB, cv2.remap arithmetic, RGB->YUV fused), written as
    ``<folder>/<name>_ss1_fr<N>_warp.npy`` ([N-1, 2, h, w, 3] float32, 0..255 -- the array the reference stores in its
    ``.mat``).  Frames are taken in sorted order (see FISRnet.FISR_for_video)."""
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
    warp_file_name = folder + '/' + folder.split('/')[-1] + '_ss{}_fr{}_warp.npy'.format(1, num_fr)
    np.save(warp_file_name, pred)
    print('[*] Warp file saved!')
    return warp_file_name


def FISR_for_video_Compute_Flow(args):
    """PWC-Net flow estimation (FISR_for_video_pwcnet_predict_from_img_test.py:84-147) is outside this hot path: the
    reference's PWC-Net copy misses 8 un-vendored modules and its checkpoint.  Supply the 5-D
    ``.flo`` file it would have written: ``<folder>/<name>_test_ss1_fr<N>.flo`` (``utils.write_flo_file_5dim``)."""
    folder = args.frame_folder_path.rstrip('/')
    path = folder + '/' + folder.split('/')[-1] + '_test_ss{}_fr{}.flo'.format(1, args.frame_num)
    if os.path.exists(path):
        return path
    raise NotImplementedError("PWC-Net is not part of the B200 hot path; expected a precomputed flow file at " + path)
