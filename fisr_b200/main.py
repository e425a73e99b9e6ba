"""Command line of the reference (``main.py:23-121``) on top of the B200 ``FISRnet``: same flags, same defaults.

    python -m fisr_b200.main --phase FISR_for_video --frame_folder_path <folder> --frame_num 5
"""
from __future__ import annotations

import argparse
import os

from .FISRnet import FISRnet
from .utils import check_folder
from .video import FISR_for_video_Compute_Flow, FISR_for_video_Warp_Img


def _pair(s):
    """The reference declares these flags ``type=tuple`` (main.py:89-103), which turns '2,2' into a tuple of characters;
    only the defaults are usable there.  Here 'H,W' parses to (H, W)."""
    if isinstance(s, tuple):
        return s
    a, b = s.replace('(', '').replace(')', '').split(',')
    return int(a), int(b)


def parse_args(argv=None):
    desc = "FISR: Deep Joint Frame Interpolation and Super-Resolution with A Multi-scale Temporal Loss (B200 path)"
    p = argparse.ArgumentParser(description=desc)
    p.add_argument('--net_type', type=str, default='FISRnet', choices=['FISRnet'])
    p.add_argument('--fraction_gpu', type=float, default=1.0)
    p.add_argument('--phase', type=str, default='FISR_for_video', choices=['train', 'test', 'FISR_for_video'])
    p.add_argument('--scale_factor', type=int, default=2)
    p.add_argument('--train_data_path', type=str, default='./data/train/LR_LFR/LR_Surfing_SlamDunk_5seq.mat')
    p.add_argument('--train_flow_data_path', type=str, default='./data/train/flow/LR_Surfing_SlamDunk_5seq_ss1.flo')
    p.add_argument('--train_flow_ss2_data_path', type=str, default='./data/train/flow/LR_Surfing_SlamDunk_5seq_ss2.flo')
    p.add_argument('--train_warped_data_path', type=str, default='./data/train/warped/LR_Surfing_SlamDunk_5seq_ss1_warp.mat')
    p.add_argument('--train_wapred_ss2_data_path', type=str, default='./data/train/warped/LR_Surfing_SlamDunk_5seq_ss2_warp.mat')
    p.add_argument('--train_label_path', type=str, default='./data/train/HR_HFR/HR_Surfing_SlamDunk_5seq.mat')
    p.add_argument('--test_data_path', type=str, default='./data/test/LR_LFR')
    p.add_argument('--test_flow_data_path', type=str, default='./data/test/flow/LR_Surfing_SlamDunk_test_ss1.flo')
    p.add_argument('--test_warped_data_path', type=str, default='./data/test/warped/LR_Surfing_SlamDunk_test_ss1_warp.mat')
    p.add_argument('--test_label_path', type=str, default='./data/test/HR_HFR')
    p.add_argument('--test_img_dir', type=str, default='./test_img_dir')
    p.add_argument('--text_dir', type=str, default='./text_dir')
    p.add_argument('--checkpoint_dir', type=str, default='./checkpoint_dir')
    p.add_argument('--log_dir', type=str, default='./logdir')
    p.add_argument('--exp_num', type=int, default=1)
    p.add_argument('--epoch', type=int, default=100)
    p.add_argument('--freq_display', type=int, default=100)
    p.add_argument('--init_lr', type=float, default=0.0001)
    p.add_argument('--lr_type', type=str, default='stair_decay', choices=['linear_decay', 'stair_decay', 'no_decay'])
    p.add_argument('--lr_stair_decay_points', type=int, nargs='+', default=[80, 90])
    p.add_argument('--lr_decreasing_factor', type=float, default=0.1)
    p.add_argument('--lr_linear_decay_point', type=int, default=50)
    p.add_argument('--batch_size', type=int, default=8)
    p.add_argument('--n_train_img_showed', type=int, default=3)
    p.add_argument('--val_batch_size', type=int, default=2)
    p.add_argument('--val_data_size', type=int, default=320)
    p.add_argument('--recn_lambda', type=float, default=1.0)
    p.add_argument('--tm1_lambda', type=float, default=1.0)
    p.add_argument('--tm2_lambda', type=float, default=0.1)
    p.add_argument('--tmm_lambda', type=float, default=1.0)
    p.add_argument('--td_lambda', type=float, default=0.1)
    p.add_argument('--ss2_lambda', type=float, default=1.0)
    p.add_argument('--test_patch', type=_pair, default=(2, 2))
    p.add_argument('--test_input_size', type=_pair, default=(1080, 1920))
    p.add_argument('--frame_folder_path', type=str, default='./FISR_test_folder/scene1')
    p.add_argument('--FISR_input_size', type=_pair, default=(1080, 1920))
    p.add_argument('--frame_num', type=int, default=5)
    p.add_argument('--FISR_test_patch', type=_pair, default=(2, 2))
    p.add_argument('--precision', type=str, default='f16x3', choices=['f16x3', 'f16', 'f16f8'],
                   help='B200 path only: f16x3 = fp32-class split operands (default; training always uses it), f16f8 = fp16 main '
                        'term + fp8 cross terms (inference, ~3e-5 max-abs, 1.2x faster), f16 = single-fp16 fast mode')
    p.add_argument('--device', type=int, default=0, help='B200 path only: CUDA device index')
    p.add_argument('--pwcnet_ckpt_path', type=str, default=None,
                   help='FISR_for_video: tfoptflow PWC-Net checkpoint prefix (default: the path hard-wired in the reference, '
                        'FISR_for_video_pwcnet_predict_from_img_test.py:31); not needed when the .flo file already exists')
    args = p.parse_args(argv)
    for d in (args.checkpoint_dir, args.text_dir, args.log_dir, args.test_img_dir):         # main.py:108-121
        check_folder(d)
    return args


def main(argv=None):
    args = parse_args(argv)
    net = FISRnet(args.device, args)
    if args.phase == 'train':
        with open(args.text_dir + '/exp_' + str(args.exp_num) + '.txt', 'a') as log:        # main.py:131-134
            log.write('----- Model parameters -----\n')
            for arg in vars(args):
                log.write('{} : {}\n'.format(arg, getattr(args, arg)))
        print("[*] Exp: ", args.exp_num)
        net.build_model()
        print("[*] Training starts")
        net.train()
        print("[*] Training finished! ")
        if args.test_data_path and os.path.isdir(args.test_data_path):                        # "test after training", main.py:159-181
            print("[*] Testing starts")
            net.test()
            print("[*] Testing finished! ")
    elif args.phase == 'test':
        net.test()
    else:                                                                                    # main.py:206-236
        flow_file_name = FISR_for_video_Compute_Flow(args)
        warp_file_name = FISR_for_video_Warp_Img(args, flow_file_name, net.engine)
        net.FISR_for_video(flow_file_name, warp_file_name)
        print(" [*] FISR_for_video finished!")


if __name__ == '__main__':
    main()
