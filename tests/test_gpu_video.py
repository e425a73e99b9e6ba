"""-m gpu: the drop-in surface -- FISRnet(sess, args).FISR_for_video / model / save+load and the warp driver -- vs the oracle."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import fisrnet_oracle as O
from oracle import pipeline_oracle as P
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def _args(tmp, H, W, n):
    return SimpleNamespace(checkpoint_dir=str(tmp / "ckpt"), test_img_dir=str(tmp / "out"), text_dir=str(tmp / "txt"),
                           log_dir=str(tmp / "log"), exp_num=1, scale_factor=2, frame_folder_path=str(tmp / "scene"),
                           FISR_input_size=(H, W), frame_num=n, FISR_test_patch=(2, 2), test_patch=(2, 2),
                           test_input_size=(H, W))


def test_fisr_for_video_end_to_end(engine, tmp_path):
    import fisr_b200
    from fisr_b200 import utils
    from fisr_b200.video import FISR_for_video_Warp_Img
    engine.set_precision("f16x3")
    H, W, n = 128, 192, 4
    frames = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"][:n]          # real LR YUV crops [n,128,192,3]
    args = _args(tmp_path, H, W, n)
    os.makedirs(args.frame_folder_path)
    for i, f in enumerate(frames):
        Image.fromarray(f).save(os.path.join(args.frame_folder_path, f"LR_seq_{2 * i + 1}.png"))
    rng = np.random.default_rng(0)
    flow = (rng.standard_normal((n - 1, 2, H, W, 2)) * 2).astype(np.float32)
    flow_file = os.path.join(args.frame_folder_path, "scene_test_ss1_fr4.flo")
    utils.write_flo_file_5dim(flow, flow_file)
    assert np.array_equal(utils.read_flo_file_5dim(flow_file), flow)

    net = fisr_b200.FISRnet(engine, args)
    params = O.init_params(12)
    engine.set_params(params)
    net._initialized = True
    net.save_checkpoint(args.checkpoint_dir, 122000)                                     # released ckpt step (README.md:62)
    engine.set_params(O.init_params(13))                                                 # clobber, load() must restore
    warp_file = FISR_for_video_Warp_Img(args, flow_file, engine)
    net.FISR_for_video(flow_file, warp_file)
    assert net.load(args.checkpoint_dir) == (True, 122000)

    # oracle: same pipeline with cv2.remap + torch-CPU network
    warp_ref = np.stack([P.warp_pair_yuv(frames[k], frames[k + 1], flow[k, 0], flow[k, 1]) for k in range(n - 1)])
    assert warp_file.endswith("_ss1_fr4_warp.mat")                                       # the reference's file name (:130)
    assert np.abs(utils.read_mat_file_warp(warp_file, 'pred') * np.float32(255.) - warp_ref).max() < 2e-3
    out_dir = os.path.join(args.frame_folder_path, "FISR_frames")
    names = sorted(os.listdir(out_dir))
    assert len(names) == 2 * (2 * n - 3)                                                 # RGB + YUV for 2n-3 frames
    fl = utils.merge_seq_dim(np.concatenate((flow[0:n - 2], flow[1:n - 1]), axis=1))
    wp = utils.merge_seq_dim(np.concatenate((warp_ref[0:n - 2], warp_ref[1:n - 1]), axis=1)) / np.float32(255.)
    for fr in range(n - 2):
        img = np.concatenate([frames[fr + s] for s in range(3)], axis=2)
        ref = P.window_forward_u8(params, img, fl[fr], wp[fr].astype(np.float32), (2, 2))
        for s in range(3):
            got = np.array(Image.open(os.path.join(out_dir, f"pred_YUV_{fr * 2 + s}.png")))
            if fr + 1 < n - 2 and s == 2:
                continue                                                                # overwritten by the next window (:1066)
            d = np.abs(got.astype(int) - ref[:, :, 3 * s:3 * s + 3].astype(int))
            assert d.max() <= 1 and (d == 0).mean() > 0.999
            rgb = np.array(Image.open(os.path.join(out_dir, f"pred_{fr * 2 + s}.png")))
            assert np.array_equal(rgb, P.yuv2rgb_matlab(got).astype("uint8"))


def test_model_method_numpy_and_torch(engine, tmp_path):
    import fisr_b200
    engine.set_precision("f16x3")
    net = fisr_b200.FISRnet(engine, _args(tmp_path, 64, 64, 3))
    params = O.init_params(14)
    engine.set_params(params)
    net._initialized = True
    x = O.synthetic_input(1, 32, 64, 5)
    ref = O.model(params, x)
    a = net.model(x.numpy(), 2, reuse=False, scope="FISRnet")
    b = net.model(x.cuda(), 2, reuse=True, scope="FISRnet")
    for r, u, v in zip(ref, a, b):
        assert isinstance(u, np.ndarray) and v.is_cuda
        assert np.abs(u - r.numpy()).max() < 1e-4 and (v.cpu() - r).abs().max() < 1e-4
    with pytest.raises(ValueError):
        net.model(x.numpy(), 4)
    assert net.model_dir == "FISRnet_exp1"
    assert net.load(str(tmp_path / "nowhere")) == (False, 0)                            # missing ckpt is not an error (:1113)


def test_cli_fisr_for_video_whole_pipeline(tmp_path):
    """``python -m fisr_b200.main --phase FISR_for_video`` on a 4-frame clip, everything from files like the reference's main.py:206-236:
    PWC-Net weights from a TensorFlow V2 bundle (tfoptflow names), FISRnet weights from a Saver bundle, flow -> ``.flo``, warp ->
    MATLAB v7.3 ``_warp.mat``, network -> PNGs; checked against the same chain on the oracles (PWC-Net oracle, cv2.remap, FISRnet oracle)."""
    from fisr_b200 import main as cli
    from fisr_b200 import tf_checkpoint as T
    from fisr_b200 import utils
    from oracle import pwcnet_oracle as W
    H, Wd, n = 64, 96, 4
    frames = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"][:n, :H, :Wd]
    scene = tmp_path / "scene9"
    os.makedirs(scene)
    for i, f in enumerate(frames):
        Image.fromarray(np.ascontiguousarray(f)).save(str(scene / f"LR_seq_{i}.png"))
    fisr_params, pwc_params = O.init_params(21), W.init_params(22)
    ck = tmp_path / "ckpt" / "FISRnet_exp1"
    T.save_fisrnet_checkpoint(str(ck / "FISRnet-122000"), {k: v.numpy() for k, v in fisr_params.items()}, None, 122000)
    (ck / "checkpoint").write_text('model_checkpoint_path: "FISRnet-122000"\n')
    pwc_prefix = str(tmp_path / "pwc" / "pwcnet.ckpt-595000")
    T.save_checkpoint(pwc_prefix, {k: v.numpy() for k, v in pwc_params.items()})
    cli.main(["--phase", "FISR_for_video", "--frame_folder_path", str(scene), "--frame_num", str(n), "--FISR_input_size", f"{H},{Wd}",
              "--FISR_test_patch", "1,1", "--checkpoint_dir", str(tmp_path / "ckpt"), "--test_img_dir", str(tmp_path / "img"),
              "--text_dir", str(tmp_path / "txt"), "--log_dir", str(tmp_path / "log"), "--pwcnet_ckpt_path", pwc_prefix])
    # ---- the same chain on the oracles
    rgb = [utils.YUV2RGB_matlab(f.astype(np.float32)) for f in frames]
    flow_ref = np.zeros((n - 1, 2, H, Wd, 2), np.float32)
    for fr in range(n - 1):
        a, b, hw0 = W.prepare_pair(rgb[fr], rgb[fr + 1])
        ff = W.forward(pwc_params, torch.from_numpy(np.stack([a, b])), torch.from_numpy(np.stack([b, a]))).numpy()
        flow_ref[fr] = np.stack([W.finish_flow(ff[k], hw0, (H, Wd)) for k in range(2)])
    flow = utils.read_flo_file_5dim(str(scene / "scene9_test_ss1_fr4.flo"))
    assert np.abs(flow - flow_ref).max() < 2e-4 * max(1.0, np.abs(flow_ref).max())
    warp_ref = np.stack([P.warp_pair_yuv(frames[k], frames[k + 1], flow[k, 0], flow[k, 1]) for k in range(n - 1)])
    warp = utils.read_mat_file_warp(str(scene / "scene9_ss1_fr4_warp.mat"), 'pred')
    assert np.abs(warp * np.float32(255.) - warp_ref).max() < 2e-3
    fl = utils.merge_seq_dim(np.concatenate((flow[0:n - 2], flow[1:n - 1]), axis=1))
    wp = utils.merge_seq_dim(np.concatenate((warp[0:n - 2], warp[1:n - 1]), axis=1))
    out_dir = scene / "FISR_frames"
    assert len(os.listdir(out_dir)) == 2 * (2 * n - 3)
    for fr in range(n - 2):
        img = np.concatenate([frames[fr + s] for s in range(3)], axis=2)
        ref = P.window_forward_u8(fisr_params, img, fl[fr], wp[fr].astype(np.float32), (1, 1))
        got = np.array(Image.open(str(out_dir / f"pred_YUV_{fr * 2}.png")))
        d = np.abs(got.astype(int) - ref[:, :, 0:3].astype(int))
        assert d.max() <= 1 and (d == 0).mean() > 0.999
