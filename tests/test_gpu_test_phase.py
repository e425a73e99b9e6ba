"""-m gpu: ``FISRnet.test()`` (reference FISRnet.py:746-935) on a one-scene synthetic test set in the reference's on-disk
layout: the printed PSNR must be the reference's -- scored on the CLIPPED FLOAT prediction (FISRnet.py:883-887), not on the
truncated uint8 canvas -- within 0.01 dB of the oracle-side pipeline."""
import os
import re
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import fisrnet_oracle as O
from oracle import pipeline_oracle as P
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


def test_test_phase_psnr_is_scored_on_the_float_canvas(engine, tmp_path, capsys):
    import fisr_b200
    from fisr_b200 import utils
    engine.set_precision("f16x3")
    H, W, grid = 128, 192, (2, 2)
    frames = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"][:5]           # 5 LR YUV frames [5,128,192,3]
    rng = np.random.default_rng(0)
    flow5 = (rng.standard_normal((1, 8, H, W, 2)) * 2).astype(np.float32)                  # [scene, 8, h, w, 2]
    warp5 = rng.uniform(0, 255, (1, 8, H, W, 3)).astype(np.float32)                        # [scene, 8, h, w, 3], 0..255
    params = O.init_params(17)
    flow = utils.merge_seq_dim(flow5)
    warp = utils.merge_seq_dim(warp5 / np.float32(255.))

    # oracle side: the three windows of the scene, float canvases (FISRnet.py:844-880)
    def fn(tile):
        return O.model(params, torch.from_numpy(tile.astype(np.float32)))[2].numpy().astype(np.float64)
    canv = []
    for s in range(3):
        img = np.concatenate([frames[s + k] for k in range(3)], axis=2)
        inp = P.normalise_window(img, flow[0, :, :, 4 * s:4 * s + 8], warp[0, :, :, 6 * s:6 * s + 12], H, W)
        canv.append(P.tiled_window(fn, inp, grid))
    # labels: 7 HR frames ~50 dB away from the prediction, where uint8 truncation of the prediction would cost ~1 dB
    label = []
    for j in range(7):
        s = min(j // 2, 2)
        k = j - 2 * s
        clean = np.clip(canv[s][:, :, 3 * k:3 * k + 3], 0, 1)
        label.append(np.uint8(np.clip(clean + rng.normal(0, 0.003, clean.shape), 0, 1) * 255 + 0.5))
    d = tmp_path
    os.makedirs(d / "lr"); os.makedirs(d / "hr")
    for i, f in enumerate(frames):
        Image.fromarray(f).save(str(d / "lr" / f"LR_seq_{i:02d}.png"))
    for j, f in enumerate(label):
        Image.fromarray(f).save(str(d / "hr" / f"HR_seq_{j:02d}.png"))
    utils.write_flo_file_5dim(flow5, str(d / "flow.flo"))
    np.save(str(d / "warp.npy"), warp5)
    args = SimpleNamespace(checkpoint_dir=str(d / "ckpt"), test_img_dir=str(d / "out"), text_dir=str(d / "txt"), log_dir=str(d / "log"),
                           exp_num=1, scale_factor=2, test_data_path=str(d / "lr"), test_label_path=str(d / "hr"),
                           test_flow_data_path=str(d / "flow.flo"), test_warped_data_path=str(d / "warp.npy"),
                           test_patch=grid, test_input_size=(H, W))
    net = fisr_b200.FISRnet(engine, args)
    engine.set_params(params)
    net._initialized = True
    capsys.readouterr()
    net.test()
    out = capsys.readouterr().out
    got = [[float(v) for v in m] for m in re.findall(r"fr1 \(FI-SR\) ([\d.]+)\[dB\], fr2 \(SR\) ([\d.]+)\[dB\], fr3 \(FI-SR\) ([\d.]+)\[dB\]", out)]
    assert len(got) == 3
    worst, cost = 0.0, []
    for s in range(3):
        gt = np.concatenate([label[2 * s + k] for k in range(3)], axis=2).astype(np.double) / 255.
        ref_pred = np.clip(canv[s], 0, 1)
        for k in range(3):
            ref = utils._compute_psnr(ref_pred[:, :, 3 * k:3 * k + 3], gt[:, :, 3 * k:3 * k + 3], 1.)
            trunc = utils._compute_psnr(np.uint8(ref_pred[:, :, 3 * k:3 * k + 3] * 255) / 255., gt[:, :, 3 * k:3 * k + 3], 1.)
            worst = max(worst, abs(got[s][k] - ref))
            cost.append(ref - trunc)
    print("test(): worst |PSNR - oracle pipeline| = %.4f dB; scoring the uint8 canvas instead would cost %.2f dB" % (worst, np.mean(cost)))
    assert worst < 0.01
    assert np.mean(cost) > 0.3            # the case really distinguishes the two ways of scoring
    avg = re.search(r"Test \(average\) test_PSNR: FISR ([\d.]+)\[dB\], SR ([\d.]+)\[dB\]", out)
    fisr = [got[0][0], got[1][0], got[2][0], got[2][2]]                                   # FISRnet.py:913-920
    assert float(avg.group(1)) == pytest.approx(np.mean(fisr), abs=1e-6)
    assert float(avg.group(2)) == pytest.approx(np.mean([g[1] for g in got]), abs=1e-6)
    # PNGs: uint8 truncation of the clipped float canvas, through YUV2RGB_matlab (FISRnet.py:901-909)
    rgb = np.array(Image.open(os.path.join(args.test_img_dir, "FISRnet_exp1", "pred_seq_00.png")))
    want = P.yuv2rgb_matlab(P.quantise(canv[0])[:, :, :3]).astype("uint8")
    dd = np.abs(rgb.astype(int) - want.astype(int))
    assert dd.max() <= 3 and (dd == 0).mean() > 0.99          # one YUV level flips up to ~2 RGB levels
