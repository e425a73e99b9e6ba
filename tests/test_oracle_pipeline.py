"""Oracle checks for the callers of the hot path: windows, tile geometry, colour, quantisation, warp (CPU)."""
import os

import numpy as np
import torch

from oracle import pipeline_oracle as P
from conftest import GOLDEN


def test_window_assembly_channels():
    b, h, w = 1, 4, 4
    data = torch.arange(15.).expand(b, h, w, 15)
    flow = 100 + torch.arange(16.).expand(b, h, w, 16)
    warp = 200 + torch.arange(24.).expand(b, h, w, 24)
    for i in range(3):
        x = P.window_input(data, flow, warp, i)
        assert x.shape == (b, h, w, 29)
        assert x[0, 0, 0, :9].tolist() == list(range(3 * i, 3 * i + 9))                   # ops.py:90-96
        assert x[0, 0, 0, 9:17].tolist() == [100 + c for c in range(4 * i, 4 * i + 8)]    # ops.py:99-106
        assert x[0, 0, 0, 17:].tolist() == [200 + c for c in range(6 * i, 6 * i + 12)]    # ops.py:109-116


def test_groups2ovlp():
    g = torch.arange(9.).reshape(1, 9, 1, 1, 1).expand(1, 9, 2, 2, 3)
    o = P.groups2ovlp(g)
    assert o.shape == (1, 7, 2, 2, 3)
    assert o[0, :, 0, 0, 0].tolist() == [0, 1, 2.5, 4, 5.5, 7, 8]                         # ops.py:119-144


def test_tile_geometry_default_grid():
    # (2,2) grid on 1080x1920 (main.py:100-103): h = 1024, w = 1920, four 544x992 tiles, canvas 2048x3840
    h, w = P.crop_hw(1080, 1920, (2, 2))
    assert (h, w) == (1024, 1920)
    sizes = []
    for p in range(4):
        pH, pW = p // 2, p % 2
        lo_h, hi_h, lo_w, hi_w, add_h, add_w = P.get_hw_boundary(32, h, w, pH, h // 2, pW, w // 2)
        sizes.append((hi_h - lo_h, hi_w - lo_w))
        assert (add_h, add_w) == (32, 32)
    assert sizes == [(544, 992)] * 4
    assert P.crop_hw(1080, 1920, (1, 1)) == (1056, 1920)      # output height depends on the grid


def test_tiled_window_reassembles_identity_network():
    rng = np.random.default_rng(0)
    inp = rng.random((1, 128, 192, 29))

    def fake_model(tile):        # "network" = x2 nearest upsample of the first 9 channels: tiling must be invisible
        return np.repeat(np.repeat(tile[..., :9], 2, axis=1), 2, axis=2)

    for grid in ((2, 2), (1, 1), (2, 3), (4, 2)):
        h, w = P.crop_hw(128, 192, grid)
        full = P.tiled_window(fake_model, inp[:, :h, :w], grid)
        assert np.array_equal(full, fake_model(inp[:, :h, :w])[0])


def test_quantise_truncates():
    x = np.array([[-0.2, 0.0, 0.999 / 255, 1.0 / 255, 0.5, 1.0, 1.7]])
    assert P.quantise(x).tolist() == [[0, 0, 0, 1, 127, 255, 255]]                         # FISRnet.py:1060-1064


def test_yuv2rgb_matches_shipped_frames():
    # The reference ships pred_YUV_k.png and pred_k.png = uint8(YUV2RGB_matlab(yuv)) (FISRnet.py:1066-1077).
    # They come from different runs, so a small fraction differs by a few LSB: require >= 99.9 %.
    g = np.load(os.path.join(GOLDEN, "scene1_yuv_rgb.npz"))
    keys = [k for k in g.files if k.startswith("yuv_")]
    assert len(keys) == 4
    for k in keys:
        rgb = P.yuv2rgb_matlab(g[k]).astype("uint8")
        ref = g["rgb_" + k[4:]]
        same = np.mean(rgb == ref)
        assert same >= 0.999, (k, same)
        assert np.abs(rgb.astype(int) - ref.astype(int)).max() <= 6


def test_fixture_geometry():
    g = np.load(os.path.join(GOLDEN, "scene1_yuv_rgb.npz"))
    H, W = g["input_hw"]
    h, w = P.crop_hw(int(H), int(W), (2, 2))
    assert tuple(g["output_hw"]) == (2 * h, 2 * w) == (2048, 3840)
    assert int(g["n_outputs"]) == 2 * int(g["n_inputs"]) - 3                                # FISRnet.py:1066-1077


def test_colour_round_trip():
    g = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"]
    yuv = g[0].astype(np.float32)
    back = P.rgb2yuv(P.yuv2rgb_matlab(yuv))
    # values that clip in RGB do not round-trip; the rest must
    ok = (P.yuv2rgb_matlab(yuv) > 0.5).all(-1) & (P.yuv2rgb_matlab(yuv) < 254.5).all(-1)
    assert ok.mean() > 0.9
    assert np.abs(back - yuv)[ok].max() < 0.05


def test_warp_fixedpoint_restatement_matches_cv2():
    g = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"]
    rng = np.random.default_rng(1)
    img = P.yuv2rgb_matlab(g[1].astype(np.float32))
    flow = (rng.standard_normal(g[1].shape[:2] + (2,)) * 3).astype(np.float32)
    a = P.warp_flow_cv2(img, flow)
    b = P.warp_flow_fixedpoint(img, flow)
    assert np.abs(a - b).max() < 1e-3                       # 0..255 scale
    # zero flow is the identity, large flow replicates the border
    assert np.abs(P.warp_flow_cv2(img, np.zeros_like(flow)) - img).max() < 1e-4
    far = np.full_like(flow, 1e4)
    assert np.allclose(P.warp_flow_cv2(img, far), img[-1, -1], atol=1e-4)


def test_warp_pair_direction():
    # slot 0 = frame 2 sampled with 0.5 * flow(1->2), slot 1 = frame 1 with 0.5 * flow(2->1) (..warp_img_with_flo.py:121-128)
    g = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"]
    z = np.zeros(g[0].shape[:2] + (2,), np.float32)
    out = P.warp_pair_yuv(g[0], g[1], z, z)
    for slot, src in ((0, g[1]), (1, g[0])):
        ok = (P.yuv2rgb_matlab(src.astype(np.float32)) > 0.5).all(-1) & (P.yuv2rgb_matlab(src.astype(np.float32)) < 254.5).all(-1)
        assert np.abs(out[slot] - src)[ok].max() < 0.05
