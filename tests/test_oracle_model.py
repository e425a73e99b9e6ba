"""Oracle self-checks (CPU): inventory, TF-1.13 op restatements, fp32-vs-fp64 consistency, golden regression."""
import os

import numpy as np
import torch

from oracle import fisrnet_oracle as O
from conftest import GOLDEN


def test_inventory_matches_survey():
    inv = O.conv_inventory()
    assert len(inv) == 138                                             # 46 convs x 3 levels (FISRnet.py:78-171)
    n_params = sum(9 * ci * co + co for ci, co in inv.values())
    assert n_params == 48_316_251
    assert inv["FISRnet/level_1/enc/level_0/conv/0"] == (29, 64)       # FISRnet.py:84
    assert inv["FISRnet/level_2/enc/level_0/conv/0"] == (38, 64)       # FISRnet.py:116 (sz[-1]+9)
    assert inv["FISRnet/level_3/dec/level_2/conv/0"] == (512, 256)     # ops.py:73 (c*2 -> c)
    assert inv["FISRnet/level_3/FI-SR/conv/2"] == (64, 6) and inv["FISRnet/level_3/SR/conv/2"] == (64, 3)
    names = list(inv)
    assert names[0] == "FISRnet/level_1/enc/level_0/conv/0" and names[-1] == "FISRnet/level_3/SR/conv/2"


def test_depth_to_space_is_tf_order():
    # tf.depth_to_space(x, 2) NHWC: out[n, 2h+i, 2w+j, c] = x[n, h, w, (2i+j)*C + c]  (NOT torch PixelShuffle order)
    n, h, w, c = 2, 3, 4, 5
    x = torch.arange(n * h * w * 4 * c, dtype=torch.float32).reshape(n, h, w, 4 * c)
    y = O.to_nhwc(O.depth_to_space2(O.to_nchw(x)))
    for i in range(2):
        for j in range(2):
            assert torch.equal(y[:, i::2, j::2, :], x[..., (2 * i + j) * c:(2 * i + j + 1) * c])
    ps = torch.nn.functional.pixel_shuffle(O.to_nchw(x), 2)
    assert not torch.equal(O.to_nchw(y), ps)


def test_legacy_bilinear_upsample():
    # resize_images(BILINEAR), TF 1.13 legacy: src = dst * 0.5, last sample replicated
    x = torch.rand(1, 2, 5, 7)
    y = O.upsample2_legacy_bilinear(x)
    ref = torch.empty(1, 2, 10, 14)
    for Y in range(10):
        for X in range(14):
            y0, x0 = Y // 2, X // 2
            y1 = min(y0 + 1, 4) if Y % 2 else y0
            x1 = min(x0 + 1, 6) if X % 2 else x0
            ref[:, :, Y, X] = 0.25 * (x[:, :, y0, x0] + x[:, :, y1, x0] + x[:, :, y0, x1] + x[:, :, y1, x1])
    assert torch.allclose(y, ref, atol=1e-6)
    assert torch.equal(y[:, :, ::2, ::2], x)
    # differs from both torch alignments
    assert not torch.allclose(y, torch.nn.functional.interpolate(x, scale_factor=2, mode="bilinear", align_corners=False), atol=1e-3)


def test_subsample_is_strided_slice():
    x = torch.rand(1, 3, 8, 12)
    assert torch.equal(O.subsample(x, 4), x[:, :, ::4, ::4])


def test_model_shapes_config1_and_dtype_consistency():
    p64 = O.init_params(0, torch.float64)
    p32 = O.cast_params(p64, torch.float32)
    x = O.synthetic_input(1, 64, 64, 5)
    o64 = O.model(p64, x)
    o32 = O.model(p32, x)
    assert [tuple(o.shape) for o in o32] == [(1, 32, 32, 9), (1, 64, 64, 9), (1, 128, 128, 9)]
    assert o32[0].dtype == torch.float32 and o64[0].dtype == torch.float64
    for a, b in zip(o32, o64):
        assert (a.double() - b).abs().max() < 5e-6        # fp32 noise floor of the 138-conv cascade


def test_model_rejects_non_multiple_of_32():
    p = O.init_params(0)
    try:
        O.model(p, torch.zeros(1, 48, 64, 29))
    except AssertionError:
        return
    raise AssertionError("48 is not a multiple of 32 (FISRnet.py:818-824)")


def test_pred_channel_order():
    # pred = concat(FISR[..., :3], SR, FISR[..., 3:])  (FISRnet.py:107-108): zero all convs but SR/conv/2 bias
    p = O.init_params(1, bias_std=0.0)
    p["FISRnet/level_3/SR/conv/2/b"] = torch.tensor([1.0, 2.0, 3.0])
    for k in ("w",):
        p[f"FISRnet/level_3/SR/conv/2/{k}"] = torch.zeros_like(p[f"FISRnet/level_3/SR/conv/2/{k}"])
        p[f"FISRnet/level_3/FI-SR/conv/2/{k}"] = torch.zeros_like(p[f"FISRnet/level_3/FI-SR/conv/2/{k}"])
    p["FISRnet/level_3/FI-SR/conv/2/b"] = torch.tensor([10., 11., 12., 13., 14., 15.])
    o = O.model(p, O.synthetic_input(1, 32, 32, 2))[2]
    assert torch.allclose(o[0, 5, 7], torch.tensor([10., 11., 12., 1., 2., 3., 13., 14., 15.]))


def test_golden_regression():
    g = np.load(os.path.join(GOLDEN, "model_fp64_64x96.npz"))
    p64 = O.init_params(7, torch.float64)
    assert abs(float(sum(v.abs().sum() for v in p64.values())) - float(g["param_abs_sum"])) < 1e-6
    x = O.synthetic_input(1, 64, 96, 8)
    assert abs(float(x.double().sum()) - float(g["input_sum"])) < 1e-6
    o = O.model(O.cast_params(p64, torch.float32), x)
    for a, k in zip(o, ("pred_l1", "pred_l2", "pred_l3")):
        assert np.abs(a.numpy() - g[k]).max() < 5e-6
