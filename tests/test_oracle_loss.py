"""Oracle checks (CPU) for the multi-scale temporal loss, train PSNR and the TF-1.13 Adam formula."""
import math

import torch

from oracle import fisrnet_oracle as O
from oracle import loss_oracle as L


def _perfect_preds(label, B):
    """Predictions that reproduce the labels at every scale: window w frame f = GT[2w+f], stride-2 frame f = GT[2f+1]."""
    preds = []
    for st in (4, 2, 1):
        g = label[:, ::st, ::st, :]                                   # [B,h,w,21]
        passes = [torch.cat([g[..., 3 * (2 * w + f):3 * (2 * w + f) + 3] for f in range(3)], dim=3) for w in range(3)]
        passes.append(torch.cat([g[..., 3 * (2 * f + 1):3 * (2 * f + 1) + 3] for f in range(3)], dim=3))
        preds.append(torch.cat(passes, dim=0))                         # [4B,h,w,9]
    return preds


def test_loss_is_zero_for_perfect_predictions_and_known_for_a_shift():
    B = 2
    label = torch.rand(B, 32, 48, 21, dtype=torch.float64)
    preds = _perfect_preds(label, B)
    s = L.temporal_loss(preds, label)
    for k in L.SCALAR_NAMES[:10]:
        assert abs(float(s[k])) < 1e-24, k
    assert math.isinf(float(s["train_PSNR"]))
    # shift window 0 / frame 0 of the finest scale by delta: recn = d^2/3 (one of 3 frames of one window term),
    # td = d^2 (only O[1]-O[0] changes), nothing else moves; total = recn + 0.1 td (main.py:80-85)
    d = 0.05
    preds[2] = preds[2].clone()
    preds[2][:B, :, :, 0:3] += d
    s = L.temporal_loss(preds, label)
    assert abs(float(s["recnLoss"]) - d * d / 3) < 1e-12
    assert abs(float(s["tdLoss"]) - d * d) < 1e-12
    assert abs(float(s["tmLoss"])) < 1e-24 and abs(float(s["recnLoss_ss2"])) < 1e-24
    assert abs(float(s["total_loss"]) - (d * d / 3 + 0.1 * d * d)) < 1e-12
    # PSNR: frame 0 has mse d^2, the other six are exact -> inf dominates the mean like tf.reduce_mean would
    assert math.isinf(float(s["train_PSNR"]))


def test_scale_weights_and_ss2_terms():
    B = 1
    label = torch.rand(B, 16, 16, 21, dtype=torch.float64)
    preds = _perfect_preds(label, B)
    d = 0.1
    preds[0] = preds[0].clone()
    preds[0][3 * B:, :, :, :] += d                      # stride-2 pass at level 1 (weight 4, FISRnet.py:326-328)
    s = L.temporal_loss(preds, label)
    assert abs(float(s["recnLoss_ss2"]) - 4 * d * d) < 1e-12           # all 3 frames off by d
    assert abs(float(s["tmLoss_ss2"]) - 4 * d * d) < 1e-12
    assert abs(float(s["tdLoss_ss2"])) < 1e-20                          # differences unchanged
    assert abs(float(s["totalLoss_ss2"]) - (4 * d * d + 0.1 * 4 * d * d)) < 1e-12
    assert abs(float(s["totalLoss_s1"])) < 1e-24


def test_pass_assembly_matches_reference_slices():
    B, h, w = 1, 4, 4
    data = torch.arange(15.).expand(B, h, w, 15)
    flow = 100 + torch.arange(16.).expand(B, h, w, 16)
    flow2 = 300 + torch.arange(8.).expand(B, h, w, 8)
    warp = 200 + torch.arange(24.).expand(B, h, w, 24)
    warp2 = 400 + torch.arange(12.).expand(B, h, w, 12)
    x = L.batch_inputs(data, flow, flow2, warp, warp2)
    assert x.shape == (4, h, w, 29)
    assert x[3, 0, 0, :9].tolist() == [0, 1, 2, 6, 7, 8, 12, 13, 14]      # frames 0, 2, 4 (FISRnet.py:394-398)
    assert x[3, 0, 0, 9:17].tolist() == [300 + c for c in range(8)]
    assert x[1, 0, 0, :9].tolist() == list(range(3, 12))


def test_adam_tf1_formula():
    p = {"a": torch.tensor([1.0, -2.0])}
    g = {"a": torch.tensor([0.5, 0.25])}
    m = {"a": torch.zeros(2)}
    v = {"a": torch.zeros(2)}
    p, m, v = L.adam_step_tf1(p, g, m, v, t=1, lr=1e-3)
    # step 1: m = 0.1 g, v = 0.001 g^2, lr_t = lr*sqrt(0.001)/0.1 -> update = lr * g/|g| / (1 + eps/(sqrt(0.001)|g|)) ~ lr
    lr_t = 1e-3 * math.sqrt(1 - 0.999) / (1 - 0.9)
    exp = torch.tensor([1.0, -2.0]) - lr_t * (0.1 * g["a"]) / (torch.sqrt(0.001 * g["a"] ** 2) + 1e-8)
    assert torch.allclose(p["a"], exp, atol=1e-9)
    assert L.piecewise_lr(0, 1220) == 1e-4 and abs(L.piecewise_lr(81 * 1220, 1220) - 1e-5) < 1e-12


def test_gradient_oracle_matches_finite_difference():
    torch.manual_seed(0)
    p64 = O.init_params(3, torch.float64)
    B, h, w = 1, 32, 32
    data, flow, flow2 = torch.rand(B, h, w, 15), torch.randn(B, h, w, 16) * 0.02, torch.randn(B, h, w, 8) * 0.02
    warp, warp2, label = torch.rand(B, h, w, 24), torch.rand(B, h, w, 12), torch.rand(B, 2 * h, 2 * w, 21)
    s, preds, grads = L.training_forward(p64, data, flow, flow2, warp, warp2, label, grad=True)
    assert preds[2].shape == (4 * B, 2 * h, 2 * w, 9) and len(grads) == 276
    name = "FISRnet/level_3/SR/conv/2/b"
    eps = 1e-5
    pp = dict(p64); pp[name] = p64[name].clone(); pp[name][1] += eps
    s2, _, _ = L.training_forward(pp, data, flow, flow2, warp, warp2, label)
    fd = (float(s2["total_loss"]) - float(s["total_loss"].detach())) / eps
    assert abs(fd - float(grads[name][1])) < 1e-4 * max(1.0, abs(fd))
