"""Oracle (test infrastructure): multi-scale temporal loss, train PSNR and the TF-1.13 Adam update.

Restates ``FISRnet.build_model`` (``FISRnet.py:250-491``) with torch so that it is differentiable:
  * window assembly for stride 1 and stride 2          -- ``FISRnet.py:281-306, 392-409``
  * ``Groups2Ovlp``                                      -- ``ops.py:119-144``
  * the seven loss terms at three scales                -- ``FISRnet.py:312-484`` (scale weights 1, 2, 4 at :326-328)
  * ``train_PSNR``                                       -- ``FISRnet.py:485-486`` (``tf.image.psnr`` per image, then mean)
  * ``tf.train.AdamOptimizer(lr).minimize``              -- ``FISRnet.py:489-491`` (TF-1.13 formula, epsilon outside the
                                                           bias correction: not ``torch.optim.Adam``)
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Dict, Sequence

import torch

from . import fisrnet_oracle as net
from .pipeline_oracle import groups2ovlp, split_seq_dim, window_input

LAMBDAS = dict(recn=1.0, tm1=1.0, tm2=0.1, tmm=1.0, td=0.1, ss2=1.0)      # main.py:80-85 defaults
SCALAR_NAMES = ("recnLoss", "tmLoss", "tmmLoss", "tdLoss", "totalLoss_s1", "recnLoss_ss2", "tdLoss_ss2", "tmLoss_ss2",
                "totalLoss_ss2", "total_loss", "train_PSNR")


def l2(x: torch.Tensor, y: torch.Tensor) -> torch.Tensor:
    """``L2_loss`` ops.py:30-32."""
    return torch.mean((x - y) ** 2)


def ss2_input(data: torch.Tensor, flow_ss2: torch.Tensor, warp_ss2: torch.Tensor) -> torch.Tensor:
    """Stride-2 network input (FISRnet.py:392-399): frames 0, 2, 4 + 8 flow + 12 warp channels."""
    frames = torch.cat([data[..., 0:3], data[..., 6:9], data[..., 12:15]], dim=3)
    return torch.cat([frames, flow_ss2, warp_ss2], dim=3)


def batch_inputs(data, flow, flow_ss2, warp, warp_ss2) -> torch.Tensor:
    """The four weight-shared passes of one training step as one batch [4B,h,w,29]: windows 0, 1, 2 then stride 2."""
    return torch.cat([window_input(data, flow, warp, i) for i in range(3)] + [ss2_input(data, flow_ss2, warp_ss2)], dim=0)


def temporal_loss(preds: Sequence[torch.Tensor], label: torch.Tensor, lambdas: Dict[str, float] = LAMBDAS) -> "OrderedDict[str, torch.Tensor]":
    """preds = (pred_l1, pred_l2, pred_l3) of the 4B-batch of :func:`batch_inputs` (pass-major), label [B,2h,2w,21].
    Returns the 11 scalars the reference logs every step (FISRnet.py:651-657)."""
    B = label.shape[0]
    gts = [split_seq_dim(label[:, ::4, ::4, :]), split_seq_dim(label[:, ::2, ::2, :]), split_seq_dim(label)]   # :263-264,277-279
    weights = (4.0, 2.0, 1.0)                                                   # l1, l2, l3  (:326-328)
    z = label.new_zeros(())
    recn, tm, tmm, td, recn2, td2, tm2 = z, z, z, z, z, z, z
    ovlp_l3 = None
    for pred, gt, wgt in zip(preds, gts, weights):
        P = torch.cat([split_seq_dim(pred[i * B:(i + 1) * B]) for i in range(3)], dim=1)     # [B,9,h,w,3]  (:291-306)
        S = split_seq_dim(pred[3 * B:4 * B])                                                  # [B,3,h,w,3]  (:405-409)
        O = groups2ovlp(P)                                                                    # [B,7,h,w,3]  (:308-310)
        for i in range(3):
            recn = recn + wgt * l2(P[:, 3 * i:3 * i + 3], gt[:, 2 * i:2 * i + 3])             # :315-328
        for i in range(2):
            tm = tm + wgt * l2(P[:, 3 * i + 2:3 * i + 3], P[:, 3 * i + 3:3 * i + 4])          # :331-341
            tmm = tmm + wgt * l2((P[:, 3 * i + 2:3 * i + 3] + P[:, 3 * i + 3:3 * i + 4]) / 2, gt[:, 2 * (i + 1):2 * (i + 1) + 1])  # :344-357
        for i in range(6):
            td = td + wgt * l2(O[:, i + 1:i + 2] - O[:, i:i + 1], gt[:, i + 1:i + 2] - gt[:, i:i + 1])   # :360-385
        gt2 = gt[:, 1::2]                                                                     # GT 1, 3, 5  (:412-423)
        recn2 = recn2 + wgt * l2(S, gt2)                                                      # :426-428
        for i in range(2):
            td2 = td2 + wgt * l2(S[:, i + 1:i + 2] - S[:, i:i + 1], gt2[:, i + 1:i + 2] - gt2[:, i:i + 1])   # :431-459
        tm2 = tm2 + wgt * l2(S, O[:, 1::2])                                                   # :462-477
        ovlp_l3 = O
    s1 = lambdas["recn"] * recn + lambdas["tm1"] * tm + lambdas["tmm"] * tmm + lambdas["td"] * td       # :388-389
    s2 = lambdas["recn"] * recn2 + lambdas["td"] * td2 + lambdas["tm2"] * tm2                           # :480-481
    total = s1 + lambdas["ss2"] * s2                                                                    # :484
    mse = torch.mean((ovlp_l3 - gts[2]) ** 2, dim=(2, 3, 4))                                            # per (image, frame)
    psnr = torch.mean(10.0 * torch.log10(1.0 / mse))                                                    # :485-486, max_val 1
    return OrderedDict(zip(SCALAR_NAMES, (recn, tm, tmm, td, s1, recn2, td2, tm2, s2, total, psnr)))


def training_forward(params, data, flow, flow_ss2, warp, warp_ss2, label, lambdas=LAMBDAS, grad: bool = False):
    """Whole forward half of a training step on the oracle network; with ``grad=True`` also d total_loss / d params."""
    dtype = next(iter(params.values())).dtype
    x = batch_inputs(data, flow, flow_ss2, warp, warp_ss2).to(dtype)
    if grad:
        params = OrderedDict((k, v.clone().requires_grad_(True)) for k, v in params.items())
    with torch.set_grad_enabled(grad):
        outs = net.Net(params).model_nchw(net.to_nchw(x))
        preds = [net.to_nhwc(o) for o in outs]
        scalars = temporal_loss(preds, label.to(dtype), lambdas)
        grads = None
        if grad:
            g = torch.autograd.grad(scalars["total_loss"], list(params.values()))
            grads = OrderedDict(zip(params.keys(), g))
    return scalars, [p.detach() for p in preds], grads


def adam_step_tf1(params, grads, m, v, t: int, lr: float, beta1=0.9, beta2=0.999, eps=1e-8):
    """One ``tf.train.AdamOptimizer`` update (TF 1.13 ``_apply_dense``), step counter ``t`` >= 1:
    lr_t = lr * sqrt(1 - beta2^t) / (1 - beta1^t);  m, v moving averages;  theta -= lr_t * m / (sqrt(v) + eps)."""
    lr_t = lr * math.sqrt(1.0 - beta2 ** t) / (1.0 - beta1 ** t)
    for k in params:
        g = grads[k]
        m[k] = beta1 * m[k] + (1.0 - beta1) * g
        v[k] = beta2 * v[k] + (1.0 - beta2) * g * g
        params[k] = params[k] - lr_t * m[k] / (torch.sqrt(v[k]) + eps)
    return params, m, v


def piecewise_lr(step: int, train_iter: int, init_lr=1e-4, points=(80, 90), factor=0.1) -> float:
    """``tf.train.piecewise_constant`` schedule of FISRnet.py:232-240 (boundaries inclusive on the left value)."""
    lr = init_lr
    for k, p in enumerate(points):
        if step > p * train_iter:
            lr = init_lr * factor ** (k + 1)
    return lr
