"""-m gpu: training-side forward half (window assembly, Groups2Ovlp, temporal loss, train PSNR) and Adam vs the oracle."""
import numpy as np
import pytest
import torch

from oracle import fisrnet_oracle as O
from oracle import loss_oracle as L
from oracle import pipeline_oracle as P

pytestmark = pytest.mark.gpu


def _batch(B, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    data = torch.rand(B, h, w, 15, generator=g)
    flow = (torch.randn(B, h, w, 16, generator=g) * 4 / 96 / 2).clamp(-1, 1)
    flow2 = (torch.randn(B, h, w, 8, generator=g) * 8 / 96 / 2).clamp(-1, 1)
    warp = torch.rand(B, h, w, 24, generator=g)
    warp2 = torch.rand(B, h, w, 12, generator=g)
    label = torch.rand(B, 2 * h, 2 * w, 21, generator=g)
    return data, flow, flow2, warp, warp2, label


def test_groups2ovlp(engine):
    pred = torch.rand(6, 20, 28, 9)
    got = engine.groups2ovlp(pred.cuda()).cpu()
    ref = P.groups2ovlp(torch.cat([P.split_seq_dim(pred[i * 2:(i + 1) * 2]) for i in range(3)], dim=1))
    assert got.shape == (2, 7, 20, 28, 3) and torch.equal(got, ref)


def test_temporal_loss_kernel_on_given_predictions(engine):
    B, h, w = 2, 24, 40
    label = torch.rand(B, 2 * h, 2 * w, 21)
    preds = [torch.rand(4 * B, h // 2, w // 2, 9), torch.rand(4 * B, h, w, 9), torch.rand(4 * B, 2 * h, 2 * w, 9)]
    ref = L.temporal_loss([p.double() for p in preds], label.double())
    got = engine.temporal_loss([p.cuda() for p in preds], label.cuda())
    for k in L.SCALAR_NAMES:
        assert abs(got[k] - float(ref[k])) < 2e-6 * max(1.0, abs(float(ref[k]))), k
    lam = dict(recn=0.5, tm1=2.0, tm2=0.3, tmm=0.25, td=1.5, ss2=0.7)
    ref = L.temporal_loss([p.double() for p in preds], label.double(), lam)
    got = engine.temporal_loss([p.cuda() for p in preds], label.cuda(), lam)
    assert abs(got["total_loss"] - float(ref["total_loss"])) < 2e-6 * float(ref["total_loss"])


def test_train_forward_reference_native_shape(engine):
    """Reference-native training shape scaled down in batch: LR 96x96 (main.py:33-37), B=2 -> 8 network passes."""
    engine.set_precision("f16x3")
    params = O.init_params(21)
    engine.set_params(params)
    batch = _batch(2, 96, 96, seed=5)
    ref, _, _ = L.training_forward(params, *batch)
    got = engine.train_forward(*[t.cuda() for t in batch])
    for k in L.SCALAR_NAMES:
        assert abs(got[k] - float(ref[k])) < 1e-4 * max(1.0, abs(float(ref[k]))), (k, got[k], float(ref[k]))


def test_adam_matches_tf1_formula(engine):
    engine.set_precision("f16x3")
    params = {k: v.clone() for k, v in O.init_params(22).items()}
    engine.set_params(params)
    engine.adam_reset(0)
    m = {k: torch.zeros_like(v) for k, v in params.items()}
    v = {k: torch.zeros_like(p) for k, p in params.items()}
    g = torch.Generator().manual_seed(3)
    for t in (1, 2, 3):
        grads = {k: torch.randn(p.shape, generator=g) * 1e-3 for k, p in params.items()}
        params, m, v = L.adam_step_tf1(params, grads, m, v, t, lr=1e-4)
        assert engine.adam_step({k: x.cuda() for k, x in grads.items()}, lr=1e-4) == t
    got = engine.get_params()
    worst = max(float(np.abs(got[k] - params[k].numpy()).max()) for k in params)
    assert worst < 2e-7, worst
    # the re-packed operand planes follow the update: forward equals the oracle on the updated weights
    x = O.synthetic_input(1, 32, 32, 9)
    ref = O.model(params, x)[2]
    assert (engine.forward(x.cuda())[2].cpu() - ref).abs().max() < 1e-4
    engine.adam_reset(0)


def test_config3_full_batch_forward_and_loss(engine):
    """BASELINE configs[2] at its full shape -- B = 16, LR 192x192, label 384x384x21, i.e. a 64-image forward of the four
    weight-shared passes: the 11 scalars of FISRnet.py:651-657 against the fp32 oracle on the same batch (the forward half of
    the step; gradients at this patch size are checked in test_gpu_backward.py at batch 1, float64 autograd over 16 samples
    takes ~10 min of host time)."""
    engine.set_precision("f16x3")
    params = O.init_params(61)
    engine.set_params(params)
    batch = _batch(16, 192, 192, 62)
    ref, _, _ = L.training_forward(params, *batch)
    got = engine.train_forward(*[t.cuda() for t in batch])
    for k in L.SCALAR_NAMES:
        assert abs(got[k] - float(ref[k])) < 2e-5 * max(1.0, abs(float(ref[k]))), (k, got[k], float(ref[k]))
