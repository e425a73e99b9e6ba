"""-m gpu: precision mode f16f8 (fp16 main term + both cross terms as fp8 MMAs, 2 MMA units per K slice instead of 3).

The operand scheme is restated on the CPU in tools/precision_study_fp8.py (2.5e-5 max-abs on the cascade vs float64);
here the CUDA path is held to the north-star bar (1e-3 max-abs, PSNR within 0.01 dB) with a 5x margin."""
import numpy as np
import pytest
import torch
import torch.nn.functional as F

from oracle import fisrnet_oracle as O

pytestmark = pytest.mark.gpu

TOL_MODEL = 2e-4           # measured ~3e-5; the north-star bar is 1e-3
SHAPES = [
    (2, 32, 32, 64, 64), (1, 24, 40, 64, 128), (1, 136, 248, 128, 128), (1, 12, 20, 256, 256), (1, 17, 31, 256, 512),
    (2, 64, 96, 29, 64), (1, 48, 48, 38, 64), (1, 1, 1, 64, 64), (1, 3, 130, 64, 64), (1, 130, 3, 128, 64),
]


@pytest.fixture()
def eng8(engine):
    engine.set_precision("f16f8")
    yield engine
    engine.set_precision("f16x3")


def _ref(x, w, b, res):
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(), padding=1).permute(0, 2, 3, 1)
    return y + res.double() if res is not None else y


def _case(n, h, w, cin, cout, seed, res=True):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, h, w, cin, generator=g)
    wt = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    r = torch.randn(n, h, w, cout, generator=g) if res else None
    return x, wt, b, r


@pytest.mark.parametrize("shape", SHAPES)
def test_conv_layer(eng8, shape):
    x, w, b, r = _case(*shape, seed=sum(shape))
    y = _ref(x, w, b, r)
    raw, act = eng8.conv3x3(x.cuda(), w.cuda(), b.cuda(), r.cuda(), relu=True)
    # per layer: the fp8 cross terms leave ~2^-14 relative error per product, averaged down by the K-sum
    tol = 1e-4 * max(1.0, float(y.abs().max()))
    err = float((raw.cpu().double() - y).abs().max())
    assert err < tol, err
    # the stored activation is (fp16 hi, e5m2 lo): ~2^-14 relative on top
    assert (act.cpu().double() - torch.relu(y)).abs().max() < tol + 1.3e-4 * float(y.abs().max())


def test_conv_is_not_plain_fp16(eng8):
    x, w, b, r = _case(1, 32, 48, 128, 128, seed=9)
    y = _ref(x, w, b, r)
    raw8, _ = eng8.conv3x3(x.cuda(), w.cuda(), b.cuda(), r.cuda(), relu=True)
    eng8.set_precision("f16")
    raw16, _ = eng8.conv3x3(x.cuda(), w.cuda(), b.cuda(), r.cuda(), relu=True)
    eng8.set_precision("f16f8")
    e8, e16 = float((raw8.cpu().double() - y).abs().max()), float((raw16.cpu().double() - y).abs().max())
    assert e8 < e16 / 4, (e8, e16)


def test_depth_to_space_and_narrow_heads(eng8):
    x, w, b, _ = _case(2, 24, 40, 64, 256, seed=7, res=False)
    y = torch.relu(_ref(x, w, b, None))
    exp = O.to_nhwc(O.depth_to_space2(O.to_nchw(y)))
    _, act = eng8.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=True, d2s=True, want_raw=False)
    assert tuple(act.shape) == (2, 48, 80, 64)
    assert (act.cpu().double() - exp).abs().max() < 2.5e-4 * float(y.abs().max())
    for cout in (6, 3):
        x, w, b, _ = _case(1, 64, 96, 64, cout, seed=cout, res=False)
        y = _ref(x, w, b, None)
        raw, act = eng8.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=False)
        assert (raw.cpu().double() - y).abs().max() < 1e-4 * float(y.abs().max())
        assert (act.cpu().double() - y).abs().max() < 2.5e-4 * float(y.abs().max())


def test_model_config2_like(eng8):
    """192x192 patches (batch 2 of BASELINE config 2's 8): max-abs and PSNR parity against the fp32 oracle."""
    params = O.init_params(1)
    eng8.set_params(params)
    x = O.synthetic_input(2, 192, 192, 1)
    out = eng8.forward(x.cuda())
    ref = O.model(params, x)
    errs = [float((a.cpu().double() - b.double()).abs().max()) for a, b in zip(out, ref)]
    assert max(errs) < TOL_MODEL, errs
    gt = (ref[2] + 0.01 * torch.randn(ref[2].shape, generator=torch.Generator().manual_seed(3))).clamp(0, 1)
    assert abs(O.psnr(out[2].cpu().clamp(0, 1), gt) - O.psnr(ref[2].clamp(0, 1), gt)) < 0.01
    assert O.psnr(out[2].cpu(), ref[2]) > 90.0


def test_model_ragged_levels(eng8):
    """64x96 input: level-1 feature maps shrink to 2x3 pixels (tile-edge masking in every chunk layout)."""
    params = O.init_params(7)
    eng8.set_params(params)
    x = O.synthetic_input(1, 64, 96, 8)
    out = eng8.forward(x.cuda())
    ref = O.model(O.cast_params(params, torch.float64), x.double())
    errs = [float((a.cpu().double() - b).abs().max()) for a, b in zip(out, ref)]
    assert max(errs) < TOL_MODEL, errs


def test_window_matches_f16x3_to_one_grey_level(eng8):
    """Tiled video path: uint8 frames of the two precision modes differ by at most one level, on < 1 % of samples."""
    g = torch.Generator().manual_seed(5)
    H, W = 200, 330
    frames = torch.randint(0, 256, (H, W, 9), generator=g, dtype=torch.uint8)
    flow = torch.randn(H, W, 8, generator=g) * 4
    warp = torch.rand(H, W, 12, generator=g)
    eng8.set_params(O.init_params(3))
    a = eng8.window(frames.cuda(), flow.cuda(), warp.cuda(), (2, 2)).cpu().numpy().astype(np.int16)
    eng8.set_precision("f16x3")
    b = eng8.window(frames.cuda(), flow.cuda(), warp.cuda(), (2, 2)).cpu().numpy().astype(np.int16)
    eng8.set_precision("f16f8")
    d = np.abs(a - b)
    assert d.max() <= 1 and (d > 0).mean() < 0.01


def test_training_refuses_f16f8(eng8):
    import fisr_b200
    z = lambda c, s=1: torch.zeros(1, 32 * s, 32 * s, c, device="cuda")
    with pytest.raises(fisr_b200.FisrError):
        eng8.train_backward(z(15), z(16), z(8), z(24), z(12), z(21, 2))


def test_cta_pair_mode_is_bit_identical(eng8, monkeypatch):
    """FISR_PAIR=1 runs the wide f16f8 layers on CTA pairs (cluster of 2, tcgen05 cta_group::2, M = 256 MMAs, half of each tap's
    weight rows per CTA).  Every accumulator sees the same MMAs in the same order, so the outputs must not change by a bit;
    shapes with odd tile counts exercise the all-padding right tile of the last pair."""
    params = O.init_params(11)
    eng8.set_params(params)
    x = O.synthetic_input(1, 96, 160, 5)          # 160 / 16 = 10 tiles, 80 / 16 = 5 (odd), 40 / 16 -> 3 (odd)
    base = [t.clone() for t in eng8.forward(x.cuda())]
    monkeypatch.setenv("FISR_PAIR", "1")
    eng8.set_precision("f16x3"); eng8.set_precision("f16f8")       # drops the cached plans: geometry is re-planned with pairs
    paired = eng8.forward(x.cuda())
    monkeypatch.delenv("FISR_PAIR")
    for a, b in zip(base, paired):
        assert torch.equal(a, b)
    eng8.set_precision("f16x3"); eng8.set_precision("f16f8")


def test_full_size_window_deterministic_and_close_to_f16x3(eng8):
    """BASELINE config 4 at full size (1080x1920 -> 2048x3840x9): the f16f8 window is bit-reproducible, equals its pipelined host
    path, and differs from the fp32-class f16x3 frames by at most one grey level on < 1 % of the samples."""
    import os
    from conftest import GOLDEN
    rng = np.random.default_rng(2)
    base = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"]
    H, W = 1080, 1920
    reps = (-(-H // base.shape[1]), -(-W // base.shape[2]), 1)
    frames = np.concatenate([np.tile(base[i], reps)[:H, :W] for i in range(3)], axis=2).astype(np.uint8)
    flow = (rng.standard_normal((H, W, 8)) * 4).astype(np.float32)
    warp = (frames[..., [3, 4, 5, 0, 1, 2, 6, 7, 8, 3, 4, 5]].astype(np.float32) / 255.
            + 0.05 * rng.standard_normal((H, W, 12))).astype(np.float32)
    eng8.set_params(O.init_params(10))
    f, fl, wp = (torch.from_numpy(a).cuda() for a in (frames, flow, warp))
    a = eng8.window(f, fl, wp, (2, 2))
    assert tuple(a.shape) == (2048, 3840, 9)
    assert torch.equal(a, eng8.window(f, fl, wp, (2, 2)))
    host = eng8.window_host(frames, flow, warp, (2, 2))
    assert np.array_equal(host, a.cpu().numpy())
    eng8.set_precision("f16x3")
    b = eng8.window(f, fl, wp, (2, 2))
    eng8.set_precision("f16f8")
    d = (a.to(torch.int16) - b.to(torch.int16)).abs()
    assert int(d.max()) <= 1 and float((d > 0).float().mean()) < 0.01
