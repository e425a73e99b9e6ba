"""-m gpu: the production tcgen05 conv kernel, one layer at a time, through the C ABI vs a float64 CPU conv."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fisrnet_oracle as O

pytestmark = pytest.mark.gpu

# (N, H, W, Cin, Cout): every (Cin, Cout) class of the network plus ragged / tiny geometries
SHAPES = [
    (1, 8, 16, 64, 64), (2, 32, 32, 64, 64), (8, 96, 96, 64, 64), (1, 24, 40, 64, 128), (1, 16, 16, 128, 128),
    (1, 136, 248, 128, 128), (1, 12, 20, 256, 256), (1, 6, 6, 512, 512), (1, 17, 31, 256, 512), (1, 34, 62, 512, 256),
    (2, 64, 96, 29, 64), (1, 48, 48, 38, 64), (1, 1, 1, 64, 64), (1, 3, 130, 64, 64), (1, 130, 3, 128, 64),
]


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
def test_conv_bias_residual_relu(engine, shape):
    engine.set_precision("f16x3")
    x, w, b, r = _case(*shape, seed=sum(shape))
    y = _ref(x, w, b, r)
    raw, act = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), r.cuda(), relu=True)
    tol = 2e-5 * max(1.0, float(y.abs().max()))          # fp32-class: the (hi, lo) split carries 22 mantissa bits
    assert (raw.cpu().double() - y).abs().max() < tol
    assert (act.cpu().double() - torch.relu(y)).abs().max() < tol


def test_conv_act_only_no_raw(engine):
    engine.set_precision("f16x3")
    x, w, b, _ = _case(2, 40, 56, 64, 64, seed=5, res=False)
    y = _ref(x, w, b, None)
    raw, act = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=True, want_raw=False)
    assert raw is None
    assert (act.cpu().double() - torch.relu(y)).abs().max() < 2e-5 * float(y.abs().max())


def test_conv_no_relu(engine):
    engine.set_precision("f16x3")
    x, w, b, _ = _case(1, 16, 24, 64, 64, seed=6, res=False)
    y = _ref(x, w, b, None)
    _, act = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=False)
    assert (act.cpu().double() - y).abs().max() < 2e-5 * float(y.abs().max())
    assert float(act.min()) < 0


def test_conv_depth_to_space_epilogue(engine):
    # conv/1 of the heads: relu + tf.depth_to_space(2) fused into the store (FISRnet.py:98-99)
    engine.set_precision("f16x3")
    x, w, b, _ = _case(2, 24, 40, 64, 256, seed=7, res=False)
    y = torch.relu(_ref(x, w, b, None))
    exp = O.to_nhwc(O.depth_to_space2(O.to_nchw(y)))
    _, act = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=True, d2s=True, want_raw=False)
    assert tuple(act.shape) == (2, 48, 80, 64)
    assert (act.cpu().double() - exp).abs().max() < 2e-5 * float(y.abs().max())


@pytest.mark.parametrize("cout", [6, 3])
def test_conv_narrow_head_output(engine, cout):
    engine.set_precision("f16x3")
    x, w, b, _ = _case(1, 64, 96, 64, cout, seed=cout, res=False)
    y = _ref(x, w, b, None)
    raw, act = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), None, relu=False)
    assert (raw.cpu().double() - y).abs().max() < 2e-5 * float(y.abs().max())
    assert (act.cpu().double() - y).abs().max() < 2e-5 * float(y.abs().max())


def test_conv_linearity_at_full_tile_size(engine):
    # size-independent property at a BASELINE-sized layer (64->64 @ 544x992): conv(x1 + x2) = conv(x1) + conv(x2) - b
    engine.set_precision("f16x3")
    g = torch.Generator().manual_seed(11)
    x1 = torch.rand(1, 544, 992, 64, generator=g).cuda()
    x2 = torch.rand(1, 544, 992, 64, generator=g).cuda()
    w = (torch.randn(3, 3, 64, 64, generator=g) * 0.06).cuda()
    b = (torch.randn(64, generator=g) * 0.1).cuda()
    r12, _ = engine.conv3x3(x1 + x2, w, b, None, relu=False, want_act=False)
    r1, _ = engine.conv3x3(x1, w, b, None, relu=False, want_act=False)
    r2, _ = engine.conv3x3(x2, w, b, None, relu=False, want_act=False)
    assert (r12 - (r1 + r2 - b)).abs().max() < 5e-5
    # and a spot check of 64 random pixels against the float64 reference
    xs = (x1 + x2).cpu()
    ys = _ref(xs[:, 100:110, 200:210], w.cpu(), b.cpu(), None)
    assert (r12.cpu()[:, 101:109, 201:209].double() - ys[:, 1:9, 1:9]).abs().max() < 5e-5


def test_fast_mode_is_fp16_accurate(engine):
    engine.set_precision("f16")
    try:
        x, w, b, r = _case(1, 32, 48, 128, 128, seed=9)
        y = _ref(x, w, b, r)
        raw, _ = engine.conv3x3(x.cuda(), w.cuda(), b.cuda(), r.cuda(), relu=True)
        err = float((raw.cpu().double() - y).abs().max())
        assert 1e-6 < err < 5e-3          # single fp16 operands: ~1e-3 relative, visibly not the split path
    finally:
        engine.set_precision("f16x3")
