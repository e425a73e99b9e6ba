"""-m gpu: the tcgen05 weight-gradient kernel (MN-major operands, split-K) through the C ABI vs float64 autograd."""
import pytest
import torch
import torch.nn.functional as F

pytestmark = pytest.mark.gpu

# (N, H, W, Cin, Cout): channel classes of the network plus ragged geometries (tile = 8 x 16 pixels)
SHAPES = [
    (1, 8, 16, 64, 64), (2, 32, 32, 64, 64), (4, 96, 96, 64, 64), (1, 24, 40, 64, 128), (1, 16, 16, 128, 128),
    (1, 12, 20, 256, 256), (1, 6, 6, 512, 512), (1, 17, 31, 256, 512), (1, 34, 62, 512, 256), (2, 64, 96, 29, 64),
    (1, 48, 48, 38, 64), (1, 1, 1, 64, 64), (1, 3, 130, 64, 64), (1, 130, 3, 128, 64), (1, 64, 96, 64, 6), (1, 40, 24, 64, 3),
    (1, 24, 24, 64, 256),
]


def _ref(x, dy):
    xx = x.double().permute(0, 3, 1, 2)
    w = torch.zeros(dy.shape[3], x.shape[3], 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(xx, w, padding=1)
    (gw,) = torch.autograd.grad(y, w, dy.double().permute(0, 3, 1, 2))
    return gw.permute(2, 3, 1, 0), dy.double().sum(dim=(0, 1, 2))


@pytest.mark.parametrize("shape", SHAPES)
@pytest.mark.parametrize("exact", [True, False])
def test_wgrad_matches_autograd(engine, shape, exact):
    engine.set_precision("f16x3")
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    x = torch.rand(n, h, w, cin, generator=g)
    dy = torch.randn(n, h, w, cout, generator=g) * 0.05
    gw_ref, gb_ref = _ref(x, dy)
    engine.set_wgrad_exact(exact)
    try:
        gw, gb = engine.wgrad3x3(x.cuda(), dy.cuda())
    finally:
        engine.set_wgrad_exact(False)
    # exact: both operands carry 22 bits.  Default: layers with >= 16384 pixels round x to fp16 in this product (2^-12).
    fast = (not exact) and n * h * w >= 16384
    tol = (3e-4 if fast else 2e-5) * max(1.0, float(gw_ref.abs().max()))
    err = float((gw.cpu().double() - gw_ref).abs().max())
    assert err < tol
    if fast:
        assert err > 1e-7            # the fast path really ran
    assert (gb.cpu().double() - gb_ref).abs().max() < 2e-5 * max(1.0, float(gb_ref.abs().max()))


def test_wgrad_scale_and_linearity(engine):
    engine.set_precision("f16x3")
    g = torch.Generator().manual_seed(3)
    x = torch.rand(2, 40, 56, 64, generator=g).cuda()
    d1 = (torch.randn(2, 40, 56, 64, generator=g) * 0.1).cuda()
    d2 = (torch.randn(2, 40, 56, 64, generator=g) * 0.1).cuda()
    g1, b1 = engine.wgrad3x3(x, d1)
    g2, b2 = engine.wgrad3x3(x, d2)
    g12, b12 = engine.wgrad3x3(x, d1 + d2, scale=0.5)
    assert (2 * g12 - (g1 + g2)).abs().max() < 1e-4 * float(g1.abs().max())
    assert (2 * b12 - (b1 + b2)).abs().max() < 1e-4 * float(b1.abs().max())
