"""-m gpu: data gradient of the 3x3 conv (forward tcgen05 kernel on rotated-transposed planes, gate / residual /
space-to-depth epilogues) through the C ABI vs float64 autograd."""
import pytest
import torch
import torch.nn.functional as F

from oracle import fisrnet_oracle as O

pytestmark = pytest.mark.gpu

SHAPES = [
    (1, 8, 16, 64, 64), (2, 32, 32, 64, 64), (1, 24, 40, 64, 128), (1, 16, 16, 128, 64), (1, 12, 20, 256, 256),
    (1, 6, 6, 512, 512), (1, 17, 31, 256, 512), (1, 34, 62, 512, 256), (2, 64, 96, 38, 64), (1, 48, 48, 64, 6),
    (1, 40, 24, 64, 3), (1, 24, 24, 64, 256), (1, 3, 130, 64, 64),
]


def _ref(dy, w, mask, res):
    cin = w.shape[2]
    x = torch.zeros(dy.shape[0], cin, dy.shape[1], dy.shape[2], dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x, w.double().permute(3, 2, 0, 1), padding=1)
    (gx,) = torch.autograd.grad(y, x, dy.double().permute(0, 3, 1, 2))
    gx = gx.permute(0, 2, 3, 1)
    if mask is not None:
        gx = gx * (mask > 0).double()
    if res is not None:
        gx = gx + res.double()
    return gx


@pytest.mark.parametrize("shape", SHAPES)
def test_dgrad_gate_and_residual(engine, shape):
    engine.set_precision("f16x3")
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(sum(shape))
    dy = torch.randn(n, h, w, cout, generator=g) * 0.1
    wt = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cout)) ** 0.5
    mask = torch.relu(torch.randn(n, h, w, cin, generator=g))
    res = torch.randn(n, h, w, cin, generator=g) * 0.1
    ref = _ref(dy, wt, mask, res)
    raw, act = engine.dgrad3x3(dy.cuda(), wt.cuda(), mask.cuda(), res.cuda())
    tol = 2e-5 * max(1.0, float(ref.abs().max()))
    assert (raw.cpu().double() - ref).abs().max() < tol
    assert (act.cpu().double() - ref).abs().max() < tol


def test_dgrad_plain(engine):
    engine.set_precision("f16x3")
    g = torch.Generator().manual_seed(1)
    dy = torch.randn(2, 24, 40, 128, generator=g) * 0.1
    wt = torch.randn(3, 3, 64, 128, generator=g) * 0.05
    ref = _ref(dy, wt, None, None)
    _, act = engine.dgrad3x3(dy.cuda(), wt.cuda(), want_raw=False)
    assert (act.cpu().double() - ref).abs().max() < 2e-5 * float(ref.abs().max())


def test_dgrad_space_to_depth(engine):
    # conv/2 of the heads reads depth_to_space(relu(conv/1)): its data gradient is gated by the shuffled activation and
    # stored space-to-depth, the adjoint of FISRnet.py:99
    engine.set_precision("f16x3")
    g = torch.Generator().manual_seed(2)
    dy = torch.randn(2, 32, 48, 6, generator=g) * 0.1
    wt = torch.randn(3, 3, 64, 6, generator=g) * 0.1
    mask = torch.relu(torch.randn(2, 32, 48, 64, generator=g))
    ref = _ref(dy, wt, mask, None)                                     # [2,32,48,64] at the shuffled resolution
    # adjoint of depth_to_space: out[n, y/2, x/2, (2*(y%2) + x%2)*64 + c] = in[n, y, x, c]
    exp = torch.zeros(2, 16, 24, 256, dtype=torch.float64)
    for i in range(2):
        for j in range(2):
            exp[..., (2 * i + j) * 64:(2 * i + j + 1) * 64] = ref[:, i::2, j::2, :]
    chk = O.to_nhwc(O.depth_to_space2(O.to_nchw(exp)))
    assert torch.equal(chk, ref)                                       # the layout really is d2s's inverse
    _, act = engine.dgrad3x3(dy.cuda(), wt.cuda(), mask.cuda(), s2d=True)
    assert tuple(act.shape) == (2, 16, 24, 256)
    assert (act.cpu().double() - exp).abs().max() < 2e-5 * max(1.0, float(ref.abs().max()))
