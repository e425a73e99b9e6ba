"""The algebraic identities pwc_api.cu relies on to run every PWC-Net conv as a stride-1, dilation-1 3x3 conv on the tensor cores,
checked on the CPU with torch: same index formulas as build_fused / build_stride2 / build_packed4 and the polyphase launches.
(The product code itself is checked on the GPU against the oracle network, tests/test_gpu_pwcnet.py.)"""
import torch
import torch.nn.functional as F

torch.manual_seed(0)


def conv3x3(x, w, dil=1):
    """x [N,H,W,C], w [3,3,Cin,Cout] (HWIO), stride 1, zero 'same' padding -> [N,H,W,Cout]."""
    return F.conv2d(x.permute(0, 3, 1, 2), w.permute(3, 2, 0, 1), padding=dil, dilation=dil).permute(0, 2, 3, 1)


def test_dilated_conv_is_undilated_convs_on_polyphase_images():
    x = torch.randn(1, 16, 24, 5, dtype=torch.float64)
    w = torch.randn(3, 3, 5, 4, dtype=torch.float64)
    for d in (2, 4, 8):
        want = conv3x3(x, w, dil=d)
        got = torch.empty_like(want)
        for py in range(d):
            for px in range(d):
                got[:, py::d, px::d] = conv3x3(x[:, py::d, px::d], w)          # zero padding of the sub-image = zero padding at distance d
        assert torch.allclose(got, want, atol=1e-12)


def test_stride2_conv_is_two_row_phase_convs_on_superpixels():
    n, H, W, C, Co = 1, 12, 16, 3, 4
    x = torch.randn(n, H, W, C, dtype=torch.float64)
    w = torch.randn(3, 3, C, Co, dtype=torch.float64)
    # TF 'same', stride 2, even size: pad 0 before, 1 after -> taps x[2o + k]
    xp = F.pad(x.permute(0, 3, 1, 2), (0, 1, 0, 1))
    want = F.conv2d(xp, w.permute(3, 2, 0, 1), stride=2).permute(0, 2, 3, 1)
    got = torch.zeros_like(want)
    for py in range(2):
        view = x[:, py::2].reshape(n, H // 2, W // 2, 2 * C)                    # super-pixel X: channels of pixels 2X and 2X + 1
        wv = torch.zeros(3, 3, 2 * C, Co, dtype=torch.float64)
        for ky in range(py, 3, 2):
            for kx in range(3):
                ty, tx, half = ky // 2 + 1, kx // 2 + 1, kx & 1
                wv[ty, tx, half * C:(half + 1) * C] = w[ky, kx]
        got += conv3x3(view, wv)
    assert torch.allclose(got, want, atol=1e-12)


def test_transposed_conv_4x4_s2_is_a_3x3_conv_with_subpixel_columns():
    n, h, w_, C = 1, 6, 7, 5
    x = torch.randn(n, h, w_, C, dtype=torch.float64)
    wt = torch.randn(4, 4, 2, C, dtype=torch.float64)                          # TF conv2d_transpose kernel [4,4,out,in]
    # out[2i + k - 1] += in[i] w[k]  ('same', stride 2): torch's ConvTranspose2d with padding 1
    want = F.conv_transpose2d(x.permute(0, 3, 1, 2), wt.permute(3, 2, 0, 1), stride=2, padding=1).permute(0, 2, 3, 1)
    tap = [[3, 1, -1], [-1, 2, 0]]                                             # [sub-pixel parity][dy + 1] -> transposed-conv tap
    w3 = torch.zeros(3, 3, C, 8, dtype=torch.float64)
    for sa in range(2):
        for sb in range(2):
            for dy in range(3):
                for dx in range(3):
                    ky, kx = tap[sa][dy], tap[sb][dx]
                    if ky < 0 or kx < 0:
                        continue
                    for co in range(2):
                        w3[dy, dx, :, (2 * sa + sb) * 2 + co] = wt[ky, kx, co]
    cols = conv3x3(x, w3)                                                      # [n,h,w,8]: column (2a + b) * 2 + co
    got = cols.reshape(n, h, w_, 2, 2, 2).permute(0, 1, 3, 2, 4, 5).reshape(n, 2 * h, 2 * w_, 2)
    assert torch.allclose(got, want, atol=1e-12)


def test_narrow_conv_on_4_pixel_superpixels():
    n, H, W, C = 1, 6, 16, 16
    x = torch.randn(n, H, W, C, dtype=torch.float64)
    w = torch.randn(3, 3, C, C, dtype=torch.float64)
    want = conv3x3(x, w)
    wp = torch.zeros(3, 3, 4 * C, 4 * C, dtype=torch.float64)
    for ky in range(3):
        for kx in range(3):
            for q in range(4):
                s = q + kx - 1
                dx = -1 if s < 0 else (1 if s > 3 else 0)
                qi = s - 4 * dx
                wp[ky, dx + 1, qi * C:(qi + 1) * C, q * C:(q + 1) * C] = w[ky, kx]
    got = conv3x3(x.reshape(n, H, W // 4, 4 * C), wp).reshape(n, H, W, C)
    assert torch.allclose(got, want, atol=1e-12)
