"""-m gpu: PWC-Net inference through the C ABI against the torch-CPU restatement (oracle/pwcnet_oracle.py).
PARITY UNPINNED: the reference's PWC-Net copy misses eight modules and its checkpoint, so the oracle restates the published
architecture of model_pwcnet.py and cannot itself be pinned."""
import os
from types import SimpleNamespace

import numpy as np
import pytest
import torch
from PIL import Image

from oracle import pwcnet_oracle as W
from conftest import GOLDEN

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def pwc():
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    from fisr_b200.pwcnet import PWCNet
    net = PWCNet(0)
    yield net
    net.close()


def _pair(n, h, w, seed):
    g = torch.Generator().manual_seed(seed)
    base = torch.rand(n, h // 8 + 2, w // 8 + 2, 3, generator=g)
    img = torch.nn.functional.interpolate(base.permute(0, 3, 1, 2), scale_factor=8, mode="bilinear").permute(0, 2, 3, 1)
    a = img[:, 4:4 + h, 4:4 + w].contiguous()
    b = img[:, 6:6 + h, 1:1 + w].contiguous()             # the same scene shifted by (-3, +2) px
    return a, b


# 256 x 320 and larger: the dilation-16 layer of the context network runs on the tensor cores too (polyphase images of >= 4 x 4)
@pytest.mark.parametrize("n,h,w", [(1, 64, 64), (2, 128, 192), (1, 192, 320), (1, 256, 320), (2, 320, 512)])
def test_forward_matches_oracle(pwc, n, h, w):
    params = W.init_params(5)
    pwc.set_params(params)
    a, b = _pair(n, h, w, seed=h + w)
    taps = {}
    ref = W.forward(W.init_params(5, torch.float64), a.double(), b.double(), taps)
    before = pwc.launch_count
    got = pwc.forward(a.cuda(), b.cuda()).cpu()
    assert pwc.launch_count > before
    assert tuple(got.shape) == (n, h, w, 2)
    for lvl in range(6, 1, -1):                             # level by level: where a mismatch first appears
        f = pwc.debug_flow(lvl, n, h, w)
        r = taps[f"flow{lvl}"].numpy()
        err = np.abs(f - r).max()
        assert err < 2e-4 * max(1.0, np.abs(r).max()), (lvl, err)
    scale = max(1.0, float(ref.abs().max()))
    assert float((got.double() - ref).abs().max()) < 2e-4 * scale


def test_tensor_core_path_matches_cuda_core_path():
    """The same forward with every conv on the CUDA-core kernel (FISR_PWC_UMMA=0), with the undilated stride-1 convs on the tcgen05
    kernel (1), with the dilated ones as polyphase launches too (2) and with those launches / the two pyramids spread over streams (3, the
    default), at a size where every dilation qualifies.  Modes 2 and 3 run the same kernels on the same data: bit-identical."""
    if not torch.cuda.is_available():
        pytest.skip("no CUDA device in this container (GPU tests run under gpurun)")
    from fisr_b200.pwcnet import PWCNet
    params = W.init_params(11)
    a, b = _pair(1, 512, 768, seed=3)
    a, b = a.cuda(), b.cuda()
    flows = {}
    old = os.environ.get("FISR_PWC_UMMA")
    try:
        for mode in ("0", "1", "2", "3"):
            os.environ["FISR_PWC_UMMA"] = mode
            net = PWCNet(0)
            net.set_params(params)
            out = net.forward(a, b).cpu().numpy()
            flows[mode] = [net.debug_flow(lvl, 1, 512, 768) for lvl in range(6, 1, -1)] + [out]
            net.close()
    finally:
        if old is None:
            os.environ.pop("FISR_PWC_UMMA", None)
        else:
            os.environ["FISR_PWC_UMMA"] = old
    for mode in ("1", "2", "3"):
        for i, (x, y) in enumerate(zip(flows[mode], flows["0"])):
            err = float(np.abs(x - y).max())
            assert err < 2e-4 * max(1.0, float(np.abs(y).max())), (mode, i, err)
    for x, y in zip(flows["3"], flows["2"]):
        assert np.array_equal(x, y)


@pytest.mark.parametrize("h,w", [(72, 104), (75, 101), (64, 64)])
def test_driver_pre_and_post_processing_on_the_device(pwc, h, w):
    """fisr_pwc_prepare_pair / fisr_pwc_finish_flow against the host restatement of the reference's driver (skimage resize, uint8
    truncation, padding; crop, scipy Gaussian, resize): the float64 arithmetic is evaluated in numpy's order, so the network input
    is bit-identical and the finished flow agrees to float32 rounding."""
    from fisr_b200 import utils
    rng = np.random.default_rng(h * w)
    yuv = rng.integers(0, 256, (2, h, w, 3), dtype=np.uint8)
    rgb = [utils.YUV2RGB_matlab(f.astype(np.float32)) for f in yuv]
    a, b, hw0 = W.prepare_pair(rgb[0], rgb[1])
    for frames in ([torch.from_numpy(f).cuda() for f in yuv], [torch.from_numpy(np.ascontiguousarray(f, dtype=np.float64)).cuda() for f in rgb]):
        img1, img2 = pwc.prepare_pair(frames[0], frames[1], 2)
        assert tuple(img1.shape) == (2,) + a.shape
        assert np.array_equal(img1[0].cpu().numpy(), a) and np.array_equal(img1[1].cpu().numpy(), b)
        assert np.array_equal(img2[0].cpu().numpy(), b) and np.array_equal(img2[1].cpu().numpy(), a)
    flow = (rng.standard_normal((2,) + a.shape[:2] + (2,)) * 5).astype(np.float32)
    got = pwc.finish_flow(torch.from_numpy(flow).cuda(), hw0, (h, w), 2).cpu().numpy()
    for k in range(2):
        want = W.finish_flow(flow[k], hw0, (h, w))
        assert got[k].shape == want.shape
        assert np.abs(got[k] - want).max() <= 1e-6, np.abs(got[k] - want).max()
        assert np.mean(got[k] == want) > 0.999


def test_flow_sequence_is_the_pairwise_loop(pwc):
    """The pipelined sequence API (frames uploaded once, downloads overlapped) returns exactly what the one-pair call returns."""
    pwc.set_params(W.init_params(4))
    yuv = np.random.default_rng(8).integers(0, 256, (4, 40, 72, 3), dtype=np.uint8)
    seq = [f.copy() for f in pwc.flow_sequence_yuv(iter(yuv))]
    assert len(seq) == 3
    for k in range(3):
        assert np.array_equal(seq[k], pwc.flow_pair_yuv(yuv[k], yuv[k + 1]))


def test_each_building_block(pwc):
    """The TF-specific semantics one by one, on the oracle side against plain formulas (runs without the GPU too)."""
    g = torch.Generator().manual_seed(0)
    x = torch.rand(1, 5, 8, 8, generator=g)
    w = torch.randn(3, 3, 5, 4, generator=g)
    b = torch.zeros(4)
    y = W.conv_same(x, w, b, stride=2)
    assert tuple(y.shape) == (1, 4, 4, 4)
    # stride 2 'same' on an even size: no padding before -> output (0, 0) sees input rows / cols 0..2
    want = (x[0, :, 0:3, 0:3].permute(1, 2, 0).unsqueeze(-1) * w).sum(dim=(0, 1, 2))
    assert torch.allclose(y[0, :, 0, 0], want, atol=1e-5)
    f = torch.zeros(1, 2, 8, 8); f[:, 0] = 1.0              # u = +1: sample one pixel to the right
    img = torch.arange(64.).reshape(1, 1, 8, 8)
    wp = W.dense_image_warp(img, f)
    assert torch.equal(wp[0, 0, :, :7], img[0, 0, :, 1:]) and torch.equal(wp[0, 0, :, 7], img[0, 0, :, 7])
    cv = W.cost_volume(torch.ones(1, 4, 6, 6), torch.ones(1, 4, 6, 6))
    assert tuple(cv.shape) == (1, 81, 6, 6) and float(cv[0, 40, 3, 3]) == 1.0 and float(cv[0, 0, 0, 0]) == 0.0
    up = W.resize_bilinear_legacy(torch.tensor([[[[0., 4.], [8., 12.]]]]), 4)
    assert up[0, 0, 0, :5].tolist() == [0., 1., 2., 3., 4.] and float(up[0, 0, 7, 7]) == 12.0


def test_compute_flow_driver_end_to_end(pwc, tmp_path):
    """FISR_for_video_Compute_Flow: YUV frames -> RGB -> x2 resize -> uint8 -> network -> crop -> anti-aliased x1/2 -> .flo, against
    the same pipeline on the oracle network."""
    from fisr_b200 import utils
    from fisr_b200.video import FISR_for_video_Compute_Flow
    frames = np.load(os.path.join(GOLDEN, "scene1_lr_crop.npz"))["frames"][:3, :72, :104]     # 72 x 104: x2 = 144 x 208 -> padded to 192 x 256
    folder = tmp_path / "scene"
    os.makedirs(folder)
    for i, f in enumerate(frames):
        Image.fromarray(np.ascontiguousarray(f)).save(str(folder / f"LR_{i}.png"))
    params = W.init_params(9)
    pwc.set_params(params)
    args = SimpleNamespace(frame_folder_path=str(folder), FISR_input_size=(72, 104), frame_num=3)
    path = FISR_for_video_Compute_Flow(args, pwcnet=pwc)
    flow = utils.read_flo_file_5dim(path)
    assert flow.shape == (2, 2, 72, 104, 2)
    rgb = [utils.YUV2RGB_matlab(f.astype(np.float32)) for f in frames]
    a, b, hw0 = W.prepare_pair(rgb[0], rgb[1])
    assert a.shape == (192, 256, 3) and hw0 == (144, 208)
    ref = W.forward(params, torch.from_numpy(np.stack([a, b])), torch.from_numpy(np.stack([b, a]))).numpy()
    for k in range(2):
        want = W.finish_flow(ref[k], hw0, (72, 104))
        assert np.abs(flow[0, k] - want).max() < 2e-4 * max(1.0, np.abs(want).max())
