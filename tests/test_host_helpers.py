"""Host-side helpers around the device paths (no GPU needed): the Gaussian weights handed to fisr_pwc_finish_flow, the colour
constants handed to fisr_pwc_prepare_pair, the threaded PNG writer of the test / FISR_for_video loops."""
import os

import numpy as np
from PIL import Image
from scipy import ndimage as ndi

from fisr_b200 import utils
from oracle import pipeline_oracle as P


def test_gauss_weights_are_scipys_kernel():
    from fisr_b200.pwcnet import _gauss_weights
    for sigma in (0.5, 1.0, 1.5):
        w = _gauss_weights(sigma)
        r = len(w) - 1
        assert r == int(4.0 * sigma + 0.5)
        x = np.zeros(4 * r + 1)
        x[2 * r] = 1.0                                     # impulse response of scipy's filter = its kernel
        k = ndi.gaussian_filter(x, sigma, mode="mirror")
        assert np.array_equal(k[2 * r:3 * r + 1], w) and np.array_equal(k[r:2 * r + 1][::-1], w)
    assert np.array_equal(_gauss_weights(0.0), np.ones(1))


def test_yuv2rgb_constants_reproduce_the_reference_formula():
    k = utils.yuv2rgb_constants()
    assert k.shape == (12,) and k.dtype == np.float64
    yuv = np.random.default_rng(0).integers(0, 256, (5, 7, 3)).astype(np.float32)
    T, off = k[:9].reshape(3, 3), k[9:]
    want = np.clip(np.stack([T[p, 0] * yuv[..., 0] + T[p, 1] * yuv[..., 1] + T[p, 2] * yuv[..., 2] - off[p] for p in range(3)], axis=-1), 0, 255)
    assert np.array_equal(utils.YUV2RGB_matlab(yuv), want)
    assert np.abs(P.yuv2rgb_matlab(yuv) - want).max() < 1e-9          # the oracle's own restatement


def test_frame_writer_writes_what_the_sequential_loop_writes(tmp_path):
    from fisr_b200.FISRnet import _FrameWriter
    rng = np.random.default_rng(1)
    frames = rng.integers(0, 256, (5, 24, 40, 3), dtype=np.uint8)
    w = _FrameWriter(workers=3, max_pending=2)              # fewer slots than frames: save() must block, not drop
    for i, f in enumerate(frames):
        w.save(f, str(tmp_path / f"pred_{i}.png"), str(tmp_path / f"pred_YUV_{i}.png"))
    w.close()
    assert sorted(os.listdir(tmp_path)) == sorted([f"pred_{i}.png" for i in range(5)] + [f"pred_YUV_{i}.png" for i in range(5)])
    for i, f in enumerate(frames):
        assert np.array_equal(np.array(Image.open(tmp_path / f"pred_YUV_{i}.png")), f)
        assert np.array_equal(np.array(Image.open(tmp_path / f"pred_{i}.png")), utils.YUV2RGB_matlab(f).astype("uint8"))


def test_frame_writer_reports_failures(tmp_path):
    from fisr_b200.FISRnet import _FrameWriter
    w = _FrameWriter(workers=1)
    w.save(np.zeros((4, 4, 3), np.uint8), str(tmp_path / "no_such_dir" / "x.png"))
    try:
        w.close()
    except (FileNotFoundError, OSError):
        return
    raise AssertionError("a failed write must surface in close()")
