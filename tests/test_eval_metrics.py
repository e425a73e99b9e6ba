"""CPU: the evaluation metrics of ``FISRnet.test`` (reference FISRnet.py:886-935, utils.py:23-26): PSNR and the tile SSIM."""
import numpy as np
import pytest

from fisr_b200 import utils


def test_psnr_known_answer():
    a = np.zeros((8, 8, 3)); b = np.full((8, 8, 3), 0.1)
    assert utils._compute_psnr(a, b, 1.) == pytest.approx(20.0, abs=1e-9)          # mse 0.01 -> 10 log10(1 / 0.01)


def test_ssim_identity_symmetry_and_bounds():
    rng = np.random.default_rng(0)
    a = rng.integers(0, 256, (50, 71, 3), dtype=np.uint8)
    b = np.clip(a.astype(int) + rng.integers(-20, 21, a.shape), 0, 255).astype(np.uint8)
    assert utils.compare_ssim(a, a) == pytest.approx(1.0, abs=1e-12)
    s = utils.compare_ssim(a, b)
    assert 0.5 < s < 1.0 and s == pytest.approx(utils.compare_ssim(b, a), abs=1e-12)
    assert utils.compare_ssim(a, 255 - a) < 0.2
    with pytest.raises(AttributeError):
        utils.compare_ssim(a, a[:-1])


def test_ssim_single_tile_hand_computed():
    # one 7x7 tile, one channel: x = ramp 0..48, y = x + 10  ->  equal variances, cov = var, means differ by 10
    x = np.arange(49, dtype=np.float64).reshape(7, 7)
    y = x + 10
    m0, m1, var = x.mean(), y.mean(), x.var()
    c1, c2 = 6.5025, 58.5225
    want = (2 * m0 * m1 + c1) * (2 * var + c2) / ((m0 ** 2 + m1 ** 2 + c1) * (2 * var + c2))
    assert utils.compare_ssim(x.astype(np.uint8), y.astype(np.uint8)) == pytest.approx(want, rel=1e-12)


def test_ssim_tiles_are_non_overlapping_and_border_is_dropped():
    rng = np.random.default_rng(1)
    a = rng.integers(0, 256, (14, 14), dtype=np.uint8)
    b = rng.integers(0, 256, (14, 14), dtype=np.uint8)
    per_tile = [utils.compare_ssim(a[y:y + 7, x:x + 7], b[y:y + 7, x:x + 7]) for y in (0, 7) for x in (0, 7)]
    assert utils.compare_ssim(a, b) == pytest.approx(np.mean(per_tile), rel=1e-12)
    pad_a = np.pad(a, ((0, 3), (0, 5)), constant_values=7)
    pad_b = np.pad(b, ((0, 3), (0, 5)), constant_values=200)
    assert utils.compare_ssim(pad_a, pad_b) == pytest.approx(utils.compare_ssim(a, b), rel=1e-12)
