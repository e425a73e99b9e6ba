"""Oracle (test infrastructure): the callers either side of ``FISRnet.model``.

Restates, with numpy/torch on the CPU:
  * window assembly            -- ``ops.py:90-116`` + ``FISRnet.py:281-306``
  * tile geometry              -- ``utils.py:118-159`` (get_HW_boundary / trim_patch_boundary)
  * tiled window inference     -- ``FISRnet.py:1003-1065`` (crop, normalise, tile loop, paste, clip, uint8)
  * colour conversion          -- ``utils.py:106-115`` (YUV2RGB_matlab), warp-side ``YUV2RGB`` / ``RGB2YUV``
                                  ``FISR_tfoptflow/FISR_for_video_warp_img_with_flo.py:35-57``
  * flow warp                  -- ``warp_flow`` ``..warp_img_with_flo.py:61-67`` (cv2.remap itself, plus a
                                  numpy restatement of OpenCV's fixed-point bilinear remap)
"""
from __future__ import annotations

from typing import Callable, Dict, List, Tuple

import numpy as np
import torch

from . import fisrnet_oracle as net

PATCH_BOUNDARY = 32     # FISRnet.py:779,984


# ------------------------------------------------------------------ window assembly
def window_input(data: torch.Tensor, flow: torch.Tensor, warp: torch.Tensor, i: int) -> torch.Tensor:
    """Stride-1 window ``i`` of a 5-frame sample (``FISRnet.py:281-306``).

    data [B,h,w,15], flow [B,h,w,16], warp [B,h,w,24] -> [B,h,w,29]:
    frames ch [3i,3i+9) (``Tensor_slicer_recurrent`` ops.py:90-96), flows ch [4i,4i+8) (ops.py:99-106),
    warped frames ch [6i,6i+12) (ops.py:109-116)."""
    return torch.cat([data[..., 3 * i:3 * i + 9], flow[..., 4 * i:4 * i + 8], warp[..., 6 * i:6 * i + 12]], dim=3)


def groups2ovlp(g: torch.Tensor) -> torch.Tensor:
    """``Groups2Ovlp`` ops.py:119-144: [B,9,h,w,3] -> [B,7,h,w,3] = [f0,f1,(f2+f3)/2,f4,(f5+f6)/2,f7,f8]."""
    return torch.stack([g[:, 0], g[:, 1], (g[:, 2] + g[:, 3]) / 2, g[:, 4], (g[:, 5] + g[:, 6]) / 2, g[:, 7], g[:, 8]],
                       dim=1)


def split_seq_dim(x: torch.Tensor) -> torch.Tensor:
    """``tf_split_seq_dim`` ops.py:155-160: [N,H,W,3*S] -> [N,S,H,W,3]."""
    n, h, w, c = x.shape
    return x.reshape(n, h, w, c // 3, 3).permute(0, 3, 1, 2, 4)


# ------------------------------------------------------------------ tile geometry
def get_hw_boundary(pb: int, h: int, w: int, pH: int, sH: int, pW: int, sW: int):
    """``utils.get_HW_boundary`` utils.py:118-135."""
    h_lo = max(pH * sH - pb, 0)
    h_hi = min((pH + 1) * sH + pb, h)
    w_lo = max(pW * sW - pb, 0)
    w_hi = min((pW + 1) * sW + pb, w)
    add_h = (pb if pH * sH >= pb else 0) + (pb if (pH + 1) * sH + pb <= h else 0)
    add_w = (pb if pW * sW >= pb else 0) + (pb if (pW + 1) * sW + pb <= w else 0)
    return h_lo, h_hi, w_lo, w_hi, add_h, add_w


def trim_patch_boundary(img: np.ndarray, pb: int, h: int, w: int, pH: int, sH: int, pW: int, sW: int, sf: int):
    """``utils.trim_patch_boundary`` utils.py:138-159 (img is [1,H,W,C])."""
    if pb == 0:
        return img
    if not pH * sH < pb:
        img = img[:, pb * sf:, :, :]
    if not (pH + 1) * sH + pb > h:
        img = img[:, :-pb * sf, :, :]
    if not pW * sW < pb:
        img = img[:, :, pb * sf:, :]
    if not (pW + 1) * sW + pb > w:
        img = img[:, :, :-pb * sf, :]
    return img


def crop_hw(H: int, W: int, num_patch: Tuple[int, int]) -> Tuple[int, int]:
    """``FISRnet.py:1006-1007``: crop so that every tile side is a multiple of 32."""
    return H - H % (32 * num_patch[0]), W - W % (32 * num_patch[1])


# ------------------------------------------------------------------ tiled window inference
def normalise_window(frames_u8: np.ndarray, flow: np.ndarray, warp: np.ndarray, h: int, w: int) -> np.ndarray:
    """``FISRnet.py:1008-1024``: frames uint8 [H,W,9]; flow f32 [H,W,8] in LR pixels; warp f32 [H,W,12]
    already /255 (``utils.read_mat_file_warp`` utils.py:51).  Returns float64 [1,h,w,29] like the reference."""
    img = np.clip(np.array(frames_u8[:h, :w, :], dtype=np.double) / 255., 0, 1)
    fl = np.clip(flow[:h, :w, :] / 96 / 2, -1, 1)
    wp = np.clip(warp[:h, :w, :], 0, 1)
    return np.concatenate([img, fl, wp], axis=2)[None]


def tiled_window(model_fn: Callable[[np.ndarray], np.ndarray], inp: np.ndarray, num_patch=(2, 2), sf: int = 2) -> np.ndarray:
    """Tile loop of ``FISRnet.py:1025-1057``.  ``model_fn`` maps a [1,th,tw,29] array to pred_l3 [1,2th,2tw,9].
    Returns the float64 [h*sf, w*sf, 9] canvas before clipping."""
    _, h, w, _ = inp.shape
    full = np.zeros((h * sf, w * sf, 9))
    for p in range(num_patch[0] * num_patch[1]):
        pH, pW = p // num_patch[1], p % num_patch[1]
        sH, sW = h // num_patch[0], w // num_patch[1]
        h_lo, h_hi, w_lo, w_hi, _, _ = get_hw_boundary(PATCH_BOUNDARY, h, w, pH, sH, pW, sW)
        pred = model_fn(inp[:, h_lo:h_hi, w_lo:w_hi, :])
        trim = trim_patch_boundary(pred, PATCH_BOUNDARY, h, w, pH, sH, pW, sW, sf)
        full[pH * sH * sf:(pH + 1) * sH * sf, pW * sW * sf:(pW + 1) * sW * sf, :] = np.squeeze(trim, 0)
    return full


def quantise(full: np.ndarray) -> np.ndarray:
    """``FISRnet.py:1060-1064``: clip to [0,1], ``np.uint8(x*255)`` (truncation)."""
    return np.uint8(np.clip(full, 0, 1) * 255)


def window_forward_u8(params: Dict[str, torch.Tensor], frames_u8, flow, warp, num_patch=(2, 2)) -> np.ndarray:
    """One window of ``FISR_for_video`` end to end on the oracle network: -> uint8 [2h,2w,9] (YUV)."""
    H, W = frames_u8.shape[:2]
    h, w = crop_hw(H, W, num_patch)
    inp = normalise_window(frames_u8, flow, warp, h, w)

    def fn(tile):
        return net.model(params, torch.from_numpy(tile.astype(np.float32)))[2].numpy().astype(np.float64)

    return quantise(tiled_window(fn, inp, num_patch))


# ------------------------------------------------------------------ colour conversion
_TINV = np.array([[0.00456621, 0., 0.00625893], [0.00456621, -0.00153632, -0.00318811], [0.00456621, 0.00791071, 0.]])


def yuv2rgb_matlab(yuv: np.ndarray) -> np.ndarray:
    """``utils.YUV2RGB_matlab`` utils.py:106-115 (float64, clipped to [0,255], not rounded)."""
    T = 255 * _TINV
    off = T @ np.array([[16.], [128.], [128.]])
    rgb = np.zeros(yuv.shape)
    for p in range(3):
        rgb[:, :, p] = T[p, 0] * yuv[:, :, 0] + T[p, 1] * yuv[:, :, 1] + T[p, 2] * yuv[:, :, 2] - off[p]
    return np.clip(rgb, 0, 255)


def rgb2yuv(rgb: np.ndarray) -> np.ndarray:
    """``RGB2YUV`` ..warp_img_with_flo.py:48-57."""
    T = np.array([[65.481, 128.553, 24.966], [-37.797, -74.203, 112], [112, -93.786, -18.214]]) / 255
    off = [16., 128., 128.]
    yuv = np.zeros(rgb.shape)
    for p in range(3):
        yuv[:, :, p] = T[p, 0] * rgb[:, :, 0] + T[p, 1] * rgb[:, :, 1] + T[p, 2] * rgb[:, :, 2] + off[p]
    return np.clip(yuv, 0, 255)


# ------------------------------------------------------------------ flow warp
def warp_flow_cv2(img: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """``warp_flow`` ..warp_img_with_flo.py:61-67, calling OpenCV exactly as the reference does.
    img f32 [h,w,3]; flow f32 [h,w,2] (already scaled by 0.5 by the caller, :123,127).  Does not mutate ``flow``."""
    import cv2
    h, w = flow.shape[:2]
    m = flow.astype(np.float32).copy()
    m[:, :, 0] += np.arange(w)
    m[:, :, 1] += np.arange(h)[:, np.newaxis]
    return cv2.remap(img, m, None, cv2.INTER_LINEAR, None, cv2.BORDER_REPLICATE)


def warp_flow_fixedpoint(img: np.ndarray, flow: np.ndarray) -> np.ndarray:
    """numpy restatement of what ``cv2.remap(INTER_LINEAR, BORDER_REPLICATE)`` computes for float images
    (OpenCV semantics, not in the reference tree): map coordinates are converted to fixed point with
    INTER_BITS = 5 fractional bits by ``cvRound(x * 32)`` (round-half-even, saturating), the four taps are
    weighted with the 32x32 float table ``(1-fy)(1-fx), (1-fy)fx, fy(1-fx), fy*fx`` and border taps are
    index-clamped."""
    h, w = flow.shape[:2]
    mx = flow[:, :, 0].astype(np.float32) + np.arange(w, dtype=np.float32)
    my = flow[:, :, 1].astype(np.float32) + np.arange(h, dtype=np.float32)[:, None]
    ix = np.rint(mx * np.float32(32)).astype(np.int64)
    iy = np.rint(my * np.float32(32)).astype(np.int64)
    sx, sy = ix >> 5, iy >> 5
    fx = (ix & 31).astype(np.float32) / np.float32(32)
    fy = (iy & 31).astype(np.float32) / np.float32(32)
    x0, x1 = np.clip(sx, 0, w - 1), np.clip(sx + 1, 0, w - 1)
    y0, y1 = np.clip(sy, 0, h - 1), np.clip(sy + 1, 0, h - 1)
    w00 = ((1 - fy) * (1 - fx))[..., None]
    w01 = ((1 - fy) * fx)[..., None]
    w10 = (fy * (1 - fx))[..., None]
    w11 = (fy * fx)[..., None]
    img = img.astype(np.float32)
    return img[y0, x0] * w00 + img[y0, x1] * w01 + img[y1, x0] * w10 + img[y1, x1] * w11


def warp_pair_yuv(yuv1_u8: np.ndarray, yuv2_u8: np.ndarray, flow12: np.ndarray, flow21: np.ndarray,
                  warp_fn=warp_flow_cv2) -> np.ndarray:
    """One frame pair of ``FISR_for_video_Warp_Img`` (..warp_img_with_flo.py:112-128).
    Returns float32 [2,h,w,3] in 0..255 (slot 0: frame 2 sampled with 0.5*flow(1->2); slot 1: frame 1 with
    0.5*flow(2->1)), exactly what is written to the ``.mat`` (:121-128,134)."""
    rgb1 = yuv2rgb_matlab(np.array(yuv1_u8, dtype=np.float32))
    rgb2 = yuv2rgb_matlab(np.array(yuv2_u8, dtype=np.float32))
    out = np.zeros((2,) + yuv1_u8.shape, dtype=np.float32)
    out[0] = rgb2yuv(warp_fn(rgb2, flow12 * 0.5))
    out[1] = rgb2yuv(warp_fn(rgb1, flow21 * 0.5))
    return out
