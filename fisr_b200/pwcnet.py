"""PWC-Net inference on the B200 path: the flow estimator the reference runs in front of its warp
(FISR_tfoptflow/model_pwcnet.py, driven by FISR_for_video_pwcnet_predict_from_img_test.py:84-147).

``PWCNet`` owns one ``fisr_pwc`` context (C ABI); every kernel is in libfisr_b200.so.  The host side mirrors the reference's
driver: YUV -> RGB, x2 ``skimage.transform.resize``, uint8, /255, zero pad to multiples of 64, crop, anti-aliased x1/2 resize, /2
(skimage is restated with numpy / scipy, it is not installable here).  PARITY UNPINNED: see include/fisr_b200.h.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FisrError


def param_inventory() -> "OrderedDict[str, Tuple[int, ...]]":
    lib = _lib.load()
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    dims = (C.c_int * 4)()
    for k in range(lib.fisr_pwc_num_params()):
        rank = lib.fisr_pwc_param_shape(k, dims)
        out[lib.fisr_pwc_param_name(k).decode()] = tuple(dims[j] for j in range(rank))
    return out


class PWCNet:
    def __init__(self, device: int = 0):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise FisrError("fisr_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.fisr_pwc_create(self.device, C.byref(h))
        if rc != 0:
            raise FisrError(f"fisr_pwc_create failed ({rc}): {self.lib.fisr_pwc_last_error(None).decode()}")
        self.h = h
        # library work runs on this side stream, ordered against torch's current stream on entry and exit (torch's default
        # stream has handle 0, which the C ABI reads as "the context's own stream")
        self.stream = torch.cuda.Stream(self.device)

    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise FisrError(f"{what} failed ({rc}): {self.lib.fisr_pwc_last_error(self.h).decode()}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.fisr_pwc_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    @property
    def launch_count(self) -> int:
        return int(self.lib.fisr_pwc_launch_count(self.h))

    def set_params(self, params: Dict[str, "np.ndarray | torch.Tensor"]) -> None:
        for name, shape in param_inventory().items():
            if name not in params:
                raise FisrError(f"missing PWC-Net parameter {name}")
            a = params[name]
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            a = np.ascontiguousarray(a, dtype=np.float32)
            if tuple(a.shape) != shape:
                raise FisrError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
            self._check(self.lib.fisr_pwc_set_param(self.h, name.encode(), a.ctypes.data, a.size), f"fisr_pwc_set_param({name})")

    def load_checkpoint(self, prefix: str) -> None:
        """A tfoptflow checkpoint (TensorFlow V2 bundle, e.g. ``pwcnet.ckpt-595000``) read without TensorFlow."""
        from .tf_checkpoint import load_checkpoint
        names = list(param_inventory())
        got = load_checkpoint(prefix, names, verify_crc=True)
        self.set_params({n: got[n] for n in names})

    def forward(self, img1: torch.Tensor, img2: torch.Tensor) -> torch.Tensor:
        """``nn()`` (model_pwcnet.py:1525-1593): img1, img2 f32 [N,H,W,3] in 0..1 on the GPU, H, W multiples of 64 -> flow [N,H,W,2]."""
        for t in (img1, img2):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 4 and t.shape[3] == 3):
                raise FisrError("PWCNet.forward takes contiguous float32 CUDA tensors [N,H,W,3]")
        n, h, w, _ = img1.shape
        out = torch.empty((n, h, w, 2), dtype=torch.float32, device=img1.device)
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        for t in (img1, img2, out):
            t.record_stream(self.stream)
        self._check(self.lib.fisr_pwc_forward(self.h, img1.data_ptr(), img2.data_ptr(), n, h, w, out.data_ptr(), self.stream.cuda_stream),
                    "fisr_pwc_forward")
        cur.wait_stream(self.stream)
        return out

    def debug_flow(self, lvl: int, n: int, h: int, w: int) -> np.ndarray:
        a = np.empty((n, h >> lvl, w >> lvl, 2), np.float32)
        self._check(self.lib.fisr_pwc_debug_flow(self.h, lvl, a.ctypes.data, a.size), "fisr_pwc_debug_flow")
        return a

    # ------------------------------------------------------------------ the reference's driver around the network
    def flow_pair(self, rgb1: np.ndarray, rgb2: np.ndarray, scale: int = 2) -> np.ndarray:
        """Bidirectional flow of one frame pair as ..predict_from_img_test.py:126-138 computes it: rgb float [h,w,3] in 0..255 ->
        float32 [2,h,w,2] (1 -> 2, 2 -> 1) at the input resolution."""
        h, w = rgb1.shape[:2]
        a, b, hw0 = prepare_pair(rgb1, rgb2, scale)
        dev = torch.device("cuda", self.device)
        ta, tb = torch.from_numpy(a).to(dev), torch.from_numpy(b).to(dev)
        flow = self.forward(torch.stack([ta, tb]), torch.stack([tb, ta])).cpu().numpy()       # both directions as one batch
        return np.stack([finish_flow(flow[k], hw0, (h, w), scale) for k in range(2)])


# ---------------------------------------------------------------------------------------- skimage.transform.resize, restated
def skimage_resize(img: np.ndarray, out_hw: Tuple[int, int], anti_aliasing: bool = False) -> np.ndarray:
    """``skimage.transform.resize`` along axes (-3, -2): order 1, mode 'reflect' (scipy 'mirror'), pixel-centre-aligned
    coordinates, Gaussian pre-filter of sigma (factor - 1) / 2 per down-scaled axis when ``anti_aliasing``."""
    from scipy import ndimage as ndi
    img = np.asarray(img, dtype=np.float64)
    h, w = img.shape[-3], img.shape[-2]
    oh, ow = out_hw
    if anti_aliasing:
        sig = [0.0] * img.ndim
        sig[-3], sig[-2] = max(0.0, (h / oh - 1) / 2), max(0.0, (w / ow - 1) / 2)
        if any(sig):
            img = ndi.gaussian_filter(img, sig, mode="mirror")

    def lerp_axis(a, coords, axis):
        n = a.shape[axis]
        period = 2 * (n - 1) if n > 1 else 1
        i0 = np.floor(coords).astype(np.int64)
        fr = coords - i0

        def mirror(i):
            if n == 1:
                return np.zeros_like(i)
            i = np.mod(i, period)
            return np.where(i >= n, period - i, i)

        shape = [1] * a.ndim
        shape[axis] = -1
        fr = fr.reshape(shape)
        return np.take(a, mirror(i0), axis=axis) * (1 - fr) + np.take(a, mirror(i0 + 1), axis=axis) * fr

    ys = (np.arange(oh) + 0.5) * (h / oh) - 0.5
    xs = (np.arange(ow) + 0.5) * (w / ow) - 0.5
    return lerp_axis(lerp_axis(img, ys, img.ndim - 3), xs, img.ndim - 2)


def prepare_pair(rgb1: np.ndarray, rgb2: np.ndarray, scale: int = 2):
    """..predict_from_img_test.py:126-131 + adapt_x (model_pwcnet.py:371-409): x`scale` resize, uint8 truncation, /255, zero pad
    to multiples of 64.  Returns (img1, img2) float32 [H,W,3] and the unpadded size."""
    h, w = rgb1.shape[:2]
    outs = []
    for a in (rgb1, rgb2):
        u8 = np.array(skimage_resize(a, (h * scale, w * scale)), dtype=np.uint8)
        x = u8.astype(np.float32) / np.float32(255.)
        outs.append(np.ascontiguousarray(np.pad(x, [(0, (-x.shape[0]) % 64), (0, (-x.shape[1]) % 64), (0, 0)], mode="constant")))
    return outs[0], outs[1], (h * scale, w * scale)


def finish_flow(flow: np.ndarray, hw0: Tuple[int, int], out_hw: Tuple[int, int], scale: int = 2) -> np.ndarray:
    """postproc_y_hat_test crop (model_pwcnet.py:449-470) + ..predict_from_img_test.py:137: anti-aliased resize, / scale."""
    return (skimage_resize(flow[:hw0[0], :hw0[1]], out_hw, anti_aliasing=True) / scale).astype(np.float32)
