"""PWC-Net inference on the B200 path: the flow estimator the reference runs in front of its warp
(FISR_tfoptflow/model_pwcnet.py, driven by FISR_for_video_pwcnet_predict_from_img_test.py:84-147).

``PWCNet`` owns one ``fisr_pwc`` context (C ABI); every kernel is in libfisr_b200.so, including the reference driver's pre- and
post-processing (YUV -> RGB, x2 ``skimage.transform.resize``, uint8, /255, zero pad to multiples of 64; crop, anti-aliased x1/2
resize, /2), which runs in float64 on the device in numpy's evaluation order.  PARITY UNPINNED: see include/fisr_b200.h.
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

    def forward(self, img1: torch.Tensor, img2: torch.Tensor, out: "torch.Tensor | None" = None) -> torch.Tensor:
        """``nn()`` (model_pwcnet.py:1525-1593): img1, img2 f32 [N,H,W,3] in 0..1 on the GPU, H, W multiples of 64 -> flow [N,H,W,2]."""
        for t in (img1, img2):
            if not (t.is_cuda and t.dtype == torch.float32 and t.is_contiguous() and t.dim() == 4 and t.shape[3] == 3):
                raise FisrError("PWCNet.forward takes contiguous float32 CUDA tensors [N,H,W,3]")
        n, h, w, _ = img1.shape
        if out is None:
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
    def _run(self, tensors, call, what):
        """One library call on the side stream, ordered against torch's current stream on entry and exit."""
        cur = torch.cuda.current_stream(self.device)
        self.stream.wait_stream(cur)
        for t in tensors:
            t.record_stream(self.stream)
        self._check(call(self.stream.cuda_stream), what)
        cur.wait_stream(self.stream)

    def prepare_pair(self, f1: torch.Tensor, f2: torch.Tensor, scale: int = 2, out=None):
        """..predict_from_img_test.py:113-131 + adapt_x on the device: two frames [h,w,3] (uint8 YUV, or float64 RGB in 0..255) ->
        (img1, img2) f32 [2,Hp,Wp,3] = the batch of both directions, ready for ``forward``."""
        from .utils import yuv2rgb_constants
        h, w = int(f1.shape[0]), int(f1.shape[1])
        kind = {torch.uint8: 0, torch.float64: 1}.get(f1.dtype)
        if kind is None or f2.dtype != f1.dtype or tuple(f1.shape) != (h, w, 3) or tuple(f2.shape) != (h, w, 3) or scale < 1 \
                or not (f1.is_cuda and f2.is_cuda and f1.is_contiguous() and f2.is_contiguous()):
            raise FisrError(f"prepare_pair: frames {tuple(f1.shape)} {f1.dtype} / {tuple(f2.shape)} {f2.dtype}, scale {scale}")
        Hp, Wp = -(-h * scale // 64) * 64, -(-w * scale // 64) * 64
        if out is None:
            img1 = torch.empty((2, Hp, Wp, 3), dtype=torch.float32, device=f1.device)
            img2 = torch.empty_like(img1)
        else:
            img1, img2 = out
        k = np.ascontiguousarray(yuv2rgb_constants(), dtype=np.float64)
        self._run((f1, f2, img1, img2), lambda st: self.lib.fisr_pwc_prepare_pair(
            self.h, f1.data_ptr(), f2.data_ptr(), kind, k.ctypes.data, h, w, scale, img1.data_ptr(), img2.data_ptr(), st), "fisr_pwc_prepare_pair")
        return img1, img2

    def finish_flow(self, flow: torch.Tensor, hw0: Tuple[int, int], out_hw: Tuple[int, int], scale: int = 2,
                    out: "torch.Tensor | None" = None) -> torch.Tensor:
        """postproc_y_hat_test crop + anti-aliased resize + / scale (..predict_from_img_test.py:137) on the device:
        flow f32 [N,Hp,Wp,2] -> [N,h,w,2]."""
        n, Hp, Wp, _ = flow.shape
        (h0, w0), (h, w) = hw0, out_hw
        if not (flow.is_cuda and flow.dtype == torch.float32 and flow.is_contiguous() and flow.shape[3] == 2):
            raise FisrError("finish_flow takes a contiguous float32 CUDA tensor [N,Hp,Wp,2]")
        wy, wx = _gauss_weights(max(0.0, (h0 / h - 1) / 2)), _gauss_weights(max(0.0, (w0 / w - 1) / 2))
        if out is None:
            out = torch.empty((n, h, w, 2), dtype=torch.float32, device=flow.device)
        self._run((flow, out), lambda st: self.lib.fisr_pwc_finish_flow(
            self.h, flow.data_ptr(), n, Hp, Wp, h0, w0, h, w, wy.ctypes.data, len(wy) - 1, wx.ctypes.data, len(wx) - 1, float(scale),
            out.data_ptr(), st), "fisr_pwc_finish_flow")
        return out

    def _flow_pair(self, f1: torch.Tensor, f2: torch.Tensor, scale: int) -> np.ndarray:
        """Pre-processing, network and post-processing of one frame pair on the device: only the two frames go up and the
        [2,h,w,2] flow comes down."""
        h, w = int(f1.shape[0]), int(f1.shape[1])
        img1, img2 = self.prepare_pair(f1, f2, scale)
        flow = self.forward(img1, img2)
        return self.finish_flow(flow, (h * scale, w * scale), (h, w), scale).cpu().numpy()

    def flow_sequence_yuv(self, frames, scale: int = 2):
        """Bidirectional flow of every adjacent pair of a frame sequence (the loop of ..predict_from_img_test.py:111-139): ``frames``
        yields uint8 YUV frames [h,w,3]; yields float32 [2,h,w,2] per pair (valid until the next item is requested).  Every frame
        is uploaded once, and the download of pair k overlaps the kernels of pair k + 1 (pinned buffers, one pair in flight)."""
        dev = torch.device("cuda", self.device)
        prev, pending, k, work = None, None, 0, None
        hosts = [None, None]
        stage, staged, j = [None, None], [None, None], 0          # pinned upload buffers, reused every other frame
        for f in frames:
            if f.dtype != np.uint8:
                raise FisrError("flow_sequence_yuv takes uint8 frames")
            if stage[j] is None or tuple(stage[j].shape) != tuple(f.shape):
                stage[j] = torch.empty(tuple(f.shape), dtype=torch.uint8, pin_memory=True)
            if staged[j] is not None:
                staged[j].synchronize()                             # the upload that last used this buffer (two frames ago)
            stage[j].numpy()[...] = f
            cur = stage[j].to(dev, non_blocking=True)
            staged[j] = torch.cuda.Event()
            staged[j].record()
            j ^= 1
            if prev is not None:
                h, w = int(cur.shape[0]), int(cur.shape[1])
                if work is None or work[3].shape[1:3] != (h, w):      # device buffers of the whole sequence (every call is stream ordered)
                    Hp, Wp = -(-h * scale // 64) * 64, -(-w * scale // 64) * 64
                    work = [torch.empty((2, Hp, Wp, 3), dtype=torch.float32, device=dev), torch.empty((2, Hp, Wp, 3), dtype=torch.float32, device=dev),
                            torch.empty((2, Hp, Wp, 2), dtype=torch.float32, device=dev), torch.empty((2, h, w, 2), dtype=torch.float32, device=dev)]
                img1, img2 = self.prepare_pair(prev, cur, scale, out=(work[0], work[1]))
                out = self.finish_flow(self.forward(img1, img2, out=work[2]), (h * scale, w * scale), (h, w), scale, out=work[3])
                if hosts[k] is None or hosts[k].shape != out.shape:
                    hosts[k] = torch.empty(out.shape, dtype=out.dtype, pin_memory=True)
                hosts[k].copy_(out, non_blocking=True)
                ev = torch.cuda.Event()
                ev.record()
                if pending is not None:
                    pending[1].synchronize()
                    yield pending[0].numpy()
                pending = (hosts[k], ev)
                k ^= 1
            prev = cur
        if pending is not None:
            pending[1].synchronize()
            yield pending[0].numpy()

    def flow_pair(self, rgb1: np.ndarray, rgb2: np.ndarray, scale: int = 2) -> np.ndarray:
        """Bidirectional flow of one frame pair as ..predict_from_img_test.py:126-138 computes it: rgb float [h,w,3] in 0..255 ->
        float32 [2,h,w,2] (1 -> 2, 2 -> 1) at the input resolution."""
        dev = torch.device("cuda", self.device)
        f1 = torch.from_numpy(np.ascontiguousarray(rgb1, dtype=np.float64)).to(dev)
        f2 = torch.from_numpy(np.ascontiguousarray(rgb2, dtype=np.float64)).to(dev)
        return self._flow_pair(f1, f2, scale)

    def flow_pair_yuv(self, yuv1: np.ndarray, yuv2: np.ndarray, scale: int = 2) -> np.ndarray:
        """The same from the uint8 YUV frames the reference's driver reads (..predict_from_img_test.py:113-120): the YUV -> RGB
        conversion of utils.YUV2RGB_matlab runs on the device too."""
        if yuv1.dtype != np.uint8 or yuv2.dtype != np.uint8:
            raise FisrError("flow_pair_yuv takes uint8 frames")
        dev = torch.device("cuda", self.device)
        f1 = torch.from_numpy(np.ascontiguousarray(yuv1)).to(dev)
        f2 = torch.from_numpy(np.ascontiguousarray(yuv2)).to(dev)
        return self._flow_pair(f1, f2, scale)


def _gauss_weights(sigma: float) -> np.ndarray:
    """Weights at distance 0..radius of the kernel ``scipy.ndimage.gaussian_filter`` builds for ``sigma`` (truncate = 4,
    _gaussian_kernel1d); a single 1 when there is nothing to filter."""
    if sigma <= 1e-15:
        return np.ones(1, dtype=np.float64)
    radius = int(4.0 * float(sigma) + 0.5)
    x = np.arange(-radius, radius + 1)
    phi = np.exp(-0.5 / (sigma * sigma) * x ** 2)
    phi = phi / phi.sum()
    return np.ascontiguousarray(phi[radius:], dtype=np.float64)
