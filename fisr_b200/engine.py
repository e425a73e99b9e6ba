"""Device context of the B200 FISRnet path: owns one ``fisr_ctx`` (C ABI) on one GPU.

PyTorch is used only as the container of device memory and for the current CUDA stream;
every kernel that runs is in ``libfisr_b200.so``.
"""
from __future__ import annotations

import ctypes as C
from collections import OrderedDict
from typing import Dict, Optional, Tuple

import numpy as np
import torch

from . import _lib
from ._lib import FisrError, PREC_F16, PREC_F16F8, PREC_F16X3

_PREC = {"f16x3": PREC_F16X3, "fp32": PREC_F16X3, "f16": PREC_F16, "fast": PREC_F16, "f16f8": PREC_F16F8}
_PREC_NAME = {PREC_F16X3: "f16x3", PREC_F16: "f16", PREC_F16F8: "f16f8"}


def param_inventory() -> "OrderedDict[str, Tuple[int, ...]]":
    """name -> shape of the 276 tensors ``FISRnet.model`` creates (ops.py:8-9), in creation order."""
    lib = _lib.load()
    out: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()
    dims = (C.c_int * 4)()
    for k in range(lib.fisr_num_params()):
        rank = lib.fisr_param_shape(k, dims)
        out[lib.fisr_param_name(k).decode()] = tuple(dims[j] for j in range(rank))
    return out


class Engine:
    def __init__(self, device: int = 0, precision: str = "f16x3"):
        self.lib = _lib.load()
        if not torch.cuda.is_available():
            raise FisrError("fisr_b200 needs a CUDA device (sm_100a); there is no CPU path")
        self.device = int(device)
        h = C.c_void_p()
        rc = self.lib.fisr_create(self.device, C.byref(h))
        if rc != 0:
            raise FisrError(f"fisr_create failed ({rc}): {self.lib.fisr_last_error(None).decode()}")
        self.h = h
        # All library work is issued on this side stream, ordered against torch's current stream on entry and
        # exit (torch's default stream has handle 0, which the C ABI reads as "the context's own stream").
        self.stream = torch.cuda.Stream(self.device)
        self.set_precision(precision)

    # ------------------------------------------------------------------ plumbing
    def _check(self, rc: int, what: str) -> None:
        if rc != 0:
            raise FisrError(f"{what} failed ({rc}): {self.lib.fisr_last_error(self.h).decode()}")

    def close(self) -> None:
        if getattr(self, "h", None):
            self.lib.fisr_destroy(self.h)
            self.h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def _stream(self) -> int:
        return self.stream.cuda_stream

    def _enter(self, *tensors) -> None:
        self.stream.wait_stream(torch.cuda.current_stream(self.device))
        for t in tensors:
            if t is not None:
                t.record_stream(self.stream)

    def _exit(self) -> None:
        torch.cuda.current_stream(self.device).wait_stream(self.stream)

    def _dev(self, t: torch.Tensor, dtype: torch.dtype) -> torch.Tensor:
        if not (t.is_cuda and t.device.index == self.device and t.dtype == dtype and t.is_contiguous()):
            raise FisrError(f"expected a contiguous {dtype} tensor on cuda:{self.device}, got {t.dtype} on {t.device}")
        return t

    def set_precision(self, precision: str) -> None:
        if precision not in _PREC:
            raise FisrError(f"unknown precision {precision!r}; use one of {sorted(_PREC)}")
        self._check(self.lib.fisr_set_precision(self.h, _PREC[precision]), "fisr_set_precision")
        self.precision = _PREC_NAME[_PREC[precision]]

    @property
    def launch_count(self) -> int:
        return int(self.lib.fisr_launch_count(self.h))

    # ------------------------------------------------------------------ parameters
    def set_params(self, params: Dict[str, "np.ndarray | torch.Tensor"]) -> None:
        for name, shape in param_inventory().items():
            if name not in params:
                raise FisrError(f"missing parameter {name}")
            a = params[name]
            a = a.detach().cpu().numpy() if isinstance(a, torch.Tensor) else np.asarray(a)
            a = np.ascontiguousarray(a, dtype=np.float32)
            if tuple(a.shape) != shape:
                raise FisrError(f"{name}: expected shape {shape}, got {tuple(a.shape)}")
            self._check(self.lib.fisr_set_param(self.h, name.encode(), a.ctypes.data, a.size), f"fisr_set_param({name})")

    def get_params(self) -> "OrderedDict[str, np.ndarray]":
        out: "OrderedDict[str, np.ndarray]" = OrderedDict()
        for name, shape in param_inventory().items():
            a = np.empty(shape, dtype=np.float32)
            self._check(self.lib.fisr_get_param(self.h, name.encode(), a.ctypes.data, a.size), f"fisr_get_param({name})")
            out[name] = a
        return out

    # ------------------------------------------------------------------ FISRnet.model
    def forward(self, img: torch.Tensor, want=(True, True, True)):
        """``FISRnet.model`` (FISRnet.py:73-173): img [N,H,W,29] fp32 on the GPU -> (pred_l1, pred_l2, pred_l3)."""
        img = self._dev(img, torch.float32)
        n, h, w, c = img.shape
        if c != 29:
            raise FisrError(f"FISRnet.model takes 29 input channels, got {c}")
        outs = [torch.empty((n, h // 2, w // 2, 9), device=img.device) if want[0] else None,
                torch.empty((n, h, w, 9), device=img.device) if want[1] else None,
                torch.empty((n, 2 * h, 2 * w, 9), device=img.device) if want[2] else None]
        ptr = [o.data_ptr() if o is not None else None for o in outs]
        self._enter(img, *outs)
        self._check(self.lib.fisr_forward(self.h, img.data_ptr(), n, h, w, ptr[0], ptr[1], ptr[2], self._stream()),
                    "fisr_forward")
        self._exit()
        return tuple(outs)

    def forward_host(self, img: np.ndarray):
        """Same through host buffers (the ``sess.run(feed_dict)`` of FISRnet.py:1048): numpy in, numpy out."""
        img = np.ascontiguousarray(img, dtype=np.float32)
        n, h, w, c = img.shape
        if c != 29:
            raise FisrError(f"FISRnet.model takes 29 input channels, got {c}")
        o1 = np.empty((n, h // 2, w // 2, 9), np.float32)
        o2 = np.empty((n, h, w, 9), np.float32)
        o3 = np.empty((n, 2 * h, 2 * w, 9), np.float32)
        self._check(self.lib.fisr_forward_host(self.h, img.ctypes.data, n, h, w, o1.ctypes.data, o2.ctypes.data,
                                               o3.ctypes.data), "fisr_forward_host")
        return o1, o2, o3

    # ------------------------------------------------------------------ tiled window (FISRnet.py:994-1065)
    @staticmethod
    def canvas_shape(H: int, W: int, num_patch=(2, 2)) -> Tuple[int, int, int]:
        h = H - H % (32 * num_patch[0])
        w = W - W % (32 * num_patch[1])
        return 2 * h, 2 * w, 9

    def window(self, frames: torch.Tensor, flow: torch.Tensor, warp: torch.Tensor, num_patch=(2, 2),
               tiles: Optional[Tuple[int, int]] = None, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """One sliding window on device tensors: frames u8 [H,W,9], flow f32 [H,W,8], warp f32 [H,W,12]
        -> uint8 canvas [2h,2w,9].  ``tiles=(first, count)`` restricts the work to part of the tile grid."""
        frames = self._dev(frames, torch.uint8)
        flow = self._dev(flow, torch.float32)
        warp = self._dev(warp, torch.float32)
        H, W, _ = frames.shape
        if out is None:
            out = torch.zeros(self.canvas_shape(H, W, num_patch), dtype=torch.uint8, device=frames.device)
        first, count = tiles if tiles is not None else (0, num_patch[0] * num_patch[1])
        self._enter(frames, flow, warp, out)
        self._check(self.lib.fisr_window_device(self.h, frames.data_ptr(), flow.data_ptr(), warp.data_ptr(), H, W,
                                                num_patch[0], num_patch[1], first, count, self._dev(out, torch.uint8).data_ptr(),
                                                self._stream()), "fisr_window_device")
        self._exit()
        return out

    def units(self, frames: torch.Tensor, flow: torch.Tensor, warp: torch.Tensor, units, num_patch=(2, 2),
              layout: str = "units", out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """Batched (window, tile) units: frames u8 [B,H,W,9], flow [B,H,W,8], warp [B,H,W,12]; unit id =
        window * tiles_per_window + tile.  layout "frames" -> [B,2h,2w,9]; "units" -> [len(units),2h/pH,2w/pW,9]."""
        frames = self._dev(frames, torch.uint8)
        flow = self._dev(flow, torch.float32)
        warp = self._dev(warp, torch.float32)
        B, H, W, _ = frames.shape
        oh, ow, _ = self.canvas_shape(H, W, num_patch)
        units = [int(u) for u in units]
        if out is None:
            shape = (B, oh, ow, 9) if layout == "frames" else (len(units), oh // num_patch[0], ow // num_patch[1], 9)
            out = torch.zeros(shape, dtype=torch.uint8, device=frames.device)
        arr = (C.c_int * max(1, len(units)))(*units)
        self._enter(frames, flow, warp, out)
        self._check(self.lib.fisr_units_device(self.h, frames.data_ptr(), flow.data_ptr(), warp.data_ptr(), B, H, W,
                                               num_patch[0], num_patch[1], arr, len(units), 0 if layout == "frames" else 1,
                                               self._dev(out, torch.uint8).data_ptr(), self._stream()), "fisr_units_device")
        self._exit()
        return out

    def window_f32(self, frames: torch.Tensor, flow: torch.Tensor, warp: torch.Tensor, num_patch=(2, 2)) -> torch.Tensor:
        frames = self._dev(frames, torch.uint8)
        H, W, _ = frames.shape
        out = torch.zeros(self.canvas_shape(H, W, num_patch), dtype=torch.float32, device=frames.device)
        self._enter(frames, flow, warp, out)
        self._check(self.lib.fisr_window_device_f32(self.h, frames.data_ptr(), self._dev(flow, torch.float32).data_ptr(),
                                                    self._dev(warp, torch.float32).data_ptr(), H, W, num_patch[0],
                                                    num_patch[1], out.data_ptr(), self._stream()), "fisr_window_device_f32")
        self._exit()
        return out

    def window_host(self, frames: np.ndarray, flow: np.ndarray, warp: np.ndarray, num_patch=(2, 2),
                    out: Optional[np.ndarray] = None) -> np.ndarray:
        """Host-buffer form of :meth:`window` (H2D + tiles + D2H inside the call)."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        warp = np.ascontiguousarray(warp, dtype=np.float32)
        H, W, _ = frames.shape
        if flow.shape != (H, W, 8) or warp.shape != (H, W, 12) or frames.shape[2] != 9:
            raise FisrError(f"window shapes: frames {frames.shape}, flow {flow.shape}, warp {warp.shape}")
        if out is None:
            out = np.empty(self.canvas_shape(H, W, num_patch), np.uint8)
        self._check(self.lib.fisr_window_host(self.h, frames.ctypes.data, flow.ctypes.data, warp.ctypes.data, H, W,
                                              num_patch[0], num_patch[1], out.ctypes.data), "fisr_window_host")
        return out

    def window_host_f32(self, frames: np.ndarray, flow: np.ndarray, warp: np.ndarray, num_patch=(2, 2)) -> np.ndarray:
        """Float canvas [2h,2w,9] of one window before clipping (``test_Pred_full``, FISRnet.py:844-880), host buffers."""
        frames = np.ascontiguousarray(frames, dtype=np.uint8)
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        warp = np.ascontiguousarray(warp, dtype=np.float32)
        H, W, _ = frames.shape
        if flow.shape != (H, W, 8) or warp.shape != (H, W, 12) or frames.shape[2] != 9:
            raise FisrError(f"window shapes: frames {frames.shape}, flow {flow.shape}, warp {warp.shape}")
        out = np.empty(self.canvas_shape(H, W, num_patch), np.float32)
        self._check(self.lib.fisr_window_host_f32(self.h, frames.ctypes.data, flow.ctypes.data, warp.ctypes.data, H, W,
                                                  num_patch[0], num_patch[1], out.ctypes.data), "fisr_window_host_f32")
        return out

    def window_submit(self, slot: int, frames: np.ndarray, flow: np.ndarray, warp: np.ndarray, num_patch=(2, 2),
                      out: Optional[np.ndarray] = None) -> np.ndarray:
        """Pipelined :meth:`window_host`: enqueue one window on slot 0/1 and return; :meth:`window_wait` delivers ``out``.
        Arrays must be C-contiguous uint8 / float32 (ideally pinned) and are kept alive until the wait."""
        for a, dt in ((frames, np.uint8), (flow, np.float32), (warp, np.float32)):
            if a.dtype != dt or not a.flags.c_contiguous:
                raise FisrError("window_submit needs C-contiguous uint8 frames and float32 flow / warp")
        H, W, _ = frames.shape
        if out is None:
            out = np.empty(self.canvas_shape(H, W, num_patch), np.uint8)
        self._check(self.lib.fisr_window_submit(self.h, slot, frames.ctypes.data, flow.ctypes.data, warp.ctypes.data, H, W,
                                                num_patch[0], num_patch[1], out.ctypes.data), "fisr_window_submit")
        self._inflight = getattr(self, "_inflight", {})
        self._inflight[slot] = (frames, flow, warp, out)
        return out

    def window_wait(self, slot: int) -> np.ndarray:
        self._check(self.lib.fisr_window_wait(self.h, slot), "fisr_window_wait")
        return self._inflight.pop(slot)[3]

    def _pinned(self, slot: int, idx: int, shape, dtype) -> np.ndarray:
        """Page-locked staging array of one pipeline slot (cudaMemcpyAsync from pageable memory blocks the host, so the two
        windows in flight would not overlap)."""
        cache = self.__dict__.setdefault("_pin_cache", {})
        key = (slot, idx)
        t = cache.get(key)
        tdt = {np.dtype(np.uint8): torch.uint8, np.dtype(np.float32): torch.float32}[np.dtype(dtype)]
        if t is None or tuple(t.shape) != tuple(shape) or t.dtype != tdt:
            t = torch.empty(tuple(shape), dtype=tdt).pin_memory()
            cache[key] = t
        return t.numpy()

    def video_windows(self, windows, num_patch=(2, 2)):
        """Generator over uint8 canvases for an iterable of (frames, flow, warp) host windows, two windows in flight:
        each window is staged through the slot's pinned buffers, so its H2D copy overlaps the previous window's kernels
        (the yielded canvas is a private copy of the slot's pinned output buffer)."""
        pending = []
        for k, (fr, fl, wp) in enumerate(windows):
            slot = k & 1
            stage = []
            for idx, (a, dt) in enumerate(((fr, np.uint8), (fl, np.float32), (wp, np.float32))):
                buf = self._pinned(slot, idx, a.shape, dt)
                np.copyto(buf, a, casting="same_kind")
                stage.append(buf)
            H, W, _ = stage[0].shape
            out = self._pinned(slot, 3, self.canvas_shape(H, W, num_patch), np.uint8)
            self.window_submit(slot, stage[0], stage[1], stage[2], num_patch, out=out)
            pending.append(slot)
            if len(pending) == 2:
                yield self.window_wait(pending.pop(0)).copy()
        while pending:
            yield self.window_wait(pending.pop(0)).copy()

    # ------------------------------------------------------------------ flow warp
    def warp(self, yuv: torch.Tensor, flow: torch.Tensor, flow_scale: float = 0.5, out_scale: float = 1.0) -> torch.Tensor:
        """``warp_flow`` with the colour round trip (..warp_img_with_flo.py:61-67,112-128) on device tensors."""
        yuv = self._dev(yuv, torch.uint8)
        flow = self._dev(flow, torch.float32)
        h, w, _ = yuv.shape
        out = torch.empty((h, w, 3), dtype=torch.float32, device=yuv.device)
        self._enter(yuv, flow, out)
        self._check(self.lib.fisr_warp_device(self.h, yuv.data_ptr(), flow.data_ptr(), flow_scale, out.data_ptr(), h, w,
                                              out_scale, self._stream()), "fisr_warp_device")
        self._exit()
        return out

    def warp_batch(self, frames: torch.Tensor, flow: torch.Tensor, src_index, flow_scale: float = 0.5,
                   out_scale: float = 1.0, out: Optional[torch.Tensor] = None) -> torch.Tensor:
        """All warps of a clip in one launch: frames u8 [F,h,w,3], flow f32 [J,h,w,2], ``src_index[j]`` = the frame job j
        samples -> f32 [J,h,w,3] (the loop of ..warp_img_with_flo.py:112-128)."""
        frames = self._dev(frames, torch.uint8)
        flow = self._dev(flow, torch.float32)
        J, h, w, _ = flow.shape
        src = [int(i) for i in src_index]                 # checked on the host: no device round trip before the launch
        if len(src) != J or min(src, default=0) < 0 or max(src, default=0) >= frames.shape[0] or tuple(frames.shape[1:]) != (h, w, 3):
            raise FisrError(f"warp_batch: {J} flows, {len(src)} source indices, frames {tuple(frames.shape)}")
        key = (tuple(src), frames.device)
        if getattr(self, "_warp_idx_key", None) != key:      # the index table of a clip is uploaded once and kept
            self._warp_idx, self._warp_idx_key = torch.tensor(src, dtype=torch.int32, device=frames.device), key
        idx = self._warp_idx
        if out is None:
            out = torch.empty((J, h, w, 3), dtype=torch.float32, device=frames.device)
        elif tuple(self._dev(out, torch.float32).shape) != (J, h, w, 3):
            raise FisrError(f"warp_batch: out has shape {tuple(out.shape)}, expected {(J, h, w, 3)}")
        self._enter(frames, flow, idx, out)
        self._check(self.lib.fisr_warp_batch_device(self.h, frames.data_ptr(), flow.data_ptr(), idx.data_ptr(), J, flow_scale,
                                                    out.data_ptr(), h, w, out_scale, self._stream()), "fisr_warp_batch_device")
        self._exit()
        return out

    def warp_host(self, yuv: np.ndarray, flow: np.ndarray, flow_scale: float = 0.5, out_scale: float = 1.0) -> np.ndarray:
        yuv = np.ascontiguousarray(yuv, dtype=np.uint8)
        flow = np.ascontiguousarray(flow, dtype=np.float32)
        h, w, _ = yuv.shape
        out = np.empty((h, w, 3), np.float32)
        self._check(self.lib.fisr_warp_host(self.h, yuv.ctypes.data, flow.ctypes.data, flow_scale, out.ctypes.data, h, w,
                                            out_scale), "fisr_warp_host")
        return out

    # ------------------------------------------------------------------ training-side forward half
    LOSS_NAMES = ("recnLoss", "tmLoss", "tmmLoss", "tdLoss", "totalLoss_s1", "recnLoss_ss2", "tdLoss_ss2", "tmLoss_ss2",
                  "totalLoss_ss2", "total_loss", "train_PSNR")

    @staticmethod
    def _lambdas(lambdas):
        if lambdas is None:
            return None
        vals = [float(lambdas[k]) for k in ("recn", "tm1", "tm2", "tmm", "td", "ss2")]
        return (C.c_float * 6)(*vals)

    def groups2ovlp(self, pred: torch.Tensor) -> torch.Tensor:
        """``Groups2Ovlp`` (ops.py:119-144): pred [3B,H,W,9] of windows 0,1,2 -> [B,7,H,W,3]."""
        pred = self._dev(pred, torch.float32)
        n, H, W, _ = pred.shape
        out = torch.empty((n // 3, 7, H, W, 3), device=pred.device)
        self._enter(pred, out)
        self._check(self.lib.fisr_groups2ovlp(self.h, pred.data_ptr(), n // 3, H, W, out.data_ptr(), self._stream()),
                    "fisr_groups2ovlp")
        self._exit()
        return out

    def temporal_loss(self, preds, label: torch.Tensor, lambdas=None) -> dict:
        """The 11 scalars of FISRnet.py:651-657 from the three outputs of the 4B-batch forward and label [B,2h,2w,21]."""
        p = [self._dev(t, torch.float32) for t in preds]
        label = self._dev(label, torch.float32)
        B, H2, W2, _ = label.shape
        out = (C.c_float * 11)()
        self._enter(*p, label)
        self._check(self.lib.fisr_temporal_loss(self.h, p[0].data_ptr(), p[1].data_ptr(), p[2].data_ptr(), label.data_ptr(),
                                                B, H2 // 2, W2 // 2, self._lambdas(lambdas), out, self._stream()),
                    "fisr_temporal_loss")
        self._exit()
        return dict(zip(self.LOSS_NAMES, (float(v) for v in out)))

    def train_forward(self, data, flow, flow_ss2, warp, warp_ss2, label, lambdas=None) -> dict:
        """Forward half of one training step (FISRnet.py:281-486): 4 weight-shared passes as one batch + loss scalars."""
        ts = [self._dev(t, torch.float32) for t in (data, flow, flow_ss2, warp, warp_ss2, label)]
        B, h, w, _ = ts[0].shape
        out = (C.c_float * 11)()
        self._enter(*ts)
        self._check(self.lib.fisr_train_forward(self.h, *[t.data_ptr() for t in ts], B, h, w, self._lambdas(lambdas), out,
                                                self._stream()), "fisr_train_forward")
        self._exit()
        return dict(zip(self.LOSS_NAMES, (float(v) for v in out)))

    def train_backward(self, data, flow, flow_ss2, warp, warp_ss2, label, lambdas=None) -> dict:
        """Forward + loss + backward of one step (FISRnet.py:281-491 up to the gradients): returns the 11 scalars;
        the gradients stay in the context (:meth:`get_grads`, :meth:`adam_apply`)."""
        ts = [self._dev(t, torch.float32) for t in (data, flow, flow_ss2, warp, warp_ss2, label)]
        B, h, w, _ = ts[0].shape
        out = (C.c_float * 11)()
        self._enter(*ts)
        self._check(self.lib.fisr_train_backward(self.h, *[t.data_ptr() for t in ts], B, h, w, self._lambdas(lambdas), out,
                                                 self._stream()), "fisr_train_backward")
        self._exit()
        return dict(zip(self.LOSS_NAMES, (float(v) for v in out)))

    def train_step(self, data, flow, flow_ss2, warp, warp_ss2, label, lr: float, lambdas=None) -> dict:
        """One ``sess.run([self.optim, ...])`` (FISRnet.py:651): forward, loss, backward, TF-1.13 Adam update."""
        ts = [self._dev(t, torch.float32) for t in (data, flow, flow_ss2, warp, warp_ss2, label)]
        B, h, w, _ = ts[0].shape
        out = (C.c_float * 11)()
        self._enter(*ts)
        self._check(self.lib.fisr_train_step(self.h, *[t.data_ptr() for t in ts], B, h, w, self._lambdas(lambdas), lr, out,
                                             self._stream()), "fisr_train_step")
        self._exit()
        return dict(zip(self.LOSS_NAMES, (float(v) for v in out)))

    def get_grads(self) -> "OrderedDict[str, np.ndarray]":
        out: "OrderedDict[str, np.ndarray]" = OrderedDict()
        for name, shape in param_inventory().items():
            a = np.empty(shape, dtype=np.float32)
            self._check(self.lib.fisr_get_grad(self.h, name.encode(), a.ctypes.data, a.size), f"fisr_get_grad({name})")
            out[name] = a
        return out

    def adam_apply(self, lr: float, beta1=0.9, beta2=0.999, eps=1e-8) -> int:
        self._check(self.lib.fisr_adam_apply(self.h, lr, beta1, beta2, eps), "fisr_adam_apply")
        return int(self.lib.fisr_adam_steps(self.h))

    def set_loss_scale(self, scale: float) -> None:
        self._check(self.lib.fisr_set_loss_scale(self.h, float(scale)), "fisr_set_loss_scale")

    def set_wgrad_exact(self, exact: bool) -> None:
        """Weight gradients with both planes of the forward activation everywhere (default: hi plane only for layers
        with >= 16384 pixels)."""
        self._check(self.lib.fisr_set_wgrad_exact(self.h, int(bool(exact))), "fisr_set_wgrad_exact")

    def profile_train(self, B: int, h: int, w: int, reps: int = 2):
        """Per-op device time of the backward pass (after one :meth:`train_backward` at that shape)."""
        torch.cuda.synchronize(self.device)
        cnt = self.lib.fisr_profile_train(self.h, B, h, w, 1, 0, None, None, None, 0)
        if cnt < 0:
            self._check(cnt, "fisr_profile_train")
        ms = (C.c_float * cnt)()
        fl = (C.c_double * cnt)()
        names = C.create_string_buffer(cnt * 128)
        rc = self.lib.fisr_profile_train(self.h, B, h, w, reps, cnt, ms, fl, names, 128)
        if rc < 0:
            self._check(rc, "fisr_profile_train")
        out = []
        for k in range(cnt):
            nm = names.raw[k * 128:(k + 1) * 128].split(b"\0", 1)[0].decode()
            kind = "dgrad" if nm.endswith("[dgrad]") else ("wgrad" if nm.endswith("[wgrad]") else "aux")
            out.append({"name": nm, "kind": kind, "ms": float(ms[k]), "flops": float(fl[k])})
        return out

    def adam_step(self, grads: Dict[str, torch.Tensor], lr: float, beta1=0.9, beta2=0.999, eps=1e-8) -> int:
        """``tf.train.AdamOptimizer(lr)`` update (TF-1.13 formula) from device gradients keyed by variable name."""
        names = list(param_inventory())
        keep = [self._dev(grads[n].contiguous(), torch.float32) for n in names]
        arr = (C.c_void_p * len(keep))(*[t.data_ptr() for t in keep])
        torch.cuda.synchronize(self.device)
        self._check(self.lib.fisr_adam_step(self.h, arr, len(keep), lr, beta1, beta2, eps), "fisr_adam_step")
        return int(self.lib.fisr_adam_steps(self.h))

    def adam_reset(self, step: int = 0) -> None:
        """Zero moments and step counter ``step`` (bias correction matches zero moments only for ``step = 0``)."""
        self._check(self.lib.fisr_adam_reset(self.h, step), "fisr_adam_reset")

    @property
    def adam_steps(self) -> int:
        return int(self.lib.fisr_adam_steps(self.h))

    def get_adam_state(self) -> dict:
        """Optimizer state as the reference's Saver stores it: ``<var>/Adam`` (m), ``<var>/Adam_1`` (v), step counter t."""
        out = {"t": self.adam_steps, "m": OrderedDict(), "v": OrderedDict()}
        for name, shape in param_inventory().items():
            for which, key in ((0, "m"), (1, "v")):
                a = np.empty(shape, dtype=np.float32)
                self._check(self.lib.fisr_get_adam_slot(self.h, name.encode(), which, a.ctypes.data, a.size), "fisr_get_adam_slot")
                out[key][name] = a
        return out

    def set_adam_state(self, m: Dict[str, np.ndarray], v: Dict[str, np.ndarray], t: int) -> None:
        for name, shape in param_inventory().items():
            for which, src in ((0, m), (1, v)):
                a = np.ascontiguousarray(src[name], dtype=np.float32)
                if tuple(a.shape) != shape:
                    raise FisrError(f"Adam slot of {name}: expected shape {shape}, got {tuple(a.shape)}")
                self._check(self.lib.fisr_set_adam_slot(self.h, name.encode(), which, a.ctypes.data, a.size), "fisr_set_adam_slot")
        self._check(self.lib.fisr_adam_set_steps(self.h, int(t)), "fisr_adam_set_steps")

    # ------------------------------------------------------------------ test hooks
    def conv3x3(self, x: torch.Tensor, w: torch.Tensor, b: torch.Tensor, res: Optional[torch.Tensor] = None,
                relu: bool = True, d2s: bool = False, want_raw: bool = True, want_act: bool = True):
        x = self._dev(x, torch.float32)
        w = self._dev(w, torch.float32)
        b = self._dev(b, torch.float32)
        n, h, wd, cin = x.shape
        cout = w.shape[3]
        raw = torch.empty((n, h, wd, cout), device=x.device) if want_raw else None
        act = None
        if want_act:
            act = torch.empty((n, 2 * h, 2 * wd, cout // 4) if d2s else (n, h, wd, cout), device=x.device)
        torch.cuda.synchronize(self.device)
        self._check(self.lib.fisr_conv3x3(self.h, x.data_ptr(), w.data_ptr(), b.data_ptr(),
                                          self._dev(res, torch.float32).data_ptr() if res is not None else None,
                                          n, h, wd, cin, cout, int(relu), int(d2s),
                                          raw.data_ptr() if raw is not None else None,
                                          act.data_ptr() if act is not None else None), "fisr_conv3x3")
        return raw, act

    def dgrad3x3(self, dy: torch.Tensor, w: torch.Tensor, mask: Optional[torch.Tensor] = None,
                 res: Optional[torch.Tensor] = None, s2d: bool = False, want_raw: bool = True):
        """Data gradient of one 3x3 SAME conv (production kernel on rotated-transposed planes):
        dx = conv3x3(dy, rot180(w)^T) * [mask > 0] + res; dy [N,H,W,Cout], w HWIO forward filter -> (raw, act)."""
        dy = self._dev(dy, torch.float32)
        w = self._dev(w, torch.float32)
        n, h, wd, cout = dy.shape
        cin = w.shape[2]
        raw = torch.empty((n, h, wd, cin), device=dy.device) if (want_raw and not s2d) else None
        act = torch.empty((n, h // 2, wd // 2, 4 * cin) if s2d else (n, h, wd, cin), device=dy.device)
        torch.cuda.synchronize(self.device)
        self._check(self.lib.fisr_dgrad3x3(self.h, dy.data_ptr(), w.data_ptr(),
                                           self._dev(mask, torch.float32).data_ptr() if mask is not None else None,
                                           self._dev(res, torch.float32).data_ptr() if res is not None else None,
                                           n, h, wd, cin, cout, int(s2d), raw.data_ptr() if raw is not None else None,
                                           act.data_ptr()), "fisr_dgrad3x3")
        return raw, act

    def wgrad3x3(self, x: torch.Tensor, dy: torch.Tensor, scale: float = 1.0, want_bias: bool = True):
        """Weight / bias gradient of one 3x3 SAME conv (production wgrad kernel): x [N,H,W,Cin], dy [N,H,W,Cout]
        -> (gw HWIO [3,3,Cin,Cout], gb [Cout])."""
        x = self._dev(x, torch.float32)
        dy = self._dev(dy, torch.float32)
        n, h, wd, cin = x.shape
        cout = dy.shape[3]
        gw = torch.empty((3, 3, cin, cout), device=x.device)
        gb = torch.empty((cout,), device=x.device) if want_bias else None
        torch.cuda.synchronize(self.device)
        self._check(self.lib.fisr_wgrad3x3(self.h, x.data_ptr(), dy.data_ptr(), n, h, wd, cin, cout, scale, gw.data_ptr(),
                                           gb.data_ptr() if gb is not None else None), "fisr_wgrad3x3")
        return gw, gb

    def debug_conv_output(self, conv_name: str, shape) -> np.ndarray:
        a = np.empty(shape, np.float32)
        self._check(self.lib.fisr_debug_conv_output(self.h, conv_name.encode(), a.ctypes.data, a.size),
                    "fisr_debug_conv_output")
        return a

    def profile_ops(self, n: int, h: int, w: int, reps: int = 3):
        """Per-launch device time of the (n,h,w) plan: list of dicts {name, kind, ms, flops, bytes}."""
        torch.cuda.synchronize(self.device)
        cnt = self.lib.fisr_profile_ops(self.h, n, h, w, 1, 0, None, None, None, None, None, 0)
        if cnt < 0:
            self._check(cnt, "fisr_profile_ops")
        ms = (C.c_float * cnt)()
        fl = (C.c_double * cnt)()
        by = (C.c_double * cnt)()
        kd = (C.c_int * cnt)()
        names = C.create_string_buffer(cnt * 96)
        rc = self.lib.fisr_profile_ops(self.h, n, h, w, reps, cnt, ms, fl, by, kd, names, 96)
        if rc < 0:
            self._check(rc, "fisr_profile_ops")
        out = []
        for k in range(cnt):
            nm = names.raw[k * 96:(k + 1) * 96].split(b"\0", 1)[0].decode()
            kind = "conv" if kd[k] < 1000 else ("upsample" if kd[k] < 2000 else "pred2next")
            out.append({"name": nm, "kind": kind, "nt": kd[k] % 1000, "ms": float(ms[k]), "flops": float(fl[k]),
                        "bytes": float(by[k])})
        return out

    def plan_info(self, n: int, h: int, w: int) -> dict:
        fl, eff, nl, ws = C.c_double(), C.c_double(), C.c_int(), C.c_size_t()
        self._check(self.lib.fisr_plan_info(self.h, n, h, w, C.byref(fl), C.byref(eff), C.byref(nl), C.byref(ws)),
                    "fisr_plan_info")
        return {"flops": fl.value, "mma_row_efficiency": eff.value, "launches": nl.value, "workspace_bytes": ws.value}
