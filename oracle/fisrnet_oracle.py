"""Oracle (test infrastructure): torch-CPU restatement of ``FISRnet.model`` and its blocks.

Follows the reference line by line:
  * blocks            -- ``ops.py:7-76``   (Conv2d, relu, res_block, Enc/Bottleneck/Dec level)
  * network topology  -- ``FISRnet.py:73-173``
  * variable names    -- scopes at ``ops.py:8-9,40,49,60,68`` under ``FISRnet/`` (``FISRnet.py:289,750``)

TF-1.13 op semantics that are NOT in the reference tree and are restated here
(SURVEY.md section 8c -- parity unpinned, see ``oracle/__init__.py``):
  * ``tf.nn.conv2d(NHWC, HWIO, stride 1, 'SAME')``   = zero-pad-1 cross-correlation
  * ``resize_images(BICUBIC)`` at integer down-scale  = ``x[:, ::s, ::s, :]`` (legacy kernel,
    align_corners=False, no half-pixel centres: the cubic weights collapse to [0,1,0,0])
  * ``resize_images(BILINEAR)`` x2 (legacy)           = ``out[2k]=in[k]``,
    ``out[2k+1]=(in[k]+in[min(k+1,n-1)])/2`` applied along H then W
  * ``tf.depth_to_space(x, 2)`` NHWC                  = ``out[n,2h+i,2w+j,c]=x[n,h,w,(2i+j)*C+c]``
  * ``max_pool 2x2 s2 SAME`` on even dims             = plain 2x2 max

All tensors at the API are NHWC like the reference; internally convs run NCHW through
``torch.nn.functional.conv2d`` (oneDNN), in the dtype of the parameters (fp32 or fp64).
"""
from __future__ import annotations

import math
from collections import OrderedDict
from typing import Callable, Dict, Iterator, Optional, Tuple

import torch
import torch.nn.functional as F

CH = 64          # FISRnet.py:74
IN_CH = 29       # 9 (3 YUV frames) + 8 (4 flows) + 12 (4 warped YUV frames), FISRnet.py:287
SF = 2           # main.py:29 default

# ---------------------------------------------------------------------------------------
# parameter inventory (creation order of tf.get_variable inside FISRnet.model)
# ---------------------------------------------------------------------------------------

def _res_block_names(prefix: str, c: int) -> Iterator[Tuple[str, int, int]]:
    yield f"{prefix}/conv/0", c, c          # ops.py:41
    yield f"{prefix}/conv/1", c, c          # ops.py:42


def _enc_names(prefix: str, c1: int, c: int) -> Iterator[Tuple[str, int, int]]:
    yield f"{prefix}/conv/0", c1, c         # ops.py:50
    yield from _res_block_names(f"{prefix}/res_block/0", c)
    yield from _res_block_names(f"{prefix}/res_block/1", c)


def _dec_names(prefix: str, c1: int, c: int) -> Iterator[Tuple[str, int, int]]:
    yield f"{prefix}/resize", c1, c         # ops.py:70
    yield f"{prefix}/conv/0", 2 * c, c      # ops.py:73
    yield from _res_block_names(f"{prefix}/res_block/0", c)
    yield from _res_block_names(f"{prefix}/res_block/1", c)


def _head_names(prefix: str, cout: int) -> Iterator[Tuple[str, int, int]]:
    yield f"{prefix}/conv/0", CH, CH                   # FISRnet.py:96,102
    yield from _res_block_names(f"{prefix}/res_block/0", CH)
    yield f"{prefix}/conv/1", CH, CH * SF * SF         # FISRnet.py:98,104
    yield f"{prefix}/conv/2", CH, cout                 # FISRnet.py:100,106


def conv_inventory(scope: str = "FISRnet", in_ch: int = IN_CH) -> "OrderedDict[str, Tuple[int, int]]":
    """name -> (Cin, Cout) for the 138 convs, in graph-creation order (FISRnet.py:78-171)."""
    inv: "OrderedDict[str, Tuple[int, int]]" = OrderedDict()
    for lvl in (1, 2, 3):
        p = f"{scope}/level_{lvl}"
        cin0 = in_ch if lvl == 1 else in_ch + 9          # FISRnet.py:84,116,147
        items = []
        items += _enc_names(f"{p}/enc/level_0", cin0, CH)
        items += _enc_names(f"{p}/enc/level_1", CH, CH * 2)
        items += _enc_names(f"{p}/enc/level_2", CH * 2, CH * 4)
        items.append((f"{p}/bottleneck/conv/0", CH * 4, CH * 8))           # ops.py:61
        items += _res_block_names(f"{p}/bottleneck/res_block/0", CH * 8)
        items += _dec_names(f"{p}/dec/level_2", CH * 8, CH * 4)
        items += _dec_names(f"{p}/dec/level_1", CH * 4, CH * 2)
        items += _dec_names(f"{p}/dec/level_0", CH * 2, CH)
        items += _head_names(f"{p}/FI-SR", 6)
        items += _head_names(f"{p}/SR", 3)
        for name, ci, co in items:
            inv[name] = (ci, co)
    return inv


def init_params(seed: int = 0, dtype: torch.dtype = torch.float32, bias_std: float = 0.01,
                scope: str = "FISRnet") -> "OrderedDict[str, torch.Tensor]":
    """Seeded stand-in for the reference initialiser (weights are not shipped).

    ``ops.py:8``: xavier_initializer(uniform=False) = truncated normal (+-2 sigma) with
    sigma = sqrt(1.3 / n), n = (fan_in + fan_out) / 2, fan = 9*C  (TF-1.13 semantics).
    ``ops.py:9``: biases are zero in the reference; the oracle draws N(0, bias_std) so that
    bias bugs are visible (pass ``bias_std=0`` for the reference initial state).
    Values are generated in fp64 and cast, so fp32 and fp64 parameter sets agree.
    """
    g = torch.Generator().manual_seed(seed)
    params: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, (ci, co) in conv_inventory(scope).items():
        n = (9 * ci + 9 * co) / 2.0
        sigma = math.sqrt(1.3 / n)
        w = torch.empty(3, 3, ci, co, dtype=torch.float64)
        torch.nn.init.trunc_normal_(w, mean=0.0, std=sigma, a=-2 * sigma, b=2 * sigma, generator=g)
        b = torch.randn(co, dtype=torch.float64, generator=g) * bias_std
        params[name + "/w"] = w.to(dtype)
        params[name + "/b"] = b.to(dtype)
    return params


def cast_params(params: Dict[str, torch.Tensor], dtype: torch.dtype) -> "OrderedDict[str, torch.Tensor]":
    return OrderedDict((k, v.to(dtype)) for k, v in params.items())


# ---------------------------------------------------------------------------------------
# TF-1.13 op restatements (NCHW inside)
# ---------------------------------------------------------------------------------------

def to_nchw(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 3, 1, 2).contiguous(memory_format=torch.channels_last)


def to_nhwc(x: torch.Tensor) -> torch.Tensor:
    return x.permute(0, 2, 3, 1).contiguous()


def subsample(x: torch.Tensor, s: int) -> torch.Tensor:
    """legacy ``resize_images(BICUBIC)`` to 1/s (FISRnet.py:81,112,263,264) on NCHW."""
    return x[:, :, ::s, ::s]


def upsample2_legacy_bilinear(x: torch.Tensor) -> torch.Tensor:
    """legacy ``resize_images(BILINEAR)`` to exactly 2x (ops.py:69) on NCHW."""
    n, c, h, w = x.shape
    nxt = torch.cat([x[:, :, 1:, :], x[:, :, -1:, :]], dim=2)
    xh = torch.stack([x, 0.5 * (x + nxt)], dim=3).reshape(n, c, 2 * h, w)
    nxt = torch.cat([xh[:, :, :, 1:], xh[:, :, :, -1:]], dim=3)
    return torch.stack([xh, 0.5 * (xh + nxt)], dim=4).reshape(n, c, 2 * h, 2 * w)


def depth_to_space2(x: torch.Tensor) -> torch.Tensor:
    """``tf.depth_to_space(x, 2)`` (FISRnet.py:99,105) on NCHW: in-channel (2i+j)*C + c."""
    n, c4, h, w = x.shape
    c = c4 // 4
    x = x.reshape(n, 2, 2, c, h, w)            # [n, i, j, c, h, w]
    x = x.permute(0, 3, 4, 1, 5, 2)            # [n, c, h, i, w, j]
    return x.reshape(n, c, 2 * h, 2 * w)


class Net:
    """One parameter set + the block functions of ``ops.py``, NCHW inside."""

    def __init__(self, params: Dict[str, torch.Tensor], scope: str = "FISRnet",
                 operand_hook: Optional[Callable[[torch.Tensor], torch.Tensor]] = None,
                 tap: Optional[Dict[str, torch.Tensor]] = None):
        self.p = params
        self.scope = scope
        # operand_hook: applied to both conv operands (study tool for reduced-precision
        # operand formats; identity in the oracle proper).
        self.hook = operand_hook
        self.tap = tap      # optional dict that receives named intermediates (NHWC)

    def conv(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """``Conv2d`` ops.py:7-11."""
        w = self.p[name + "/w"]                 # HWIO
        b = self.p[name + "/b"]
        wk = w.permute(3, 2, 0, 1)              # OIHW
        if self.hook is not None:
            x, wk = self.hook(x), self.hook(wk)
        y = F.conv2d(x, wk, None, stride=1, padding=1) + b.view(1, -1, 1, 1)
        if self.tap is not None:
            self.tap[name] = to_nhwc(y)
        return y

    def res_block(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """ops.py:39-44 (pre-activation)."""
        n = self.conv(F.relu(x), name + "/conv/0")
        n = self.conv(F.relu(n), name + "/conv/1")
        return x + n

    def enc_level(self, x: torch.Tensor, name: str):
        """ops.py:48-55."""
        n = self.conv(x, name + "/conv/0")
        n = self.res_block(n, name + "/res_block/0")
        n = F.relu(self.res_block(n, name + "/res_block/1"))
        skip = n
        n = F.max_pool2d(n, 2, 2)
        return n, skip

    def bottleneck(self, x: torch.Tensor, name: str) -> torch.Tensor:
        """ops.py:59-63."""
        n = self.conv(x, name + "/conv/0")
        return F.relu(self.res_block(n, name + "/res_block/0"))

    def dec_level(self, x: torch.Tensor, skip: torch.Tensor, name: str) -> torch.Tensor:
        """ops.py:67-76; ``size`` is always exactly 2x the input (FISRnet.py:91-93)."""
        n = upsample2_legacy_bilinear(x)
        assert n.shape[2:] == skip.shape[2:]
        n = F.relu(self.conv(n, name + "/resize"))
        n = torch.cat([n, skip], dim=1)
        n = self.conv(n, name + "/conv/0")
        n = self.res_block(n, name + "/res_block/0")
        return F.relu(self.res_block(n, name + "/res_block/1"))

    def head(self, n: torch.Tensor, name: str, relu_before_last: bool) -> torch.Tensor:
        """FISRnet.py:95-106."""
        n2 = self.conv(n, name + "/conv/0")
        n2 = self.res_block(n2, name + "/res_block/0")
        n2 = self.conv(F.relu(n2), name + "/conv/1")
        n2 = depth_to_space2(F.relu(n2))
        if relu_before_last:
            n2 = F.relu(n2)
        return self.conv(n2, name + "/conv/2")

    def level(self, x: torch.Tensor, lvl: int) -> torch.Tensor:
        """One U-Net of the cascade, FISRnet.py:82-108 (identical for the 3 levels)."""
        p = f"{self.scope}/level_{lvl}"
        n, s0 = self.enc_level(x, p + "/enc/level_0")
        n, s1 = self.enc_level(n, p + "/enc/level_1")
        n, s2 = self.enc_level(n, p + "/enc/level_2")
        n = self.bottleneck(n, p + "/bottleneck")
        n = self.dec_level(n, s2, p + "/dec/level_2")
        n = self.dec_level(n, s1, p + "/dec/level_1")
        n = self.dec_level(n, s0, p + "/dec/level_0")
        fisr = self.head(n, p + "/FI-SR", True)
        sr = self.head(n, p + "/SR", False)
        return torch.cat([fisr[:, :3], sr, fisr[:, 3:]], dim=1)       # FISRnet.py:107-108

    def model_nchw(self, img: torch.Tensor):
        """``FISRnet.model`` FISRnet.py:73-173 on NCHW input [N,29,H,W]."""
        assert img.shape[2] % 32 == 0 and img.shape[3] % 32 == 0, "H, W must be multiples of 32"
        pred_l1 = self.level(subsample(img, 4), 1)                                  # :81
        pred_l2 = self.level(torch.cat([subsample(img, 2), pred_l1], dim=1), 2)     # :112-113
        pred_l3 = self.level(torch.cat([img, pred_l2], dim=1), 3)                   # :144
        return pred_l1, pred_l2, pred_l3


def model(params: Dict[str, torch.Tensor], img_nhwc: torch.Tensor, sf: int = 2, scope: str = "FISRnet",
          operand_hook=None, tap=None):
    """``FISRnet.model(img, sf)``: img [N,H,W,29] -> (pred_l1 [N,H/2,W/2,9], pred_l2 [N,H,W,9],
    pred_l3 [N,2H,2W,9]), all NHWC, in the dtype of ``params``."""
    assert sf == 2
    dtype = next(iter(params.values())).dtype
    net = Net(params, scope, operand_hook, tap)
    with torch.no_grad():
        outs = net.model_nchw(to_nchw(img_nhwc.to(dtype)))
    return tuple(to_nhwc(o) for o in outs)


# ---------------------------------------------------------------------------------------
# synthetic inputs shared by tests / smoke / bench (SURVEY.md section 8d value distributions)
# ---------------------------------------------------------------------------------------

def synthetic_input(n: int, h: int, w: int, seed: int = 1) -> torch.Tensor:
    """[n,h,w,29] fp32: frames U[0,1], flow N(0,4px)/96/2 clipped [-1,1], warp = frame+N(0,.02)."""
    g = torch.Generator().manual_seed(seed)
    frames = torch.rand(n, h, w, 9, generator=g)
    flow = (torch.randn(n, h, w, 8, generator=g) * 4.0 / 96.0 / 2.0).clamp(-1, 1)
    idx = [3, 4, 5, 0, 1, 2, 6, 7, 8, 3, 4, 5]
    warp = (frames[..., idx] + 0.02 * torch.randn(n, h, w, 12, generator=g)).clamp(0, 1)
    return torch.cat([frames, flow, warp], dim=3).contiguous()


def psnr(a: torch.Tensor, b: torch.Tensor, peak: float = 1.0) -> float:
    """``utils._compute_psnr`` utils.py:23-26."""
    mse = torch.mean((a.double() - b.double()) ** 2).item()
    return 10.0 * math.log10(peak * peak / mse) if mse > 0 else float("inf")
