"""Oracle (test infrastructure): torch-CPU restatement of the PWC-Net the reference runs in front of its flow warp.

PARITY UNPINNED.  The reference's PWC-Net is a partial copy of philferriere/tfoptflow: ``model_pwcnet.py`` imports eight modules
that are not vendored (``core_warp``, ``core_costvol``, ...; model_pwcnet.py:21-28) and no checkpoint is shipped, so neither the
code nor any output of it can run here.  What IS in the tree is followed line by line:

  * configuration      -- FISR_tfoptflow/FISR_for_video_pwcnet_predict_from_img_test.py:96-110: PWC-Net-large, dense + residual
                          connections, 6 pyramid levels, flow predicted at level 2, search range 4
  * feature pyramid    -- model_pwcnet.py:1012-1101 (conv{l}a stride 2, conv{l}aa, conv{l}b; 16..196 channels; leaky ReLU 0.1)
  * flow estimator     -- model_pwcnet.py:1282-1448 (DenseNet block 128,128,96,64,32 with ``concat([act, x])``, then flow{l})
  * context network    -- model_pwcnet.py:1453-1522 (dilations 1,2,4,8,16,1,1; residual add)
  * cascade            -- model_pwcnet.py:1525-1593 (scaler 20 / 2^l, conv2d_transpose 4x4 s2 with 2 filters for up_flow AND up_feat,
                          final ``resize_bilinear`` x4 times 4)
  * pre / post         -- ..predict_from_img_test.py:112-139 and model_pwcnet.py:371-409,449-470 (x2 ``skimage.transform.resize``,
                          uint8 truncation, /255, zero pad to a multiple of 64, crop, anti-aliased x1/2 resize, /2)

and the two un-vendored ops are restated from their published definitions:

  * ``cost_volume(c1, warp, 4)``    -- 81 channels, channel d = (dy+4)*9 + (dx+4): mean over features of c1[y,x] * warp[y+dy,x+dx]
                                        (zero outside), followed by leaky ReLU 0.1 (tfoptflow core_costvol.py)
  * ``dense_image_warp(c2, flow)``  -- tf.contrib.image.dense_image_warp as tfoptflow vendors it: output[y,x] = c2 sampled
                                        bilinearly at (y - flow_y', x - flow_x') in TF's convention; tfoptflow feeds it the (u, v)
                                        flow so that c2 is sampled at (x + u, y + v); floor clamped to [0, size-2], weights to [0,1]
  * TF ``'same'`` padding           -- stride-2 convs on even sizes pad 0 before / 1 after (not PyTorch's symmetric padding);
                                        conv2d_transpose 4x4 s2 ``'same'``: out[o] += in[i] w[k] for o = 2 i + k - 1
  * ``skimage.transform.resize``    -- order-1 interpolation at pixel-centre-aligned coordinates, mode 'reflect' (= scipy
                                        'mirror'), Gaussian pre-filter sigma (s-1)/2 when down-scaling with anti_aliasing
"""
from __future__ import annotations

from collections import OrderedDict
from typing import Dict, List, Tuple

import numpy as np
import torch
import torch.nn.functional as F

PYR_LVLS, FLOW_PRED_LVL, SEARCH_RANGE = 6, 2, 4            # ..predict_from_img_test.py:104-107, model_pwcnet.py defaults
NUM_CHANN = [None, 16, 32, 64, 96, 128, 196]               # model_pwcnet.py:1083
DENSE = [128, 128, 96, 64, 32]                             # model_pwcnet.py:1415-1433
CTXT = [(128, 1), (128, 2), (128, 4), (96, 8), (64, 16), (32, 1), (2, 1)]      # model_pwcnet.py:1506-1519


def param_inventory() -> "OrderedDict[str, Tuple[int, ...]]":
    """TF variable names (scope ``pwcnet/``) -> shapes, in graph-creation order.  conv kernels HWIO, transpose kernels
    [4,4,out,in] (tf.layers.conv2d_transpose)."""
    inv: "OrderedDict[str, Tuple[int, ...]]" = OrderedDict()

    def conv(name, cin, cout):
        inv[name + "/kernel"] = (3, 3, cin, cout)
        inv[name + "/bias"] = (cout,)

    for lvl in range(1, PYR_LVLS + 1):
        cin = 3 if lvl == 1 else NUM_CHANN[lvl - 1]
        conv(f"pwcnet/featpyr/conv{lvl}a", cin, NUM_CHANN[lvl])
        conv(f"pwcnet/featpyr/conv{lvl}aa", NUM_CHANN[lvl], NUM_CHANN[lvl])
        conv(f"pwcnet/featpyr/conv{lvl}b", NUM_CHANN[lvl], NUM_CHANN[lvl])
    for lvl in range(PYR_LVLS, FLOW_PRED_LVL - 1, -1):
        c = (2 * SEARCH_RANGE + 1) ** 2 + (0 if lvl == PYR_LVLS else NUM_CHANN[lvl] + 4)
        for k, f in enumerate(DENSE):
            conv(f"pwcnet/predict_flow/conv{lvl}_{k}", c, f)
            c += f
        conv(f"pwcnet/predict_flow/flow{lvl}", c, 2)
        cc = c
        for k, (f, _) in enumerate(CTXT):
            conv(f"pwcnet/ctxt/dc_conv{lvl}{k + 1}", cc, f)
            cc = f
        if lvl != FLOW_PRED_LVL:
            for nm, ci in (("up_flow", 2), ("up_feat", c)):
                inv[f"pwcnet/upsample/{nm}{lvl}/kernel"] = (4, 4, 2, ci)
                inv[f"pwcnet/upsample/{nm}{lvl}/bias"] = (2,)
    return inv


def init_params(seed: int = 0, dtype=torch.float32) -> "OrderedDict[str, torch.Tensor]":
    """Seeded stand-in for a trained checkpoint (none is shipped): he_normal kernels (model_pwcnet.py:1085), small random biases;
    the transpose kernels start near a bilinear up-sampler so that the cascade carries signal."""
    g = torch.Generator().manual_seed(seed)
    p: "OrderedDict[str, torch.Tensor]" = OrderedDict()
    for name, shape in param_inventory().items():
        if name.endswith("/bias"):
            p[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * 0.02).to(dtype)
        elif "/upsample/" in name:
            p[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * (1.0 / (4 * shape[3]) ** 0.5)).to(dtype)
        else:
            fan_in = shape[0] * shape[1] * shape[2]
            scale = 0.35 if name.endswith(("flow2/kernel", "flow3/kernel", "flow4/kernel", "flow5/kernel", "flow6/kernel")) else 1.0
            p[name] = (torch.randn(shape, generator=g, dtype=torch.float64) * (2.0 / fan_in) ** 0.5 * scale).to(dtype)
    return p


# ------------------------------------------------------------------------------------------ ops (NCHW inside)
def conv_same(x: torch.Tensor, w_hwio: torch.Tensor, b: torch.Tensor, stride: int = 1, dilation: int = 1) -> torch.Tensor:
    """tf.layers.conv2d(..., padding='same'): total padding max((ceil(n/s) - 1) s + (k-1) d + 1 - n, 0), the smaller half first."""
    k = w_hwio.shape[0]
    h, w = x.shape[2:]
    pads = []
    for n in (w, h):                                       # F.pad order: last dim first
        out = -(-n // stride)
        tot = max((out - 1) * stride + (k - 1) * dilation + 1 - n, 0)
        pads += [tot // 2, tot - tot // 2]
    x = F.pad(x, pads)
    return F.conv2d(x, w_hwio.permute(3, 2, 0, 1), b, stride=stride, dilation=dilation)


def lrelu(x: torch.Tensor) -> torch.Tensor:
    return F.leaky_relu(x, 0.1)


def cost_volume(c1: torch.Tensor, warp: torch.Tensor, r: int = SEARCH_RANGE) -> torch.Tensor:
    n, c, h, w = c1.shape
    pw = F.pad(warp, [r, r, r, r])
    out = [torch.mean(c1 * pw[:, :, dy:dy + h, dx:dx + w], dim=1, keepdim=True) for dy in range(2 * r + 1) for dx in range(2 * r + 1)]
    return lrelu(torch.cat(out, dim=1))


def dense_image_warp(img: torch.Tensor, flow_uv: torch.Tensor) -> torch.Tensor:
    """img [N,C,H,W] sampled at (x + u, y + v), flow_uv [N,2,H,W] = (u, v); TF's _interpolate_bilinear clamping."""
    n, c, h, w = img.shape
    ys, xs = torch.meshgrid(torch.arange(h, dtype=img.dtype), torch.arange(w, dtype=img.dtype), indexing="ij")
    qx, qy = xs + flow_uv[:, 0], ys + flow_uv[:, 1]
    out = []
    x0 = torch.clamp(torch.floor(qx), 0, w - 2)
    y0 = torch.clamp(torch.floor(qy), 0, h - 2)
    ax = torch.clamp(qx - x0, 0, 1).unsqueeze(1)
    ay = torch.clamp(qy - y0, 0, 1).unsqueeze(1)
    x0, y0 = x0.long(), y0.long()
    flat = img.reshape(n, c, h * w)

    def take(yy, xx):
        idx = (yy * w + xx).reshape(n, 1, h * w).expand(n, c, h * w)
        return torch.gather(flat, 2, idx).reshape(n, c, h, w)

    tl, tr, bl, br = take(y0, x0), take(y0, x0 + 1), take(y0 + 1, x0), take(y0 + 1, x0 + 1)
    top = tl + ax * (tr - tl)
    bot = bl + ax * (br - bl)
    return top + ay * (bot - top)


def deconv_same(x: torch.Tensor, w_hwoi: torch.Tensor, b: torch.Tensor) -> torch.Tensor:
    """tf.layers.conv2d_transpose(x, 2, 4, 2, 'same'): kernel [4,4,out,in]; out[2i + k - 1] += x[i] w[k]."""
    y = F.conv_transpose2d(x, w_hwoi.permute(3, 2, 0, 1), b, stride=2, padding=1)
    return y


def resize_bilinear_legacy(x: torch.Tensor, scale: int) -> torch.Tensor:
    """tf.image.resize_bilinear (align_corners=False, no half-pixel centres): src = dst / scale, right / bottom index clamped."""
    n, c, h, w = x.shape

    def axis(v, size, dim):
        dst = torch.arange(size * scale, dtype=x.dtype) / scale
        i0 = torch.floor(dst).long().clamp(max=size - 1)
        i1 = (i0 + 1).clamp(max=size - 1)
        fr = (dst - i0.to(x.dtype))
        shape = [1, 1, 1, 1]
        shape[dim] = -1
        return v.index_select(dim, i0) * (1 - fr.view(shape)) + v.index_select(dim, i1) * fr.view(shape)

    return axis(axis(x, h, 2), w, 3)


def forward(params: Dict[str, torch.Tensor], img1: torch.Tensor, img2: torch.Tensor, taps: Dict[str, torch.Tensor] = None) -> torch.Tensor:
    """``ModelPWCNet.nn`` (model_pwcnet.py:1525-1593): img1, img2 [N,H,W,3] in 0..1, H, W multiples of 64 -> flow [N,H,W,2]."""
    dt = next(iter(params.values())).dtype
    P = lambda n: (params[n + "/kernel"], params[n + "/bias"])

    def pyramid(x):
        out = [None]
        for lvl in range(1, PYR_LVLS + 1):
            x = lrelu(conv_same(x, *P(f"pwcnet/featpyr/conv{lvl}a"), stride=2))
            x = lrelu(conv_same(x, *P(f"pwcnet/featpyr/conv{lvl}aa")))
            x = lrelu(conv_same(x, *P(f"pwcnet/featpyr/conv{lvl}b")))
            out.append(x)
        return out

    with torch.no_grad():
        c1 = pyramid(img1.to(dt).permute(0, 3, 1, 2))
        c2 = pyramid(img2.to(dt).permute(0, 3, 1, 2))
        up_flow = up_feat = flow = None
        for lvl in range(PYR_LVLS, FLOW_PRED_LVL - 1, -1):
            if lvl == PYR_LVLS:
                x = cost_volume(c1[lvl], c2[lvl])
            else:
                warp = dense_image_warp(c2[lvl], up_flow * (20.0 / 2 ** lvl))
                x = torch.cat([cost_volume(c1[lvl], warp), c1[lvl], up_flow, up_feat], dim=1)
            for k in range(len(DENSE)):
                x = torch.cat([lrelu(conv_same(x, *P(f"pwcnet/predict_flow/conv{lvl}_{k}"))), x], dim=1)
            upfeat = x
            flow = conv_same(upfeat, *P(f"pwcnet/predict_flow/flow{lvl}"))
            y = upfeat
            for k, (_, d) in enumerate(CTXT):
                y = conv_same(y, *P(f"pwcnet/ctxt/dc_conv{lvl}{k + 1}"), dilation=d)
                if k + 1 < len(CTXT):
                    y = lrelu(y)
            flow = flow + y
            if taps is not None:
                taps[f"flow{lvl}"] = flow.permute(0, 2, 3, 1)
            if lvl != FLOW_PRED_LVL:
                up_flow = deconv_same(flow, *P(f"pwcnet/upsample/up_flow{lvl}"))
                up_feat = deconv_same(upfeat, *P(f"pwcnet/upsample/up_feat{lvl}"))
        scaler = 2 ** FLOW_PRED_LVL
        return (resize_bilinear_legacy(flow, scaler) * scaler).permute(0, 2, 3, 1).contiguous()


# ------------------------------------------------------------------------------------------ pre / post (host side in the reference too)
def skimage_resize(img: np.ndarray, out_hw: Tuple[int, int], anti_aliasing: bool = False) -> np.ndarray:
    """``skimage.transform.resize(img, out_shape)`` for [..., H, W, C]-like arrays resized along two axes (H, W = axes -3, -2):
    order 1, mode 'reflect', optional Gaussian pre-filter of sigma (factor - 1) / 2 per down-scaled axis."""
    from scipy import ndimage as ndi
    img = np.asarray(img, dtype=np.float64)
    h, w = img.shape[-3], img.shape[-2]
    oh, ow = out_hw
    if anti_aliasing:
        sig = [0.0] * img.ndim
        sig[-3], sig[-2] = max(0.0, (h / oh - 1) / 2), max(0.0, (w / ow - 1) / 2)
        if any(sig):
            img = ndi.gaussian_filter(img, sig, mode="mirror")
    ys = (np.arange(oh) + 0.5) * (h / oh) - 0.5
    xs = (np.arange(ow) + 0.5) * (w / ow) - 0.5

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

        a0, a1 = np.take(a, mirror(i0), axis=axis), np.take(a, mirror(i0 + 1), axis=axis)
        shape = [1] * a.ndim
        shape[axis] = -1
        fr = fr.reshape(shape)
        return a0 * (1 - fr) + a1 * fr

    return lerp_axis(lerp_axis(img, ys, img.ndim - 3), xs, img.ndim - 2)


def prepare_pair(rgb1: np.ndarray, rgb2: np.ndarray, scale: int = 2):
    """..predict_from_img_test.py:126-131 + adapt_x: x2 resize, uint8 truncation, /255, zero pad to multiples of 64.
    rgb float [h,w,3] 0..255 -> (img1, img2) float32 [H,W,3], and the unpadded (H0, W0)."""
    h, w = rgb1.shape[:2]
    outs = []
    for a in (rgb1, rgb2):
        u8 = np.array(skimage_resize(a, (h * scale, w * scale)), dtype=np.uint8)
        x = u8.astype(np.float32) / np.float32(255.)
        ph, pw = (-x.shape[0]) % 64, (-x.shape[1]) % 64
        outs.append(np.pad(x, [(0, ph), (0, pw), (0, 0)], mode="constant"))
    return outs[0], outs[1], (h * scale, w * scale)


def finish_flow(flow: np.ndarray, hw0: Tuple[int, int], out_hw: Tuple[int, int], scale: int = 2) -> np.ndarray:
    """postproc_y_hat_test crop + ..predict_from_img_test.py:137: anti-aliased resize back to (h, w), divided by the scale."""
    flow = flow[:hw0[0], :hw0[1]]
    return (skimage_resize(flow, out_hw, anti_aliasing=True) / scale).astype(np.float32)
