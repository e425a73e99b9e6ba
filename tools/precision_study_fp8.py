"""Design study (not product): error of the "fp16 main term + fp8 cross terms" operand scheme vs fp64.

  x = xh + xl, w = wh + wl  (fp16 pairs).  conv(x, w) ~= conv(xh, wh)                      [fp16 MMA, full rate]
                                                     + conv(e4m3(xl), e4m3(wh)) + conv(e4m3(xh), e4m3(wl))   [fp8 MMA, K-concatenated, 2x rate]
The cross terms are ~2^-11 of the main term, so 3 mantissa bits on them leave ~2^-15 relative error.
"""
import sys, time, torch
import torch.nn.functional as F
sys.path.insert(0, '.')
from oracle import fisrnet_oracle as O

F8 = torch.float8_e4m3fn

def e4m3(x, scale):
    """round x*scale to e4m3 (saturating), return the de-scaled value in x's dtype"""
    y = (x * scale).clamp(-448.0, 448.0).to(torch.float32).to(F8).to(x.dtype)
    return y / scale

def e5m2(x, scale):
    y = (x * scale).clamp(-57344.0, 57344.0).to(torch.float32).to(torch.float8_e5m2).to(x.dtype)
    return y / scale

class Cfg:
    mode = "f16+f8"
    sx = 1.0     # activation scale before e4m3
    sw = 64.0    # weight scale before e4m3
    lo_shift = 2048.0
    q8 = staticmethod(e4m3)

def conv_hook(self, x, name):
    w = self.p[name + "/w"]; b = self.p[name + "/b"]
    wk = w.permute(3, 2, 0, 1)
    dt = x.dtype
    xh = x.to(torch.float16).to(dt); xl = (x - xh).to(torch.float16).to(dt)
    wh = wk.to(torch.float16).to(dt); wl = (wk - wh).to(torch.float16).to(dt)
    conv = lambda a, c: F.conv2d(a, c, None, stride=1, padding=1)
    q8 = Cfg.q8
    if Cfg.mode == "f16":
        y = conv(xh, wh)
    elif Cfg.mode == "f16x3":
        y = conv(xh, wh) + conv(xl, wh) + conv(xh, wl)
    elif Cfg.mode == "f16x2a":      # activation split only
        y = conv(xh, wh) + conv(xl, wh)
    elif Cfg.mode == "built":       # the scheme as built (common.cuh): x = xh + xl; 8-bit row [e5m2(16 xl) | e5m2(xh)];
        # weights W = 128 w: fp16 plane Wh, 8-bit row [e4m3(Wh / 16) | e5m2(W - Wh)]; accumulator holds 128 x the result
        W = wk * 128.0
        Wh = W.to(torch.float16).to(dt); Wl = W - Wh
        y = (conv(xh, Wh) + conv(e5m2(xl, 16.0), e4m3(Wh, 1.0 / 16.0)) + conv(e5m2(xh, 1.0), e5m2(Wl, 1.0))) / 128.0
    elif Cfg.mode == "f16+f8":
        y = conv(xh, wh) + conv(q8(xl, Cfg.sx * Cfg.lo_shift), q8(wh, Cfg.sw)) + conv(q8(xh, Cfg.sx), q8(wl, Cfg.sw * Cfg.lo_shift))
    else:
        raise ValueError(Cfg.mode)
    y = y + b.view(1, -1, 1, 1)
    if self.tap is not None:
        self.tap[name] = O.to_nhwc(y)
    return y

H = int(sys.argv[1]) if len(sys.argv) > 1 else 96
N = int(sys.argv[2]) if len(sys.argv) > 2 else 1
orig = O.Net.conv
for seed in (0, 1):
    p64 = O.init_params(seed, torch.float64)
    x = O.synthetic_input(N, H, H, seed + 10)
    O.Net.conv = orig
    tap = {}
    ref = O.model(p64, x, tap=tap)
    amax = max(float(v.abs().max()) for v in tap.values())
    wmax = max(float(v.abs().max()) for k, v in p64.items() if k.endswith('/w'))
    print(f"seed {seed} H={H}: max |pre-activation| {amax:.2f}, max |w| {wmax:.3f}")
    O.Net.conv = conv_hook
    def rep(tag):
        outs = O.model(p64, x)
        errs = [(a.double() - b).abs().max().item() for a, b in zip(outs, ref)]
        print(f"  {tag:34s} maxabs l1/l2/l3 = {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}")
    for mode in ("f16", "f16x2a", "f16x3", "built"):
        Cfg.mode = mode; rep(mode)
    Cfg.mode = "f16+f8"
    for q8, qn in ((e4m3, "e4m3"), (e5m2, "e5m2")):
        for sx, sw in ((1.0, 64.0), (0.25, 64.0), (4.0, 256.0), (1 / 16.0, 16.0)):
            Cfg.q8 = staticmethod(q8); Cfg.sx = sx; Cfg.sw = sw
            rep(f"f16+{qn} sx={sx} sw={sw}")
O.Net.conv = orig
