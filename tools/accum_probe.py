"""Does the tcgen05 fp32 accumulator round or truncate?  Sums of identical positive products through the production
conv and wgrad kernels vs the exact value (products of fp16 values and their integer multiples are exact in fp64)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fisr_b200  # noqa: E402

eng = fisr_b200.Engine(0)


def split(v):
    hi = torch.tensor(v, dtype=torch.float32).half()
    lo = (torch.tensor(v, dtype=torch.float32) - hi.float()).half()
    return hi.double() + lo.double()


for cin in (64, 512):
    xv, wv = 0.7373, 0.0213
    x = torch.full((1, 16, 16, cin), xv).cuda()
    w = torch.full((3, 3, cin, 64), wv).cuda()
    b = torch.zeros(64).cuda()
    raw, _ = eng.conv3x3(x, w, b, None, relu=False, want_act=False)
    xs, ws = split(xv), split(wv)
    exact = float(9 * cin * (xs * ws - (xs - split(xv).half().double() if False else 0)))
    # the kernel drops lo*lo: hi*hi + lo*hi + hi*lo
    xh = torch.tensor(xv).half().double(); xl = xs - xh
    wh = torch.tensor(wv).half().double(); wl = ws - wh
    exact = float(9 * cin * (xh * wh + xl * wh + xh * wl))
    got = float(raw[0, 8, 8, 0])
    print(f"conv Cin={cin}: K={9 * cin} interior value {got:.9f} exact {exact:.9f} rel err {(got - exact) / exact:+.3e} "
          f"(fp32 eps 6e-8; {9 * cin // 16 * 3} accumulating MMAs)")

for n, hw in ((1, 64), (4, 192), (16, 192)):
    xv, dv = 0.7373, 0.0213
    x = torch.full((n, hw, hw, 64), xv).cuda()
    dy = torch.full((n, hw, hw, 64), dv).cuda()
    gw, gb = eng.wgrad3x3(x, dy)
    exact = float(n * hw * hw * split(xv) * split(dv))
    got = float(gw[1, 1, 0, 0])
    print(f"wgrad {n}x{hw}x{hw}: centre tap {got:.6f} exact {exact:.6f} rel err {(got - exact) / exact:+.3e}; "
          f"bias {float(gb[0]):.6f} exact {float(n * hw * hw * split(dv)):.6f}")
