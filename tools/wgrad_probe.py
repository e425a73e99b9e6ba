"""GPU bring-up of the MN-major wgrad kernel: runs one small case per descriptor variant in a fresh process."""
import os
import subprocess
import sys

CASE = r'''
import os, sys, torch
sys.path.insert(0, os.getcwd())
import torch.nn.functional as F
import fisr_b200
eng = fisr_b200.Engine(0)
for shape in [(1, 8, 16, 64, 64), (2, 32, 48, 64, 64), (1, 16, 16, 128, 128)]:
    n, h, w, cin, cout = shape
    g = torch.Generator().manual_seed(1)
    x = torch.rand(n, h, w, cin, generator=g)
    dy = torch.randn(n, h, w, cout, generator=g) * 0.05
    wz = torch.zeros(cout, cin, 3, 3, dtype=torch.float64, requires_grad=True)
    y = F.conv2d(x.double().permute(0, 3, 1, 2), wz, padding=1)
    (gw_ref,) = torch.autograd.grad(y, wz, dy.double().permute(0, 3, 1, 2))
    gw_ref = gw_ref.permute(2, 3, 1, 0)
    try:
        gw, gb = eng.wgrad3x3(x.cuda(), dy.cuda())
        err = float((gw.cpu().double() - gw_ref).abs().max())
        print("variant", os.environ.get("FISR_WGRAD_VARIANT"), shape, "max-abs err %.3e (ref max %.3e)" % (err, float(gw_ref.abs().max())),
              "PASS" if err < 1e-4 else "FAIL", flush=True)
    except Exception as e:
        print("variant", os.environ.get("FISR_WGRAD_VARIANT"), shape, "ERROR", e, flush=True)
        break
'''
for v in sys.argv[1:] or ["0", "1"]:
    env = dict(os.environ, FISR_WGRAD_VARIANT=v)
    try:
        subprocess.run([sys.executable, "-c", CASE], env=env, timeout=120)
    except subprocess.TimeoutExpired:
        print("variant", v, "TIMEOUT", flush=True)
