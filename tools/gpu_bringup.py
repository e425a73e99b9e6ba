"""Bring-up on the GPU box: single-conv checks across shapes/epilogues, then whole-model parity with per-layer bisect."""
import sys, time, os
sys.path.insert(0, '.')
import numpy as np, torch
import torch.nn.functional as F
from oracle import fisrnet_oracle as O
import fisr_b200

torch.backends.cuda.matmul.allow_tf32 = False
torch.backends.cudnn.allow_tf32 = False
eng = fisr_b200.Engine(0)
dev = torch.device('cuda:0')

def ref_conv(x, w, b, res=None):
    y = F.conv2d(x.double().permute(0, 3, 1, 2), w.double().permute(3, 2, 0, 1), b.double(), padding=1).permute(0, 2, 3, 1)
    if res is not None: y = y + res.double()
    return y

def check_conv(n, h, w, cin, cout, res=False, relu=True, d2s=False, seed=0):
    g = torch.Generator().manual_seed(seed)
    x = torch.rand(n, h, w, cin, generator=g)
    wt = torch.randn(3, 3, cin, cout, generator=g) * (2.0 / (9 * cin)) ** 0.5
    b = torch.randn(cout, generator=g) * 0.1
    r = torch.randn(n, h, w, cout, generator=g) if res else None
    y = ref_conv(x, wt, b, r)
    raw, act = eng.conv3x3(x.to(dev), wt.to(dev), b.to(dev), r.to(dev) if res else None, relu=relu, d2s=d2s,
                           want_raw=cout > 16 or True, want_act=True)
    e_raw = (raw.cpu().double() - y).abs().max().item()
    ya = torch.relu(y) if relu else y
    if d2s:
        ya = O.to_nhwc(O.depth_to_space2(O.to_nchw(ya)))
    e_act = (act.cpu().double() - ya).abs().max().item()
    ok = e_raw < 2e-5 * max(1, y.abs().max().item()) and e_act < 2e-5 * max(1, y.abs().max().item()) if eng.precision == 'f16x3' else e_raw < 2e-2
    print(f"conv n{n} {h}x{w} {cin}->{cout} res={int(res)} relu={int(relu)} d2s={int(d2s)}: raw {e_raw:.2e} act {e_act:.2e} {'OK' if ok else 'FAIL'}", flush=True)
    return ok

allok = True
for prec in ('f16x3', 'f16'):
    eng.set_precision(prec)
    print('precision', prec)
    cases = [(1, 8, 16, 64, 64), (2, 32, 32, 64, 64), (1, 24, 40, 64, 128), (1, 16, 16, 128, 128), (1, 12, 20, 256, 256),
             (1, 6, 6, 512, 512), (1, 17, 31, 256, 512), (2, 64, 96, 29, 64), (1, 48, 48, 38, 64), (1, 64, 64, 64, 6), (1, 33, 47, 64, 3),
             (8, 96, 96, 64, 64), (1, 136, 248, 128, 128)]
    for c in cases:
        allok &= check_conv(*c, res=c[4] > 16, relu=True)
    
    allok &= check_conv(1, 16, 24, 64, 64, relu=False)
eng.set_precision('f16x3')
print('single conv checks', 'PASS' if allok else 'FAIL', flush=True)

# whole model
for (n, hh, ww, seed) in ((1, 96, 96, 0), (2, 64, 160, 1), (8, 192, 192, 1), (1, 544, 992, 2), (4, 544, 992, 2)):
    p32 = O.init_params(seed)
    eng.set_params(p32)
    x = O.synthetic_input(n, hh, ww, seed + 10)
    tap = {}
    if n == 4:
        ref = None; t_cpu = 0
    else:
        t = time.time(); ref = O.model(p32, x, tap=tap); t_cpu = time.time() - t
    t = time.time(); out = eng.forward(x.to(dev)); torch.cuda.synchronize(); t_gpu = time.time() - t
    errs = [(a.cpu() - b).abs().max().item() for a, b in zip(out, ref)] if ref is not None else [0, 0, 0]
    print(f"model n{n} {hh}x{ww}: maxabs l1/l2/l3 {errs[0]:.2e} {errs[1]:.2e} {errs[2]:.2e}  cpu {t_cpu:.2f}s gpu(first) {t_gpu:.3f}s", flush=True)
    if max(errs) > 1e-4 or any(np.isnan(e) for e in errs):
        shown = 0
        for name, refv in tap.items():
            try:
                got = eng.debug_conv_output(name, tuple(refv.shape))
            except Exception as ex:
                continue
            e = np.abs(got - refv.numpy()).max()
            if e > 1e-4 or np.isnan(e):
                print(f"   first bad conv with fp32 output: {name} maxabs {e:.3e} shape {tuple(refv.shape)}")
                shown += 1
                if shown >= 3: break
    print(eng.plan_info(n, hh, ww))
    # timing
    xd = x.to(dev)
    for _ in range(3): eng.forward(xd)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.forward(xd)
    e1.record(); torch.cuda.synchronize()
    print(f"   gpu steady {e0.elapsed_time(e1)/10:.3f} ms/forward", flush=True)
