"""Small workload for compute-sanitizer: one conv of each kernel family, a whole small forward in every precision mode, a tiled
window and one training step."""
import os, sys
os.environ["FISR_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fisr_b200
from fisr_b200.init import xavier_params
eng = fisr_b200.Engine(0)
g = torch.Generator().manual_seed(0)
def conv(n, h, w, ci, co, res=True, d2s=False):
    x = torch.rand(n, h, w, ci, generator=g).cuda(); wt = (torch.randn(3, 3, ci, co, generator=g) * 0.05).cuda()
    b = torch.zeros(co).cuda(); r = torch.rand(n, h, w, co, generator=g).cuda() if res else None
    eng.conv3x3(x, wt, b, r, relu=True, d2s=d2s, want_raw=not d2s); torch.cuda.synchronize(); print("ok", eng.precision, n, h, w, ci, co, res, d2s, flush=True)
for prec in ("f16x3", "f16f8", "f16"):
    eng.set_precision(prec)
    conv(1, 72, 120, 128, 128)
    conv(1, 40, 56, 64, 64)
    conv(1, 17, 31, 64, 64, res=False)
    conv(1, 24, 40, 64, 256, res=False, d2s=True)
    conv(1, 40, 56, 64, 6, res=False)
    eng.set_params(xavier_params(0, 0.01))
    for shape in ((1, 64, 96), (2, 128, 128)):
        out = eng.forward(torch.rand(*shape, 29, generator=g).cuda()); torch.cuda.synchronize(); print("forward ok", prec, shape, float(out[2].abs().max()), flush=True)
    fr = torch.randint(0, 256, (136, 200, 9), generator=g, dtype=torch.uint8).cuda()
    out = eng.window(fr, torch.randn(136, 200, 8, generator=g).cuda(), torch.rand(136, 200, 12, generator=g).cuda(), (2, 2))
    torch.cuda.synchronize(); print("window ok", prec, tuple(out.shape), flush=True)
eng.set_precision("f16x3")
mk = lambda c, s=32: torch.rand(1, s, s, c, generator=g).cuda()
s = eng.train_backward(mk(15), (mk(16) - 0.5) * 0.1, (mk(8) - 0.5) * 0.1, mk(24), mk(12), mk(21, 64))
torch.cuda.synchronize(); print("train ok", s["total_loss"], flush=True)
s = eng.train_step(mk(15), (mk(16) - 0.5) * 0.1, (mk(8) - 0.5) * 0.1, mk(24), mk(12), mk(21, 64), lr=1e-4)      # + multi-tensor Adam / re-pack
torch.cuda.synchronize(); print("train step ok", s["total_loss"], eng.adam_steps, flush=True)
st = eng.get_adam_state(); eng.set_adam_state(st["m"], st["v"], st["t"]); eng.adam_reset(0)
for w in (200, 190):                                                    # word-gather and byte-gather paths of the warp kernel
    fr = torch.randint(0, 256, (3, 72, w, 3), generator=g, dtype=torch.uint8).cuda()
    fl = (torch.randn(4, 72, w, 2, generator=g) * 30).cuda()            # far out-of-frame taps: BORDER_REPLICATE
    out = eng.warp_batch(fr, fl, [1, 0, 2, 1], 0.5, 1.0); torch.cuda.synchronize(); print("warp ok", w, float(out.mean()), flush=True)
from fisr_b200.pwcnet import PWCNet
from oracle import pwcnet_oracle as W
net = PWCNet(0); net.set_params(W.init_params(0))
f = net.forward(torch.rand(2, 64, 128, 3, generator=g).cuda(), torch.rand(2, 64, 128, 3, generator=g).cuda())
torch.cuda.synchronize(); print("pwcnet ok", float(f.abs().mean()), flush=True)
f = net.forward(torch.rand(1, 256, 320, 3, generator=g).cuda(), torch.rand(1, 256, 320, 3, generator=g).cuda())      # every dilation as polyphase launches
torch.cuda.synchronize(); print("pwcnet polyphase ok", float(f.abs().mean()), flush=True)
import numpy as np
yuv = np.random.default_rng(0).integers(0, 256, (2, 43, 61, 3), dtype=np.uint8)                                     # driver pre / post-processing kernels
f = net.flow_pair_yuv(yuv[0], yuv[1]); print("flow_pair ok", f.shape, float(np.abs(f).mean()), flush=True)
net.close()
