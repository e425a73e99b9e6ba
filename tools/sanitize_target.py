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
