"""Small workload for compute-sanitizer: one conv of each kernel family + a whole small forward."""
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
    eng.conv3x3(x, wt, b, r, relu=True, d2s=d2s); torch.cuda.synchronize(); print("ok", n, h, w, ci, co, res, d2s, flush=True)
conv(1, 72, 120, 128, 128)
conv(1, 40, 56, 64, 64)
conv(1, 24, 40, 64, 256, res=False, d2s=True)
conv(1, 40, 56, 64, 6, res=False)
eng.set_params(xavier_params(0, 0.01))
for shape in ((1, 64, 96), (2, 128, 128)):
    out = eng.forward(torch.rand(*shape, 29, generator=g).cuda()); torch.cuda.synchronize(); print("forward ok", shape, float(out[2].abs().max()), flush=True)
