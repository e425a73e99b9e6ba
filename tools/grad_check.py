"""Per-tensor gradient error table of the CUDA backward pass vs float64 autograd on the oracle (bring-up / debugging)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fisr_b200  # noqa: E402
from oracle import fisrnet_oracle as O  # noqa: E402
from oracle import loss_oracle as L  # noqa: E402

B, h, w, seed = (int(a) for a in (sys.argv[1:5] + ["1", "32", "32", "31"][len(sys.argv) - 1:]))
g = torch.Generator().manual_seed(seed + 100)
data = torch.rand(B, h, w, 15, generator=g)
flow = (torch.randn(B, h, w, 16, generator=g) * 4 / 96 / 2).clamp(-1, 1)
flow2 = (torch.randn(B, h, w, 8, generator=g) * 8 / 96 / 2).clamp(-1, 1)
warp = torch.rand(B, h, w, 24, generator=g)
warp2 = torch.rand(B, h, w, 12, generator=g)
label = torch.rand(B, 2 * h, 2 * w, 21, generator=g)
batch = (data, flow, flow2, warp, warp2, label)
params = O.init_params(seed)
eng = fisr_b200.Engine(0)
eng.set_params(params)
if os.environ.get("FISR_WGRAD_EXACT") == "1":
    eng.set_wgrad_exact(True)
p64 = {k: v.double() for k, v in params.items()}
ref_s, _, ref_g = L.training_forward(p64, *[t.double() for t in batch], grad=True)
got_s = eng.train_backward(*[t.cuda() for t in batch])
print("total_loss", got_s["total_loss"], float(ref_s["total_loss"]))
got = eng.get_grads()
_, _, g32 = L.training_forward(params, *batch, grad=True)          # torch fp32 autograd: the reference-class noise floor
num = den = num32 = 0.0
print(f"{'tensor':60s} {'max-rel':>9s} {'l2-rel':>9s} | fp32 oracle: {'max-rel':>9s} {'l2-rel':>9s}")
for k, r in ref_g.items():
    r = r.numpy()
    d = got[k] - r
    d32 = g32[k].numpy() - r
    e = np.abs(d).max() / max(np.abs(r).max(), 1e-30)
    e32 = np.abs(d32).max() / max(np.abs(r).max(), 1e-30)
    l2, l232 = np.linalg.norm(d) / np.linalg.norm(r), np.linalg.norm(d32) / np.linalg.norm(r)
    num += float((d.astype(np.float64) ** 2).sum()); den += float((r ** 2).sum()); num32 += float((d32.astype(np.float64) ** 2).sum())
    print(f"{k:60s} {e:9.2e} {l2:9.2e} | {e32:9.2e} {l232:9.2e}{'  <<<' if e > 1e-3 else ''}")
print(f"whole-gradient relative L2 error: cuda {np.sqrt(num / den):.3e}   torch-fp32 {np.sqrt(num32 / den):.3e}")
