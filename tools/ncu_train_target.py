"""ncu target: one training step (forward + loss + backward) for batch B at LR size h x w; first call builds the plan."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fisr_b200  # noqa: E402
from fisr_b200.init import xavier_params  # noqa: E402

B, h, w = (int(a) for a in (sys.argv[1:4] + ["16", "192", "192"][len(sys.argv) - 1:]))
steps = int(sys.argv[4]) if len(sys.argv) > 4 else 1
eng = fisr_b200.Engine(0)
eng.set_params(xavier_params(0))
g = torch.Generator().manual_seed(2)
mk = lambda c, hh=h, ww=w: torch.rand(B, hh, ww, c, generator=g).cuda()
batch = (mk(15), (mk(16) - 0.5) * 0.1, (mk(8) - 0.5) * 0.1, mk(24), mk(12), mk(21, 2 * h, 2 * w))
for _ in range(steps):
    s = eng.train_backward(*batch)
torch.cuda.synchronize()
print("total_loss", s["total_loss"])
