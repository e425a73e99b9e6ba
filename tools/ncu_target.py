"""Short target for ncu: two batched forwards (no CUDA graph) of the bench workload's tile batch."""
import os, sys
os.environ["FISR_NO_GRAPH"] = "1"
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fisr_b200
from fisr_b200.init import xavier_params
n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 544, 992)
eng = fisr_b200.Engine(0)
eng.set_params(xavier_params(0, 0.01))
x = torch.rand(n, h, w, 29, device="cuda")
for _ in range(2):
    eng.forward(x, want=(False, False, False))
torch.cuda.synchronize()
