"""ncu / timing target: one PWC-Net forward on both directions of a 1080p pair after the x2 pre-upscale (2 x 2176 x 3840)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch
from fisr_b200.pwcnet import PWCNet
from oracle import pwcnet_oracle as W
net = PWCNet(0); net.set_params(W.init_params(0))
a = torch.rand(2, 2176, 3840, 3).cuda(); b = torch.rand(2, 2176, 3840, 3).cuda()
for _ in range(3): f = net.forward(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): f = net.forward(a, b)
e1.record()
torch.cuda.synchronize()
print(f"PWC-Net 2 x 2176 x 3840 (FISR_PWC_UMMA={os.environ.get('FISR_PWC_UMMA', 'default')}): {e0.elapsed_time(e1) / 10:.2f} ms per forward, mean |flow| {float(f.abs().mean()):.5f}")
