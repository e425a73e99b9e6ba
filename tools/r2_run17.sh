#!/bin/bash
timeout 600 python -m pytest tests/test_gpu_pwcnet.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
from fisr_b200.pwcnet import PWCNet
from oracle import pwcnet_oracle as W
net = PWCNet(0); net.set_params(W.init_params(0))
a = torch.rand(2, 2176, 3840, 3).cuda(); b = torch.rand(2, 2176, 3840, 3).cuda()
for _ in range(2): f = net.forward(a, b)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(3): f = net.forward(a, b)
e1.record(); torch.cuda.synchronize()
print(f"PWC-Net 2 x 2176x3840: {e0.elapsed_time(e1)/3:.1f} ms per forward")
PY
