#!/bin/bash
timeout 1200 python -m pytest tests/test_gpu_pwcnet.py tests/test_gpu_video.py -x -q 2>&1 | tail -12
timeout 300 python - <<'P'
import time, numpy as np, torch
from fisr_b200.pwcnet import PWCNet
from oracle import pwcnet_oracle as W
net = PWCNet(0); net.set_params(W.init_params(0))
rng = np.random.default_rng(0)
y1 = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8); y2 = rng.integers(0, 256, (1080, 1920, 3), dtype=np.uint8)
net.flow_pair_yuv(y1, y2); torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(3): f = net.flow_pair_yuv(y1, y2)
print("flow_pair_yuv 1080p host to host: %.1f ms" % ((time.perf_counter() - t) / 3 * 1e3), f.shape)
a = torch.from_numpy(y1).cuda(); b = torch.from_numpy(y2).cuda()
for name, fn in (("prepare", lambda: net.prepare_pair(a, b, 2)),):
    fn(); torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
    e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize(); print(name, e0.elapsed_time(e1) / 5, "ms")
i1, i2 = net.prepare_pair(a, b, 2); fl = net.forward(i1, i2)
fn = lambda: net.finish_flow(fl, (2160, 3840), (1080, 1920), 2)
fn(); torch.cuda.synchronize(); e0 = torch.cuda.Event(enable_timing=True); e1 = torch.cuda.Event(enable_timing=True)
e0.record(); [fn() for _ in range(5)]; e1.record(); torch.cuda.synchronize(); print("finish", e0.elapsed_time(e1) / 5, "ms")
P
