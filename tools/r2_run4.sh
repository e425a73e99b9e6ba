#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_f16f8.py tests/test_gpu_window.py tests/test_gpu_conv.py tests/test_gpu_video.py -q -m gpu -x 2>&1 | tail -3
(
FISR_CHUNKS=0 python tools/window_time.py f16f8
python tools/window_time.py f16f8
python tools/window_time.py f16x3
) 2>&1 | grep -v Warning | tee gpurun_out/r2_ab_run4.txt
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_tile_f16f8_c.txt 2>&1; head -1 gpurun_out/r2_layers_tile_f16f8_c.txt
timeout 300 python tools/profile_layers.py 8 192 192 f16f8 > gpurun_out/r2_layers_cfg2_f16f8_c.txt 2>&1; head -1 gpurun_out/r2_layers_cfg2_f16f8_c.txt
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, '.')
import fisr_b200
eng = fisr_b200.Engine(0)
g = torch.Generator().manual_seed(5)
nfr, H, W = 9, 1080, 1920
jobs = 2 * (nfr - 1)
yuv = torch.randint(0, 256, (nfr, H, W, 3), dtype=torch.uint8, generator=g).cuda()
flo = (torch.randn(jobs, H, W, 2, generator=g) * 4).cuda()
src = [fr + 1 - (j & 1) for fr in range(nfr - 1) for j in range(2)]
for _ in range(3): eng.warp_batch(yuv, flo, src, 0.5, 1 / 255.)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10): eng.warp_batch(yuv, flo, src, 0.5, 1 / 255.)
e1.record(); torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
by = jobs * H * W * 23
print(f"warp batch: {ms:.4f} ms per 16 warps = {ms*1e3/jobs:.2f} us per 1080p warp, {by/ms/1e6:.0f} GB/s (23 B/px)")
PY
