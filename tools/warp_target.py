"""ncu / timing target for the flow-warp kernel: the 16 warps of a 9-frame 1080p clip in one launch (fisr_warp_batch_device)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fisr_b200  # noqa: E402

eng = fisr_b200.Engine(0)
g = torch.Generator().manual_seed(5)
nfr, H, W = 9, 1080, 1920
jobs = 2 * (nfr - 1)
sigma = float(sys.argv[1]) if len(sys.argv) > 1 else 4.0
yuv = torch.randint(0, 256, (nfr, H, W, 3), dtype=torch.uint8, generator=g).cuda()
flo = (torch.randn(jobs, H, W, 2, generator=g) * sigma).cuda()
src = [fr + 1 - (j & 1) for fr in range(nfr - 1) for j in range(2)]
out = torch.empty((jobs, H, W, 3), dtype=torch.float32, device='cuda')
for _ in range(3):
    eng.warp_batch(yuv, flo, src, 0.5, 1 / 255., out=out)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
for _ in range(10):
    eng.warp_batch(yuv, flo, src, 0.5, 1 / 255., out=out)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / 10
by = jobs * H * W * 23
print(f"warp batch (flow sigma {sigma} px): {ms:.4f} ms per {jobs} warps = {ms * 1e3 / jobs:.2f} us per 1080p warp, {by / ms / 1e6:.0f} GB/s (23 B/px)")
