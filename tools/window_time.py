"""ms per batched forward of the bench workload (4 tiles of 544x992, CUDA graph replay, inputs in HBM) and of config 2:
the A/B number for kernel / plan changes (toggle with FISR_NO_PDL / FISR_NO_POOL_FUSION / FISR_NO_HEAD_MERGE)."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fisr_b200  # noqa: E402
from fisr_b200.init import xavier_params  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
eng = fisr_b200.Engine(0, precision=prec)
eng.set_params(xavier_params(0, 0.01))
g = torch.Generator().manual_seed(1)
frames = torch.randint(0, 256, (1080, 1920, 9), dtype=torch.uint8, generator=g).cuda()
flow = (torch.randn(1080, 1920, 8, generator=g) * 4).cuda()
warp = torch.rand(1080, 1920, 12, generator=g).cuda()
out = torch.zeros((2048, 3840, 9), dtype=torch.uint8, device="cuda")


def timed(fn, warm=3, reps=20):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


ms = timed(lambda: eng.window(frames, flow, warp, (2, 2), out=out))
x = torch.rand(8, 192, 192, 29, generator=g).cuda()
ms2 = timed(lambda: eng.forward(x))
flags = " ".join(f"{k}={os.environ[k]}" for k in ("FISR_NO_PDL", "FISR_NO_POOL_FUSION", "FISR_NO_HEAD_MERGE", "FISR_HEAD_F16") if k in os.environ)
print(f"{prec} [{flags or 'defaults'}]: window (4 x 544x992) {ms:.3f} ms = {2000 / ms:.2f} frames/s; config 2 (8x192x192) {ms2:.3f} ms; "
      f"checksum {int(out.sum())}")
