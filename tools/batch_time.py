"""ms per window of the tiled video path when B windows (4 B tiles) run as one batched forward."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch, fisr_b200
from fisr_b200.init import xavier_params
eng = fisr_b200.Engine(0, precision="f16f8")
eng.set_params(xavier_params(0, 0.01))
g = torch.Generator().manual_seed(1)
for B in (1, 2, 3, 4):
    frames = torch.randint(0, 256, (B, 1080, 1920, 9), dtype=torch.uint8, generator=g).cuda()
    flow = (torch.randn(B, 1080, 1920, 8, generator=g) * 4).cuda()
    warp = torch.rand(B, 1080, 1920, 12, generator=g).cuda()
    out = torch.zeros((B, 2048, 3840, 9), dtype=torch.uint8, device="cuda")
    units = list(range(4 * B))
    for _ in range(3): eng.units(frames, flow, warp, units, (2, 2), layout="frames", out=out)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(10): eng.units(frames, flow, warp, units, (2, 2), layout="frames", out=out)
    e1.record(); torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 10
    print(f"B={B} windows per forward ({4*B} tiles): {ms:.3f} ms per step, {ms/B:.3f} ms per window, {2000*B/ms:.2f} frames/s, workspace {eng.plan_info(4*B,544,992)['workspace_bytes']/1e9:.1f} GB", flush=True)
    del frames, flow, warp, out
    eng.set_precision("f16"); eng.set_precision("f16f8")     # free the plan before the next size
