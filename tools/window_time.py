"""In-process A/B of plan-time settings on the bench workload (4 tiles of 544x992, CUDA graph replay, inputs in HBM) and on
config 2 (8x192x192).  Variants are environment settings the planner reads (getenv at plan time; add one temporarily for the
experiment at hand -- none is left in the library); they are measured interleaved, several rounds, in ONE process (clock /
power state drifts by several percent between processes and boxes), best and median reported.

    python tools/window_time.py f16f8 "" "FISR_PAIR=1"
"""
import os
import statistics
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import torch  # noqa: E402
import fisr_b200  # noqa: E402
from fisr_b200.init import xavier_params  # noqa: E402

prec = sys.argv[1] if len(sys.argv) > 1 else "f16f8"
variants = sys.argv[2:] or [""]
rounds = int(os.environ.get("AB_ROUNDS", "3"))
eng = fisr_b200.Engine(0, precision=prec)
eng.set_params(xavier_params(0, 0.01))
g = torch.Generator().manual_seed(1)
frames = torch.randint(0, 256, (1080, 1920, 9), dtype=torch.uint8, generator=g).cuda()
flow = (torch.randn(1080, 1920, 8, generator=g) * 4).cuda()
warp = torch.rand(1080, 1920, 12, generator=g).cuda()
out = torch.zeros((2048, 3840, 9), dtype=torch.uint8, device="cuda")
x = torch.rand(8, 192, 192, 29, generator=g).cuda()


def timed(fn, warm=3, reps=15):
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


res = {v: ([], [], None) for v in variants}
for r in range(rounds):
    for v in variants:
        keys = []
        for kv in v.split():
            k, val = kv.split("=")
            os.environ[k] = val
            keys.append(k)
        other = "f16" if prec != "f16" else "f16x3"
        eng.set_precision(other)
        eng.set_precision(prec)               # drops the cached plans: geometry is re-planned under this variant's settings
        w_ms = timed(lambda: eng.window(frames, flow, warp, (2, 2), out=out))
        c_ms = timed(lambda: eng.forward(x))
        res[v][0].append(w_ms)
        res[v][1].append(c_ms)
        res[v] = (res[v][0], res[v][1], int(out.sum()))
        for k in keys:
            del os.environ[k]
for v in variants:
    w, c, chk = res[v]
    print(f"{prec} [{v or 'defaults'}]: window best {min(w):.3f} median {statistics.median(w):.3f} ms ({2000 / min(w):.2f} frames/s); "
          f"config 2 best {min(c):.3f} median {statistics.median(c):.3f} ms; checksum {chk}")
