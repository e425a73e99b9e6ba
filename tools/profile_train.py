"""Per-op table of one training step (forward ops + backward ops) for batch B at LR size h x w (default config 3)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fisr_b200  # noqa: E402
from fisr_b200.init import xavier_params as init_params  # noqa: E402

B, h, w = (int(a) for a in (sys.argv[1:4] + ["16", "192", "192"][len(sys.argv) - 1:]))
eng = fisr_b200.Engine(0)
eng.set_params(init_params(0))
g = torch.Generator().manual_seed(2)
mk = lambda c, hh=h, ww=w: torch.rand(B, hh, ww, c, generator=g).cuda()
batch = (mk(15), (mk(16) - 0.5) * 0.1, (mk(8) - 0.5) * 0.1, mk(24), mk(12), mk(21, 2 * h, 2 * w))
torch.cuda.synchronize()
t0 = time.time()
s = eng.train_backward(*batch)
torch.cuda.synchronize()
print(f"# first call (plan build + run): {time.time() - t0:.2f} s, total_loss {s['total_loss']:.5f}")
ev0, ev1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
reps = 3
ev0.record()
for _ in range(reps):
    eng.train_backward(*batch)
ev1.record()
torch.cuda.synchronize()
step_ms = ev0.elapsed_time(ev1) / reps
fwd = eng.profile_ops(4 * B, h, w, reps=2)
bwd = eng.profile_train(B, h, w, reps=2)
f_ms = sum(o["ms"] for o in fwd)
f_fl = sum(o["flops"] for o in fwd)
kinds = {}
for o in bwd:
    k = kinds.setdefault(o["kind"], [0.0, 0.0, 0])
    k[0] += o["ms"]; k[1] += o["flops"]; k[2] += 1
b_ms = sum(o["ms"] for o in bwd)
print(f"# train step B={B} {h}x{w}: forward+loss+backward {step_ms:.2f} ms/step (no Adam); forward ops {f_ms:.2f} ms "
      f"({f_fl / f_ms / 1e9:.1f} TFLOP/s alg); backward ops {b_ms:.2f} ms")
for k, (ms, fl, n) in kinds.items():
    print(f"#   backward {k:6s}: {n:4d} ops {ms:8.2f} ms  {fl / max(ms, 1e-9) / 1e9:8.1f} TFLOP/s alg")
print(f"{'name':70s} {'kind':6s} {'ms':>8s} {'TF/s alg':>9s}")
for o in bwd:
    print(f"{o['name']:70s} {o['kind']:6s} {o['ms']:8.4f} {o['flops'] / max(o['ms'], 1e-9) / 1e9:9.1f}")
