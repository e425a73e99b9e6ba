"""Per-launch device time of one batched forward (CUDA events inside libfisr_b200): the optimisation worklist."""
import json, sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import fisr_b200
from fisr_b200.init import xavier_params

n, h, w = (int(a) for a in sys.argv[1:4]) if len(sys.argv) >= 4 else (4, 544, 992)
prec = sys.argv[4] if len(sys.argv) > 4 else "f16x3"
eng = fisr_b200.Engine(0, precision=prec)
eng.set_params(xavier_params(0, 0.01))
ops = eng.profile_ops(n, h, w, reps=3)
mult = {"f16x3": 3, "f16f8": 2}.get(prec, 1)
tot = sum(o["ms"] for o in ops)
print(f"# plan {n}x{h}x{w} {prec}: {len(ops)} launches, {tot:.3f} ms launch-by-launch, "
      f"{sum(o['flops'] for o in ops)/tot/1e9:.1f} TFLOP/s algorithmic")
print(f"{'name':58s} {'kind':8s} {'ms':>8s} {'%':>5s} {'TF/s alg':>9s} {'TF/s iss':>9s} {'GB/s':>8s}")
agg = {}
for o in ops:
    tf = o["flops"] / o["ms"] / 1e9 if o["ms"] > 0 else 0
    gb = o["bytes"] / o["ms"] / 1e6 if o["ms"] > 0 else 0
    print(f"{o['name'][-58:]:58s} {o['kind']+str(o['nt'] or ''):8s} {o['ms']:8.4f} {100*o['ms']/tot:5.1f} {tf:9.1f} {tf*mult:9.1f} {gb:8.0f}")
json.dump({"plan": [n, h, w], "precision": prec, "total_ms": tot, "ops": ops}, open(f"gpurun_out/layers_{n}x{h}x{w}_{prec}.json", "w"))
