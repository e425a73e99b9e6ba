#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_pwcnet.py -q -m gpu -x 2>&1 | tail -2
start=$(date +%s)
timeout 900 python bench.py > gpurun_out/r02_bench_n1_b.json 2> gpurun_out/r02_bench_n1_b.err; echo "bench rc=$? in $(( $(date +%s) - start )) s"
python - <<'PY'
import json
d=json.load(open('gpurun_out/r02_bench_n1_b.json'))
print('value',round(d['value'],2),'ms',round(d['ms_per_step'],3),'e2e',round(d['e2e']['value'],2),'frac',round(d['roofline']['frac'],3))
for k,v in d['extra'].items(): print(k, {a:(round(b,3) if isinstance(b,float) else b) for a,b in v.items() if a not in('note','input')})
PY
tail -2 gpurun_out/r02_bench_n1_b.err
