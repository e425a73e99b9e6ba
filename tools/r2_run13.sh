#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu 2>&1 | tail -3
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 600 python bench.py --no-extras > gpurun_out/r02_bench_quick.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r02_bench_quick.json')); print('N=1 value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2), 'launches', d['gpu_launches'])"
