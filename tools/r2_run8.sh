#!/bin/bash
mkdir -p gpurun_out
python tools/warp_target.py 4 2>&1 | grep -v Warn
timeout 300 python -m pytest tests/test_gpu_window.py -q -m gpu -x -k warp 2>&1 | tail -2
python bench.py --no-extras --steps 10 > gpurun_out/r2_bench_n1_quick.json 2>/dev/null; python -c "
import json; d=json.load(open('gpurun_out/r2_bench_n1_quick.json')); print('N=1 value', round(d['value'],2), 'ms', round(d['ms_per_step'],3), 'e2e', round(d['e2e']['value'],2))"
bash tools/gpu_multi.sh 2
