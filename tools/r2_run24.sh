#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-200 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --precision f16x3 --no-extras > gpurun_out/r02_bench_n1_f16x3.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_n1_f16x3.json
timeout 600 python tools/profile_train.py > gpurun_out/r02_train_cfg3_B16_192.txt 2>&1; head -3 gpurun_out/r02_train_cfg3_B16_192.txt
