#!/bin/bash
# round 2, GPU run 1: full GPU suite (new parity tests), gradient-noise diagnosis, bench with the `extra` block, reference arm
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
nproc
timeout 1500 python -m pytest tests -q -m gpu --durations=20 -s > gpurun_out/r2_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -5 gpurun_out/r2_pytest_gpu.log
timeout 300 python tools/accum_probe.py > gpurun_out/r2_accum_probe.txt 2>&1; tail -6 gpurun_out/r2_accum_probe.txt
timeout 300 python tools/grad_check.py 2 64 64 71 > gpurun_out/r2_grad_default.txt 2>&1; tail -1 gpurun_out/r2_grad_default.txt
FISR_WGRAD_EXACT=1 timeout 300 python tools/grad_check.py 2 64 64 71 > gpurun_out/r2_grad_exact.txt 2>&1; tail -1 gpurun_out/r2_grad_exact.txt
timeout 600 python bench.py > gpurun_out/r2_bench.json 2> gpurun_out/r2_bench.err; echo "bench rc=$?"; cut -c1-400 gpurun_out/r2_bench.json; tail -3 gpurun_out/r2_bench.err
timeout 300 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/r2_bench_ref.json 2>/dev/null; cut -c1-300 gpurun_out/r2_bench_ref.json
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_tile_f16f8.txt 2>&1; head -1 gpurun_out/r2_layers_tile_f16f8.txt
