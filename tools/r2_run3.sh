#!/bin/bash
# round 2, GPU run 3: full GPU suite (-x, as the driver runs it) + ncu full captures of four small deep layers
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r2_pytest_gpu_b.log 2>&1; echo "pytest rc=$?"; tail -3 gpurun_out/r2_pytest_gpu_b.log
tools/gpu_ncu.sh f16f8 l1bott 16 1 l1enc2 11 1 l2bott 62 1 l1enc1 6 1
for n in l1bott l1enc2 l2bott l1enc1; do ncu -i gpurun_out/prof_f16f8_$n.ncu-rep --page raw --csv > gpurun_out/r2_ncu_$n.csv 2>/dev/null; done
ls -la gpurun_out/*.ncu-rep | head
