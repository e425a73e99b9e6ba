#!/bin/bash
# quick loop: GPU tests + per-layer tables (no ncu)
mkdir -p gpurun_out
timeout 900 python -m pytest tests -x -q -m gpu 2>&1 | tail -4
for ms in 3 2; do
  FISR_MIN_SLOTS=$ms timeout 300 python tools/profile_layers.py 4 544 992 f16x3 > gpurun_out/layers_f16x3_ms$ms.txt 2>&1; head -1 gpurun_out/layers_f16x3_ms$ms.txt
done
timeout 300 python tools/profile_layers.py 8 192 192 f16x3 > gpurun_out/layers_cfg2_f16x3.txt 2>&1; head -1 gpurun_out/layers_cfg2_f16x3.txt
timeout 300 python tools/profile_layers.py 4 544 992 f16 > gpurun_out/layers_f16.txt 2>&1; head -1 gpurun_out/layers_f16.txt
