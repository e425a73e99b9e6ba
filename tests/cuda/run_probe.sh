#!/bin/bash
# Runs every probe mode/variant in its own process under a timeout (a hang must not take the box down).
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,memory.total --format=csv
for mode in 0 1 2; do for variant in 0 1 2; do
  timeout 60 ./build/umma_probe $mode $variant; echo "exit=$?"
done; done 2>&1 | tee gpurun_out/umma_probe.log
