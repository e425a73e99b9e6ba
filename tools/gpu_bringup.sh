#!/bin/bash
mkdir -p gpurun_out
timeout 900 python tools/gpu_bringup.py 2>&1 | tee gpurun_out/bringup.log
