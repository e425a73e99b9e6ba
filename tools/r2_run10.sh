#!/bin/bash
mkdir -p gpurun_out
FISR_ASTAGES=2 timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_astages2.txt 2>&1; head -1 gpurun_out/r2_layers_astages2.txt
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_astages1.txt 2>&1; head -1 gpurun_out/r2_layers_astages1.txt
python tools/window_time.py f16f8 "" "FISR_ASTAGES=2" 2>&1 | grep -v Warning
