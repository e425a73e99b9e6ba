#!/bin/bash
mkdir -p gpurun_out
python tools/warp_target.py 4 2>&1 | grep -v Warn
python tools/warp_target.py 0.5 2>&1 | grep -v Warn
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_yuv -s 3 -c 1 -o gpurun_out/prof_warp -f python tools/warp_target.py 4 > gpurun_out/ncu_warp.log 2>&1; tail -2 gpurun_out/ncu_warp.log
ncu -i gpurun_out/prof_warp.ncu-rep --page details > gpurun_out/r2_ncu_warp_details.txt 2>/dev/null
ncu -i gpurun_out/prof_warp.ncu-rep --page source --csv > gpurun_out/r2_ncu_warp_source.csv 2>/dev/null
