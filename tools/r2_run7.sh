#!/bin/bash
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_gpu_window.py tests/test_gpu_video.py -q -m gpu -x 2>&1 | tail -3
python tools/warp_target.py 4 2>&1 | grep -v Warn
python tools/warp_target.py 0.5 2>&1 | grep -v Warn
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_yuv -s 3 -c 1 -o gpurun_out/prof_warp2 -f python tools/warp_target.py 4 > gpurun_out/ncu_warp2.log 2>&1; tail -1 gpurun_out/ncu_warp2.log
ncu -i gpurun_out/prof_warp2.ncu-rep --page details > gpurun_out/r2_ncu_warp2_details.txt 2>/dev/null
