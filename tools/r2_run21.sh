#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_pwc_launches.csv python tools/pwc_target.py > gpurun_out/ncu_pwc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_pwc_launches.csv > gpurun_out/r02_pwc_launches.txt; cat gpurun_out/r02_pwc_launches.txt
python tools/warp_target.py 4 2>&1 | grep -v Warn
