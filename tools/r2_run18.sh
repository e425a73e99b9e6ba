#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -s 111 -c 111 --csv --log-file gpurun_out/r02_pwc_launches.csv python tools/pwc_target.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/r02_pwc_launches.csv
