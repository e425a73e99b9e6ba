#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pwcnet.py -x -q 2>&1 | tail -3
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cost_volume -c 10 --csv --log-file gpurun_out/cv.csv python tools/pwc_target.py > /dev/null 2>&1
python tools/summarize_launches.py gpurun_out/cv.csv; grep cost_volume gpurun_out/cv.csv | awk -F'","' '{print $NF}' | head -5
timeout 300 python tools/pwc_target.py 2>&1 | tail -1
