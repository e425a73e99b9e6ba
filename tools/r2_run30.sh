#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_pwcnet.py tests/test_gpu_video.py -x -q 2>&1 | tail -5
timeout 300 python tools/pwc_target.py 2>&1 | tail -1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:cost_volume -c 10 python tools/pwc_target.py 2>&1 | grep -A3 "cost_volume" | grep "gpu__time" | head -10
