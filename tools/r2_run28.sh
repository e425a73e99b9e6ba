#!/bin/bash
mkdir -p gpurun_out
timeout 1500 python -m pytest tests/test_gpu_pwcnet.py -x -q 2>&1 | tail -8
for m in 2 3; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_pwc_launches.csv python tools/pwc_target.py > gpurun_out/ncu_pwc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_pwc_launches.csv > gpurun_out/r02_pwc_launches.txt; cat gpurun_out/r02_pwc_launches.txt
