#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_pwcnet.py tests/test_gpu_model.py -x -q 2>&1 | tail -12
for m in 0 2; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/r02_pwc_launches.csv python tools/pwc_target.py > gpurun_out/ncu_pwc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_pwc_launches.csv > gpurun_out/r02_pwc_launches.txt; cat gpurun_out/r02_pwc_launches.txt
