#!/bin/bash
for m in 2 3; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
for m in 2 3; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
timeout 1500 python -m pytest tests/test_gpu_pwcnet.py tests/test_gpu_video.py tests/test_gpu_test_phase.py -x -q 2>&1 | tail -5
