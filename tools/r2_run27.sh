#!/bin/bash
timeout 1500 python -m pytest tests/test_gpu_video.py tests/test_gpu_test_phase.py tests/test_gpu_pwcnet.py -x -q 2>&1 | tail -5
