#!/bin/bash
timeout 900 python -m pytest tests/test_gpu_video.py -q -m gpu -x 2>&1 | tail -12
