#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_f16f8.py tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_parity_headline.py -q -m gpu -x 2>&1 | tail -2
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_dbl.txt 2>&1; head -1 gpurun_out/r2_layers_dbl.txt
python tools/window_time.py f16f8 "" 2>&1 | grep -v Warn
