#!/bin/bash
# round 2, GPU run 2: PDL + fused max-pool + merged FI-SR/SR conv/0 -- parity subset, A/B timing, per-layer table
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_f16f8.py tests/test_gpu_window.py tests/test_gpu_parity_headline.py tests/test_gpu_train_loop.py tests/test_gpu_backward.py -q -m gpu -x 2>&1 | tail -5
(
FISR_NO_PDL=1 FISR_NO_POOL_FUSION=1 FISR_NO_HEAD_MERGE=1 python tools/window_time.py f16f8
FISR_NO_POOL_FUSION=1 FISR_NO_HEAD_MERGE=1 python tools/window_time.py f16f8
FISR_NO_PDL=1 FISR_NO_HEAD_MERGE=1 python tools/window_time.py f16f8
FISR_NO_PDL=1 FISR_NO_POOL_FUSION=1 python tools/window_time.py f16f8
python tools/window_time.py f16f8
FISR_NO_PDL=1 FISR_NO_POOL_FUSION=1 python tools/window_time.py f16x3
python tools/window_time.py f16x3
) 2>&1 | grep -v Warning | tee gpurun_out/r2_ab_run2.txt
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/r2_layers_tile_f16f8_b.txt 2>&1; head -1 gpurun_out/r2_layers_tile_f16f8_b.txt
