#!/bin/bash
# PWC-Net on the tensor cores: parity in the three conv modes, FISRnet regression subset, timing
mkdir -p gpurun_out
for m in 0 1 2; do
  echo "== FISR_PWC_UMMA=$m"
  FISR_PWC_UMMA=$m timeout 600 python -m pytest tests/test_gpu_pwcnet.py -x -q 2>&1 | tail -12
done
timeout 900 python -m pytest tests/test_gpu_model.py tests/test_gpu_conv.py tests/test_gpu_parity_headline.py -x -q 2>&1 | tail -5
for m in 0 1 2; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -2; done
