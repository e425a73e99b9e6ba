python -m pytest tests/test_gpu_f16f8.py tests/test_gpu_model.py tests/test_gpu_window.py -x -q 2>&1 | tail -2
python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/l_base.txt 2>&1; head -1 gpurun_out/l_base.txt; grep "level_3/FI-SR/conv/1\|level_3/FI-SR/conv/2\|level_3/SR/conv/2\|level_3/FI-SR/conv/0\|level_3/dec/level_2/res_block/0/conv/0" gpurun_out/l_base.txt
for v in 1 2; do echo "ASTAGES=$v"; FISR_ASTAGES=$v python tools/profile_layers.py 4 544 992 f16f8 2>&1 | grep "^# plan\|level_3/FI-SR/conv/2\|level_3/SR/conv/2\|level_3/FI-SR/conv/0\|level_3/dec/level_2/res_block/0/conv/0"; done
python tools/profile_layers.py 8 192 192 f16f8 2>&1 | head -1
