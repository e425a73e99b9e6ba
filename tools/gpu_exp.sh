python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for p in f16f8 f16x3; do python tools/profile_layers.py 4 544 992 $p > gpurun_out/x_$p.txt 2>&1; head -1 gpurun_out/x_$p.txt; done
grep "level_3/FI-SR/conv/1\|level_3/FI-SR/conv/2\|level_3/FI-SR/conv/0\|level_3/dec/level_2/res_block/0/conv/0\|level_3/FI-SR/res_block/0/conv/0\|level_3/enc/level_0/res_block/0/conv/1" gpurun_out/x_f16f8.txt
echo KB1_NT=64; FISR_KB1_NT=64 python tools/profile_layers.py 4 544 992 f16f8 2>&1 | grep "^# plan\|level_3/FI-SR/conv/1\|level_2/FI-SR/conv/1"
python tools/profile_train.py 2>&1 | sed -n 2,5p
