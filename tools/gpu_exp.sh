python -m pytest tests -m gpu -x -q 2>&1 | tail -2
for p in f16f8 f16x3; do python tools/profile_layers.py 4 544 992 $p > gpurun_out/layers_tile_$p.txt 2>&1; head -1 gpurun_out/layers_tile_$p.txt; done
grep "level_3/FI-SR/conv/1\|level_3/FI-SR/conv/2\|level_3/SR/conv/2\|level_3/FI-SR/conv/0\|level_3/dec/level_2/res_block/0/conv/0\|level_3/FI-SR/res_block/0/conv/0\|level_3/dec/level_2/resize" gpurun_out/layers_tile_f16f8.txt
python tools/profile_layers.py 8 192 192 f16f8 2>&1 | head -1
python tools/profile_train.py 2>&1 | sed -n 2,5p
