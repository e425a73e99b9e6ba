#!/bin/bash
# One GPU call: bench line, per-layer table, ncu launch list, ncu --set full of the top conv kernel.
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.limit --format=csv
timeout 600 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench rc=$?"; tail -c 3000 gpurun_out/bench.json; tail -5 gpurun_out/bench.err
timeout 300 python tools/profile_layers.py 4 544 992 f16x3 > gpurun_out/layers_f16x3.txt 2>&1; head -3 gpurun_out/layers_f16x3.txt
timeout 300 python tools/profile_layers.py 8 192 192 f16x3 > gpurun_out/layers_cfg2_f16x3.txt 2>&1; head -2 gpurun_out/layers_cfg2_f16x3.txt
timeout 300 python tools/profile_layers.py 4 544 992 f16 > gpurun_out/layers_f16.txt 2>&1; head -2 gpurun_out/layers_f16.txt
if [ "$1" != "nonCU" ]; then
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches.csv python tools/ncu_target.py 4 544 992 > gpurun_out/ncu_launch.log 2>&1; echo "ncu launches rc=$?"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s 232 -c 4 -o gpurun_out/prof_conv -f python tools/ncu_target.py 4 544 992 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"
fi
ls -la gpurun_out
