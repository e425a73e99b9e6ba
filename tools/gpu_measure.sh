#!/bin/bash
# round-end measurement set: GPU tests, bench lines (both precisions + reference arm), per-layer tables, training table,
# ncu launch list, DRAM traffic of every conv launch of one forward, full captures of four representative convs
mkdir -p gpurun_out
python -m pytest tests -m gpu -x -q 2>&1 | tail -2
python bench.py > gpurun_out/bench_f16f8.json 2> gpurun_out/bench_f16f8.err; cut -c1-200 gpurun_out/bench_f16f8.json
python bench.py --precision f16x3 > gpurun_out/bench_f16x3.json 2>/dev/null
python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2>/dev/null
for p in f16f8 f16x3; do python tools/profile_layers.py 4 544 992 $p > gpurun_out/layers_tile_$p.txt 2>&1; python tools/profile_layers.py 8 192 192 $p > gpurun_out/layers_cfg2_$p.txt 2>&1; done
python tools/profile_train.py > gpurun_out/train_cfg3.txt 2>&1
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 800 --csv --log-file gpurun_out/launches_f16f8.csv python bench.py --steps 2 --warmup 1 > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/launches_f16f8.csv > gpurun_out/launches_f16f8.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv3x3_umma -s 138 -c 138 --csv --log-file gpurun_out/traffic_f16f8.csv python tools/ncu_target.py 4 544 992 f16f8 2 > gpurun_out/ncu_traffic.log 2>&1
tools/gpu_ncu.sh f16f8 conv64 93 2 conv128 98 2 head 131 2
