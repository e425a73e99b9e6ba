#!/bin/bash
# round-2 measurement set: GPU tests, smoke, bench lines (both precisions, reference arm), per-layer tables, training table,
# ncu launch list, DRAM traffic of every conv launch of one forward, full captures of representative kernels
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,power.limit --format=csv,noheader
timeout 1500 python -m pytest tests -x -q -m gpu > gpurun_out/r02_pytest_gpu.log 2>&1; echo "pytest rc=$?"; tail -2 gpurun_out/r02_pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" 2>&1 | grep smoke
timeout 900 python bench.py > gpurun_out/r02_bench_n1.json 2> gpurun_out/r02_bench_n1.err; echo "bench rc=$?"; cut -c1-160 gpurun_out/r02_bench_n1.json
timeout 600 python bench.py --precision f16x3 --no-extras > gpurun_out/r02_bench_n1_f16x3.json 2>/dev/null
timeout 600 python bench.py --impl reference --steps 3 --warmup 1 > gpurun_out/r02_bench_reference_arm.json 2>/dev/null; cut -c1-200 gpurun_out/r02_bench_reference_arm.json
for p in f16f8 f16x3; do
  timeout 300 python tools/profile_layers.py 4 544 992 $p > gpurun_out/r02_layers_4x544x992_$p.txt 2>&1; head -1 gpurun_out/r02_layers_4x544x992_$p.txt
  timeout 300 python tools/profile_layers.py 8 192 192 $p > gpurun_out/r02_layers_8x192x192_$p.txt 2>&1; head -1 gpurun_out/r02_layers_8x192x192_$p.txt
done
timeout 600 python tools/profile_train.py > gpurun_out/r02_train_cfg3_B16_192.txt 2>&1; head -5 gpurun_out/r02_train_cfg3_B16_192.txt
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 900 --csv --log-file gpurun_out/r02_launches_f16f8.csv python bench.py --steps 2 --warmup 1 --no-extras > gpurun_out/ncu_bench.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_launches_f16f8.csv > gpurun_out/r02_launches_f16f8.txt; head -8 gpurun_out/r02_launches_f16f8.txt
timeout 900 ncu --metrics dram__bytes_read.sum,dram__bytes_write.sum,gpu__time_duration.sum --clock-control none -k regex:conv3x3_umma -s 138 -c 138 --csv --log-file gpurun_out/r02_traffic_f16f8.csv python tools/ncu_target.py 4 544 992 f16f8 2 > gpurun_out/ncu_traffic.log 2>&1
tools/gpu_ncu.sh f16f8 conv64 93 2 pool64 96 1 conv128 98 2 head 131 2
timeout 600 ncu --set full --clock-control none --import-source on -k regex:warp_yuv -s 3 -c 1 -o gpurun_out/prof_warp -f python tools/warp_target.py 4 > gpurun_out/ncu_warp.log 2>&1
python tools/warp_target.py 4 2>&1 | grep -v Warn
# PWC-Net: launch list of 5 forwards on both directions of a 1080p pair, full capture of the level-2 estimator convs
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 2500 --csv --log-file gpurun_out/r02_pwc_launches.csv python tools/pwc_target.py > gpurun_out/ncu_pwc.log 2>&1
python tools/summarize_launches.py gpurun_out/r02_pwc_launches.csv > gpurun_out/r02_pwc_launches.txt; head -6 gpurun_out/r02_pwc_launches.txt
for m in 0 1 2 3; do FISR_PWC_UMMA=$m timeout 300 python tools/pwc_target.py 2>&1 | tail -1; done
# level-2 estimator convs + fused flow conv + dc_conv1 of the fourth forward (3 warm-up forwards of 248 tcgen05 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s $((3 * 248 + 180)) -c 7 -o gpurun_out/prof_pwc_level2 -f python tools/pwc_target.py > gpurun_out/ncu_pwc_level2.log 2>&1; tail -1 gpurun_out/ncu_pwc_level2.log
