#!/bin/bash
mkdir -p gpurun_out
# level-2 dense convs + fused flow conv + dc_conv1 of the third forward (2 warm-up forwards of 228 tcgen05 launches each)
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s $((2 * 228 + 160)) -c 7 -o gpurun_out/prof_pwc_level2 -f python tools/pwc_target.py > gpurun_out/ncu_pwc_level2.log 2>&1
echo "ncu rc=$?"; tail -2 gpurun_out/ncu_pwc_level2.log
timeout 1500 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_target.py > gpurun_out/r02_sanitize.log 2>&1; echo "sanitize rc=$?"; tail -8 gpurun_out/r02_sanitize.log
