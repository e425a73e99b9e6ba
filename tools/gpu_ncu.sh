#!/bin/bash
# ncu --set full (+source counters) of selected conv launches of the second forward: $1 = skip count, $2 = count
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s ${1:-231} -c ${2:-2} -o gpurun_out/prof_conv -f python tools/ncu_target.py 4 544 992 > gpurun_out/ncu_full.log 2>&1; echo "ncu full rc=$?"; tail -3 gpurun_out/ncu_full.log
