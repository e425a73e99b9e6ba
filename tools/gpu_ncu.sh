#!/bin/bash
# ncu --set full (+source counters) of selected conv launches of the SECOND forward of tools/ncu_target.py
#   $1 = precision, then pairs "<name> <conv index in launch order> <count>" (see tools/ncu_target.py for the order)
mkdir -p gpurun_out
prec=${1:-f16f8}; shift
while [ $# -ge 3 ]; do
  name=$1; idx=$2; cnt=$3; shift 3
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:conv3x3_umma -s $((138 + idx)) -c $cnt \
      -o gpurun_out/prof_${prec}_${name} -f python tools/ncu_target.py 4 544 992 $prec 2 > gpurun_out/ncu_${prec}_${name}.log 2>&1
  echo "ncu full $prec $name rc=$?"; tail -2 gpurun_out/ncu_${prec}_${name}.log
done
