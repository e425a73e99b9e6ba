timeout 300 python -m pytest tests/test_gpu_f16f8.py -x -q 2>&1 | tail -5
timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/x_pair.txt 2>&1; head -1 gpurun_out/x_pair.txt
FISR_PAIR=0 timeout 300 python tools/profile_layers.py 4 544 992 f16f8 > gpurun_out/x_nopair.txt 2>&1; head -1 gpurun_out/x_nopair.txt
paste <(grep "level_3" gpurun_out/x_nopair.txt | awk '{printf "%-52s %-8s %8s\n", substr($1,17), $2, $3}') <(grep "level_3" gpurun_out/x_pair.txt | awk '{printf "%8s\n", $3}') | head -50
