#!/bin/bash
# multi-GPU check: sharded result == single-GPU result (NCCL gather and peer-memory exchange), then the bench at N ranks in
# both exchange modes
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 tools/dist_check.py 2>&1 | grep -v Warning | tail -5
for mode in p2p nccl; do
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29532 bench.py --gpus $N --steps 10 --warmup 3 --exchange $mode > gpurun_out/r2_bench_n${N}_$mode.json 2> gpurun_out/r2_bench_n${N}_$mode.err; echo "bench $mode rc=$?"
python - <<PY
import json
d=json.load(open("gpurun_out/r2_bench_n${N}_$mode.json"))
print("$mode: value", round(d["value"],2), "ms/step", round(d["ms_per_step"],3), "e2e", round(d["e2e"]["value"],2), "|", d["config"]["sharding"][:60], "| strong:", {k: d["strong_scaling"][k] for k in ("windows_per_step","tiles_per_rank_per_step","ms_per_step","value")})
PY
tail -2 gpurun_out/r2_bench_n${N}_$mode.err
done
