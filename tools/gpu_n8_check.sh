#!/bin/bash
# 8-GPU validation of the bench (what the driver's SCALE run does at N = 8), with its own timeout.
N=${1:-8}
nvidia-smi topo -m > gpurun_out/r2_topo_n$N.txt
timeout 700 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  bench.py --gpus $N --steps 5 --warmup 3 > gpurun_out/r2_bench_n$N.log 2> gpurun_out/r2_bench_n$N.err
echo bench_rc=$?
tail -c 800 gpurun_out/r2_bench_n$N.err
python - <<PY
import json
l = json.loads(open("gpurun_out/r2_bench_n$N.log").read().strip().splitlines()[-1])
print("value", l["value"], "frac", l["roofline"]["frac"])
print("e2e", l["e2e"]["value"], l["e2e"]["roofline"])
print("scan", l["scan"])
print("loglik", l["loglik"])
PY
