#!/bin/bash
# 2-GPU check of the communicator: world-2 tests (NCCL + peer-mailbox scan) and a short torchrun bench.
nvidia-smi topo -m > gpurun_out/r2_topo_n2.txt
timeout 300 python -m pytest tests/test_gpu_comm.py -m gpu -x -q 2>&1 | tail -30 > gpurun_out/r2_pytest_comm_n2.log
cat gpurun_out/r2_pytest_comm_n2.log
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 \
  bench.py --gpus 2 --steps 5 --warmup 3 > gpurun_out/r2_bench_n2.log 2> gpurun_out/r2_bench_n2.err
echo bench_rc=$?
tail -c 1500 gpurun_out/r2_bench_n2.err
python - <<'PY'
import json
l = json.loads(open("gpurun_out/r2_bench_n2.log").read().strip().splitlines()[-1])
print(l["value"], l["e2e"]["value"], l["e2e"]["roofline"])
print(l["scan"])
print(l["loglik"])
PY
