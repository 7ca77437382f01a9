#!/bin/bash
# sharded registration over peer memory: distributed test (all GPUs of the box) + bench at N = all GPUs (sharded leg)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2e}; N=${2:-2}
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/${TAG}_dist.log 2>&1; echo "dist rc=$?"; tail -15 $OUT/${TAG}_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus $N --steps 3 --warmup 3 > $OUT/${TAG}_bench_n$N.log 2>&1; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n$N.log").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
    print(json.dumps(d.get("sharded"), indent=1))
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n$N.log").read()[-3000:])
PY
AICP_B200_COMM=nccl timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29612 bench.py --gpus $N --steps 1 --warmup 3 > $OUT/${TAG}_bench_n${N}_nccl.log 2>&1; echo "bench nccl rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n${N}_nccl.log").read().strip().splitlines()[-1])
    print("NCCL carrier:", json.dumps(d.get("sharded")))
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n${N}_nccl.log").read()[-3000:])
PY
