#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2m}
timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/${TAG}_dist.log 2>&1; echo "dist rc=$?"; tail -3 $OUT/${TAG}_dist.log
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_goldens.py tests/test_gpu_configs.py -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
python tools/loop_probe.py 0 6 > $OUT/${TAG}_probe.log 2>&1; cat $OUT/${TAG}_probe.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 3 --warmup 3 > $OUT/${TAG}_bench_n2.log 2>&1; echo "bench n2 rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n2.log").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f lat %.3f" % (d["value"], d["e2e"]["value"], d["latency_single_stream"]["ms_per_registration"]))
    for k in ("c3", "c4"):
        c = d["sharded"][k]
        print(k, "sharded %.3f single %.3f speedup %.3f identical %s" % (c["ms_per_registration"], c["ms_single_gpu"], c["speedup"], c["bit_identical_to_single_gpu"]), c["stage_ms_rank0_last_rep"]["sharded"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n2.log").read()[-3000:])
PY
