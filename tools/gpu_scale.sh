#!/bin/bash
# bench.py at N GPUs exactly as the driver launches it (the line carries the sharded leg); optional: distributed test first
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-scale}; N=${2:-2}; DIST=${3:-0}
if [ "$DIST" = "1" ]; then
  timeout 900 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/${TAG}_dist_n$N.log 2>&1; echo "dist rc=$?"; tail -3 $OUT/${TAG}_dist_n$N.log
fi
timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29621 bench.py --gpus $N --steps 10 --warmup 3 > $OUT/${TAG}_bench_n$N.log 2>&1; echo "bench n$N rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_n$N.log").read().strip().splitlines()[-1])
    json.dump(d, open("$OUT/${TAG}_bench_n$N.json", "w"))
    print("N=$N value %.1f e2e %.1f ms/step %.2f" % (d["value"], d["e2e"]["value"], d["ms_per_step"]))
    sh = d.get("sharded", {})
    print(sh.get("exchange"))
    for k in ("c3", "c4"):
        c = sh.get(k)
        if c: print(k, "sharded %.3f single %.3f speedup %.3f identical %s it %d" % (c["ms_per_registration"], c["ms_single_gpu"], c["speedup"], c["bit_identical_to_single_gpu"], c["iterations"]), c["stage_ms_rank0_last_rep"]["sharded"])
    if "error" in sh: print("ERROR", sh["error"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_n$N.log").read()[-3000:])
PY
if [ "${4:-0}" = "1" ]; then
  timeout 900 python tools/bench_devices.py --pairs 64 --steps 5 > $OUT/${TAG}_devices_n$N.json 2>&1; echo "bench_devices rc=$?"; tail -1 $OUT/${TAG}_devices_n$N.json | cut -c1-400
fi
