#!/bin/bash
# round 2: persistent loop kernel -- parity + bench A/B against the multi-launch loop
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2b}
timeout 1200 python -m pytest tests/test_gpu_parity.py tests/test_goldens.py -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -5 $OUT/${TAG}_tests.log
for LS in 0 1; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --loop-schedule $LS > $OUT/${TAG}_bench_ls$LS.log 2>&1; echo "bench ls=$LS rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_ls$LS.log").read().strip().splitlines()[-1])
    print("LS=$LS value %.1f e2e %.1f lat %.3f ms" % (d["value"], d["e2e"]["value"], d["latency_single_stream"]["ms_per_registration"]))
    print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["stage_ms_per_registration"].items() if k != "note"}, "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_ls$LS.log").read()[-3000:])
PY
done
