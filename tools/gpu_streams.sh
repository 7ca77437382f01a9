#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2y}
for S in 6 8 10 12 16; do
  timeout 600 python bench.py --steps 5 --warmup 3 --no-cpu --streams $S > $OUT/${TAG}_bench_s$S.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_s$S.log").read().strip().splitlines()[-1])
    print("S=$S value %.1f e2e %.1f" % (d["value"], d["e2e"]["value"]))
except Exception as e:
    print("failed", e)
PY
done
