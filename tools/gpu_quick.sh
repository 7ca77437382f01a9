#!/bin/bash
# quick GPU check: parity tests + bench at a few stream counts.  gpurun --timeout 900 -- 'bash tools/gpu_quick.sh <tag> [streams...]'
TAG=${1:-q}; shift
STREAMS=${@:-1 8}
OUT=gpurun_out; mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
for S in $STREAMS; do
  python bench.py --steps 5 --warmup 3 --streams $S --no-cpu > $OUT/${TAG}_bench_s$S.log 2>&1; echo "bench s=$S rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_s$S.log").read().strip().splitlines()[-1])
    print("S=$S value %.1f e2e %.1f lat %.3f ms" % (d["value"], d["e2e"]["value"], d["latency_single_stream"]["ms_per_registration"]))
    print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["stage_ms_per_registration"].items() if k != "note"}, "launches", d["gpu_launches"], "k_match ms", round(d["roofline"]["avg_launch_ms"], 4))
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_s$S.log").read()[-2000:])
PY
done
