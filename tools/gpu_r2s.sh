#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2s}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_goldens.py tests/test_gpu_configs.py tests/test_append.py -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
for F in 1 0; do
  AICP_B200_FUSED_TAIL=$F timeout 600 python bench.py --steps 6 --warmup 3 --no-cpu > $OUT/${TAG}_bench_f$F.log 2>&1; echo "bench fused=$F rc=$?"
  python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench_f$F.log").read().strip().splitlines()[-1])
    print("fused=$F value %.1f e2e %.1f lat %.3f ms" % (d["value"], d["e2e"]["value"], d["latency_single_stream"]["ms_per_registration"]))
    print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["stage_ms_per_registration"].items() if k != "note"}, "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench_f$F.log").read()[-3000:])
PY
done
