#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2z}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_goldens.py tests/test_gpu_configs.py -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
timeout 600 python bench.py --steps 8 --warmup 3 --no-cpu > $OUT/${TAG}_bench.log 2>&1; echo "bench rc=$?"
python - <<PY
import json
try:
    d = json.loads(open("$OUT/${TAG}_bench.log").read().strip().splitlines()[-1])
    print("value %.1f e2e %.1f lat %.3f ms" % (d["value"], d["e2e"]["value"], d["latency_single_stream"]["ms_per_registration"]))
    print("   ", {k: (round(v, 4) if isinstance(v, float) else v) for k, v in d["stage_ms_per_registration"].items() if k != "note"}, "launches", d["gpu_launches"])
except Exception as e:
    print("failed", e); print(open("$OUT/${TAG}_bench.log").read()[-3000:])
PY
python tools/append_probe.py > $OUT/${TAG}_append_probe.json 2>&1; tail -1 $OUT/${TAG}_append_probe.json | cut -c1-700
python tools/loop_probe.py 0 8 > $OUT/${TAG}_loop_probe.log 2>&1; cat $OUT/${TAG}_loop_probe.log
