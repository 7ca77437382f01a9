#!/bin/bash
# full GPU suite + default bench + reference arm + ncu launch list and full capture of the loop kernel (one registration)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2q}
timeout 2400 python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
timeout 900 python bench.py > $OUT/${TAG}_bench_default.log 2>&1; echo "bench rc=$?"
timeout 600 python bench.py --impl reference --steps 5 --warmup 1 > $OUT/${TAG}_bench_reference.log 2>&1; echo "ref rc=$?"
python tools/append_probe.py > $OUT/${TAG}_append_probe.json 2>&1; tail -1 $OUT/${TAG}_append_probe.json | cut -c1-700
python tools/loop_probe.py 0 6 > $OUT/${TAG}_loop_probe.log 2>&1; cat $OUT/${TAG}_loop_probe.log
PCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule 1 --knn-schedule 1"
$PCMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/${TAG}_launches.csv $PCMD > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$PCMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_icp_loop|k_knn_warp" -c 2 -o $OUT/${TAG}_prof -f $PCMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
python - <<PY
import json
for s in ("default", "reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.log" % s).read().strip().splitlines()[-1])
        print(s, "value %.2f e2e %.2f" % (d["value"], d["e2e"]["value"]), d.get("latency_single_stream"), d.get("cpu_baseline", {}).get("value"))
    except Exception as e:
        print(s, "failed", e)
PY
