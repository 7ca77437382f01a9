#!/bin/bash
# Round-2 evidence run on ONE B200: whole GPU suite, smoke, bench lines, the secondary configurations, probes, ncu launch lists and
# --set full captures of (a) the schedule the timed batch runs and (b) one registration at a time (persistent loop kernel).
#   gpurun --timeout 3000 -- 'bash tools/gpu_final2.sh <tag>'
set -u
TAG=${1:-final2}
OUT=gpurun_out
mkdir -p $OUT
(time python -m pytest tests -q -m gpu --timeout 300) > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -3 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench_default.log 2>&1; echo "bench default rc=$?"
python bench.py --impl reference --steps 20 --warmup 3 > $OUT/${TAG}_bench_reference.log 2>&1; echo "bench reference rc=$?"
python tools/bench_configs.py > $OUT/${TAG}_bench_configs.log 2>&1; echo "bench configs rc=$?"
python tools/bench_pipeline.py > $OUT/${TAG}_bench_pipeline.log 2>&1; echo "bench pipeline rc=$?"
python tools/append_probe.py > $OUT/${TAG}_append_probe.json 2>&1; echo "append rc=$?"
python tools/crop_probe.py > $OUT/${TAG}_crop_probe.json 2>&1; echo "crop rc=$?"
python tools/loop_probe.py 0 8 > $OUT/${TAG}_loop_probe.log 2>&1; echo "loop probe rc=$?"
# (a) what a batch worker runs: tile kernels, two launches per iteration
ACMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule 2 --knn-schedule 2 --loop-schedule 1"
$ACMD > $OUT/${TAG}_plain_a.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/${TAG}_launches_batch_schedule.csv $ACMD > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches (a) rc=$?"
$ACMD > $OUT/${TAG}_plain_a2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_match_tile|k_knn_tile|k_quantile_accumulate" -c 8 -o $OUT/${TAG}_prof_batch_schedule -f $ACMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full (a) rc=$?"
# (b) one registration at a time: per-thread search inside the persistent loop kernel
BCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run"
$BCMD > $OUT/${TAG}_plain_b.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/${TAG}_launches_single.csv $BCMD > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu launches (b) rc=$?"
$BCMD > $OUT/${TAG}_plain_b2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_icp_loop|k_knn_warp" -c 2 -o $OUT/${TAG}_prof_single -f $BCMD > $OUT/${TAG}_ncu4.log 2>&1
echo "ncu full (b) rc=$?"
CCMD="python tools/crop_probe.py"
$CCMD > /dev/null 2>&1 &&
ncu --set full --clock-control none -k "regex:k_crop" -s 6 -c 4 -o $OUT/${TAG}_prof_crop -f $CCMD > $OUT/${TAG}_ncu5.log 2>&1
echo "ncu full (crop) rc=$?"
python - <<PY
import json
for s in ("default", "reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.log" % s).read().strip().splitlines()[-1])
        print(s, round(d["value"], 2), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"), d.get("latency_single_stream"))
    except Exception as e:
        print(s, "failed", e)
PY
