#!/bin/bash
# One gpurun call: GPU tests, bench at a few concurrency levels, ncu launch list and one `--set full` capture.
#   gpurun --timeout 1500 -- 'bash tools/gpu_profile.sh <tag> [full-kernel-regex]'
# Outputs land in gpurun_out/<tag>_*; summarise them into profiles/ with tools/ncu_summary.py / tools/launch_shares.py.
set -u
TAG=${1:-run}
KREGEX=${2:-'k_match|k_knn_warp|k_knn_tile|k_accumulate|k_normals_from_knn|k_select23|k_refit|k_radix_tree'}
OUT=gpurun_out
mkdir -p $OUT
python -m pytest tests -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1
echo "tests rc=$?" | tee -a $OUT/${TAG}_tests.log
tail -3 $OUT/${TAG}_tests.log
for S in 1 4 8; do
  python bench.py --steps 5 --warmup 3 --streams $S --no-cpu > $OUT/${TAG}_bench_s$S.log 2>&1
  echo "bench s=$S rc=$?"
done
python bench.py > $OUT/${TAG}_bench_default.log 2>&1
echo "bench default rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.log 2>&1
echo "bench reference rc=$?"
# profiled command: ONE registration (the timed step is bracketed by cudaProfilerStart/Stop); schedule 2 = the tile kernels
# that the batched path uses, schedule 1 = the per-thread kernels of the single-registration path
PCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule ${SCHED:-2} --knn-schedule ${SCHED:-2}"
$PCMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $PCMD > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
$PCMD > $OUT/${TAG}_plain2.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:$KREGEX" -c 40 -o $OUT/${TAG}_prof -f $PCMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full rc=$?"
python - <<EOF
import json
for s in ("s1", "s4", "s8", "default", "reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.log" % s).read().strip().splitlines()[-1])
        print(s, round(d["value"], 2), round(d.get("e2e", {}).get("value", 0), 2), d.get("stage_ms_per_registration"), d.get("roofline", {}).get("avg_launch_ms"))
    except Exception as e:
        print(s, "failed", e)
EOF
