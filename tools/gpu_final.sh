#!/bin/bash
# Round-end evidence run: whole GPU suite, smoke, bench lines, ncu launch lists and one --set full capture per path.
#   gpurun --timeout 1500 -- 'bash tools/gpu_final.sh <tag>'
set -u
TAG=${1:-final}
OUT=gpurun_out
mkdir -p $OUT
(time python -m pytest tests -q -m gpu --timeout 300) > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
python -c "import __graft_entry__ as g; g.smoke()" > $OUT/${TAG}_smoke.log 2>&1; echo "smoke rc=$?"; tail -2 $OUT/${TAG}_smoke.log
python bench.py > $OUT/${TAG}_bench_default.log 2>&1; echo "bench default rc=$?"
python bench.py --impl reference --steps 2 --warmup 1 > $OUT/${TAG}_bench_reference.log 2>&1; echo "bench reference rc=$?"
python tools/bench_configs.py --configs prefilter,risk,5risk > $OUT/${TAG}_bench_filters.log 2>&1; echo "bench filters rc=$?"
PCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule 2 --knn-schedule 2"
$PCMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --metrics gpu__time_duration.sum --clock-control none --profile-from-start off -c 1500 --csv --log-file $OUT/${TAG}_launches.csv $PCMD > $OUT/${TAG}_ncu1.log 2>&1
echo "ncu launches rc=$?"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_match_tile|k_knn_tile|k_accumulate|k_select23" -c 8 -o $OUT/${TAG}_prof -f $PCMD > $OUT/${TAG}_ncu2.log 2>&1
echo "ncu full (registration) rc=$?"
FCMD="python tools/prefilter_profile.py 3 1"
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_knn_warp|k_pf_cc_union|k_vg_centroids|k_pf_edges|k_pf_normals|k_radix_pass" -c 9 -o $OUT/${TAG}_prof_prefilter -f $FCMD > $OUT/${TAG}_ncu3.log 2>&1
echo "ncu full (prefilter) rc=$?"
python - <<PY
import json
for s in ("default", "reference"):
    try:
        d = json.loads(open("$OUT/${TAG}_bench_%s.log" % s).read().strip().splitlines()[-1])
        print(s, round(d["value"], 2), d.get("e2e", {}).get("value"), d.get("roofline", {}).get("frac"))
    except Exception as e:
        print(s, "failed", e)
PY
