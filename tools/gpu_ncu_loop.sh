#!/bin/bash
# ncu --set full of the persistent loop kernel on ONE C3 registration (per-thread search schedule)
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2h}
PCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule 1 --knn-schedule 1"
$PCMD > $OUT/${TAG}_plain.log 2>&1 &&
ncu --set full --clock-control none --import-source on --profile-from-start off -k "regex:k_icp_loop|k_knn_warp" -c 4 -o $OUT/${TAG}_prof -f $PCMD > $OUT/${TAG}_ncu.log 2>&1
echo "ncu rc=$?"; tail -5 $OUT/${TAG}_ncu.log
