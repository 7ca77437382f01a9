#!/bin/bash
# warp instructions per launch of one C3 registration, both schedules (for the issue-slot line of bench.py)
set -u
OUT=gpurun_out; mkdir -p $OUT
ACMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run --match-schedule 2 --knn-schedule 2 --loop-schedule 1"
$ACMD > $OUT/inst_plain_a.log 2>&1 &&
ncu --metrics smsp__inst_executed.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/inst_batch_schedule.csv $ACMD > $OUT/inst_ncu_a.log 2>&1
echo "ncu inst (a) rc=$?"
BCMD="python bench.py --pairs 1 --streams 1 --steps 1 --warmup 2 --no-cpu --profile-run"
$BCMD > $OUT/inst_plain_b.log 2>&1 &&
ncu --metrics smsp__inst_executed.sum --clock-control none --profile-from-start off -c 400 --csv --log-file $OUT/inst_single.csv $BCMD > $OUT/inst_ncu_b.log 2>&1
echo "ncu inst (b) rc=$?"
