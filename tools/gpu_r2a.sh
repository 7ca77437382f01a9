#!/bin/bash
# round 2, first GPU call (2 GPUs): cooperative-launch probe, distributed test, bench at N=2 (with the sharded leg), N=1, reference arm
OUT=gpurun_out; mkdir -p $OUT
timeout 120 tools/probes/coop_probe > $OUT/r2a_coop.log 2>&1; echo "coop rc=$?"; cat $OUT/r2a_coop.log
timeout 600 python -m pytest tests/test_distributed.py -x -q -m gpu > $OUT/r2a_dist.log 2>&1; echo "dist rc=$?"; tail -3 $OUT/r2a_dist.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29611 bench.py --gpus 2 --steps 5 --warmup 3 > $OUT/r2a_bench_n2.log 2>&1; echo "bench n2 rc=$?"; tail -c 3000 $OUT/r2a_bench_n2.log
timeout 900 python bench.py --steps 5 --warmup 3 > $OUT/r2a_bench_n1.log 2>&1; echo "bench n1 rc=$?"; tail -c 2500 $OUT/r2a_bench_n1.log
timeout 600 python bench.py --impl reference --steps 4 --warmup 1 > $OUT/r2a_bench_ref.log 2>&1; echo "bench ref rc=$?"; tail -c 1500 $OUT/r2a_bench_ref.log
