#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2o}
for V in base crop23; do
  LIBV=""; [ $V != base ] && LIBV="AICP_B200_LIB=$PWD/aicp_mapping_b200/lib/libaicp_b200_$V.so"
  env $LIBV python tools/crop_probe.py > $OUT/${TAG}_crop_$V.json 2>&1; echo $V; cut -c1-330 $OUT/${TAG}_crop_$V.json
done
timeout 1200 python tools/c4_scene_probe.py > $OUT/${TAG}_c4scene.log 2>&1; cat $OUT/${TAG}_c4scene.log
timeout 900 python -m pytest tests/test_gpu_configs.py -x -q -m gpu -k "c4" > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
