#!/bin/bash
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2u}
timeout 1500 python -m pytest tests/test_gpu_parity.py tests/test_goldens.py tests/test_crop_box.py -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -3 $OUT/${TAG}_tests.log
python tools/loop_probe.py 0 8 > $OUT/${TAG}_probe_side.log 2>&1; grep "loop 0 match 1" $OUT/${TAG}_probe_side.log
AICP_B200_LIB=$PWD/aicp_mapping_b200/lib/libaicp_b200_noside.so python tools/loop_probe.py 0 8 > $OUT/${TAG}_probe_noside.log 2>&1; grep "loop 0 match 1" $OUT/${TAG}_probe_noside.log
python tools/loop_probe.py 5 8 > $OUT/${TAG}_probe_side5.log 2>&1; grep "loop 0 match 1" $OUT/${TAG}_probe_side5.log
AICP_B200_LIB=$PWD/aicp_mapping_b200/lib/libaicp_b200_noside.so python tools/loop_probe.py 5 8 > $OUT/${TAG}_probe_noside5.log 2>&1; grep "loop 0 match 1" $OUT/${TAG}_probe_noside5.log
