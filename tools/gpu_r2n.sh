#!/bin/bash
# TMA crop: parity tests, probe legacy vs bulk, ncu time of the kernels; multi-device batch test; adapter demo
OUT=gpurun_out; mkdir -p $OUT; TAG=${1:-r2n}
timeout 900 python -m pytest tests/test_crop_box.py tests/test_ingest.py tests/test_host.py "tests/test_gpu_configs.py::test_register_batch_over_all_devices" -x -q -m gpu > $OUT/${TAG}_tests.log 2>&1; echo "tests rc=$?"; tail -4 $OUT/${TAG}_tests.log
AICP_B200_CROP=legacy python tools/crop_probe.py > $OUT/${TAG}_crop_legacy.json 2>&1; cat $OUT/${TAG}_crop_legacy.json
python tools/crop_probe.py > $OUT/${TAG}_crop_bulk.json 2>&1; cat $OUT/${TAG}_crop_bulk.json
python tools/crop_probe.py > /dev/null 2>&1 &&
ncu --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum --clock-control none -k regex:k_crop -c 6 --csv --log-file $OUT/${TAG}_crop_ncu.csv python tools/crop_probe.py > $OUT/${TAG}_crop_ncu.log 2>&1; echo "ncu rc=$?"; grep -E "k_crop" $OUT/${TAG}_crop_ncu.csv | tail -6 | cut -c1-300
