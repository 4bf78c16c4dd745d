#!/bin/bash
# round 2, visit B: long-sequence kernel with shared-memory rings: parity, calibration, cfg4 probe, whole GPU suite
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python -m pytest tests/test_gpu_long_sequences.py -m gpu -q -x --timeout 300 > $O/b_pytest_long.log 2>&1; echo "pytest long exit $?"
tail -n 5 $O/b_pytest_long.log
timeout 900 python tools/calibrate_xw.py 4000 > $O/b_calib_xw.json 2> $O/b_calib_xw.err; echo "calib exit $?"
cat $O/b_calib_xw.json
timeout 600 python tools/cfg4_probe.py > $O/b_cfg4.txt 2>&1; echo "cfg4 exit $?"
grep -v "^\[swg\]" $O/b_cfg4.txt
timeout 1500 python -m pytest tests -m gpu -q -x --timeout 600 > $O/b_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 8 $O/b_pytest_all.log
