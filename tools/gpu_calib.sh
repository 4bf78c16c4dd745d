#!/bin/bash
# Calibration of the planner's rate tables (all three search kernels) on the GPU box; then, here:
#   python tools/make_rates.py gpurun_out/r2/calib.json gpurun_out/r2/calib_q2.json gpurun_out/r2/calib_xw.json
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 900 python tools/calibrate_q2.py 0.35 > $O/calib_q2.json 2> $O/calib_q2.err; echo "calib q2 exit $?"
timeout 900 python tools/calibrate.py 0.35 > $O/calib.json 2> $O/calib.err; echo "calib exit $?"
timeout 900 python tools/calibrate_xw.py 4000 4,8,16 > $O/calib_xw.json 2> $O/calib_xw.err; echo "calib xw exit $?"
tail -n 2 $O/calib_q2.err $O/calib.err $O/calib_xw.err
