#!/bin/bash
# Calibration of the planner's rate tables (both kernels) on the GPU box; then tools/make_rates.py here.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_gpu_query_pairs.py -m gpu -q -x > gpurun_out/pytest_q2.log 2>&1; echo "pytest q2 exit $?"; tail -n 3 gpurun_out/pytest_q2.log
timeout 900 python tools/calibrate_q2.py 0.35 > gpurun_out/calib_q2.json 2> gpurun_out/calib_q2.err; echo "calib q2 exit $?"
timeout 900 python tools/calibrate.py 0.35 > gpurun_out/calib.json 2> gpurun_out/calib.err; echo "calib exit $?"
tail -n 2 gpurun_out/calib_q2.err gpurun_out/calib.err
