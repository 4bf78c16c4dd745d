#!/bin/bash
# Full-size cross-checks on the GPU box: all scores identical between the plainest path (sequence-pair kernel, no
# long-tile split, no column chunks) and the planner's mix -- cfg2, cfg2 with a Swiss-Prot-like tail of very long
# sequences, cfg4, cfg3 -- then the GPU test suite.
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/full_crosscheck.py 1.0 cfg2 > $O/x_cross_cfg2.txt 2>&1; echo "cfg2 exit $?"; tail -n 6 $O/x_cross_cfg2.txt
SWG_TITIN=1 SWG_VERBOSE=1 timeout 600 python tools/full_crosscheck.py 1.0 cfg2 > $O/x_cross_cfg2_titin.txt 2>&1; echo "cfg2+tail exit $?"; grep -v "^\[swg\]   G=" $O/x_cross_cfg2_titin.txt | tail -n 16
timeout 600 python tools/full_crosscheck.py 1.0 cfg4 > $O/x_cross_cfg4.txt 2>&1; echo "cfg4 exit $?"; tail -n 4 $O/x_cross_cfg4.txt
timeout 900 python tools/full_crosscheck.py 1.0 cfg3 > $O/x_cross_cfg3.txt 2>&1; echo "cfg3 exit $?"; tail -n 4 $O/x_cross_cfg3.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/x_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 5 $O/x_pytest_all.log
