#!/bin/bash
# The full-size cross-checks without cfg3 and without the test suite (see visit_crosscheck.sh)
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/full_crosscheck.py 1.0 cfg2 > $O/x_cross_cfg2.txt 2>&1; echo "cfg2 exit $?"; tail -n 6 $O/x_cross_cfg2.txt
SWG_TITIN=1 SWG_VERBOSE=1 timeout 600 python tools/full_crosscheck.py 1.0 cfg2 > $O/x_cross_cfg2_titin.txt 2>&1; echo "cfg2+tail exit $?"; grep -v "^\[swg\]   G=" $O/x_cross_cfg2_titin.txt | tail -n 8
timeout 600 python tools/full_crosscheck.py 1.0 cfg4 > $O/x_cross_cfg4.txt 2>&1; echo "cfg4 exit $?"; tail -n 4 $O/x_cross_cfg4.txt
