#!/bin/bash
# round 2, visit D: long-tile planner probe + the GPU tests that changed
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 900 python tools/long_probe.py > $O/d_long_probe.txt 2>&1; echo "probe exit $?"
cat $O/d_long_probe.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/d_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 12 $O/d_pytest_all.log
timeout 300 python __graft_entry__.py smoke > $O/d_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 $O/d_smoke.log
