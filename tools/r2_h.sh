#!/bin/bash
# round 2, visit H: column chunks -- parity, probe
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 1200 python -m pytest tests/test_gpu_long_sequences.py -m gpu -q --timeout 900 -x > $O/h_pytest.log 2>&1; echo "pytest exit $?"
tail -n 6 $O/h_pytest.log
timeout 900 python tools/long_probe.py > $O/h_long_probe.txt 2>&1; echo "probe exit $?"
grep -v "^\[swg\]   G=" $O/h_long_probe.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x --deselect tests/test_gpu_long_sequences.py > $O/h_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 6 $O/h_pytest_all.log
