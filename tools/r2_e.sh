#!/bin/bash
# round 2, visit E: long-tile probe with the wide-shape alternative, changed GPU tests, then the ncu evidence
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 900 python tools/long_probe.py > $O/e_long_probe.txt 2>&1; echo "probe exit $?"
grep -v "^\[swg\]   G=" $O/e_long_probe.txt
timeout 1800 python -m pytest tests/test_gpu_long_sequences.py tests/test_gpu_parity.py tests/test_gpu_batches.py -m gpu -q --timeout 900 -x > $O/e_pytest.log 2>&1; echo "pytest exit $?"
tail -n 6 $O/e_pytest.log
timeout 2400 bash tools/r2_profile.sh > $O/e_profile.log 2>&1; echo "profile exit $?"
tail -n 16 $O/e_profile.log
