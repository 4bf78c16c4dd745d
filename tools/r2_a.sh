#!/bin/bash
# round 2, visit A: long-sequence kernel parity + cfg4 probe
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1; nproc >> $O/gpu.txt
timeout 900 python -m pytest tests/test_gpu_long_sequences.py -m gpu -q -x --timeout 300 > $O/a_pytest_long.log 2>&1; echo "pytest long exit $?"
tail -n 25 $O/a_pytest_long.log
timeout 600 python -m pytest tests/test_gpu_parity.py -m gpu -q -x --timeout 300 -k "long or golden" > $O/a_pytest_parity.log 2>&1; echo "pytest parity exit $?"
tail -n 5 $O/a_pytest_parity.log
timeout 600 python tools/cfg4_probe.py > $O/a_cfg4.txt 2>&1; echo "cfg4 exit $?"
cat $O/a_cfg4.txt
