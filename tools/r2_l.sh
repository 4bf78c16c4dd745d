#!/bin/bash
# round 2, visit L: full-size CLI parity against the reference (cfg2, cfg4, cfg5 sweep), then the ncu evidence of the final build
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 900 python tools/full_parity.py 1.0 20000 1 cfg2 > $O/l_parity_cfg2.txt 2>&1; echo "cfg2 exit $?"; cat $O/l_parity_cfg2.txt
timeout 600 python tools/full_parity.py 1.0 100 1 cfg4 > $O/l_parity_cfg4.txt 2>&1; echo "cfg4 exit $?"; cat $O/l_parity_cfg4.txt
timeout 900 python tools/full_parity.py 1.0 100 1 cfg5 > $O/l_parity_cfg5.txt 2>&1; echo "cfg5 exit $?"; cat $O/l_parity_cfg5.txt
timeout 600 python -m pytest tests/test_gpu_cli_stream.py -m gpu -q -x > $O/l_pytest_cli.log 2>&1; echo "cli tests exit $?"; tail -n 3 $O/l_pytest_cli.log
timeout 2400 bash tools/r2_profile.sh > $O/l_profile.log 2>&1; echo "profile exit $?"
tail -n 14 $O/l_profile.log
