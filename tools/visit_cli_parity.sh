#!/bin/bash
# Both command lines side by side at full size (cfg2 top 20000, cfg4 top 100, the cfg5 matrix / penalty sweep), then the
# gap-penalty probe
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/full_parity.py 1.0 20000 1 cfg2 > $O/p_cfg2.txt 2>&1; echo "cfg2 exit $?"; tail -n 5 $O/p_cfg2.txt
timeout 600 python tools/full_parity.py 1.0 100 1 cfg4 > $O/p_cfg4.txt 2>&1; echo "cfg4 exit $?"; tail -n 5 $O/p_cfg4.txt
timeout 900 python tools/full_parity.py 1.0 100 1 cfg5 > $O/p_cfg5.txt 2>&1; echo "cfg5 exit $?"; tail -n 14 $O/p_cfg5.txt
timeout 600 python tools/penalty_probe.py > $O/p_penalty.txt 2>&1; echo "penalty exit $?"; tail -n 12 $O/p_penalty.txt
