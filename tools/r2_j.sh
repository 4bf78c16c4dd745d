#!/bin/bash
# round 2, visit J: race-free long-tile scheduling -- probes, all GPU tests, full bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 500 python tools/slow_probe.py 2>&1 | tail -7 > $O/j_slow.txt; cat $O/j_slow.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/j_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 6 $O/j_pytest_all.log
timeout 300 python __graft_entry__.py smoke > $O/j_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/j_smoke.log
timeout 1500 python bench.py > $O/j_bench.json 2> $O/j_bench.err; echo "bench exit $?"
tail -n 12 $O/j_bench.err; cut -c1-300 $O/j_bench.json
