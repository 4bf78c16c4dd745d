#!/bin/bash
# round 2, visit I: late task fetch -- parity, trace probe, headline bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/trace_probe.py 2>&1 | grep -v "G=32 K=" | grep -E "====|search|sequence-pair kernel|long tiles: 32-thread|long-sequence" > $O/i_trace.txt; cat $O/i_trace.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/i_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 6 $O/i_pytest_all.log
timeout 900 python bench.py --no-cpu-baseline > $O/i_bench.json 2> $O/i_bench.err; echo "bench exit $?"
tail -n 8 $O/i_bench.err; cut -c1-400 $O/i_bench.json
