#!/bin/bash
# round 2, visit C: whole GPU suite, smoke, default bench (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q -x --timeout 900 > $O/c_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 12 $O/c_pytest_all.log
timeout 300 python __graft_entry__.py smoke > $O/c_smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 $O/c_smoke.log
timeout 1500 python bench.py > $O/c_bench.json 2> $O/c_bench.err; echo "bench exit $?"
tail -n 12 $O/c_bench.err; cat $O/c_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/c_bench_ref.json 2> $O/c_bench_ref.err; echo "ref exit $?"
cat $O/c_bench_ref.json
