#!/bin/bash
# One visit with everything the round's last commit must pass: GPU tests, smoke, bench (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/z_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 5 $O/z_pytest_all.log
timeout 300 python __graft_entry__.py smoke > $O/z_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/z_smoke.log
timeout 1500 python bench.py > $O/z_bench.json 2> $O/z_bench.err; echo "bench exit $?"
tail -n 8 $O/z_bench.err; cut -c1-200 $O/z_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/z_bench_ref.json 2> $O/z_bench_ref.err; echo "ref exit $?"; cut -c1-200 $O/z_bench_ref.json
