#!/bin/bash
# round 2, visit N: gap-extend immediates in the sequence-pair kernel -- probe, all tests, smoke, bench (both arms)
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/penalty_probe.py > $O/n_penalties.txt 2>&1; echo "probe exit $?"; cat $O/n_penalties.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/n_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 5 $O/n_pytest_all.log
timeout 300 python __graft_entry__.py smoke > $O/n_smoke.log 2>&1; echo "smoke exit $?"; tail -n 2 $O/n_smoke.log
timeout 1500 python bench.py > $O/n_bench.json 2> $O/n_bench.err; echo "bench exit $?"
tail -n 8 $O/n_bench.err; cut -c1-200 $O/n_bench.json
timeout 600 python bench.py --impl reference --steps 2 --warmup 1 > $O/n_bench_ref.json 2> $O/n_bench_ref.err; echo "ref exit $?"; cut -c1-300 $O/n_bench_ref.json
