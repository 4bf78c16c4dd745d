#!/bin/bash
# round 2, visit K: extend-immediate probes, shard floor, all tests, bench
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 120 ./swimm_b200/pipebench > $O/k_pipebench.json 2> $O/k_pipebench.err; echo "pipebench exit $?"; grep -E "mix_" $O/k_pipebench.json
timeout 300 python tools/shard_probe.py 2>&1 | grep -v "^\[swg\]   G=" > $O/k_shard.txt; cat $O/k_shard.txt
timeout 1800 python -m pytest tests -m gpu -q --timeout 900 -x > $O/k_pytest_all.log 2>&1; echo "pytest all exit $?"
tail -n 5 $O/k_pytest_all.log
timeout 1500 python bench.py > $O/k_bench.json 2> $O/k_bench.err; echo "bench exit $?"
tail -n 8 $O/k_bench.err; cut -c1-200 $O/k_bench.json
