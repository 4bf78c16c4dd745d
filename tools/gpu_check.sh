#!/bin/bash
# First-contact GPU run: pipe microbenchmark, parity tests, smoke, a short bench.  Everything is bounded by timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt; lscpu | grep -E "Model name|^CPU\(s\)|Thread|Core|Socket" >> gpurun_out/gpu.txt
timeout 120 ./swimm_b200/pipebench > gpurun_out/pipebench.json 2> gpurun_out/pipebench.err; echo "pipebench exit $?"
timeout 900 python -m pytest tests -m gpu -q --maxfail=30 -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 40 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout 600 python bench.py --scale ${BENCH_SCALE:-0.1} --steps 2 --warmup 1 > gpurun_out/bench_small.json 2> gpurun_out/bench_small.err; echo "bench exit $?"
tail -n 5 gpurun_out/bench_small.err; cat gpurun_out/bench_small.json
cat gpurun_out/pipebench.json
