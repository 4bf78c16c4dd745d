#!/bin/bash
# Query-pair kernel bring-up: its parity tests first (bounded), then the whole GPU suite and a bench.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_gpu_query_pairs.py -m gpu -q -x > gpurun_out/pytest_q2.log 2>&1; echo "pytest q2 exit $?"
tail -n 25 gpurun_out/pytest_q2.log
timeout 900 python -m pytest tests -m gpu -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest all exit $?"
tail -n 8 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --no-cpu-baseline > gpurun_out/bench_q2.json 2> gpurun_out/bench_q2.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench_q2.err; cat gpurun_out/bench_q2.json
