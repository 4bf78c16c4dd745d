#!/bin/bash
# One GPU-box visit: pipe microbenchmark, parity tests, smoke, the default bench (both arms).  Bounded by timeouts.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > gpurun_out/gpu.txt 2>&1
nproc >> gpurun_out/gpu.txt
timeout 120 ./swimm_b200/pipebench > gpurun_out/pipebench.json 2> gpurun_out/pipebench.err; echo "pipebench exit $?"
timeout 1200 python -m pytest tests -m gpu -q -x > gpurun_out/pytest_gpu.log 2>&1; echo "pytest exit $?"
tail -n 15 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "smoke exit $?"; tail -n 3 gpurun_out/smoke.log
timeout 900 python bench.py > gpurun_out/bench.json 2> gpurun_out/bench.err; echo "bench exit $?"
tail -n 3 gpurun_out/bench.err; cat gpurun_out/bench.json
if [ -n "$WITH_REF" ]; then
  timeout 900 python bench.py --impl reference --steps 2 --warmup 1 > gpurun_out/bench_ref.json 2> gpurun_out/bench_ref.err; echo "ref exit $?"
  cat gpurun_out/bench_ref.json
fi
grep -E "mix_|prmt" gpurun_out/pipebench.json
