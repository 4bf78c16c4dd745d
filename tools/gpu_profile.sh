#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe): launch list, then one full capture of the top kernel.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --scale ${PROF_SCALE:-0.1} --steps 2 --warmup 3 --no-cpu-baseline --no-pipebench"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 2000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
$CMD > gpurun_out/prof_plain2.json 2> gpurun_out/prof_plain2.err &&
ncu --set full --clock-control none --import-source on -k regex:wavefront_kernel -s ${PROF_SKIP:-200} -c ${PROF_COUNT:-4} -o gpurun_out/prof $CMD > gpurun_out/ncu_full.log 2>&1
echo "full capture exit $?"
ls -la gpurun_out | tail -12
