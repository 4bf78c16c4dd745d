#!/bin/bash
# ncu evidence for the bench command (B200_PROFILING.md recipe) with the query-pair kernel in the mix:
# launch list at scale 0.25, then full-scale full captures of four representative search launches.
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
if [ -z "$SKIP_LIST" ]; then
CMD="python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu-baseline --no-pipebench"
$CMD > gpurun_out/prof_plain.json 2> gpurun_out/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file gpurun_out/launches.csv $CMD > gpurun_out/ncu_list.log 2>&1
echo "launch list exit $?"
fi
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipebench"
$CMD > gpurun_out/prof3_plain.json 2> gpurun_out/prof3_plain.err || exit 1
i=0
for pat in "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.0, .bool.1" "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.1, .bool.1" "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.1, .bool.0" "wavefront_kernel<swg::Lane16, .int.8, "; do
  i=$((i+1))
  if [ -n "$ONLY" ] && [ "$ONLY" != "$i" ] && [ "$ONLY" != "23" -o $i -lt 2 -o $i -gt 3 ]; then continue; fi
  # the first launches of every instantiation are the empty warm-up launches of the first run: skip past them
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s 10 -c 1 -f -o gpurun_out/prof3_$i $CMD > gpurun_out/ncu3_$i.log 2>&1
  echo "capture $i exit $?"
done
ls -la gpurun_out | grep prof3
