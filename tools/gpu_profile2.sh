#!/bin/bash
# Full-scale ncu captures of three representative search launches (one query each of 144, 1000 and 2005 residues).
cd "$(dirname "$0")/.."
mkdir -p gpurun_out
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipebench"
$CMD > gpurun_out/prof2_plain.json 2> gpurun_out/prof2_plain.err || exit 1
i=0
for pat in "Lane16, .int.8, .int.18" "Lane16, .int.32, .int.32, .bool.0" "Lane16, .int.32, .int.32, .bool.1"; do
  i=$((i+1))
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s 1 -c 1 -o gpurun_out/prof2_$i $CMD > gpurun_out/ncu2_$i.log 2>&1
  echo "capture $i exit $?"
done
ls -la gpurun_out | grep prof2
