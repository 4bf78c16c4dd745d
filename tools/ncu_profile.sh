#!/bin/bash
# round 2 ncu evidence (B200_PROFILING.md recipe): launch list of the bench command at scale 0.25, then full-scale full
# captures of representative launches: query-pair kernel first / middle / last pass, sequence-pair kernel (q 144), and the
# long-sequence kernel on cfg4.  Every command first runs to completion WITHOUT ncu.
cd "$(dirname "$0")/.."
O=gpurun_out/r2p; mkdir -p $O
CMD="python bench.py --scale 0.25 --steps 2 --warmup 3 --no-cpu-baseline --no-pipebench --no-extra"
$CMD > $O/prof_plain.json 2> $O/prof_plain.err &&
ncu --metrics gpu__time_duration.sum --clock-control none -c 6000 --csv --log-file $O/launches.csv $CMD > $O/ncu_list.log 2>&1
echo "launch list exit $?"
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipebench --no-extra"
$CMD > $O/prof3_plain.json 2> $O/prof3_plain.err || exit 1
i=0
for pat in "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.0, .bool.1" "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.1, .bool.1" "wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.1, .bool.0" "wavefront_kernel<swg::Lane16, .int.8, "; do
  i=$((i+1)); SKIP=10; if [ $i = 4 ]; then SKIP=3; fi
  if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " $i "; then continue; fi
  # the first launches of every instantiation are the empty warm-up launches of the first run: skip past them
  ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"$pat" -s $SKIP -c 1 -f -o $O/prof3_$i $CMD > $O/ncu3_$i.log 2>&1
  echo "capture $i exit $?"
done
if [ -n "$ONLY" ] && ! echo " $ONLY " | grep -q " 5 "; then ls -la $O | grep prof3; exit 0; fi
CMD="python tools/xw_profile_target.py"
$CMD > $O/xw_plain.txt 2>&1 || exit 1
tail -n 3 $O/xw_plain.txt
# the long-sequence kernel's launch for the 5478-row query (16 warps x 11 rows): launch 0 of that instantiation is the
# empty warm-up launch
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wavefront_xw_kernel<swg::Lane16, .int.11," -s 1 -c 1 -f -o $O/prof3_5 $CMD > $O/ncu3_5.log 2>&1
echo "capture 5 exit $?"
ls -la $O | grep prof3
