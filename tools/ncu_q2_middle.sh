#!/bin/bash
# one full ncu capture of a middle-pass launch of the query-pair kernel on the benchmark batch (after a plain run)
cd "$(dirname "$0")/.."
O=gpurun_out/r2q; mkdir -p $O
CMD="python bench.py --steps 1 --warmup 1 --no-cpu-baseline --no-pipebench --no-extra"
$CMD > $O/plain.json 2> $O/plain.err || exit 1
ncu --set full --clock-control none --import-source on --kernel-name-base demangled -k regex:"wavefront_q2_kernel<.int.32, .int.[0-9]+, .bool.1, .bool.1" -s 10 -c 1 -f -o $O/q2_middle $CMD > $O/ncu.log 2>&1
echo "capture exit $?"; ls -la $O
