#!/bin/bash
# A/B of two builds of the library on the benchmark batch (kernel-only legs), after the parity tests of the new one.
# usage: tools/visit_ab.sh [pytest-selection]
cd "$(dirname "$0")/.."
O=gpurun_out/r2ab; mkdir -p $O
if [ -z "$SKIP_TESTS" ]; then timeout 1500 python -m pytest tests -m gpu -q --timeout 900 -x > $O/pytest.log 2>&1; echo "pytest exit $?"; fi
tail -n 6 $O/pytest.log
B="python bench.py --steps 3 --warmup 3 --no-cpu-baseline --no-pipebench --no-extra"
for lib in ${LIBS:-new base new base}; do
  if [ $lib = new ]; then unset SWIMM_B200_LIB; else export SWIMM_B200_LIB=$PWD/gpurun_in/libswimm_cuda_$lib.so; fi
  timeout 600 $B > $O/bench_$lib.json 2> $O/bench_$lib.err; echo "$lib exit $?"
  python - "$O/bench_$lib.json" <<'P'
import json,sys
d=json.loads(open(sys.argv[1]).read().strip().splitlines()[-1])
print(d["value"], d["ms_per_step"], d["roofline"]["frac"], d.get("e2e",{}).get("value"), d.get("clocks"))
P
done
