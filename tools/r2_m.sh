#!/bin/bash
# round 2, visit M: does the query-pair kernel hold its issue rate with 12 warps per SM (384-thread CTAs)?
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/calibrate_q2.py 0.35 32 24,28,32 > $O/m_calib_512.json 2> $O/m_calib_512.err; echo "512 exit $?"; cat $O/m_calib_512.json
SWIMM_B200_LIB=$PWD/swimm_b200/variant384/libswimm_cuda.so timeout 600 python tools/calibrate_q2.py 0.35 32 24,28,32 > $O/m_calib_384.json 2> $O/m_calib_384.err; echo "384 exit $?"; cat $O/m_calib_384.json
