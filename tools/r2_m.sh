#!/bin/bash
# round 2, visit M: does the query-pair kernel hold its issue rate with 12 warps per SM (384-thread CTAs)?
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
timeout 600 python tools/calibrate_q2.py 0.35 32 24,28,32 > $O/m_calib_512.json 2> $O/m_calib_512.err; echo "512 exit $?"; cat $O/m_calib_512.json
SWIMM_B200_LIB=$PWD/swimm_b200/variant384/libswimm_cuda.so timeout 600 python tools/calibrate_q2.py 0.35 32 24,28,32 > $O/m_calib_384.json 2> $O/m_calib_384.err; echo "384 exit $?"; cat $O/m_calib_384.json
timeout 900 python -m pytest tests/test_gpu_batches.py -m gpu -q -x --timeout 600 > $O/m_pytest_batches.log 2>&1; echo "batches exit $?"; tail -n 4 $O/m_pytest_batches.log
timeout 600 python tools/full_parity.py 1.0 100 1 cfg4 > $O/m_parity_cfg4.txt 2>&1; echo "cfg4 exit $?"; tail -n 3 $O/m_parity_cfg4.txt
