#!/bin/bash
# round 2, visit F (2 GPUs): multi-GPU entry points, CLI -x 0, the bench at N = 2
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/f_gpus.txt 2>&1
timeout 900 python -m pytest tests/test_gpu_batches.py tests/test_gpu_cli_stream.py -m gpu -q --timeout 600 -x > $O/f_pytest.log 2>&1; echo "pytest exit $?"
tail -n 5 $O/f_pytest.log
timeout 600 python tools/cli_multi_gpu.py 0.05 0 10 > $O/f_cli_x2.txt 2>&1; echo "cli exit $?"; cat $O/f_cli_x2.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 2 --steps 3 --warmup 3 > $O/f_bench_2gpu.json 2> $O/f_bench_2gpu.err; echo "bench exit $?"
tail -n 8 $O/f_bench_2gpu.err; cut -c1-1500 $O/f_bench_2gpu.json
