#!/bin/bash
# round 2, visit G (8 GPUs): swimm -m 3 -x 8 against the reference CLI on a cfg3-shaped database, and the bench at N = 8
cd "$(dirname "$0")/.."
O=gpurun_out/r2; mkdir -p $O
nvidia-smi --query-gpu=index,name --format=csv > $O/g_gpus.txt 2>&1; nproc >> $O/g_gpus.txt
timeout 900 python tools/cli_multi_gpu.py ${CLI_SCALE:-0.25} 8 10 > $O/g_cli_x8.txt 2>&1; echo "cli exit $?"; cat $O/g_cli_x8.txt
timeout 1500 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus 8 --steps 2 --warmup 3 > $O/g_bench_8gpu.json 2> $O/g_bench_8gpu.err; echo "bench exit $?"
tail -n 12 $O/g_bench_8gpu.err; cut -c1-1500 $O/g_bench_8gpu.json
