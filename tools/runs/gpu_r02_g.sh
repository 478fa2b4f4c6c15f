#!/bin/bash
# 8-GPU call: C multi-GPU layer on 8 peers, bench at N=8 (weak 4K + C4 + C5 strong + e2e ceiling)
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/g_topo.txt 2>&1
timeout 600 python tools/check_multi.py > gpurun_out/g_check_single.log 2>&1; echo "rc $?" >> gpurun_out/g_check_single.log; tail -8 gpurun_out/g_check_single.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py > gpurun_out/g_check_ranks.log 2>&1; echo "rc $?" >> gpurun_out/g_check_ranks.log; tail -7 gpurun_out/g_check_ranks.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/g_bench_n8.log 2>&1; echo "rc $?" >> gpurun_out/g_bench_n8.log; tail -c 3000 gpurun_out/g_bench_n8.log
