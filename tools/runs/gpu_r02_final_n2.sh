#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/final_bench_n2.log 2>&1; echo "rc $?"; tail -c 400 gpurun_out/final_bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --impl reference --gpus 2 --steps 2 --warmup 1 > gpurun_out/final_bench_ref_n2.log 2>&1; echo "rc ref $?"; tail -c 300 gpurun_out/final_bench_ref_n2.log
