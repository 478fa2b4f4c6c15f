#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/multi8_topo.txt 2>&1
timeout 600 python tools/check_multi.py > gpurun_out/check_multi_n8_single_process.txt 2>&1; tail -3 gpurun_out/check_multi_n8_single_process.txt
timeout 600 python -m torch.distributed.run --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py > gpurun_out/check_multi_n8_per_gpu.txt 2>&1; tail -3 gpurun_out/check_multi_n8_per_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 8 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 8 --steps 20 --warmup 5 > gpurun_out/bench_n8.log 2>&1; tail -c 600 gpurun_out/bench_n8.log
