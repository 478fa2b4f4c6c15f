#!/bin/bash
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/multi2_topo.txt 2>&1
timeout 600 python tools/check_multi.py > gpurun_out/check_multi_n2_single_process.txt 2>&1; tail -4 gpurun_out/check_multi_n2_single_process.txt
timeout 600 python -m torch.distributed.run --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py > gpurun_out/check_multi_n2_per_gpu.txt 2>&1; tail -4 gpurun_out/check_multi_n2_per_gpu.txt
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/bench_n2.log 2>&1; tail -c 1500 gpurun_out/bench_n2.log
