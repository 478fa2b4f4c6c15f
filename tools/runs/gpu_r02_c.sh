#!/bin/bash
# 2-GPU call: C multi-GPU layer on real peers + bench at N=2
set -x
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/c_topo.txt 2>&1
timeout 600 python tools/check_multi.py > gpurun_out/c_check_single.log 2>&1; echo "rc $?" >> gpurun_out/c_check_single.log; tail -8 gpurun_out/c_check_single.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29511 tools/check_multi.py > gpurun_out/c_check_ranks.log 2>&1; echo "rc $?" >> gpurun_out/c_check_ranks.log; tail -8 gpurun_out/c_check_ranks.log
timeout 600 python -m pytest tests/test_multi_gpu.py -q > gpurun_out/c_pytest_multi.log 2>&1; tail -3 gpurun_out/c_pytest_multi.log
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus 2 --steps 20 --warmup 5 > gpurun_out/c_bench_n2.log 2>&1; echo "rc $?" >> gpurun_out/c_bench_n2.log; tail -c 2500 gpurun_out/c_bench_n2.log
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513 bench.py --gpus 2 --steps 10 --warmup 3 --workload gigapixel > gpurun_out/c_bench_giga_n2.log 2>&1; echo "rc $?" >> gpurun_out/c_bench_giga_n2.log; tail -c 1200 gpurun_out/c_bench_giga_n2.log
