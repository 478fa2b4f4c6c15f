#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
for wc in 0 1 0 1; do
DCTC_E2E_WC=$wc timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/av_wc${wc}.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/av_wc${wc}.log'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; print("wc=${wc} N=", d['n_gpus'], "e2e", round(e['value']), e.get('pcie'), e.get('frac_of_pcie_ceiling'))
PY
done
