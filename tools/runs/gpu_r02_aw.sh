#!/bin/bash
mkdir -p gpurun_out
N=${1:-8}
for cfg in "33554432 48" "268435456 6" "1073741824 2"; do
set -- $cfg
DCTC_PROBE_BYTES=$1 DCTC_PROBE_ITERS=$2 timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 bench.py --gpus $N --steps 20 --warmup 5 --no-configs --no-cpu-baseline > gpurun_out/aw.log 2>&1
python - <<PY
import json
for l in open('gpurun_out/aw.log'):
    if l.startswith('{'):
        d=json.loads(l); e=d['e2e']; p=e.get('pcie'); print("probe $1 x $2: e2e", round(e['value']), "h2d %.1f d2h %.1f bidir/dir %.1f" % (p['h2d_gbs'], p['d2h_gbs'], p['bidir_gbs_per_dir']), e.get('frac_of_pcie_ceiling'))
PY
done
