#!/bin/bash
# N-GPU sanity run of the bench line (the driver's scaling run): usage  gpurun --gpus N -- bash tools/gpu_r2x.sh N
N=${1:-8}
out=gpurun_out; mkdir -p $out
nvidia-smi -L | head -8
echo "[r2x] bench N=$N"
timeout 500 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29531 \
    bench.py --gpus $N --steps 20 --warmup 3 > $out/r2x_bench$N.json 2> $out/r2x_bench$N.err
echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('$out/r2x_bench$N.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['clocks'], d.get('head_alive_frac'), d.get('remeasured'))
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['replicas_identical'])
"; tail -4 $out/r2x_bench$N.err | cut -c1-300
echo "[r2x] done"
