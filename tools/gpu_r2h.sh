#!/bin/bash
# 2-GPU visit: NCCL data-parallel tests + the 2-rank bench (captured async all-reduces, clean teardown)
out=gpurun_out; mkdir -p $out
nvidia-smi -L
echo "[r2h] pytest tests/test_dp_gpu.py"
timeout 900 python -m pytest tests/test_dp_gpu.py -q -rf -x > $out/r2h_pytest_dp.log 2>&1; tail -15 $out/r2h_pytest_dp.log | cut -c1-300
echo "[r2h] bench N=2"
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2h_bench2.json 2> $out/r2h_bench2.err
echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('$out/r2h_bench2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['clocks'])
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['replicas_identical'])
"; tail -5 $out/r2h_bench2.err | cut -c1-300
echo "[r2h] bench N=1 on the same box"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
echo "[r2h] done"
