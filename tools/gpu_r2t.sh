#!/bin/bash
# 2-GPU visit after the lane / host-path / new-kernel changes: NCCL data-parallel tests, then the 2-rank bench with its
# secondary workloads (joint step on its lanes with the captured all-reduces)
out=gpurun_out; mkdir -p $out
nvidia-smi -L
echo "[r2t] pytest tests/test_dp_gpu.py"
timeout 400 python -m pytest tests/test_dp_gpu.py -q -rf -x --timeout 240 > $out/r2t_pytest_dp.log 2>&1; tail -6 $out/r2t_pytest_dp.log | cut -c1-300
echo "[r2t] bench N=2"
timeout 400 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29517 \
    bench.py --gpus 2 --steps 20 --warmup 3 > $out/r2t_bench2.json 2> $out/r2t_bench2.err
echo "rc=$?"; python -c "
import json
d=json.loads([l for l in open('$out/r2t_bench2.json').read().strip().splitlines() if l.startswith('{')][-1])
print(d['n_gpus'], d['ms_per_step'], d['value'], d['e2e']['value'], d.get('replicas_identical'), d['clocks'], d.get('head_alive_frac'))
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['replicas_identical'])
"; tail -5 $out/r2t_bench2.err | cut -c1-300
echo "[r2t] reference arm under torchrun (rank 0 only)"
timeout 300 python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29519 \
    bench.py --impl reference --gpus 2 --steps 1 --warmup 0 --cpu-sample 1 2>/dev/null | tail -1 | cut -c1-400
echo "[r2t] done"
