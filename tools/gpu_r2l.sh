#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2l] thin d2s epilogue: probes"
timeout 200 python tools/tc_probe.py perf_up2 2>&1 | tail -3
timeout 300 python -m pytest tests/test_tc_gpu.py -q -x --timeout 200 -k "up2 or deconv or thin or stride2" 2>&1 | tail -3 | cut -c1-300
echo "[r2l] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2l_pytest.log 2>&1; tail -8 $out/r2l_pytest.log | cut -c1-300
echo "[r2l] bench (full default line)"
timeout 500 python bench.py --steps 20 --warmup 3 > $out/r2l_bench.json 2> $out/r2l_bench.err
python -c "
import json
d=json.loads(open('$out/r2l_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['step_frac_of_burst'])
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['step_frac_of_burst'])
print(d['cpu_baseline'])
"; tail -3 $out/r2l_bench.err
echo "[r2l] done"
