#!/bin/bash
# visit V: the library without the CTA-pair kernel, the rewritten host-path test, the default bench line
out=gpurun_out; mkdir -p $out
echo "[r2v] smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300
echo "[r2v] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2v_pytest.log 2>&1; tail -6 $out/r2v_pytest.log | cut -c1-300
echo "[r2v] bench (default command)"
timeout 600 python bench.py > $out/r2v_bench.json 2> $out/r2v_bench.err
python -c "
import json
d=json.loads(open('$out/r2v_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['step_frac_of_burst'], d.get('head_alive_frac'), d.get('remeasured'))
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['step_frac_of_burst'], v.get('head_alive_frac'))
print(d.get('cpu_baseline'))
"; tail -3 $out/r2v_bench.err
echo "[r2v] done"
