#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2n] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2n_pytest.log 2>&1; tail -8 $out/r2n_pytest.log | cut -c1-300
echo "[r2n] bench"
timeout 500 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/r2n_bench.json 2> $out/r2n_bench.err
python -c "
import json
d=json.loads(open('$out/r2n_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['step_frac_of_burst'])
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['step_frac_of_burst'])
"; tail -3 $out/r2n_bench.err
echo "[r2n] launch list of one step period"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2n_launches_step.txt 2>&1; head -40 $out/r2n_launches_step.txt
echo "[r2n] per-op times of an eager pix2pix step"
timeout 300 python tools/op_times.py 16 p2p > $out/r2n_op_times_p2p.txt 2>&1; head -45 $out/r2n_op_times_p2p.txt
echo "[r2n] done"
