#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2o] pack tests"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -x --timeout 200 -k "pack" 2>&1 | tail -3 | cut -c1-300
echo "[r2o] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2o_pytest.log 2>&1; tail -8 $out/r2o_pytest.log | cut -c1-300
echo "[r2o] bench"
timeout 500 python bench.py --steps 20 --warmup 3 > $out/r2o_bench.json 2> $out/r2o_bench.err
python -c "
import json
d=json.loads(open('$out/r2o_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['step_frac_of_burst'])
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['step_frac_of_burst'])
"; tail -3 $out/r2o_bench.err
echo "[r2o] launch list of one step period"
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2o_launches_step.txt 2>&1; head -44 $out/r2o_launches_step.txt
echo "[r2o] ncu --set full of the final kernels (one eager step)"
timeout 600 ncu --set full --clock-control none -k regex:'tc_conv_rb|tc_wgrad_rb|c1s2|bn_bwd|maxpool2_bwd|col_reduce|pack_tile' --launch-skip 120 -c 60 \
    -o /tmp/r2o_step -f python tools/op_times.py > /tmp/ncu_step.log 2>&1
ncu -i /tmp/r2o_step.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second \
    > $out/r2o_ncu_final_kernels.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2o_ncu_final_kernels.csv')))
h=rows[0]
ki=h.index('Kernel Name')
cols=[i for i,c in enumerate(h) if c in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active')]
print([h[i].split('.')[0][-22:] for i in cols], rows[1][cols[0]:cols[0]+1])
for r in rows[2:]:
    print(r[ki].split('(')[0][:30].ljust(30), [r[i] for i in cols])
PY
echo "[r2o] done"
