#!/bin/bash
# visit U: validation of the tree as it stands -- smoke(), the whole GPU suite, the default bench line, refreshed per-op
# times / launch list / ncu capture of the final kernel set
out=gpurun_out; mkdir -p $out
echo "[r2u] smoke"
timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" 2>&1 | tail -3 | cut -c1-300
echo "[r2u] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2u_pytest.log 2>&1; tail -6 $out/r2u_pytest.log | cut -c1-300
echo "[r2u] bench (default command)"
timeout 600 python bench.py > $out/r2u_bench.json 2> $out/r2u_bench.err
python -c "
import json
d=json.loads(open('$out/r2u_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'], d['roofline']['frac'], d['roofline']['step_frac_of_burst'], d.get('head_alive_frac'), d.get('remeasured'))
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'], v['step_frac_of_burst'], v.get('head_alive_frac'))
print(d.get('cpu_baseline'))
"; tail -3 $out/r2u_bench.err
echo "[r2u] reference arm"
timeout 600 python bench.py --impl reference 2>/dev/null | tail -1 | cut -c1-600
echo "[r2u] per-op times"
timeout 200 python tools/op_times.py 32 dcgan > $out/r2u_op_times_dcgan.txt 2>&1; head -12 $out/r2u_op_times_dcgan.txt | cut -c1-160
timeout 200 python tools/op_times.py 16 p2p > $out/r2u_op_times_p2p.txt 2>&1; head -8 $out/r2u_op_times_p2p.txt | cut -c1-160
echo "[r2u] launch list of one step period"
timeout 400 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2u_launches_step.txt 2>&1; head -24 $out/r2u_launches_step.txt
echo "[r2u] ncu --set full of the final kernels (one eager step)"
timeout 600 ncu --set full --clock-control none -k regex:'tc_conv_rb|tc_wgrad_rb|c1s2|bn_bwd|maxpool2_bwd|col_reduce' --launch-skip 100 -c 50 \
    -o /tmp/r2u_step -f python tools/op_times.py > /tmp/ncu_step.log 2>&1
ncu -i /tmp/r2u_step.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,lts__t_sector_hit_rate.pct,launch__grid_size,launch__registers_per_thread,sm__cycles_elapsed.avg.per_second \
    > $out/r2u_ncu_final_kernels.csv 2>&1
python - <<'PY'
import csv
rows=list(csv.reader(open('gpurun_out/r2u_ncu_final_kernels.csv')))
h=rows[0]
ki=h.index('Kernel Name')
cols=[i for i,c in enumerate(h) if c in ('gpu__time_duration.sum','dram__bytes_read.sum','dram__bytes_write.sum','sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active','dram__throughput.avg.pct_of_peak_sustained_elapsed','smsp__issue_active.avg.pct_of_peak_sustained_active')]
print([h[i].split('.')[0][-22:] for i in cols], rows[1][cols[0]:cols[0]+1])
for r in rows[2:]:
    print(r[ki].split('(')[0][:30].ljust(30), [r[i] for i in cols])
PY
echo "[r2u] done"
