#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2f] c1s2 kernel tests first"
timeout 600 python -m pytest tests/test_tc_gpu.py -q -k "c1s2" -x > $out/r2f_pytest_c1.log 2>&1; tail -4 $out/r2f_pytest_c1.log | cut -c1-300
echo "[r2f] pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q -rf > $out/r2f_pytest.log 2>&1; tail -10 $out/r2f_pytest.log | cut -c1-300
echo "[r2f] bench"
timeout 600 python bench.py --steps 20 --warmup 3 --no-cpu-baseline > $out/r2f_bench.json 2> $out/r2f_bench.err
python -c "
import json
d=json.loads(open('$out/r2f_bench.json').read().strip().splitlines()[-1])
print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'], d['gpu_launches'])
for k,v in d.get('secondary',{}).items(): print(k, v['ms_per_step'], v['value'], v['e2e']['value'])
"; tail -3 $out/r2f_bench.err
echo "[r2f] kernel perf probes"
timeout 300 python tools/tc_probe.py c1 2>&1 | tail -4
timeout 300 python tools/tc_probe.py c1bwd 2>&1 | tail -4
echo "[r2f] launch list of one step (eager, serialised)"
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 4000 --csv --log-file /tmp/launches.csv \
    python bench.py --steps 1 --warmup 3 --no-cpu-baseline --no-secondary > /tmp/launches.log 2>&1
python tools/launch_summary.py /tmp/launches.csv 0.2 > $out/r2f_launches_step.txt 2>&1; head -45 $out/r2f_launches_step.txt
echo "[r2f] ncu --set full on the G-out forward launch and the c1s2 kernels"
timeout 900 ncu --set full --clock-control none --import-source on -k regex:'c1s2_conv|c1s2_bwd_kernel|tc_conv_rb' --launch-skip 36 -c 14 \
    -o /tmp/r2f_step -f python tools/op_times.py > /tmp/ncu_step.log 2>&1
ncu -i /tmp/r2f_step.ncu-rep --page raw --csv \
    --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,l1tex__data_pipe_lsu_wavefronts.sum,l1tex__t_requests_pipe_lsu_mem_global_op_st.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_st.sum,l1tex__t_requests_pipe_lsu_mem_global_op_ld.sum,l1tex__t_sectors_pipe_lsu_mem_global_op_ld.sum,launch__grid_size,sm__cycles_elapsed.avg.per_second \
    > $out/r2f_ncu_step.txt 2>&1
cut -c1-400 $out/r2f_ncu_step.txt | head -20
ncu -i /tmp/r2f_step.ncu-rep --page details --csv 2>/dev/null | grep -i "stall\|Kernel Name\|Warp Cycles Per Issued\|No Eligible\|Issue Slots Busy" | cut -c1-300 | head -120 > $out/r2f_ncu_details.txt
wc -l $out/r2f_ncu_details.txt
echo "[r2f] done"
