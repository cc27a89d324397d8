#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2k] ncu source-level capture of the G-out forward launch"
timeout 300 ncu --set full --clock-control none --import-source on -k regex:tc_conv_rb -c 1 -o /tmp/gout -f python tools/tc_probe.py perf_up2 > /tmp/ncu_gout.log 2>&1
timeout 120 python tools/ncu_top_stalls.py /tmp/gout.ncu-rep 45 > $out/r2k_gout_stalls.txt 2>&1; head -110 $out/r2k_gout_stalls.txt | cut -c1-200
ncu -i /tmp/gout.ncu-rep --page raw --csv --metrics gpu__time_duration.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,smsp__issue_active.avg.pct_of_peak_sustained_active,dram__bytes_read.sum,l1tex__data_pipe_lsu_wavefronts.sum,smsp__cycles_active.avg 2>/dev/null | tail -2 | cut -c1-400
echo "[r2k] compute-sanitizer memcheck"
timeout 420 compute-sanitizer --tool memcheck --print-limit 20 python tools/sanitize_step.py > $out/r2k_memcheck.log 2>&1; tail -8 $out/r2k_memcheck.log | cut -c1-200
echo "[r2k] compute-sanitizer racecheck"
timeout 420 compute-sanitizer --tool racecheck --print-limit 20 python tools/sanitize_step.py > $out/r2k_racecheck.log 2>&1; tail -8 $out/r2k_racecheck.log | cut -c1-200
echo "[r2k] done"
