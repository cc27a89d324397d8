#!/bin/bash
# One GPU-box visit that collects everything a round needs, so that box acquisition is paid once:
#
#   gpurun --timeout 1500 -- 'bash tools/gpu_round.sh r2'
#
# stages (each under its own timeout; a failing stage does not stop the next):
#   1. pytest -m gpu                                    -> gpurun_out/<tag>_pytest_gpu.log
#   2. smoke()                                          -> gpurun_out/<tag>_smoke.log
#   3. bench.py, default arguments (N=1)                -> gpurun_out/<tag>_bench.json
#   4. bench.py --workload both (joint step)            -> gpurun_out/<tag>_bench_joint.json
#   5. ncu launch list of one bench step                -> gpurun_out/<tag>_launches.csv  (+ per-kernel summary .txt)
#   6. ncu --set full of the dominant kernels           -> gpurun_out/<tag>_ncu_full.txt  (tools/tc_probe.py perf)
#   8. ncu --set full of one eager step's HBM-side kernels -> gpurun_out/<tag>_ncu_step.{ncu-rep,txt}
#   7. opt-in variants not yet measured (A/B)           -> gpurun_out/<tag>_pytest_gpu_experimental.log, _bench_experimental.txt
# Skip stages with SKIP="1 6" (space-separated numbers).  Numbers printed under ncu are never bench values.
tag=${1:-rX}
out=gpurun_out
mkdir -p $out
skip=" ${SKIP:-} "
run() { case "$skip" in *" $1 "*) echo "[gpu_round] stage $1 skipped"; return 1;; esac; echo "[gpu_round] stage $1: $2"; return 0; }

run 1 "pytest -m gpu" && { timeout 900 python -m pytest tests -m gpu -x -q > $out/${tag}_pytest_gpu.log 2>&1; tail -3 $out/${tag}_pytest_gpu.log; }
run 2 "smoke" && { timeout 300 python -c "import __graft_entry__ as g; g.smoke(); print('smoke ok')" > $out/${tag}_smoke.log 2>&1; tail -2 $out/${tag}_smoke.log; }
run 3 "bench N=1" && { timeout 600 python bench.py > $out/${tag}_bench.json 2> $out/${tag}_bench.err; tail -c 600 $out/${tag}_bench.json; }
run 4 "bench joint" && { timeout 600 python bench.py --workload both --no-cpu-baseline > $out/${tag}_bench_joint.json 2> $out/${tag}_bench_joint.err; tail -c 300 $out/${tag}_bench_joint.json; }
run 5 "ncu launch list" && {
  timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 3000 --csv --log-file $out/${tag}_launches.csv \
      python bench.py --steps 1 --warmup 3 --no-cpu-baseline > $out/${tag}_launches.log 2>&1
  python tools/launch_summary.py $out/${tag}_launches.csv > $out/${tag}_launches_step.txt 2>&1; head -12 $out/${tag}_launches_step.txt; }
run 6 "ncu --set full (dominant kernels)" && {
  timeout 900 ncu --set full --clock-control none --import-source on -k regex:tc_ -s 2 -c 10 -o $out/${tag}_ncu_full -f \
      python tools/tc_probe.py perf > $out/${tag}_ncu_full.log 2>&1
  ncu -i $out/${tag}_ncu_full.ncu-rep --page raw --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__cycles_elapsed.avg.per_second \
      > $out/${tag}_ncu_full.txt 2>&1; head -5 $out/${tag}_ncu_full.txt; }
# 8 (listed first so that SKIP numbering stays stable). `ncu --set full` of the HBM-side and row-box kernels of ONE eager
# training step (tools/op_times.py runs two warm-up steps, then one; 36 matching launches per step): the input for the
# next fusions
run 8 "ncu --set full, one eager step, HBM-side kernels" && {
  timeout 900 ncu --set full --clock-control none -k regex:'c1s2|bn_bwd|maxpool2_bwd|tc_conv_rb' --launch-skip 72 -c 36 \
      -o $out/${tag}_ncu_step -f python tools/op_times.py > $out/${tag}_ncu_step.log 2>&1
  ncu -i $out/${tag}_ncu_step.ncu-rep --page raw --csv \
      --metrics gpu__time_duration.sum,dram__bytes_read.sum,dram__bytes_write.sum,dram__throughput.avg.pct_of_peak_sustained_elapsed,sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active,sm__throughput.avg.pct_of_peak_sustained_elapsed,smsp__issue_active.avg.pct_of_peak_sustained_active,launch__grid_size,launch__registers_per_thread \
      > $out/${tag}_ncu_step.txt 2>&1; head -5 $out/${tag}_ncu_step.txt; }
# 7. the written-but-unmeasured variants: full GPU suite and the bench with them switched on
run 7 "experimental variants (EW_HOIST, WGRAD_STREAM, TC_SPLITK, C1_EPI2)" && {
  HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 HMGAN_C1_EPI2=1 HMGAN_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q \
      > $out/${tag}_pytest_gpu_experimental.log 2>&1; tail -3 $out/${tag}_pytest_gpu_experimental.log
  for v in "HMGAN_EW_HOIST=1" "HMGAN_WGRAD_STREAM=1" "HMGAN_TC_SPLITK=1" "HMGAN_C1_EPI2=1" "HMGAN_EW_HOIST=1 HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 HMGAN_C1_EPI2=1"; do
    echo "[gpu_round] bench with $v"
    env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
        python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
  done > $out/${tag}_bench_experimental.txt 2>&1; cat $out/${tag}_bench_experimental.txt; }
echo "[gpu_round] done"
