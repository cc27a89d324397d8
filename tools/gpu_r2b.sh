#!/bin/bash
# Round-2 GPU visit B: tc32 bring-up + per-variant test runs.  Outputs (small text files only) under gpurun_out/.
out=gpurun_out; mkdir -p $out
echo "[r2b] tc32 kernel probes"
timeout 600 python tools/tc_probe.py tc32 > $out/r2b_tc32_probe.log 2>&1; tail -5 $out/r2b_tc32_probe.log
echo "[r2b] tc32 step diagnostics"
timeout 900 python tools/dbg_tc32.py precision=tc32 wide64 gate64 dcgan512 joint512 B=2 > $out/r2b_tc32_steps.log 2>&1; tail -30 $out/r2b_tc32_steps.log
for v in HMGAN_EW_HOIST HMGAN_WGRAD_STREAM HMGAN_TC_SPLITK; do
  echo "[r2b] pytest -m gpu with $v=1"
  env $v=1 HMGAN_TEST_EXPERIMENTAL=1 timeout 900 python -m pytest tests -m gpu -q -rf --deselect tests/test_tc_gpu.py -k "not tc32" > $out/r2b_pytest_$v.log 2>&1
  tail -8 $out/r2b_pytest_$v.log
done
for v in "HMGAN_X=0" "HMGAN_WGRAD_STREAM=1" "HMGAN_TC_SPLITK=1" "HMGAN_WGRAD_STREAM=1 HMGAN_TC_SPLITK=1 HMGAN_EW_HOIST=1"; do
  echo "[r2b] joint bench with $v"
  env $v timeout 300 python bench.py --workload both --steps 20 --warmup 3 --no-cpu-baseline 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'])"
done > $out/r2b_bench_joint_variants.txt 2>&1; cat $out/r2b_bench_joint_variants.txt
echo "[r2b] done"
