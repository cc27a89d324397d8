#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2i] BN-from-a kernel test"
timeout 300 python -m pytest tests/test_kernels_gpu.py -q -k batchnorm -x 2>&1 | tail -4 | cut -c1-300
echo "[r2i] pytest -m gpu"
timeout 1200 python -m pytest tests -m gpu -q -rf --timeout 600 --deselect tests/test_dp_gpu.py > $out/r2i_pytest.log 2>&1; tail -8 $out/r2i_pytest.log | cut -c1-300
for v in "HMGAN_BN_FROM_A=0" "HMGAN_BN_FROM_A=1"; do
  echo "[r2i] bench $v"
  env $v timeout 300 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2>/dev/null | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'])"
done > $out/r2i_bench_variants.txt 2>&1; cat $out/r2i_bench_variants.txt
echo "[r2i] G-out forward under the tuning knobs"
for v in "HMGAN_X=0" "HMGAN_TC_SMAX=1" "HMGAN_TC_SMAX=2" "HMGAN_RB_ASLOTS=3" "HMGAN_TC_SMAX=2 HMGAN_RB_ASLOTS=4" "HMGAN_TC_SMAX=1 HMGAN_RB_ASLOTS=6" "HMGAN_TC_ROWBOX=0" "HMGAN_TC_SPLITK=0"; do
  echo "-- $v"; env $v timeout 120 python tools/tc_probe.py perf_up2 2>&1 | tail -3
done > $out/r2i_perf_up2.txt 2>&1; cat $out/r2i_perf_up2.txt
echo "[r2i] done"
