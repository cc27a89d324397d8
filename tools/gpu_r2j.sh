#!/bin/bash
out=gpurun_out; mkdir -p $out
echo "[r2j] kernel probes: pool, up2 (interleaved MMA order), perf"
timeout 200 python tools/tc_probe.py pool 2>&1 | tail -6 | cut -c1-200
timeout 200 python tools/tc_probe.py perf_up2 2>&1 | tail -3
timeout 200 python tools/tc_probe.py perf 2>&1 | tail -10
echo "[r2j] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2j_pytest.log 2>&1; tail -12 $out/r2j_pytest.log | cut -c1-300
for v in "HMGAN_POOL_TC=0 HMGAN_BN_FROM_A=0" "HMGAN_POOL_TC=0 HMGAN_BN_FROM_A=1" "HMGAN_POOL_TC=1 HMGAN_BN_FROM_A=1"; do
  echo "[r2j] bench $v"
  env $v timeout 200 python bench.py --steps 20 --warmup 3 --no-cpu-baseline --no-secondary 2> $out/r2j_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], d['losses'][:2], d['clocks'])" || tail -5 $out/r2j_bench.err
done > $out/r2j_bench_variants.txt 2>&1; cat $out/r2j_bench_variants.txt | cut -c1-300
echo "[r2j] done"
