#!/bin/bash
# visit Z: generator packs deferred to the start of the next step (HMGAN_DEFER_GPACK), same-box A/B + the graph / staleness tests
out=gpurun_out; mkdir -p $out
echo "[r2z] staleness / graph tests"
timeout 600 python -m pytest tests/test_step_gpu.py tests/test_entry_gpu.py -q -x --timeout 300 -k "replayed or async or host_path or multi_stream or entry or experiment" 2>&1 | tail -5 | cut -c1-300
for v in "HMGAN_DEFER_GPACK=0 --workload dcgan" "HMGAN_DEFER_GPACK=1 --workload dcgan" "HMGAN_DEFER_GPACK=0 --workload dcgan" "HMGAN_DEFER_GPACK=1 --workload dcgan" "HMGAN_DEFER_GPACK=0 --workload both" "HMGAN_DEFER_GPACK=1 --workload both"; do
  set -- $v
  envs=""; args=""
  for w in "$@"; do case $w in HMGAN_*) envs="$envs $w";; *) args="$args $w";; esac; done
  echo "[r2z] bench $v"
  env $envs timeout 200 python bench.py --steps 30 --warmup 3 --no-cpu-baseline --no-secondary $args 2> $out/r2z_bench.err | tail -1 | \
      python -c "import json,sys; d=json.loads(sys.stdin.read()); print(d['ms_per_step'], d['value'], d['e2e']['value'], [round(v,4) for v in d['losses']], d['clocks']['sm_mhz'], d.get('head_alive_frac'))" || tail -5 $out/r2z_bench.err
done > $out/r2z_bench_variants.txt 2>&1; cat $out/r2z_bench_variants.txt | cut -c1-300
echo "[r2z] pytest -m gpu"
timeout 900 python -m pytest tests -m gpu -q -rf --timeout 300 --deselect tests/test_dp_gpu.py > $out/r2z_pytest.log 2>&1; tail -6 $out/r2z_pytest.log | cut -c1-300
echo "[r2z] done"
